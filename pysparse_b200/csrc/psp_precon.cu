// psp_precon.cu -- Jacobi and SSOR / symmetric Gauss-Seidel preconditioners.
//
// Replaces jacobi() (pysparse/precon/src/preconmodule.c:35-54), ssor_kernel
// (:95-143) and symgs_kernel (:149-193) plus their constructors (:352-459).
//
// SSOR keeps the reference's lexicographic sweep order.  The analysis phase
// (psp_ssor_create) computes, on the host, the level sets of the dependency
// graph of L (forward sweep: row i needs x[j] for every stored L[i,j], j < i)
// and of L^T (backward sweep: row i needs x[i'] for every stored L[i',i],
// i' > i), and builds LEVEL-MAJOR copies of the operands: the rows of a level
// are stored back to back (coalesced streaming), forward rows list their
// entries in ascending column order, backward rows in DESCENDING original-row
// order -- the order in which the reference's backward loop scatters into
// y[j] -- so the scatter becomes an atomic-free gather with the same
// accumulation order and the results are bit-identical to the reference.
//
// Rows inside a level are independent; levels are separated by a barrier.
// Two engines, one launch per preconditioner application either way:
//   * narrow levels (lexicographic 2-D grids: 2n-1 levels of <= n rows): ONE
//     thread-block cluster; the per-level barrier is the hardware cluster
//     barrier (barrier.cluster arrive/wait, ~0.2 us) and the next level's row
//     data is prefetched between arrive and wait;
//   * wide levels: persistent cooperative grid with grid.sync() per level.
#include <cooperative_groups.h>
#include "psp_mat.cuh"
#include "psp_vec.cuh"
#include "psp_correq.cuh"
#include "psp_sweep_p2p.cuh"

namespace cg = cooperative_groups;

namespace psp {

struct FJacobiCorr {        // y = (x - y) .* dinv + temp      (preconmodule.c:50-51)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *x, *dinv, *temp; double *y;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *, double &) { y[i] = (x[i] - y[i]) * dinv[i] + temp[i]; }
};
struct FJacobiScale {       // y = x .* dinv                   (preconmodule.c:41-42)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *x, *dinv; double *y;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *, double &) { y[i] = x[i] * dinv[i]; }
};

struct SweepArgs {
  int n;
  const double *diag;
  const int *rows_f, *lptr_f, *find, *fcol; const double *fval; int nlev_f;
  const int *rows_b, *lptr_b, *bind, *bcol; const double *bval; int nlev_b;
  const double *b;              // right-hand side (the vector being preconditioned)
  double *x;                    // result
  double *w;                    // work: L x / L^T x   (reference: y resp. h)
  double omega;
  int steps;
  const State *S;               // optional gate
};

// One row of a sweep.  RELAX = false: symgs_kernel (preconmodule.c:149-193);
// RELAX = true: ssor_kernel (:95-143).  fwd selects the half-step.  The
// expression order is the reference's, every operation rounded separately.
template <bool RELAX>
__device__ __forceinline__ void sweep_row(const SweepArgs &a, bool fwd, int step, int r, int k0, int k1,
                                          const int *col, const double *val) {
  const double d = a.diag[r], br = a.b[r];
  if (!RELAX) {
    double s = 0.0;
    for (int k = k0; k < k1; k++) s = __dadd_rn(s, __dmul_rn(val[k], __ldcg(a.x + col[k])));
    // forward: x_i = (b_i - w_i - s)/d_i, w_i = s  (w = 0 in the first step)
    // backward: x_i = (b_i - (Lx)_i - t)/d_i, w_i = t  with (Lx)_i left in w by the forward sweep
    const double wr = (fwd && step == 0) ? 0.0 : a.w[r];
    a.x[r] = __ddiv_rn(__dsub_rn(__dsub_rn(br, wr), s), d);
    a.w[r] = s;
  } else {
    const double om = a.omega, om1 = 1.0 - a.omega;
    double temp;
    if (fwd && step == 0) temp = __dmul_rn(om, br);
    else temp = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(om1, a.x[r]), d), a.w[r]), __dmul_rn(om, br));
    double s = 0.0;
    for (int k = k0; k < k1; k++) s = __dsub_rn(s, __dmul_rn(val[k], __ldcg(a.x + col[k])));
    const double h = __dmul_rn(om, s);
    a.w[r] = h;
    a.x[r] = __ddiv_rn(__dadd_rn(temp, h), d);
  }
}

// ---------------------------------------------------------------- grid-wide engine
template <bool RELAX>
__global__ void __launch_bounds__(kThreads) k_sweep_grid(SweepArgs a) {
  if (a.S && *(const volatile int *)&a.S->done) return;
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gsz = gridDim.x * blockDim.x;
  for (int step = 0; step < a.steps; step++) {
    for (int half = 0; half < 2; half++) {
      const bool fwd = (half == 0);
      const int nlev = fwd ? a.nlev_f : a.nlev_b;
      const int *lptr = fwd ? a.lptr_f : a.lptr_b, *rows = fwd ? a.rows_f : a.rows_b;
      const int *ind = fwd ? a.find : a.bind, *col = fwd ? a.fcol : a.bcol;
      const double *val = fwd ? a.fval : a.bval;
      for (int l = 0; l < nlev; l++) {
        const int p0 = lptr[l], p1 = lptr[l + 1];
        for (int p = p0 + gtid; p < p1; p += gsz) sweep_row<RELAX>(a, fwd, step, rows[p], ind[p], ind[p + 1], col, val);
        grid.sync();
      }
    }
  }
}

// ---------------------------------------------------------------- cluster engine
constexpr int kSweepThreads = 512;

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Per-thread software pipeline in shared memory (cp.async / LDGSTS): every thread owns
// private ring slots, so no CTA-level synchronisation is involved.  While level l is
// being computed, the row METADATA (row index, entry range) of level l+4 and the row
// DATA (diagonal, right-hand side, work value, first kPre entries) of level l+2 are in
// flight: after the cluster barrier only the x gathers (L2 hits: written one level ago)
// and ~10 dependent flops remain on the critical path.
constexpr int kPre = 4;         // entries per row prefetched into shared memory
constexpr int kMetaRing = 5;
constexpr int kDataRing = 3;
constexpr int kLptrChunk = 8192;

__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr size_t kSweepSmem = sizeof(int) * (kLptrChunk + 8 + 3 * kMetaRing * kSweepThreads + kDataRing * kPre * kSweepThreads) +
                              sizeof(double) * (4 * kDataRing * kSweepThreads + kDataRing * kPre * kSweepThreads);

template <bool RELAX>
__global__ void __launch_bounds__(kSweepThreads, 1) k_sweep_cluster(SweepArgs a) {
  if (a.S && *(const volatile int *)&a.S->done) return;
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  constexpr int T = kSweepThreads;
  double *s_d = reinterpret_cast<double *>(sweep_smem);          // [kDataRing][T]
  double *s_b = s_d + kDataRing * T;
  double *s_w = s_b + kDataRing * T;
  double *s_xo = s_w + kDataRing * T;
  double *s_v = s_xo + kDataRing * T;                             // [kDataRing][kPre][T]
  int *s_c = reinterpret_cast<int *>(s_v + kDataRing * kPre * T); // [kDataRing][kPre][T]
  int *s_r = s_c + kDataRing * kPre * T;                          // [kMetaRing][T]
  int *s_k0 = s_r + kMetaRing * T;
  int *s_k1 = s_k0 + kMetaRing * T;
  int *s_lptr = s_k1 + kMetaRing * T;                             // [kLptrChunk + 8]

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int ctid = (int)cluster.block_rank() * T + tid;
  const int csz = (int)cluster.num_blocks() * T;

  for (int step = 0; step < a.steps; step++) {
    for (int half = 0; half < 2; half++) {
      const bool fwd = (half == 0);
      const int nlev = fwd ? a.nlev_f : a.nlev_b;
      const int *lptr = fwd ? a.lptr_f : a.lptr_b, *rows = fwd ? a.rows_f : a.rows_b;
      const int *ind = fwd ? a.find : a.bind, *col = fwd ? a.fcol : a.bcol;
      const double *val = fwd ? a.fval : a.bval;
      const bool need_w = RELAX ? !(fwd && step == 0) : !(fwd && step == 0);
      const bool need_xo = RELAX && !(fwd && step == 0);

      for (int base = 0; base < nlev; base += kLptrChunk) {
        const int nl = (nlev - base < kLptrChunk) ? nlev - base : kLptrChunk;   // levels in this chunk
        __syncthreads();
        for (int i = tid; i <= nl; i += T) s_lptr[i] = __ldg(lptr + base + i);
        __syncthreads();

        auto issue_meta = [&](int L) {        // L relative to the chunk
          if (L < nl) {
            const int p = s_lptr[L] + ctid, slot = (L % kMetaRing) * T + tid;
            if (p < s_lptr[L + 1]) {
              cp_async4(&s_r[slot], rows + p);
              cp_async4(&s_k0[slot], ind + p);
              cp_async4(&s_k1[slot], ind + p + 1);
            } else {
              s_r[slot] = -1;
            }
          }
          cp_commit();
        };
        auto issue_data = [&](int L) {        // the metadata of level L has landed
          if (L >= 0 && L < nl) {
            const int ms = (L % kMetaRing) * T + tid, ds = (L % kDataRing) * T + tid;
            const int r = s_r[ms];
            if (r >= 0) {
              const int k0 = s_k0[ms], k1 = s_k1[ms];
              cp_async8(&s_d[ds], a.diag + r);
              cp_async8(&s_b[ds], a.b + r);
              if (need_w) cp_async8(&s_w[ds], a.w + r);
              if (need_xo) cp_async8(&s_xo[ds], a.x + r);
#pragma unroll
              for (int j = 0; j < kPre; j++)
                if (k0 + j < k1) {
                  const int es = ((L % kDataRing) * kPre + j) * T + tid;
                  cp_async4(&s_c[es], col + k0 + j);
                  cp_async8(&s_v[es], val + k0 + j);
                }
            }
          }
          cp_commit();
        };

        // (A variant that issues untracked prefetch.global.L2 hints far ahead and copies
        // L2-resident data at the start of each iteration measured slower on B200:
        // 4.4 vs 3.55 us per level, profiles/r1_ssor_notes.md.)
        for (int l = -4; l < nl; l++) {
          if (l >= 0) {
            cp_wait<2>();                     // data of level l is in shared memory
            const int ms = (l % kMetaRing) * T + tid, ds = (l % kDataRing) * T + tid;
            const int r = s_r[ms];
            if (r >= 0) {
              const int k0 = s_k0[ms], k1 = s_k1[ms];
              const double d = s_d[ds], br = s_b[ds];
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < kPre; j++)
                if (k0 + j < k1) {
                  const int es = ((l % kDataRing) * kPre + j) * T + tid;
                  const double pr = __dmul_rn(s_v[es], __ldcg(a.x + s_c[es]));
                  s = RELAX ? __dsub_rn(s, pr) : __dadd_rn(s, pr);
                }
              for (int k = k0 + kPre; k < k1; k++) {
                const double pr = __dmul_rn(__ldg(val + k), __ldcg(a.x + __ldg(col + k)));
                s = RELAX ? __dsub_rn(s, pr) : __dadd_rn(s, pr);
              }
              if (!RELAX) {
                const double wr = need_w ? s_w[ds] : 0.0;
                a.x[r] = __ddiv_rn(__dsub_rn(__dsub_rn(br, wr), s), d);
                a.w[r] = s;
              } else {
                const double om = a.omega, om1 = 1.0 - a.omega;
                double temp;
                if (!need_xo) temp = __dmul_rn(om, br);
                else temp = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(om1, s_xo[ds]), d), s_w[ds]), __dmul_rn(om, br));
                const double h = __dmul_rn(om, s);
                a.w[r] = h;
                a.x[r] = __ddiv_rn(__dadd_rn(temp, h), d);
              }
            }
            // levels wider than the cluster: remaining rows straight from global memory
            for (int pp = s_lptr[l] + ctid + csz; pp < s_lptr[l + 1]; pp += csz)
              sweep_row<RELAX>(a, fwd, step, __ldg(rows + pp), __ldg(ind + pp), __ldg(ind + pp + 1), col, val);
            cluster_arrive();
          }
          issue_meta(l + 4);
          cp_wait<4>();                       // metadata of level l+2 has landed
          issue_data(l + 2);
          if (l >= 0) cluster_wait();
        }
        cp_wait<0>();
      }
    }
  }
}

// ---------------------------------------------------------------- single-CTA engine
// For long chains of narrow levels the cheapest barrier by far is bar.sync
// (measured on B200: 0.02 us vs 0.24 us for an empty cluster barrier and ~0.85 us once
// a global store has to be released to another SM).  One CTA walks all levels; a row
// reads the x values produced in the PREVIOUS level from shared memory (loc[k] >= 0 is
// the position of column k inside that level) and older ones from global memory (same
// SM, ordered by the barrier).  DRAM latency is taken off the critical path with L2
// prefetch hints issued a few levels ahead.
constexpr int kCtaThreads = 1024;
constexpr int kCtaRows = 8;          // rows per thread and level that get scattered-vector hints

template <bool RELAX>
__global__ void __launch_bounds__(kCtaThreads, 1) k_sweep_cta(SweepArgs a, const int *floc, const int *bloc, int maxw) {
  if (a.S && *(const volatile int *)&a.S->done) return;
  extern __shared__ double xs[];      // [2][maxw]
  constexpr int T = kCtaThreads;
  const int tid = threadIdx.x;
  for (int step = 0; step < a.steps; step++) {
    for (int half = 0; half < 2; half++) {
      const bool fwd = (half == 0);
      const int nlev = fwd ? a.nlev_f : a.nlev_b;
      const int *lptr = fwd ? a.lptr_f : a.lptr_b, *rows = fwd ? a.rows_f : a.rows_b;
      const int *ind = fwd ? a.find : a.bind, *col = fwd ? a.fcol : a.bcol, *loc = fwd ? floc : bloc;
      const double *val = fwd ? a.fval : a.bval;
      const bool need_w = !(fwd && step == 0);
      for (int l = 0; l < nlev; l++) {
        const int p0 = __ldg(lptr + l), p1 = __ldg(lptr + l + 1);
        double *cur = xs + (size_t)(l & 1) * maxw;
        const double *prev = xs + (size_t)((l & 1) ^ 1) * maxw;
        // rows of level l+2 whose scattered vector entries get an L2 hint below (loaded
        // now so that the latency overlaps the compute)
        int hr[kCtaRows];
        {
          const int q0 = (l + 2 < nlev) ? __ldg(lptr + l + 2) : 0, q1 = (l + 2 < nlev) ? __ldg(lptr + l + 3) : 0;
#pragma unroll
          for (int j = 0; j < kCtaRows; j++) { const int q = q0 + tid + j * T; hr[j] = (q < q1) ? __ldg(rows + q) : -1; }
        }
        for (int p = p0 + tid; p < p1; p += T) {
          const int r = __ldg(rows + p), k0 = __ldg(ind + p), k1 = __ldg(ind + p + 1);
          const double d = __ldg(a.diag + r), br = __ldg(a.b + r);
          double s = 0.0;
          for (int k = k0; k < k1; k++) {
            const int lc = __ldg(loc + k);
            const double xv = (lc >= 0) ? prev[lc] : a.x[__ldg(col + k)];
            const double pr = __dmul_rn(__ldg(val + k), xv);
            s = RELAX ? __dsub_rn(s, pr) : __dadd_rn(s, pr);
          }
          double xn;
          if (!RELAX) {
            const double wr = need_w ? a.w[r] : 0.0;
            xn = __ddiv_rn(__dsub_rn(__dsub_rn(br, wr), s), d);
            a.w[r] = s;
          } else {
            const double om = a.omega, om1 = 1.0 - a.omega;
            double temp;
            if (!need_w) temp = __dmul_rn(om, br);
            else temp = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(om1, a.x[r]), d), a.w[r]), __dmul_rn(om, br));
            const double h = __dmul_rn(om, s);
            a.w[r] = h;
            xn = __ddiv_rn(__dadd_rn(temp, h), d);
          }
          a.x[r] = xn;
          cur[p - p0] = xn;
        }
        // L2 hints: scattered vector entries of level l+2, streamed arrays of level l+4
#pragma unroll
        for (int j = 0; j < kCtaRows; j++)
          if (hr[j] >= 0) {
            prefetch_l2(a.diag + hr[j]); prefetch_l2(a.b + hr[j]);
            if (need_w) prefetch_l2(a.w + hr[j]);
          }
        if (l + 4 < nlev) {
          const int q0 = __ldg(lptr + l + 4), q1 = __ldg(lptr + l + 5);
          for (int q = q0 + tid * 32; q < q1; q += T * 32) { prefetch_l2(rows + q); prefetch_l2(ind + q); }
          const int e0 = __ldg(ind + q0), e1 = __ldg(ind + q1);
          for (int e = e0 + tid * 16; e < e1; e += T * 16) { prefetch_l2(col + e); prefetch_l2(loc + e); prefetch_l2(val + e); }
        }
        __syncthreads();
      }
    }
  }
}

static int grid_engine_blocks(const void *kernel) {
  int *slot = (kernel == (const void *)k_sweep_grid<false>) ? &ctx().occ_sweep_grid[0] : &ctx().occ_sweep_grid[1];
  if (*slot == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    *slot = per_sm * ctx().sms;
  }
  return *slot;
}

static int upload_i(const std::vector<int> &v, int **d);
static int upload_d(const std::vector<double> &v, double **d);

// ---------------------------------------------------------------- point-to-point engine (host side)
__global__ void k_fill_bits(unsigned long long *p, size_t n, unsigned long long v, const State *S) {
  if (S && *(const volatile int *)&S->done) return;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

template <bool RELAX>
static const void *p2p_kernel(int D) {
  if (D <= 2) return (const void *)k_sweep_p2p<RELAX, 2>;
  if (D <= 4) return (const void *)k_sweep_p2p<RELAX, 4>;
  return (const void *)k_sweep_p2p<RELAX, 8>;
}
static int p2p_width(int D) { return D <= 2 ? 2 : (D <= 4 ? 4 : 8); }
static size_t p2p_smem(int D) { return D <= 2 ? p2p_smem_bytes<2>() : (D <= 4 ? p2p_smem_bytes<4>() : p2p_smem_bytes<8>()); }

// One application: 2 * steps half sweeps, each one cooperative launch of nwarps single-warp
// CTAs.  The half sweeps ping-pong between a scratch vector and dy (forward -> scratch,
// backward -> dy), each target pre-filled with the sentinel.
static int p2p_apply(psp_prec *K, const double *dx, double *dy, State *gateS, bool relax) {
  Ctx &c = ctx();
  const size_t n = (size_t)K->n;
  if (!K->d_t2) {
    if (cudaMalloc((void **)&K->d_t2, sizeof(double) * n) != cudaSuccess) { cudaGetLastError(); set_error("ssor: out of device memory"); return PSP_ERR_NOMEM; }
  }
  PSP_CUDA(cudaMemsetAsync(K->d_abort, 0, sizeof(int), c.stream));
  for (int step = 0; step < K->steps; step++)
    for (int half = 0; half < 2; half++) {
      const psp_prec::P2PSched &sc = K->p2p[half];
      double *xnew = half == 0 ? K->d_t2 : dy;
      const double *xold = half == 0 ? dy : K->d_t2;
      k_fill_bits<<<grid_for(n), kThreads, 0, c.stream>>>(reinterpret_cast<unsigned long long *>(xnew), n, kSweepSentinel, gateS);
      PSP_LAUNCH_CHECK();
      P2PArgs a;
      a.dir.wbase = sc.d_wbase; a.dir.rec = sc.d_rec;
      a.dir.nwarps = sc.nwarps; a.dir.D = p2p_width(sc.D);
      a.b = dx; a.xold = xold; a.xnew = xnew; a.w = K->d_t1; a.omega = K->omega;
      a.first = (step == 0 && half == 0) ? 1 : 0;
      a.reverse = half; a.n = K->n;
      a.abort_flag = K->d_abort; a.S = gateS;
      const void *kern = relax ? p2p_kernel<true>(sc.D) : p2p_kernel<false>(sc.D);
      void *args[] = {&a};
      PSP_CUDA(cudaLaunchCooperativeKernel(kern, dim3(sc.nwarps), dim3(64), args, p2p_smem(sc.D), c.stream));
      c.launches++;
    }
  return PSP_OK;
}

// Chain schedule of one sweep direction.  ind/col/val: the rows of the dependency matrix (L
// for the forward sweep; L^T with every row's entries by descending original row for the
// backward sweep); reverse: rows are swept n-1 .. 0.  Returns false when the matrix does not
// suit the engine (long rows, too many warps to be co-resident, too much padding).
static bool p2p_build(int n, const std::vector<int> &ind, const std::vector<int> &col, const std::vector<double> &val,
                      const double *diag, bool reverse, int max_warps, psp_prec::P2PSched &out, int *status) {
  *status = PSP_OK;
  if (n == 0) return false;
  int D = 0;
  for (int i = 0; i < n; i++) D = std::max(D, ind[i + 1] - ind[i]);
  if (D > kP2PMaxDeps) return false;
  if (D == 0) D = 1;
  const int W = p2p_width(D);
  // chains: a row continues the chain of its NEAREST not yet continued dependency (entries are
  // stored farthest-back first, so the row's entries are tried last to first: on a
  // lexicographic grid the nearest one is the previous point of the same grid line, a chain is
  // a grid line, all chains have the same length and consecutive chains are neighbouring lines)
  std::vector<int> chain((size_t)n, -1), pos((size_t)n, 0);
  std::vector<char> claimed((size_t)n, 0);
  std::vector<int> len;
  for (int p = 0; p < n; p++) {
    const int r = reverse ? n - 1 - p : p;
    int parent = -1;
    for (int k = ind[r + 1] - 1; k >= ind[r]; k--)
      if (!claimed[(size_t)col[k]]) { parent = col[k]; break; }
    if (parent >= 0) {
      claimed[(size_t)parent] = 1;
      chain[(size_t)r] = chain[(size_t)parent]; pos[(size_t)r] = pos[(size_t)parent] + 1;
      len[(size_t)chain[(size_t)r]] = pos[(size_t)r] + 1;
    } else {
      chain[(size_t)r] = (int)len.size(); pos[(size_t)r] = 0;
      len.push_back(1);
    }
  }
  const int nchains = (int)len.size();
  const int nwarps = (nchains + 31) / 32;
  if (nwarps > max_warps) return false;
  std::vector<int> wbase((size_t)nwarps + 1, 0);
  for (int w = 0; w < nwarps; w++) {
    int lmax = 0;
    for (int cidx = 32 * w; cidx < std::min(nchains, 32 * w + 32); cidx++) lmax = std::max(lmax, len[(size_t)cidx]);
    wbase[(size_t)w + 1] = wbase[(size_t)w] + lmax;
  }
  const long long nslots = wbase[(size_t)nwarps];
  if (nslots * 32 > 3LL * n + 4096) return false;          // mostly padding: chains too uneven
  // packed records: (row, pad, diag, W x (dep row, producing lane), W x value) per (slot, lane)
  const size_t rb = 16 + 16 * (size_t)W;
  std::vector<unsigned char> rec((size_t)nslots * 32 * rb, 0);
  auto rec_at = [&](size_t slot, int lane) { return rec.data() + (slot * 32 + (size_t)lane) * rb; };
  for (size_t e = 0; e < (size_t)nslots * 32; e++) {
    unsigned char *p = rec.data() + e * rb;
    const int minus1 = -1; const double one = 1.0;
    memcpy(p, &minus1, 4); memcpy(p + 8, &one, 8);
    for (int d = 0; d < W; d++) { memcpy(p + 16 + 8 * d, &minus1, 4); memcpy(p + 16 + 8 * d + 4, &minus1, 4); }
  }
  for (int r = 0; r < n; r++) {
    const int cidx = chain[(size_t)r], w = cidx / 32, lane = cidx % 32;
    unsigned char *p = rec_at((size_t)wbase[(size_t)w] + (size_t)pos[(size_t)r], lane);
    memcpy(p, &r, 4); memcpy(p + 8, &diag[r], 8);
    int d = 0;
    for (int k = ind[r]; k < ind[r + 1]; k++, d++) {
      const int dc = chain[(size_t)col[k]];
      // producer inside this warp: (its position in its chain << 5) | its lane
      const int srcl = (dc / 32 == w) ? ((pos[(size_t)col[k]] << 5) | (dc % 32)) : -1;
      memcpy(p + 16 + 8 * d, &col[k], 4); memcpy(p + 16 + 8 * d + 4, &srcl, 4);
      memcpy(p + 16 + 8 * (size_t)W + 8 * d, &val[k], 8);
    }
  }
  out.nwarps = nwarps; out.D = D; out.nslots = nslots;
  int st = upload_i(wbase, &out.d_wbase);
  if (st == PSP_OK) {
    if (cudaMalloc(&out.d_rec, rec.size() + 64) != cudaSuccess) { cudaGetLastError(); set_error("ssor: out of device memory"); st = PSP_ERR_NOMEM; }
    else if (cudaMemcpy(out.d_rec, rec.data(), rec.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); set_error("ssor: upload failed"); st = PSP_ERR_CUDA; }
  }
  *status = st;
  return st == PSP_OK;
}

int prec_apply_dev(psp_prec *K, const double *dx, double *dy, State *gateS) {
  Ctx &c = ctx();
  const long n = K->n;
  if (n == 0) return PSP_OK;
  if (K->ceq) return correq_apply(K->ceq, 1, dx, dy, gateS, 0);
  if (K->kind == 1) {
    const int gate = gateS ? 0 : 3;
    PSP_TRY(launch_map(FJacobiScale{dx, K->d_dinv, dy}, n, EPI_NONE, gate));
    for (int step = 1; step < K->steps; step++) {
      PSP_CUDA(cudaMemcpyAsync(K->d_t1, dy, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
      PSP_TRY(spmv_launch(K->mat, K->d_t1, dy, SpmvDots()));
      PSP_TRY(launch_map(FJacobiCorr{dx, K->d_dinv, K->d_t1, dy}, n, EPI_NONE, gate));
    }
    return PSP_OK;
  }
  SweepArgs a;
  a.n = (int)n;
  a.diag = K->mat->d_diag;
  a.rows_f = K->d_rows_f; a.lptr_f = K->d_lptr_f; a.find = K->d_find; a.fcol = K->d_fcol; a.fval = K->d_fval;
  a.nlev_f = K->nlev_f;
  a.rows_b = K->d_rows_b; a.lptr_b = K->d_lptr_b; a.bind = K->d_bind; a.bcol = K->d_bcol; a.bval = K->d_bval;
  a.nlev_b = K->nlev_b;
  a.b = dx; a.x = dy; a.w = K->d_t1; a.omega = K->omega; a.steps = K->steps; a.S = gateS;
  const bool relax = (K->omega != 1.0);
  // point-to-point engine: long chains of narrow levels (ssor_engine 0 automatic, 4 forced)
  if (K->p2p_ok && (c.ssor_engine == 4 || (c.ssor_engine == 0 && (long)K->nlev_f + K->nlev_b >= 64)))
    return p2p_apply(K, dx, dy, gateS, relax);
  // engine choice: many narrow levels -> one cluster with the hardware barrier; few wide
  // levels -> the whole GPU with grid.sync.  ssor_engine: 0 auto, 1 grid, 2 cluster.
  const int maxw = K->maxw_f > K->maxw_b ? K->maxw_f : K->maxw_b;
  const long nlev = (long)K->nlev_f + K->nlev_b;
  bool use_cluster = (maxw <= 4 * c.ssor_cluster * kSweepThreads) && nlev >= 64;
  if (c.ssor_engine == 1) use_cluster = false;
  if (c.ssor_engine == 2) use_cluster = true;
  // single CTA: chains of narrow levels (lexicographic 2-D grids); needs 2 x maxw doubles of smem
  // (experimental, never chosen automatically: one SM cannot stream the operands fast
  // enough -- 9.8 us per level at 4096^2 vs 3.6 for the cluster engine)
  const bool use_cta = (c.ssor_engine == 3) && maxw <= 12288;
  if (use_cta) {
    const void *kern = relax ? (const void *)k_sweep_cta<true> : (const void *)k_sweep_cta<false>;
    const size_t smem = sizeof(double) * 2 * (size_t)(maxw > 0 ? maxw : 1);
    size_t *prepared_smem = c.sweep_cta_smem;
    if (prepared_smem[relax ? 1 : 0] < smem) {
      PSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      prepared_smem[relax ? 1 : 0] = smem;
    }
    int mw = maxw;
    const int *fl = K->d_floc, *bl = K->d_bloc;
    void *args[] = {&a, &fl, &bl, &mw};
    PSP_CUDA(cudaLaunchKernel(kern, dim3(1), dim3(kCtaThreads), args, smem, c.stream));
    c.launches++;
    return PSP_OK;
  }
  if (use_cluster) {
    const void *kern = relax ? (const void *)k_sweep_cluster<true> : (const void *)k_sweep_cluster<false>;
    int csize = (int)c.ssor_cluster;
    bool *prepared = c.sweep_cluster_ready;
    if (!prepared[relax ? 1 : 0]) {
      PSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSweepSmem));
      PSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      prepared[relax ? 1 : 0] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize); cfg.blockDim = dim3(kSweepThreads); cfg.dynamicSmemBytes = kSweepSmem; cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    void *args[] = {&a};
    PSP_CUDA(cudaLaunchKernelExC(&cfg, kern, args));
    c.launches++;
    return PSP_OK;
  }
  const void *kern = relax ? (const void *)k_sweep_grid<true> : (const void *)k_sweep_grid<false>;
  void *args[] = {&a};
  PSP_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid_engine_blocks(kern)), dim3(kThreads), args, 0, c.stream));
  c.launches++;
  return PSP_OK;
}

static int upload_i(const std::vector<int> &v, int **d) {
  size_t bytes = sizeof(int) * (v.size() ? v.size() : 1);
  if (cudaMalloc((void **)d, bytes) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory"); return PSP_ERR_NOMEM; }
  if (!v.empty()) PSP_CUDA(cudaMemcpy(*d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
  return PSP_OK;
}
static int upload_d(const std::vector<double> &v, double **d) {
  size_t bytes = sizeof(double) * (v.size() ? v.size() : 1);
  if (cudaMalloc((void **)d, bytes) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory"); return PSP_ERR_NOMEM; }
  if (!v.empty()) PSP_CUDA(cudaMemcpy(*d, v.data(), sizeof(double) * v.size(), cudaMemcpyHostToDevice));
  return PSP_OK;
}

// level-sorted row lists (counting sort: rows ascending inside a level); returns the widest level
static int sort_by_level(const std::vector<int> &level, int nlev, std::vector<int> &rows, std::vector<int> &lptr) {
  const int n = (int)level.size();
  lptr.assign((size_t)nlev + 1, 0);
  for (int i = 0; i < n; i++) lptr[level[i] + 1]++;
  int maxw = 0;
  for (int l = 0; l < nlev; l++) { if (lptr[l + 1] > maxw) maxw = lptr[l + 1]; lptr[l + 1] += lptr[l]; }
  std::vector<int> fill(lptr.begin(), lptr.end() - 1);
  rows.resize((size_t)n);
  for (int i = 0; i < n; i++) rows[fill[level[i]]++] = i;
  return maxw;
}

// level-major copy of a CSR structure: position p holds row rows[p].  oloc[k] is the
// position of column ocol[k] inside its level when that level is the one just before
// the row's level (its x value is then still in shared memory), else -1.
static void level_major(const std::vector<int> &rows, const std::vector<int> &lptr, const std::vector<int> &level,
                        const std::vector<int> &ind, const std::vector<int> &col, const std::vector<double> &val,
                        std::vector<int> &oind, std::vector<int> &ocol, std::vector<double> &oval,
                        std::vector<int> &oloc) {
  const size_t n = rows.size();
  std::vector<int> pos(n);
  for (size_t p = 0; p < n; p++) pos[rows[p]] = (int)p;
  oind.assign(n + 1, 0);
  ocol.resize(col.size());
  oval.resize(val.size());
  oloc.resize(col.size());
  int q = 0;
  for (size_t p = 0; p < n; p++) {
    const int r = rows[p];
    for (int k = ind[r]; k < ind[r + 1]; k++) {
      const int c = col[k];
      ocol[q] = c; oval[q] = val[k];
      oloc[q] = (level[c] == level[r] - 1) ? pos[c] - lptr[level[c]] : -1;
      q++;
    }
    oind[p + 1] = q;
  }
}

int ssor_create(psp_mat *A, double omega, int steps, psp_prec **out) {
  if (!A || A->kind != 2) { set_error("ssor: matrix must be an sss_mat"); return PSP_ERR_ARG; }
  const int n = A->n;
  const std::vector<int> &ind = A->h_ind, &col = A->h_col;
  const std::vector<double> &val = A->h_val;
  psp_prec *K = new psp_prec();
  K->kind = 2; K->n = n; K->omega = omega; K->steps = steps; K->mat = A;
  // forward levels
  std::vector<int> &lf = K->h_level_f, &lb = K->h_level_b;
  lf.assign((size_t)n, 0); lb.assign((size_t)n, 0);
  int nlf = 0, nlb = 0;
  for (int i = 0; i < n; i++) {
    int lv = 0;
    for (int k = ind[i]; k < ind[i + 1]; k++) { int c1 = lf[col[k]] + 1; if (c1 > lv) lv = c1; }
    lf[i] = lv;
    if (lv + 1 > nlf) nlf = lv + 1;
  }
  // backward levels: row i waits for every later row i' with L[i',i] != 0
  for (int i = n - 1; i >= 0; i--) {
    if (lb[i] + 1 > nlb) nlb = lb[i] + 1;
    for (int k = ind[i]; k < ind[i + 1]; k++) { int j = col[k]; if (lb[i] + 1 > lb[j]) lb[j] = lb[i] + 1; }
  }
  if (n == 0) { nlf = 0; nlb = 0; }
  K->nlev_f = nlf; K->nlev_b = nlb;
  // transpose index of L, entries of each row by descending original row
  const size_t nnz = col.size();
  std::vector<int> tind((size_t)n + 1, 0), tcol(nnz), fill((size_t)(n > 0 ? n : 1));
  std::vector<double> tval(nnz);
  for (size_t k = 0; k < nnz; k++) tind[col[k] + 1]++;
  for (int j = 0; j < n; j++) tind[j + 1] += tind[j];
  for (int j = 0; j < n; j++) fill[j] = tind[j];
  for (int i = n - 1; i >= 0; i--)
    for (int k = ind[i]; k < ind[i + 1]; k++) { int j = col[k]; tcol[fill[j]] = i; tval[fill[j]] = val[k]; fill[j]++; }

  std::vector<int> rows, lptr, oind, ocol, oloc;
  std::vector<double> oval;
  int st = PSP_OK;
  K->maxw_f = sort_by_level(lf, nlf, rows, lptr);
  level_major(rows, lptr, lf, ind, col, val, oind, ocol, oval, oloc);
  if (st == PSP_OK) st = upload_i(oloc, &K->d_floc);
  if (st == PSP_OK) st = upload_i(rows, &K->d_rows_f);
  if (st == PSP_OK) st = upload_i(lptr, &K->d_lptr_f);
  if (st == PSP_OK) st = upload_i(oind, &K->d_find);
  if (st == PSP_OK) st = upload_i(ocol, &K->d_fcol);
  if (st == PSP_OK) st = upload_d(oval, &K->d_fval);
  K->maxw_b = sort_by_level(lb, nlb, rows, lptr);
  level_major(rows, lptr, lb, tind, tcol, tval, oind, ocol, oval, oloc);
  if (st == PSP_OK) st = upload_i(oloc, &K->d_bloc);
  if (st == PSP_OK) st = upload_i(rows, &K->d_rows_b);
  if (st == PSP_OK) st = upload_i(lptr, &K->d_lptr_b);
  if (st == PSP_OK) st = upload_i(oind, &K->d_bind);
  if (st == PSP_OK) st = upload_i(ocol, &K->d_bcol);
  if (st == PSP_OK) st = upload_d(oval, &K->d_bval);
  if (st == PSP_OK && cudaMalloc((void **)&K->d_t1, sizeof(double) * (size_t)(n > 0 ? n : 1)) != cudaSuccess) {
    cudaGetLastError(); st = PSP_ERR_NOMEM;
  }
  if (st == PSP_OK && ctx().ssor_engine != 1 && ctx().ssor_engine != 2 && ctx().ssor_engine != 3) {
    // point-to-point engine: chain schedules of both sweeps, if the matrix suits it and all
    // warps can be co-resident
    int D = 0;
    for (int i = 0; i < n; i++) D = std::max(D, ind[i + 1] - ind[i]);
    for (int j = 0; j < n; j++) D = std::max(D, tind[j + 1] - tind[j]);
    if (n > 0 && D <= kP2PMaxDeps) {
      int per_sm = 0;
      const void *k0 = p2p_kernel<false>(D), *k1 = p2p_kernel<true>(D);
      int a0 = 0, a1 = 0;
      const size_t smem = p2p_smem(D);
      if (cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
          cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a0, k0, 64, smem) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a1, k1, 64, smem) == cudaSuccess)
        per_sm = std::min(a0, a1);
      else cudaGetLastError();
      const int max_warps = per_sm * ctx().sms;
      int s1 = PSP_OK, s2 = PSP_OK;
      const bool ok = max_warps > 0 &&
                      p2p_build(n, ind, col, val, A->h_diag.data(), false, max_warps, K->p2p[0], &s1) &&
                      p2p_build(n, tind, tcol, tval, A->h_diag.data(), true, max_warps, K->p2p[1], &s2);
      if (s1 != PSP_OK || s2 != PSP_OK) st = (s1 != PSP_OK ? s1 : s2);
      K->p2p_ok = ok && st == PSP_OK;
      if (K->p2p_ok && cudaMalloc((void **)&K->d_abort, sizeof(int)) != cudaSuccess) { cudaGetLastError(); st = PSP_ERR_NOMEM; }
    }
  }
  if (st != PSP_OK) { psp_prec_destroy(K); if (st == PSP_ERR_NOMEM) set_error("ssor: out of device memory"); return st; }
  *out = K;
  return PSP_OK;
}

}  // namespace psp

using namespace psp;

extern "C" int psp_jacobi_create(psp_mat *A, int n, const double *diag, double omega, int steps,
                                 psp_prec **out) {
  if (!out || n < 0 || (n > 0 && !diag)) { set_error("psp_jacobi_create: bad arguments"); return PSP_ERR_ARG; }
  if (steps > 1 && !A) { set_error("psp_jacobi_create: steps > 1 needs the matrix"); return PSP_ERR_ARG; }
  std::vector<double> dinv((size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    const double d = diag[i];
    volatile double t = 1.0 + d;             // preconmodule.c:395
    if (t == 1.0) { set_error("diagonal element close to zero"); return PSP_ERR_ARG; }
    dinv[i] = omega / d;                     // preconmodule.c:400
  }
  PSP_TRY(ensure_init());
  psp_prec *K = new psp_prec();
  K->kind = 1; K->n = n; K->omega = omega; K->steps = steps; K->mat = A;
  size_t bytes = sizeof(double) * (size_t)(n > 0 ? n : 1);
  if (cudaMalloc((void **)&K->d_dinv, bytes) != cudaSuccess ||
      (steps > 1 && cudaMalloc((void **)&K->d_t1, bytes) != cudaSuccess)) {
    cudaGetLastError();
    psp_prec_destroy(K);
    set_error("psp_jacobi_create: out of device memory");
    return PSP_ERR_NOMEM;
  }
  if (n > 0 && cudaMemcpy(K->d_dinv, dinv.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) {
    psp_prec_destroy(K);
    set_error("psp_jacobi_create: upload failed");
    return PSP_ERR_CUDA;
  }
  if (multi_gpus() > 1) K->h_dinv.assign(dinv.begin(), dinv.begin() + n);
  *out = K;
  return PSP_OK;
}

namespace psp {
// single-step Jacobi from a ready dinv slice (one rank of the multi-GPU path)
int jacobi_from_dinv(int n, const double *dinv, psp_prec **out) {
  psp_prec *K = new psp_prec();
  K->kind = 1; K->n = n; K->omega = 1.0; K->steps = 1;
  const size_t bytes = sizeof(double) * (size_t)(n > 0 ? n : 1);
  if (cudaMalloc((void **)&K->d_dinv, bytes) != cudaSuccess) { cudaGetLastError(); delete K; set_error("jacobi: out of device memory"); return PSP_ERR_NOMEM; }
  if (n > 0 && cudaMemcpy(K->d_dinv, dinv, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError(); cudaFree(K->d_dinv); delete K; set_error("jacobi: upload failed"); return PSP_ERR_CUDA;
  }
  *out = K;
  return PSP_OK;
}
}  // namespace psp

extern "C" int psp_ssor_create(psp_mat *sss, double omega, int steps, psp_prec **out) {
  if (!out) { set_error("psp_ssor_create: NULL out"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  return ssor_create(sss, omega, steps, out);
}

extern "C" int psp_prec_set_matrix(psp_prec *K, psp_mat *A) {
  if (!K || !A) { set_error("psp_prec_set_matrix: NULL argument"); return PSP_ERR_ARG; }
  if (K->kind != 1) { set_error("psp_prec_set_matrix: only Jacobi preconditioners re-bind their matrix"); return PSP_ERR_ARG; }
  if (A->m != K->n || A->n != K->n) { set_error("psp_prec_set_matrix: matrix order differs from the preconditioner's"); return PSP_ERR_ARG; }
  K->mat = A;
  return PSP_OK;
}

extern "C" void psp_prec_destroy(psp_prec *K) {
  if (!K || K->ceq) return;                        // a facade belongs to its psp_correq
  multi_prec_release(K);
  cudaFree(K->d_dinv); cudaFree(K->d_t1); cudaFree(K->d_t2);
  cudaFree(K->d_rows_f); cudaFree(K->d_lptr_f); cudaFree(K->d_rows_b); cudaFree(K->d_lptr_b);
  cudaFree(K->d_find); cudaFree(K->d_fcol); cudaFree(K->d_fval);
  cudaFree(K->d_bind); cudaFree(K->d_bcol); cudaFree(K->d_bval); cudaFree(K->d_floc); cudaFree(K->d_bloc);
  cudaFree(K->d_x); cudaFree(K->d_y); cudaFree(K->d_bar); cudaFree(K->d_abort);
  for (int h = 0; h < 2; h++) {
    cudaFree(K->p2p[h].d_wbase); cudaFree(K->p2p[h].d_rec);
  }
  delete K;
}

extern "C" int psp_prec_apply_dev(psp_prec *K, const double *dx, double *dy) {
  if (!K || !dx || !dy) { set_error("psp_prec_apply_dev: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  return prec_apply_dev(K, dx, dy, nullptr);
}

extern "C" int psp_prec_apply(psp_prec *K, const double *x, double *y) {
  if (!K || (K->n > 0 && (!x || !y))) { set_error("psp_prec_apply: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  Ctx &c = ctx();
  const size_t n = (size_t)K->n;
  if (!K->d_x) {
    if (cudaMalloc((void **)&K->d_x, sizeof(double) * (n ? n : 1)) != cudaSuccess ||
        cudaMalloc((void **)&K->d_y, sizeof(double) * (n ? n : 1)) != cudaSuccess) {
      cudaGetLastError();
      set_error("psp_prec_apply: out of device memory");
      return PSP_ERR_NOMEM;
    }
  }
  PSP_TRY(psp_vec_upload(K->d_x, x, n, 1));
  PSP_CUDA(cudaEventRecord(c.ev0, c.stream));
  PSP_TRY(prec_apply_dev(K, K->d_x, K->d_y, nullptr));
  PSP_CUDA(cudaEventRecord(c.ev1, c.stream));
  PSP_TRY(psp_vec_download(K->d_y, y, n, 1));
  if (K->d_abort) {
    int gave_up = 0;
    PSP_CUDA(cudaMemcpy(&gave_up, K->d_abort, sizeof(int), cudaMemcpyDeviceToHost));
    if (gave_up) { set_error("ssor: a sweep gave up waiting for a row it depends on (NaN payload clash in the input?)"); return PSP_ERR_CUDA; }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  c.last_ms = ms;
  return PSP_OK;
}

extern "C" int psp_ssor_levels(const psp_prec *K, int *nlev_fwd, int *nlev_bwd, int *level_fwd,
                               int *level_bwd) {
  if (!K || K->kind != 2) { set_error("psp_ssor_levels: not an SSOR preconditioner"); return PSP_ERR_ARG; }
  if (nlev_fwd) *nlev_fwd = K->nlev_f;
  if (nlev_bwd) *nlev_bwd = K->nlev_b;
  if (level_fwd) memcpy(level_fwd, K->h_level_f.data(), sizeof(int) * (size_t)K->n);
  if (level_bwd) memcpy(level_bwd, K->h_level_b.data(), sizeof(int) * (size_t)K->n);
  return PSP_OK;
}
