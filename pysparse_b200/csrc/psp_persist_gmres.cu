// psp_persist_gmres.cu -- single-launch restarted GMRES(m): all restart cycles of a solve in
// one persistent cooperative kernel (infrastructure: psp_persist.cuh).
//
// STATUS: written at the end of round 1, compiled for sm_100a but NOT yet run on a GPU.
// OFF by default (option gmres_persist = 0; solve_gmres keeps the validated per-phase kernel
// sequence).  Bring-up: scripts/gmres_persist_probe.py.
//
// Why: the per-phase path launches i + 4 kernels for inner step i (preconditioner scaling,
// SpMV + dot, i + 1 fused "axpy + next dot" modified Gram-Schmidt steps, normalisation), each
// ending in a grid reduction -- at 8 GPUs (1 M rows per GPU, vectors resident in L2) launch
// and all-reduce latency dominate (2 617 it/s on 8 GPUs vs 897 on one: 36 % efficiency).
// Here every one of those steps is a vector phase followed by the ticket/flag all-reduce,
// which is also the grid barrier (~2-3 us instead of a kernel boundary + MEMBAR.SYS).
//
// Differences from the PCG kernel, forced by the algorithm:
//   * the scalar epilogues of GMRES (gmres.c:143-174: Hessenberg column, Givens rotations,
//     back substitution) update GLOBAL arrays and must run exactly once: they run in the CTA
//     that draws the last ticket, on the global State, before the others are released
//     (grid_allreduce_op); all threads then read the few scalars they need from the global
//     State (ld.cg);
//   * halo exchange by "late push" + block-local unpack: the vector to multiply is complete
//     at a barrier, the boundary entries are then sent as they are (self-validating words)
//     and every boundary row block unpacks what it reads right before it is multiplied --
//     no recomputation of the sent values, hence no double buffering of the Krylov basis.
#include "psp_persist.cuh"

namespace psp {

struct GmresArgs : PersistArgs {
  const double *b;
  double *x;
  double *V, *W;                // basis (m + 1 vectors) and preconditioned basis (m vectors; == V if K = I)
  double *xs;                   // sharded: x with halo room (x itself otherwise)
  const double *dinv;           // nullptr: identity preconditioner
  const double *g;              // GmresSmall::g (the rotated right-hand side / solution of the small system)
  size_t N;                     // distance between basis vectors
  int m;                        // restart length
};

constexpr int kGmresMaxM = 256;                   // restart lengths the kernel caches g[] for

template <class T>
__device__ __forceinline__ T sload(const T *p) { return *reinterpret_cast<const volatile T *>(p); }

// exchange (sharded only) of `vec`, complete at the preceding barrier; returns true when the
// boundary blocks have to unpack block-locally (push model)
__device__ __forceinline__ bool gmres_exchange(cg::grid_group &grid, PersistCtx &c, const GmresArgs &a, double *vec,
                                               bool push) {
  if (!c.dd) return false;
  if (push) { halo_push_late(c, a, vec); return true; }
  publish_and_pull(grid, c, a, vec, (size_t)(vec - c.dd->heap[c.dd->rank]));
  return false;
}

template <int FMT>
__global__ void __launch_bounds__(kThreads, persist_ctas(FMT)) k_gmres_persist(GmresArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[FMT == 0 ? kPersistStages : 1];
  __shared__ __align__(16) unsigned char sseg_raw[FMT == 1 ? sizeof(RunSmem) : 16];   // (format 0 stages its descriptors in the TMA ring)
  RunSmem *sseg = reinterpret_cast<RunSmem *>(sseg_raw);
  __shared__ double red_sh[8 * (kRedSlots + 1)];
  __shared__ double tot[kRedSlots + 1];
  __shared__ double gsh[kGmresMaxM];
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const unsigned n = (unsigned)a.n;
  const unsigned gstride = gridDim.x * blockDim.x;
  const unsigned g0 = blockIdx.x * blockDim.x + tid;
  State *S = a.S;                                   // global: updated by one thread per barrier
  void *aux = (void *)(a.S + 1);                    // GmresSmall (psp_solvers.cu: state + 1)

  PersistCtx c;
  c.sS = nullptr; c.tot = tot; c.red_sh = red_sh; c.gen = 0; c.unpack_target = 0;
  c.dd = a.dd;
  c.debug = a.dist_debug;
  c.seq = c.pubs = 0;
  if (c.dd) { c.seq = (unsigned)c.dd->red_seq; c.pubs = (unsigned)c.dd->pubs_done; }
  const bool push = c.dd && a.push;
  const bool ident = (a.dinv == nullptr);

  // all-reduce + barrier + scalar epilogue `epi` (once, in the last CTA)
  auto reduce1 = [&](double v, int epi) {
    double acc[1] = {v};
    grid_allreduce_op<1, false>(c, a, acc, 0.0, [&](const double *t) {
      S->red[0] = t[0];
      run_epilogue(epi, S, aux);
      __threadfence();
    });
  };

  MatPass<FMT> ring;
  ring.init(smem, full, a.l2_stream != 0, sseg);
  ring.prefetch(a);

  // ---- ||b||  (gmres.c:100-108)
  {
    double acc = 0.0;
    for (unsigned i = g0; i < n; i += gstride) { double t = a.b[i]; acc += t * t; }
    reduce1(acc, EPI_GMRES_NORMB);
  }

  for (;;) {                                        // restart cycles (gmres.c:110-185) + the exit pass
    // ---- V0 = A x - b, beta = |V0|   (gmres.c:112-116; phase 3: true residual at exit :187-190)
    double *xin = a.x;
    if (c.dd) {
      for (unsigned i = g0; i < n; i += gstride) a.xs[i] = a.x[i];
      grid.sync();
      xin = a.xs;
    }
    {
      const bool unpack = gmres_exchange(grid, c, a, xin, push);
      if (unpack)
        ring.template pass<true>(a, xin, red_sh, [&](int r, double sum) { a.V[r] = sum; },
                                        [&](int b) { halo_unpack_block(c, a, xin, b); });
      else
        ring.template pass<false>(a, xin, red_sh, [&](int r, double sum) { a.V[r] = sum; }, [](int) {});
      ring.prefetch(a);
    }
    grid.sync();
    {
      double acc = 0.0;
      const bool zero_rhs = (sload(&S->phase) == 1);
      for (unsigned i = g0; i < n; i += gstride) {
        if (zero_rhs) { a.x[i] = 0.0; continue; }
        const double t = a.V[i] + (-1.0) * a.b[i];
        a.V[i] = t;
        acc += t * t;
      }
      reduce1(acc, EPI_GMRES_BETA);
    }
    if (sload(&S->done)) break;

    // ---- inner iterations  (gmres.c:127-166)
    const int m = a.m;
    for (int i = 0; i < m; i++) {
      if (sload(&S->cyc_done)) break;
      double *vi = a.V + (size_t)i * a.N, *vn = a.V + (size_t)(i + 1) * a.N;
      double *wi = ident ? vi : a.W + (size_t)i * a.N;
      {                                             // V(i) *= scale (V0: -1/beta, else 1/H(i,i-1)) ; W(i) = K^-1 V(i)
        const double f = sload(&S->v[i == 0 ? G_SCALE0 : G_INVH]);
        for (unsigned e = g0; e < n; e += gstride) {
          const double t = f * vi[e];
          vi[e] = t;
          if (!ident) wi[e] = t * a.dinv[e];
        }
      }
      grid.sync();
      {                                             // V(i+1) = A W(i) ; <V(i+1), V(0)>
        double acc = 0.0;
        const bool unpack = gmres_exchange(grid, c, a, wi, push);
        if (unpack)
          ring.template pass<true>(a, wi, red_sh, [&](int r, double sum) { vn[r] = sum; acc += sum * a.V[r]; },
                                          [&](int b) { halo_unpack_block(c, a, wi, b); });
        else
          ring.template pass<false>(a, wi, red_sh, [&](int r, double sum) { vn[r] = sum; acc += sum * a.V[r]; }, [](int) {});
        ring.prefetch(a);
        reduce1(acc, EPI_GMRES_H);                  // H(0,i)
      }
      for (int k = 0; k <= i; k++) {                // modified Gram-Schmidt (gmres.c:141-148)
        const double h = sload(&S->v[G_HK]);
        const double *vk = a.V + (size_t)k * a.N;
        const double *vnext = (k < i) ? a.V + (size_t)(k + 1) * a.N : nullptr;
        double acc = 0.0;
        for (unsigned e = g0; e < n; e += gstride) {
          const double t = vn[e] + (-h) * vk[e];
          vn[e] = t;
          acc += t * (vnext ? vnext[e] : t);
        }
        reduce1(acc, k < i ? EPI_GMRES_H : EPI_GMRES_HLAST);
      }
    }
    // the last V(i+1) of the cycle stays unnormalised: it is not used again (the per-phase path
    // normalises it, which changes nothing either)

    // ---- x += sum_j g[j] W(j)   (gmres.c:177-181), then the outer loop test (:183)
    {
      const int last = sload(&S->last_i);
      for (int j = tid; j <= last; j += blockDim.x) gsh[j] = sload(&a.g[j]);
      __syncthreads();
      const double *Wb = ident ? a.V : a.W;
      for (unsigned e = g0; e < n; e += gstride) {
        double xi = a.x[e];
        for (int j = 0; j <= last; j++) xi = xi + gsh[j] * Wb[(size_t)j * a.N + e];
        a.x[e] = xi;
      }
      double none[1] = {0.0};
      grid_allreduce_op<0, false>(c, a, none, 0.0, [&](const double *) {
        run_epilogue(EPI_GMRES_CYCLE, S, aux);
        __threadfence();
      });
    }
  }

  ring.drain(a);
  if (blockIdx.x == 0 && tid == 0 && c.dd) {        // 64-bit sequence numbers: add what this launch did
    const unsigned nred = c.seq - (unsigned)c.dd->red_seq, nexch = c.pubs - (unsigned)c.dd->pubs_done;
    c.dd->red_seq += nred; c.dd->pubs_done += nexch; c.dd->pulls_done += nexch;
  }
}

// Host entry: returns PSP_OK after ENQUEUEING the solve (one cooperative launch).
int gmres_persist_launch(const psp_mat *M, int n, const double *b, double *x, double *V, double *W, double *xs,
                         size_t N, int m, const double *dinv, const double *g, DistMat *dm) {
  Ctx &c = ctx();
  if (m > kGmresMaxM) { set_error("gmres_persist: restart length %d exceeds %d", m, kGmresMaxM); return PSP_ERR_ARG; }
  const int fmt = M->fmt;
  const void *kern = fmt == 1 ? (const void *)k_gmres_persist<1> : (const void *)k_gmres_persist<0>;
  const int smem_bytes = fmt == 1 ? 0 : kPersistStages * kStageBytes;
  int &resident = c.occ_gmres[fmt];
  if (resident == 0) {
    if (smem_bytes) PSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem_bytes) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (per_sm > persist_ctas(fmt)) per_sm = persist_ctas(fmt);
    resident = per_sm * c.sms;
    if (resident > kMaxGrid) resident = kMaxGrid;
  }
  GmresArgs a;
  PSP_TRY(persist_prepare(a, M, n, dm));
  if (dm && !M->d_blk_halo) a.push = 0;             // block-local unpack needs the per-block halo ranges
  a.b = b; a.x = x; a.V = V; a.W = W; a.xs = xs; a.dinv = dinv; a.g = g; a.N = N; a.m = m;
  const int grid = persist_grid(M, n, resident);
  a.R.chunk = run_chunk(M->nb, grid);
  void *args[] = {&a};
  PSP_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kThreads), args, smem_bytes, c.stream));
  c.launches++;
  return PSP_OK;
}

}  // namespace psp
