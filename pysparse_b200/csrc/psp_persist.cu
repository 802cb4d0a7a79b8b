// psp_persist.cu -- single-launch PCG: the whole solve in one persistent
// cooperative kernel.
//
// Replaces the per-iteration launch sequence of psp_solvers.cu for the most
// common case (native matrix, identity or single-step Jacobi preconditioner):
// one launch per solve, the iteration loop, the reductions and the convergence
// test (pcg.c:91-165) all run on the device; the host is not involved until
// the kernel ends.  Per iteration the grid passes three barriers (after the
// p update, after the SpMV partial sums, after the x/r update partial sums)
// instead of three kernel boundaries; every CTA reduces the per-CTA partials
// redundantly in the same order, so all CTAs -- and, multi-GPU, all ranks --
// take bit-identical decisions.  The matrix slices of the NEXT SpMV pass are
// prefetched into the TMA ring while the vector phases run.
//
// Multi-GPU: the halo of p is pulled from the neighbours' heaps between the p
// update and the SpMV (publish / poll as in psp_dist.cu) and the scalar
// all-reduces go through the peer mailboxes from inside the kernel.
//
// The scalar logic is NOT duplicated: thread 0 of every CTA runs the same
// run_epilogue() cases as the multi-kernel path on a shared-memory copy of the
// solver state.
#include <cooperative_groups.h>
#include "psp_mat.cuh"
#include "psp_epilogue.cuh"
#include "psp_tma_ring.cuh"
#include "psp_dist.cuh"

namespace cg = cooperative_groups;

namespace psp {

constexpr int kPersistStages = 2;

struct PcgArgs {
  MatView A;
  int n;                        // local rows
  const double *b;
  double *x, *r, *p, *q;
  const double *dinv;           // nullptr: identity preconditioner
  State *S;
  double *partials;             // 2 x (kRedSlots + 1) x kMaxGrid, double-buffered
  int l2_stream;
  int dist_debug;               // timing diagnostics (see Ctx::dist_debug)
  // multi-GPU
  const int *halo_owner, *halo_ridx;
  int nhalo;
  unsigned nbmask;
  // push model (send map present)
  int push, nsend;
  const int *send_row, *send_rank, *send_off;
  unsigned sendmask;
};

struct PersistCtx {
  State *sS;                    // shared-memory copy of the solver state
  double *tot;                  // shared: kRedSlots + 1 totals
  double *red_sh;               // shared: 8 doubles
  double *partials;
  DistDev *dd;
  unsigned long long seq, pubs, pulls;
  int parity;
  int debug;
};

// Grid-wide deterministic reduction; on return tot[0..NS) (and tot[kRedSlots] for the
// max) are valid in every thread of every CTA.
template <int NS, bool HASMAX>
__device__ __forceinline__ void grid_allreduce(cg::grid_group &grid, PersistCtx &c, const double *acc, double mx) {
  double *buf = c.partials + (size_t)c.parity * (kRedSlots + 1) * kMaxGrid;
  c.parity ^= 1;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    double v = block_sum(acc[s], c.red_sh);
    if (threadIdx.x == 0) buf[s * kMaxGrid + blockIdx.x] = v;
  }
  if (HASMAX) {
    double v = block_max(mx, c.red_sh);
    if (threadIdx.x == 0) buf[kRedSlots * kMaxGrid + blockIdx.x] = v;
  }
  grid.sync();
  const volatile double *vp = buf;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    double a = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a += vp[s * kMaxGrid + i];
    a = block_sum(a, c.red_sh);
    if (threadIdx.x == 0) c.tot[s] = a;
  }
  if (HASMAX) {
    double a = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a = fmax(a, vp[kRedSlots * kMaxGrid + i]);
    a = block_max(a, c.red_sh);
    if (threadIdx.x == 0) c.tot[kRedSlots] = a;
  }
  __syncthreads();
  if (c.dd && c.debug < 2) {
    // cross-GPU: CTA 0 pushes this rank's totals to every peer's mailbox; every CTA
    // reads the W contributions from the LOCAL mailbox and adds them in rank order
    __shared__ double box[kMaxRanks][kRedSlots + 1];
    DistDev *dd = c.dd;
    const int W = dd->world, me = dd->rank;
    const unsigned long long seq = ++c.seq;
    const int slot = (int)(seq % kMailRing);
    if ((int)threadIdx.x < W) {
      const int w = threadIdx.x;
      if (blockIdx.x == 0) {
        DistMail *dst = &dd->ctl[w]->mail[slot][me];
#pragma unroll
        for (int s = 0; s < NS; s++) st_relaxed_sys_f64(&dst->val[s], c.tot[s]);
        if (HASMAX) st_relaxed_sys_f64(&dst->val[kRedSlots], c.tot[kRedSlots]);
        st_release_sys(&dst->seq, seq);
      }
      const DistMail *src = &dd->ctl[me]->mail[slot][w];
      while (ld_acquire_sys(&src->seq) != seq) { }
#pragma unroll
      for (int s = 0; s < NS; s++) box[w][s] = ld_relaxed_sys_f64(&src->val[s]);
      if (HASMAX) box[w][kRedSlots] = ld_relaxed_sys_f64(&src->val[kRedSlots]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < NS; s++) {
        double a = 0.0;
        for (int w = 0; w < W; w++) a += box[w][s];
        c.tot[s] = a;
      }
      if (HASMAX) {
        double a = 0.0;
        for (int w = 0; w < W; w++) a = fmax(a, box[w][kRedSlots]);
        c.tot[kRedSlots] = a;
      }
    }
    __syncthreads();
  }
}

// thread 0 of every CTA advances its copy of the state; all copies stay identical
template <int NS, bool HASMAX>
__device__ __forceinline__ void epilogue(PersistCtx &c, int id) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; s++) c.sS->red[s] = c.tot[s];
    if (HASMAX) c.sS->redmax = c.tot[kRedSlots];
    run_epilogue(id, c.sS, nullptr);
  }
  __syncthreads();
}

// after a grid barrier: make this rank's vector visible to the peers, fetch their halo
__device__ __forceinline__ void publish_and_pull(cg::grid_group &grid, PersistCtx &c, const PcgArgs &a, double *vec,
                                                 size_t slot_off) {
  DistDev *dd = c.dd;
  if (c.debug >= 1) { grid.sync(); return; }
  c.pubs += 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(&dd->ctl[dd->rank]->pub, c.pubs);
  }
  const unsigned long long want = c.pulls + 1;
  if (threadIdx.x == 0) {
    for (int w = 0; w < dd->world; w++)
      if ((a.nbmask >> w) & 1u)
        while (ld_acquire_sys(&dd->ctl[w]->pub) < want) { }
  }
  __syncthreads();
  const int stride = gridDim.x * blockDim.x;
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < a.nhalo; h += stride)
    vec[a.n + h] = ld_relaxed_sys_f64(dd->heap[a.halo_owner[h]] + slot_off + a.halo_ridx[h]);
  c.pulls += 1;
  grid.sync();
}

// Push model (needs the send map, psp_dist_set_sendmap): called after the grid barrier
// that completes `vec`.  Every thread stores its share of the boundary entries straight
// into the neighbours' heaps (posted NVLink stores, nobody waits for them); the CTA that
// finishes last raises the "arrived" counters in the neighbours' control blocks; then the
// CTA waits -- on LOCAL memory -- until its own neighbours' entries have arrived.  No
// remote loads, no second grid barrier.  The wait (halo_wait) is deferred to the first
// row block that reads a halo column, so the transfer overlaps the interior blocks.
__device__ __forceinline__ void halo_push(PersistCtx &c, const PcgArgs &a, const double *vec, size_t slot_off) {
  DistDev *dd = c.dd;
  const int me = dd->rank;
  c.pubs += 1;
  c.pulls += 1;
  if (c.debug >= 1) return;
  const int stride = gridDim.x * blockDim.x;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < a.nsend; s += stride)
    st_relaxed_sys_f64(dd->heap[a.send_rank[s]] + slot_off + a.send_off[s], vec[a.send_row[s]]);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned t = atomicAdd(&dd->push_ticket, 1u);
    if (t == gridDim.x - 1) {
      dd->push_ticket = 0u;
      __threadfence_system();
      for (int w = 0; w < dd->world; w++)
        if ((a.sendmask >> w) & 1u) st_release_sys(&dd->ctl[w]->arrived[me], c.pubs);
    }
  }
}

__device__ __forceinline__ void halo_wait(PersistCtx &c, const PcgArgs &a) {
  DistDev *dd = c.dd;
  if (c.debug >= 1) return;
  if (threadIdx.x == 0) {
    for (int w = 0; w < dd->world; w++)
      if ((a.nbmask >> w) & 1u)
        while (ld_acquire_sys(&dd->ctl[dd->rank]->arrived[w]) < c.pulls) { }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 4) k_pcg_persist(PcgArgs a, size_t p_slot_off) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[kPersistStages];
  __shared__ double red_sh[8];
  __shared__ double tot[kRedSlots + 1];
  __shared__ State sS;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
  const long n = a.n;
  const long gstride = (long)gridDim.x * blockDim.x;
  const long g0 = (long)blockIdx.x * blockDim.x + tid;

  if (tid == 0) sS = *a.S;
  __syncthreads();
  PersistCtx c;
  c.sS = &sS; c.tot = tot; c.red_sh = red_sh; c.partials = a.partials; c.parity = 0;
  c.dd = sS.dist;
  c.debug = a.dist_debug;
  c.seq = c.pubs = c.pulls = 0;
  if (c.dd) { c.seq = c.dd->red_seq; c.pubs = c.dd->pubs_done; c.pulls = c.dd->pulls_done; }

  Ring<kPersistStages> ring;
  ring.init(smem, full, a.l2_stream != 0);
  ring.prefetch(a.A);

  // ---- ||b||  (pcg.c:57-67)
  {
    double acc[1] = {0.0};
    for (long i = g0; i < n; i += gstride) { double t = a.b[i]; acc[0] += t * t; }
    grid_allreduce<1, false>(grid, c, acc, 0.0);
    epilogue<1, false>(c, EPI_PCG_NORMB);
  }
  // ---- r = b - A x0  (pcg.c:72-89)
  const double *xin = a.x;
  if (c.dd) {
    for (long i = g0; i < n; i += gstride) a.p[i] = a.x[i];
    grid.sync();
    if (a.push) halo_push(c, a, a.p, p_slot_off);
    else publish_and_pull(grid, c, a, a.p, p_slot_off);
    xin = a.p;
  }
  ring.template pass<false>(a.A, xin, red_sh, [&](int r, double sum) { a.r[r] = sum; },
                            [&] { if (a.push) halo_wait(c, a); });
  ring.prefetch(a.A);
  grid.sync();
  {
    double acc[2] = {0.0, 0.0};
    const bool zero_rhs = (sS.phase == 1);
    for (long i = g0; i < n; i += gstride) {
      if (zero_rhs) { a.x[i] = 0.0; continue; }
      double t = a.b[i] - a.r[i];
      a.r[i] = t;
      acc[0] += t * t;
      double z = a.dinv ? t * a.dinv[i] : t;
      acc[1] += t * z;
    }
    grid_allreduce<2, false>(grid, c, acc, 0.0);
    epilogue<2, false>(c, EPI_PCG_RESID0_J);
  }

  // ---- iterations  (pcg.c:91-165)
  while (!sS.done) {
    {                                   // p = z + beta p   (first: p = z)
      const double beta = sS.v[V_BETA];
      const bool first = (sS.it == 1);
      for (long i = g0; i < n; i += gstride) {
        double ri = a.r[i];
        double zi = a.dinv ? ri * a.dinv[i] : ri;
        a.p[i] = first ? zi : zi + beta * a.p[i];
      }
    }
    grid.sync();
    if (c.dd) {
      if (a.push) halo_push(c, a, a.p, p_slot_off);
      else publish_and_pull(grid, c, a, a.p, p_slot_off);
    }
    {                                   // q = A p ; pq = p.q
      double acc[1] = {0.0};
      ring.template pass<false>(a.A, a.p, red_sh, [&](int r, double sum) { a.q[r] = sum; acc[0] += a.p[r] * sum; },
                                [&] { if (a.push) halo_wait(c, a); });
      ring.prefetch(a.A);
      grid_allreduce<1, false>(grid, c, acc, 0.0);
      epilogue<1, false>(c, EPI_PCG_PQ);
    }
    if (sS.done) break;
    {                                   // stagnation scan, x += alpha p, r -= alpha q, |r|^2, r.K^-1 r
      const double alpha = sS.v[V_ALPHA];
      const bool scan = (sS.stag == 0);
      double acc[2] = {0.0, 0.0}, mx = 0.0;
      for (long i = g0; i < n; i += gstride) {
        double xi = a.x[i], pi = a.p[i];
        if (scan) {
          if (xi != 0.0) mx = fmax(mx, fabs(alpha * pi / xi));
          else if (pi != 0.0) mx = fmax(mx, 1.0);
        }
        a.x[i] = xi + alpha * pi;
        double t = a.r[i] + (-alpha) * a.q[i];
        a.r[i] = t;
        acc[0] += t * t;
        double z = a.dinv ? t * a.dinv[i] : t;
        acc[1] += t * z;
      }
      grid_allreduce<2, true>(grid, c, acc, mx);
      epilogue<2, true>(c, EPI_PCG_UPDATE_J);
    }
  }

  ring.drain(a.A);
  if (blockIdx.x == 0 && tid == 0) {
    *a.S = sS;
    if (c.dd) { c.dd->red_seq = c.seq; c.dd->pubs_done = c.pubs; c.dd->pulls_done = c.pulls; }
  }
}

static double *g_partials2 = nullptr;

// Host entry: returns PSP_OK after ENQUEUEING the solve (one cooperative launch).
int pcg_persist_launch(const psp_mat *M, int n, const double *b, double *x, double *r, double *p, double *q,
                       const double *dinv, DistMat *dm) {
  Ctx &c = ctx();
  static int resident = 0;
  const int smem_bytes = kPersistStages * kStageBytes;
  if (resident == 0) {
    PSP_CUDA(cudaFuncSetAttribute(k_pcg_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist, kThreads, smem_bytes) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    resident = per_sm * c.sms;
    if (resident > kMaxGrid) resident = kMaxGrid;
  }
  if (!g_partials2) PSP_CUDA(cudaMalloc((void **)&g_partials2, sizeof(double) * 2 * (kRedSlots + 1) * kMaxGrid));
  PcgArgs a;
  a.A.ind = M->d_ind; a.A.col = M->d_col; a.A.val = M->d_val; a.A.nb = M->nb;
  a.A.blk = reinterpret_cast<const int4 *>(M->d_blk);
  a.A.nint = M->nint;
  a.n = n; a.b = b; a.x = x; a.r = r; a.p = p; a.q = q; a.dinv = dinv;
  a.S = c.state; a.partials = g_partials2; a.l2_stream = l2_stream_for(n); a.dist_debug = (int)c.dist_debug;
  a.halo_owner = dm ? dm->d_owner : nullptr; a.halo_ridx = dm ? dm->d_ridx : nullptr;
  a.nhalo = dm ? dm->nhalo : 0; a.nbmask = dm ? dm->nbmask : 0u;
  a.push = (dm && dm->has_sendmap && c.dist_push) ? 1 : 0;
  a.nsend = dm ? dm->nsend : 0; a.sendmask = dm ? dm->sendmask : 0u;
  a.send_row = dm ? dm->d_send_row : nullptr; a.send_rank = dm ? dm->d_send_rank : nullptr;
  a.send_off = dm ? dm->d_send_off : nullptr;
  size_t p_slot_off = dm ? (size_t)(p - dist().heap) : 0;
  // enough CTAs for the vector phases and the row blocks, never more than are co-resident
  long long want = (n + kThreads - 1) / kThreads;
  if (M->nb > want) want = M->nb;
  int grid = (int)(want < resident ? (want > 0 ? want : 1) : resident);
  void *args[] = {&a, &p_slot_off};
  PSP_CUDA(cudaLaunchCooperativeKernel((const void *)k_pcg_persist, dim3(grid), dim3(kThreads), args, smem_bytes, c.stream));
  c.launches++;
  return PSP_OK;
}

}  // namespace psp
