// psp_persist.cu -- single-launch PCG: the whole solve in one persistent
// cooperative kernel.
//
// Replaces the per-iteration launch sequence of psp_solvers.cu for the most
// common case (native matrix, identity or single-step Jacobi preconditioner):
// one launch per solve, the iteration loop, the reductions and the convergence
// test (pcg.c:91-165) all run on the device; the host is not involved until
// the kernel ends.  Per iteration the grid synchronises three times (after the
// p update, with the SpMV partial sums, with the x/r update partial sums)
// instead of crossing three kernel boundaries.  The two reductions ARE the
// barriers: every CTA deposits its partial sums and draws a ticket, the CTA
// with the last ticket adds the partials in index order, publishes the totals
// and releases the others -- all CTAs, and multi-GPU all ranks, see the same
// bits and take identical decisions.  The matrix slices of the NEXT SpMV pass
// are prefetched into the TMA ring while the vector phases run.
//
// Multi-GPU (row slabs, one process per GPU): the boundary entries of the new
// search direction are pushed into the neighbours' landing zones as
// self-validating words at the start of the p update and unpacked after the
// barrier, the row blocks that read halo columns run last (push model; the
// pull model of psp_dist.cu is the fallback without a send map); the scalar
// all-reduces go through the peer mailboxes from inside the reduction, one
// NVLink trip each.  No system-scope fences, no collective library.
//
// The scalar logic is NOT duplicated: thread 0 of every CTA runs the same
// run_epilogue() cases as the multi-kernel path on a shared-memory copy of the
// solver state.
#include "psp_persist.cuh"

namespace psp {

struct PcgArgs : PersistArgs {
  const double *b;
  double *x, *r, *p, *p2, *q;    // p, p2: the two buffers of the search direction
  const double *dinv;           // nullptr: identity preconditioner
};

template <int FMT>
__global__ void __launch_bounds__(kThreads, persist_ctas(FMT)) k_pcg_persist(PcgArgs a, size_t p_slot_off) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[FMT == 0 ? kPersistStages : 1];
  __shared__ __align__(16) unsigned char sseg_raw[FMT == 1 ? sizeof(RunSmem) : 16];   // (format 0 stages its descriptors in the TMA ring)
  RunSmem *sseg = reinterpret_cast<RunSmem *>(sseg_raw);
  __shared__ double red_sh[8 * (kRedSlots + 1)];
  __shared__ double tot[kRedSlots + 1];
  __shared__ State sS;
  cg::grid_group grid = cg::this_grid();
  constexpr int VU = vec_unroll(FMT);      // elements per thread and trip in the vector phases
  const int tid = threadIdx.x;
  // 32-bit indices: n < 2^31 and gridDim * blockDim < 2^20, so i + gstride cannot wrap
  const unsigned n = (unsigned)a.n;
  const unsigned gstride = gridDim.x * blockDim.x;
  const unsigned g0 = blockIdx.x * blockDim.x + tid;

  if (tid == 0) sS = *a.S;
  __syncthreads();
  PersistCtx c;
  c.sS = &sS; c.tot = tot; c.red_sh = red_sh; c.gen = 0; c.unpack_target = 0;
  c.dd = a.dd;
  c.debug = a.dist_debug;
  c.seq = c.pubs = 0;
  if (c.dd) { c.seq = (unsigned)c.dd->red_seq; c.pubs = (unsigned)c.dd->pubs_done; }
  const bool push = c.dd && a.push;
  const bool do_push = push && c.debug < 1;

  MatPass<FMT> ring;
  ring.init(smem, full, a.l2_stream != 0, sseg);
  ring.prefetch(a);

  // ---- ||b||  (pcg.c:57-67)
  {
    double acc[1] = {0.0};
    for (unsigned i = g0; i < n; i += gstride) { double t = a.b[i]; acc[0] += t * t; }
    grid_allreduce<1, false>(c, a, acc, 0.0);
    epilogue<1, false>(c, EPI_PCG_NORMB);
  }
  // the two buffers of the search direction (sharded: two slots of the symmetric heap)
  // (only the push model needs two: without it the update is done in place)
  double *p_cur = a.p, *p_new = push ? a.p2 : a.p;
  // ---- r = b - A x0  (pcg.c:72-89)
  const double *xin = a.x;
  if (c.dd) {
    if (do_push)
      for (unsigned s = g0; s < (unsigned)a.nsend; s += gstride) halo_send(c, a, s, a.x[a.send_row[s]]);
    for (unsigned i = g0; i < n; i += gstride) p_cur[i] = a.x[i];
    grid.sync();
#ifdef PSP_BLOCK_UNPACK
    if (push) c.pubs += 1;
#else
    if (push) halo_unpack(c, a, p_cur);
#endif
    else publish_and_pull(grid, c, a, a.p, p_slot_off);
    xin = p_cur;
  }
#ifdef PSP_BLOCK_UNPACK
  ring.template pass<true>(a, xin, red_sh, [&](int r, double sum) { a.r[r] = sum; },
                                  [&](int b) { if (push) halo_unpack_block(c, a, p_cur, b); });
#else
  ring.template pass<false>(a, xin, red_sh, [&](int r, double sum) { a.r[r] = sum; },
                            [&](int) { if (push) halo_wait(c, a); });
#endif
  ring.prefetch(a);
  grid.sync();
  {
    double acc[2] = {0.0, 0.0};
    const bool zero_rhs = (sS.phase == 1);
    for (unsigned i = g0; i < n; i += gstride) {
      if (zero_rhs) { a.x[i] = 0.0; continue; }
      double t = a.b[i] - a.r[i];
      a.r[i] = t;
      acc[0] += t * t;
      double z = a.dinv ? t * a.dinv[i] : t;
      acc[1] += t * z;
    }
    grid_allreduce<2, false>(c, a, acc, 0.0);
    epilogue<2, false>(c, EPI_PCG_RESID0_J);
  }

  // ---- iterations  (pcg.c:91-165)
#ifdef PSP_PERSIST_PROF
  for (int k = 0; k < 16; k++) c.prof[k] = 0;
  c.t_last = clock64();
#endif
  while (!sS.done) {
    {                                   // p_new = z + beta p_cur   (first: p_new = z)
      const double beta = sS.v[V_BETA];
      const bool first = (sS.it == 1);
      if (do_push) {
        // boundary entries first, straight to the neighbours (same arithmetic as below)
        for (unsigned s = g0; s < (unsigned)a.nsend; s += gstride) {
          const int row = a.send_row[s];
          const double ri = a.r[row];
          const double zi = a.dinv ? ri * a.dinv[row] : ri;
          halo_send(c, a, s, first ? zi : zi + beta * p_cur[row]);
        }
      }
      for (unsigned i0 = g0; i0 < n; i0 += VU * gstride) {
        double ri[VU], di[VU], pi[VU];
#pragma unroll
        for (int u = 0; u < VU; u++) {
          const unsigned i = i0 + u * gstride;
          const bool ok = i < n;
          ri[u] = ok ? a.r[i] : 0.0;
          di[u] = (ok && a.dinv) ? a.dinv[i] : 1.0;
          pi[u] = (ok && !first) ? p_cur[i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < VU; u++) {
          const unsigned i = i0 + u * gstride;
          if (i < n) {
            const double zi = a.dinv ? ri[u] * di[u] : ri[u];
            p_new[i] = first ? zi : zi + beta * pi[u];
          }
        }
      }
    }
    PROF_T(c, 0);
    grid.sync();
    PROF_T(c, 1);
    if (c.dd) {
#ifdef PSP_BLOCK_UNPACK
      if (push) c.pubs += 1;
#else
      if (push) halo_unpack(c, a, p_new);
#endif
      else publish_and_pull(grid, c, a, a.p, p_slot_off);    // pull model: p is updated in place
    }
    PROF_T(c, 2);
    {                                   // q = A p ; pq = p.q
      double acc[1] = {0.0};
#ifdef PSP_BLOCK_UNPACK
      ring.template pass<true>(a, p_new, red_sh, [&](int r, double sum) { a.q[r] = sum; acc[0] += p_new[r] * sum; },
                                      [&](int b) { if (push) halo_unpack_block(c, a, p_new, b); });
#else
      ring.template pass<false>(a, p_new, red_sh, [&](int r, double sum) { a.q[r] = sum; acc[0] += p_new[r] * sum; },
                                [&](int) { if (push) { PROF_T(c, 3); halo_wait(c, a); PROF_T(c, 13); } });
#endif
      ring.prefetch(a);
      PROF_T(c, 3);
      grid_allreduce<1, false>(c, a, acc, 0.0);
      PROF_T(c, 4);
      epilogue<1, false>(c, EPI_PCG_PQ);
      PROF_T(c, 5);
    }
    if (sS.done) break;
    {                                   // stagnation scan, x += alpha p, r -= alpha q, |r|^2, r.K^-1 r
      const double alpha = sS.v[V_ALPHA];
      const bool scan = (sS.stag == 0);
      double acc[2] = {0.0, 0.0}, mx = 0.0;
      for (unsigned i0 = g0; i0 < n; i0 += VU * gstride) {
        double xv[VU], pv[VU], rv[VU], qv[VU], dv[VU];
#pragma unroll
        for (int u = 0; u < VU; u++) {
          const unsigned i = i0 + u * gstride;
          const bool ok = i < n;
          xv[u] = ok ? a.x[i] : 0.0;
          pv[u] = ok ? p_new[i] : 0.0;
          rv[u] = ok ? a.r[i] : 0.0;
          qv[u] = ok ? a.q[i] : 0.0;
          dv[u] = (ok && a.dinv) ? a.dinv[i] : 1.0;
        }
#pragma unroll
        for (int u = 0; u < VU; u++) {
          const unsigned i = i0 + u * gstride;
          if (i < n) {
            const double xi = xv[u], pi = pv[u];
            if (scan) {
              if (xi != 0.0) mx = fmax(mx, fabs(alpha * pi / xi));
              else if (pi != 0.0) mx = fmax(mx, 1.0);
            }
            a.x[i] = xi + alpha * pi;
            const double t = rv[u] + (-alpha) * qv[u];
            a.r[i] = t;
            acc[0] += t * t;
            const double z = a.dinv ? t * dv[u] : t;
            acc[1] += t * z;
          }
        }
      }
      PROF_T(c, 6);
      grid_allreduce<2, true>(c, a, acc, mx);
      PROF_T(c, 4);
      epilogue<2, true>(c, EPI_PCG_UPDATE_J);
      PROF_T(c, 7);
    }
    { double *t = p_cur; p_cur = p_new; p_new = t; }
  }
#ifdef PSP_PERSIST_PROF
  if (blockIdx.x == 0 && tid == 0 && sS.it > 100) {
    const double k = 1.0 / (double)sS.it;
    printf("persist prof (cycles/iter, it=%d): p-update %.0f | sync %.0f | halo unpack %.0f | pass %.0f | "
           "allreduce[partials %.0f rest %.0f] | epi pq %.0f | update %.0f | epi upd %.0f | halo fence %.0f wait %.0f\n",
           sS.it, c.prof[0] * k, c.prof[1] * k, c.prof[2] * k, c.prof[3] * k, c.prof[8] * k, (c.prof[9] + c.prof[4]) * k,
           c.prof[5] * k, c.prof[6] * k, c.prof[7] * k, c.prof[11] * k, c.prof[13] * k);
  }
#endif

  ring.drain(a);
  if (blockIdx.x == 0 && tid == 0) {
    *a.S = sS;
    if (c.dd) {                       // 64-bit sequence numbers: add what this launch did (wrap-safe)
      const unsigned nred = c.seq - (unsigned)c.dd->red_seq, nexch = c.pubs - (unsigned)c.dd->pubs_done;
      c.dd->red_seq += nred; c.dd->pubs_done += nexch; c.dd->pulls_done += nexch;
    }
  }
}


int persist_prepare(PersistArgs &a, const psp_mat *M, int n, DistMat *dm) {
  Ctx &c = ctx();
  if (!c.persist_partials) PSP_CUDA(cudaMalloc((void **)&c.persist_partials, sizeof(double) * 2 * (kRedSlots + 1) * kMaxGrid));
  if (!c.persist_sync) PSP_CUDA(cudaMalloc(&c.persist_sync, sizeof(PersistSync)));
  PSP_CUDA(cudaMemsetAsync(c.persist_sync, 0, sizeof(PersistSync), c.stream));
  a.A.ind = M->d_ind; a.A.col = M->d_col; a.A.val = M->d_val; a.A.nb = M->nb;
  a.A.blk = reinterpret_cast<const int4 *>(M->d_blk);
  a.A.nint = M->nint;
  a.A.blk_halo = reinterpret_cast<const int2 *>(M->d_blk_halo);
  a.R.val = M->d_rval; a.R.seg = reinterpret_cast<const int4 *>(M->d_rseg);
  a.R.blk = reinterpret_cast<const int4 *>(M->d_rblk); a.R.nb = M->nb; a.R.nint = M->nint;
  a.R.blk_halo = reinterpret_cast<const int2 *>(M->d_blk_halo);
  a.n = n;
  a.S = c.state; a.partials = c.persist_partials; a.sync = static_cast<PersistSync *>(c.persist_sync); a.dd = dm ? dist().dev : nullptr;
  a.l2_stream = l2_stream_for(n); a.dist_debug = (int)c.dist_debug;
  a.halo_owner = dm ? dm->d_owner : nullptr; a.halo_ridx = dm ? dm->d_ridx : nullptr;
  a.nhalo = dm ? dm->nhalo : 0; a.nbmask = dm ? dm->nbmask : 0u;
  a.push = (dm && dm->has_sendmap && c.dist_push) ? 1 : 0;
  a.nsend = dm ? dm->nsend : 0;
  a.send_row = dm ? dm->d_send_row : nullptr; a.send_dst = dm ? dm->d_send_dst : nullptr;
  return PSP_OK;
}

int persist_grid(const psp_mat *M, int n, int resident) {
  long long want = (n + kThreads - 1) / kThreads;
  if (M->nb > want) want = M->nb;
  return (int)(want < resident ? (want > 0 ? want : 1) : resident);
}

// Host entry: returns PSP_OK after ENQUEUEING the solve (one cooperative launch).
int pcg_persist_launch(const psp_mat *M, int n, const double *b, double *x, double *r, double *p, double *p2,
                       double *q, const double *dinv, DistMat *dm) {
  Ctx &c = ctx();
  const int fmt = M->fmt;
  const void *kern = fmt == 1 ? (const void *)k_pcg_persist<1> : (const void *)k_pcg_persist<0>;
  const int smem_bytes = fmt == 1 ? 0 : kPersistStages * kStageBytes;
  int &resident = c.occ_pcg[fmt];
  if (resident == 0) {
    if (smem_bytes) PSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem_bytes) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (per_sm > persist_ctas(fmt)) per_sm = persist_ctas(fmt);
    resident = per_sm * c.sms;
    if (resident > kMaxGrid) resident = kMaxGrid;
  }
  PcgArgs a;
  PSP_TRY(persist_prepare(a, M, n, dm));
  a.b = b; a.x = x; a.r = r; a.p = p; a.p2 = p2; a.q = q; a.dinv = dinv;
  size_t p_slot_off = dm ? (size_t)(p - dist().heap) : 0;
  const int grid = persist_grid(M, n, resident);
  a.R.chunk = run_chunk(M->nb, grid);
  void *args[] = {&a, &p_slot_off};
  PSP_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kThreads), args, smem_bytes, c.stream));
  c.launches++;
  return PSP_OK;
}

}  // namespace psp
