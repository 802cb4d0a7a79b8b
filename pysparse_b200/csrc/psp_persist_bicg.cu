// psp_persist_bicg.cu -- single-launch BiCGSTAB: the whole solve in one persistent
// cooperative kernel (same infrastructure as the PCG kernel, psp_persist.cuh).
//
// STATUS: written at the end of round 1, compiled for sm_100a but NOT yet run on a GPU.
// It is OFF by default (option bicg_persist = 0; solve_bicgstab keeps the validated
// per-phase kernel sequence) and is to be validated against the oracle and the per-phase
// path before the default changes.
//
// One iteration of right-preconditioned BiCGSTAB (bicgstab.c:244-293) with a native matrix
// and an identity / 1-step Jacobi preconditioner:
//   P   p = r + beta (p - omega v), phat = K^-1 p          -> halo exchange of phat
//       v = A phat, rhat.v            (all-reduce)  alpha = rho / rhat.v
//   S   s = r - alpha v, shat = K^-1 s                     -> halo exchange of shat
//       t = A shat, s.t, t.t          (all-reduce)  omega = s.t / t.t
//   U   x += alpha phat + omega shat, r = s - omega t, |r|^2, rhat.r   (all-reduce)
//       convergence test, rho, beta
// i.e. two SpMV passes, three all-reduces (each also the grid barrier) and two plain grid
// barriers per iteration instead of five kernel boundaries.  The scalar work runs in the
// same run_epilogue() cases as the per-phase path.  Multi-GPU: as in the PCG kernel the
// boundary entries of phat / shat are computed and sent (self-validating words) at the
// start of their vector phase; p is double-buffered because the sender recomputes its
// entries from the OLD p while the owner thread overwrites it.
#include "psp_persist.cuh"

namespace psp {

struct BicgArgs : PersistArgs {
  const double *b;
  double *x, *r, *rhat, *p, *p2, *phat, *v, *s, *shat, *t;   // phat == p / shat == s when K is the identity
  const double *dinv;                                         // nullptr: identity preconditioner
};

template <int FMT>
__global__ void __launch_bounds__(kThreads, persist_ctas(FMT))
k_bicgstab_persist(BicgArgs a, size_t p_slot_off, size_t phat_slot_off, size_t shat_slot_off) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[FMT == 0 ? kPersistStages : 1];
  __shared__ __align__(16) unsigned char sseg_raw[FMT == 1 ? sizeof(RunSmem) : 16];   // (format 0 stages its descriptors in the TMA ring)
  RunSmem *sseg = reinterpret_cast<RunSmem *>(sseg_raw);
  __shared__ double red_sh[8 * (kRedSlots + 1)];
  __shared__ double tot[kRedSlots + 1];
  __shared__ State sS;
  cg::grid_group grid = cg::this_grid();
  const int tid = threadIdx.x;
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
  const bool ident = (a.dinv == nullptr);

  MatPass<FMT> ring;
  ring.init(smem, full, a.l2_stream != 0, sseg);
  ring.prefetch(a);

  // ---- ||b||  (bicgstab.c:203-216)
  {
    double acc[1] = {0.0};
    for (unsigned i = g0; i < n; i += gstride) { double t = a.b[i]; acc[0] += t * t; }
    grid_allreduce<1, false>(c, a, acc, 0.0);
    epilogue<1, false>(c, EPI_BICG_NORMB);
  }
  // p is double-buffered in the push model only (otherwise updated in place)
  double *p_cur = a.p, *p_new = push ? a.p2 : a.p;

  // ---- r = A x0  (bicgstab.c:221-233): sharded, x0 travels in the p buffer
  const double *xin = a.x;
  if (c.dd) {
    if (do_push)
      for (unsigned s = g0; s < (unsigned)a.nsend; s += gstride) halo_send(c, a, s, a.x[a.send_row[s]]);
    for (unsigned i = g0; i < n; i += gstride) p_cur[i] = a.x[i];
    grid.sync();
    if (push) halo_unpack(c, a, p_cur);
    else publish_and_pull(grid, c, a, a.p, p_slot_off);
    xin = p_cur;
  }
  ring.template pass<false>(a, xin, red_sh, [&](int r, double sum) { a.r[r] = sum; },
                            [&](int) { if (push) halo_wait(c, a); });
  ring.prefetch(a);
  grid.sync();
  // ---- r = b - r, rhat = r, |r|^2   (bicgstab.c:235-248)
  {
    double acc[1] = {0.0};
    const bool zero_rhs = (sS.phase == 1);
    for (unsigned i = g0; i < n; i += gstride) {
      if (zero_rhs) { a.x[i] = 0.0; continue; }
      double t = a.b[i] + -a.r[i];
      a.r[i] = t; a.rhat[i] = t;
      acc[0] += t * t;
    }
    grid_allreduce<1, false>(c, a, acc, 0.0);
    epilogue<1, false>(c, EPI_BICG_RES0);
  }

  // ---- iterations  (bicgstab.c:244-293)
  while (!sS.done) {
    // the vectors the two products read: K^-1 p and K^-1 s (p and s themselves if K = I)
    double *x1 = ident ? p_new : a.phat;
    double *x2 = ident ? a.s : a.shat;
    {                                   // P: p = r + beta (p - omega v)   (first: p = r) ; phat = K^-1 p
      const double beta = sS.v[B_BETA], omega = sS.v[B_OMEGA];
      const bool first = (sS.it == 1);
      if (do_push)
        for (unsigned s = g0; s < (unsigned)a.nsend; s += gstride) {
          const int row = a.send_row[s];
          const double pn = first ? a.r[row] : a.r[row] + beta * (p_cur[row] - omega * a.v[row]);
          halo_send(c, a, s, ident ? pn : pn * a.dinv[row]);
        }
      for (unsigned i = g0; i < n; i += gstride) {
        const double pn = first ? a.r[i] : a.r[i] + beta * (p_cur[i] - omega * a.v[i]);
        p_new[i] = pn;
        if (!ident) a.phat[i] = pn * a.dinv[i];
      }
    }
    grid.sync();
    if (c.dd) {
      if (push) halo_unpack(c, a, x1);
      else publish_and_pull(grid, c, a, x1, ident ? p_slot_off : phat_slot_off);
    }
    {                                   // v = A phat ; rhat.v
      double acc[1] = {0.0};
      ring.template pass<false>(a, x1, red_sh, [&](int r, double sum) { a.v[r] = sum; acc[0] += a.rhat[r] * sum; },
                                [&](int) { if (push) halo_wait(c, a); });
      ring.prefetch(a);
      grid_allreduce<1, false>(c, a, acc, 0.0);
      epilogue<1, false>(c, EPI_BICG_ALPHA);
    }
    {                                   // S: s = r - alpha v ; shat = K^-1 s   (bicgstab.c:265)
      const double alpha = sS.v[B_ALPHA];
      if (do_push)
        for (unsigned s = g0; s < (unsigned)a.nsend; s += gstride) {
          const int row = a.send_row[s];
          const double sn = a.r[row] + (-alpha) * a.v[row];
          halo_send(c, a, s, ident ? sn : sn * a.dinv[row]);
        }
      for (unsigned i = g0; i < n; i += gstride) {
        const double sn = a.r[i] + (-alpha) * a.v[i];
        a.s[i] = sn;
        if (!ident) a.shat[i] = sn * a.dinv[i];
      }
    }
    grid.sync();
    if (c.dd) {
      if (push) halo_unpack(c, a, x2);
      else publish_and_pull(grid, c, a, x2, shat_slot_off);
    }
    {                                   // t = A shat ; s.t ; t.t
      double acc[2] = {0.0, 0.0};
      ring.template pass<false>(a, x2, red_sh,
                                [&](int r, double sum) { a.t[r] = sum; acc[0] += a.s[r] * sum; acc[1] += sum * sum; },
                                [&](int) { if (push) halo_wait(c, a); });
      ring.prefetch(a);
      grid_allreduce<2, false>(c, a, acc, 0.0);
      epilogue<2, false>(c, EPI_BICG_OMEGA);
    }
    {                                   // U: x += alpha phat + omega shat ; r = s - omega t ; |r|^2 ; rhat.r
      const double alpha = sS.v[B_ALPHA], omega = sS.v[B_OMEGA];
      double acc[2] = {0.0, 0.0};
      for (unsigned i = g0; i < n; i += gstride) {
        a.x[i] = a.x[i] + alpha * x1[i] + omega * x2[i];
        const double rn = a.s[i] - omega * a.t[i];
        a.r[i] = rn;
        acc[0] += rn * rn;
        acc[1] += a.rhat[i] * rn;
      }
      grid_allreduce<2, false>(c, a, acc, 0.0);
      epilogue<2, false>(c, EPI_BICG_RES);
    }
    { double *tmp = p_cur; p_cur = p_new; p_new = tmp; }
  }

  ring.drain(a);
  if (blockIdx.x == 0 && tid == 0) {
    *a.S = sS;
    if (c.dd) {                         // 64-bit sequence numbers: add what this launch did (wrap-safe)
      const unsigned nred = c.seq - (unsigned)c.dd->red_seq, nexch = c.pubs - (unsigned)c.dd->pubs_done;
      c.dd->red_seq += nred; c.dd->pubs_done += nexch; c.dd->pulls_done += nexch;
    }
  }
}

// Host entry: returns PSP_OK after ENQUEUEING the solve (one cooperative launch).
// p2: second buffer for p (nullptr: no room -- the push model is then not used).
int bicg_persist_launch(const psp_mat *M, int n, const double *b, double *x, double *r, double *rhat, double *p,
                        double *p2, double *phat, double *v, double *s, double *shat, double *t, const double *dinv,
                        DistMat *dm) {
  Ctx &c = ctx();
  const int fmt = M->fmt;
  const void *kern = fmt == 1 ? (const void *)k_bicgstab_persist<1> : (const void *)k_bicgstab_persist<0>;
  const int smem_bytes = fmt == 1 ? 0 : kPersistStages * kStageBytes;
  int &resident = c.occ_bicg[fmt];
  if (resident == 0) {
    if (smem_bytes) PSP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem_bytes) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (per_sm > persist_ctas(fmt)) per_sm = persist_ctas(fmt);
    resident = per_sm * c.sms;
    if (resident > kMaxGrid) resident = kMaxGrid;
  }
  BicgArgs a;
  PSP_TRY(persist_prepare(a, M, n, dm));
  if (!p2) a.push = 0;
  a.b = b; a.x = x; a.r = r; a.rhat = rhat; a.p = p; a.p2 = p2; a.phat = phat; a.v = v; a.s = s; a.shat = shat; a.t = t;
  a.dinv = dinv;
  size_t p_off = 0, phat_off = 0, shat_off = 0;
  if (dm) {
    p_off = (size_t)(p - dist().heap);
    phat_off = (size_t)(phat - dist().heap);
    shat_off = (size_t)(shat - dist().heap);
  }
  const int grid = persist_grid(M, n, resident);
  a.R.chunk = run_chunk(M->nb, grid);
  void *args[] = {&a, &p_off, &phat_off, &shat_off};
  PSP_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kThreads), args, smem_bytes, c.stream));
  c.launches++;
  return PSP_OK;
}

}  // namespace psp
