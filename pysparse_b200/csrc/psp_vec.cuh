// psp_vec.cuh -- the fused BLAS-1 kernel template.
//
// One persistent, grid-stride kernel per "phase" of a solver iteration: a
// functor F describes what happens to element i (any number of vector reads
// and writes), which sums / max it contributes to, and which scalar epilogue
// (psp_epilogue.cuh) runs once the grid-wide sums are known.  This is what
// replaces the reference's sequences of ddot / daxpy / dcopy / dnrm2 calls
// (pysparse/include/blas.h:97-126) with one pass over memory per phase.
//
// Multi-GPU: when State::dist is set, the sums are all-reduced across the
// ranks inside the same kernel (dist_allreduce, psp_common.cuh), and a kernel
// launched with kPublish announces -- after all its stores -- that the vector it
// produced may now be read by the peers (halo pull, psp_dist.cu).
#pragma once
#include "psp_common.cuh"
#include "psp_epilogue.cuh"

namespace psp {

// gate: 0 = skip when S->done; 1 = also skip when S->cyc_done (GMRES inner
// loop); 3 = never skip.
constexpr int kGateMask = 0xff;
constexpr int kPublish = 0x100;

__device__ __forceinline__ bool gated_out(const State *S, int gate) {
  gate &= kGateMask;
  if (gate == 3) return false;   // runs unconditionally (post-loop clean-up kernels)
  const volatile State *V = S;
  if (V->done) return true;
  if (gate == 1 && V->cyc_done) return true;
  return false;
}

// Announce to the peers that this rank's freshly written vector is complete.
// Called by thread 0 of the last block (all blocks' stores are ordered before
// the ticket it drew).
__device__ __forceinline__ void dist_publish(DistDev *dd) {
  __threadfence_system();
  dd->pubs_done += 1;
  st_release_sys(&dd->ctl[dd->rank]->pub, dd->pubs_done);
}

// F must provide:
//   static constexpr int NS        number of sums (0..kRedSlots)
//   static constexpr bool HASMAX   contributes to a max reduction
//   static constexpr bool TICKET   run the epilogue even though NS == 0
//   void prepare(const State*)     load the scalars it needs (uniform)
//   void elem(long i, double *acc, double &mx)
template <class F>
__global__ void __launch_bounds__(kThreads) k_map(F f, long n, State *S, double *partials,
                                                   unsigned *ticket, int epilogue, int gate) {
  if (gated_out(S, gate)) return;
  __shared__ double sh[8];
  f.prepare(S);
  constexpr int NA = F::NS > 0 ? F::NS : 1;
  double acc[NA];
#pragma unroll
  for (int s = 0; s < NA; s++) acc[s] = 0.0;
  double mx = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  // (two elements per trip with __restrict__ members was measured slower: 596 vs 579 us per
  // PCG iteration at 256^3 -- the extra registers cost a resident CTA)
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f.elem(i, acc, mx);
  const bool publish = (gate & kPublish) != 0;
  if (F::NS > 0 || F::HASMAX || F::TICKET || publish) {
    double mine[NA];
#pragma unroll
    for (int s = 0; s < F::NS; s++) mine[s] = block_sum(acc[s], sh);
    double mymax = 0.0;
    if constexpr (F::HASMAX) mymax = block_max(mx, sh);
    double tot[NA], totmax = 0.0;
    if (grid_reduce_finish<F::NS, F::HASMAX>(mine, mymax, partials, ticket, tot, &totmax, sh, S->dist)) {
      if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < F::NS; s++) S->red[s] = tot[s];
        if constexpr (F::HASMAX) S->redmax = totmax;
        run_epilogue(epilogue, S, (void *)(S + 1));
        if (publish && S->dist) dist_publish(S->dist);
      }
    }
  }
}

template <class F>
inline int launch_map(const F &f, long n, int epilogue = EPI_NONE, int gate = 0) {
  Ctx &c = ctx();
  // persistent grid = exactly the CTAs of this instantiation that are resident at once
  static int per_sm = 0;
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_map<F>, kThreads, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (per_sm > 8) per_sm = 8;
  }
  int grid = grid_for((size_t)(n > 0 ? n : 1), per_sm);
  k_map<F><<<grid, kThreads, 0, c.stream>>>(f, n, c.state, c.partials, c.tickets, epilogue, gate);
  PSP_LAUNCH_CHECK();
  return PSP_OK;
}

}  // namespace psp
