// psp_common.cuh -- shared internals of libpsp_b200.so (context, error handling,
// deterministic reductions, the device-side solver state).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "psp_b200.h"

namespace psp {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);

#define PSP_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      psp::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                     cudaGetErrorString(e_));                                       \
      return PSP_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

#define PSP_TRY(call)                 \
  do {                                \
    int s_ = (call);                  \
    if (s_ != PSP_OK) return s_;      \
  } while (0)

// ---------------------------------------------------------------- context
constexpr int kMaxGrid = 148 * 16;  // upper bound on persistent grid sizes
constexpr int kRedSlots = 8;        // sums a single kernel may reduce
constexpr int kThreads = 256;

// Device-side solver state.  One instance lives in device memory for the
// duration of a solve; every kernel of the solve reads `done` first and
// returns immediately when it is set, so iterations enqueued past the point
// of convergence cost a few microseconds and change nothing.
// ---- multi-GPU device-side structures (psp_dist.cu) ------------------------
// One process per GPU; every rank maps every peer's "symmetric heap" and
// control block through CUDA IPC, so that
//   * scalar all-reduces are done INSIDE the reduction kernels: the last block
//     pushes its partial sums into every peer's mailbox (NVLink stores) and
//     sums the W mailbox entries it receives, in rank order -- the same bits
//     on every rank, no NCCL launch, no host involvement;
//   * x-halo values are pulled straight out of the neighbours' vectors by a
//     small gather kernel (NVLink loads) once the owner has published them.
constexpr int kMaxRanks = 8;
constexpr int kMailRing = 4;

struct DistMail {                       // one all-reduce contribution
  double val[kRedSlots + 1];            // sums, then the max
  unsigned long long seq;               // written last (release); reduction sequence number
  // flag-in-data copy used by the persistent kernels: word 2k / 2k+1 = low / high half of
  // value k in the low 32 bits, the low 32 bits of the sequence number in the high ones --
  // every 8-byte word validates itself, so no release fence and a single NVLink trip
#ifdef PSP_LL_ALLREDUCE
  unsigned long long ll[2 * (kRedSlots + 1)];   // A/B build: the per-phase kernels use the words too
#else
  unsigned long long ll[8];
#endif
};

struct DistCtl {                        // lives in each rank's IPC-shared control block
  unsigned long long pub;               // number of vector publications completed by this rank
  unsigned long long pad[7];
  DistMail mail[kMailRing][kMaxRanks];  // mail[seq % ring][source rank]
};

struct DistDev {                        // device-resident descriptor (private to the rank)
  int rank, world;
  DistCtl *ctl[kMaxRanks];              // ctl[rank] is local, the others are peer mappings
  double *heap[kMaxRanks];              // symmetric heaps (same layout on every rank)
  unsigned long long red_seq;           // reductions performed so far (identical on all ranks)
  unsigned long long pubs_done;         // publications performed by this rank
  unsigned long long pulls_done;        // halo pulls performed by this rank
  // push-model halo exchange (persistent kernels): element e of a vector sent to this rank
  // lands in the two words 2e, 2e+1 at heap + ll_off -- (exchange number << 32) | low / high
  // half of the value; self-validating, so the sender needs no fence and no flag
  size_t ll_off;
};

struct State {
  DistDev *dist; // nullptr on a single GPU
  int done;      // 1: solve finished, remaining kernels are no-ops
  int info;      // solver flag / info as the reference reports it
  int iter;      // iteration count as the reference reports it
  int it;        // running iteration counter
  int maxit;
  int stag;      // PCG stagnation latch
  int inner;     // GMRES: index inside the restart cycle
  int phase;     // solver-specific phase latch
  int mdim;      // GMRES restart length
  int cyc_done;  // GMRES: inner loop of the current cycle finished
  int last_i;    // GMRES: last inner index of the cycle
  int pad0;
  double tol;
  double relres;
  double red[kRedSlots];   // raw results of the last reduction
  double redmax;           // raw result of the last max-reduction
  double v[48];            // solver scalars (alpha, beta, rho, ...)
};

struct Ctx {
  bool ready = false;
  int device = -1;
  int sms = 0;
  cudaStream_t stream = nullptr;
  double *partials = nullptr;       // [kRedSlots+1][kMaxGrid]
  unsigned *tickets = nullptr;      // [16] zero-initialised ticket counters
  State *state = nullptr;           // device solver state
  State *h_state = nullptr;         // pinned mirror
  double *stage = nullptr;          // pinned staging for strided / small copies
  size_t stage_elems = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evpoll[2] = {nullptr, nullptr};
  double last_ms = 0.0;
  long long launches = 0;
  int solve_depth = 0;              // > 0 while a solve is running (nested solves get their own frame)
  // device scratch kept between calls (b, x and the work vectors of psp_solve): a solve of a
  // small system costs a few kernel launches, three cudaMalloc/cudaFree pairs would dominate it
  double *pool[3] = {nullptr, nullptr, nullptr};
  size_t pool_cap[3] = {0, 0, 0};
  // options
  long long spmv_variant = 0;
  long long gpus = 1;         // > 1: matrices are also cut into row slabs over that many GPUs and pcg / bicgstab / gmres run sharded (PSP_GPUS)
  long long spmv_format = -1; // device layout chosen at matrix creation: -1 automatic, 0 CSR row blocks, 1 diagonal runs
  // per-device launch caches (occupancy, "attribute already set"): a context is one device, and
  // cudaFuncSetAttribute is per device -- nothing of this kind may live in a function-static
  int occ_tma[3][2] = {{0, 0}, {0, 0}, {0, 0}}, occ_tma_cap[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};   // k_spmv_tma<STAGES-2, DOTS>
  int occ_stream[2][2] = {{0, 0}, {0, 0}};                                          // k_spmv_stream<VEC, DOTS>
  int occ_sweep_grid[2] = {0, 0};
  size_t sweep_cta_smem[2] = {0, 0};
  bool sweep_cluster_ready[2] = {false, false};
  double *persist_partials = nullptr;           // scratch of the persistent kernels' all-reduce
  void *persist_sync = nullptr;
  int occ_runs[2] = {0, 0};   // resident CTAs of k_spmv_runs<false/true> on this device
  int occ_pcg[2] = {0, 0};    // resident CTAs of k_pcg_persist<format> on this device
  int occ_bicg[2] = {0, 0}, occ_gmres[2] = {0, 0};
  long long spmv_ctas = 0;    // cap on resident CTAs per SM for the TMA SpMV (0: occupancy limit)
  long long l2_stream = -1;   // L2 evict-first hint on the matrix stream: 1 on, 0 off, -1 automatic
  long long pcg_persist = -1; // single-launch PCG: 1 always, 0 never, -1 automatic
  long long ssor_engine = 0;  // SSOR sweeps: 0 automatic, 1 grid-wide barrier, 2 cluster barrier, 3 single CTA, 4 point-to-point (read when the preconditioner is created)
  long long ssor_cluster = 8; // CTAs in the sweep cluster (8 portable, 16 non-portable)
  // single-launch kernels of the other two sharded solvers: 1 always, 0 never, -1 automatic.
  // Measured on conv-diff 200^3 (it/s, per-phase vs single launch, 1/2/4/8 GPUs): BiCGSTAB
  // 1757/2819/4471/6680 vs 1771/3190/5765/10733 -> always; GMRES(30) 937/1567/2292/2638 vs
  // 762/1447/2771/3712 -> from 4 GPUs on (profiles/r2_multi_notes.md)
  long long gmres_persist = -1;
  long long bicg_persist = -1;
  long long dist_push = 1;    // sharded single-launch PCG: push the halo to the neighbours (needs the send map)
  long long dist_debug = 0;   // timing diagnostics only (results invalid): 1 skip halo pull, 2 also skip the cross-GPU all-reduce
  long long poll_chunk = 0;   // 0: automatic
  // sampled SpMV timing (bench.py roofline): event pairs around every Nth launch
  static constexpr int kSpmvPairs = 512;
  long long spmv_sample = 0, spmv_seen = 0;
  int spmv_pairs = 0;
  cudaEvent_t spmv_ev[2 * kSpmvPairs] = {};
  // general-purpose timing events on the library stream (psp_event_record)
  cudaEvent_t user_ev[8] = {};
};

Ctx &ctx();                          // the calling thread's context (a rank of the in-process multi-GPU layer has its own)
void bind_thread(Ctx *c, struct DistHost *d);   // psp_multi.cu: make a worker thread act as one rank
int init_ctx(Ctx &c, int device);    // create streams / scratch of a context on `device`
int ensure_init();                   // PSP_OK or error; lazily initialises device 0
int ensure_stage(size_t elems);      // pinned staging buffer of at least `elems` doubles
int pool_get(int slot, size_t elems, double **out);   // context-owned device scratch (grows, never shrinks below use)

// The matrix is streamed once per product, the vectors are re-used: when the vectors of
// the problem (<= 3 M rows: 6 vectors <= 144 MB) can live in the 126 MB L2, the TMA loads
// of the matrix carry an evict-first hint so that they do not push the vectors out.
inline int l2_stream_for(long rows) {
  Ctx &c = ctx();
  if (c.l2_stream >= 0) return c.l2_stream ? 1 : 0;
  return rows <= 3L * 1000 * 1000 ? 1 : 0;
}

inline int grid_for(size_t work_items, int per_sm = 8) {
  Ctx &c = ctx();
  long long g = (long long)((work_items + kThreads - 1) / kThreads);
  long long cap = (long long)c.sms * per_sm;
  if (cap > kMaxGrid) cap = kMaxGrid;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

#define PSP_LAUNCH_CHECK()                                                        \
  do {                                                                            \
    psp::ctx().launches++;                                                        \
    cudaError_t e_ = cudaGetLastError();                                          \
    if (e_ != cudaSuccess) {                                                      \
      psp::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,     \
                     cudaGetErrorString(e_));                                     \
      return PSP_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic block-wide sum for blockDim.x == kThreads (8 warps).  Result
// valid in thread 0.  `sh` must hold 8 doubles; contains a __syncthreads().
__device__ __forceinline__ double block_sum(double v, double *sh) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

__device__ __forceinline__ double block_max(double v, double *sh) {
  v = warp_max(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
    r = warp_max(r);
  }
  return r;
}

// Second stage of a grid-wide deterministic reduction.  Every block has put
// its NS partial sums (and optionally one partial max) into
// partials[s * kMaxGrid + blockIdx.x]; the block that draws the last ticket
// adds them up in a fixed order (independent of which block that is) and
// returns true in ALL its threads with totals[] valid in thread 0.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Cross-GPU part of a reduction, run by the LAST block of the reduction kernel
// on every rank: lanes 0..W-1 push this rank's totals into rank w's mailbox
// (remote NVLink stores, value first, sequence number last with release
// semantics), then wait for rank w's contribution in the local mailbox and the
// sums are formed in rank order, so every rank gets bit-identical results.
template <int NS, bool HASMAX>
__device__ __forceinline__ void dist_allreduce(DistDev *dd, double *totals, double *totmax, double *sh) {
  __shared__ double box[kMaxRanks][kRedSlots + 1];
  const int W = dd->world, me = dd->rank;
  const unsigned long long seq = dd->red_seq + 1;
  const int slot = (int)(seq % kMailRing);
  __syncthreads();
#ifdef PSP_LL_ALLREDUCE
  // A/B build (-DPSP_LL_ALLREDUCE, prepared for round 2, not yet run): the flag-in-data words of
  // the persistent kernels for the per-phase kernels too -- no st.release.sys (MEMBAR.SYS) per
  // reduction.  Thread (w, j) sends word j of this rank's totals to rank w and fetches rank w's.
  {
    constexpr int NV = NS + (HASMAX ? 1 : 0);
    __shared__ unsigned lbox[kMaxRanks][2 * (kRedSlots + 1)];
    const unsigned tag = (unsigned)seq;
    if ((int)threadIdx.x < W * 2 * NV) {
      const int w = threadIdx.x / (2 * NV), j = threadIdx.x % (2 * NV), k = j >> 1;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(k < NS ? totals[k] : *totmax);
      const unsigned half = (j & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&dd->ctl[w]->mail[slot][me].ll[j]),
                   "l"(((unsigned long long)tag << 32) | half) : "memory");
      const unsigned long long *src = &dd->ctl[me]->mail[slot][w].ll[j];
      unsigned long long got;
      do {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(src) : "memory");
      } while ((unsigned)(got >> 32) != tag);
      lbox[w][j] = (unsigned)got;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < NV; k++) {
        double a = 0.0;
        for (int w = 0; w < W; w++) {
          const double v = __longlong_as_double((long long)(((unsigned long long)lbox[w][2 * k + 1] << 32) | lbox[w][2 * k]));
          a = (k < NS) ? a + v : fmax(a, v);
        }
        if (k < NS) totals[k] = a; else *totmax = a;
      }
      dd->red_seq = seq;
    }
    (void)sh; (void)box;
    return;
  }
#endif
  if ((int)threadIdx.x < W) {
    const int w = threadIdx.x;
    DistMail *dst = &dd->ctl[w]->mail[slot][me];
#pragma unroll
    for (int s = 0; s < NS; s++) st_relaxed_sys_f64(&dst->val[s], totals[s]);
    if (HASMAX) st_relaxed_sys_f64(&dst->val[kRedSlots], *totmax);
    st_release_sys(&dst->seq, seq);
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
      totals[s] = a;
    }
    if (HASMAX) {
      double a = 0.0;
      for (int w = 0; w < W; w++) a = fmax(a, box[w][kRedSlots]);
      *totmax = a;
    }
    dd->red_seq = seq;
  }
  (void)sh;
}

template <int NS, bool HASMAX>
__device__ __forceinline__ bool grid_reduce_finish(const double *mine, double mymax,
                                                   double *partials, unsigned *ticket,
                                                   double *totals, double *totmax, double *sh,
                                                   DistDev *dd = nullptr) {
  __shared__ bool is_last;
  __shared__ double tot_sh[kRedSlots + 1];
  __syncthreads();   // every thread's global stores precede thread 0's fence + ticket
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; s++) partials[s * kMaxGrid + blockIdx.x] = mine[s];
    if (HASMAX) partials[kRedSlots * kMaxGrid + blockIdx.x] = mymax;
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  const volatile double *vp = partials;
#pragma unroll
  for (int s = 0; s < NS; s++) {
    double a = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a += vp[s * kMaxGrid + i];
    a = block_sum(a, sh);
    if (threadIdx.x == 0) totals[s] = a;
  }
  if (HASMAX) {
    double a = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
      a = fmax(a, vp[kRedSlots * kMaxGrid + i]);
    a = block_max(a, sh);
    if (threadIdx.x == 0) *totmax = a;
  }
  if (threadIdx.x == 0) *ticket = 0u;  // re-arm for the next kernel (stream-ordered)
  if (dd != nullptr && (NS > 0 || HASMAX)) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < NS; s++) tot_sh[s] = totals[s];
      if (HASMAX) tot_sh[kRedSlots] = *totmax;
    }
    dist_allreduce<NS, HASMAX>(dd, tot_sh, &tot_sh[kRedSlots], sh);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < NS; s++) totals[s] = tot_sh[s];
      if (HASMAX) *totmax = tot_sh[kRedSlots];
    }
  }
  return true;
}

#endif  // __CUDACC__

}  // namespace psp
