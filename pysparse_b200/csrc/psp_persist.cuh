// psp_persist.cuh -- infrastructure shared by the single-launch (persistent, cooperative)
// solver kernels: the ticket/flag grid all-reduce that doubles as the grid barrier, the
// device-side scalar epilogue on a shared-memory copy of the solver state, and the multi-GPU
// halo exchange (push model with self-validating words; pull model as the fallback).
// See psp_persist.cu for the PCG kernel built on it.
#pragma once
#include <cooperative_groups.h>
#include "psp_mat.cuh"
#include "psp_epilogue.cuh"
#include "psp_tma_ring.cuh"
#include "psp_runs.cuh"
#include "psp_dist.cuh"

namespace cg = cooperative_groups;

namespace psp {

// tuning constants (overridable for A/B builds: PSP_NVCC_EXTRA="-DPSP_PERSIST_CTAS=3 ...")
#ifndef PSP_PERSIST_STAGES
#define PSP_PERSIST_STAGES 2
#endif
#ifndef PSP_PERSIST_CTAS
#define PSP_PERSIST_CTAS 4
#endif
#ifndef PSP_VEC_UNROLL
#define PSP_VEC_UNROLL 1
#endif
// Halo unpack of the sharded single-launch PCG: block-local by default (every boundary row block
// copies the landing-zone words it needs right before it is multiplied; no unpack pass, no
// tickets): 83.0 -> 80.5 us per iteration on 8 GPUs, 163 -> 159 on 4 (profiles/r2_multi_notes.md).
// -DPSP_NO_BLOCK_UNPACK restores the separate unpack pass.
#if !defined(PSP_NO_BLOCK_UNPACK) && !defined(PSP_BLOCK_UNPACK)
#define PSP_BLOCK_UNPACK 1
#endif
constexpr int kPersistStages = PSP_PERSIST_STAGES;   // TMA ring depth
constexpr int kPersistCtas = PSP_PERSIST_CTAS;       // CTAs per SM the register budget is sized for
constexpr int kVecUnroll = PSP_VEC_UNROLL;           // elements per thread and trip in the vector phases
#ifndef PSP_PERSIST_RUN_UNROLL
#define PSP_PERSIST_RUN_UNROLL 8
#endif
constexpr int kPersistRunUnroll = PSP_PERSIST_RUN_UNROLL;   // runs per memory round trip in the diagonal-run pass
#ifndef PSP_PERSIST_RUN_CTAS
#define PSP_PERSIST_RUN_CTAS 4
#endif
// CTAs per SM the register budget is sized for, by matrix format.  Measured at 256^3 (PCG +
// Jacobi, it/s CSR / SSS): 4 CTAs x 64 registers, 8 runs per batch 2030 / 2163; 3 CTAs x 80
// registers with 2 elements per thread in the vector phases 1976 / 2165; 2 CTAs x 128 registers
// 1686 / 1852 (profiles/r2_runs_notes.md).
__host__ __device__ constexpr int persist_ctas(int fmt) { return fmt == 1 ? PSP_PERSIST_RUN_CTAS : kPersistCtas; }
#ifndef PSP_RUN_VEC_UNROLL
#define PSP_RUN_VEC_UNROLL 1
#endif
// fewer resident threads (format 1) need more loads in flight per thread in the vector phases
__host__ __device__ constexpr int vec_unroll(int fmt) { return fmt == 1 ? PSP_RUN_VEC_UNROLL : kVecUnroll; }

struct PersistSync;

// The part of a persistent solver's kernel arguments the shared helpers use (lives in the
// parameter bank: costs no registers).
struct PersistArgs {
  MatView A;                    // CSR row blocks (matrix format 0)
  RunView R;                    // diagonal runs (matrix format 1); exactly one of the two is set
  int n;                        // local rows
  State *S;
  double *partials;             // 2 x (kRedSlots + 1) x kMaxGrid, double-buffered
  PersistSync *sync;
  DistDev *dd;                  // nullptr on a single GPU
  int l2_stream;
  int dist_debug;               // timing diagnostics (see Ctx::dist_debug)
  // multi-GPU
  const int *halo_owner, *halo_ridx;
  int nhalo;
  unsigned nbmask;
  // push model (send map present)
  int push, nsend;
  const int *send_row;
  unsigned long long *const *send_dst;    // per entry: its two words in the neighbour's landing zone
};

struct PersistSync {            // global scratch of the all-reduce / barrier (zeroed before each launch)
  unsigned ticket, pad;
  unsigned long long flag;      // generation of the last completed all-reduce
  unsigned long long unpacked;  // halo exchanges x CTAs that have unpacked their share
  double result[2][kRedSlots + 1];
};

// Per-thread view of the kernel's bookkeeping.  Kept small on purpose: the kernel runs
// at 64 registers per thread (4 CTAs per SM), everything that can live in the parameter
// bank (PcgArgs) or be recomputed does.
struct PersistCtx {
  State *sS;                    // shared-memory copy of the solver state
  double *tot;                  // shared: kRedSlots + 1 totals
  double *red_sh;               // shared: 8 * (kRedSlots + 1) doubles
  DistDev *dd;
  // low 32 bits of the reduction / halo-exchange sequence numbers (DistDev holds the
  // 64-bit values between launches), all-reduce generation and unpack tickets expected
  unsigned seq, pubs, gen, unpack_target;
  int debug;
#ifdef PSP_PERSIST_PROF
  long long prof[16], t_last;
#endif
};

// diagnostic build (PSP_NVCC_EXTRA=-DPSP_PERSIST_PROF): CTA 0 / thread 0 accumulates the SM
// cycles it spends in each segment of the iteration and prints them when the kernel ends
#ifdef PSP_PERSIST_PROF
#define PROF_T(c, k)                                                     \
  do {                                                                   \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                           \
      const long long t_ = clock64();                                    \
      (c).prof[k] += t_ - (c).t_last;                                    \
      (c).t_last = t_;                                                   \
    }                                                                    \
  } while (0)
#else
#define PROF_T(c, k)
#endif

// Sums (and one max) of up to kRedSlots + 1 per-thread values over the CTA in ONE pass:
// warp shuffles, one shared-memory exchange, two CTA barriers.  Result in thread 0.
// sh: 8 * (kRedSlots + 1) doubles.  Fixed order, so every CTA / rank / run agrees.
template <int NS, bool HASMAX>
__device__ __forceinline__ void block_reduce_multi(double *v, double &mx, double *sh) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int s = 0; s < NS; s++) v[s] = warp_sum(v[s]);
  if (HASMAX) mx = warp_max(mx);
  __syncthreads();
  if (l == 0) {
#pragma unroll
    for (int s = 0; s < NS; s++) sh[s * 8 + w] = v[s];
    if (HASMAX) sh[kRedSlots * 8 + w] = mx;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int s = 0; s < NS; s++) {
      double r = (l < 8) ? sh[s * 8 + l] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      v[s] = r;
    }
    if (HASMAX) {
      double r = (l < 8) ? sh[kRedSlots * 8 + l] : 0.0;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
      mx = r;
    }
  }
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Grid-wide deterministic all-reduce AND grid barrier in one: every CTA deposits its
// partial sums and draws a ticket; the CTA that draws the last one adds the partials up
// in a fixed order (so the result does not depend on which CTA that is), does the
// cross-GPU exchange through the peer mailboxes when the solve is sharded, publishes the
// totals and releases the others, which have been spinning on one flag.  Only one CTA
// reads the partials and only one polls the mailboxes.  On return tot[0..NS) (and
// tot[kRedSlots] for the max) are valid in every thread of every CTA, and all global
// writes made before the call are visible grid-wide.
//
// last_op(tot) is run by thread 0 of the CTA with the last ticket after the totals (cross-GPU
// included) are known and BEFORE the others are released: the place for scalar work that
// must happen exactly once (solvers whose epilogue updates global arrays, e.g. GMRES).
template <int NS, bool HASMAX, class LastOp>
__device__ __forceinline__ void grid_allreduce_op(PersistCtx &c, const PersistArgs &a, const double *acc, double mx,
                                                  LastOp last_op) {
  __shared__ int is_last;
  __shared__ unsigned box[kMaxRanks][8];
  const unsigned gen = ++c.gen;
  double *buf = a.partials + (size_t)(gen & 1u) * (kRedSlots + 1) * kMaxGrid;
  double *res = a.sync->result[gen & 1u];
  const bool xgpu = c.dd && c.debug < 2;
  const unsigned seq = xgpu ? ++c.seq : c.seq;                // every CTA keeps count
  double v[NS > 0 ? NS : 1];
#pragma unroll
  for (int s = 0; s < NS; s++) v[s] = acc[s];
  block_reduce_multi<NS, HASMAX>(v, mx, c.red_sh);
  PROF_T(c, 8);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; s++) buf[s * kMaxGrid + blockIdx.x] = v[s];
    if (HASMAX) buf[kRedSlots * kMaxGrid + blockIdx.x] = mx;
    __threadfence();
    const unsigned t = atomicAdd(&a.sync->ticket, 1u);
    __threadfence();
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    double sm[NS > 0 ? NS : 1], m = 0.0;
#pragma unroll
    for (int s = 0; s < NS; s++) sm[s] = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
#pragma unroll
      for (int s = 0; s < NS; s++) sm[s] += __ldcg(buf + s * kMaxGrid + i);
      if (HASMAX) m = fmax(m, __ldcg(buf + kRedSlots * kMaxGrid + i));
    }
    block_reduce_multi<NS, HASMAX>(sm, m, c.red_sh);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < NS; s++) c.tot[s] = sm[s];
      if (HASMAX) c.tot[kRedSlots] = m;
    }
    if (xgpu) {
      // cross-GPU, one NVLink trip: thread (w, j) stores word j of this rank's totals into
      // rank w's mailbox (payload + sequence tag in one 8-byte word, see DistMail::ll) and
      // spins on the LOCAL mailbox until rank w's word j of the same sequence number is
      // there; thread 0 then adds the W contributions in rank order.
      constexpr int NV = NS + (HASMAX ? 1 : 0);
      static_assert(2 * NV <= 8, "DistMail::ll holds four values");
      __syncthreads();
      DistDev *dd = c.dd;
      const int W = dd->world, me = dd->rank;
      const int slot = (int)(seq % kMailRing);
      const unsigned tag = seq;
      if constexpr (NV > 0) {                       // NV == 0: a pure barrier, nothing to exchange
        if ((int)threadIdx.x < W * 2 * NV) {
          const int w = threadIdx.x / (2 * NV), j = threadIdx.x % (2 * NV);
          const int k = j >> 1;
          const unsigned long long bits = (unsigned long long)__double_as_longlong(k < NS ? c.tot[k] : c.tot[kRedSlots]);
          const unsigned half = (j & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
          st_relaxed_sys_u64(&dd->ctl[w]->mail[slot][me].ll[j], ((unsigned long long)tag << 32) | half);
          const unsigned long long *src = &dd->ctl[me]->mail[slot][w].ll[j];
          unsigned long long got;
          do { got = ld_relaxed_sys_u64(src); } while ((unsigned)(got >> 32) != tag);
          box[w][j] = (unsigned)got;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
          double t = 0.0;
          for (int w = 0; w < W; w++) {
            const double v = __longlong_as_double((long long)(((unsigned long long)box[w][2 * k + 1] << 32) | box[w][2 * k]));
            t = (k < NS) ? t + v : fmax(t, v);
          }
          c.tot[k < NS ? k : kRedSlots] = t;
        }
      }
    }
    if (threadIdx.x == 0) {
      last_op(c.tot);
#pragma unroll
      for (int s = 0; s < NS; s++) res[s] = c.tot[s];
      if (HASMAX) res[kRedSlots] = c.tot[kRedSlots];
      a.sync->ticket = 0u;               // nobody draws again before the flag moves
      __threadfence();
      st_release_gpu_u64(&a.sync->flag, (unsigned long long)gen);
    }
  } else if (threadIdx.x == 0) {
    while (ld_relaxed_gpu_u64(&a.sync->flag) < (unsigned long long)gen) { }
    __threadfence();
#pragma unroll
    for (int s = 0; s < NS; s++) c.tot[s] = __ldcg(res + s);
    if (HASMAX) c.tot[kRedSlots] = __ldcg(res + kRedSlots);
  }
  __syncthreads();
  PROF_T(c, 9);
}

template <int NS, bool HASMAX>
__device__ __forceinline__ void grid_allreduce(PersistCtx &c, const PersistArgs &a, const double *acc, double mx) {
  grid_allreduce_op<NS, HASMAX>(c, a, acc, mx, [](const double *) {});
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
__device__ __forceinline__ void publish_and_pull(cg::grid_group &grid, PersistCtx &c, const PersistArgs &a, double *vec,
                                                 size_t slot_off) {
  DistDev *dd = c.dd;
  if (c.debug >= 1) { grid.sync(); return; }
  // 64-bit publication number: the value at launch plus the exchanges done since (inside
  // this kernel every publication is followed by exactly one pull)
  const unsigned long long want = dd->pubs_done + (unsigned)(c.pubs - (unsigned)dd->pubs_done) + 1;
  c.pubs += 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(&dd->ctl[dd->rank]->pub, want);
  }
  if (threadIdx.x == 0) {
    for (int w = 0; w < dd->world; w++)
      if ((a.nbmask >> w) & 1u)
        while (ld_acquire_sys(&dd->ctl[w]->pub) < want) { }
  }
  __syncthreads();
  const int stride = gridDim.x * blockDim.x;
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < a.nhalo; h += stride)
    vec[a.n + h] = ld_relaxed_sys_f64(dd->heap[a.halo_owner[h]] + slot_off + a.halo_ridx[h]);
  grid.sync();
}

// Push model (needs the send map, psp_dist_set_sendmap).  The search direction is double-
// buffered (p_cur is read, p_new is written), so the boundary entries of p_new can be
// computed and sent at the START of the p-update phase: each one goes straight into the
// neighbour's landing zone as two self-validating 8-byte words (half of the value + the
// exchange number, DistDev::ll_off) -- posted NVLink stores, no fence, no flag, nobody
// waits for them.  After the grid barrier that ends the phase every CTA unpacks its share
// of the landing zone into the halo part of p_new (spinning on LOCAL memory until the
// words of this exchange are there -- they left a whole phase ago) and draws a ticket;
// the row blocks that read halo columns are scheduled last and wait for all tickets, so
// the exchange overlaps the p update and the interior blocks.  No remote loads, no
// system-scope fences (MEMBAR.SYS costs ~5 us here), no extra grid barrier.
__device__ __forceinline__ void halo_send(const PersistCtx &c, const PersistArgs &a, unsigned s, double v) {
  const unsigned long long tag = (unsigned long long)(c.pubs + 1u) << 32;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  unsigned long long *dst = a.send_dst[s];
  st_relaxed_sys_u64(dst, tag | (bits & 0xffffffffull));
  st_relaxed_sys_u64(dst + 1, tag | (bits >> 32));
}

// after the grid barrier that completes the own part of `vec`
__device__ __forceinline__ void halo_unpack(PersistCtx &c, const PersistArgs &a, double *vec) {
  DistDev *dd = c.dd;
  c.pubs += 1;
  c.unpack_target += gridDim.x;
  if (c.debug >= 1) return;
  const unsigned tag = c.pubs;
  const unsigned long long *ll =
      reinterpret_cast<const unsigned long long *>(dd->heap[dd->rank] + dd->ll_off) + 2 * (size_t)a.n;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long h = (long)blockIdx.x * blockDim.x + threadIdx.x; h < a.nhalo; h += stride) {
    unsigned long long w0, w1;                      // both words in one 16-byte load; each validates itself
    do {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(ll + 2 * h) : "memory");
    } while ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag);
    vec[a.n + h] = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&a.sync->unpacked, 1ull);
  }
}

// Block-local unpack (the PCG kernel uses it in the A/B build -DPSP_BLOCK_UNPACK only): no
// unpack pass and no tickets.  A boundary block's CTA unpacks, right before the block, exactly
// the halo positions that block reads (MatView::blk_halo); ranges of different blocks may
// overlap, the CTAs then write identical values.
__device__ __forceinline__ void halo_unpack_block(PersistCtx &c, const PersistArgs &a, double *vec, int b) {
  if (c.debug >= 1) return;
  DistDev *dd = c.dd;
  const int2 hr = __ldg(a.A.blk_halo + b);
  const unsigned tag = c.pubs;
  const unsigned long long *ll =
      reinterpret_cast<const unsigned long long *>(dd->heap[dd->rank] + dd->ll_off) + 2 * (size_t)a.n;
  for (int h = hr.x + (int)threadIdx.x; h < hr.y; h += (int)blockDim.x) {
    unsigned long long w0, w1;
    do {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(ll + 2 * (size_t)h) : "memory");
    } while ((unsigned)(w0 >> 32) != tag || (unsigned)(w1 >> 32) != tag);
    vec[a.n + h] = __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
  }
  __syncthreads();
}

// Late push: after the grid barrier that completes `vec`, every thread sends its share of the
// boundary entries as they are (no recomputation, so no double buffering); to be paired with
// halo_unpack_block.  Counts as one exchange.
__device__ __forceinline__ void halo_push_late(PersistCtx &c, const PersistArgs &a, const double *vec) {
  if (c.debug < 1) {
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < (unsigned)a.nsend; s += stride)
      halo_send(c, a, s, vec[a.send_row[s]]);
  }
  c.pubs += 1;
}

// before the first row block that reads a halo column
__device__ __forceinline__ void halo_wait(PersistCtx &c, const PersistArgs &a) {
  if (c.debug >= 1) return;
  if (threadIdx.x == 0) {
    while (ld_relaxed_gpu_u64(&a.sync->unpacked) < (unsigned long long)c.unpack_target) { }
    __threadfence();
  }
  __syncthreads();
}

// The SpMV pass of a persistent kernel, by matrix format.  FMT 0: the TMA-fed row-block ring
// (prefetches the first blocks of the NEXT pass while the vector phases run); FMT 1: the
// diagonal-run pass (plain coalesced loads, nothing to prefetch or drain, no shared memory).
template <int FMT> struct MatPass;
template <> struct MatPass<0> {
  Ring<kPersistStages> ring;
  __device__ __forceinline__ void init(unsigned char *smem, uint64_t *full, bool l2_stream, RunSmem *) { ring.init(smem, full, l2_stream); }
  __device__ __forceinline__ void prefetch(const PersistArgs &a) { ring.prefetch(a.A); }
  __device__ __forceinline__ void drain(const PersistArgs &a) { ring.drain(a.A); }
  template <bool PER_BLOCK, class RowOp, class HaloOp>
  __device__ __forceinline__ void pass(const PersistArgs &a, const double *x, double *red_sh, RowOp rowop, HaloOp haloop) {
    ring.template pass<false, PER_BLOCK>(a.A, x, red_sh, rowop, haloop);
  }
};
template <> struct MatPass<1> {
  RunSmem *sseg;                // shared: descriptor staging of run_pass
  __device__ __forceinline__ void init(unsigned char *, uint64_t *, bool, RunSmem *sseg_) { sseg = sseg_; }
  __device__ __forceinline__ void prefetch(const PersistArgs &) {}
  __device__ __forceinline__ void drain(const PersistArgs &) {}
  template <bool PER_BLOCK, class RowOp, class HaloOp>
  __device__ __forceinline__ void pass(const PersistArgs &a, const double *x, double *, RowOp rowop, HaloOp haloop) {
    run_pass<false, kPersistRunUnroll, PER_BLOCK>(a.R, x, *sseg, rowop, haloop);
  }
};

// ---- host side (psp_persist.cu)
// fills the shared part of the arguments; zeroes the barrier scratch on the stream
int persist_prepare(PersistArgs &a, const psp_mat *M, int n, DistMat *dm);
// co-resident grid for a persistent kernel: enough CTAs for the vector phases and the row
// blocks, never more than fit (`resident` = CTAs per SM x SMs, from the occupancy API)
int persist_grid(const psp_mat *M, int n, int resident);

}  // namespace psp
