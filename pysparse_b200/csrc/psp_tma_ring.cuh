// psp_tma_ring.cuh -- the TMA-fed row-block SpMV pass, shared by the stand-alone
// SpMV kernel (psp_spmv.cu) and the persistent solver kernels (psp_persist.cu).
//
// A CTA walks the row blocks b = blockIdx.x, blockIdx.x + gridDim.x, ...  The
// block's slices of val / col / ind are brought into a STAGES-deep shared-
// memory ring by cp.async.bulk (TMA, completion on an mbarrier); one thread
// per row then walks its row in shared memory, gathers x through the read-only
// path and adds the rounded products in row order (the reference's order).
#pragma once
#include "psp_common.cuh"

namespace psp {

constexpr int kBlkNnz = 2044;             // entries per row block (+4 slack for alignment)
constexpr int kStageVal = 2048 * 8;       // bytes
constexpr int kStageCol = 2048 * 4;
constexpr int kStageInd = 272 * 4;        // kThreads + 1 rows, +3 alignment, rounded to 16 B
constexpr int kStageBytes = kStageVal + kStageCol + kStageInd;   // 25664, multiple of 128

__device__ __forceinline__ double ldg_f64(const double *p) { return __ldg(p); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

struct MatView {                          // device view of a CSR matrix + its row-block schedule
  const int *ind, *col;
  const int4 *blk;                        // (first row, first entry, end row, end entry) per block
  const double *val;
  int nb;
  int nint;                               // blocks [nint, nb) read halo columns (sharded matrices; else nb)
  const int2 *blk_halo;                   // per block: [first, end) halo positions its entries read (or nullptr)
};

template <int STAGES>
struct Ring {
  unsigned char *smem;                    // STAGES * kStageBytes, 128-byte aligned
  uint64_t *full;                         // STAGES mbarriers
  bool l2_stream;                         // bulk copies carry the L2 evict-first hint
  uint32_t phases;                        // parity bit per stage
  int cursor;                             // next stage to consume
  int prefetched;                         // number of blocks of the next pass already in flight

  __device__ __forceinline__ void init(unsigned char *smem_, uint64_t *full_, bool l2_stream_) {
    smem = smem_; full = full_; phases = 0; cursor = 0; prefetched = 0; l2_stream = l2_stream_;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  // thread 0 only: start the bulk copies of row block b into `stage`
  __device__ __forceinline__ void issue(const MatView &A, int b, int stage) {
    const int4 d = __ldg(A.blk + b);
    if (d.w - d.y > kBlkNnz) return;                // over-long row: read straight from global
    const int base = d.y & ~3, rbase = d.x & ~3;
    const uint32_t ne = (uint32_t)((d.w - base + 3) & ~3);
    const uint32_t nr = (uint32_t)((d.z + 1 - rbase + 3) & ~3);
    unsigned char *st = smem + (size_t)stage * kStageBytes;
    mbar_expect_tx(&full[stage], ne * 12u + nr * 4u);
    if (l2_stream) {
      // the matrix is read once per product: evict-first keeps the (re-used) vectors in L2
      uint64_t policy;                                // made on the spot: not worth two registers in every thread
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      tma_load_1d_hint(st, A.val + base, ne * 8u, &full[stage], policy);
      tma_load_1d_hint(st + kStageVal, A.col + base, ne * 4u, &full[stage], policy);
      tma_load_1d_hint(st + kStageVal + kStageCol, A.ind + rbase, nr * 4u, &full[stage], policy);
    } else {
      tma_load_1d(st, A.val + base, ne * 8u, &full[stage]);
      tma_load_1d(st + kStageVal, A.col + base, ne * 4u, &full[stage]);
      tma_load_1d(st + kStageVal + kStageCol, A.ind + rbase, nr * 4u, &full[stage]);
    }
  }

  // Start the first STAGES blocks of the NEXT pass (the matrix never changes, so this
  // can be done long before the vector to multiply is ready).  All threads call it.
  __device__ __forceinline__ void prefetch(const MatView &A) {
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // see pass(): in-place products
#pragma unroll
      for (int s = 0; s < STAGES; s++) {
        const int b = blockIdx.x + s * gridDim.x;
        if (b < A.nb) issue(A, b, (cursor + s) % STAGES);
      }
    }
    prefetched = 1;
  }

  // Consume a prefetch that will not be used (before the CTA exits).
  __device__ __forceinline__ void drain(const MatView &A) {
    if (!prefetched) return;
    for (int s = 0; s < STAGES; s++) {
      const int b = blockIdx.x + s * gridDim.x;
      if (b >= A.nb) break;
      const int4 d = __ldg(A.blk + b);
      if (d.w - d.y > kBlkNnz) continue;
      const int stage = (cursor + s) % STAGES;
      mbar_wait(&full[stage], (phases >> stage) & 1u);
      phases ^= (1u << stage);
    }
    prefetched = 0;
  }

  // One SpMV pass over this CTA's row blocks.  rowop(r, sum) is called exactly once
  // per row by the thread that holds the row sum.  red_sh: 8 doubles of scratch.
  // NC: gather x through the non-coherent read-only path (ld.global.nc) -- only legal
  // when x is not written during the kernel; the persistent kernels, which rewrite x
  // between grid barriers, use ordinary loads.
  //
  // The schedule lists the blocks that touch no halo column first (MatView::nint of
  // them); haloop(b) is called by all threads before this CTA's first block that does
  // (PER_BLOCK: before every such block) -- the sharded persistent kernels wait there for
  // the neighbours' boundary values, which have been travelling while the interior blocks
  // were multiplied.
  template <bool NC, class RowOp>
  __device__ __forceinline__ void pass(const MatView &A, const double *x, double *red_sh, RowOp rowop) {
    pass<NC>(A, x, red_sh, rowop, [](int) {});
  }
  template <bool NC, bool PER_BLOCK = false, class RowOp, class HaloOp>
  __device__ __forceinline__ void pass(const MatView &A, const double *x, double *red_sh, RowOp rowop, HaloOp haloop) {
    const int tid = threadIdx.x;
    if (!prefetched) prefetch(A);
    prefetched = 0;
    bool halo_ready = false;
    for (int b = blockIdx.x; b < A.nb; b += gridDim.x) {
      if (b >= A.nint && (PER_BLOCK || !halo_ready)) { haloop(b); halo_ready = true; }
      const int stage = cursor;
      cursor = (cursor + 1) % STAGES;
      const int4 d = __ldg(A.blk + b);
      const int r0 = d.x, r1 = d.z, k0 = d.y, k1 = d.w;
      const int nrows = r1 - r0, cnt = k1 - k0;
      bool wrote_stage = false;                     // stage buffer was written through the generic proxy
      if (cnt <= kBlkNnz) {
        const unsigned char *st = smem + (size_t)stage * kStageBytes;
        const double *sval = reinterpret_cast<const double *>(st);
        const int *scol = reinterpret_cast<const int *>(st + kStageVal);
        const int *sind = reinterpret_cast<const int *>(st + kStageVal + kStageCol);
        const int base = k0 & ~3, roff = r0 - (r0 & ~3);
        mbar_wait(&full[stage], (phases >> stage) & 1u);
        phases ^= (1u << stage);                    // flips only when the stage was really used
        if (cnt <= 16 * nrows) {
          // short rows: one thread per row, products added in row (= reference) order
          if (tid < nrows) {
            const int s = sind[roff + tid] - base, e = sind[roff + tid + 1] - base;
            double sum = 0.0;
            for (int k = s; k < e; k += 8) {
              double pr[8];
#pragma unroll
              for (int j = 0; j < 8; j++)
                pr[j] = (k + j < e) ? __dmul_rn(sval[k + j], NC ? ldg_f64(x + scol[k + j]) : x[scol[k + j]]) : 0.0;
#pragma unroll
              for (int j = 0; j < 8; j++)
                if (k + j < e) sum = __dadd_rn(sum, pr[j]);
            }
            rowop(r0 + tid, sum);
          }
        } else {
          // longer / irregular rows (mean > 16 entries): two phases per block.
          // (1) entry-parallel: thread t takes entries t, t + 256, ... -- its x gathers are
          //     independent and go out four at a time (a row-per-warp walk has ONE gather in flight
          //     per warp and spends its time waiting for L2: 113 G gathers/s on the 2^22-row
          //     power-law matrix, where the chip sustains 200); the rounded products replace the
          //     values in the stage buffer.
          // (2) one thread per row adds its products in row order -- the reference's order, so
          //     these rows are bit-identical to it as well; rows longer than kSeqRow are left to
          //     whole warps (strided partial sums, then a shuffle tree).
          // Measured on the 2^21-row power-law matrix (profiles/r2_irregular_notes.md): the gather
          // flavour (ld.global.nc or plain) and the batch (2 / 4 / 8) make no difference, the ring
          // depth does (2 / 3 / 4 stages: 538 / 462 / 657 us), kSeqRow 96 / 48 / 32: 462 / 437 / 446 us.
#ifndef PSP_IRR_SEQ
#define PSP_IRR_SEQ 48
#endif
          constexpr int kSeqRow = PSP_IRR_SEQ;
          __shared__ int long_rows[kBlkNnz / kSeqRow + 2];
          __shared__ int n_long;
          double *pv = const_cast<double *>(sval);
          const int kb = k0 - base, ke = k1 - base;
          if (tid == 0) n_long = 0;
#ifndef PSP_IRR_BATCH
#define PSP_IRR_BATCH 4                              // eight at once spill under the 32-register cap
#endif
#ifdef PSP_IRR_PLAIN
#define PSP_IRR_LOAD(p) (*(p))
#else
#define PSP_IRR_LOAD(p) (NC ? ldg_f64(p) : *(p))
#endif
#pragma unroll 1
          for (int h = 0; h < 8; h += PSP_IRR_BATCH) {
            double xv[PSP_IRR_BATCH];
#pragma unroll
            for (int j = 0; j < PSP_IRR_BATCH; j++) {
              const int k = kb + tid + (h + j) * kThreads;
              xv[j] = (k < ke) ? PSP_IRR_LOAD(x + scol[k]) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < PSP_IRR_BATCH; j++) {
              const int k = kb + tid + (h + j) * kThreads;
              if (k < ke) pv[k] = __dmul_rn(sval[k], xv[j]);
            }
          }
          wrote_stage = true;
          __syncthreads();
          if (tid < nrows) {
            const int s = sind[roff + tid] - base, e = sind[roff + tid + 1] - base;
            if (e - s <= kSeqRow) {
              double sum = 0.0;
              for (int k = s; k < e; k += 8) {
                double pr[8];
#pragma unroll
                for (int j = 0; j < 8; j++) pr[j] = (k + j < e) ? pv[k + j] : 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++)
                  if (k + j < e) sum = __dadd_rn(sum, pr[j]);
              }
              rowop(r0 + tid, sum);
            } else {
              long_rows[atomicAdd(&n_long, 1)] = tid;
            }
          }
          __syncthreads();
          const int nl = n_long;
          if (nl > 0) {
            const int w = tid >> 5, l = tid & 31;
            for (int q = w; q < nl; q += kThreads / 32) {
              const int rr = long_rows[q];
              const int s = sind[roff + rr] - base, e = sind[roff + rr + 1] - base;
              double sum = 0.0;
              for (int k = s + l; k < e; k += 32) sum += pv[k];
              sum = warp_sum(sum);
              if (l == 0) rowop(r0 + rr, sum);
            }
          }
        }
      } else {
        // one over-long row: whole CTA straight from global memory
        double sum = 0.0;
        for (int k = k0 + tid; k < k1; k += kThreads)
          sum += __ldg(A.val + k) * (NC ? ldg_f64(x + __ldg(A.col + k)) : x[__ldg(A.col + k)]);
        sum = block_sum(sum, red_sh);
        if (tid == 0) rowop(r0, sum);
      }
      __syncthreads();                              // everyone is done with this stage's buffers
      if (tid == 0) {
        const int nb2 = b + STAGES * gridDim.x;
        // our in-place writes (generic proxy) must be ordered before the TMA (async proxy)
        // overwrites the same buffer
        if (wrote_stage) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (nb2 < A.nb) issue(A, nb2, stage);
      }
    }
  }
};

}  // namespace psp
