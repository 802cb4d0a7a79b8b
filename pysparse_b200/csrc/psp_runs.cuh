// psp_runs.cuh -- the "diagonal run" device format of a sparse matrix and the SpMV pass over
// it, shared by the stand-alone kernel (psp_runs.cu) and the persistent solver kernels.
//
// The matrices of the path (finite-difference / finite-element operators assembled through
// ll_mat) store long stretches of consecutive rows whose entries sit on the same diagonal:
// row r has an entry in column c, row r+1 in column c+1, ...  A RUN is such a stretch inside
// one block of kThreads consecutive rows: (first local row t0, length, first column, position of
// its first value).  The values of a run are stored back to back, so thread t of the CTA (row
// r0 + t) reads val[kpos + t - t0] and x[xoff + t - t0]: both streams are fully coalesced
// 8-byte-per-lane loads, and NO column index is read at all -- the index arrays of the CSR
// layout (4 bytes per entry, a third of the matrix traffic) are replaced by 16 bytes per run.
// The format is lossless (psp_runs.cu converts both ways) and is chosen by the analysis in
// psp_csr_create / psp_sss_create only when the runs are long; irregular matrices keep the
// row-block CSR stream of psp_tma_ring.cuh.
//
// Symmetric (SSS) matrices need it for a second reason: sss_matvec (sss_mat.c:40-56) uses
// every stored entry L[i,j] twice, y[i] += v x[j] and y[j] += v x[i].  Here the value is STORED
// ONCE, in the run of its row i; the mirrored use is a second run descriptor in the block of
// row j that points at the same values (kpos) -- the transpose index costs 16 bytes per run
// instead of a second copy of the values, nothing is scattered, no atomics, and the second
// read of a value comes from L2 (the blocks that share values are processed close in time).
//
// Order of accumulation: the runs of a block are sorted by diagonal offset, so every row adds
// its products in ascending column order, each product rounded before it is added (no FMA):
// bit-identical to csr_matvec_kernel (csr_mat.c:49-54); for SSS the order is lower entries,
// diagonal, upper entries ascending -- the order in which sss_matvec builds y[i].
#pragma once
#include <cstring>
#include <type_traits>
#include "psp_common.cuh"

namespace psp {

// A run may bridge a short gap (a row that has no entry on the run's diagonal: the first node of
// every grid line on the -1 diagonal of a stencil matrix): the gap's slot of the value array holds
// this quiet-NaN pattern and the kernel skips it -- without it every such gap would split the
// run and a 256-row block of a 384-wide grid would need 9-10 descriptors instead of 7, i.e. two
// batches (two memory round trips) instead of one.  The analysis does not bridge when a stored
// value of the matrix has exactly this bit pattern.
constexpr unsigned long long kRunSkip = 0xFFF8C0DE5EED0002ull;
constexpr int kRunGapMax = 2;             // rows a run may bridge
constexpr int kRunGapFlag = 1 << 30;      // in a block descriptor's row count: some run of the block covers a gap
constexpr int kRunBatch = 8;              // descriptors per batch of loads in run_pass (its UNROLL)
constexpr int kRunGapMinLen = 16;         // ... once it is at least this long
__host__ __device__ __forceinline__ bool run_skip(double v) {
#ifdef __CUDA_ARCH__
  return (unsigned long long)__double_as_longlong(v) == kRunSkip;
#else
  unsigned long long b; memcpy(&b, &v, 8); return b == kRunSkip;
#endif
}

struct RunView {
  const double *val;        // run-ordered values (SSS: strict lower triangle, then the diagonal)
  const int4 *seg;          // (kpos, xoff, t0, len) per run, in the order of the block schedule
  const int4 *blk;          // (first row, rows, first run, end run) per row block, schedule order
  int nb;
  int nint;                 // blocks [nint, nb) read halo columns (row slabs of sharded matrices)
  const int2 *blk_halo;     // per block: [first, end) halo positions its runs read (or nullptr)
  int chunk;                // consecutive blocks a CTA takes at a time (1 .. kRunChunkMax)
};

constexpr int kRunChunkMax = 8;      // blocks per chunk
constexpr int kRunSegCache = 96;     // run descriptors of one chunk kept in shared memory

struct RunSmem {                     // descriptor staging of run_pass (3.4 KB)
  int4 blk[3 * kRunChunkMax];        // block descriptors of the current chunk and the next two
  int4 seg[2 * kRunSegCache];        // run descriptors of the current and the next chunk
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// One SpMV pass.  A CTA takes CHUNKS of A.chunk consecutive row blocks, chunk blockIdx.x,
// +gridDim.x, ...; thread t owns row r0 + t of a block.  rowop(r, sum) is called once per row
// by the thread holding the sum.
//
// The pass is organised around memory latency.  Descriptors never sit on the critical path:
// the block descriptors of chunk c+2 and the run descriptors of chunk c+1 are copied into
// shared memory with cp.async while chunk c is being multiplied (the CTA synchronises once per
// chunk, not per block).  For a block, the loads of UNROLL runs -- 2 * UNROLL independent
// 8-byte loads per thread -- are all issued before the first product is added, so a block with
// up to UNROLL runs costs ONE memory round trip and the warps of a CTA drift freely inside a
// chunk.  Chunks of consecutive blocks interleaved over the CTAs keep the reuse distance of
// the mirrored runs of an SSS matrix (second use of a value) short: it is served from L2.
// NC: gather x through the read-only path (only legal when x is not written by this kernel).
// haloop: see Ring::pass -- called before the first (PER_BLOCK: every) block that reads halo
// columns.
template <bool NC, int UNROLL, bool PER_BLOCK, class RowOp, class HaloOp>
__device__ __forceinline__ void run_pass(const RunView &A, const double *x, RunSmem &sm, RowOp rowop, HaloOp haloop) {
  const int t = threadIdx.x;
  const int C = A.chunk, G = gridDim.x;
  const int nch = (A.nb + C - 1) / C;
  int ci = blockIdx.x;
  if (ci >= nch) return;
  bool halo_ready = false;
  auto fetch_blk = [&](int chunk, int slot) {
    if (chunk < nch) {
      const int b0 = chunk * C, cnt = (A.nb - b0 < C) ? A.nb - b0 : C;
      if (t < cnt) cp_async16(&sm.blk[slot * kRunChunkMax + t], A.blk + b0 + t);
    }
  };
  auto fetch_seg = [&](int chunk, int bslot, int sslot) {      // the chunk's block descriptors have landed
    if (chunk < nch) {
      const int b0 = chunk * C, cnt = (A.nb - b0 < C) ? A.nb - b0 : C;
      const int s0 = sm.blk[bslot * kRunChunkMax].z, s1 = sm.blk[bslot * kRunChunkMax + cnt - 1].w;
      const int ns = (s1 - s0 < kRunSegCache) ? s1 - s0 : kRunSegCache;
      for (int i = t; i < ns; i += blockDim.x) cp_async16(&sm.seg[sslot * kRunSegCache + i], A.seg + s0 + i);
    }
  };
  fetch_blk(ci, 0);
  fetch_blk(ci + G, 1);
  cp_async_wait_all();
  __syncthreads();
  fetch_seg(ci, 0, 0);
  cp_async_wait_all();
  __syncthreads();
  int bs = 0, ss = 0;                                // slots of the current chunk
  for (; ci < nch; ci += G) {
    const int bs1 = bs == 2 ? 0 : bs + 1, bs2 = bs1 == 2 ? 0 : bs1 + 1;
    fetch_blk(ci + 2 * G, bs2);
    fetch_seg(ci + G, bs1, ss ^ 1);
    const int4 *sb = sm.blk + bs * kRunChunkMax;
    const int4 *sg = sm.seg + ss * kRunSegCache;
    const int b0 = ci * C, cnt = (A.nb - b0 < C) ? A.nb - b0 : C;
    const int sbase = sb[0].z;
    for (int jb = 0; jb < cnt; jb++) {
      const int b = b0 + jb;
      if (b >= A.nint && (PER_BLOCK || !halo_ready)) { haloop(b); halo_ready = true; }
      const int4 d = sb[jb];
      const bool gaps = (d.y & kRunGapFlag) != 0;
      const int nrows = d.y & ~kRunGapFlag;
      const int o = d.z - sbase, ns = d.w - d.z;       // this block's runs: cache positions o .. o+ns-1
      double sum = 0.0;
      if (o + ns <= kRunSegCache) {
        // Three separate phases per batch of UNROLL runs.  (1) descriptors -> element indices,
        // (2) all loads, (3) the adds.  Nothing in (2) depends on a load in flight: a warp has six
        // scoreboards, and a descriptor read placed between the global loads ends up waiting
        // on a scoreboard it shares with them -- one DRAM round trip per run instead of per batch.
        // (ptxas still forms the first product before the last loads are issued; forcing all
        // loads first through an artificial data dependency measured slower, profiles/r2_runs_notes.md.)
        // GAPS: some run of this block bridges a gap -- only then are the loaded values tested for
        // the marker (the test costs two instructions per run in an issue-bound loop: 3 % at 256^3)
        auto batches = [&](auto GAPS) {
          for (int s = 0; s < ns; s += UNROLL) {
            unsigned iv[UNROLL], ix[UNROLL];
            bool in[UNROLL];
            double v[UNROLL], xv[UNROLL];
#pragma unroll
            for (int j = 0; j < UNROLL; j++) {
              const int i = s + j;
              const int4 g = sg[o + (i < ns ? i : 0)];
              const unsigned u = (unsigned)(t - g.z);
              in[j] = (i < ns) && u < (unsigned)g.w;
              iv[j] = (unsigned)g.x + u;
              ix[j] = (unsigned)g.y + u;
            }
#pragma unroll
            for (int j = 0; j < UNROLL; j++) {
              v[j] = in[j] ? __ldg(A.val + iv[j]) : 0.0;
              xv[j] = in[j] ? (NC ? __ldg(x + ix[j]) : x[ix[j]]) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < UNROLL; j++) {           // conditional: adding a zero product would turn -0.0 into +0.0
              bool use = in[j];
              if (decltype(GAPS)::value) use = use && !run_skip(v[j]);
              if (use) sum = __dadd_rn(sum, __dmul_rn(v[j], xv[j]));
            }
          }
        };
        if (gaps) batches(std::true_type());
        else batches(std::false_type());
      } else {
        // more runs than the cache holds (irregular block): descriptors straight from global memory
        for (int i = 0; i < ns; i++) {
          const int4 g = __ldg(A.seg + d.z + i);
          const unsigned u = (unsigned)(t - g.z);
          if (u < (unsigned)g.w) {
            const double *xp = x + (size_t)(unsigned)g.y + u;
            const double av = __ldg(A.val + (size_t)(unsigned)g.x + u);
            if (!run_skip(av)) sum = __dadd_rn(sum, __dmul_rn(av, NC ? __ldg(xp) : *xp));
          }
        }
      }
      if (t < nrows) rowop(d.x + t, sum);
    }
    cp_async_wait_all();
    __syncthreads();
    bs = bs1;
    ss ^= 1;
  }
}

}  // namespace psp
