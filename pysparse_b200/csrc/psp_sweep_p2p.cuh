// psp_sweep_p2p.cuh -- SSOR / symmetric Gauss-Seidel sweeps without any global barrier:
// point-to-point dependencies between rows (the "wavefront strip" engine).
//
// The level-scheduled engines of psp_precon.cu synchronise the whole cluster / grid once per
// level: 16 382 dependent barriers per application on a 4096^2 grid, 3.5 us each.  The real
// critical path of symgs_kernel / ssor_kernel (pysparse/precon/src/preconmodule.c:95-193) is
// much shorter than that: one row update (a handful of flops and a division) per level.
//
// Here every THREAD owns a chain of rows r0 -> r1 -> ... in which each row depends on its
// predecessor (so that value stays in a register), chains are grouped 32 to a warp in the order
// of their first row, and a warp advances all its chains together, one row per trip:
//   * a dependency produced by a lane of the SAME warp in the previous trip is read with a
//     shuffle from that lane's "last result" register (no memory at all) -- on a lexicographic
//     grid that is the left neighbour, and a warp becomes a strip of 32 grid columns swept as
//     a skewed wavefront;
//   * every other dependency is read from the result vector in global memory, which is
//     pre-filled with a NaN sentinel: a value is ready when it is not the sentinel (8-byte
//     stores are atomic, so no flags and no fences).  Producers run ahead of consumers, a
//     lagging warp polls -- nobody ever waits at a barrier.  All warps are co-resident
//     (cooperative launch), a lane that cannot proceed just skips the trip, and every lane
//     takes its rows in ascending (sweep) order, so the earliest unfinished row can always
//     proceed: no deadlock.
// The arithmetic of a row is the reference's, operation by operation (shared with the
// level-scheduled engines: sweep_row), so the results are bit-identical to it.
#pragma once
#include "psp_common.cuh"

namespace psp {

constexpr unsigned long long kSweepSentinel = 0xFFF8DEADBEEF0001ull;   // a quiet NaN no computation produces
constexpr int kP2PMaxDeps = 8;

struct P2PDir {                 // one sweep direction, device arrays
  const int *wbase;             // [nwarps + 1] first slot of every warp (slot = wbase[w] + position in the chain)
  const void *rec;              // P2PRec<W> of (slot, lane) at index slot * 32 + lane; row -1: padding
  int nwarps, D;
};

struct P2PArgs {
  P2PDir dir;
  const double *b;              // right-hand side
  const double *xold;           // result of the previous half sweep (read by the relaxed variant only)
  double *xnew;                 // result of this half sweep, pre-filled with the sentinel
  double *w;                    // L x / L^T x of the previous half sweep, overwritten
  double omega;
  int first;                    // forward sweep of the first step: zero initial guess (w, xold not read)
  int reverse;                  // backward sweep: rows are taken in descending order
  int n;
  int *abort_flag;              // set when a dependency never arrived (sentinel in the input?)
  const State *S;               // optional gate
};

// Result values travel between SMs through L2: relaxed gpu-scope accesses (L1 is bypassed; no
// fence is needed because a value is its own flag).  (ld/st.volatile would be system scope:
// three times the latency.)
__device__ __forceinline__ double ld_result(const double *p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_result(double *p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ bool is_sentinel(double v) {
  return (unsigned long long)__double_as_longlong(v) == kSweepSentinel;
}
__device__ __forceinline__ void prefetch_l2_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Operand pipeline.  A single warp issues one instruction every few cycles, so the instructions
// of a trip ARE its latency: the CTA therefore has two warps.  Warp 0 (compute) only gathers the
// dependencies and updates the rows.  Warp 1 (loader) keeps a ring of kP2PRing row slots per
// lane in shared memory filled with cp.async, lane l serving compute lane l: stage A, kP2PAhead
// slots ahead of stage B, streams the packed schedule record (row, diagonal, dependencies);
// stage B uses the row index stage A brought to gather b[r], w[r], the old x[r] and to read the
// out-of-warp dependencies early.  The two lanes talk through shared memory: `progress` (rows the
// compute lane has finished) throttles the loader, `filled` (slots whose stage B has landed)
// releases the compute lane.
constexpr unsigned kP2PSpinLimit = 1u << 21;     // idle trips before a warp gives up (~0.2 s): a dependency that never arrives
constexpr int kP2PAhead = 8;
constexpr int kP2PRing = 32;
constexpr int kP2PHist = 16;

template <int D>
struct P2PRec {                 // schedule record of one (slot, lane): 16 + 16 D bytes, 16-byte aligned
  int row, pad;
  double diag;
  int2 dep[D];                  // (row of the dependency or -1, producing lane of this warp or -1)
  double dval[D];
};

template <int D>
struct P2PSmem {
  P2PRec<D> rec[kP2PRing][32];
  double b[kP2PRing][32], w[kP2PRing][32], xo[kP2PRing][32];
  double2 pre[kP2PRing][D][32];  // early read of the out-of-warp dependencies (aligned pair holding the value)
  double hist_x[kP2PHist][32];   // the last kP2PHist results of every lane and the rows they belong to:
  int hist_row[kP2PHist][32];    // in-warp dependencies are read from here (position in the chain % kP2PHist)
  int progress[32], filled[32];
};

__device__ __forceinline__ void cpa8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16_l2(void *dst, const void *src) {     // .cg: L2 only, never a stale L1 line
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int lds_volatile(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ void sts_volatile(int *p, int v) { *(volatile int *)p = v; }

// RELAX = false: symgs_kernel (preconmodule.c:149-193); true: ssor_kernel (:95-143)
template <bool RELAX, int D>
__global__ void __launch_bounds__(64) k_sweep_p2p(P2PArgs a) {
  if (a.S && *(const volatile int *)&a.S->done) return;
  extern __shared__ __align__(16) unsigned char p2p_smem_raw[];
  P2PSmem<D> &sm = *reinterpret_cast<P2PSmem<D> *>(p2p_smem_raw);
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x;
  const int s0 = __ldg(a.dir.wbase + w), s1 = __ldg(a.dir.wbase + w + 1);
  const int nslot = s1 - s0;
  const bool need_w = !a.first;
  const bool need_xo = RELAX && !a.first;
  const double sentinel = __longlong_as_double((long long)kSweepSentinel);
  const P2PRec<D> *recs = reinterpret_cast<const P2PRec<D> *>(a.dir.rec);
  if (threadIdx.x < 32) {
    sm.progress[lane] = 0; sm.filled[lane] = 0;
#pragma unroll
    for (int h = 0; h < kP2PHist; h++) sm.hist_row[h][lane] = -1;
  }
  __syncthreads();

  if (threadIdx.x >= 32) {
    // ================================================================ loader warp
    // step j of a lane: wait for the group that holds stage A of slot j, issue stage A of slot
    // j + kP2PAhead and stage B of slot j as one group, publish the slots whose group is complete
    int j = 0;                                     // next step (slot index relative to s0)
    auto stage_a = [&](int q) {
      if (q < nslot) {
        const char *src = reinterpret_cast<const char *>(recs + ((size_t)(s0 + q) * 32 + lane));
        char *dst = reinterpret_cast<char *>(&sm.rec[q % kP2PRing][lane]);
#pragma unroll
        for (int c = 0; c < (int)sizeof(P2PRec<D>); c += 16) cpa16(dst + c, src + c);
      }
    };
    for (int q = 0; q < kP2PAhead; q++) stage_a(q);
    cpa_commit();
    for (int q = 1; q < kP2PAhead; q++) cpa_commit();        // so that the group count is uniform from step 0 on
    unsigned spins = 0;
    while (true) {
      const bool more = j < nslot;
      if (__ballot_sync(0xffffffffu, more) == 0u) break;
      // ring space: slot j + kP2PAhead overwrites slot j + kP2PAhead - kP2PRing, which must be finished
      const bool room = more && (j + kP2PAhead - kP2PRing < lds_volatile(&sm.progress[lane]));
      if (room) {
        cpa_wait<kP2PAhead - 1>();                 // the group with stage A of slot j is complete
        const int q = j % kP2PRing;
        const P2PRec<D> &rc = sm.rec[q][lane];
        const int r = rc.row;
        if (r < 0) {
          // end of this lane's chain (the rest of its slots is padding): release everything
          cpa_wait<0>();
          __threadfence_block();
          sts_volatile(&sm.filled[lane], nslot);
          j = nslot;
        } else {
          cpa8(&sm.b[q][lane], a.b + r);
          if (need_w) cpa8(&sm.w[q][lane], a.w + r);
          if (need_xo) cpa8(&sm.xo[q][lane], a.xold + r);
#pragma unroll
          for (int k = 0; k < D; k++) {
            // out-of-warp dependencies, read early; if the producer is not that far ahead the
            // sentinel comes back and the compute lane polls when the row is due
            const int2 cs = rc.dep[k];
            if (cs.x >= 0 && cs.y < 0) {
              if ((cs.x | 1) < a.n) cpa16_l2(&sm.pre[q][k][lane], a.xnew + (cs.x & ~1));
              else sm.pre[q][k][lane] = make_double2(sentinel, sentinel);    // last element of an odd-length vector
            }
          }
        }
        if (r >= 0) {
          stage_a(j + kP2PAhead);
          cpa_commit();
          j += 1;
          // groups complete in order: after this wait everything up to step j - kP2PAhead has landed
          cpa_wait<kP2PAhead - 1>();
          if (j == nslot) cpa_wait<0>();           // last step of this lane: nothing more to overlap with
          const int ok = (j == nslot) ? nslot : j - (kP2PAhead - 1);
          if (ok > 0) { __threadfence_block(); sts_volatile(&sm.filled[lane], ok); }
        }
      }
      if (__ballot_sync(0xffffffffu, room) == 0u) {
        __nanosleep(200);                          // nothing to do: leave the issue slots and the shared-memory pipe to the compute warp
        if (++spins > kP2PSpinLimit || ((spins & 1023u) == 0u && *(volatile int *)a.abort_flag)) {
#ifdef PSP_P2P_DEBUG
          printf("loader w%d lane%d gave up: j=%d nslot=%d progress=%d filled=%d\n", w, lane, j, nslot, sm.progress[lane], sm.filled[lane]);
#endif
          break;
        }
      } else spins = 0;
    }
    cpa_wait<0>();
    __threadfence_block();
    sts_volatile(&sm.filled[lane], nslot);
    return;
  }

  // ================================================================== compute warp
  const double om = a.omega, om1 = 1.0 - a.omega;
  int r = -1, col[D], src[D];                     // src: (position in the producer's chain << 5) | producing lane, or -1
  double dg = 1.0, bv = 0.0, wv = 0.0, xo = 0.0, val[D], pre[D];
  int q = 0;                                       // current slot (relative); nslot: chain finished
  bool have = false;                               // operands of slot q are in registers
  unsigned idle = 0;

  while (true) {
    // ---- take the current row's operands out of the ring once the loader has them
    if (!have && q < nslot && lds_volatile(&sm.filled[lane]) > q) {
      __threadfence_block();
      const int qs = q % kP2PRing;
      const P2PRec<D> &rc = sm.rec[qs][lane];
      r = rc.row;
      if (r >= 0) {
        dg = rc.diag; bv = sm.b[qs][lane];
        wv = need_w ? sm.w[qs][lane] : 0.0;
        xo = need_xo ? sm.xo[qs][lane] : 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) {
          const int2 cs = rc.dep[k];
          col[k] = cs.x; src[k] = cs.y; val[k] = rc.dval[k];
          const double2 pv = sm.pre[qs][k][lane];
          pre[k] = (cs.x >= 0 && cs.y < 0) ? ((cs.x & 1) ? pv.y : pv.x) : sentinel;
          // out-of-warp value that the early read missed: ask again right away rather than a trip later
          if (cs.x >= 0 && cs.y < 0 && is_sentinel(pre[k])) pre[k] = ld_result(a.xnew + cs.x);
        }
        have = true;
      } else {                                     // padding: this lane's chain is finished
        q = nslot;
        sts_volatile(&sm.progress[lane], nslot);   // (its loader lane stops waiting for ring space)
      }
    }
    const bool active = q < nslot;
    if (__ballot_sync(0xffffffffu, active) == 0u) break;
    // ---- gather this row's dependencies.  Produced by a lane of this warp: from that lane's
    // history ring in shared memory (tagged with the row, so a slot that is not written yet or
    // already overwritten is recognised); a producer more than kP2PHist rows ahead has the
    // value in memory.  Anything else: the early read, or the poll of the previous trip.
    bool ready = have;
    double xv[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
      xv[k] = 0.0;
      if (have && col[k] >= 0) {
        if (src[k] >= 0) {
          const int sl = src[k] & 31, sp = src[k] >> 5;
          const int hrow = *(volatile int *)&sm.hist_row[sp % kP2PHist][sl];
          if (hrow == col[k]) { xv[k] = *(volatile double *)&sm.hist_x[sp % kP2PHist][sl]; }
          else {
            ready = false;
            // overwritten already (the producer is past position sp): read it from memory from now on
            if (lds_volatile(&sm.progress[sl]) > sp) { src[k] = -1; pre[k] = sentinel; }
          }
        } else if (!is_sentinel(pre[k])) xv[k] = pre[k];
        else ready = false;
      }
    }
    // ---- the row update (same operations, same order as sweep_row in psp_precon.cu)
    if (ready) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++)
        if (col[k] >= 0) {
          const double pr = __dmul_rn(val[k], xv[k]);
          s = RELAX ? __dsub_rn(s, pr) : __dadd_rn(s, pr);
        }
      double xn, wn;
      if (!RELAX) {
        xn = __ddiv_rn(__dsub_rn(__dsub_rn(bv, wv), s), dg);
        wn = s;
      } else {
        double temp;
        if (!need_xo) temp = __dmul_rn(om, bv);
        else temp = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(om1, xo), dg), wv), __dmul_rn(om, bv));
        wn = __dmul_rn(om, s);
        xn = __ddiv_rn(__dadd_rn(temp, wn), dg);
      }
      st_result(a.xnew + r, xn);
      a.w[r] = wn;
      // history: value first, then the tag that validates it (same-warp readers follow in program order)
      *(volatile double *)&sm.hist_x[q % kP2PHist][lane] = xn;
      *(volatile int *)&sm.hist_row[q % kP2PHist][lane] = r;
      q += 1;
      have = false;
      sts_volatile(&sm.progress[lane], q);
    }
    const unsigned adv = __ballot_sync(0xffffffffu, ready);
    // ---- lanes waiting for a value that comes through memory: one poll per trip (never a
    // spin: another lane of this warp may be what the producer is waiting for)
    if (have && !ready) {
#pragma unroll
      for (int k = 0; k < D; k++)
        if (col[k] >= 0 && src[k] < 0 && is_sentinel(pre[k])) pre[k] = ld_result(a.xnew + col[k]);
    }
    if (adv == 0u) {                               // the whole warp is waiting (for other warps or its loader)
      if (++idle > kP2PSpinLimit) {
#ifdef PSP_P2P_DEBUG
        printf("compute w%d lane%d gave up: q=%d nslot=%d have=%d r=%d filled=%d col0=%d src0=%d\n", w, lane, q, nslot, (int)have, r, sm.filled[lane], col[0], src[0]);
#endif
        if (lane == 0) atomicExch(a.abort_flag, 1);
        break;
      }
      if ((idle & 1023u) == 0u && *(volatile int *)a.abort_flag) break;
    } else idle = 0;
  }
}

template <int D>
constexpr size_t p2p_smem_bytes() { return sizeof(P2PSmem<D>); }
template <int D>
constexpr size_t p2p_rec_bytes() { return sizeof(P2PRec<D>); }

}  // namespace psp
