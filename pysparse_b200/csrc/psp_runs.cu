// psp_runs.cu -- analysis, upload and stand-alone kernel of the "diagonal run" format
// (psp_runs.cuh has the format and the device pass).
//
// Replaces, for matrices whose rows line up on diagonals, the inner loops of
// csr_matvec_kernel (pysparse/sparse/src/csr_mat.c:49-54) and sss_matvec (sss_mat.c:40-56).
// Algorithmic traffic of the reference layouts is 12 bytes per stored entry (value + column
// index); this format moves 8 (the value) plus 16 bytes per run.
#include <algorithm>
#include "psp_mat.cuh"
#include "psp_epilogue.cuh"
#include "psp_runs.cuh"

namespace psp {

// ------------------------------------------------------------------ kernel
struct RunArgs {
  RunView A;
  const double *x;
  double *y;
  const double *w1, *w2;
  int w2_is_y, epilogue, gate;
  State *S;
  double *partials;
  unsigned *ticket;
  void *aux;
};

template <bool DOTS>
__global__ void __launch_bounds__(kThreads, 4) k_spmv_runs(RunArgs a) {
  __shared__ double red_sh[8];
  __shared__ RunSmem sseg;
  if (DOTS) {
    const volatile State *V = a.S;
    if (V->done || (a.gate == 1 && V->cyc_done)) return;
  }
  double acc0 = 0.0, acc1 = 0.0;
  run_pass<true, 8, false>(a.A, a.x, sseg, [&](int r, double sum) {
    a.y[r] = sum;
    if (DOTS) {
      if (a.w1) acc0 += a.w1[r] * sum;
      if (a.w2_is_y) acc1 += sum * sum;
      else if (a.w2) acc1 += a.w2[r] * sum;
    }
  }, [](int) {});
  if (DOTS) {
    double mine[2];
    mine[0] = block_sum(acc0, red_sh);
    mine[1] = block_sum(acc1, red_sh);
    double tot[2], dummy;
    if (grid_reduce_finish<2, false>(mine, 0.0, a.partials, a.ticket, tot, &dummy, red_sh, a.S->dist)) {
      if (threadIdx.x == 0) {
        a.S->red[0] = tot[0];
        a.S->red[1] = tot[1];
        run_epilogue(a.epilogue, a.S, a.aux);
      }
    }
  }
}

int runs_launch(const psp_mat *A, const double *dx, double *dy, const SpmvDots &d) {
  Ctx &c = ctx();
  RunArgs a;
  a.A.val = A->d_rval; a.A.seg = reinterpret_cast<const int4 *>(A->d_rseg);
  a.A.blk = reinterpret_cast<const int4 *>(A->d_rblk); a.A.nb = A->nb; a.A.nint = A->nb; a.A.blk_halo = nullptr;
  a.x = dx; a.y = dy;
  a.w1 = d.w1; a.w2 = d.w2; a.w2_is_y = d.w2_is_y; a.epilogue = d.epilogue; a.gate = d.gate; a.S = d.S;
  a.partials = c.partials; a.ticket = c.tickets;
  a.aux = d.S ? (void *)(d.S + 1) : nullptr;
  // persistent grid: what is resident at once (the DOTS variant needs it for its ticket
  // reduction buffers, the plain one simply has no tail of tiny waves)
  int &res = c.occ_runs[d.S ? 1 : 0];
  if (res == 0) {
    int per_sm = 0;
    const void *fn = d.S ? (const void *)k_spmv_runs<true> : (const void *)k_spmv_runs<false>;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    res = per_sm * c.sms;
    if (res > kMaxGrid) res = kMaxGrid;
  }
  a.A.chunk = run_chunk(A->nb, res);
  const int nch = (A->nb + a.A.chunk - 1) / a.A.chunk;
  const int grid = nch < res ? nch : res;
  if (d.S) k_spmv_runs<true><<<grid, kThreads, 0, c.stream>>>(a);
  else k_spmv_runs<false><<<grid, kThreads, 0, c.stream>>>(a);
  PSP_LAUNCH_CHECK();
  return PSP_OK;
}

// ------------------------------------------------------------------ host analysis
namespace {

struct Seg { int key; int t0, len; int x0; int kpos; };     // key: (global) diagonal offset = sort key

struct RunBuild {
  std::vector<int> blk;        // 4 ints per block, natural order
  std::vector<int> seg;        // 4 ints per run
  std::vector<double> val;     // run values; kRunSkip in the bridged gaps
  long long diag_off = 0;      // SSS: position of the diagonal in val
};

double skip_value() { double v; const unsigned long long b = kRunSkip; memcpy(&v, &b, 8); return v; }

// Runs of the rows [r0, r1) of a CSR structure.  A run continues to row t when the row has an
// entry on the same (global) diagonal whose local column continues the run -- directly, or
// (bridge) after at most kRunGapMax rows without one, once the run is kRunGapMinLen rows long.
// entry_seg[k - ind[r0]] receives the block-local run number of entry k.
void scan_block(const int *ind, const int *col, const int *gcol, int grow0, int r0, int r1, bool bridge,
                std::vector<Seg> &segs, std::vector<int> *entry_seg) {
  struct Open { int key, id; };
  static thread_local std::vector<Open> open, keep;     // runs that ended at most kRunGapMax rows ago, by key
  open.clear();
  segs.clear();
  const int k0 = ind[r0];
  if (entry_seg) entry_seg->resize((size_t)(ind[r1] - k0));
  for (int r = r0; r < r1; r++) {
    const int t = r - r0;
    keep.clear();
    size_t p = 0;
    for (int k = ind[r]; k < ind[r + 1]; k++) {
      const int key = (gcol ? gcol[k] : col[k]) - (grow0 + r);
      while (p < open.size() && open[p].key < key) {
        const Seg &o = segs[(size_t)open[p].id];
        if (bridge && t - (o.t0 + o.len) < kRunGapMax) keep.push_back(open[p]);     // may still be continued by a later row
        p++;
      }
      int id = -1;
      if (p < open.size() && open[p].key == key) {
        Seg &s = segs[(size_t)open[p].id];
        const int gap = t - (s.t0 + s.len);
        if (s.x0 + s.len + gap == col[k] && (gap == 0 || (bridge && gap <= kRunGapMax && s.len >= kRunGapMinLen))) {
          id = open[p].id;
          s.len += gap + 1;
        }
        p++;
      }
      if (id < 0) {
        id = (int)segs.size();
        segs.push_back(Seg{key, t, 1, col[k], 0});
      }
      if (entry_seg) (*entry_seg)[(size_t)(k - k0)] = id;
      keep.push_back(Open{key, id});
    }
    for (; p < open.size(); p++) {
      const Seg &o = segs[(size_t)open[p].id];
      if (bridge && t - (o.t0 + o.len) < kRunGapMax) keep.push_back(open[p]);
    }
    open.swap(keep);
  }
}

// no stored value may look like the gap marker
bool can_bridge(const double *val, long long nnz, const double *diag, int n) {
  for (long long k = 0; k < nnz; k++)
    if (run_skip(val[k])) return false;
  if (diag)
    for (int i = 0; i < n; i++)
      if (run_skip(diag[i])) return false;
  return true;
}

// cheap look at a sample of blocks: is the format worth a full analysis?
bool sample_accepts(int m, const int *ind, const int *col, const int *gcol, int grow0) {
  const int nb = (m + kThreads - 1) / kThreads;
  if (nb == 0) return false;
  std::vector<Seg> segs;
  long long nseg = 0, nent = 0;
  const int nsample = nb < 48 ? nb : 48;
  for (int i = 0; i < nsample; i++) {
    const int b = (int)((long long)i * nb / nsample);
    const int r0 = b * kThreads, r1 = std::min(m, r0 + kThreads);
    if ((long long)ind[r1] - ind[r0] > 64 * kThreads) return false;    // long rows: not this format's case
    scan_block(ind, col, gcol, grow0, r0, r1, true, segs, nullptr);
    nseg += (long long)segs.size();
    nent += ind[r1] - ind[r0];
  }
  return nent > 0 && nseg * 8 <= nent;
}

int upload_raw(const void *src, size_t bytes, void **dst) {
  void *p = nullptr;
  const size_t pad = 64;
  cudaError_t e = cudaMalloc(&p, bytes + pad);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes + pad, cudaGetErrorString(e));
    return PSP_ERR_NOMEM;
  }
  if (cudaMemset(p, 0, bytes + pad) != cudaSuccess || (bytes && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess)) {
    cudaGetLastError();
    cudaFree(p);
    set_error("upload of %zu bytes failed", bytes);
    return PSP_ERR_CUDA;
  }
  *dst = p;
  return PSP_OK;
}

// own_cols >= 0: blocks that read no halo column first (their number -> *nint), and per
// scheduled block the range of halo positions it reads (bh, 2 ints per block)
int upload_runs(psp_mat *A, RunBuild &rb, int own_cols) {
  const int nb = (int)(rb.blk.size() / 4);
  std::vector<int> sched;
  std::vector<int> bh;
  int nint = nb;
  if (own_cols >= 0) {
    std::vector<int> interior, boundary;
    std::vector<int> lo((size_t)nb, INT32_MAX), hi((size_t)nb, -1);
    for (int b = 0; b < nb; b++) {
      for (int s = rb.blk[4 * (size_t)b + 2]; s < rb.blk[4 * (size_t)b + 3]; s++) {
        const int x0 = rb.seg[4 * (size_t)s + 1], len = rb.seg[4 * (size_t)s + 3];
        if (x0 + len > own_cols) {
          const int a0 = (x0 > own_cols ? x0 : own_cols) - own_cols, a1 = x0 + len - own_cols;
          if (a0 < lo[b]) lo[b] = a0;
          if (a1 > hi[b]) hi[b] = a1;
        }
      }
      (hi[b] >= 0 ? boundary : interior).push_back(b);
    }
    nint = (int)interior.size();
    sched.reserve((size_t)nb * 4);
    bh.assign(2 * (size_t)nb, 0);
    // the run descriptors follow the schedule: consecutive blocks of the schedule have
    // consecutive descriptor ranges (run_pass fetches them chunk-wise)
    std::vector<int> seg2;
    seg2.reserve(rb.seg.size());
    int i = 0;
    for (const std::vector<int> *lst : {&interior, &boundary})
      for (int b : *lst) {
        const int s0 = rb.blk[4 * (size_t)b + 2], s1 = rb.blk[4 * (size_t)b + 3];
        sched.push_back(rb.blk[4 * (size_t)b]); sched.push_back(rb.blk[4 * (size_t)b + 1]);
        sched.push_back((int)(seg2.size() / 4));
        seg2.insert(seg2.end(), rb.seg.begin() + 4 * (size_t)s0, rb.seg.begin() + 4 * (size_t)s1);
        sched.push_back((int)(seg2.size() / 4));
        if (hi[b] >= 0) { bh[2 * (size_t)i] = lo[b]; bh[2 * (size_t)i + 1] = hi[b]; }
        i++;
      }
    rb.seg.swap(seg2);
  }
  const std::vector<int> &blk = own_cols >= 0 ? sched : rb.blk;
  PSP_TRY(upload_raw(rb.val.data(), sizeof(double) * rb.val.size(), (void **)&A->d_rval));
  PSP_TRY(upload_raw(rb.seg.data(), sizeof(int) * rb.seg.size(), (void **)&A->d_rseg));
  PSP_TRY(upload_raw(blk.data(), sizeof(int) * blk.size(), (void **)&A->d_rblk));
  if (own_cols >= 0) PSP_TRY(upload_raw(bh.data(), sizeof(int) * bh.size(), (void **)&A->d_blk_halo));
  A->fmt = 1;
  A->nb = nb; A->nint = nint;
  A->rval_len = (long long)rb.val.size();
  A->nseg = (int)(rb.seg.size() / 4);
  return PSP_OK;
}

bool by_key_then_row(const Seg &a, const Seg &b) { return a.key != b.key ? a.key < b.key : a.t0 < b.t0; }

}  // namespace

// A block whose runs cover a bridged gap is flagged (bit kRunGapFlag of its row count): only there
// does the kernel test the loaded values for the marker.
static void flag_gap_blocks(RunBuild &rb) {
  const size_t nb = rb.blk.size() / 4;
  const unsigned long long *bits = reinterpret_cast<const unsigned long long *>(rb.val.data());
  for (size_t b = 0; b < nb; b++) {
    bool gap = false;
    for (int sgi = rb.blk[4 * b + 2]; sgi < rb.blk[4 * b + 3] && !gap; sgi++) {
      const size_t k0 = (size_t)rb.seg[4 * (size_t)sgi], len = (size_t)rb.seg[4 * (size_t)sgi + 3];
      // (a lower run never starts or ends on a gap, but a mirrored slice of it can: the slices are
      // cut at the target block's boundaries)
      for (size_t u = 0; u < len; u++)
        if (bits[k0 + u] == kRunSkip) { gap = true; break; }
    }
    if (gap) rb.blk[4 * b + 1] |= kRunGapFlag;
  }
}

// Automatic choice: the run kernels win once the matrix stream is far beyond what L2 and the
// latency of one pass can hide (measured on 3-D Poisson, PCG us/iteration, runs vs row blocks:
// 192^3 255 vs 238, 256^3 493 vs 545; SpMV 256^3 228 vs 288 us) -- below ~1 GB of CSR stream
// the TMA-fed row blocks stay the default.  `bytes`: 12 bytes per entry of the (expanded) matrix.
static bool runs_pay_off(long long bytes) { return bytes >= (1LL << 30); }

// host-only analysis of a CSR structure; `force`: build whatever the run lengths are
static bool build_csr_runs(int m, const int *ind, const int *col, const double *val, const int *gcol, int grow0,
                           bool force, RunBuild &rb) {
  const long long nnz = ind[m];
  if (m == 0 || nnz == 0) return false;
  if (!force && !sample_accepts(m, ind, col, gcol, grow0)) return false;
  const int nb = (m + kThreads - 1) / kThreads;
  const bool bridge = can_bridge(val, nnz, nullptr, 0);
  const double skip = skip_value();
  rb.blk.resize(4 * (size_t)nb);
  rb.val.reserve((size_t)nnz + (size_t)nnz / 64);
  std::vector<Seg> segs;
  std::vector<int> entry_seg, order;
  long long kpos = 0;
  for (int b = 0; b < nb; b++) {
    const int r0 = b * kThreads, r1 = std::min(m, r0 + kThreads);
    scan_block(ind, col, gcol, grow0, r0, r1, bridge, segs, &entry_seg);
    order.resize(segs.size());
    for (size_t i = 0; i < segs.size(); i++) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return by_key_then_row(segs[(size_t)x], segs[(size_t)y]); });
    rb.blk[4 * (size_t)b] = r0; rb.blk[4 * (size_t)b + 1] = r1 - r0;
    rb.blk[4 * (size_t)b + 2] = (int)(rb.seg.size() / 4);
    for (int id : order) {
      Seg &s = segs[(size_t)id];
      if (kpos + s.len >= (1LL << 31)) return false;
      s.kpos = (int)kpos;
      kpos += s.len;
      rb.seg.push_back(s.kpos); rb.seg.push_back(s.x0); rb.seg.push_back(s.t0); rb.seg.push_back(s.len);
    }
    rb.blk[4 * (size_t)b + 3] = (int)(rb.seg.size() / 4);
    rb.val.resize((size_t)kpos, skip);               // bridged gaps keep the marker
    const int k0 = ind[r0];
    for (int r = r0; r < r1; r++)
      for (int k = ind[r]; k < ind[r + 1]; k++) {
        const Seg &s = segs[(size_t)entry_seg[(size_t)(k - k0)]];
        rb.val[(size_t)s.kpos + (size_t)(r - r0 - s.t0)] = val[k];
      }
    // give up early on matrices the sample misjudged
    if (!force && (b & 1023) == 1023 && (long long)(rb.seg.size() / 4) * 8 > (long long)ind[r1]) return false;
  }
  if (!force && (long long)(rb.seg.size() / 4) * 16 > nnz) return false;     // mean run shorter than 16 rows
  flag_gap_blocks(rb);
  return true;
}

int runs_try_csr(psp_mat *A, const int *ind, const int *col, const double *val, const int *gcol, int grow0,
                 int own_cols, bool *accepted) {
  *accepted = false;
  Ctx &c = ctx();
  if (c.spmv_format == 0) return PSP_OK;
  if (c.spmv_format < 0 && !runs_pay_off(12LL * ind[A->m])) return PSP_OK;
  RunBuild rb;
  if (!build_csr_runs(A->m, ind, col, val, gcol, grow0, c.spmv_format > 0, rb)) return PSP_OK;
  PSP_TRY(upload_runs(A, rb, own_cols));
  *accepted = true;
  return PSP_OK;
}

// SSS: the strict lower triangle is scanned block by block.  Two value layouts:
//  * block-major (global_major == false): a block's lower runs are stored back to back; the
//    mirrored (upper) runs of a target block are slices of the source blocks' runs, so a run that
//    crosses a source-block boundary inside the target block gives two descriptors;
//  * global-run-major: runs that continue across a block boundary are chained into GLOBAL runs and
//    the values are stored global-run-major (sorted by diagonal, then row); lower and mirrored
//    runs of a block are slices of the global runs -- ONE descriptor per diagonal for an interior
//    block of a stencil matrix however the grid lines fall into the blocks.
// over_budget: number of blocks that need more than one batch (8 descriptors) in run_pass.
static bool build_sss_runs_layout(int n, const int *ind, const int *col, const double *val, const double *diag, bool force,
                                  bool global_major, RunBuild &rb, long long *over_budget) {
  const long long nnz = ind[n];
  const int nb = (n + kThreads - 1) / kThreads;
  const bool bridge = can_bridge(val, nnz, diag, n);
  const double skip = skip_value();
  struct GRun { int key, row0, len; long long pos; };
  struct LSeg { Seg s; int grun; };                  // s.kpos is filled in once the positions are known
  std::vector<GRun> gruns;
  std::vector<std::vector<LSeg>> lower((size_t)nb);
  std::vector<Seg> segs;
  struct Tail { int key, grun, xend; };              // runs that reach the last row of the previous block
  std::vector<Tail> tails, tails_next;
  // pass 1: structure
  for (int b = 0; b < nb; b++) {
    const int r0 = b * kThreads, r1 = std::min(n, r0 + kThreads);
    scan_block(ind, col, nullptr, 0, r0, r1, bridge, segs, nullptr);
    tails_next.clear();
    for (const Seg &s : segs) {
      int g = -1;
      if (global_major && s.t0 == 0)
        for (const Tail &t : tails)
          if (t.key == s.key && t.xend == s.x0) { g = t.grun; break; }
      if (g < 0) {
        g = (int)gruns.size();
        gruns.push_back(GRun{s.key, r0 + s.t0, 0, 0});
      }
      gruns[(size_t)g].len += s.len;
      lower[(size_t)b].push_back(LSeg{s, g});
      if (s.t0 + s.len == r1 - r0) tails_next.push_back(Tail{s.key, g, s.x0 + s.len});
    }
    tails.swap(tails_next);
  }
  // positions: global-run-major by (diagonal, row); block-major by (block, diagonal, row) --
  // without chaining every run is its own "global" run and starts in exactly one block
  {
    std::vector<int> order(gruns.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
      const GRun &a = gruns[(size_t)x], &c = gruns[(size_t)y];
      if (!global_major && a.row0 / kThreads != c.row0 / kThreads) return a.row0 < c.row0;
      return a.key != c.key ? a.key < c.key : a.row0 < c.row0;
    });
    long long pos = 0;
    for (int id : order) { gruns[(size_t)id].pos = pos; pos += gruns[(size_t)id].len; }
    if (pos + n >= (1LL << 31)) return false;
    rb.diag_off = pos;
    rb.val.assign((size_t)pos + (size_t)n, skip);    // bridged gaps keep the marker
    std::copy(diag, diag + n, rb.val.begin() + pos);
  }
  // pass 2: values (the scan is deterministic: same runs, same numbering) and the mirrored slices
  struct Upper { int tb; Seg s; };
  std::vector<Upper> upper;
  std::vector<int> entry_seg;
  for (int b = 0; b < nb; b++) {
    const int r0 = b * kThreads, r1 = std::min(n, r0 + kThreads);
    scan_block(ind, col, nullptr, 0, r0, r1, bridge, segs, &entry_seg);
    std::vector<LSeg> &ls = lower[(size_t)b];
    for (LSeg &l : ls) {
      const GRun &g = gruns[(size_t)l.grun];
      l.s.kpos = (int)(g.pos + (r0 + l.s.t0 - g.row0));
    }
    const int k0 = ind[r0];
    for (int r = r0; r < r1; r++)
      for (int k = ind[r]; k < ind[r + 1]; k++) {
        const Seg &sg = ls[(size_t)entry_seg[(size_t)(k - k0)]].s;
        rb.val[(size_t)sg.kpos + (size_t)(r - r0 - sg.t0)] = val[k];
      }
    std::sort(ls.begin(), ls.end(), [](const LSeg &a, const LSeg &c) { return by_key_then_row(a.s, c.s); });
  }
  for (const GRun &g : gruns) {
    // entry (i, j) adds val * x[i] to y[j]: rows i0 .. i0+len-1 of the run are the columns of the
    // target rows j0 .. j0+len-1, j0 = i0 + key (a bridged gap travels along: both uses skip it)
    int i0 = g.row0, j0 = g.row0 + g.key, left = g.len;
    long long kp = g.pos;
    while (left > 0) {
      const int tb = j0 / kThreads, tt = j0 % kThreads;
      const int piece = std::min(left, kThreads - tt);
      upper.push_back(Upper{tb, Seg{-g.key, tt, piece, i0, (int)kp}});
      i0 += piece; j0 += piece; kp += piece; left -= piece;
    }
  }
  long long nseg = (long long)upper.size() * 2 + nb;
  if (!force && nnz > 0 && nseg * 16 > 2 * nnz + n) return false;
  // counting sort of the mirrored runs by target block, then by (diagonal, row) inside a block
  std::vector<int> start((size_t)nb + 1, 0);
  for (const Upper &u : upper) start[(size_t)u.tb + 1]++;
  for (int b = 0; b < nb; b++) start[(size_t)b + 1] += start[(size_t)b];
  std::vector<Seg> usorted(upper.size());
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (const Upper &u : upper) usorted[(size_t)fill[(size_t)u.tb]++] = u.s;
  }
  rb.blk.resize(4 * (size_t)nb);
  rb.seg.clear();
  *over_budget = 0;
  for (int b = 0; b < nb; b++) {
    const int r0 = b * kThreads, r1 = std::min(n, r0 + kThreads);
    std::sort(usorted.begin() + start[(size_t)b], usorted.begin() + start[(size_t)b + 1], by_key_then_row);
    rb.blk[4 * (size_t)b] = r0; rb.blk[4 * (size_t)b + 1] = r1 - r0;
    rb.blk[4 * (size_t)b + 2] = (int)(rb.seg.size() / 4);
    auto emit = [&](const Seg &s) { rb.seg.push_back(s.kpos); rb.seg.push_back(s.x0); rb.seg.push_back(s.t0); rb.seg.push_back(s.len); };
    for (const LSeg &l : lower[(size_t)b]) emit(l.s);
    emit(Seg{0, 0, r1 - r0, r0, (int)(rb.diag_off + r0)});            // the diagonal
    for (int i = start[(size_t)b]; i < start[(size_t)b + 1]; i++) emit(usorted[(size_t)i]);
    rb.blk[4 * (size_t)b + 3] = (int)(rb.seg.size() / 4);
    if (rb.blk[4 * (size_t)b + 3] - rb.blk[4 * (size_t)b + 2] > kRunBatch) (*over_budget)++;
  }
  flag_gap_blocks(rb);
  return true;
}

// Block-major values stream best (one contiguous region per block: 256^3 SSS product 186 us
// against 199 us global-run-major) as long as the blocks stay within one batch of descriptors;
// when the grid lines do not line up with the blocks (384^3: 9-10 descriptors, 804 us) the
// global-run-major layout brings them back to 7 (675 us).
static bool build_sss_runs(int n, const int *ind, const int *col, const double *val, const double *diag, bool force,
                           RunBuild &rb) {
  const long long nnz = ind[n];
  if (n == 0) return false;
  if (nnz > 0 && !force && !sample_accepts(n, ind, col, nullptr, 0)) return false;
  if (nnz + n >= (1LL << 31) - (nnz >> 6)) return false;
  const int nb = (n + kThreads - 1) / kThreads;
  long long over = 0;
  if (!build_sss_runs_layout(n, ind, col, val, diag, force, false, rb, &over)) return false;
  if (over * 8 <= nb) return true;
  RunBuild alt;
  long long over_alt = 0;
  if (build_sss_runs_layout(n, ind, col, val, diag, force, true, alt, &over_alt) && over_alt < over) {
    rb.blk.swap(alt.blk); rb.seg.swap(alt.seg); rb.val.swap(alt.val); rb.diag_off = alt.diag_off;
  }
  return true;
}

int runs_try_sss(psp_mat *A, const int *ind, const int *col, const double *val, const double *diag, bool *accepted) {
  *accepted = false;
  Ctx &c = ctx();
  if (c.spmv_format == 0) return PSP_OK;
  if (c.spmv_format < 0 && !runs_pay_off(12LL * (2LL * ind[A->n] + A->n))) return PSP_OK;
  RunBuild rb;
  if (!build_sss_runs(A->n, ind, col, val, diag, c.spmv_format > 0, rb)) return PSP_OK;
  PSP_TRY(upload_runs(A, rb, -1));
  A->d_diag = A->d_rval + rb.diag_off;  // the SSOR sweeps read the diagonal from here
  *accepted = true;
  return PSP_OK;
}

int runs_to_csr(const psp_mat *A, std::vector<int> &ind, std::vector<int> &col, std::vector<double> &val) {
  if (A->fmt != 1 || A->kind != 1) { set_error("runs_to_csr: not a CSR matrix in run format"); return PSP_ERR_ARG; }
  std::vector<int> blk(4 * (size_t)A->nb), seg(4 * (size_t)A->nseg);
  std::vector<double> rv((size_t)A->rval_len);
  PSP_CUDA(cudaMemcpy(blk.data(), A->d_rblk, sizeof(int) * blk.size(), cudaMemcpyDeviceToHost));
  PSP_CUDA(cudaMemcpy(seg.data(), A->d_rseg, sizeof(int) * seg.size(), cudaMemcpyDeviceToHost));
  PSP_CUDA(cudaMemcpy(rv.data(), A->d_rval, sizeof(double) * rv.size(), cudaMemcpyDeviceToHost));
  for (int b = 0; b < A->nb; b++) blk[4 * (size_t)b + 1] &= ~kRunGapFlag;
  const int m = A->m;
  ind.assign((size_t)m + 1, 0);
  for (int b = 0; b < A->nb; b++)
    for (int s = blk[4 * (size_t)b + 2]; s < blk[4 * (size_t)b + 3]; s++)
      for (int u = 0; u < seg[4 * (size_t)s + 3]; u++)
        if (!run_skip(rv[(size_t)seg[4 * (size_t)s] + (size_t)u])) ind[(size_t)(blk[4 * (size_t)b] + seg[4 * (size_t)s + 2] + u) + 1]++;
  for (int r = 0; r < m; r++) ind[(size_t)r + 1] += ind[(size_t)r];
  col.resize((size_t)ind[(size_t)m]); val.resize((size_t)ind[(size_t)m]);
  std::vector<int> fill(ind.begin(), ind.end() - 1);
  for (int b = 0; b < A->nb; b++)                   // runs of a block are in ascending-column order per row
    for (int s = blk[4 * (size_t)b + 2]; s < blk[4 * (size_t)b + 3]; s++)
      for (int u = 0; u < seg[4 * (size_t)s + 3]; u++) {
        if (run_skip(rv[(size_t)seg[4 * (size_t)s] + (size_t)u])) continue;       // bridged gap
        const int r = blk[4 * (size_t)b] + seg[4 * (size_t)s + 2] + u;
        const int p = fill[(size_t)r]++;
        col[(size_t)p] = seg[4 * (size_t)s + 1] + u;
        val[(size_t)p] = rv[(size_t)seg[4 * (size_t)s] + (size_t)u];
      }
  return PSP_OK;
}

}  // namespace psp

// ---------------------------------------------------------------- C-ABI: host-only analysis
// The structure work of the format (bit-exact integer work, no GPU needed; used by the CPU tests).
using namespace psp;

extern "C" int psp_runs_analyze(int sss, int m, const int *ind, const int *col, const double *val, const double *diag,
                                int *nblk, int *nseg, long long *nval, int **blk, int **seg, double **rval) {
  if (!ind || !nblk || !nseg || !nval || !blk || !seg || !rval) { set_error("psp_runs_analyze: NULL argument"); return PSP_ERR_ARG; }
  RunBuild rb;
  const bool ok = sss ? build_sss_runs(m, ind, col, val, diag, true, rb) : build_csr_runs(m, ind, col, val, nullptr, 0, true, rb);
  if (!ok) { *nblk = 0; *nseg = 0; *nval = 0; *blk = nullptr; *seg = nullptr; *rval = nullptr; return PSP_OK; }
  *nblk = (int)(rb.blk.size() / 4); *nseg = (int)(rb.seg.size() / 4); *nval = (long long)rb.val.size();
  *blk = (int *)malloc(sizeof(int) * (rb.blk.size() + 1));
  *seg = (int *)malloc(sizeof(int) * (rb.seg.size() + 1));
  *rval = (double *)malloc(sizeof(double) * (rb.val.size() + 1));
  if (!*blk || !*seg || !*rval) { free(*blk); free(*seg); free(*rval); set_error("psp_runs_analyze: out of memory"); return PSP_ERR_NOMEM; }
  memcpy(*blk, rb.blk.data(), sizeof(int) * rb.blk.size());
  memcpy(*seg, rb.seg.data(), sizeof(int) * rb.seg.size());
  memcpy(*rval, rb.val.data(), sizeof(double) * rb.val.size());
  return PSP_OK;
}

extern "C" void psp_runs_free(void *p) { free(p); }
