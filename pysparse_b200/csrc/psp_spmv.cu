// psp_spmv.cu -- fp64 CSR / SSS sparse matrix-vector product for sm_100a.
//
// Replaces csr_matvec_kernel / csr_matvec_transp_kernel (pysparse/sparse/src/
// csr_mat.c:17-106) and sss_matvec (sss_mat.c:40-76).
//
// Kernel design ("row-block streaming"):
//   * an analysis pass on the host bins consecutive rows into row blocks of
//     at most kBlkNnz stored entries and kThreads rows; a row longer than
//     kBlkNnz gets a block of its own;
//   * a persistent grid (a multiple of the SM count) walks the row blocks.
//     For a block, ALL threads stream the block's contiguous val/col range
//     with aligned 128-bit loads (fully coalesced, independent of the row
//     lengths), gather x through the read-only path, and park the products
//     in shared memory; then one thread per row adds its products up IN ROW
//     ORDER.  The row sums therefore have exactly the accumulation order of
//     the reference's serial loop, and since the products are rounded before
//     they are added (no FMA) short-row results are bit-identical to it;
//   * blocks whose rows are long on average are reduced warp-per-row, and a
//     single over-long row is reduced by the whole CTA (tree order; within
//     the 1e-13 contract);
//   * optional fused epilogue: up to two dot products with the result vector
//     (p.q for PCG, rhat.v / t.s / t.t for BiCGSTAB ...), reduced
//     deterministically, followed by the solver's scalar epilogue.
//
// Transposed product: a one-time transpose index (CSR of A^T, built on the
// host with a counting sort that keeps ascending row order) turns the
// reference's scatter loop into the same gather kernel -- deterministic, no
// fp64 atomics, and again the reference's accumulation order.
//
// SSS: the strict lower triangle is expanded once into the symmetric CSR in
// ll_mat.to_csr() order (lower row | diag | upper = transposed lower, rows
// ascending), which is exactly the order in which sss_matvec accumulates
// y[i]; the lower triangle + diagonal are kept as well for the SSOR sweeps.
#include <climits>
#include "psp_mat.cuh"
#include "psp_correq.cuh"
#include "psp_epilogue.cuh"
#include "psp_tma_ring.cuh"

namespace psp {

constexpr int kProdSlots = kBlkNnz + 4;

struct SpmvArgs {
  const int *ind, *col, *blk;
  const double *val, *x;
  double *y;
  int nb;
  const double *w1, *w2;
  int w2_is_y;
  int epilogue;
  int gate;
  int l2_stream;      // TMA loads of the matrix carry an L2 evict-first hint
  State *S;
  double *partials;
  unsigned *ticket;
  void *aux;
};

// VEC = 0: scalar 8/4-byte loads;  VEC = 1: aligned 128-bit loads (4 entries / thread / step)
template <int VEC, bool DOTS>
__global__ void __launch_bounds__(kThreads, 4) k_spmv_stream(SpmvArgs a) {
  __shared__ __align__(16) double prod[kProdSlots];
  __shared__ double red_sh[8];
  if (DOTS) {
    const volatile State *V = a.S;
    if (V->done || (a.gate == 1 && V->cyc_done)) return;
  }
  const int tid = threadIdx.x;
  double acc0 = 0.0, acc1 = 0.0;

  for (int b = blockIdx.x; b < a.nb; b += gridDim.x) {
    const int4 d = __ldg(reinterpret_cast<const int4 *>(a.blk) + b);
    const int r0 = d.x, r1 = d.z, k0 = d.y, k1 = d.w;
    const int nrows = r1 - r0, cnt = k1 - k0;

    if (cnt <= kBlkNnz) {
      int base;   // prod[j] holds the product of entry base + j
      if (VEC) {
        base = k0 & ~3;
        const int ngroups = (k1 - base + 3) >> 2;
        for (int g = tid; g < ngroups; g += kThreads) {
          const int k = base + 4 * g;
          const int4 c = __ldg(reinterpret_cast<const int4 *>(a.col + k));
          const double2 v01 = __ldg(reinterpret_cast<const double2 *>(a.val + k));
          const double2 v23 = __ldg(reinterpret_cast<const double2 *>(a.val + k + 2));
          double2 p01, p23;
          p01.x = __dmul_rn(v01.x, ldg_f64(a.x + c.x));
          p01.y = __dmul_rn(v01.y, ldg_f64(a.x + c.y));
          p23.x = __dmul_rn(v23.x, ldg_f64(a.x + c.z));
          p23.y = __dmul_rn(v23.y, ldg_f64(a.x + c.w));
          *reinterpret_cast<double2 *>(prod + 4 * g) = p01;
          *reinterpret_cast<double2 *>(prod + 4 * g + 2) = p23;
        }
      } else {
        base = k0;
#pragma unroll 4
        for (int k = tid; k < cnt; k += kThreads)
          prod[k] = __dmul_rn(__ldg(a.val + k0 + k), ldg_f64(a.x + __ldg(a.col + k0 + k)));
      }
      __syncthreads();
      if (cnt <= 16 * nrows) {
        // short rows: one thread per row, sequential (reference) order
        if (tid < nrows) {
          const int r = r0 + tid;
          const int s = __ldg(a.ind + r) - base, e = __ldg(a.ind + r + 1) - base;
          double sum = 0.0;
          for (int k = s; k < e; k++) sum = __dadd_rn(sum, prod[k]);
          a.y[r] = sum;
          if (DOTS) {
            if (a.w1) acc0 += a.w1[r] * sum;
            if (a.w2_is_y) acc1 += sum * sum;
            else if (a.w2) acc1 += a.w2[r] * sum;
          }
        }
      } else {
        // long rows: one warp per row
        const int w = tid >> 5, l = tid & 31;
        for (int rr = w; rr < nrows; rr += kThreads / 32) {
          const int r = r0 + rr;
          const int s = __ldg(a.ind + r) - base, e = __ldg(a.ind + r + 1) - base;
          double sum = 0.0;
          for (int k = s + l; k < e; k += 32) sum += prod[k];
          sum = warp_sum(sum);
          if (l == 0) {
            a.y[r] = sum;
            if (DOTS) {
              if (a.w1) acc0 += a.w1[r] * sum;
              if (a.w2_is_y) acc1 += sum * sum;
              else if (a.w2) acc1 += a.w2[r] * sum;
            }
          }
        }
      }
      __syncthreads();
    } else {
      // one over-long row: whole CTA, strided, tree reduction
      double sum = 0.0;
      for (int k = k0 + tid; k < k1; k += kThreads)
        sum += __ldg(a.val + k) * ldg_f64(a.x + __ldg(a.col + k));
      sum = block_sum(sum, red_sh);
      if (tid == 0) {
        a.y[r0] = sum;
        if (DOTS) {
          if (a.w1) acc0 += a.w1[r0] * sum;
          if (a.w2_is_y) acc1 += sum * sum;
          else if (a.w2) acc1 += a.w2[r0] * sum;
        }
      }
      __syncthreads();
    }
  }

  if (DOTS) {
    double mine[2];
    mine[0] = block_sum(acc0, red_sh);
    mine[1] = block_sum(acc1, red_sh);
    double tot[2], dummy;
    if (grid_reduce_finish<2, false>(mine, 0.0, a.partials, a.ticket, tot, &dummy, red_sh, a.S->dist)) {
      if (tid == 0) {
        a.S->red[0] = tot[0];
        a.S->red[1] = tot[1];
        run_epilogue(a.epilogue, a.S, a.aux);
      }
    }
  }
}


// ---------------------------------------------------------------- TMA-fed variant
// Same row-block schedule and the same arithmetic, but the block's slices of
// val / col / ind are brought into shared memory by the TMA unit
// (cp.async.bulk, completion on an mbarrier) through a STAGES-deep ring, so
// the streaming part of the traffic costs no LSU issue slots or registers and
// the next blocks are in flight while the current one is reduced.  One thread
// per row then walks its row in shared memory: consecutive lanes gather
// x[col] for consecutive rows, which for banded matrices touches 2-3 cache
// lines per warp instruction instead of one line per diagonal per few lanes.
// The pass itself lives in psp_tma_ring.cuh (shared with the persistent solver
// kernels).
// (shared memory admits at most 4 CTAs per SM -- 2 stages of 25.7 KB each -- so 64 registers per
// thread cost no occupancy; with the default bound ptxas aims at 8 CTAs, stops at 32 registers and
// spills the two running dot products around the irregular-row phase)
template <int STAGES, bool DOTS>
__global__ void __launch_bounds__(kThreads, 4) k_spmv_tma(SpmvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ double red_sh[8];
  if (DOTS) {
    const volatile State *V = a.S;
    if (V->done || (a.gate == 1 && V->cyc_done)) return;
  }
  MatView A;
  A.ind = a.ind; A.col = a.col; A.val = a.val; A.nb = a.nb; A.nint = a.nb; A.blk_halo = nullptr;
  A.blk = reinterpret_cast<const int4 *>(a.blk);
  Ring<STAGES> ring;
  ring.init(smem, full, a.l2_stream != 0);
  double acc0 = 0.0, acc1 = 0.0;
  ring.template pass<true>(A, a.x, red_sh, [&](int r, double sum) {
    a.y[r] = sum;
    if (DOTS) {
      if (a.w1) acc0 += a.w1[r] * sum;
      if (a.w2_is_y) acc1 += sum * sum;
      else if (a.w2) acc1 += a.w2[r] * sum;
    }
  });
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

#ifndef PSP_IRR_VARIANT
#define PSP_IRR_VARIANT 4
#endif

template <int STAGES, bool DOTS>
static int launch_tma(const SpmvArgs &a, int nb, int ctas_per_sm) {
  Ctx &c = ctx();
  static_assert(STAGES >= 2 && STAGES <= 4, "launch cache is indexed by STAGES - 2");
  int &resident = c.occ_tma[STAGES - 2][DOTS ? 1 : 0], &cached_cap = c.occ_tma_cap[STAGES - 2][DOTS ? 1 : 0];
  const int smem_bytes = STAGES * kStageBytes;
  if (resident == 0 || cached_cap != ctas_per_sm) {
    cached_cap = ctas_per_sm;
    PSP_CUDA(cudaFuncSetAttribute(k_spmv_tma<STAGES, DOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_spmv_tma<STAGES, DOTS>, kThreads, smem_bytes) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (ctas_per_sm > 0 && per_sm > ctas_per_sm) per_sm = ctas_per_sm;
    resident = per_sm * c.sms;
    if (resident > kMaxGrid) resident = kMaxGrid;
  }
  const int grid = nb < resident ? nb : resident;
  k_spmv_tma<STAGES, DOTS><<<grid, kThreads, smem_bytes, c.stream>>>(a);
  PSP_LAUNCH_CHECK();
  return PSP_OK;
}

// ---------------------------------------------------------------- host side

static int upload(const void *src, size_t bytes, void **dst, size_t pad_bytes = 64) {
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes + pad_bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes + pad_bytes, cudaGetErrorString(e));
    return PSP_ERR_NOMEM;
  }
  PSP_CUDA(cudaMemsetAsync((char *)p + bytes, 0, pad_bytes, ctx().stream));
  if (bytes) PSP_CUDA(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, ctx().stream));
  *dst = p;
  return PSP_OK;
}

// Greedy row-block schedule: consecutive rows while the block stays within
// kBlkNnz entries and kThreads rows; an over-long row is a block by itself.
static void build_row_blocks(int m, const int *ind, std::vector<int> &blk) {
  blk.clear();
  blk.push_back(0);
  int r = 0;
  while (r < m) {
    int start = r;
    long long base = ind[r];
    int e = r;
    while (e < m && (e - start) < kThreads && (long long)ind[e + 1] - base <= kBlkNnz) e++;
    if (e == start) e = start + 1;   // single row longer than a block
    blk.push_back(e);
    r = e;
  }
}

static int check_csr(int m, int n, long long nnz, const int *ind, const int *col, bool strict_lower) {
  if (m < 0 || n < 0 || nnz < 0) { set_error("negative matrix dimension"); return PSP_ERR_ARG; }
  if (ind[0] != 0 || ind[m] != nnz) { set_error("ind[0] != 0 or ind[m] != nnz"); return PSP_ERR_STRUCTURE; }
  for (int i = 0; i < m; i++) {
    if (ind[i + 1] < ind[i]) { set_error("row pointer array not monotone at row %d", i); return PSP_ERR_STRUCTURE; }
    for (int k = ind[i]; k < ind[i + 1]; k++) {
      int c = col[k];
      if (c < 0 || c >= n || (strict_lower && c >= i)) {
        set_error("column index %d out of range in row %d", c, i);
        return PSP_ERR_STRUCTURE;
      }
    }
  }
  return PSP_OK;
}

// own_cols >= 0 (row slab of a sharded matrix, halo columns numbered from own_cols on):
// the blocks that touch no halo column are scheduled first, A->nint of them.
static int finish_csr_upload(psp_mat *A, const std::vector<int> &ind, const int *col, const double *val,
                             int own_cols = -1) {
  std::vector<int> rows, blk, order;
  build_row_blocks(A->m, ind.data(), rows);
  A->nb = (int)rows.size() - 1;
  A->nint = A->nb;
  A->irregular = A->m > 0 && ind[A->m] > 16LL * A->m;
  order.reserve((size_t)A->nb);
  if (own_cols >= 0) {
    std::vector<int> boundary;
    for (int b = 0; b < A->nb; b++) {
      bool halo = false;
      for (int k = ind[rows[b]]; k < ind[rows[b + 1]] && !halo; k++) halo = col[k] >= own_cols;
      (halo ? boundary : order).push_back(b);
    }
    A->nint = (int)order.size();
    order.insert(order.end(), boundary.begin(), boundary.end());
  } else {
    for (int b = 0; b < A->nb; b++) order.push_back(b);
  }
  blk.resize(4 * (size_t)A->nb);        // (first row, first entry, end row, end entry) per block
  for (int i = 0; i < A->nb; i++) {
    const int b = order[i];
    blk[4 * i] = rows[b]; blk[4 * i + 1] = ind[rows[b]];
    blk[4 * i + 2] = rows[b + 1]; blk[4 * i + 3] = ind[rows[b + 1]];
  }
  PSP_TRY(upload(ind.data(), sizeof(int) * (size_t)(A->m + 1), (void **)&A->d_ind));
  PSP_TRY(upload(col, sizeof(int) * (size_t)A->fnnz, (void **)&A->d_col));
  PSP_TRY(upload(val, sizeof(double) * (size_t)A->fnnz, (void **)&A->d_val));
  PSP_TRY(upload(blk.data(), sizeof(int) * blk.size(), (void **)&A->d_blk));
  std::vector<int> bh;                             // outlives the asynchronous copy below
  if (own_cols >= 0) {
    // per block (schedule order): the range of halo positions its entries read -- lets a
    // persistent kernel unpack exactly what a boundary block needs right before it
    bh.assign(2 * (size_t)A->nb, 0);
    for (int i = A->nint; i < A->nb; i++) {
      const int b = order[i];
      int lo = INT_MAX, hi = -1;
      for (int k = ind[rows[b]]; k < ind[rows[b + 1]]; k++)
        if (col[k] >= own_cols) { lo = col[k] < lo ? col[k] : lo; hi = col[k] > hi ? col[k] : hi; }
      bh[2 * (size_t)i] = lo - own_cols; bh[2 * (size_t)i + 1] = hi + 1 - own_cols;
    }
    PSP_TRY(upload(bh.data(), sizeof(int) * bh.size(), (void **)&A->d_blk_halo));
  }
  PSP_CUDA(cudaStreamSynchronize(ctx().stream));   // host staging vectors die with this frame
  return PSP_OK;
}

int csr_create(int m, int n, long long nnz, const int *ind, const int *col, const double *val,
               psp_mat **out, bool keep_host, int own_cols, const int *gcol, int grow0) {
  PSP_TRY(check_csr(m, n, nnz, ind, col, false));
  psp_mat *A = new psp_mat();
  A->kind = 1; A->m = m; A->n = n; A->nnz = nnz; A->fnnz = nnz;
  std::vector<int> hind(ind, ind + m + 1);
  // rows lined up on diagonals: the run format (no column indices on the device at all)
  bool runs = false;
  int st = runs_try_csr(A, ind, col, val, gcol, grow0, own_cols, &runs);
  if (st == PSP_OK && !runs) st = finish_csr_upload(A, hind, col, val, own_cols);
  if (st != PSP_OK) { psp_mat_destroy(A); return st; }
  if (keep_host) {
    A->h_ind.swap(hind);
    A->h_col.assign(col, col + nnz);
    A->h_val.assign(val, val + nnz);
  }
  *out = A;
  return PSP_OK;
}

// symmetric expansion of an SSS matrix in to_csr() order (ll_mat.c:1601-1623)
void sss_expand(int n, long long nnz, const int *ind, const int *col, const double *val, const double *diag,
                std::vector<int> &find, std::vector<int> &fcol, std::vector<double> &fval) {
  const long long fnnz = 2 * nnz + n;
  find.assign((size_t)n + 1, 0); fcol.resize((size_t)fnnz); fval.resize((size_t)fnnz);
  std::vector<int> fill((size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) find[i + 1] = ind[i + 1] - ind[i] + 1;
  for (long long k = 0; k < nnz; k++) find[col[k] + 1]++;
  for (int i = 0; i < n; i++) find[i + 1] += find[i];
  for (int i = 0; i < n; i++) {
    int p = find[i];
    for (int k = ind[i]; k < ind[i + 1]; k++) { fcol[p] = col[k]; fval[p] = val[k]; p++; }
    fcol[p] = i; fval[p] = diag[i]; p++;
    fill[i] = p;
  }
  for (int i = 0; i < n; i++)               // rows ascending => upper parts ascending
    for (int k = ind[i]; k < ind[i + 1]; k++) {
      int j = col[k];
      fcol[fill[j]] = i; fval[fill[j]] = val[k]; fill[j]++;
    }
}

int sss_create(int n, long long nnz, const int *ind, const int *col, const double *val,
               const double *diag, psp_mat **out) {
  PSP_TRY(check_csr(n, n, nnz, ind, col, true));
  long long fnnz = 2 * nnz + n;
  psp_mat *A = new psp_mat();
  A->kind = 2; A->m = n; A->n = n; A->nnz = nnz; A->fnnz = fnnz;
  {
    // SSS-native: every stored entry of the lower triangle is kept ONCE on the device and used
    // for both products of sss_matvec (sss_mat.c:47-54) through two run descriptors
    bool runs = false;
    int st = runs_try_sss(A, ind, col, val, diag, &runs);
    if (st != PSP_OK) { psp_mat_destroy(A); return st; }
    if (runs) {
      A->h_ind.assign(ind, ind + n + 1);
      A->h_col.assign(col, col + nnz);
      A->h_val.assign(val, val + nnz);
      A->h_diag.assign(diag, diag + n);
      *out = A;
      return PSP_OK;
    }
  }
  // irregular symmetric matrices: fall back to the one-time symmetric expansion
  if (fnnz >= (1LL << 31)) { psp_mat_destroy(A); set_error("expanded SSS matrix exceeds 2^31 entries"); return PSP_ERR_ARG; }
  std::vector<int> find, fcol;
  std::vector<double> fval;
  sss_expand(n, nnz, ind, col, val, diag, find, fcol, fval);
  int st = finish_csr_upload(A, find, fcol.data(), fval.data());
  if (st == PSP_OK) st = upload(diag, sizeof(double) * (size_t)n, (void **)&A->d_diag);
  if (st == PSP_OK && cudaStreamSynchronize(ctx().stream) != cudaSuccess) st = PSP_ERR_CUDA;
  if (st != PSP_OK) { psp_mat_destroy(A); return st; }
  A->h_ind.assign(ind, ind + n + 1);
  A->h_col.assign(col, col + nnz);
  A->h_val.assign(val, val + nnz);
  A->h_diag.assign(diag, diag + n);
  *out = A;
  return PSP_OK;
}

// One-time transpose index: CSR of A^T via counting sort over the columns;
// rows are visited in ascending order, so every row of A^T lists its entries
// by ascending original row -- the order csr_matvec_transp_kernel
// (csr_mat.c:74-88) accumulates y[col] in.
int mat_transpose(psp_mat *A) {
  if (A->transp) return PSP_OK;
  if (A->kind != 1) { set_error("transpose index is only built for CSR matrices"); return PSP_ERR_ARG; }
  const int m = A->m, n = A->n;
  const long long nnz = A->nnz;
  std::vector<int> ind, col;
  std::vector<double> val;
  if (!A->h_ind.empty()) { ind = A->h_ind; col = A->h_col; val = A->h_val; }
  else if (A->fmt == 1) { PSP_TRY(runs_to_csr(A, ind, col, val)); }
  else {
    ind.resize((size_t)m + 1); col.resize((size_t)nnz); val.resize((size_t)nnz);
    PSP_CUDA(cudaMemcpy(ind.data(), A->d_ind, sizeof(int) * (size_t)(m + 1), cudaMemcpyDeviceToHost));
    PSP_CUDA(cudaMemcpy(col.data(), A->d_col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
    PSP_CUDA(cudaMemcpy(val.data(), A->d_val, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost));
  }
  std::vector<int> tind((size_t)n + 1, 0), tcol((size_t)nnz), fill((size_t)n);
  std::vector<double> tval((size_t)nnz);
  for (long long k = 0; k < nnz; k++) tind[col[k] + 1]++;
  for (int j = 0; j < n; j++) tind[j + 1] += tind[j];
  for (int j = 0; j < n; j++) fill[j] = tind[j];
  for (int i = 0; i < m; i++)
    for (int k = ind[i]; k < ind[i + 1]; k++) {
      int j = col[k];
      tcol[fill[j]] = i; tval[fill[j]] = val[k]; fill[j]++;
    }
  psp_mat *T = nullptr;
  PSP_TRY(csr_create(n, m, nnz, tind.data(), tcol.data(), tval.data(), &T, false, -1, nullptr, 0));
  A->transp = T;
  return PSP_OK;
}

static int spmv_dispatch(const psp_mat *A, const double *dx, double *dy, const SpmvDots &d);

// Every `spmv_sample`-th SpMV launch is bracketed by a CUDA-event pair on the
// library stream (bench.py's live roofline measurement: the kernel is timed
// where it runs, inside the solver loop).  Off by default.
int spmv_launch(const psp_mat *A, const double *dx, double *dy, const SpmvDots &d) {
  Ctx &c = ctx();
  if (A->m == 0) return PSP_OK;
  if (A->ceq) {
    // operator facade (psp_correq.cu); the solvers unwrap it themselves and add their reductions
    if (d.w1 || d.w2 || d.w2_is_y || d.epilogue != EPI_NONE) { set_error("spmv: fused reductions on an operator handle"); return PSP_ERR_ARG; }
    return correq_apply(A->ceq, 0, dx, dy, d.S, d.gate);
  }
  bool sample = false;
  if (c.spmv_sample > 0) {
    sample = (c.spmv_seen % c.spmv_sample == 0) && c.spmv_pairs < Ctx::kSpmvPairs;
    c.spmv_seen++;
  }
  if (!sample) return spmv_dispatch(A, dx, dy, d);
  if (!c.spmv_ev[0]) {
    for (int i = 0; i < 2 * Ctx::kSpmvPairs; i++) PSP_CUDA(cudaEventCreate(&c.spmv_ev[i]));
  }
  PSP_CUDA(cudaEventRecord(c.spmv_ev[2 * c.spmv_pairs], c.stream));
  PSP_TRY(spmv_dispatch(A, dx, dy, d));
  PSP_CUDA(cudaEventRecord(c.spmv_ev[2 * c.spmv_pairs + 1], c.stream));
  c.spmv_pairs++;
  return PSP_OK;
}

static int spmv_dispatch(const psp_mat *A, const double *dx, double *dy, const SpmvDots &d) {
  Ctx &c = ctx();
  if (A->fmt == 1) return runs_launch(A, dx, dy, d);
  SpmvArgs a;
  a.ind = A->d_ind; a.col = A->d_col; a.blk = A->d_blk; a.val = A->d_val;
  a.x = dx; a.y = dy; a.nb = A->nb;
  a.w1 = d.w1; a.w2 = d.w2; a.w2_is_y = d.w2_is_y; a.epilogue = d.epilogue; a.S = d.S;
  a.gate = d.gate;
  a.l2_stream = l2_stream_for(A->m);
  a.partials = c.partials; a.ticket = c.tickets; a.aux = nullptr;
  if (d.S) a.aux = (void *)(d.S + 1);   // solver-specific small state sits right behind the State
  // ring depth: 2 stages (4 CTAs per SM) for stencil-like rows; matrices whose blocks take the
  // entry-parallel path (mean row > 16 entries) spend a block's time between barriers, not in the
  // stream, and want the next blocks to be there already: deeper ring, fewer CTAs
  // (power-law 2^21 rows: 538 us with 2 stages, 460 with 3; profiles/r2_irregular_notes.md)
  int variant = c.spmv_variant;
  if (variant == 0) variant = A->irregular ? PSP_IRR_VARIANT : 3;
  if (variant == 3) {
    return d.S ? launch_tma<2, true>(a, A->nb, (int)c.spmv_ctas) : launch_tma<2, false>(a, A->nb, (int)c.spmv_ctas);
  }
  if (variant == 4) {
    return d.S ? launch_tma<3, true>(a, A->nb, (int)c.spmv_ctas) : launch_tma<3, false>(a, A->nb, (int)c.spmv_ctas);
  }
  if (variant == 5) {
    return d.S ? launch_tma<4, true>(a, A->nb, (int)c.spmv_ctas) : launch_tma<4, false>(a, A->nb, (int)c.spmv_ctas);
  }
  const bool vec = (c.spmv_variant != 1);
  // persistent grid: exactly the number of CTAs that are resident at once
  int &res = c.occ_stream[vec ? 1 : 0][d.S ? 1 : 0];
  if (res == 0) {
    const void *fn = vec ? (d.S ? (const void *)k_spmv_stream<1, true> : (const void *)k_spmv_stream<1, false>)
                         : (d.S ? (const void *)k_spmv_stream<0, true> : (const void *)k_spmv_stream<0, false>);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    res = per_sm * c.sms;
    if (res > kMaxGrid) res = kMaxGrid;
  }
  int grid = A->nb < res ? A->nb : res;
  if (d.S) {
    if (vec) k_spmv_stream<1, true><<<grid, kThreads, 0, c.stream>>>(a);
    else k_spmv_stream<0, true><<<grid, kThreads, 0, c.stream>>>(a);
  } else {
    if (vec) k_spmv_stream<1, false><<<grid, kThreads, 0, c.stream>>>(a);
    else k_spmv_stream<0, false><<<grid, kThreads, 0, c.stream>>>(a);
  }
  PSP_LAUNCH_CHECK();
  return PSP_OK;
}

}  // namespace psp

// ---------------------------------------------------------------- C-ABI
using namespace psp;

namespace psp {
int csr_create(int, int, long long, const int *, const int *, const double *, psp_mat **, bool, int, const int *, int);
int sss_create(int, long long, const int *, const int *, const double *, const double *, psp_mat **);
}

extern "C" int psp_csr_create(int m, int n, int nnz, const int *ind, const int *col,
                              const double *val, psp_mat **out) {
  if (!out || !ind || (nnz > 0 && (!col || !val))) { set_error("psp_csr_create: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  PSP_TRY(csr_create(m, n, nnz, ind, col, val, out, false, -1, nullptr, 0));
  if (multi_gpus() > 1) {                // additionally: row slabs on the other GPUs (psp_multi.cu)
    int st = multi_attach_csr(*out, ind, col, val);
    if (st != PSP_OK) { psp_mat_destroy(*out); *out = nullptr; return st; }
  }
  return PSP_OK;
}

extern "C" int psp_sss_create(int n, int nnz, const int *ind, const int *col, const double *val,
                              const double *diag, psp_mat **out) {
  if (!out || !ind || (n > 0 && !diag) || (nnz > 0 && (!col || !val))) {
    set_error("psp_sss_create: NULL argument");
    return PSP_ERR_ARG;
  }
  PSP_TRY(ensure_init());
  PSP_TRY(sss_create(n, nnz, ind, col, val, diag, out));
  if (multi_gpus() > 1 && 2LL * nnz + n < (1LL << 31)) {
    // sharded form: rows of the symmetric expansion (every slab gathers over its own rows, the
    // entries of the lower triangle it needs from later rows are replicated into it)
    std::vector<int> find, fcol;
    std::vector<double> fval;
    sss_expand(n, nnz, ind, col, val, diag, find, fcol, fval);
    int st = multi_attach_csr(*out, find.data(), fcol.data(), fval.data());
    if (st != PSP_OK) { psp_mat_destroy(*out); *out = nullptr; return st; }
  }
  return PSP_OK;
}

extern "C" void psp_mat_destroy(psp_mat *A) {
  if (!A || A->ceq) return;                        // a facade belongs to its psp_correq
  if (A->multi) { multi_detach(A->multi); A->multi = nullptr; }
  if (A->transp) psp_mat_destroy(A->transp);
  cudaFree(A->d_ind); cudaFree(A->d_col); cudaFree(A->d_val); cudaFree(A->d_blk); cudaFree(A->d_blk_halo);
  if (!(A->fmt == 1 && A->kind == 2)) cudaFree(A->d_diag);      // run format: the diagonal lives inside d_rval
  cudaFree(A->d_rval); cudaFree(A->d_rseg); cudaFree(A->d_rblk);
  cudaFree(A->d_x); cudaFree(A->d_y);
  delete A;
}

extern "C" int psp_mat_info(const psp_mat *A, int *m, int *n, long long *nnz, int *kind) {
  if (!A) { set_error("psp_mat_info: NULL matrix"); return PSP_ERR_ARG; }
  if (m) *m = A->m;
  if (n) *n = A->n;
  if (nnz) *nnz = A->nnz;
  if (kind) *kind = A->kind;
  return PSP_OK;
}

extern "C" int psp_matvec_dev(psp_mat *A, int transp, const double *dx, double *dy) {
  if (!A || !dx || !dy) { set_error("psp_matvec_dev: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  const psp_mat *M = A;
  if (transp && A->kind == 1) { PSP_TRY(mat_transpose(A)); M = A->transp; }
  return spmv_launch(M, dx, dy, SpmvDots());
}

static int ensure_vec(double **p, size_t *cap, size_t n) {
  if (*cap >= n && *p) return PSP_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  cudaError_t e = cudaMalloc((void **)p, sizeof(double) * (n > 0 ? n : 1));
  if (e != cudaSuccess) { set_error("cudaMalloc of a vector of %zu doubles failed", n); return PSP_ERR_NOMEM; }
  *cap = n;
  return PSP_OK;
}

extern "C" int psp_matvec(psp_mat *A, int transp, const double *x, long incx, double *y, long incy) {
  if (!A || (!x && A->n) || (!y && A->m)) { set_error("psp_matvec: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  Ctx &c = ctx();
  const bool t = transp && A->kind == 1;
  const size_t nx = t ? A->m : A->n, ny = t ? A->n : A->m;
  PSP_TRY(ensure_vec(&A->d_x, &A->x_cap, nx));
  PSP_TRY(ensure_vec(&A->d_y, &A->y_cap, ny));
  PSP_TRY(psp_vec_upload(A->d_x, x, nx, incx));
  PSP_CUDA(cudaEventRecord(c.ev0, c.stream));
  PSP_TRY(psp_matvec_dev(A, transp, A->d_x, A->d_y));
  PSP_CUDA(cudaEventRecord(c.ev1, c.stream));
  PSP_TRY(psp_vec_download(A->d_y, y, ny, incy));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  c.last_ms = ms;
  return PSP_OK;
}
