// psp_precon.cu -- Jacobi and SSOR / symmetric Gauss-Seidel preconditioners.
//
// Replaces jacobi() (pysparse/precon/src/preconmodule.c:35-54), ssor_kernel
// (:95-143) and symgs_kernel (:149-193) plus their constructors (:352-459).
//
// SSOR keeps the reference's lexicographic sweep order.  The analysis phase
// (psp_ssor_create) computes, on the host, the level sets of the dependency
// graph of L (forward sweep: row i needs x[j] for every stored L[i,j], j < i)
// and of L^T (backward sweep: row i needs x[i'] for every stored L[i',i],
// i' > i), plus a transpose index of L whose rows list their entries by
// DESCENDING original row -- the order in which the reference's backward loop
// scatters into y[j] -- so the scatter becomes an atomic-free gather with the
// same accumulation order.  Rows inside one level are independent; a single
// persistent cooperative kernel walks the levels with a grid-wide barrier in
// between, so one preconditioner application is ONE launch.
#include <cooperative_groups.h>
#include "psp_mat.cuh"
#include "psp_vec.cuh"

namespace cg = cooperative_groups;

namespace psp {

struct FJacobiCorr {        // y = (x - y) .* dinv + temp      (preconmodule.c:50-51)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *x, *dinv, *temp; double *y;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *, double &) { y[i] = (x[i] - y[i]) * dinv[i] + temp[i]; }
};
struct FJacobiScale {       // y = x .* dinv                   (preconmodule.c:41-42)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *x, *dinv; double *y;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *, double &) { y[i] = x[i] * dinv[i]; }
};

struct SweepArgs {
  int n;
  const int *lind, *lcol;       // L rows (ascending columns)
  const double *lval, *diag;
  const int *tind, *tcol;       // L^T rows, entries by descending original row
  const double *tval;
  const int *rows_f, *lptr_f; int nlev_f;
  const int *rows_b, *lptr_b; int nlev_b;
  const double *b;              // right-hand side (the vector being preconditioned)
  double *x;                    // result
  double *w;                    // work: L x / L^T x   (reference: y resp. h)
  double omega;
  int steps;
  const State *S;               // optional gate
};

// symmetric Gauss-Seidel, omega == 1  (symgs_kernel, preconmodule.c:149-193)
__global__ void __launch_bounds__(kThreads) k_symgs(SweepArgs a) {
  if (a.S && *(const volatile int *)&a.S->done) return;
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gsz = gridDim.x * blockDim.x;
  for (int step = 0; step < a.steps; step++) {
    // forward: x_i = (b_i - w_i - s)/d_i ; w_i = s = sum_{j<i} L_ij x_j
    for (int l = 0; l < a.nlev_f; l++) {
      const int p0 = a.lptr_f[l], p1 = a.lptr_f[l + 1];
      for (int p = p0 + gtid; p < p1; p += gsz) {
        const int r = a.rows_f[p];
        double s = 0.0;
        for (int k = a.lind[r]; k < a.lind[r + 1]; k++)
          s = __dadd_rn(s, __dmul_rn(a.lval[k], __ldcg(a.x + a.lcol[k])));
        const double wr = (step == 0) ? 0.0 : a.w[r];
        a.x[r] = __ddiv_rn(__dsub_rn(__dsub_rn(a.b[r], wr), s), a.diag[r]);
        a.w[r] = s;
      }
      grid.sync();
    }
    // backward: t = sum_{i'>i} L_i'i x_i' (descending i') ; x_i = (b_i - (Lx)_i - t)/d_i ; w_i = t
    for (int l = 0; l < a.nlev_b; l++) {
      const int p0 = a.lptr_b[l], p1 = a.lptr_b[l + 1];
      for (int p = p0 + gtid; p < p1; p += gsz) {
        const int r = a.rows_b[p];
        double t = 0.0;
        for (int k = a.tind[r]; k < a.tind[r + 1]; k++)
          t = __dadd_rn(t, __dmul_rn(a.tval[k], __ldcg(a.x + a.tcol[k])));
        const double s = a.w[r];
        a.x[r] = __ddiv_rn(__dsub_rn(__dsub_rn(a.b[r], s), t), a.diag[r]);
        a.w[r] = t;
      }
      grid.sync();
    }
  }
}

// SSOR with relaxation, omega != 1  (ssor_kernel, preconmodule.c:95-143)
__global__ void __launch_bounds__(kThreads) k_ssor(SweepArgs a) {
  if (a.S && *(const volatile int *)&a.S->done) return;
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gsz = gridDim.x * blockDim.x;
  const double om = a.omega, om1 = 1.0 - a.omega;
  for (int step = 0; step < a.steps; step++) {
    for (int l = 0; l < a.nlev_f; l++) {
      const int p0 = a.lptr_f[l], p1 = a.lptr_f[l + 1];
      for (int p = p0 + gtid; p < p1; p += gsz) {
        const int r = a.rows_f[p];
        const double d = a.diag[r], br = a.b[r];
        double temp;
        if (step == 0) temp = __dmul_rn(om, br);
        else temp = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(om1, a.x[r]), d), a.w[r]), __dmul_rn(om, br));
        double s = 0.0;
        for (int k = a.lind[r]; k < a.lind[r + 1]; k++)
          s = __dsub_rn(s, __dmul_rn(a.lval[k], __ldcg(a.x + a.lcol[k])));
        const double h = __dmul_rn(om, s);
        a.w[r] = h;
        a.x[r] = __ddiv_rn(__dadd_rn(temp, h), d);
      }
      grid.sync();
    }
    for (int l = 0; l < a.nlev_b; l++) {
      const int p0 = a.lptr_b[l], p1 = a.lptr_b[l + 1];
      for (int p = p0 + gtid; p < p1; p += gsz) {
        const int r = a.rows_b[p];
        const double d = a.diag[r];
        const double temp = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(om1, a.x[r]), d), a.w[r]),
                                      __dmul_rn(om, a.b[r]));
        double t = 0.0;
        for (int k = a.tind[r]; k < a.tind[r + 1]; k++)
          t = __dsub_rn(t, __dmul_rn(a.tval[k], __ldcg(a.x + a.tcol[k])));
        const double h = __dmul_rn(om, t);
        a.w[r] = h;
        a.x[r] = __ddiv_rn(__dadd_rn(temp, h), d);
      }
      grid.sync();
    }
  }
}

static int sweep_grid(const void *kernel) {
  static int cached_symgs = 0, cached_ssor = 0;
  int *slot = (kernel == (const void *)k_symgs) ? &cached_symgs : &cached_ssor;
  if (*slot == 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    *slot = per_sm * ctx().sms;
  }
  return *slot;
}

int prec_apply_dev(psp_prec *K, const double *dx, double *dy, State *gateS) {
  Ctx &c = ctx();
  const long n = K->n;
  if (n == 0) return PSP_OK;
  if (K->kind == 1) {
    const int gate = gateS ? 0 : 3;
    PSP_TRY(launch_map(FJacobiScale{dx, K->d_dinv, dy}, n, EPI_NONE, gate));
    for (int step = 1; step < K->steps; step++) {
      PSP_CUDA(cudaMemcpyAsync(K->d_t1, dy, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
      PSP_TRY(spmv_launch(K->mat, K->d_t1, dy, SpmvDots()));
      PSP_TRY(launch_map(FJacobiCorr{dx, K->d_dinv, K->d_t1, dy}, n, EPI_NONE, gate));
    }
    return PSP_OK;
  }
  const psp_mat *A = K->mat;
  SweepArgs a;
  a.n = (int)n;
  a.lind = A->d_lind; a.lcol = A->d_lcol; a.lval = A->d_lval; a.diag = A->d_diag;
  a.tind = K->d_tind; a.tcol = K->d_tcol; a.tval = K->d_tval;
  a.rows_f = K->d_rows_f; a.lptr_f = K->d_lptr_f; a.nlev_f = K->nlev_f;
  a.rows_b = K->d_rows_b; a.lptr_b = K->d_lptr_b; a.nlev_b = K->nlev_b;
  a.b = dx; a.x = dy; a.w = K->d_t1; a.omega = K->omega; a.steps = K->steps; a.S = gateS;
  const void *kern = (K->omega == 1.0) ? (const void *)k_symgs : (const void *)k_ssor;
  void *args[] = {&a};
  PSP_CUDA(cudaLaunchCooperativeKernel(kern, dim3(sweep_grid(kern)), dim3(kThreads), args, 0, c.stream));
  c.launches++;
  return PSP_OK;
}

static int upload_i(const std::vector<int> &v, int **d) {
  size_t bytes = sizeof(int) * (v.size() ? v.size() : 1);
  if (cudaMalloc((void **)d, bytes) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory"); return PSP_ERR_NOMEM; }
  if (!v.empty()) PSP_CUDA(cudaMemcpy(*d, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice));
  return PSP_OK;
}

// level sets + level-sorted row lists (counting sort: rows ascending inside a level)
static void sort_by_level(const std::vector<int> &level, int nlev, std::vector<int> &rows, std::vector<int> &lptr) {
  const int n = (int)level.size();
  lptr.assign((size_t)nlev + 1, 0);
  for (int i = 0; i < n; i++) lptr[level[i] + 1]++;
  for (int l = 0; l < nlev; l++) lptr[l + 1] += lptr[l];
  std::vector<int> fill(lptr.begin(), lptr.end() - 1);
  rows.resize((size_t)n);
  for (int i = 0; i < n; i++) rows[fill[level[i]]++] = i;
}

int ssor_create(psp_mat *A, double omega, int steps, psp_prec **out) {
  if (!A || A->kind != 2) { set_error("ssor: matrix must be an sss_mat"); return PSP_ERR_ARG; }
  const int n = A->n;
  const std::vector<int> &ind = A->h_ind, &col = A->h_col;
  const std::vector<double> &val = A->h_val;
  psp_prec *K = new psp_prec();
  K->kind = 2; K->n = n; K->omega = omega; K->steps = steps; K->mat = A;
  // forward levels
  std::vector<int> &lf = K->h_level_f, &lb = K->h_level_b;
  lf.assign((size_t)n, 0); lb.assign((size_t)n, 0);
  int nlf = 0, nlb = 0;
  for (int i = 0; i < n; i++) {
    int lv = 0;
    for (int k = ind[i]; k < ind[i + 1]; k++) { int c1 = lf[col[k]] + 1; if (c1 > lv) lv = c1; }
    lf[i] = lv;
    if (lv + 1 > nlf) nlf = lv + 1;
  }
  // backward levels: row i waits for every later row i' with L[i',i] != 0
  for (int i = n - 1; i >= 0; i--) {
    if (lb[i] + 1 > nlb) nlb = lb[i] + 1;
    for (int k = ind[i]; k < ind[i + 1]; k++) { int j = col[k]; if (lb[i] + 1 > lb[j]) lb[j] = lb[i] + 1; }
  }
  if (n == 0) { nlf = 0; nlb = 0; }
  K->nlev_f = nlf; K->nlev_b = nlb;
  std::vector<int> rows, lptr;
  int st = PSP_OK;
  sort_by_level(lf, nlf, rows, lptr);
  if (st == PSP_OK) st = upload_i(rows, &K->d_rows_f);
  if (st == PSP_OK) st = upload_i(lptr, &K->d_lptr_f);
  sort_by_level(lb, nlb, rows, lptr);
  if (st == PSP_OK) st = upload_i(rows, &K->d_rows_b);
  if (st == PSP_OK) st = upload_i(lptr, &K->d_lptr_b);
  // transpose index of L, entries of each row by descending original row
  const size_t nnz = col.size();
  std::vector<int> tind((size_t)n + 1, 0), tcol(nnz), fill((size_t)(n > 0 ? n : 1));
  std::vector<double> tval(nnz);
  for (size_t k = 0; k < nnz; k++) tind[col[k] + 1]++;
  for (int j = 0; j < n; j++) tind[j + 1] += tind[j];
  for (int j = 0; j < n; j++) fill[j] = tind[j];
  for (int i = n - 1; i >= 0; i--)
    for (int k = ind[i]; k < ind[i + 1]; k++) { int j = col[k]; tcol[fill[j]] = i; tval[fill[j]] = val[k]; fill[j]++; }
  if (st == PSP_OK) st = upload_i(tind, &K->d_tind);
  if (st == PSP_OK) st = upload_i(tcol, &K->d_tcol);
  if (st == PSP_OK) {
    size_t bytes = sizeof(double) * (nnz ? nnz : 1);
    if (cudaMalloc((void **)&K->d_tval, bytes) != cudaSuccess) { cudaGetLastError(); st = PSP_ERR_NOMEM; }
    else if (nnz && cudaMemcpy(K->d_tval, tval.data(), sizeof(double) * nnz, cudaMemcpyHostToDevice) != cudaSuccess) st = PSP_ERR_CUDA;
  }
  if (st == PSP_OK && cudaMalloc((void **)&K->d_t1, sizeof(double) * (size_t)(n > 0 ? n : 1)) != cudaSuccess) {
    cudaGetLastError(); st = PSP_ERR_NOMEM;
  }
  if (st != PSP_OK) { psp_prec_destroy(K); if (st == PSP_ERR_NOMEM) set_error("ssor: out of device memory"); return st; }
  *out = K;
  return PSP_OK;
}

}  // namespace psp

using namespace psp;

extern "C" int psp_jacobi_create(psp_mat *A, int n, const double *diag, double omega, int steps,
                                 psp_prec **out) {
  if (!out || n < 0 || (n > 0 && !diag)) { set_error("psp_jacobi_create: bad arguments"); return PSP_ERR_ARG; }
  if (steps > 1 && !A) { set_error("psp_jacobi_create: steps > 1 needs the matrix"); return PSP_ERR_ARG; }
  std::vector<double> dinv((size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    const double d = diag[i];
    volatile double t = 1.0 + d;             // preconmodule.c:395
    if (t == 1.0) { set_error("diagonal element close to zero"); return PSP_ERR_ARG; }
    dinv[i] = omega / d;                     // preconmodule.c:400
  }
  PSP_TRY(ensure_init());
  psp_prec *K = new psp_prec();
  K->kind = 1; K->n = n; K->omega = omega; K->steps = steps; K->mat = A;
  size_t bytes = sizeof(double) * (size_t)(n > 0 ? n : 1);
  if (cudaMalloc((void **)&K->d_dinv, bytes) != cudaSuccess ||
      (steps > 1 && cudaMalloc((void **)&K->d_t1, bytes) != cudaSuccess)) {
    cudaGetLastError();
    psp_prec_destroy(K);
    set_error("psp_jacobi_create: out of device memory");
    return PSP_ERR_NOMEM;
  }
  if (n > 0 && cudaMemcpy(K->d_dinv, dinv.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) {
    psp_prec_destroy(K);
    set_error("psp_jacobi_create: upload failed");
    return PSP_ERR_CUDA;
  }
  *out = K;
  return PSP_OK;
}

extern "C" int psp_ssor_create(psp_mat *sss, double omega, int steps, psp_prec **out) {
  if (!out) { set_error("psp_ssor_create: NULL out"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  return ssor_create(sss, omega, steps, out);
}

extern "C" void psp_prec_destroy(psp_prec *K) {
  if (!K) return;
  cudaFree(K->d_dinv); cudaFree(K->d_t1); cudaFree(K->d_t2);
  cudaFree(K->d_rows_f); cudaFree(K->d_lptr_f); cudaFree(K->d_rows_b); cudaFree(K->d_lptr_b);
  cudaFree(K->d_tind); cudaFree(K->d_tcol); cudaFree(K->d_tval);
  cudaFree(K->d_x); cudaFree(K->d_y); cudaFree(K->d_bar);
  delete K;
}

extern "C" int psp_prec_apply_dev(psp_prec *K, const double *dx, double *dy) {
  if (!K || !dx || !dy) { set_error("psp_prec_apply_dev: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  return prec_apply_dev(K, dx, dy, nullptr);
}

extern "C" int psp_prec_apply(psp_prec *K, const double *x, double *y) {
  if (!K || (K->n > 0 && (!x || !y))) { set_error("psp_prec_apply: NULL argument"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  Ctx &c = ctx();
  const size_t n = (size_t)K->n;
  if (!K->d_x) {
    if (cudaMalloc((void **)&K->d_x, sizeof(double) * (n ? n : 1)) != cudaSuccess ||
        cudaMalloc((void **)&K->d_y, sizeof(double) * (n ? n : 1)) != cudaSuccess) {
      cudaGetLastError();
      set_error("psp_prec_apply: out of device memory");
      return PSP_ERR_NOMEM;
    }
  }
  PSP_TRY(psp_vec_upload(K->d_x, x, n, 1));
  PSP_CUDA(cudaEventRecord(c.ev0, c.stream));
  PSP_TRY(prec_apply_dev(K, K->d_x, K->d_y, nullptr));
  PSP_CUDA(cudaEventRecord(c.ev1, c.stream));
  PSP_TRY(psp_vec_download(K->d_y, y, n, 1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  c.last_ms = ms;
  return PSP_OK;
}

extern "C" int psp_ssor_levels(const psp_prec *K, int *nlev_fwd, int *nlev_bwd, int *level_fwd,
                               int *level_bwd) {
  if (!K || K->kind != 2) { set_error("psp_ssor_levels: not an SSOR preconditioner"); return PSP_ERR_ARG; }
  if (nlev_fwd) *nlev_fwd = K->nlev_f;
  if (nlev_bwd) *nlev_bwd = K->nlev_b;
  if (level_fwd) memcpy(level_fwd, K->h_level_f.data(), sizeof(int) * (size_t)K->n);
  if (level_bwd) memcpy(level_bwd, K->h_level_b.data(), sizeof(int) * (size_t)K->n);
  return PSP_OK;
}
