// psp_correq.cu -- the Jacobi-Davidson correction-equation operator, resident on the device.
//
// JDSYM hands its inner solver one object that plays both roles (pysparse/eigen/src/jdsym.c:570,
// correq.c:117-221):
//   matvec   SYM  : y = (I - Qb Q') (A - theta M) x
//            UNSYM: y = (I - Y inv(H) Qb') K^-1 (A - theta M) x      (with a preconditioner)
//                   y = (I - Q Qb') (A - theta M) x                  (without)
//   precon   SYM  : y = (I - Y inv(H) Qb') K^-1 x   /   (I - Q Qb') x   /   x
//            UNSYM: y = x
// with Qb = M Q for a generalised problem and Qb = Q otherwise, k locked + active Ritz vectors.
// In the reference these are dgemv calls on host arrays around SpMatrix_Matvec / _Precon callbacks;
// round 1 of this library reached the solver through the host-callback boundary, i.e. two PCIe
// crossings of a full vector per inner matvec.  Here Q, M Q and Y = K^-1 Qb live in HBM next to the
// matrix, the projections are two small kernels (k dot products, then a rank-k update) and the
// operator is handed to psp_solve as an ordinary psp_mat / psp_prec handle (a facade whose `ceq`
// field points here), so the whole inner solve runs without touching the host.
#include "psp_correq.cuh"
#include "psp_vec.cuh"

#include <new>

namespace psp {

constexpr int kCeqThreads = 256;
constexpr int kCeqCols = 8;           // columns per pass of the dot-product kernel

// part[blockIdx.x][j] = sum over this block's rows of B[i, j] * x[i]
__global__ void __launch_bounds__(kCeqThreads) k_ceq_dots(const double *B, long ldb, int k, const double *x, long n,
                                                           double *part, const State *S, int gate) {
  if (gated_out(S, gate)) return;
  __shared__ double red[kCeqCols][kCeqThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long stride = (long)gridDim.x * kCeqThreads;
  for (int c0 = 0; c0 < k; c0 += kCeqCols) {
    const int nc = (k - c0 < kCeqCols) ? k - c0 : kCeqCols;
    double acc[kCeqCols];
#pragma unroll
    for (int j = 0; j < kCeqCols; j++) acc[j] = 0.0;
    for (long i = (long)blockIdx.x * kCeqThreads + threadIdx.x; i < n; i += stride) {
      const double xi = x[i];
#pragma unroll
      for (int j = 0; j < kCeqCols; j++)
        if (j < nc) acc[j] += B[(size_t)(c0 + j) * ldb + i] * xi;
    }
#pragma unroll
    for (int j = 0; j < kCeqCols; j++) {
      const double s = warp_sum(acc[j]);
      if (lane == 0) red[j][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < nc) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kCeqThreads / 32; w++) s += red[threadIdx.x][w];
      part[(size_t)blockIdx.x * kCeqMaxK + c0 + threadIdx.x] = s;
    }
    __syncthreads();
  }
}

// c = sum of the block partials (fixed order: deterministic), w = G c (G == nullptr: w = c)
__global__ void __launch_bounds__(kCeqThreads) k_ceq_small(const double *part, int nblk, int k, const double *G, int ldg,
                                                            double *w, const State *S, int gate) {
  if (gated_out(S, gate)) return;
  __shared__ double c[kCeqMaxK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = warp; j < k; j += kCeqThreads / 32) {
    double s = 0.0;
    for (int b = lane; b < nblk; b += 32) s += part[(size_t)b * kCeqMaxK + j];
    s = warp_sum(s);
    if (lane == 0) c[j] = s;
  }
  __syncthreads();
  if (threadIdx.x < k) {
    double s = c[threadIdx.x];
    if (G) {
      s = 0.0;
      for (int j = 0; j < k; j++) s += G[(size_t)threadIdx.x * ldg + j] * c[j];
    }
    w[threadIdx.x] = s;
  }
}

// y = x - A[:, :k] w
__global__ void __launch_bounds__(kCeqThreads) k_ceq_update(const double *A, long lda, int k, const double *w, const double *x,
                                                             double *y, long n, const State *S, int gate) {
  if (gated_out(S, gate)) return;
  __shared__ double ws[kCeqMaxK];
  if (threadIdx.x < k) ws[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const long stride = (long)gridDim.x * kCeqThreads;
  for (long i = (long)blockIdx.x * kCeqThreads + threadIdx.x; i < n; i += stride) {
    double s = 0.0;
    for (int j = 0; j < k; j++) s += A[(size_t)j * lda + i] * ws[j];
    y[i] = x[i] - s;
  }
}

struct FCeqShift {          // t = t - theta v   (two roundings, like the host expression it replaces)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  double *t; const double *v; double theta;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *, double &) { t[i] = __dsub_rn(t[i], __dmul_rn(theta, v[i])); }
};

static int ceq_grid(const psp_correq *E) {
  long g = (E->n + kCeqThreads - 1) / kCeqThreads;
  return (int)(g < 1 ? 1 : (g > E->nblk ? E->nblk : g));
}

// y = (I - Acols G Bcols') x      (G == nullptr: identity);  x == y is allowed
static int ceq_project(psp_correq *E, const double *Acols, const double *Bcols, const double *G, const double *x,
                       double *y, const State *S, int gate) {
  Ctx &c = ctx();
  const long n = E->n;
  if (E->k == 0) {
    if (x != y) PSP_CUDA(cudaMemcpyAsync(y, x, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
    return PSP_OK;
  }
  const int grid = ceq_grid(E);
  k_ceq_dots<<<grid, kCeqThreads, 0, c.stream>>>(Bcols, n, E->k, x, n, E->d_part, S, gate);
  PSP_LAUNCH_CHECK();
  k_ceq_small<<<1, kCeqThreads, 0, c.stream>>>(E->d_part, grid, E->k, G, kCeqMaxK, E->d_w, S, gate);
  PSP_LAUNCH_CHECK();
  k_ceq_update<<<grid, kCeqThreads, 0, c.stream>>>(Acols, n, E->k, E->d_w, x, y, n, S, gate);
  PSP_LAUNCH_CHECK();
  return PSP_OK;
}

// t = (A - theta M) x   /   (A - theta I) x        (correq.c:117-143)
static int ceq_shifted(psp_correq *E, const double *dx, double *t, State *S, int gate) {
  SpmvDots d;
  d.S = S; d.gate = gate & kGateMask;
  if (!S) d = SpmvDots();
  PSP_TRY(spmv_launch(E->A, dx, t, d));
  if (E->M) {
    if (E->theta != 0.0) {
      PSP_TRY(spmv_launch(E->M, dx, E->d_t2, d));
      PSP_TRY(launch_map(FCeqShift{t, E->d_t2, E->theta}, E->n, EPI_NONE, gate));
    }
  } else {
    PSP_TRY(launch_map(FCeqShift{t, dx, E->theta}, E->n, EPI_NONE, gate));
  }
  return PSP_OK;
}

int correq_apply(psp_correq *E, int role, const double *dx, double *dy, State *S, int gate) {
  Ctx &c = ctx();
  const long n = E->n;
  if (n == 0) return PSP_OK;
  if (!S) gate = 3;                                  // outside a solve: nothing to be gated by
  const State *gs = S ? S : c.state;
  const double *Qb = E->M ? E->dQm : E->dQ;
  const bool sym = (E->optype == 2);
  if (role == 0) {
    PSP_TRY(ceq_shifted(E, dx, E->d_t, S, gate));
    if (sym) return ceq_project(E, Qb, E->dQ, nullptr, E->d_t, dy, gs, gate);
    if (E->K) {
      PSP_TRY(prec_apply_dev(E->K, E->d_t, E->d_t2, S));
      return ceq_project(E, E->dY, Qb, E->dG, E->d_t2, dy, gs, gate);
    }
    return ceq_project(E, E->dQ, Qb, nullptr, E->d_t, dy, gs, gate);
  }
  if (sym && E->K) {
    PSP_TRY(prec_apply_dev(E->K, dx, E->d_t2, S));
    return ceq_project(E, E->dY, Qb, E->dG, E->d_t2, dy, gs, gate);
  }
  if (sym && E->M) return ceq_project(E, E->dQ, E->dQm, nullptr, dx, dy, gs, gate);
  if (dx != dy) PSP_CUDA(cudaMemcpyAsync(dy, dx, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
  return PSP_OK;
}

}  // namespace psp

using namespace psp;

static void ceq_free(psp_correq *E) {
  if (!E) return;
  cudaFree(E->dQ); cudaFree(E->dQm); cudaFree(E->dY); cudaFree(E->dG); cudaFree(E->d_w);
  cudaFree(E->d_t); cudaFree(E->d_t2); cudaFree(E->d_part);
  if (E->as_mat) { cudaFree(E->as_mat->d_x); cudaFree(E->as_mat->d_y); delete E->as_mat; }
  if (E->as_prec) { cudaFree(E->as_prec->d_x); cudaFree(E->as_prec->d_y); delete E->as_prec; }
  delete E;
}

extern "C" int psp_correq_create(int optype, psp_mat *A, psp_mat *M, psp_prec *K, int kmax, psp_correq **out) {
  if (!out) { set_error("psp_correq_create: NULL output"); return PSP_ERR_ARG; }
  *out = nullptr;
  if (!A || A->ceq || A->m != A->n) { set_error("psp_correq_create: A must be a square device matrix"); return PSP_ERR_ARG; }
  if (M && (M->ceq || M->m != A->m || M->n != A->n)) { set_error("psp_correq_create: M does not match A"); return PSP_ERR_ARG; }
  if (K && (K->ceq || K->n != A->n)) { set_error("psp_correq_create: K does not match A"); return PSP_ERR_ARG; }
  if (optype != 1 && optype != 2) { set_error("psp_correq_create: optype must be 1 (UNSYM) or 2 (SYM)"); return PSP_ERR_ARG; }
  if (kmax < 1 || kmax > kCeqMaxK) { set_error("psp_correq_create: kmax must be in 1..%d", kCeqMaxK); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  psp_correq *E = new (std::nothrow) psp_correq();
  if (!E) { set_error("psp_correq_create: out of memory"); return PSP_ERR_NOMEM; }
  E->optype = optype; E->n = A->n; E->kmax = kmax; E->A = A; E->M = M; E->K = K;
  E->nblk = 4 * ctx().sms;
  const size_t n1 = (size_t)(E->n > 0 ? E->n : 1), cols = n1 * (size_t)kmax * sizeof(double);
  bool ok = cudaMalloc((void **)&E->dQ, cols) == cudaSuccess;
  if (ok && M) ok = cudaMalloc((void **)&E->dQm, cols) == cudaSuccess;
  if (ok && K) ok = cudaMalloc((void **)&E->dY, cols) == cudaSuccess &&
                    cudaMalloc((void **)&E->dG, sizeof(double) * kCeqMaxK * kCeqMaxK) == cudaSuccess;
  ok = ok && cudaMalloc((void **)&E->d_w, sizeof(double) * kCeqMaxK) == cudaSuccess &&
       cudaMalloc((void **)&E->d_t, n1 * sizeof(double)) == cudaSuccess &&
       cudaMalloc((void **)&E->d_t2, n1 * sizeof(double)) == cudaSuccess &&
       cudaMalloc((void **)&E->d_part, sizeof(double) * (size_t)E->nblk * kCeqMaxK) == cudaSuccess;
  E->as_mat = new (std::nothrow) psp_mat();
  E->as_prec = new (std::nothrow) psp_prec();
  if (!ok || !E->as_mat || !E->as_prec) {
    cudaGetLastError();
    ceq_free(E);
    set_error("psp_correq_create: out of device memory for %d vectors of length %d", kmax, A->n);
    return PSP_ERR_NOMEM;
  }
  E->as_mat->kind = 3; E->as_mat->m = E->as_mat->n = E->n; E->as_mat->ceq = E;
  E->as_prec->kind = 3; E->as_prec->n = E->n; E->as_prec->ceq = E;
  *out = E;
  return PSP_OK;
}

extern "C" void psp_correq_destroy(psp_correq *E) { ceq_free(E); }

extern "C" int psp_correq_set_columns(psp_correq *E, int which, int col0, int ncols, const double *host, long ld) {
  if (!E || which < 0 || which > 2 || col0 < 0 || ncols < 0 || col0 + ncols > E->kmax || (ncols && !host) || ld < E->n) {
    set_error("psp_correq_set_columns: bad arguments");
    return PSP_ERR_ARG;
  }
  double *dst = which == 0 ? E->dQ : which == 1 ? E->dQm : E->dY;
  if (!dst) { set_error("psp_correq_set_columns: the operator has no %s", which == 1 ? "mass matrix" : "preconditioner"); return PSP_ERR_ARG; }
  if (ncols == 0 || E->n == 0) return PSP_OK;
  Ctx &c = ctx();
  PSP_CUDA(cudaMemcpy2DAsync(dst + (size_t)col0 * E->n, sizeof(double) * (size_t)E->n, host, sizeof(double) * (size_t)ld,
                             sizeof(double) * (size_t)E->n, (size_t)ncols, cudaMemcpyHostToDevice, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  return PSP_OK;
}

extern "C" int psp_correq_update(psp_correq *E, int k, double theta, const double *hinv) {
  if (!E || k < 0 || k > E->kmax) { set_error("psp_correq_update: k out of range"); return PSP_ERR_ARG; }
  if (E->K && k > 0 && !hinv) { set_error("psp_correq_update: inv(H) is required with a preconditioner"); return PSP_ERR_ARG; }
  E->k = k; E->theta = theta;
  if (E->K && k > 0) {
    Ctx &c = ctx();
    PSP_CUDA(cudaMemcpy2DAsync(E->dG, sizeof(double) * kCeqMaxK, hinv, sizeof(double) * (size_t)k, sizeof(double) * (size_t)k,
                               (size_t)k, cudaMemcpyHostToDevice, c.stream));
    PSP_CUDA(cudaStreamSynchronize(c.stream));
  }
  return PSP_OK;
}

extern "C" psp_mat *psp_correq_mat(psp_correq *E) { return E ? E->as_mat : nullptr; }
extern "C" psp_prec *psp_correq_prec(psp_correq *E) { return E ? E->as_prec : nullptr; }
