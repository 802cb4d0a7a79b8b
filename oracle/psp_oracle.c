/* oracle/psp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the PySparse hot path (SpMV, preconditioners,
 * Krylov solvers), written from the reference's behaviour -- every routine
 * cites the reference file:line whose arithmetic (operation order included)
 * it restates.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may load this library; the product (pysparse_b200) never does.
 *
 * PINNING: the reference's own tests hold no golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against the reference's
 * own C kernels compiled from /root/reference (oracle/_ref/libpsp_ref.so,
 * see oracle/Makefile + ref_glue.c): tests/test_oracle.py requires BIT-EXACT
 * agreement of every routine here with that library, and tests/golden/*.json
 * holds (info, iter, relres, |x|) tables generated from it.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction), see oracle/Makefile.
 * Conventions: int32 indices, fp64 values, 0-based, vectors are double[n].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* operators                                                           */
/* ------------------------------------------------------------------ */
typedef int (*orc_cb)(void *ctx, int n, double *x, double *y);

typedef struct orc_op {
  int kind;                 /* 1 csr, 2 sss, 3 jacobi, 4 ssor, 5 callback */
  int m, n;
  const double *val, *diag;
  const int *col, *ind;
  struct orc_op *mat;
  double *dinv, *t1, *t2;
  double omega;
  int steps;
  orc_cb cb;
  void *ctx;
} orc_op;

enum { K_CSR = 1, K_SSS = 2, K_JACOBI = 3, K_SSOR = 4, K_CB = 5 };

/* ------------------------------------------------------------------ */
/* BLAS-1 (external Fortran BLAS in the reference, include/blas.h:97-126) */
/* ------------------------------------------------------------------ */
static double v_dot(int n, const double *a, const double *b)
{
  /* Netlib ddot with unit strides sums left to right */
  double acc = 0.0;
  for (int i = 0; i < n; i++) acc = acc + a[i] * b[i];
  return acc;
}

static double v_nrm2(int n, const double *a)
{
  /* Netlib dnrm2: running (scale, ssq) pair, result scale*sqrt(ssq) */
  if (n < 1) return 0.0;
  if (n == 1) return fabs(a[0]);
  double big = 0.0, acc = 1.0;
  for (int i = 0; i < n; i++) {
    if (a[i] == 0.0) continue;
    double mag = fabs(a[i]);
    if (big < mag) {
      double q = big / mag;
      acc = 1.0 + acc * q * q;
      big = mag;
    } else {
      double q = mag / big;
      acc = acc + q * q;
    }
  }
  return big * sqrt(acc);
}

static void v_axpy(int n, double a, const double *x, double *y)
{
  if (a == 0.0) return;                 /* Netlib daxpy quick return */
  for (int i = 0; i < n; i++) y[i] = y[i] + a * x[i];
}

static void v_copy(int n, const double *x, double *y)
{
  memcpy(y, x, sizeof(double) * (size_t)n);
}

static void v_scal(int n, double a, double *x)
{
  for (int i = 0; i < n; i++) x[i] = a * x[i];
}

/* ------------------------------------------------------------------ */
/* SpMV                                                                */
/* ------------------------------------------------------------------ */

/* csr_mat.c:49-54 (contiguous) and :58-72 (strided): row sums, k ascending */
void orc_csr_matvec(int m, const double *val, const int *col, const int *ind,
                    const double *x, int incx, double *y, int incy)
{
  for (int r = 0; r < m; r++) {
    double acc = 0.0;
    for (int k = ind[r]; k < ind[r + 1]; k++)
      acc += val[k] * x[(long)col[k] * incx];
    y[(long)r * incy] = acc;
  }
}

/* csr_mat.c:74-88 / :90-106: y = 0, then row-major scatter */
void orc_csr_matvec_transp(int m, int n, const double *val, const int *col, const int *ind,
                           const double *x, int incx, double *y, int incy)
{
  for (int c = 0; c < n; c++) y[(long)c * incy] = 0.0;
  for (int r = 0; r < m; r++) {
    double xr = x[(long)r * incx];
    for (int k = ind[r]; k < ind[r + 1]; k++)
      y[(long)col[k] * incy] += val[k] * xr;
  }
}

/* sss_mat.c:40-56 / :58-76: lower-triangle gather + scatter, y not pre-zeroed */
void orc_sss_matvec(int n, const double *val, const double *diag, const int *col, const int *ind,
                    const double *x, int incx, double *y, int incy)
{
  for (int r = 0; r < n; r++) {
    double xr = x[(long)r * incx];
    double acc = 0.0;
    for (int k = ind[r]; k < ind[r + 1]; k++) {
      int c = col[k];
      double a = val[k];
      acc += a * x[(long)c * incx];
      y[(long)c * incy] += a * xr;
    }
    y[(long)r * incy] = acc + diag[r] * xr;
  }
}

static int op_matvec(orc_op *A, int n, double *x, double *y)
{
  switch (A->kind) {
  case K_CSR: orc_csr_matvec(A->m, A->val, A->col, A->ind, x, 1, y, 1); return 0;
  case K_SSS: orc_sss_matvec(A->n, A->val, A->diag, A->col, A->ind, x, 1, y, 1); return 0;
  case K_CB:  return A->cb(A->ctx, n, x, y);
  }
  return -1;
}

/* ------------------------------------------------------------------ */
/* preconditioners                                                     */
/* ------------------------------------------------------------------ */

/* preconmodule.c:35-54: y = x.*dinv, then (steps-1) damped-Jacobi corrections */
static int jacobi_apply(orc_op *K, int n, double *x, double *y)
{
  const double *dinv = K->dinv;
  for (int i = 0; i < n; i++) y[i] = x[i] * dinv[i];
  for (int s = 1; s < K->steps; s++) {
    v_copy(n, y, K->t1);
    if (op_matvec(K->mat, n, K->t1, y)) return -1;
    for (int i = 0; i < n; i++) y[i] = (x[i] - y[i]) * dinv[i] + K->t1[i];
  }
  return 0;
}

/* preconmodule.c:149-193 (omega == 1): symmetric Gauss-Seidel, zero start.
 * b = rhs, x = result, w = work (holds L*x then L^T*x). */
static void symgs_apply(int n, const double *b, double *x, double *w,
                        const double *val, const double *diag, const int *col, const int *ind,
                        int steps)
{
  for (int i = 0; i < n; i++) w[i] = 0.0;
  for (int s = 0; s < steps; s++) {
    for (int r = 0; r < n; r++) {                 /* forward: (L+D) x = b - w */
      double acc = 0.0;
      for (int k = ind[r]; k < ind[r + 1]; k++) acc += val[k] * x[col[k]];
      x[r] = (b[r] - w[r] - acc) / diag[r];
      w[r] = acc;
    }
    for (int i = 0; i < n; i++) { x[i] = w[i]; w[i] = 0.0; }
    for (int r = n - 1; r >= 0; r--) {            /* backward: (L^T+D) x = b - (Lx) */
      x[r] = (b[r] - x[r] - w[r]) / diag[r];
      double xr = x[r];
      for (int k = ind[r]; k < ind[r + 1]; k++) w[col[k]] += val[k] * xr;
    }
  }
}

/* preconmodule.c:95-143 (omega != 1): SSOR with relaxation.
 * h and t are work vectors (reference names: h, temp). */
static void ssor_apply(int n, const double *b, double *x, double *h, double *t,
                       const double *val, const double *diag, const int *col, const int *ind,
                       double omega, int steps)
{
  for (int s = 0; s < steps; s++) {
    if (s == 0)
      for (int i = 0; i < n; i++) t[i] = omega * b[i];
    else
      for (int i = 0; i < n; i++) t[i] = (1.0 - omega) * x[i] * diag[i] + h[i] + omega * b[i];
    for (int r = 0; r < n; r++) {
      double acc = 0.0;
      for (int k = ind[r]; k < ind[r + 1]; k++) acc -= val[k] * x[col[k]];
      h[r] = omega * acc;
      x[r] = (t[r] + h[r]) / diag[r];
    }
    for (int i = 0; i < n; i++) {
      t[i] = (1.0 - omega) * x[i] * diag[i] + h[i] + omega * b[i];
      h[i] = 0.0;
    }
    for (int r = n - 1; r >= 0; r--) {
      h[r] = omega * h[r];
      x[r] = (t[r] + h[r]) / diag[r];
      double xr = x[r];
      for (int k = ind[r]; k < ind[r + 1]; k++) h[col[k]] -= val[k] * xr;
    }
  }
}

static int op_precon(orc_op *K, int n, double *x, double *y)
{
  switch (K->kind) {
  case K_JACOBI: return jacobi_apply(K, n, x, y);
  case K_SSOR: {                         /* SSOR_precon, preconmodule.c:199-223 */
    orc_op *S = K->mat;
    if (K->omega == 1.0)
      symgs_apply(n, x, y, K->t1, S->val, S->diag, S->col, S->ind, K->steps);
    else
      ssor_apply(n, x, y, K->t1, K->t2, S->val, S->diag, S->col, S->ind, K->omega, K->steps);
    return 0;
  }
  case K_CB: return K->cb(K->ctx, n, x, y);
  }
  return -1;
}

/* ------------------------------------------------------------------ */
/* handles                                                             */
/* ------------------------------------------------------------------ */
void *orc_csr_new(int m, int n, const double *val, const int *col, const int *ind)
{
  orc_op *A = calloc(1, sizeof *A);
  A->kind = K_CSR; A->m = m; A->n = n; A->val = val; A->col = col; A->ind = ind;
  return A;
}

void *orc_sss_new(int n, const double *val, const double *diag, const int *col, const int *ind)
{
  orc_op *A = calloc(1, sizeof *A);
  A->kind = K_SSS; A->m = n; A->n = n; A->val = val; A->diag = diag; A->col = col; A->ind = ind;
  return A;
}

void *orc_callback_new(int n, orc_cb cb, void *ctx)
{
  orc_op *A = calloc(1, sizeof *A);
  A->kind = K_CB; A->m = n; A->n = n; A->cb = cb; A->ctx = ctx;
  return A;
}

/* newJacobiObject, preconmodule.c:389-401: dinv = omega / d; a diagonal entry
 * with 1.0 + d == 1.0 is rejected ("diagonal element close to zero") -> NULL */
void *orc_jacobi_new(void *mat, int n, const double *diag, double omega, int steps)
{
  orc_op *K = calloc(1, sizeof *K);
  K->kind = K_JACOBI; K->m = K->n = n; K->mat = mat; K->omega = omega; K->steps = steps;
  K->dinv = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  K->t1 = malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    if (1.0 + diag[i] == 1.0) { free(K->dinv); free(K->t1); free(K); return NULL; }
    K->dinv[i] = omega / diag[i];
  }
  return K;
}

void *orc_ssor_new(void *sss, double omega, int steps)
{
  orc_op *S = sss;
  orc_op *K = calloc(1, sizeof *K);
  K->kind = K_SSOR; K->m = K->n = S->n; K->mat = S; K->omega = omega; K->steps = steps;
  K->t1 = malloc(sizeof(double) * (size_t)(S->n > 0 ? S->n : 1));
  K->t2 = malloc(sizeof(double) * (size_t)(S->n > 0 ? S->n : 1));
  return K;
}

void orc_op_free(void *h)
{
  orc_op *A = h;
  if (!A) return;
  free(A->dinv); free(A->t1); free(A->t2); free(A);
}

int orc_matvec(void *A, int n, double *x, double *y) { return op_matvec(A, n, x, y); }
int orc_precon(void *K, int n, double *x, double *y) { return op_precon(K, n, x, y); }

/* ------------------------------------------------------------------ */
/* solvers                                                             */
/* ------------------------------------------------------------------ */
#define MV(x, y)   do { if (op_matvec(A, n, (x), (y))) return -1; } while (0)
#define PC(x, y)   do { if (K) { if (op_precon(K, n, (x), (y))) return -1; } \
                        else v_copy(n, (x), (y)); } while (0)

/* pcg.c:22-171.  flag: 0 ok, -1 maxit, -2 rho==0, -5 stagnation, -6 beta/pq==0.
 * iter == maxit+1 when the loop runs out (:91,:165). */
static int pcg(orc_op *A, orc_op *K, int n, const double *b, double *x, double tol, int maxit,
               int *flag, int *iter, double *relres, double *wk)
{
  double *r = wk, *z = wk + n, *p = wk + 2 * (size_t)n, *q = wk + 3 * (size_t)n;
  double bnorm = v_nrm2(n, b);
  if (bnorm == 0.0) {
    for (int i = 0; i < n; i++) x[i] = 0.0;
    *flag = 0; *relres = 0.0; *iter = 0;
    return 0;
  }
  *flag = -1;
  double thresh = tol * bnorm;
  MV(x, r);
  for (int i = 0; i < n; i++) r[i] = b[i] - r[i];
  double rnorm = v_nrm2(n, r);
  if (rnorm <= thresh) { *flag = 0; *relres = rnorm / bnorm; *iter = 0; return 0; }

  double rho = 1.0, rho_prev, alpha, beta, pq;
  int stalled = 0, it;
  for (it = 1; it <= maxit; it++) {
    PC(r, z);
    rho_prev = rho;
    rho = v_dot(n, r, z);
    if (rho == 0.0) { *flag = -2; break; }
    if (it == 1) {
      v_copy(n, z, p);
    } else {
      beta = rho / rho_prev;
      if (beta == 0.0) { *flag = -6; break; }
      for (int i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
    }
    MV(p, q);
    pq = v_dot(n, p, q);
    if (pq == 0.0) { *flag = -6; break; }
    alpha = rho / pq;
    if (alpha == 0.0) stalled = 1;
    if (!stalled) {                       /* :126-139 stagnation scan */
      double big = 0.0;
      for (int i = 0; i < n; i++) {
        if (x[i] != 0.0) {
          double rel = fabs(alpha * p[i] / x[i]);
          if (rel > big) big = rel;
        } else if (p[i] != 0.0) {
          big = 1.0;
        }
      }
      stalled = (1.0 + big == 1.0);
    }
    v_axpy(n, alpha, p, x);
    v_axpy(n, -alpha, q, r);
    rnorm = v_nrm2(n, r);
    if (rnorm <= thresh) { *flag = 0; break; }
    if (stalled) { *flag = -5; break; }
  }
  *iter = it;
  *relres = rnorm / bnorm;
  return 0;
}

/* minres.c:43-200.  returns 0 / -1 (not converged) / -3 (K not SPD) / -6 (r1==0);
 * on -3/-6 *relres is left untouched (the reference never writes it). */
static int minres(orc_op *A, orc_op *K, int n, const double *b, double *x, double tol, int maxit,
                  int *iter, double *relres, double *wk)
{
  size_t N = (size_t)n;
  double *vh_old = wk, *vh = wk + N, *y = wk + 2 * N, *w = wk + 3 * N, *w_old = wk + 4 * N,
         *v = wk + 5 * N, *av = wk + 6 * N;
  *iter = 0;
  for (int i = 0; i < n; i++) vh_old[i] = 0.0;
  MV(x, vh);
  for (int i = 0; i < n; i++) vh[i] = b[i] - vh[i];
  double r0 = v_nrm2(n, vh);
  PC(vh, y);
  double beta = v_dot(n, vh, y);
  if (beta < 0.0) return -3;
  beta = sqrt(beta);
  double beta_old = 1.0, c = 1.0, c_old = 1.0, s = 0.0, s_old = 0.0;
  for (int i = 0; i < n; i++) w[i] = 0.0;
  for (int i = 0; i < n; i++) w_old[i] = 0.0;
  double eta = beta, rmr = r0;

  for (;;) {
    if (*iter >= maxit || rmr < tol * r0) break;
    *iter += 1;
    for (int i = 0; i < n; i++) v[i] = y[i] / beta;
    v_copy(n, vh, y);
    MV(v, av);
    double alpha = v_dot(n, v, av);
    double k1 = alpha / beta, k2 = beta / beta_old;
    for (int i = 0; i < n; i++) vh[i] = av[i] - k1 * vh[i] - k2 * vh_old[i];
    v_copy(n, y, vh_old);
    PC(vh, y);
    beta_old = beta;
    beta = v_dot(n, vh, y);
    if (beta < 0.0) return -3;
    beta = sqrt(beta);

    double c_oold = c_old, s_oold = s_old;
    c_old = c; s_old = s;
    double r1_hat = c_old * alpha - c_oold * s_old * beta_old;
    double r1 = sqrt(r1_hat * r1_hat + beta * beta);
    double r2 = s_old * alpha + c_oold * c_old * beta_old;
    double r3 = s_oold * beta_old;
    if (r1 == 0.0) return -6;
    c = r1_hat / r1;
    s = beta / r1;
    for (int i = 0; i < n; i++) {
      double keep = w[i];
      w[i] = (v[i] - r3 * w_old[i] - r2 * keep) / r1;
      w_old[i] = keep;
    }
    double step = c * eta;
    for (int i = 0; i < n; i++) x[i] += step * w[i];
    eta = -s * eta;
    rmr *= fabs(s);
  }
  *relres = rmr / r0;
  return (rmr < tol * r0) ? 0 : -1;
}

/* qmrs.c:34-154.  x0 is ignored (x zeroed, :71-77); err = tau/|b|;
 * 0 iff err < tol else -1; -2 delta==0; -6 zero scalars. */
static int qmrs(orc_op *A, orc_op *K, int n, const double *b, double *x, double tol, int maxit,
                int *iter, double *err, double *wk)
{
  size_t N = (size_t)n;
  double *t0 = wk, *p = wk + N, *d = wk + 2 * N, *v = wk + 3 * N, *t = wk + 4 * N, *g = wk + 5 * N;
  v_copy(n, b, v);
  double rho0 = v_nrm2(n, v);
  double tau = rho0;
  for (int i = 0; i < n; i++) { v[i] /= rho0; p[i] = 0.0; g[i] = 0.0; d[i] = 0.0; x[i] = 0.0; }
  double c0 = 1.0, eps = 1.0, xi = 1.0, theta0 = 0.0, eta = -1.0, res0 = rho0;
  *err = 1.0; *iter = 0;
  while (*err > tol && *iter < maxit) {
    ++*iter;
    if (eps == 0.0) return -6;
    PC(v, t0);
    double delta = v_dot(n, t0, v);
    if (delta == 0.0) return -2;
    double cc = xi * (delta / eps);
    for (int i = 0; i < n; i++) { p[i] = v[i] - p[i] * cc; g[i] = t0[i] - g[i] * cc; }
    MV(g, t);
    eps = v_dot(n, g, t);
    double beta = eps / delta;
    for (int i = 0; i < n; i++) v[i] = t[i] - v[i] * beta;
    double rho1 = v_nrm2(n, v);
    xi = rho1;
    if (c0 * fabs(beta) == 0.0) return -6;
    double theta = rho1 / (c0 * fabs(beta));
    double c1 = 1.0 / sqrt(theta * theta + 1.0);
    if (beta * (c0 * c0) == 0.0) return -6;
    eta = -eta * rho0 * (c1 * c1) / (beta * (c0 * c0));
    tau = tau * theta * c1;
    if (rho1 == 0.0) return -6;
    double tc = theta0 * c1;
    cc = tc * tc;
    double inv = 1.0 / rho1;
    for (int i = 0; i < n; i++) { d[i] = p[i] * eta + d[i] * cc; x[i] += d[i]; v[i] *= inv; }
    if (xi == 0.0) return -6;
    rho0 = rho1;
    *err = tau / res0;
    c0 = c1;
    theta0 = theta;
  }
  if (K) { if (op_precon(K, n, x, t0)) return -1; v_copy(n, t0, x); }
  return (*err < tol) ? 0 : -1;
}

/* bicgstab.c:182-300.  *info starts at -6 and stays there on breakdown
 * (rho==0 / omega==0 -> return -1 with relres unset). */
static int bicgstab(orc_op *A, orc_op *K, int n, const double *b, double *x, double tol, int maxit,
                    int *info, int *iter, double *relres, double *wk)
{
  size_t N = (size_t)n;
  *info = -6;
  double bnorm = v_nrm2(n, b);
  if (bnorm == 0.0) {
    for (int i = 0; i < n; i++) x[i] = 0.0;
    *info = 0; *relres = 0.0; *iter = 0;
    return 0;
  }
  double *r = wk, *rh = wk + N, *p = wk + 2 * N, *ph = wk + 3 * N, *v = wk + 4 * N,
         *s = wk + 5 * N, *sh = wk + 6 * N, *t = wk + 7 * N;
  double omega = 0.0, beta = 0.0, alpha = 0.0, rho2 = 0.0, rho1, res, res0;
  MV(x, r);
  for (int i = 0; i < n; i++) r[i] = b[i] + -r[i];
  res0 = v_nrm2(n, r);
  v_copy(n, r, rh);
  *iter = 0;
  do {
    ++*iter;
    rho1 = v_dot(n, rh, r);
    if (rho1 == 0.0) return -1;
    if (*iter == 1) {
      v_copy(n, r, p);
    } else {
      beta = (rho1 / rho2) * (alpha / omega);
      for (int i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    }
    PC(p, ph);
    MV(ph, v);
    alpha = rho1 / v_dot(n, rh, v);
    for (int i = 0; i < n; i++) s[i] = r[i] + (-alpha) * v[i];
    PC(s, sh);
    MV(sh, t);
    omega = v_dot(n, t, s) / v_dot(n, t, t);
    for (int i = 0; i < n; i++) x[i] = x[i] + alpha * ph[i] + omega * sh[i];
    for (int i = 0; i < n; i++) r[i] = s[i] - omega * t[i];
    res = v_nrm2(n, r);
    if (omega == 0.0) return -1;
    rho2 = rho1;
  } while (res / res0 > tol && *iter < maxit);
  *relres = res / res0;
  *info = (*relres >= tol) ? -1 : 0;
  return 0;
}

/* cgs.c:13-111.  0-based iteration counter; stop on |r|^2 < tol^2 |b|^2. */
static int cgs(orc_op *A, orc_op *K, int n, const double *b, double *x, double tol, int maxit,
               int *iter, double *res, double *wk)
{
  size_t N = (size_t)n;
  double *r0 = wk, *r = wk + N, *p = wk + 2 * N, *q = wk + 3 * N, *u = wk + 4 * N,
         *v = wk + 5 * N, *ta = wk + 6 * N, *tb = wk + 7 * N;
  double tol2 = tol * tol;
  *iter = 0;
  MV(x, ta);
  v_copy(n, b, r0);
  v_axpy(n, -1.0, ta, r0);
  v_copy(n, r0, r); v_copy(n, r0, u); v_copy(n, r0, p);
  double rho = v_dot(n, r0, r0);
  double b2 = v_dot(n, b, b);
  if (rho < b2 * tol2) { *res = sqrt(rho / b2); return 0; }
  for (; *iter < maxit; ++*iter) {
    if (K) { if (op_precon(K, n, p, ta)) return -1; MV(ta, v); }
    else MV(p, v);
    double alpha = rho / v_dot(n, v, r0);
    double neg = -alpha;
    v_copy(n, u, q);  v_axpy(n, neg, v, q);
    v_copy(n, u, ta); v_axpy(n, 1.0, q, ta);
    PC(ta, tb);
    v_axpy(n, alpha, tb, x);
    MV(tb, ta);
    v_axpy(n, neg, ta, r);
    *res = v_dot(n, r, r);
    if (*res < b2 * tol2) { *res = sqrt(*res / b2); return 0; }
    double rho_new = v_dot(n, r, r0);
    double beta = rho_new / rho;
    rho = rho_new;
    v_copy(n, r, u);  v_axpy(n, beta, q, u);
    v_copy(n, q, ta); v_axpy(n, beta, p, ta);
    v_copy(n, u, p);  v_axpy(n, beta, ta, p);
  }
  *res = sqrt(*res / b2);
  return -1;
}

/* gmres.c:47-71 plane rotations */
static void rot_make(double a, double b, double *c, double *s)
{
  if (b == 0.0) { *c = 1.0; *s = 0.0; }
  else if (fabs(b) > fabs(a)) { double t = a / b; *s = 1.0 / sqrt(1.0 + t * t); *c = t * *s; }
  else { double t = b / a; *c = 1.0 / sqrt(1.0 + t * t); *s = t * *c; }
}
static void rot_apply(double *a, double *b, double c, double s)
{
  double t = c * *a + s * *b;
  *b = -s * *a + c * *b;
  *a = t;
}

/* gmres.c:73-206.  Restarted flexible GMRES(m), MGS; always returns 0;
 * relres = true |Ax-b| / |r0| at exit. */
static int gmres(orc_op *A, orc_op *K, int n, const double *b, double *x, double tol, int maxit,
                 int m, int *iter_out, double *relres)
{
  size_t N = (size_t)n;
  int ld = m + 1;
  double *H = malloc(sizeof(double) * (size_t)m * (size_t)(m + 1));
  double *g = malloc(sizeof(double) * (size_t)(m + 1));
  double *cs = malloc(sizeof(double) * (size_t)m), *sn = malloc(sizeof(double) * (size_t)m);
  double *V = malloc(sizeof(double) * N * (size_t)(m + 1));
  double *W = malloc(sizeof(double) * N * (size_t)m);
#define HH(i, j) H[(size_t)(j) * ld + (i)]
  int rc = 0;
  double bnorm = v_nrm2(n, b);
  if (bnorm == 0.0) {
    for (int i = 0; i < n; i++) x[i] = 0.0;
    *relres = 0.0; *iter_out = 0;
    goto done;                            /* (the reference leaks here, :100-108) */
  }
  int iter = 0, i;
  double beta, res0 = 0.0, rel = 0.0;
  do {
    if (op_matvec(A, n, x, V)) { rc = -1; goto done; }
    v_axpy(n, -1.0, b, V);
    beta = sqrt(v_dot(n, V, V));
    v_scal(n, -1.0 / beta, V);
    if (iter == 0) res0 = beta;
    for (int j = 1; j < m + 1; j++) g[j] = 0.0;
    g[0] = beta;
    i = -1;
    do {
      i++; iter++;
      double *vi = V + (size_t)i * N, *vn = V + (size_t)(i + 1) * N, *wi = W + (size_t)i * N;
      if (K) { if (op_precon(K, n, vi, wi)) { rc = -1; goto done; } }
      else v_copy(n, vi, wi);
      if (op_matvec(A, n, wi, vn)) { rc = -1; goto done; }
      for (int k = 0; k <= i; k++) {
        HH(k, i) = v_dot(n, vn, V + (size_t)k * N);
        v_axpy(n, -HH(k, i), V + (size_t)k * N, vn);
      }
      HH(i + 1, i) = sqrt(v_dot(n, vn, vn));
      v_scal(n, 1.0 / HH(i + 1, i), vn);
      for (int k = 0; k < i; k++) rot_apply(&HH(k, i), &HH(k + 1, i), cs[k], sn[k]);
      rot_make(HH(i, i), HH(i + 1, i), &cs[i], &sn[i]);
      rot_apply(&HH(i, i), &HH(i + 1, i), cs[i], sn[i]);
      rot_apply(&g[i], &g[i + 1], cs[i], sn[i]);
      rel = fabs(g[i + 1]) / res0;
      if (rel <= tol) break;
    } while (i + 1 < m && iter + 1 <= maxit);
    for (int j = i; j >= 0; j--) {
      g[j] /= HH(j, j);
      for (int k = j - 1; k >= 0; k--) g[k] -= HH(k, j) * g[j];
    }
    for (int j = 0; j <= i; j++) v_axpy(n, g[j], W + (size_t)j * N, x);
  } while (rel > tol && iter + 1 <= maxit);

  if (op_matvec(A, n, x, V)) { rc = -1; goto done; }
  v_axpy(n, -1.0, b, V);
  beta = sqrt(v_dot(n, V, V));
  *iter_out = iter;
  *relres = beta / res0;
done:
  free(H); free(g); free(cs); free(sn); free(V); free(W);
  return rc;
#undef HH
}

/* Same conventions as the wrappers in itsolversmodule.c:32-586:
 * solver ids 0 pcg, 1 minres, 2 qmrs, 3 bicgstab, 4 cgs, 5 gmres;
 * iter/relres pre-set to -1/NaN where the reference leaves them unset. */
int orc_solve(int solver, void *Ah, void *Kh, int n, const double *b, double *x,
              double tol, int maxit, int dim, int *info, int *iter, double *relres)
{
  orc_op *A = Ah, *K = Kh;
  size_t N = (size_t)(n > 0 ? n : 1);
  double *wk = NULL;
  int rc = 0;
  *info = 0; *iter = -1; *relres = NAN;
  switch (solver) {
  case 0: wk = malloc(4 * N * sizeof(double));
          rc = pcg(A, K, n, b, x, tol, maxit, info, iter, relres, wk); break;
  case 1: wk = malloc(7 * N * sizeof(double));
          *info = minres(A, K, n, b, x, tol, maxit, iter, relres, wk); break;
  case 2: wk = malloc(6 * N * sizeof(double));
          *info = qmrs(A, K, n, b, x, tol, maxit, iter, relres, wk); break;
  case 3: wk = malloc(8 * N * sizeof(double));
          rc = bicgstab(A, K, n, b, x, tol, maxit, info, iter, relres, wk); break;
  case 4: wk = malloc(8 * N * sizeof(double));
          *info = cgs(A, K, n, b, x, tol, maxit, iter, relres, wk); break;
  case 5: *info = gmres(A, K, n, b, x, tol, maxit, dim, iter, relres); break;
  default: return -100;
  }
  free(wk);
  return rc;
}

const char *orc_describe(void)
{
  return "pysparse_b200 CPU restatement (gcc -O2 -ffp-contract=off, sequential BLAS-1)";
}
