/* oracle/ref_glue.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Glue that lets the reference's OWN numerical kernels be called from ctypes
 * without a Python-2 interpreter.  Nothing in here computes anything itself
 * apart from the five BLAS-1 routines the reference expects from an external
 * Fortran BLAS (which is not part of /root/reference).
 *
 * How the reference kernels get in (see oracle/Makefile):
 *   - pysparse/itsolvers/src/{pcg,minres,qmrs,cgs,bicgstab,gmres}.c are
 *     compiled UNMODIFIED, straight from /root/reference.  They reach the
 *     operators through `itsolvers_spmatrix[7]` (SpMatrix_Matvec) and `[8]`
 *     (SpMatrix_Precon)  (pysparse/include/spmatrix_api.h:160-166); this file
 *     provides that 16-slot table with plain C functions in those two slots.
 *     The `PyObject *` the kernels carry around is then only an opaque handle
 *     (a `ref_op *` below).
 *   - the Python-free kernel line ranges
 *         pysparse/sparse/src/csr_mat.c:17-72,74-106   (csr matvec / transp)
 *         pysparse/sparse/src/sss_mat.c:40-76          (sss matvec)
 *         pysparse/precon/src/preconmodule.c:35-54     (jacobi)
 *         pysparse/precon/src/preconmodule.c:95-193    (ssor / symgs)
 *         pysparse/eigen/src/correq.c:1-236            (JDSYM correction-equation operator:
 *                                                       object layout, projections, right / matvec / precon)
 *         pysparse/eigen/src/orthopack.c:124-135       (mgs, which CorrEq_right calls)
 *     are extracted with `sed -n` into a scratch directory at build time and
 *     #included below (REF_*_INC).  No reference text is stored in this repo.
 *
 * Output: oracle/_ref/libpsp_ref.so (git-ignored, ships with gpurun).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "Python.h"          /* only for the PyObject typedef used in prototypes */
#include "fortran.h"         /* F77() name mangling, pysparse/include/fortran.h */

/* ------------------------------------------------------------------ */
/* BLAS-1: the reference links an unnamed external Fortran BLAS        */
/* (pysparse/itsolvers/setup.py:18-30).  Netlib reference semantics.   */
/* ------------------------------------------------------------------ */
double F77(ddot)(int *n, double *x, int *incx, double *y, int *incy)
{
  /* Netlib ddot, inc==1 path: clean-up loop over the first n%5 elements,
   * then blocks of five added left to right == plain sequential order. */
  double t = 0.0;
  int i, ix = 0, iy = 0;
  if (*incx == 1 && *incy == 1) {
    for (i = 0; i < *n; i++) t = t + x[i] * y[i];
    return t;
  }
  if (*incx < 0) ix = (1 - *n) * *incx;
  if (*incy < 0) iy = (1 - *n) * *incy;
  for (i = 0; i < *n; i++, ix += *incx, iy += *incy) t = t + x[ix] * y[iy];
  return t;
}

double F77(dnrm2)(int *n, double *x, int *incx)
{
  /* Netlib dnrm2 (LAPACK 3.x reference BLAS): scaled sum of squares. */
  double scale = 0.0, ssq = 1.0, a;
  int i, ix;
  if (*n < 1 || *incx < 1) return 0.0;
  if (*n == 1) return fabs(x[0]);
  for (i = 0, ix = 0; i < *n; i++, ix += *incx) {
    if (x[ix] != 0.0) {
      a = fabs(x[ix]);
      if (scale < a) {
        double q = scale / a;
        ssq = 1.0 + ssq * q * q;
        scale = a;
      } else {
        double q = a / scale;
        ssq = ssq + q * q;
      }
    }
  }
  return scale * sqrt(ssq);
}

void F77(daxpy)(int *n, double *a, double *x, int *incx, double *y, int *incy)
{
  int i, ix = 0, iy = 0;
  if (*n <= 0 || *a == 0.0) return;
  if (*incx < 0) ix = (1 - *n) * *incx;
  if (*incy < 0) iy = (1 - *n) * *incy;
  for (i = 0; i < *n; i++, ix += *incx, iy += *incy) y[iy] = y[iy] + *a * x[ix];
}

void F77(dcopy)(int *n, double *x, int *incx, double *y, int *incy)
{
  int i, ix = 0, iy = 0;
  if (*incx < 0) ix = (1 - *n) * *incx;
  if (*incy < 0) iy = (1 - *n) * *incy;
  for (i = 0; i < *n; i++, ix += *incx, iy += *incy) y[iy] = x[ix];
}

void F77(dscal)(int *n, double *a, double *x, int *incx)
{
  int i, ix;
  if (*n <= 0 || *incx <= 0) return;
  for (i = 0, ix = 0; i < *n; i++, ix += *incx) x[ix] = *a * x[ix];
}

/* ------------------------------------------------------------------ */
/* operator handles (what the kernels see as `PyObject *`)             */
/* ------------------------------------------------------------------ */
enum { OP_CSR = 1, OP_SSS = 2, OP_JACOBI = 3, OP_SSOR = 4, OP_CALLBACK = 5 };

typedef int (*ref_cb)(void *ctx, int n, double *x, double *y);

typedef struct ref_op {
  int kind;
  int m, n, nnz;
  double *val, *diag;   /* borrowed from the caller (NumPy arrays) */
  int *col, *ind;
  /* preconditioners */
  struct ref_op *mat;
  double *dinv, *temp, *temp2;
  double omega;
  int steps;
  /* callbacks (duck-typed operators in tests) */
  ref_cb cb;
  void *cb_ctx;
} ref_op;

static int glue_matvec(PyObject *h, int nx, double *x, int ny, double *y);
static int glue_precon(PyObject *h, int n, double *x, double *y);

/* the kernels extracted from the reference ------------------------- */
#define SpMatrix_Matvec glue_matvec   /* used by jacobi(), preconmodule.c:47 */
#include REF_CSR_INC
#include REF_SSS_INC
#include REF_JACOBI_INC
#include REF_SSOR_INC
#undef SpMatrix_Matvec

static int glue_matvec(PyObject *h, int nx, double *x, int ny, double *y)
{
  ref_op *a = (ref_op *)h;
  (void)nx; (void)ny;
  switch (a->kind) {
  case OP_CSR:
    csr_matvec_kernel(a->m, x, y, a->val, a->col, a->ind);
    return 0;
  case OP_SSS:
    sss_matvec(a->n, x, y, a->val, a->diag, a->col, a->ind);
    return 0;
  case OP_CALLBACK:
    return a->cb(a->cb_ctx, nx, x, y);
  }
  return -1;
}

static int glue_precon(PyObject *h, int n, double *x, double *y)
{
  ref_op *k = (ref_op *)h;
  switch (k->kind) {
  case OP_JACOBI:
    /* Jacobi_precon, preconmodule.c:60-80 */
    return jacobi((PyObject *)k->mat, k->n, k->dinv, k->steps, x, y, k->temp);
  case OP_SSOR:
    /* SSOR_precon, preconmodule.c:199-223 */
    if (k->omega == 1.0)
      symgs_kernel(k->n, x, y, k->temp, k->mat->val, k->mat->diag,
                   k->mat->col, k->mat->ind, k->steps);
    else
      ssor_kernel(k->n, x, y, k->temp, k->temp2, k->mat->val, k->mat->diag,
                  k->mat->col, k->mat->ind, k->omega, k->steps);
    return 0;
  case OP_CALLBACK:
    return k->cb(k->cb_ctx, n, x, y);
  }
  (void)n;
  return -1;
}

/* the 16-slot table of pysparse/include/spmatrix_api.h:10-72; only the two
 * slots the solver kernels dereference are populated. */
static void *glue_table[16] = {
  0, 0, 0, 0, 0, 0, 0, (void *)glue_matvec, (void *)glue_precon,
  0, 0, 0, 0, 0, 0, 0
};
void **itsolvers_spmatrix = glue_table;   /* SPMATRIX_UNIQUE_SYMBOL, pcg.c:4 */

/* ------------------------------------------------------------------ */
/* exported, ctypes-friendly surface                                   */
/* ------------------------------------------------------------------ */
void *ref_csr_new(int m, int n, int nnz, double *val, int *col, int *ind)
{
  ref_op *a = (ref_op *)calloc(1, sizeof *a);
  a->kind = OP_CSR; a->m = m; a->n = n; a->nnz = nnz;
  a->val = val; a->col = col; a->ind = ind;
  return a;
}

void *ref_sss_new(int n, int nnz, double *val, double *diag, int *col, int *ind)
{
  ref_op *a = (ref_op *)calloc(1, sizeof *a);
  a->kind = OP_SSS; a->m = n; a->n = n; a->nnz = nnz;
  a->val = val; a->diag = diag; a->col = col; a->ind = ind;
  return a;
}

void *ref_callback_new(int n, ref_cb cb, void *ctx)
{
  ref_op *a = (ref_op *)calloc(1, sizeof *a);
  a->kind = OP_CALLBACK; a->m = n; a->n = n; a->cb = cb; a->cb_ctx = ctx;
  return a;
}

/* newJacobiObject, preconmodule.c:352-412: dinv[i] = omega / A[i,i];
 * returns NULL for "diagonal element close to zero" (1.0 + d == 1.0). */
void *ref_jacobi_new(void *mat, int n, const double *diag, double omega, int steps)
{
  int i;
  ref_op *k = (ref_op *)calloc(1, sizeof *k);
  k->kind = OP_JACOBI; k->n = n; k->m = n; k->mat = (ref_op *)mat;
  k->omega = omega; k->steps = steps;
  k->dinv = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  k->temp = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  for (i = 0; i < n; i++) {
    double d = diag[i];
    if (1.0 + d == 1.0) { free(k->dinv); free(k->temp); free(k); return NULL; }
    k->dinv[i] = omega / d;
  }
  return k;
}

/* newSSORObject, preconmodule.c:414-459 */
void *ref_ssor_new(void *sss, double omega, int steps)
{
  ref_op *s = (ref_op *)sss;
  ref_op *k = (ref_op *)calloc(1, sizeof *k);
  int n = s->n;
  k->kind = OP_SSOR; k->n = n; k->m = n; k->mat = s;
  k->omega = omega; k->steps = steps;
  k->temp = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  k->temp2 = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  return k;
}

void ref_op_free(void *h)
{
  ref_op *a = (ref_op *)h;
  if (!a) return;
  if (a->kind == OP_JACOBI) free(a->dinv);
  if (a->kind == OP_JACOBI || a->kind == OP_SSOR) { free(a->temp); free(a->temp2); }
  free(a);
}

/* CSRMat_matvec / CSRMat_matvec_transp, csr_mat.c:114-163 (contiguous or
 * strided, selected like the wrappers do) */
void ref_csr_matvec(void *h, double *x, int incx, double *y, int incy)
{
  ref_op *a = (ref_op *)h;
  if (incx == 1 && incy == 1)
    csr_matvec_kernel(a->m, x, y, a->val, a->col, a->ind);
  else
    csr_matvec_kernel_stride(a->m, x, incx, y, incy, a->val, a->col, a->ind);
}

void ref_csr_matvec_transp(void *h, double *x, int incx, double *y, int incy)
{
  ref_op *a = (ref_op *)h;
  if (incx == 1 && incy == 1)
    csr_matvec_transp_kernel(a->m, a->n, x, y, a->val, a->col, a->ind);
  else
    csr_matvec_transp_kernel_stride(a->m, a->n, x, incx, y, incy, a->val, a->col, a->ind);
}

/* SSSMat_matvec, sss_mat.c:78-96 */
void ref_sss_matvec(void *h, double *x, int incx, double *y, int incy)
{
  ref_op *a = (ref_op *)h;
  if (incx == 1 && incy == 1)
    sss_matvec(a->n, x, y, a->val, a->diag, a->col, a->ind);
  else
    sss_matvec_stride(a->n, x, incx, y, incy, a->val, a->diag, a->col, a->ind);
}

int ref_precon(void *k, int n, double *x, double *y)
{
  return glue_precon((PyObject *)k, n, x, y);
}

/* the six kernels, compiled from the reference sources as they are */
int Itsolvers_pcg_kernel(int, double *, double *, double, int, int, int *, double *, int *, double *, PyObject *, PyObject *);
int Itsolvers_bicgstab_kernel(int, double *, double *, double, int, int, int *, double *, int *, double *, PyObject *, PyObject *);
int Itsolvers_minres_kernel(int, double, int, int *, double *, int, double *, double *, double *, PyObject *, PyObject *);
int Itsolvers_gmres_kernel(int, double, int, int *, double *, int, double *, double *, double *, PyObject *, PyObject *);
int Itsolvers_qmrs_kernel(int, double *, double *, double *, double, int, int *, double *, PyObject *, PyObject *);
int Itsolvers_cgs_kernel(int, double *, double *, int, double, double *, int *, double *, PyObject *, PyObject *);

enum { S_PCG = 0, S_MINRES = 1, S_QMRS = 2, S_BICGSTAB = 3, S_CGS = 4, S_GMRES = 5 };

/* Mirrors the six wrappers of pysparse/itsolvers/src/itsolversmodule.c
 * (workspace sizes :90,:182,:278,:373,:467,:560; where `info` comes from).
 * `relres` / `iter` are pre-set to NaN / -1 so that the paths on which the
 * reference leaves them uninitialised are at least deterministic here. */
int ref_solve(int solver, void *A, void *K, int n, double *b, double *x,
              double tol, int maxit, int dim, int *info, int *iter, double *relres)
{
  double *work;
  int rc = 0;
  size_t nn = (size_t)(n > 0 ? n : 1);
  *info = 0; *iter = -1; *relres = NAN;
  switch (solver) {
  case S_PCG:
    work = (double *)malloc(4 * nn * sizeof(double));
    rc = Itsolvers_pcg_kernel(n, x, b, tol, maxit, 0, iter, relres, info, work,
                              (PyObject *)A, (PyObject *)K);
    break;
  case S_BICGSTAB:
    work = (double *)malloc(8 * nn * sizeof(double));
    rc = Itsolvers_bicgstab_kernel(n, x, b, tol, maxit, 0, iter, relres, info, work,
                                   (PyObject *)A, (PyObject *)K);
    break;
  case S_MINRES:
    work = (double *)malloc(7 * nn * sizeof(double));
    *info = Itsolvers_minres_kernel(n, tol, maxit, iter, relres, 0, x, b, work,
                                    (PyObject *)A, (PyObject *)K);
    break;
  case S_GMRES:
    work = (double *)malloc(nn * sizeof(double));
    *info = Itsolvers_gmres_kernel(n, tol, maxit, iter, relres, dim, x, b, work,
                                   (PyObject *)A, (PyObject *)K);
    break;
  case S_QMRS:
    work = (double *)malloc(6 * nn * sizeof(double));
    *info = Itsolvers_qmrs_kernel(n, b, x, work, tol, maxit, iter, relres,
                                  (PyObject *)A, (PyObject *)K);
    break;
  case S_CGS:
    work = (double *)malloc(8 * nn * sizeof(double));
    *info = Itsolvers_cgs_kernel(n, b, x, maxit, tol, work, iter, relres,
                                 (PyObject *)A, (PyObject *)K);
    break;
  default:
    return -100;
  }
  free(work);
  return rc;
}

/* ------------------------------------------------------------------ */
/* JDSYM correction equation: pysparse/eigen/src/correq.c as it is      */
/* ------------------------------------------------------------------ */
#ifdef REF_CORREQ_INC
#include <assert.h>
/* what jdsymmodule.c:27-31 defines before it #includes correq.c */
static int ONE = 1;
static double DONE = 1.0, DZER = 0.0, DMONE = -1.0;

/* BLAS-2 / LAPACK pieces correq.c expects from the external Fortran libraries: Netlib reference
 * semantics (column-major, unblocked), only the cases correq.c / jdsym.c:533-537 use */
void F77(dgemv)(char *trans, int *m, int *n, double *alpha, double *a, int *lda, double *x, int *incx,
                double *beta, double *y, int *incy, int ltrans)
{
  int i, j;
  (void)incx; (void)incy; (void)ltrans;               /* always 1 here */
  if (*trans == 't' || *trans == 'T') {
    for (j = 0; j < *n; j++) {
      double t = 0.0;
      for (i = 0; i < *m; i++) t = t + a[(size_t)j * *lda + i] * x[i];
      y[j] = (*beta == 0.0) ? *alpha * t : *alpha * t + *beta * y[j];
    }
  } else {
    if (*beta == 0.0) for (i = 0; i < *m; i++) y[i] = 0.0;
    else if (*beta != 1.0) for (i = 0; i < *m; i++) y[i] = *beta * y[i];
    for (j = 0; j < *n; j++) {
      double t = *alpha * x[j];
      if (t != 0.0)
        for (i = 0; i < *m; i++) y[i] = y[i] + t * a[(size_t)j * *lda + i];
    }
  }
}

static void glue_dgetf2(int n, double *a, int lda, int *ipiv, int *info)
{
  int i, j, c;
  *info = 0;
  for (j = 0; j < n; j++) {
    int p = j;
    for (i = j + 1; i < n; i++)
      if (fabs(a[(size_t)j * lda + i]) > fabs(a[(size_t)j * lda + p])) p = i;
    ipiv[j] = p + 1;
    if (a[(size_t)j * lda + p] == 0.0) { if (!*info) *info = j + 1; continue; }
    if (p != j)
      for (c = 0; c < n; c++) { double t = a[(size_t)c * lda + j]; a[(size_t)c * lda + j] = a[(size_t)c * lda + p]; a[(size_t)c * lda + p] = t; }
    { double r = 1.0 / a[(size_t)j * lda + j]; for (i = j + 1; i < n; i++) a[(size_t)j * lda + i] = a[(size_t)j * lda + i] * r; }
    for (c = j + 1; c < n; c++) {
      double t = a[(size_t)c * lda + j];
      for (i = j + 1; i < n; i++) a[(size_t)c * lda + i] = a[(size_t)c * lda + i] - a[(size_t)j * lda + i] * t;
    }
  }
}

void F77(dgetrs)(char *trans, int *n, int *nrhs, double *a, int *lda, int *ipiv, double *b, int *ldb, int *info, int ltrans)
{
  int i, j;
  (void)trans; (void)nrhs; (void)ldb; (void)ltrans;   /* "n", one right-hand side */
  *info = 0;
  for (i = 0; i < *n; i++) { int p = ipiv[i] - 1; if (p != i) { double t = b[i]; b[i] = b[p]; b[p] = t; } }
  for (j = 0; j < *n; j++)                              /* L y = b, unit lower */
    if (b[j] != 0.0)
      for (i = j + 1; i < *n; i++) b[i] = b[i] - b[j] * a[(size_t)j * *lda + i];
  for (j = *n - 1; j >= 0; j--)                         /* U x = y */
    if (b[j] != 0.0) {
      b[j] = b[j] / a[(size_t)j * *lda + j];
      for (i = 0; i < j; i++) b[i] = b[i] - b[j] * a[(size_t)j * *lda + i];
    }
}

/* the operator sees its vectors as NumPy arrays; only ->data is dereferenced */
typedef struct { char *data; } PyArrayObject;
#define SPMATRIX_PARSE_ARGS_ARR_ARR(args, xp, yp, n1, n2) \
  do { xp = ((PyArrayObject **)(args))[0]; yp = ((PyArrayObject **)(args))[1]; } while (0)
#define SpMatrix_Matvec glue_matvec
#define SpMatrix_Precon glue_precon
#include REF_MGS_INC
#include REF_CORREQ_INC
#undef SpMatrix_Matvec
#undef SpMatrix_Precon

typedef struct { CorrEqObject ce; double *H; int kmax; } ref_correq;

/* newCorrEqObject (correq.c:307-341) + the work arrays jdsym.c:176-190 allocates around it */
void *ref_correq_new(int optype, int n, void *amat, void *mmat, void *prec, double *Q, double *Qm, double *Y, int kmax)
{
  ref_correq *r = (ref_correq *)calloc(1, sizeof *r);
  r->kmax = kmax;
  r->ce.op_type = optype; r->ce.n = n;
  r->ce.amat = (PyObject *)amat; r->ce.mmat = (PyObject *)mmat; r->ce.prec = (PyObject *)prec;
  r->ce.Q = Q; r->ce.Qm = Qm; r->ce.Y = Y;
  r->ce.Hlu = (double *)calloc((size_t)kmax * kmax, sizeof(double));
  r->ce.Hpiv = (int *)calloc((size_t)kmax, sizeof(int));
  r->ce.ldh = kmax;
  r->ce.work = (double *)calloc(2 * (size_t)(n > 0 ? n : 1), sizeof(double));
  r->ce.right = CorrEq_right; r->ce.update = CorrEq_update;
  return r;
}

/* jdsym.c:533-537 (copy H, factor its leading k x k block) followed by CorrEq_update (:569);
 * H: kmax x kmax, column-major, or NULL when there is no preconditioner */
int ref_correq_update(void *h, int k, double theta, const double *H)
{
  ref_correq *r = (ref_correq *)h;
  int info = 0;
  if (H) {
    memcpy(r->ce.Hlu, H, sizeof(double) * (size_t)r->kmax * r->kmax);
    glue_dgetf2(k, r->ce.Hlu, r->kmax, r->ce.Hpiv, &info);
  }
  r->ce.update(&r->ce, k, theta);
  return info;
}

void ref_correq_right(void *h, double *rhs) { ref_correq *r = (ref_correq *)h; r->ce.right(&r->ce, rhs); }

void ref_correq_matvec(void *h, double *x, double *y)
{
  PyArrayObject xa = {(char *)x}, ya = {(char *)y};
  PyArrayObject *args[2] = {&xa, &ya};
  CorrEq_matvec(&((ref_correq *)h)->ce, (PyObject *)args);
}

void ref_correq_precon(void *h, double *x, double *y)
{
  PyArrayObject xa = {(char *)x}, ya = {(char *)y};
  PyArrayObject *args[2] = {&xa, &ya};
  CorrEq_precon(&((ref_correq *)h)->ce, (PyObject *)args);
}

void ref_correq_free(void *h)
{
  ref_correq *r = (ref_correq *)h;
  if (!r) return;
  free(r->ce.Hlu); free(r->ce.Hpiv); free(r->ce.work); free(r);
}
#endif  /* REF_CORREQ_INC */

const char *ref_describe(void)
{
  return "reference kernels compiled from /root/reference (gcc -O2, no FMA, Netlib BLAS-1 loops)";
}
