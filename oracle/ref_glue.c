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

const char *ref_describe(void)
{
  return "reference kernels compiled from /root/reference (gcc -O2, no FMA, Netlib BLAS-1 loops)";
}
