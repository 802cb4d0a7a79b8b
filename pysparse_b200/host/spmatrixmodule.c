/* spmatrixmodule.c -- module `spmatrix` of pysparse_b200: ll_mat / csr_mat /
 * sss_mat types, the `_C_API` table, host-side structure conversion.
 *
 * Mirrors pysparse/sparse/src/spmatrixmodule.c (module functions :317-357, C-API
 * :86-307), csr_mat.c (methods :114-171, attributes :208-231), sss_mat.c
 * (:78-110, :149-172) and the parts of ll_mat.c users reach the hot path through
 * (get/set :210-457, to_csr :1577-1648, to_sss :1654-1708, matvec :1434-1465,
 * put, find, keys/values/items, copy, scale, shift, norm, generalize, compress).
 *
 * All products are computed on the GPU through include/psp_b200.h; the
 * conversion routines are integer/structure work done on the host and are
 * required to be bit-exact w.r.t. the reference's output arrays.
 */
#define PY_ARRAY_UNIQUE_SYMBOL psp_spmatrix_ARRAY_API
#include "psp_host.h"
#include <math.h>
#include <stdarg.h>

static PyTypeObject LLMatType, CSRMatType, SSSMatType;
static PyObject *SpMatrix_ErrorObject;

/* ------------------------------------------------------------------------- */
/* argument checking of matvec-like calls (include/spmatrix.h:38-54)           */
/* ------------------------------------------------------------------------- */
static int parse_vec_pair(PyObject *args, PyArrayObject **x, PyArrayObject **y, long n1, long n2, int contiguous)
{
  if (!PyArg_ParseTuple(args, "O!O!", &PyArray_Type, x, &PyArray_Type, y)) return -1;
  PyArrayObject *a[2] = {*x, *y};
  long want[2] = {n1, n2};
  for (int k = 0; k < 2; k++) {
    if (PyArray_NDIM(a[k]) != 1 || PyArray_TYPE(a[k]) != NPY_DOUBLE || PyArray_DIM(a[k], 0) != want[k] ||
        (contiguous && !PyArray_IS_C_CONTIGUOUS(a[k]))) {
      PyErr_Format(PyExc_ValueError, contiguous
                   ? "arg %d must be a contiguous 1-dimensional double array of appropriate size."
                   : "arg %d must be a 1-dimensional double array of appropriate size.", k + 1);
      return -1;
    }
  }
  return 0;
}

static long elem_stride(PyArrayObject *a)
{
  return PyArray_DIM(a, 0) > 1 ? (long)(PyArray_STRIDE(a, 0) / (npy_intp)sizeof(double)) : 1;
}

static PyObject *device_matvec(psp_mat *dev, int transp, PyArrayObject *x, PyArrayObject *y)
{
  long incx = elem_stride(x), incy = elem_stride(y);
  if (incx < 1 || incy < 1) {
    PyErr_SetString(PyExc_ValueError, "negative strides are not supported");
    return NULL;
  }
  int st = psp_matvec(dev, transp, (const double *)PyArray_DATA(x), incx, (double *)PyArray_DATA(y), incy);
  if (st != PSP_OK) return psp_raise(st);
  Py_RETURN_NONE;
}

static void mat_capsule_noop(PyObject *cap) { (void)cap; }

/* ========================================================================= */
/* csr_mat                                                                     */
/* ========================================================================= */
static PyObject *newCSRMatObject(int dim[], int nnz)
{
  CSRMatObject *op = PyObject_New(CSRMatObject, &CSRMatType);
  if (!op) return PyErr_NoMemory();
  op->dim[0] = dim[0]; op->dim[1] = dim[1]; op->nnz = nnz;
  op->dev = NULL; op->base[0] = op->base[1] = op->base[2] = NULL;
  op->val = PyMem_New(double, nnz > 0 ? nnz : 1);
  op->col = PyMem_New(int, nnz > 0 ? nnz : 1);
  op->ind = PyMem_New(int, dim[0] + 1);
  if (!op->val || !op->col || !op->ind) {
    PyMem_Free(op->val); PyMem_Free(op->col); PyMem_Free(op->ind);
    PyObject_Del(op);
    return PyErr_NoMemory();
  }
  return (PyObject *)op;
}

static psp_mat *csr_device(CSRMatObject *self)
{
  if (!self->dev) {
    int st = psp_csr_create(self->dim[0], self->dim[1], self->nnz, self->ind, self->col, self->val, &self->dev);
    if (st != PSP_OK) { psp_raise(st); return NULL; }
  }
  return self->dev;
}

static void CSRMat_dealloc(CSRMatObject *self)
{
  if (self->dev) psp_mat_destroy(self->dev);
  if (self->base[0]) { Py_XDECREF(self->base[0]); Py_XDECREF(self->base[1]); Py_XDECREF(self->base[2]); }
  else { PyMem_Free(self->val); PyMem_Free(self->col); PyMem_Free(self->ind); }
  PyObject_Del(self);
}

static PyObject *CSRMat_matvec(CSRMatObject *self, PyObject *args)
{
  PyArrayObject *x, *y;
  if (parse_vec_pair(args, &x, &y, self->dim[1], self->dim[0], 0)) return NULL;
  psp_mat *d = csr_device(self);
  return d ? device_matvec(d, 0, x, y) : NULL;
}

static PyObject *CSRMat_matvec_transp(CSRMatObject *self, PyObject *args)
{
  PyArrayObject *x, *y;
  if (parse_vec_pair(args, &x, &y, self->dim[0], self->dim[1], 0)) return NULL;
  psp_mat *d = csr_device(self);
  return d ? device_matvec(d, 1, x, y) : NULL;
}

static PyObject *CSRMat_get_shape(CSRMatObject *self, void *c) { return Py_BuildValue("(i,i)", self->dim[0], self->dim[1]); }
static PyObject *CSRMat_get_nnz(CSRMatObject *self, void *c) { return PyLong_FromLong(self->nnz); }
static PyObject *CSRMat_get_dev(CSRMatObject *self, void *c)
{
  psp_mat *d = csr_device(self);
  return d ? PyCapsule_New(d, PSP_MAT_CAPSULE, mat_capsule_noop) : NULL;
}

static PyObject *wrap_i32(int *p, npy_intp n, PyObject *owner)
{
  PyObject *a = PyArray_SimpleNewFromData(1, &n, NPY_INT32, p);
  if (a) { Py_INCREF(owner); PyArray_SetBaseObject((PyArrayObject *)a, owner); PyArray_CLEARFLAGS((PyArrayObject *)a, NPY_ARRAY_WRITEABLE); }
  return a;
}
static PyObject *wrap_f64(double *p, npy_intp n, PyObject *owner)
{
  PyObject *a = PyArray_SimpleNewFromData(1, &n, NPY_DOUBLE, p);
  if (a) { Py_INCREF(owner); PyArray_SetBaseObject((PyArrayObject *)a, owner); PyArray_CLEARFLAGS((PyArrayObject *)a, NPY_ARRAY_WRITEABLE); }
  return a;
}

/* read-only views of the storage arrays (pysparse_b200 addition; used by the
 * bit-exactness tests of the conversions) */
static PyObject *CSRMat_arrays(CSRMatObject *self, PyObject *noargs)
{
  return Py_BuildValue("(NNN)", wrap_f64(self->val, self->nnz, (PyObject *)self),
                       wrap_i32(self->col, self->nnz, (PyObject *)self),
                       wrap_i32(self->ind, self->dim[0] + 1, (PyObject *)self));
}

static PyObject *CSRMat_diagonal(CSRMatObject *self, PyObject *noargs);

static PyMethodDef CSRMat_methods[] = {
  {"matvec", (PyCFunction)CSRMat_matvec, METH_VARARGS, "a.matvec(x, y)\n\ncompute the sparse matrix-vector product y := a * x."},
  {"matvec_transp", (PyCFunction)CSRMat_matvec_transp, METH_VARARGS, "a.matvec_transp(x, y)\n\ncompute the sparse matrix-vector product y := a^T * x."},
  {"arrays", (PyCFunction)CSRMat_arrays, METH_NOARGS, "(val, col, ind) read-only views of the CSR storage"},
  {"diagonal", (PyCFunction)CSRMat_diagonal, METH_NOARGS, "dense copy of the diagonal"},
  {NULL, NULL}
};
static PyGetSetDef CSRMat_getset[] = {
  {"shape", (getter)CSRMat_get_shape, NULL, NULL, NULL},
  {"nnz", (getter)CSRMat_get_nnz, NULL, NULL, NULL},
  {"_psp_mat", (getter)CSRMat_get_dev, NULL, "device handle (PyCapsule)", NULL},
  {NULL}
};

static double csr_get(CSRMatObject *self, int i, int j)
{
  for (int k = self->ind[i]; k < self->ind[i + 1]; k++)
    if (self->col[k] == j) return self->val[k];
  return 0.0;
}

static PyObject *CSRMat_diagonal(CSRMatObject *self, PyObject *noargs)
{
  npy_intp n = self->dim[0] < self->dim[1] ? self->dim[0] : self->dim[1];
  PyObject *a = PyArray_SimpleNew(1, &n, NPY_DOUBLE);
  if (!a) return NULL;
  double *d = (double *)PyArray_DATA((PyArrayObject *)a);
  for (int i = 0; i < n; i++) d[i] = csr_get(self, i, i);
  return a;
}

/* ========================================================================= */
/* sss_mat                                                                     */
/* ========================================================================= */
static PyObject *newSSSMatObject(int n, int nnz)
{
  SSSMatObject *op = PyObject_New(SSSMatObject, &SSSMatType);
  if (!op) return PyErr_NoMemory();
  op->n = n; op->nnz = nnz; op->dev = NULL;
  op->base[0] = op->base[1] = op->base[2] = op->base[3] = NULL;
  op->val = PyMem_New(double, nnz > 0 ? nnz : 1);
  op->diag = PyMem_New(double, n > 0 ? n : 1);
  op->col = PyMem_New(int, nnz > 0 ? nnz : 1);
  op->ind = PyMem_New(int, n + 1);
  if (!op->val || !op->diag || !op->col || !op->ind) {
    PyMem_Free(op->val); PyMem_Free(op->diag); PyMem_Free(op->col); PyMem_Free(op->ind);
    PyObject_Del(op);
    return PyErr_NoMemory();
  }
  return (PyObject *)op;
}

static psp_mat *sss_device(SSSMatObject *self)
{
  if (!self->dev) {
    int st = psp_sss_create(self->n, self->nnz, self->ind, self->col, self->val, self->diag, &self->dev);
    if (st != PSP_OK) { psp_raise(st); return NULL; }
  }
  return self->dev;
}

static void SSSMat_dealloc(SSSMatObject *self)
{
  if (self->dev) psp_mat_destroy(self->dev);
  if (self->base[0]) { for (int k = 0; k < 4; k++) Py_XDECREF(self->base[k]); }
  else { PyMem_Free(self->val); PyMem_Free(self->diag); PyMem_Free(self->col); PyMem_Free(self->ind); }
  PyObject_Del(self);
}

/* matvec and matvec_transp are the same function (sss_mat.c:106-110) */
static PyObject *SSSMat_matvec(SSSMatObject *self, PyObject *args)
{
  PyArrayObject *x, *y;
  if (parse_vec_pair(args, &x, &y, self->n, self->n, 0)) return NULL;
  psp_mat *d = sss_device(self);
  return d ? device_matvec(d, 0, x, y) : NULL;
}

static PyObject *SSSMat_get_shape(SSSMatObject *self, void *c) { return Py_BuildValue("(i,i)", self->n, self->n); }
/* nnz reports lower triangle + diagonal (sss_mat.c:155) */
static PyObject *SSSMat_get_nnz(SSSMatObject *self, void *c) { return PyLong_FromLong(self->nnz + self->n); }
static PyObject *SSSMat_get_dev(SSSMatObject *self, void *c)
{
  psp_mat *d = sss_device(self);
  return d ? PyCapsule_New(d, PSP_MAT_CAPSULE, mat_capsule_noop) : NULL;
}
static PyObject *SSSMat_arrays(SSSMatObject *self, PyObject *noargs)
{
  return Py_BuildValue("(NNNN)", wrap_f64(self->val, self->nnz, (PyObject *)self),
                       wrap_f64(self->diag, self->n, (PyObject *)self),
                       wrap_i32(self->col, self->nnz, (PyObject *)self),
                       wrap_i32(self->ind, self->n + 1, (PyObject *)self));
}
static PyObject *SSSMat_diagonal(SSSMatObject *self, PyObject *noargs)
{
  npy_intp n = self->n;
  PyObject *a = PyArray_SimpleNew(1, &n, NPY_DOUBLE);
  if (a && n) memcpy(PyArray_DATA((PyArrayObject *)a), self->diag, sizeof(double) * (size_t)n);
  return a;
}

static double sss_get(SSSMatObject *self, int i, int j)
{
  if (i == j) return self->diag[i];
  if (i < j) { int t = i; i = j; j = t; }
  for (int k = self->ind[i]; k < self->ind[i + 1]; k++)
    if (self->col[k] == j) return self->val[k];
  return 0.0;
}

/* a[i,j] for integer pairs.  (In the reference snapshot this lookup always
 * raises because LLMat_parse_index is a stub, ll_mat.c:3798-3806; the intended
 * behaviour, sss_mat.c:14-28, is implemented here.) */
static PyObject *pair_subscript(PyObject *idx, int n0, int n1, long *i, long *j)
{
  if (!PyTuple_Check(idx) || PyTuple_GET_SIZE(idx) != 2) {
    PyErr_SetString(PyExc_IndexError, "index must be a pair (i, j)");
    return NULL;
  }
  *i = PyLong_AsLong(PyTuple_GET_ITEM(idx, 0));
  *j = PyLong_AsLong(PyTuple_GET_ITEM(idx, 1));
  if (PyErr_Occurred()) { PyErr_Clear(); PyErr_SetString(PyExc_IndexError, "indices must be integers"); return NULL; }
  if (*i < 0) *i += n0;
  if (*j < 0) *j += n1;
  if (*i < 0 || *j < 0 || *i >= n0 || *j >= n1) { PyErr_SetString(PyExc_IndexError, "indices out of range"); return NULL; }
  return idx;
}
static PyObject *SSSMat_subscript(SSSMatObject *self, PyObject *idx)
{
  long i, j;
  if (!pair_subscript(idx, self->n, self->n, &i, &j)) return NULL;
  return PyFloat_FromDouble(sss_get(self, (int)i, (int)j));
}
static PyObject *CSRMat_subscript(CSRMatObject *self, PyObject *idx)
{
  long i, j;
  if (!pair_subscript(idx, self->dim[0], self->dim[1], &i, &j)) return NULL;
  return PyFloat_FromDouble(csr_get(self, (int)i, (int)j));
}
static PyMappingMethods SSSMat_as_mapping = {0, (binaryfunc)SSSMat_subscript, 0};
static PyMappingMethods CSRMat_as_mapping = {0, (binaryfunc)CSRMat_subscript, 0};

static PyMethodDef SSSMat_methods[] = {
  {"matvec", (PyCFunction)SSSMat_matvec, METH_VARARGS, "a.matvec(x, y)\n\ncompute the sparse matrix-vector product y := a * x."},
  {"matvec_transp", (PyCFunction)SSSMat_matvec, METH_VARARGS, "a.matvec_transp(x, y)\n\nsame as matvec (a is symmetric)."},
  {"arrays", (PyCFunction)SSSMat_arrays, METH_NOARGS, "(val, diag, col, ind) read-only views of the SSS storage"},
  {"diagonal", (PyCFunction)SSSMat_diagonal, METH_NOARGS, "dense copy of the diagonal"},
  {NULL, NULL}
};
static PyGetSetDef SSSMat_getset[] = {
  {"shape", (getter)SSSMat_get_shape, NULL, NULL, NULL},
  {"nnz", (getter)SSSMat_get_nnz, NULL, NULL, NULL},
  {"_psp_mat", (getter)SSSMat_get_dev, NULL, "device handle (PyCapsule)", NULL},
  {NULL}
};

/* ========================================================================= */
/* ll_mat                                                                      */
/* ========================================================================= */
static PyObject *SpMatrix_NewLLMatObject(int dim[], int sym, int sizeHint, int storeZeros)
{
  if (dim[0] < 0 || dim[1] < 0) { PyErr_SetString(PyExc_ValueError, "matrix dimension must be non-negative"); return NULL; }
  if (sizeHint < 1) sizeHint = 1;
  LLMatObject *op = PyObject_New(LLMatObject, &LLMatType);
  if (!op) return PyErr_NoMemory();
  op->dim[0] = dim[0]; op->dim[1] = dim[1]; op->issym = sym; op->storeZeros = storeZeros; op->nnz = 0;
  op->version = 1; op->dev_version = 0; op->dev = NULL;
  op->rows = (LLRow *)calloc((size_t)(dim[0] > 0 ? dim[0] : 1), sizeof(LLRow));
  if (!op->rows) { PyObject_Del(op); return PyErr_NoMemory(); }
  return (PyObject *)op;
}

static void LLMat_dealloc(LLMatObject *self)
{
  if (self->dev) psp_mat_destroy(self->dev);
  for (int i = 0; i < self->dim[0]; i++) { free(self->rows[i].col); free(self->rows[i].val); }
  free(self->rows);
  PyObject_Del(self);
}

/* position of column j in row r (lower bound) */
static int row_find(const LLRow *r, int j)
{
  int lo = 0, hi = r->len;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (r->col[mid] < j) lo = mid + 1; else hi = mid; }
  return lo;
}

static int row_insert(LLRow *r, int pos, int j, double x)
{
  if (r->len == r->cap) {
    int ncap = r->cap ? r->cap * 2 : 8;
    int *nc = (int *)realloc(r->col, sizeof(int) * (size_t)ncap);
    if (!nc) return -1;
    r->col = nc;
    double *nv = (double *)realloc(r->val, sizeof(double) * (size_t)ncap);
    if (!nv) return -1;
    r->val = nv;
    r->cap = ncap;
  }
  memmove(r->col + pos + 1, r->col + pos, sizeof(int) * (size_t)(r->len - pos));
  memmove(r->val + pos + 1, r->val + pos, sizeof(double) * (size_t)(r->len - pos));
  r->col[pos] = j; r->val[pos] = x; r->len++;
  return 0;
}

static void row_remove(LLRow *r, int pos)
{
  memmove(r->col + pos, r->col + pos + 1, sizeof(int) * (size_t)(r->len - pos - 1));
  memmove(r->val + pos, r->val + pos + 1, sizeof(double) * (size_t)(r->len - pos - 1));
  r->len--;
}

/* a[i,j]   (SpMatrix_LLMatGetItem, ll_mat.c:210-245) */
static double SpMatrix_LLMatGetItem(LLMatObject *a, int i, int j)
{
  if (i < 0 || i >= a->dim[0] || j < 0 || j >= a->dim[1]) {
    PyErr_SetString(PyExc_IndexError, "indices out of range");
    return 0.0;
  }
  if (a->issym && i < j) { int t = i; i = j; j = t; }
  const LLRow *r = &a->rows[i];
  int pos = row_find(r, j);
  return (pos < r->len && r->col[pos] == j) ? r->val[pos] : 0.0;
}

/* a[i,j] = x   (SpMatrix_LLMatSetItem, ll_mat.c:251-352) */
static int SpMatrix_LLMatSetItem(LLMatObject *a, int i, int j, double x)
{
  if (a->issym && i < j) {
    PyErr_SetString(PyExc_IndexError, "write operation to upper triangle of symmetric matrix");
    return -1;
  }
  if (i < 0 || i >= a->dim[0] || j < 0 || j >= a->dim[1]) {
    PyErr_SetString(PyExc_IndexError, "indices out of range");
    return -1;
  }
  LLRow *r = &a->rows[i];
  int pos = row_find(r, j);
  int found = (pos < r->len && r->col[pos] == j);
  if (x != 0.0 || a->storeZeros == 1) {
    if (found) r->val[pos] = x;
    else {
      if (row_insert(r, pos, j, x)) { PyErr_NoMemory(); return -1; }
      a->nnz++;
    }
  } else if (found) {
    row_remove(r, pos);
    a->nnz--;
  }
  a->version++;
  return 0;
}

/* a[i,j] += x   (SpMatrix_LLMatUpdateItemAdd, ll_mat.c:358-457) */
static int SpMatrix_LLMatUpdateItemAdd(LLMatObject *a, int i, int j, double x)
{
  if (a->issym && i < j) {
    PyErr_SetString(PyExc_IndexError, "write operation to upper triangle of symmetric matrix");
    return -1;
  }
  if (i < 0 || i >= a->dim[0] || j < 0 || j >= a->dim[1]) {
    PyErr_SetString(PyExc_IndexError, "indices out of range");
    return -1;
  }
  if (a->storeZeros == 0 && x == 0.0) return 0;
  LLRow *r = &a->rows[i];
  int pos = row_find(r, j);
  if (pos < r->len && r->col[pos] == j) {
    x += r->val[pos];
    if (a->storeZeros == 0 && x == 0.0) { row_remove(r, pos); a->nnz--; }
    else r->val[pos] = x;
  } else {
    if (row_insert(r, pos, j, x)) { PyErr_NoMemory(); return -1; }
    a->nnz++;
  }
  a->version++;
  return 0;
}

/* column-linked index of the reference (ll_mat.c:135-201): tied to its linked
 * list layout, which this container does not have. */
struct llColIndex;
static int SpMatrix_LLMatBuildColIndex(struct llColIndex **idx, LLMatObject *self, int includeDiagonal)
{
  (void)idx; (void)self; (void)includeDiagonal;
  PyErr_SetString(PyExc_NotImplementedError, "llColIndex is not available in pysparse_b200 (row-array ll_mat)");
  return 1;
}
static void SpMatrix_LLMatDestroyColIndex(struct llColIndex **idx) { (void)idx; }

/* ---- conversions --------------------------------------------------------- */

/* A.to_csr()   (LLMat_to_csr, ll_mat.c:1577-1648): general matrices are copied
 * row by row; symmetric ones are expanded: row i = its stored (lower + diag)
 * entries, then the column-i entries of the rows below it, ascending. */
static PyObject *LLMat_to_csr(LLMatObject *self, PyObject *noargs)
{
  const int m = self->dim[0];
  if (!self->issym) {
    CSRMatObject *op = (CSRMatObject *)newCSRMatObject(self->dim, self->nnz);
    if (!op) return NULL;
    int r = 0;
    op->ind[0] = 0;
    for (int i = 0; i < m; i++) {
      const LLRow *row = &self->rows[i];
      memcpy(op->col + r, row->col, sizeof(int) * (size_t)row->len);
      memcpy(op->val + r, row->val, sizeof(double) * (size_t)row->len);
      r += row->len;
      op->ind[i + 1] = r;
    }
    return (PyObject *)op;
  }
  long nlo = 0, ndiag = 0;
  int *cnt = (int *)calloc((size_t)m + 1, sizeof(int));
  if (!cnt) return PyErr_NoMemory();
  for (int i = 0; i < m; i++) {
    const LLRow *row = &self->rows[i];
    cnt[i] += row->len;
    for (int k = 0; k < row->len; k++) {
      if (row->col[k] < i) { nlo++; cnt[row->col[k]]++; }
      else ndiag++;
    }
  }
  if (2 * nlo + ndiag >= 2147483647L) { free(cnt); PyErr_SetString(PyExc_OverflowError, "too many non-zeros for int32 CSR"); return NULL; }
  CSRMatObject *op = (CSRMatObject *)newCSRMatObject(self->dim, (int)(2 * nlo + ndiag));
  if (!op) { free(cnt); return NULL; }
  op->ind[0] = 0;
  for (int i = 0; i < m; i++) op->ind[i + 1] = op->ind[i] + cnt[i];
  /* cnt becomes the fill cursor of the "upper" part of each row */
  for (int i = 0; i < m; i++) {
    const LLRow *row = &self->rows[i];
    int p = op->ind[i];
    memcpy(op->col + p, row->col, sizeof(int) * (size_t)row->len);
    memcpy(op->val + p, row->val, sizeof(double) * (size_t)row->len);
    cnt[i] = p + row->len;
  }
  for (int i = 0; i < m; i++) {
    const LLRow *row = &self->rows[i];
    for (int k = 0; k < row->len; k++) {
      int j = row->col[k];
      if (j < i) { op->col[cnt[j]] = i; op->val[cnt[j]] = row->val[k]; cnt[j]++; }
    }
  }
  free(cnt);
  return (PyObject *)op;
}

/* A.to_sss()   (LLMat_to_sss, ll_mat.c:1654-1708): strict lower triangle +
 * diagonal; entries above the diagonal are silently dropped. */
static PyObject *LLMat_to_sss(LLMatObject *self, PyObject *noargs)
{
  const int n = self->dim[0];
  if (n != self->dim[1]) { PyErr_SetString(PyExc_ValueError, "Matrix must be square"); return NULL; }
  long nnz = 0;
  for (int i = 0; i < n; i++) nnz += row_find(&self->rows[i], i);
  SSSMatObject *op = (SSSMatObject *)newSSSMatObject(n, (int)nnz);
  if (!op) return NULL;
  int r = 0;
  op->ind[0] = 0;
  for (int i = 0; i < n; i++) {
    const LLRow *row = &self->rows[i];
    int nl = row_find(row, i);
    memcpy(op->col + r, row->col, sizeof(int) * (size_t)nl);
    memcpy(op->val + r, row->val, sizeof(double) * (size_t)nl);
    r += nl;
    op->diag[i] = (nl < row->len && row->col[nl] == i) ? row->val[nl] : 0.0;
    op->ind[i + 1] = r;
  }
  return (PyObject *)op;
}

/* device copy of an ll_mat: CSR for general, SSS for symmetric matrices
 * (the same accumulation orders as ll_matvec_kernel / ll_matvec_kernel_sym,
 * ll_mat.c:1262-1316); rebuilt when the matrix has been modified. */
static psp_mat *ll_device(LLMatObject *self)
{
  if (self->dev && self->dev_version == self->version) return self->dev;
  if (self->dev) { psp_mat_destroy(self->dev); self->dev = NULL; }
  int st;
  if (self->issym) {
    SSSMatObject *s = (SSSMatObject *)LLMat_to_sss(self, NULL);
    if (!s) return NULL;
    st = psp_sss_create(s->n, s->nnz, s->ind, s->col, s->val, s->diag, &self->dev);
    Py_DECREF(s);
  } else {
    CSRMatObject *c = (CSRMatObject *)LLMat_to_csr(self, NULL);
    if (!c) return NULL;
    st = psp_csr_create(c->dim[0], c->dim[1], c->nnz, c->ind, c->col, c->val, &self->dev);
    Py_DECREF(c);
  }
  if (st != PSP_OK) { psp_raise(st); return NULL; }
  self->dev_version = self->version;
  return self->dev;
}

static PyObject *LLMat_matvec(LLMatObject *self, PyObject *args)
{
  PyArrayObject *x, *y;
  if (parse_vec_pair(args, &x, &y, self->dim[1], self->dim[0], 0)) return NULL;
  psp_mat *d = ll_device(self);
  return d ? device_matvec(d, 0, x, y) : NULL;
}
static PyObject *LLMat_matvec_transp(LLMatObject *self, PyObject *args)
{
  PyArrayObject *x, *y;
  if (parse_vec_pair(args, &x, &y, self->dim[0], self->dim[1], 0)) return NULL;
  psp_mat *d = ll_device(self);
  return d ? device_matvec(d, self->issym ? 0 : 1, x, y) : NULL;
}

/* ---- bulk element access -------------------------------------------------- */
static int as_index_array(PyObject *o, PyArrayObject **out)
{
  *out = (PyArrayObject *)PyArray_FROM_OTF(o, NPY_LONG, NPY_ARRAY_IN_ARRAY);
  if (!*out) return -1;
  if (PyArray_NDIM(*out) > 1) { Py_DECREF(*out); *out = NULL; PyErr_SetString(PyExc_ValueError, "index arrays must be 1-dimensional"); return -1; }
  return 0;
}

/* a.put(b, id1=None, id2=None) / a.update_add_at(b, id1, id2)   (ll_mat.c:1847, :2498) */
static PyObject *put_like(LLMatObject *self, PyObject *args, int add)
{
  PyObject *bIn, *id1in = NULL, *id2in = NULL;
  if (!PyArg_ParseTuple(args, add ? "OOO" : "O|OO", &bIn, &id1in, &id2in)) return NULL;
  PyArrayObject *b = (PyArrayObject *)PyArray_FROM_OTF(bIn, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!b) return NULL;
  PyArrayObject *i1 = NULL, *i2 = NULL;
  PyObject *ret = NULL;
  int b_scalar = (PyArray_NDIM(b) == 0);
  if (id1in && id1in != Py_None && as_index_array(id1in, &i1)) goto done;
  if (id2in && id2in != Py_None && as_index_array(id2in, &i2)) goto done;
  long len;
  if (i1) len = (long)PyArray_SIZE(i1);
  else if (i2) len = (long)PyArray_SIZE(i2);
  else if (!b_scalar) len = (long)PyArray_SIZE(b);
  else { PyErr_SetString(PyExc_ValueError, "put: no index arrays and scalar value"); goto done; }
  if ((!b_scalar && PyArray_SIZE(b) != len) || (i1 && PyArray_SIZE(i1) != len) || (i2 && PyArray_SIZE(i2) != len)) {
    PyErr_SetString(PyExc_IndexError, "Not as many row indices as values");
    goto done;
  }
  const double *bv = (const double *)PyArray_DATA(b);
  const long *p1 = i1 ? (const long *)PyArray_DATA(i1) : NULL;
  const long *p2 = i2 ? (const long *)PyArray_DATA(i2) : NULL;
  for (long k = 0; k < len; k++) {
    long r = p1 ? p1[k] : k;
    long c = p2 ? p2[k] : r;
    double v = b_scalar ? bv[0] : bv[k];
    if (r < 0) r += self->dim[0];
    if (c < 0) c += self->dim[1];
    if (r < 0 || r >= self->dim[0] || c < 0 || c >= self->dim[1]) {   /* checked as long: no 2^31 wrap-around */
      PyErr_SetString(PyExc_IndexError, "indices out of range");
      goto done;
    }
    int rc = add ? SpMatrix_LLMatUpdateItemAdd(self, (int)r, (int)c, v) : SpMatrix_LLMatSetItem(self, (int)r, (int)c, v);
    if (rc) goto done;
  }
  Py_INCREF(Py_None);
  ret = Py_None;
done:
  Py_XDECREF(b); Py_XDECREF(i1); Py_XDECREF(i2);
  return ret;
}
static PyObject *LLMat_put(LLMatObject *self, PyObject *args) { return put_like(self, args, 0); }
static PyObject *LLMat_update_add_at(LLMatObject *self, PyObject *args) { return put_like(self, args, 1); }

/* val, irow, jcol = a.find()   (LLMat_Find, ll_mat.c:2999-3035) */
static PyObject *LLMat_find(LLMatObject *self, PyObject *noargs)
{
  npy_intp n = self->nnz;
  PyObject *ar = PyArray_SimpleNew(1, &n, NPY_INT), *ac = PyArray_SimpleNew(1, &n, NPY_INT), *av = PyArray_SimpleNew(1, &n, NPY_DOUBLE);
  if (!ar || !ac || !av) { Py_XDECREF(ar); Py_XDECREF(ac); Py_XDECREF(av); return NULL; }
  int *pi = (int *)PyArray_DATA((PyArrayObject *)ar), *pj = (int *)PyArray_DATA((PyArrayObject *)ac);
  double *pv = (double *)PyArray_DATA((PyArrayObject *)av);
  long e = 0;
  for (int i = 0; i < self->dim[0]; i++) {
    const LLRow *r = &self->rows[i];
    for (int k = 0; k < r->len; k++, e++) { pi[e] = i; pj[e] = r->col[k]; pv[e] = r->val[k]; }
  }
  return Py_BuildValue("NNN", av, ar, ac);
}

/* a.take(b, id1=None, id2=None): b[k] = a[id1[k], id2[k]]   (ll_mat.c LLMat_take) */
static PyObject *LLMat_take(LLMatObject *self, PyObject *args)
{
  PyObject *bIn, *id1in = NULL, *id2in = NULL;
  if (!PyArg_ParseTuple(args, "O!|OO", &PyArray_Type, &bIn, &id1in, &id2in)) return NULL;
  PyArrayObject *b = (PyArrayObject *)bIn, *i1 = NULL, *i2 = NULL;
  if (PyArray_TYPE(b) != NPY_DOUBLE || PyArray_NDIM(b) != 1 || !PyArray_IS_C_CONTIGUOUS(b)) {
    PyErr_SetString(PyExc_ValueError, "b must be a contiguous 1-dimensional double array");
    return NULL;
  }
  PyObject *ret = NULL;
  if (id1in && id1in != Py_None && as_index_array(id1in, &i1)) goto done;
  if (id2in && id2in != Py_None && as_index_array(id2in, &i2)) goto done;
  long len = (long)PyArray_SIZE(b);
  if ((i1 && PyArray_SIZE(i1) != len) || (i2 && PyArray_SIZE(i2) != len)) { PyErr_SetString(PyExc_IndexError, "Not as many indices as values"); goto done; }
  double *bv = (double *)PyArray_DATA(b);
  const long *p1 = i1 ? (const long *)PyArray_DATA(i1) : NULL, *p2 = i2 ? (const long *)PyArray_DATA(i2) : NULL;
  for (long k = 0; k < len; k++) {
    long r = p1 ? p1[k] : k, c = p2 ? p2[k] : r;
    if (r < 0) r += self->dim[0];
    if (c < 0) c += self->dim[1];
    if (r < 0 || r >= self->dim[0] || c < 0 || c >= self->dim[1]) {
      PyErr_SetString(PyExc_IndexError, "indices out of range");
      goto done;
    }
    bv[k] = SpMatrix_LLMatGetItem(self, (int)r, (int)c);
    if (PyErr_Occurred()) goto done;
  }
  Py_INCREF(Py_None);
  ret = Py_None;
done:
  Py_XDECREF(i1); Py_XDECREF(i2);
  return ret;
}

static PyObject *LLMat_keys_values_items(LLMatObject *self, int what)
{
  PyObject *list = PyList_New(self->nnz);
  if (!list) return NULL;
  Py_ssize_t e = 0;
  for (int i = 0; i < self->dim[0]; i++) {
    const LLRow *r = &self->rows[i];
    for (int k = 0; k < r->len; k++, e++) {
      PyObject *o;
      if (what == 0) o = Py_BuildValue("(ii)", i, r->col[k]);
      else if (what == 1) o = PyFloat_FromDouble(r->val[k]);
      else o = Py_BuildValue("((ii)d)", i, r->col[k], r->val[k]);
      if (!o) { Py_DECREF(list); return NULL; }
      PyList_SET_ITEM(list, e, o);
    }
  }
  return list;
}
static PyObject *LLMat_keys(LLMatObject *self, PyObject *n) { return LLMat_keys_values_items(self, 0); }
static PyObject *LLMat_values(LLMatObject *self, PyObject *n) { return LLMat_keys_values_items(self, 1); }
static PyObject *LLMat_items(LLMatObject *self, PyObject *n) { return LLMat_keys_values_items(self, 2); }

static PyObject *LLMat_copy(LLMatObject *self, PyObject *noargs)
{
  LLMatObject *c = (LLMatObject *)SpMatrix_NewLLMatObject(self->dim, self->issym, self->nnz, self->storeZeros);
  if (!c) return NULL;
  for (int i = 0; i < self->dim[0]; i++) {
    const LLRow *s = &self->rows[i];
    LLRow *d = &c->rows[i];
    if (!s->len) continue;
    d->col = (int *)malloc(sizeof(int) * (size_t)s->len);
    d->val = (double *)malloc(sizeof(double) * (size_t)s->len);
    if (!d->col || !d->val) { Py_DECREF(c); return PyErr_NoMemory(); }
    memcpy(d->col, s->col, sizeof(int) * (size_t)s->len);
    memcpy(d->val, s->val, sizeof(double) * (size_t)s->len);
    d->len = d->cap = s->len;
  }
  c->nnz = self->nnz;
  return (PyObject *)c;
}

static PyObject *LLMat_scale(LLMatObject *self, PyObject *args)
{
  double f;
  if (!PyArg_ParseTuple(args, "d", &f)) return NULL;
  for (int i = 0; i < self->dim[0]; i++) for (int k = 0; k < self->rows[i].len; k++) self->rows[i].val[k] *= f;
  self->version++;
  Py_RETURN_NONE;
}

/* a.shift(sigma, b): a += sigma * b   (LLMat_shift, ll_mat.c) */
static PyObject *LLMat_shift(LLMatObject *self, PyObject *args)
{
  double sigma;
  LLMatObject *b;
  if (!PyArg_ParseTuple(args, "dO!", &sigma, &LLMatType, &b)) return NULL;
  if (self->dim[0] != b->dim[0] || self->dim[1] != b->dim[1]) { PyErr_SetString(PyExc_ValueError, "matrix shapes do not match"); return NULL; }
  if (self->issym && !b->issym) { PyErr_SetString(PyExc_ValueError, "cannot shift symmetric matrix by non-symmetric matrix"); return NULL; }
  for (int i = 0; i < b->dim[0]; i++) {
    const LLRow *r = &b->rows[i];
    for (int k = 0; k < r->len; k++) {
      int j = r->col[k];
      if (SpMatrix_LLMatUpdateItemAdd(self, i, j, sigma * r->val[k])) return NULL;
      if (b->issym && !self->issym && i != j && SpMatrix_LLMatUpdateItemAdd(self, j, i, sigma * r->val[k])) return NULL;
    }
  }
  Py_RETURN_NONE;
}

/* a.norm(p): p in '1', 'inf', 'fro'   (LLMat_norm, ll_mat.c) */
static PyObject *LLMat_norm(LLMatObject *self, PyObject *args)
{
  const char *p;
  if (!PyArg_ParseTuple(args, "s", &p)) return NULL;
  const int m = self->dim[0], n = self->dim[1];
  double res = 0.0;
  if (!strcmp(p, "fro")) {
    for (int i = 0; i < m; i++) {
      const LLRow *r = &self->rows[i];
      for (int k = 0; k < r->len; k++) {
        double v = r->val[k] * r->val[k];
        res += (self->issym && r->col[k] != i) ? 2.0 * v : v;
      }
    }
    return PyFloat_FromDouble(sqrt(res));
  }
  int by_col = !strcmp(p, "1");
  if (!by_col && strcmp(p, "inf")) { PyErr_SetString(PyExc_ValueError, "unknown norm"); return NULL; }
  if (self->issym) {                              /* as the reference (test/test_spmatrix.py:193-199) */
    PyErr_SetString(PyExc_NotImplementedError, "Not implemented for symmetric matrices");
    return NULL;
  }
  int len = by_col ? n : m;
  if (self->issym) len = m;
  double *s = (double *)calloc((size_t)(len > 0 ? len : 1), sizeof(double));
  if (!s) return PyErr_NoMemory();
  for (int i = 0; i < m; i++) {
    const LLRow *r = &self->rows[i];
    for (int k = 0; k < r->len; k++) {
      double v = fabs(r->val[k]);
      int j = r->col[k];
      if (self->issym) { s[i] += v; if (i != j) s[j] += v; }
      else s[by_col ? j : i] += v;
    }
  }
  for (int i = 0; i < len; i++) if (s[i] > res) res = s[i];
  free(s);
  return PyFloat_FromDouble(res);
}

static PyObject *LLMat_generalize(LLMatObject *self, PyObject *noargs)
{
  if (self->issym) {
    self->issym = 0;
    for (int i = 0; i < self->dim[0]; i++) {
      int nl = row_find(&self->rows[i], i);
      for (int k = 0; k < nl; k++)
        if (SpMatrix_LLMatSetItem(self, self->rows[i].col[k], i, self->rows[i].val[k])) return NULL;
    }
  }
  Py_RETURN_NONE;
}

static PyObject *LLMat_compress(LLMatObject *self, PyObject *noargs)
{
  long freed = 0;
  for (int i = 0; i < self->dim[0]; i++) {
    LLRow *r = &self->rows[i];
    if (r->cap > r->len) {
      freed += r->cap - r->len;
      if (r->len == 0) { free(r->col); free(r->val); r->col = NULL; r->val = NULL; r->cap = 0; }
      else {
        r->col = (int *)realloc(r->col, sizeof(int) * (size_t)r->len);
        r->val = (double *)realloc(r->val, sizeof(double) * (size_t)r->len);
        r->cap = r->len;
      }
    }
  }
  return PyLong_FromLong(freed);
}

/* a.export_mtx(fileName, precision=16)   (MatrixMarket coordinate real) */
static PyObject *LLMat_export_mtx(LLMatObject *self, PyObject *args)
{
  const char *fname; int prec = 16;
  if (!PyArg_ParseTuple(args, "s|i", &fname, &prec)) return NULL;
  FILE *f = fopen(fname, "w");
  if (!f) return PyErr_SetFromErrnoWithFilename(PyExc_IOError, fname);
  fprintf(f, "%%%%MatrixMarket matrix coordinate real %s\n", self->issym ? "symmetric" : "general");
  fprintf(f, "%d %d %d\n", self->dim[0], self->dim[1], self->nnz);
  for (int i = 0; i < self->dim[0]; i++)
    for (int k = 0; k < self->rows[i].len; k++)
      fprintf(f, "%d %d %.*e\n", i + 1, self->rows[i].col[k] + 1, prec - 1, self->rows[i].val[k]);
  fclose(f);
  Py_RETURN_NONE;
}

static PyObject *LLMat_get_dev(LLMatObject *self, void *c)
{
  psp_mat *d = ll_device(self);
  return d ? PyCapsule_New(d, PSP_MAT_CAPSULE, mat_capsule_noop) : NULL;
}
static PyObject *LLMat_get_shape(LLMatObject *self, void *c) { return Py_BuildValue("(i,i)", self->dim[0], self->dim[1]); }
static PyObject *LLMat_get_nnz(LLMatObject *self, void *c) { return PyLong_FromLong(self->nnz); }
static PyObject *LLMat_get_issym(LLMatObject *self, void *c) { return PyLong_FromLong(self->issym); }

/* ---- subscripting: integers, slices, lists / tuples / integer arrays ----------
 * (getSubMatrix_FromList / setSubMatrix_FromList, ll_mat.c:632-1250; documented in
 * doc/pysparse/source/spmatrix.rst:94-260).  All paths walk the stored entries of the
 * selected rows -- never the dense l0 x l1 index space -- except for the assignment of
 * a non-zero scalar, which by definition fills a dense block.
 *
 * Deliberate differences from the reference snapshot (its fancy-indexing rewrite left its
 * own test/test_spmatrix.py:84-123 failing):
 *   - a stepped column slice selects (col - start) % step == 0 (the reference tests
 *     col % step, which is only right when start is a multiple of step);
 *   - A[rows, cols] = M REPLACES the block (entries of A in the block that are zero in M
 *     are removed), which is what the reference's tests and documentation expect; the
 *     snapshot only overlays the stored entries of M;
 *   - a symmetric M stands for the full symmetric matrix wherever it is assigned, and
 *     slicing a symmetric matrix returns the full (general) sub-matrix.
 */
typedef struct {
  int kind;                      /* 0 single integer, 1 slice, 2 explicit index list */
  Py_ssize_t start, step, len;   /* kinds 0 / 1 */
  long *list;                    /* kind 2: len indices, wrapped and bounds-checked (owned) */
} IndexSet;

static void index_set_free(IndexSet *s) { free(s->list); s->list = NULL; }

static int index_set_parse(PyObject *o, int dim, IndexSet *s)
{
  s->kind = 0; s->start = 0; s->step = 1; s->len = 0; s->list = NULL;
  if (PySlice_Check(o)) {
    Py_ssize_t stop;
    if (PySlice_Unpack(o, &s->start, &stop, &s->step) < 0) return -1;
    s->len = PySlice_AdjustIndices(dim, &s->start, &stop, s->step);
    s->kind = 1;
    return 0;
  }
  if (PyIndex_Check(o) && !PyArray_Check(o)) {
    Py_ssize_t v = PyNumber_AsSsize_t(o, PyExc_IndexError);
    if (v == -1 && PyErr_Occurred()) return -1;
    while (v < 0 && dim > 0) v += dim;            /* ll_mat.c:644-645 */
    if (v < 0 || v >= dim) { PyErr_SetString(PyExc_IndexError, "Index out of bounds"); return -1; }
    s->start = v; s->len = 1;
    return 0;
  }
  if (PyList_Check(o) || PyTuple_Check(o) || PyArray_Check(o)) {
    PyArrayObject *raw = (PyArrayObject *)PyArray_FROM_O(o), *a = NULL;
    if (raw && (PyArray_ISINTEGER(raw) || PyArray_SIZE(raw) == 0))     /* no silent truncation of floats */
      a = (PyArrayObject *)PyArray_FROM_OTF((PyObject *)raw, NPY_LONG, NPY_ARRAY_IN_ARRAY | NPY_ARRAY_FORCECAST);
    Py_XDECREF(raw);
    if (!a) { PyErr_Clear(); PyErr_SetString(PyExc_IndexError, "Invalid list item"); return -1; }
    if (PyArray_NDIM(a) != 1) { Py_DECREF(a); PyErr_SetString(PyExc_IndexError, "index arrays must be 1-dimensional"); return -1; }
    s->len = (Py_ssize_t)PyArray_SIZE(a);
    s->list = (long *)malloc(sizeof(long) * (size_t)(s->len > 0 ? s->len : 1));
    if (!s->list) { Py_DECREF(a); PyErr_NoMemory(); return -1; }
    const long *p = (const long *)PyArray_DATA(a);
    for (Py_ssize_t k = 0; k < s->len; k++) {
      long v = p[k];
      if (v < 0) v += dim;
      if (v < 0 || v >= dim) { Py_DECREF(a); index_set_free(s); PyErr_SetString(PyExc_IndexError, "Index out of bounds"); return -1; }
      s->list[k] = v;
    }
    Py_DECREF(a);
    s->kind = 2;
    return 0;
  }
  PyErr_SetString(PyExc_IndexError, "Index must be an integer, a slice, a list or an integer array");
  return -1;
}

static inline long index_at(const IndexSet *s, Py_ssize_t k) { return s->kind == 2 ? s->list[k] : (long)(s->start + k * s->step); }

/* column -> positions in an index set: arithmetic for slices, chains for lists (a column
 * may occur several times in a list) */
typedef struct { const IndexSet *s; int *head, *next; } ColMap;

static void colmap_free(ColMap *m) { free(m->head); free(m->next); m->head = m->next = NULL; }

static int colmap_build(ColMap *m, const IndexSet *s, int dim)
{
  m->s = s; m->head = m->next = NULL;
  if (s->kind != 2) return 0;
  m->head = (int *)malloc(sizeof(int) * (size_t)(dim > 0 ? dim : 1));
  m->next = (int *)malloc(sizeof(int) * (size_t)(s->len > 0 ? s->len : 1));
  if (!m->head || !m->next) { colmap_free(m); PyErr_NoMemory(); return -1; }
  for (int c = 0; c < dim; c++) m->head[c] = -1;
  for (Py_ssize_t k = s->len - 1; k >= 0; k--) { m->next[k] = m->head[s->list[k]]; m->head[s->list[k]] = (int)k; }
  return 0;
}

/* first position of column c in the set, or -1 */
static inline int colmap_first(const ColMap *m, int c)
{
  const IndexSet *s = m->s;
  if (s->kind == 2) return m->head[c];
  if (s->len == 0) return -1;
  Py_ssize_t d = c - s->start;
  if (s->step < 0) { d = -d; }
  const Py_ssize_t st = s->step < 0 ? -s->step : s->step;
  if (d < 0 || d % st != 0 || d / st >= s->len) return -1;
  return (int)(d / st);
}
static inline int colmap_next(const ColMap *m, int pos) { return m->s->kind == 2 ? m->next[pos] : -1; }

static PyObject *LLMat_copy(LLMatObject *self, PyObject *noargs);
static PyObject *LLMat_generalize(LLMatObject *self, PyObject *noargs);

static PyObject *ll_get_submatrix(LLMatObject *self, const IndexSet *r, const IndexSet *c)
{
  LLMatObject *src = self, *tmp = NULL, *dst = NULL;
  ColMap cm = {0};
  if (self->issym) {                              /* the full matrix, not just the stored triangle */
    tmp = (LLMatObject *)LLMat_copy(self, NULL);
    if (!tmp) return NULL;
    PyObject *none = LLMat_generalize(tmp, NULL);
    if (!none) { Py_DECREF(tmp); return NULL; }
    Py_DECREF(none);
    src = tmp;
  }
  int dim[2] = {(int)r->len, (int)c->len};
  dst = (LLMatObject *)SpMatrix_NewLLMatObject(dim, 0, 16, self->storeZeros);
  if (!dst || colmap_build(&cm, c, src->dim[1])) goto fail;
  for (Py_ssize_t a = 0; a < r->len; a++) {
    const LLRow *row = &src->rows[index_at(r, a)];
    for (int k = 0; k < row->len; k++)
      for (int j = colmap_first(&cm, row->col[k]); j >= 0; j = colmap_next(&cm, j))
        if (SpMatrix_LLMatSetItem(dst, (int)a, j, row->val[k])) goto fail;
  }
  colmap_free(&cm);
  Py_XDECREF(tmp);
  return (PyObject *)dst;
fail:
  colmap_free(&cm);
  Py_XDECREF(tmp);
  Py_XDECREF(dst);
  return NULL;
}

/* remove the stored entries of the block rows x cols */
static int ll_clear_block(LLMatObject *self, const IndexSet *r, const IndexSet *c)
{
  ColMap cm;
  if (colmap_build(&cm, c, self->dim[1])) return -1;
  for (Py_ssize_t a = 0; a < r->len; a++) {
    LLRow *row = &self->rows[index_at(r, a)];
    int w = 0;
    for (int k = 0; k < row->len; k++) {
      if (colmap_first(&cm, row->col[k]) >= 0) continue;
      row->col[w] = row->col[k]; row->val[w] = row->val[k]; w++;
    }
    self->nnz -= row->len - w;
    row->len = w;
  }
  colmap_free(&cm);
  self->version++;
  return 0;
}

static int upper_write_error(void)
{
  PyErr_SetString(PyExc_IndexError, "Writing to upper triangle of symmetric matrix");
  return -1;
}

static int ll_set_scalar(LLMatObject *self, const IndexSet *r, const IndexSet *c, double v)
{
  if (r->len == 0 || c->len == 0) return 0;
  if (self->issym) {                              /* nothing is written if any target is above the diagonal */
    long minrow = index_at(r, 0), maxcol = index_at(c, 0);
    for (Py_ssize_t a = 1; a < r->len; a++) if (index_at(r, a) < minrow) minrow = index_at(r, a);
    for (Py_ssize_t b = 1; b < c->len; b++) if (index_at(c, b) > maxcol) maxcol = index_at(c, b);
    if (minrow < maxcol) return upper_write_error();
  }
  if (v == 0.0 && !self->storeZeros) return ll_clear_block(self, r, c);
  for (Py_ssize_t a = 0; a < r->len; a++)
    for (Py_ssize_t b = 0; b < c->len; b++)
      if (SpMatrix_LLMatSetItem(self, (int)index_at(r, a), (int)index_at(c, b), v)) return -1;
  return 0;
}

static int ll_set_matrix(LLMatObject *self, const IndexSet *r, const IndexSet *c, LLMatObject *mat)
{
  if (mat->dim[0] != r->len || mat->dim[1] != c->len) { PyErr_SetString(PyExc_ValueError, "Matrix shapes are different"); return -1; }
  LLMatObject *own = NULL;
  int rc = -1;
  if (mat == self) {                              /* A[p, q] = A: read from a snapshot */
    own = (LLMatObject *)LLMat_copy(self, NULL);
    if (!own) return -1;
    mat = own;
  }
  if (self->issym && !mat->issym) {               /* a general block may not reach above the diagonal */
    for (int a = 0; a < mat->dim[0]; a++)
      for (int k = 0; k < mat->rows[a].len; k++)
        if (index_at(r, a) < index_at(c, mat->rows[a].col[k])) { upper_write_error(); goto done; }
  }
  if (ll_clear_block(self, r, c)) goto done;
  for (int a = 0; a < mat->dim[0]; a++) {
    const LLRow *row = &mat->rows[a];
    for (int k = 0; k < row->len; k++) {
      const int b = row->col[k];
      const double v = row->val[k];
      long i = index_at(r, a), j = index_at(c, b);
      if (!(self->issym && i < j) && SpMatrix_LLMatSetItem(self, (int)i, (int)j, v)) goto done;
      if (mat->issym && a != b) {                 /* the mirrored entry (b, a) of a symmetric source */
        i = index_at(r, b); j = index_at(c, a);
        if (!(self->issym && i < j) && SpMatrix_LLMatSetItem(self, (int)i, (int)j, v)) goto done;
      }
    }
  }
  rc = 0;
done:
  Py_XDECREF(own);
  return rc;
}

static int parse_index_pair(LLMatObject *self, PyObject *idx, IndexSet *r, IndexSet *c)
{
  if (!PyTuple_Check(idx) || PyTuple_GET_SIZE(idx) != 2) { PyErr_SetString(PyExc_IndexError, "There must be exactly two indices"); return -1; }
  if (index_set_parse(PyTuple_GET_ITEM(idx, 0), self->dim[0], r)) return -1;
  if (index_set_parse(PyTuple_GET_ITEM(idx, 1), self->dim[1], c)) { index_set_free(r); return -1; }
  return 0;
}

static PyObject *LLMat_subscript(LLMatObject *self, PyObject *idx)
{
  IndexSet r, c;
  if (parse_index_pair(self, idx, &r, &c)) return NULL;
  PyObject *ret;
  if (r.kind == 0 && c.kind == 0) ret = PyFloat_FromDouble(SpMatrix_LLMatGetItem(self, (int)r.start, (int)c.start));
  else ret = ll_get_submatrix(self, &r, &c);
  index_set_free(&r); index_set_free(&c);
  return ret;
}

static int LLMat_ass_subscript(LLMatObject *self, PyObject *idx, PyObject *value)
{
  if (!value) { PyErr_SetString(PyExc_TypeError, "cannot delete matrix entries"); return -1; }
  IndexSet r, c;
  if (parse_index_pair(self, idx, &r, &c)) return -1;
  int rc;
  if (PyObject_TypeCheck(value, &LLMatType)) {
    if (r.kind == 0 && c.kind == 0) { PyErr_SetString(PyExc_ValueError, "Value must be double"); rc = -1; }
    else rc = ll_set_matrix(self, &r, &c, (LLMatObject *)value);
  } else {
    double v = PyFloat_AsDouble(value);
    if (v == -1.0 && PyErr_Occurred()) { PyErr_Clear(); PyErr_SetString(PyExc_ValueError, "Value must be double"); rc = -1; }
    else if (r.kind == 0 && c.kind == 0) rc = SpMatrix_LLMatSetItem(self, (int)r.start, (int)c.start, v);
    else rc = ll_set_scalar(self, &r, &c, v);
  }
  index_set_free(&r); index_set_free(&c);
  return rc;
}

/* ---- masks, element-matrix assembly, scaling, row / column deletion ----------- */
static PyArrayObject *as_long_vector(PyObject *o, Py_ssize_t len, const char *what)
{
  PyArrayObject *a = (PyArrayObject *)PyArray_FROM_OTF(o, NPY_LONG, NPY_ARRAY_IN_ARRAY);
  if (!a || PyArray_NDIM(a) != 1 || (len >= 0 && PyArray_SIZE(a) != len)) {
    Py_XDECREF(a);
    PyErr_Clear();
    PyErr_Format(PyExc_ValueError, "%s must be a 1D integer NumPy array of appropriate length", what);
    return NULL;
  }
  return a;
}

/* A.update_add_mask(B, ind0, ind1, mask0, mask1): A[ind0[i], ind1[j]] += B[i, j] where
 * mask0[i] and mask1[j]   (LLMat_update_add_mask, ll_mat.c:2201-2291; spmatrix.rst:358-382).
 * B is read as documented, B[i, j]; the reference snapshot addresses its C-contiguous copy
 * of B column-major, i.e. reads B[j, i] of a square B -- the same for the symmetric element
 * matrices the method is meant for. */
static PyObject *LLMat_update_add_mask(LLMatObject *self, PyObject *args)
{
  PyObject *bIn, *i0In, *i1In, *m0In, *m1In, *ret = NULL;
  if (self->issym) { PyErr_SetString(SpMatrix_ErrorObject, "Method not allowed for symmetric matrices"); return NULL; }
  if (!PyArg_ParseTuple(args, "OOOOO", &bIn, &i0In, &i1In, &m0In, &m1In)) return NULL;
  PyArrayObject *b = (PyArrayObject *)PyArray_FROM_OTF(bIn, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  PyArrayObject *i0 = NULL, *i1 = NULL, *m0 = NULL, *m1 = NULL;
  if (!b) return NULL;
  if (PyArray_NDIM(b) != 2) { PyErr_SetString(PyExc_ValueError, "arg 1 must be a 2-dimensional double array"); goto done; }
  if (!(i0 = as_long_vector(i0In, -1, "arg 2")) || !(i1 = as_long_vector(i1In, -1, "arg 3"))) goto done;
  const Py_ssize_t l0 = PyArray_SIZE(i0), l1 = PyArray_SIZE(i1);
  m0 = (PyArrayObject *)PyArray_FROM_OTF(m0In, NPY_LONG, NPY_ARRAY_IN_ARRAY);
  m1 = (PyArrayObject *)PyArray_FROM_OTF(m1In, NPY_LONG, NPY_ARRAY_IN_ARRAY);
  if (!m0 || !m1) goto done;
  if (PyArray_NDIM(m0) != 1 || PyArray_NDIM(m1) != 1 || PyArray_SIZE(m0) != l0 || PyArray_SIZE(m1) != l1) {
    PyErr_SetString(PyExc_ValueError, "Shapes of index and mask array do not match");
    goto done;
  }
  if (PyArray_DIM(b, 0) != l0 || PyArray_DIM(b, 1) != l1) {
    PyErr_SetString(PyExc_ValueError, "Shapes of input matrix and index arrays do not match");
    goto done;
  }
  const double *bv = (const double *)PyArray_DATA(b);
  const long *p0 = (const long *)PyArray_DATA(i0), *p1 = (const long *)PyArray_DATA(i1);
  const long *q0 = (const long *)PyArray_DATA(m0), *q1 = (const long *)PyArray_DATA(m1);
  for (Py_ssize_t i = 0; i < l0; i++) {
    if (!q0[i]) continue;
    long r = p0[i];
    if (r < 0) r += self->dim[0];
    if (r < 0 || r >= self->dim[0]) { PyErr_SetString(PyExc_IndexError, "element of arg 2 out of range"); goto done; }
    for (Py_ssize_t j = 0; j < l1; j++) {
      if (!q1[j]) continue;
      long c = p1[j];
      if (c < 0) c += self->dim[1];
      if (c < 0 || c >= self->dim[1]) { PyErr_SetString(PyExc_IndexError, "element of arg 3 out of range"); goto done; }
      if (SpMatrix_LLMatUpdateItemAdd(self, (int)r, (int)c, bv[i * l1 + j])) goto done;
    }
  }
  Py_INCREF(Py_None);
  ret = Py_None;
done:
  Py_XDECREF(b); Py_XDECREF(i0); Py_XDECREF(i1); Py_XDECREF(m0); Py_XDECREF(m1);
  return ret;
}

/* A.update_add_mask_sym(B, ind, mask): the lower triangle of the (symmetric) element matrix
 * B is added where mask[i] and mask[j]: once, below the diagonal, for a symmetric A; to both
 * (ind[i], ind[j]) and (ind[j], ind[i]) for a general one   (ll_mat.c:2306-2391). */
static PyObject *LLMat_update_add_mask_sym(LLMatObject *self, PyObject *args)
{
  PyObject *bIn, *iIn, *mIn, *ret = NULL;
  if (!PyArg_ParseTuple(args, "OOO", &bIn, &iIn, &mIn)) return NULL;
  PyArrayObject *b = (PyArrayObject *)PyArray_FROM_OTF(bIn, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  PyArrayObject *ind = NULL, *mask = NULL;
  if (!b) return NULL;
  if (PyArray_NDIM(b) != 2) { PyErr_SetString(PyExc_ValueError, "arg 1 must be a 2-dimensional double array"); goto done; }
  if (!(ind = as_long_vector(iIn, -1, "arg 2"))) goto done;
  const Py_ssize_t len = PyArray_SIZE(ind);
  mask = (PyArrayObject *)PyArray_FROM_OTF(mIn, NPY_LONG, NPY_ARRAY_IN_ARRAY);
  if (!mask) goto done;
  if (PyArray_NDIM(mask) != 1 || PyArray_SIZE(mask) != len) { PyErr_SetString(PyExc_ValueError, "Shapes of index and mask array do not match"); goto done; }
  if (PyArray_DIM(b, 0) != len || PyArray_DIM(b, 1) != len) { PyErr_SetString(PyExc_ValueError, "Shapes of input matrix and index arrays do not match"); goto done; }
  const double *bv = (const double *)PyArray_DATA(b);
  const long *p = (const long *)PyArray_DATA(ind), *q = (const long *)PyArray_DATA(mask);
  for (Py_ssize_t i = 0; i < len; i++) {
    if (!q[i]) continue;
    long r = p[i];
    if (r < 0) r += self->dim[0];
    if (r < 0 || r >= self->dim[0]) { PyErr_SetString(PyExc_IndexError, "element of arg 2 out of range"); goto done; }
    for (Py_ssize_t j = 0; j <= i; j++) {
      if (!q[j]) continue;
      long c = p[j];
      if (c < 0) c += self->dim[1];
      if (c < 0 || c >= self->dim[1]) { PyErr_SetString(PyExc_IndexError, "element of arg 2 out of range"); goto done; }
      const double v = bv[i * len + j];
      if (self->issym) {
        if (SpMatrix_LLMatUpdateItemAdd(self, (int)(r > c ? r : c), (int)(r > c ? c : r), v)) goto done;
      } else {
        if (SpMatrix_LLMatUpdateItemAdd(self, (int)r, (int)c, v)) goto done;
        if (r != c && SpMatrix_LLMatUpdateItemAdd(self, (int)c, (int)r, v)) goto done;
      }
    }
  }
  Py_INCREF(Py_None);
  ret = Py_None;
done:
  Py_XDECREF(b); Py_XDECREF(ind); Py_XDECREF(mask);
  return ret;
}

/* A.col_scale(v): A[i, j] *= v[j];  A.row_scale(v): A[i, j] *= v[i] -- on the stored entries
 * (LLMat_col_scale / LLMat_row_scale, ll_mat.c:1473-1571) */
static PyObject *scale_by_vector(LLMatObject *self, PyObject *args, int by_col)
{
  PyObject *vIn;
  if (!PyArg_ParseTuple(args, "O", &vIn)) return NULL;
  PyArrayObject *v = (PyArrayObject *)PyArray_FROM_OTF(vIn, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!v) { PyErr_Clear(); PyErr_SetString(SpMatrix_ErrorObject, "Cannot read input vector."); return NULL; }
  if (PyArray_NDIM(v) != 1 || PyArray_SIZE(v) != self->dim[by_col ? 1 : 0]) {
    Py_DECREF(v);
    PyErr_SetString(SpMatrix_ErrorObject, "Size of scaling vector does not match the matrix.");
    return NULL;
  }
  const double *s = (const double *)PyArray_DATA(v);
  for (int i = 0; i < self->dim[0]; i++) {
    LLRow *row = &self->rows[i];
    for (int k = 0; k < row->len; k++) row->val[k] *= by_col ? s[row->col[k]] : s[i];
  }
  self->version++;
  Py_DECREF(v);
  Py_RETURN_NONE;
}
static PyObject *LLMat_col_scale(LLMatObject *self, PyObject *args) { return scale_by_vector(self, args, 1); }
static PyObject *LLMat_row_scale(LLMatObject *self, PyObject *args) { return scale_by_vector(self, args, 0); }

/* A.delete_rows(mask) / delete_cols(mask) / delete_rowcols(mask): row / column i is REMOVED
 * where mask[i] == 0, the matrix shrinks   (ll_mat.c:2764-2995, spmatrix.rst:468-481) */
static void ll_drop_rows(LLMatObject *self, const long *mask)
{
  int newm = 0;
  for (int i = 0; i < self->dim[0]; i++) {
    if (mask[i]) { self->rows[newm++] = self->rows[i]; continue; }
    self->nnz -= self->rows[i].len;
    free(self->rows[i].col); free(self->rows[i].val);
  }
  for (int i = newm; i < self->dim[0]; i++) memset(&self->rows[i], 0, sizeof(LLRow));
  self->dim[0] = newm;
}

static int ll_drop_cols(LLMatObject *self, const long *mask)
{
  const int n = self->dim[1];
  int *newcol = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (!newcol) { PyErr_NoMemory(); return -1; }
  int newn = 0;
  for (int j = 0; j < n; j++) newcol[j] = mask[j] ? newn++ : -1;
  for (int i = 0; i < self->dim[0]; i++) {
    LLRow *row = &self->rows[i];
    int w = 0;
    for (int k = 0; k < row->len; k++) {
      const int c = newcol[row->col[k]];
      if (c < 0) continue;
      row->col[w] = c; row->val[w] = row->val[k]; w++;
    }
    self->nnz -= row->len - w;
    row->len = w;
  }
  free(newcol);
  self->dim[1] = newn;
  return 0;
}

static PyObject *delete_by_mask(LLMatObject *self, PyObject *args, int rows, int cols)
{
  PyObject *mIn;
  if (!PyArg_ParseTuple(args, "O", &mIn)) return NULL;
  if (self->issym && !(rows && cols)) { PyErr_SetString(SpMatrix_ErrorObject, "method not allowed for symmetric matrices"); return NULL; }
  if (rows && cols && self->dim[0] != self->dim[1]) { PyErr_SetString(PyExc_ValueError, "delete_rowcols needs a square matrix"); return NULL; }
  PyArrayObject *mask = as_long_vector(mIn, self->dim[rows ? 0 : 1], "mask");
  if (!mask) return NULL;
  const long *m = (const long *)PyArray_DATA(mask);
  int rc = 0;
  /* columns first: a symmetric matrix keeps (i, j), i >= j, exactly when both survive, and
   * the surviving pairs stay below the diagonal because the renumbering is monotone */
  if (cols) rc = ll_drop_cols(self, m);
  if (!rc && rows) ll_drop_rows(self, m);
  self->version++;
  Py_DECREF(mask);
  if (rc) return NULL;
  Py_RETURN_NONE;
}
static PyObject *LLMat_delete_rows(LLMatObject *self, PyObject *args) { return delete_by_mask(self, args, 1, 0); }
static PyObject *LLMat_delete_cols(LLMatObject *self, PyObject *args) { return delete_by_mask(self, args, 0, 1); }
static PyObject *LLMat_delete_rowcols(LLMatObject *self, PyObject *args) { return delete_by_mask(self, args, 1, 1); }

/* len(A) = number of matrix positions (LLMat_length, ll_mat.c:3195) */
static Py_ssize_t LLMat_length(LLMatObject *self) { return (Py_ssize_t)self->dim[0] * self->dim[1]; }
static PyMappingMethods LLMat_as_mapping = {(lenfunc)LLMat_length, (binaryfunc)LLMat_subscript, (objobjargproc)LLMat_ass_subscript};

/* str(A) / print(A): the text LLMatType_print (ll_mat.c:3085-3157) writes -- a dense picture
 * for matrices of at most 500 x 20, the list of stored entries otherwise */
typedef struct { char *p; size_t len, cap; } StrBuf;
static int sb_printf(StrBuf *b, const char *fmt, ...)
{
  va_list ap;
  for (;;) {
    va_start(ap, fmt);
    int n = vsnprintf(b->p + b->len, b->cap - b->len, fmt, ap);
    va_end(ap);
    if (n < 0) return -1;
    if ((size_t)n < b->cap - b->len) { b->len += (size_t)n; return 0; }
    size_t ncap = (b->cap + (size_t)n + 64) * 2;
    char *np_ = (char *)realloc(b->p, ncap);
    if (!np_) return -1;
    b->p = np_; b->cap = ncap;
  }
}

static PyObject *LLMat_str(LLMatObject *a)
{
  StrBuf b = {(char *)malloc(256), 0, 256};
  if (!b.p) return PyErr_NoMemory();
  const char *sym = a->issym ? "symmetric" : "general";
  int rc = 0;
  if (a->dim[1] <= 20 && a->dim[0] <= 500) {
    rc |= sb_printf(&b, "ll_mat(%s, [%d,%d]):\n", sym, a->dim[0], a->dim[1]);
    for (int i = 0; i < a->dim[0] && !rc; i++) {
      const LLRow *r = &a->rows[i];
      int k = 0;
      for (int j = 0; j < a->dim[1] && !rc; j++) {
        double val = 0.0;
        if (k < r->len && r->col[k] == j) val = r->val[k++];
        if (val != 0.0) {
          int ex = (int)log10(fabs(val));
          if (abs(ex) <= 4) rc |= sb_printf(&b, "%9.*f ", ex < 0 ? 6 : 6 - ex, val);
          else rc |= sb_printf(&b, "%9.1e ", val);
        } else if (!a->issym || i > j) rc |= sb_printf(&b, " -------- ");
      }
      rc |= sb_printf(&b, "\n");
    }
  } else if (a->nnz == 0) {
    rc |= sb_printf(&b, "ll_mat(%s, [%d,%d])", sym, a->dim[0], a->dim[1]);
  } else {
    int first = 1;
    rc |= sb_printf(&b, "ll_mat(%s, [%d,%d], [", sym, a->dim[0], a->dim[1]);
    for (int i = 0; i < a->dim[0] && !rc; i++)
      for (int k = 0; k < a->rows[i].len && !rc; k++) {
        rc |= sb_printf(&b, "%s(%d,%d): %g", first ? "" : ", ", i, a->rows[i].col[k], a->rows[i].val[k]);
        first = 0;
      }
    rc |= sb_printf(&b, "])");
  }
  PyObject *o = rc ? PyErr_NoMemory() : PyUnicode_FromStringAndSize(b.p, (Py_ssize_t)b.len);
  free(b.p);
  return o;
}

static PyMethodDef LLMat_methods[] = {
  {"matvec", (PyCFunction)LLMat_matvec, METH_VARARGS, "A.matvec(x, y)\n\ncompute the sparse matrix-vector product y := A * x."},
  {"matvec_transp", (PyCFunction)LLMat_matvec_transp, METH_VARARGS, "A.matvec_transp(x, y)\n\ncompute y := A^T * x."},
  {"to_csr", (PyCFunction)LLMat_to_csr, METH_NOARGS, "A.to_csr()\n\nreturn a new CSRMatObject constructed from data of A"},
  {"to_sss", (PyCFunction)LLMat_to_sss, METH_NOARGS, "a.to_sss()\n\nreturn a new SSSMatObject constructed from the lower triangle of a"},
  {"generalize", (PyCFunction)LLMat_generalize, METH_NOARGS, "convert ll_mat object from symmetric to non-symmetric form (in-place)."},
  {"compress", (PyCFunction)LLMat_compress, METH_NOARGS, "release unused storage; returns the number of freed slots"},
  {"export_mtx", (PyCFunction)LLMat_export_mtx, METH_VARARGS, "A.export_mtx(fileName, precision=16)"},
  {"copy", (PyCFunction)LLMat_copy, METH_NOARGS, "A.copy()"},
  {"norm", (PyCFunction)LLMat_norm, METH_VARARGS, "A.norm(p), p in '1', 'inf', 'fro'"},
  {"shift", (PyCFunction)LLMat_shift, METH_VARARGS, "A.shift(sigma, B): A += sigma * B"},
  {"scale", (PyCFunction)LLMat_scale, METH_VARARGS, "A.scale(sigma)"},
  {"keys", (PyCFunction)LLMat_keys, METH_NOARGS, "list of (i, j) of the stored entries"},
  {"values", (PyCFunction)LLMat_values, METH_NOARGS, "list of the stored values"},
  {"items", (PyCFunction)LLMat_items, METH_NOARGS, "list of ((i, j), value)"},
  {"put", (PyCFunction)LLMat_put, METH_VARARGS, "a.put(b, id1, id2): a[id1[i], id2[i]] = b[i]"},
  {"take", (PyCFunction)LLMat_take, METH_VARARGS, "a.take(b, id1, id2): b[i] = a[id1[i], id2[i]]"},
  {"find", (PyCFunction)LLMat_find, METH_NOARGS, "(val, irow, jcol) = a.find()"},
  {"update_add_at", (PyCFunction)LLMat_update_add_at, METH_VARARGS, "a.update_add_at(b, id1, id2): a[id1[i], id2[i]] += b[i]"},
  {"update_add_mask", (PyCFunction)LLMat_update_add_mask, METH_VARARGS, "A.update_add_mask(B, ind0, ind1, mask0, mask1): A[ind0[i], ind1[j]] += B[i, j] where mask0[i] and mask1[j]"},
  {"update_add_mask_sym", (PyCFunction)LLMat_update_add_mask_sym, METH_VARARGS, "A.update_add_mask_sym(B, ind, mask): symmetric element-matrix assembly"},
  {"col_scale", (PyCFunction)LLMat_col_scale, METH_VARARGS, "A.col_scale(v): scale the i-th column of A by v[i] in place"},
  {"row_scale", (PyCFunction)LLMat_row_scale, METH_VARARGS, "A.row_scale(v): scale the i-th row of A by v[i] in place"},
  {"delete_rows", (PyCFunction)LLMat_delete_rows, METH_VARARGS, "A.delete_rows(mask): remove (in place) the rows with mask[i] == 0"},
  {"delete_cols", (PyCFunction)LLMat_delete_cols, METH_VARARGS, "A.delete_cols(mask): remove (in place) the columns with mask[i] == 0"},
  {"delete_rowcols", (PyCFunction)LLMat_delete_rowcols, METH_VARARGS, "A.delete_rowcols(mask): remove (in place) rows and columns with mask[i] == 0"},
  {NULL, NULL}
};
static PyObject *LLMat_get_storeZeros(LLMatObject *self, void *c) { return PyLong_FromLong(self->storeZeros); }
static PyGetSetDef LLMat_getset[] = {
  {"shape", (getter)LLMat_get_shape, NULL, NULL, NULL},
  {"nnz", (getter)LLMat_get_nnz, NULL, NULL, NULL},
  {"issym", (getter)LLMat_get_issym, NULL, NULL, NULL},
  {"storeZeros", (getter)LLMat_get_storeZeros, NULL, NULL, NULL},
  {"_psp_mat", (getter)LLMat_get_dev, NULL, "device handle (PyCapsule)", NULL},
  {NULL}
};

/* ========================================================================= */
/* C-API functions (spmatrixmodule.c:42-307)                                   */
/* ========================================================================= */
static int SpMatrix_GetShape(PyObject *op, int dim[])
{
  PyObject *sh = PyObject_GetAttrString(op, "shape");
  if (!sh) return -1;
  if (PySequence_Size(sh) != 2) { Py_DECREF(sh); PyErr_SetString(PyExc_ValueError, "invalid matrix shape"); return -1; }
  for (int k = 0; k < 2; k++) {
    PyObject *e = PySequence_GetItem(sh, k);
    dim[k] = e ? (int)PyLong_AsLong(e) : -1;
    Py_XDECREF(e);
  }
  Py_DECREF(sh);
  if (PyErr_Occurred()) { PyErr_Clear(); PyErr_SetString(PyExc_ValueError, "invalid matrix shape"); return -1; }
  return 0;
}

static int SpMatrix_GetOrder(PyObject *op, int *n)
{
  int dim[2];
  if (SpMatrix_GetShape(op, dim) == -1) return -1;
  if (dim[0] != dim[1]) { PyErr_SetString(PyExc_ValueError, "matrix is not square"); return -1; }
  *n = dim[0];
  return 0;
}

static double SpMatrix_GetItem(PyObject *op, int i, int j)
{
  if (Py_TYPE(op) == &SSSMatType) return sss_get((SSSMatObject *)op, i, j);
  if (Py_TYPE(op) == &CSRMatType) return csr_get((CSRMatObject *)op, i, j);
  if (Py_TYPE(op) == &LLMatType) return SpMatrix_LLMatGetItem((LLMatObject *)op, i, j);
  PyObject *index = Py_BuildValue("(ii)", i, j);
  PyObject *fo = index ? PyObject_GetItem(op, index) : NULL;
  Py_XDECREF(index);
  if (!fo) return 0.0;
  double d = PyFloat_AsDouble(fo);
  Py_DECREF(fo);
  return d;
}

static psp_mat *native_handle(PyObject *m)
{
  if (Py_TYPE(m) == &CSRMatType) return csr_device((CSRMatObject *)m);
  if (Py_TYPE(m) == &SSSMatType) return sss_device((SSSMatObject *)m);
  if (Py_TYPE(m) == &LLMatType) return ll_device((LLMatObject *)m);
  return NULL;
}

/* y = matrix * x on raw host buffers (slot 7).  Native matrices go to the GPU;
 * any other object is called back through its Python `matvec` method exactly
 * like the reference does (spmatrixmodule.c:169-201). */
static int SpMatrix_Matvec(PyObject *matrix, int nx, double *x, int ny, double *y)
{
  if (Py_TYPE(matrix) == &CSRMatType || Py_TYPE(matrix) == &SSSMatType || Py_TYPE(matrix) == &LLMatType) {
    psp_mat *d = native_handle(matrix);
    if (!d) return -1;
    int st = psp_matvec(d, 0, x, 1, y, 1);
    if (st != PSP_OK) { psp_raise(st); return -1; }
    return 0;
  }
  npy_intp dx = nx, dy = ny;
  PyObject *xa = PyArray_SimpleNewFromData(1, &dx, NPY_DOUBLE, x);
  PyObject *ya = xa ? PyArray_SimpleNewFromData(1, &dy, NPY_DOUBLE, y) : NULL;
  PyObject *res = ya ? PyObject_CallMethod(matrix, "matvec", "OO", xa, ya) : NULL;
  Py_XDECREF(xa); Py_XDECREF(ya);
  if (!res) return -1;
  Py_DECREF(res);
  return 0;
}

/* y = prec(x) (slot 8): always the Python `precon` method (spmatrixmodule.c:215-248);
 * the native preconditioner objects implement it on the GPU. */
static int SpMatrix_Precon(PyObject *prec, int n, double *x, double *y)
{
  npy_intp d = n;
  PyObject *xa = PyArray_SimpleNewFromData(1, &d, NPY_DOUBLE, x);
  PyObject *ya = xa ? PyArray_SimpleNewFromData(1, &d, NPY_DOUBLE, y) : NULL;
  PyObject *res = ya ? PyObject_CallMethod(prec, "precon", "OO", xa, ya) : NULL;
  Py_XDECREF(xa); Py_XDECREF(ya);
  if (!res) return -1;
  Py_DECREF(res);
  return 0;
}

static int SpMatrix_ParseVecOpArgs(PyObject *args, double **x_data, double **y_data, int n)
{
  PyArrayObject *x, *y;
  if (parse_vec_pair(args, &x, &y, n, n, 1)) return -1;
  *x_data = (double *)PyArray_DATA(x);
  *y_data = (double *)PyArray_DATA(y);
  return 0;
}

/* slot 15 (spmatrixmodule.c:266-307): call linsolver(A, b, x, tol, itmax[, K]) */
static int ItSolvers_Solve(PyObject *linsolver, PyObject *A, int n, double *b, double *x, double tol, int itmax,
                           PyObject *K, int *info, int *iter, double *relres)
{
  npy_intp d = n;
  PyObject *ba = PyArray_SimpleNewFromData(1, &d, NPY_DOUBLE, b);
  PyObject *xa = ba ? PyArray_SimpleNewFromData(1, &d, NPY_DOUBLE, x) : NULL;
  PyObject *res = NULL;
  if (xa) res = K ? PyObject_CallFunction(linsolver, "OOOdiO", A, ba, xa, tol, itmax, K)
                  : PyObject_CallFunction(linsolver, "OOOdi", A, ba, xa, tol, itmax);
  Py_XDECREF(ba); Py_XDECREF(xa);
  if (!res) return -1;
  int ok = PyArg_ParseTuple(res, "iid", info, iter, relres);
  Py_DECREF(res);
  return ok ? 0 : -1;
}

/* ========================================================================= */
/* module functions                                                            */
/* ========================================================================= */
static PyObject *LLMat_zeros(PyObject *self, PyObject *args)
{
  int dim[2], sizeHint = 1000, storeZeros = 0;
  if (!PyArg_ParseTuple(args, "ii|ii", dim, dim + 1, &sizeHint, &storeZeros)) return NULL;
  return SpMatrix_NewLLMatObject(dim, 0, sizeHint, storeZeros);
}
static PyObject *LLMat_sym_zeros(PyObject *self, PyObject *args)
{
  int dim[2], n, sizeHint = 1000, storeZeros = 0;
  if (!PyArg_ParseTuple(args, "i|ii", &n, &sizeHint, &storeZeros)) return NULL;
  dim[0] = dim[1] = n;
  return SpMatrix_NewLLMatObject(dim, 1, sizeHint, storeZeros);
}

/* ll_mat_from_mtx(fileName)   (ll_mat.c:3390-3456): MatrixMarket coordinate real */
static PyObject *LLMat_from_mtx(PyObject *self, PyObject *args)
{
  const char *fname;
  if (!PyArg_ParseTuple(args, "s", &fname)) return NULL;
  FILE *f = fopen(fname, "r");
  if (!f) return PyErr_SetFromErrnoWithFilename(PyExc_IOError, fname);
  char line[1024], obj[64], fmt[64], field[64], symm[64];
  if (!fgets(line, sizeof line, f) || sscanf(line, "%%%%MatrixMarket %63s %63s %63s %63s", obj, fmt, field, symm) != 4 ||
      strcmp(obj, "matrix") || strcmp(fmt, "coordinate")) {
    fclose(f);
    PyErr_SetString(PyExc_IOError, "error reading MatrixMarket header");
    return NULL;
  }
  if (strcmp(field, "real") && strcmp(field, "integer")) { fclose(f); PyErr_SetString(PyExc_TypeError, "Matrix type not supported"); return NULL; }
  int sym = !strcmp(symm, "symmetric");
  do { if (!fgets(line, sizeof line, f)) { fclose(f); PyErr_SetString(PyExc_IOError, "error reading MatrixMarket file"); return NULL; } } while (line[0] == '%');
  int dim[2]; long nz;
  if (sscanf(line, "%d %d %ld", dim, dim + 1, &nz) != 3) { fclose(f); PyErr_SetString(PyExc_IOError, "error reading MatrixMarket size line"); return NULL; }
  LLMatObject *A = (LLMatObject *)SpMatrix_NewLLMatObject(dim, sym, (int)nz, 0);
  if (!A) { fclose(f); return NULL; }
  for (long k = 0; k < nz; k++) {
    int i, j; double v;
    if (fscanf(f, "%d %d %lg", &i, &j, &v) != 3) { fclose(f); Py_DECREF(A); PyErr_SetString(PyExc_IOError, "error reading MatrixMarket data"); return NULL; }
    if (SpMatrix_LLMatSetItem(A, i - 1, j - 1, v)) { fclose(f); Py_DECREF(A); return NULL; }
  }
  fclose(f);
  return (PyObject *)A;
}

/* ---- pysparse_b200 additions: build csr_mat / sss_mat straight from arrays -- */
/* A private, C-contiguous copy of `o` as float64 / int32.  Integer arrays of any width are
 * accepted (NumPy's default int64 included) after an explicit range check; the copy means later
 * edits of the caller's arrays cannot desynchronise the host data from the cached device copy. */
static PyArrayObject *want(PyObject *o, int type)
{
  if (type == NPY_DOUBLE) return (PyArrayObject *)PyArray_FROM_OTF(o, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY | NPY_ARRAY_ENSURECOPY);
  if (PyArray_Check(o) && PyArray_TYPE((PyArrayObject *)o) == NPY_INT32)        /* already narrow: plain copy, */
    return (PyArrayObject *)PyArray_FROM_OTF(o, NPY_INT32, NPY_ARRAY_IN_ARRAY | NPY_ARRAY_ENSURECOPY);   /* range checked by structure_ok */
  PyArrayObject *wide = (PyArrayObject *)PyArray_FROM_OTF(o, NPY_INT64, NPY_ARRAY_IN_ARRAY);
  if (!wide) return NULL;
  const npy_int64 *w = (const npy_int64 *)PyArray_DATA(wide);
  const npy_intp len = PyArray_SIZE(wide);
  for (npy_intp k = 0; k < len; k++)
    if (w[k] < 0 || w[k] > 2147483647LL) {
      Py_DECREF(wide);
      PyErr_SetString(PyExc_ValueError, "index array holds values outside [0, 2^31)");
      return NULL;
    }
  PyArrayObject *out = (PyArrayObject *)PyArray_Cast(wide, NPY_INT32);
  Py_DECREF(wide);
  if (out && !PyArray_ISCARRAY(out)) { Py_DECREF(out); PyErr_SetString(PyExc_ValueError, "index array conversion failed"); return NULL; }
  return out;
}

/* the check the device upload would make, done at construction so that A[i,j], diagonal() ...
 * never read outside the arrays: monotone row pointers, column range, strict lower triangle */
static int structure_ok(int m, int n, const int *ind, const int *col, npy_intp nnz, int strict_lower, const char *who)
{
  if (m < 0 || n < 0 || ind[0] != 0 || ind[m] != nnz) { PyErr_Format(PyExc_ValueError, "%s: ind[0] != 0 or ind[m] != nnz", who); return 0; }
  for (int i = 0; i < m; i++) {
    if (ind[i + 1] < ind[i]) { PyErr_Format(PyExc_ValueError, "%s: row pointer array not monotone at row %d", who, i); return 0; }
    for (int k = ind[i]; k < ind[i + 1]; k++)
      if (col[k] < 0 || col[k] >= n || (strict_lower && col[k] >= i)) {
        PyErr_Format(PyExc_ValueError, "%s: column index %d out of range in row %d", who, col[k], i);
        return 0;
      }
  }
  return 1;
}

static PyObject *csr_from_arrays(PyObject *self, PyObject *args)
{
  int m, n; PyObject *oi, *oc, *ov;
  if (!PyArg_ParseTuple(args, "iiOOO", &m, &n, &oi, &oc, &ov)) return NULL;
  PyArrayObject *ind = want(oi, NPY_INT32), *col = ind ? want(oc, NPY_INT32) : NULL, *val = col ? want(ov, NPY_DOUBLE) : NULL;
  if (!val) { Py_XDECREF(ind); Py_XDECREF(col); return NULL; }
  if (PyArray_SIZE(ind) != (npy_intp)m + 1 || PyArray_SIZE(col) != PyArray_SIZE(val)) {
    Py_DECREF(ind); Py_DECREF(col); Py_DECREF(val);
    PyErr_SetString(PyExc_ValueError, "csr_from_arrays: inconsistent array lengths");
    return NULL;
  }
  if (!structure_ok(m, n, (const int *)PyArray_DATA(ind), (const int *)PyArray_DATA(col), PyArray_SIZE(val), 0, "csr_from_arrays")) {
    Py_DECREF(ind); Py_DECREF(col); Py_DECREF(val);
    return NULL;
  }
  CSRMatObject *op = PyObject_New(CSRMatObject, &CSRMatType);
  if (!op) { Py_DECREF(ind); Py_DECREF(col); Py_DECREF(val); return PyErr_NoMemory(); }
  op->dim[0] = m; op->dim[1] = n; op->nnz = (int)PyArray_SIZE(val); op->dev = NULL;
  op->val = (double *)PyArray_DATA(val); op->col = (int *)PyArray_DATA(col); op->ind = (int *)PyArray_DATA(ind);
  op->base[0] = (PyObject *)val; op->base[1] = (PyObject *)col; op->base[2] = (PyObject *)ind;
  return (PyObject *)op;
}

static PyObject *sss_from_arrays(PyObject *self, PyObject *args)
{
  int n; PyObject *oi, *oc, *ov, *od;
  if (!PyArg_ParseTuple(args, "iOOOO", &n, &oi, &oc, &ov, &od)) return NULL;
  PyArrayObject *ind = want(oi, NPY_INT32), *col = ind ? want(oc, NPY_INT32) : NULL, *val = col ? want(ov, NPY_DOUBLE) : NULL,
                *diag = val ? want(od, NPY_DOUBLE) : NULL;
  if (!diag) { Py_XDECREF(ind); Py_XDECREF(col); Py_XDECREF(val); return NULL; }
  if (PyArray_SIZE(ind) != (npy_intp)n + 1 || PyArray_SIZE(col) != PyArray_SIZE(val) || PyArray_SIZE(diag) != n) {
    Py_DECREF(ind); Py_DECREF(col); Py_DECREF(val); Py_DECREF(diag);
    PyErr_SetString(PyExc_ValueError, "sss_from_arrays: inconsistent array lengths");
    return NULL;
  }
  if (!structure_ok(n, n, (const int *)PyArray_DATA(ind), (const int *)PyArray_DATA(col), PyArray_SIZE(val), 1, "sss_from_arrays")) {
    Py_DECREF(ind); Py_DECREF(col); Py_DECREF(val); Py_DECREF(diag);
    return NULL;
  }
  SSSMatObject *op = PyObject_New(SSSMatObject, &SSSMatType);
  if (!op) { Py_DECREF(ind); Py_DECREF(col); Py_DECREF(val); Py_DECREF(diag); return PyErr_NoMemory(); }
  op->n = n; op->nnz = (int)PyArray_SIZE(val); op->dev = NULL;
  op->val = (double *)PyArray_DATA(val); op->diag = (double *)PyArray_DATA(diag);
  op->col = (int *)PyArray_DATA(col); op->ind = (int *)PyArray_DATA(ind);
  op->base[0] = (PyObject *)val; op->base[1] = (PyObject *)diag; op->base[2] = (PyObject *)col; op->base[3] = (PyObject *)ind;
  return (PyObject *)op;
}

/* ---- sparse x sparse products of the assembly container ----------------------
 * matrixmultiply(A, B) = A B, dot(A, B) = A^T B, symdot(A[, d]) = A^T diag(d) A
 * (LLMat_matrixmultiply / LLMat_dot / LLMat_symdot, ll_mat.c:3462-3796).  Like shift() and
 * scale() these are operations OF THE HOST CONTAINER (the reference's are too); their
 * result is an ll_mat, whose matvec is the GPU path.  Every entry of the result adds its
 * terms in the reference's order, so the values are bit-identical to the reference's. */
static PyObject *LLMat_matrixmultiply(PyObject *module, PyObject *args)
{
  LLMatObject *A, *B, *C = NULL, *Afull = NULL;
  double *acc = NULL;
  int *mark = NULL, *cols = NULL;
  if (!PyArg_ParseTuple(args, "O!O!", &LLMatType, &A, &LLMatType, &B)) return NULL;
  if (A->dim[1] != B->dim[0]) { PyErr_SetString(PyExc_ValueError, "matrix dimensions must agree"); return NULL; }
  if (B->issym) {
    PyErr_SetString(PyExc_NotImplementedError, A->issym ? "multiply of two symmetric matrices not supported"
                                                        : "multiply of an unsymmetric and a symmetric matrix not supported");
    return NULL;
  }
  int dim[2] = {A->dim[0], B->dim[1]};
  C = (LLMatObject *)SpMatrix_NewLLMatObject(dim, 0, 1000, A->storeZeros == 1 && B->storeZeros == 1);
  if (!C) return NULL;
  if (A->issym) {              /* row i of the full matrix, ascending columns: the order in which the
                                  reference's sym-unsym loop adds the terms of row i of the product */
    Afull = (LLMatObject *)LLMat_copy(A, NULL);
    PyObject *none = Afull ? LLMat_generalize(Afull, NULL) : NULL;
    if (!none) goto fail;
    Py_DECREF(none);
    A = Afull;
  }
  const int n = B->dim[1];
  acc = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  mark = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  cols = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (!acc || !mark || !cols) { PyErr_NoMemory(); goto fail; }
  for (int j = 0; j < n; j++) mark[j] = -1;
  for (int i = 0; i < A->dim[0]; i++) {
    const LLRow *ra = &A->rows[i];
    int cnt = 0;
    for (int k = 0; k < ra->len; k++) {
      const double va = ra->val[k];
      const LLRow *rb = &B->rows[ra->col[k]];
      for (int l = 0; l < rb->len; l++) {
        const int j = rb->col[l];
        if (mark[j] != i) { mark[j] = i; cols[cnt++] = j; acc[j] = va * rb->val[l]; }
        else acc[j] += va * rb->val[l];
      }
    }
    for (int t = 0; t < cnt; t++)
      if (SpMatrix_LLMatSetItem(C, i, cols[t], acc[cols[t]])) goto fail;
  }
  free(acc); free(mark); free(cols);
  Py_XDECREF(Afull);
  return (PyObject *)C;
fail:
  free(acc); free(mark); free(cols);
  Py_XDECREF(Afull);
  Py_XDECREF(C);
  return NULL;
}

static PyObject *LLMat_dot(PyObject *module, PyObject *args)
{
  LLMatObject *A, *B;
  if (!PyArg_ParseTuple(args, "O!O!", &LLMatType, &A, &LLMatType, &B)) return NULL;
  if (A->dim[0] != B->dim[0]) { PyErr_SetString(PyExc_ValueError, "matrix dimensions must agree"); return NULL; }
  if (A->issym || B->issym) { PyErr_SetString(PyExc_NotImplementedError, "ddot operation with symmetric matrices not supported"); return NULL; }
  int dim[2] = {A->dim[1], B->dim[1]};
  LLMatObject *C = (LLMatObject *)SpMatrix_NewLLMatObject(dim, 0, 1000, A->storeZeros == 1 && B->storeZeros == 1);
  if (!C) return NULL;
  for (int i = 0; i < A->dim[0]; i++) {
    const LLRow *ra = &A->rows[i], *rb = &B->rows[i];
    for (int k = 0; k < ra->len; k++)
      for (int l = 0; l < rb->len; l++)
        if (SpMatrix_LLMatUpdateItemAdd(C, ra->col[k], rb->col[l], ra->val[k] * rb->val[l])) { Py_DECREF(C); return NULL; }
  }
  return (PyObject *)C;
}

static PyObject *LLMat_symdot(PyObject *module, PyObject *args)
{
  LLMatObject *A;
  PyObject *dIn = NULL;
  PyArrayObject *d = NULL;
  if (!PyArg_ParseTuple(args, "O!|O", &LLMatType, &A, &dIn)) return NULL;
  if (dIn && dIn != Py_None) {
    d = (PyArrayObject *)PyArray_FROM_OTF(dIn, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
    if (!d || PyArray_NDIM(d) != 1) { Py_XDECREF(d); PyErr_Clear(); PyErr_SetString(SpMatrix_ErrorObject, "Could not read scaling vector."); return NULL; }
    if (PyArray_SIZE(d) != A->dim[0]) { Py_DECREF(d); PyErr_SetString(SpMatrix_ErrorObject, "Scaling vector has wrong dimension."); return NULL; }
  }
  if (A->issym) { Py_XDECREF(d); PyErr_SetString(PyExc_NotImplementedError, "symdot operation with symmetric matrices not supported"); return NULL; }
  int dim[2] = {A->dim[1], A->dim[1]};
  LLMatObject *C = (LLMatObject *)SpMatrix_NewLLMatObject(dim, 1, 1000, A->storeZeros == 1);
  if (!C) { Py_XDECREF(d); return NULL; }
  const double *dv = d ? (const double *)PyArray_DATA(d) : NULL;
  for (int i = 0; i < A->dim[0]; i++) {
    const LLRow *r = &A->rows[i];
    for (int k = 0; k < r->len; k++) {
      double va = r->val[k];
      if (dv) va *= dv[i];
      for (int l = 0; l < r->len && r->col[l] <= r->col[k]; l++)
        if (SpMatrix_LLMatUpdateItemAdd(C, r->col[k], r->col[l], va * r->val[l])) { Py_XDECREF(d); Py_DECREF(C); return NULL; }
    }
  }
  Py_XDECREF(d);
  return (PyObject *)C;
}

static PyMethodDef spmatrix_methods[] = {
  {"ll_mat", LLMat_zeros, METH_VARARGS, "ll_mat(n, m, sizeHint=1000, storeZeros=0)"},
  {"ll_mat_sym", LLMat_sym_zeros, METH_VARARGS, "ll_mat_sym(n, sizeHint=1000, storeZeros=0)"},
  {"ll_mat_from_mtx", LLMat_from_mtx, METH_VARARGS, "ll_mat_from_mtx(fileName)"},
  {"matrixmultiply", LLMat_matrixmultiply, METH_VARARGS, "matrixmultiply(A, B)\n\nReturns a new ll_mat object representing the matrix A*B"},
  {"dot", LLMat_dot, METH_VARARGS, "dot(A, B)\n\nReturns a new ll_mat object representing the matrix transpose(A)*B"},
  {"symdot", LLMat_symdot, METH_VARARGS, "symdot(A) or symdot(A, d)\n\nReturns a new symmetric ll_mat object representing transpose(A)*A, or transpose(A)*diag(d)*A"},
  {"csr_from_arrays", csr_from_arrays, METH_VARARGS, "csr_from_arrays(m, n, ind, col, val) -> csr_mat (pysparse_b200 addition)"},
  {"sss_from_arrays", sss_from_arrays, METH_VARARGS, "sss_from_arrays(n, ind, col, val, diag) -> sss_mat (pysparse_b200 addition)"},
  {NULL, NULL}
};

static void *SpMatrix_API[PSP_API_pointers];

static struct PyModuleDef spmatrix_def = {PyModuleDef_HEAD_INIT, "spmatrix",
  "pysparse_b200.sparse.spmatrix -- ll_mat / csr_mat / sss_mat with GPU matrix-vector products", -1, spmatrix_methods};

static int init_type(PyTypeObject *t, const char *name, size_t size, destructor dealloc, PyMethodDef *methods,
                     PyGetSetDef *getset, PyMappingMethods *mapping)
{
  t->tp_name = name;
  t->tp_basicsize = (Py_ssize_t)size;
  t->tp_dealloc = dealloc;
  t->tp_flags = Py_TPFLAGS_DEFAULT;
  t->tp_methods = methods;
  t->tp_getset = getset;
  t->tp_as_mapping = mapping;
  return PyType_Ready(t);
}

PyMODINIT_FUNC PyInit_spmatrix(void)
{
  import_array();
  static PyTypeObject zero = {PyVarObject_HEAD_INIT(NULL, 0)};
  LLMatType = zero; CSRMatType = zero; SSSMatType = zero;
  LLMatType.tp_str = (reprfunc)LLMat_str;
  if (init_type(&LLMatType, "pysparse_b200.sparse.spmatrix.ll_mat", sizeof(LLMatObject), (destructor)LLMat_dealloc, LLMat_methods, LLMat_getset, &LLMat_as_mapping) < 0 ||
      init_type(&CSRMatType, "pysparse_b200.sparse.spmatrix.csr_mat", sizeof(CSRMatObject), (destructor)CSRMat_dealloc, CSRMat_methods, CSRMat_getset, &CSRMat_as_mapping) < 0 ||
      init_type(&SSSMatType, "pysparse_b200.sparse.spmatrix.sss_mat", sizeof(SSSMatObject), (destructor)SSSMat_dealloc, SSSMat_methods, SSSMat_getset, &SSSMat_as_mapping) < 0)
    return NULL;
  PyObject *m = PyModule_Create(&spmatrix_def);
  if (!m) return NULL;
  Py_INCREF(&LLMatType); PyModule_AddObject(m, "LLMatType", (PyObject *)&LLMatType);
  Py_INCREF(&CSRMatType); PyModule_AddObject(m, "CSRMatType", (PyObject *)&CSRMatType);
  Py_INCREF(&SSSMatType); PyModule_AddObject(m, "SSSMatType", (PyObject *)&SSSMatType);
  /* the reference's "spmatrix.error" is a Python-2 string exception (spmatrixmodule.c:373);
   * here it is an exception class of that name */
  SpMatrix_ErrorObject = PyErr_NewException("spmatrix.error", NULL, NULL);
  if (!SpMatrix_ErrorObject) return NULL;
  Py_INCREF(SpMatrix_ErrorObject);
  PyModule_AddObject(m, "error", SpMatrix_ErrorObject);

  SpMatrix_API[PSP_API_LLMatType] = (void *)&LLMatType;
  SpMatrix_API[PSP_API_CSRMatType] = (void *)&CSRMatType;
  SpMatrix_API[PSP_API_SSSMatType] = (void *)&SSSMatType;
  SpMatrix_API[PSP_API_ParseVecOpArgs] = (void *)SpMatrix_ParseVecOpArgs;
  SpMatrix_API[PSP_API_GetShape] = (void *)SpMatrix_GetShape;
  SpMatrix_API[PSP_API_GetOrder] = (void *)SpMatrix_GetOrder;
  SpMatrix_API[PSP_API_GetItem] = (void *)SpMatrix_GetItem;
  SpMatrix_API[PSP_API_Matvec] = (void *)SpMatrix_Matvec;
  SpMatrix_API[PSP_API_Precon] = (void *)SpMatrix_Precon;
  SpMatrix_API[PSP_API_NewLLMatObject] = (void *)SpMatrix_NewLLMatObject;
  SpMatrix_API[PSP_API_LLMatGetItem] = (void *)SpMatrix_LLMatGetItem;
  SpMatrix_API[PSP_API_LLMatSetItem] = (void *)SpMatrix_LLMatSetItem;
  SpMatrix_API[PSP_API_LLMatUpdateItemAdd] = (void *)SpMatrix_LLMatUpdateItemAdd;
  SpMatrix_API[PSP_API_LLMatBuildColIndex] = (void *)SpMatrix_LLMatBuildColIndex;
  SpMatrix_API[PSP_API_LLMatDestroyColIndex] = (void *)SpMatrix_LLMatDestroyColIndex;
  SpMatrix_API[PSP_API_ItSolvers_Solve] = (void *)ItSolvers_Solve;
  PyObject *cap = PyCapsule_New((void *)SpMatrix_API, PSP_CAPI_NAME, NULL);
  if (!cap) return NULL;
  PyModule_AddObject(m, "_C_API", cap);
  return m;
}
