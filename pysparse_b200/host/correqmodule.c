/* correqmodule.c -- module `correq` of pysparse_b200.eigen: the Jacobi-Davidson
 * correction-equation operator as a device-resident object.
 *
 * In the reference the operator is a C object created by jdsym.c:570 (CorrEq_New,
 * pysparse/eigen/src/correq.c:28-92) and handed to the Python-level inner solver as both
 * `A` (matvec, :144-221) and `K` (precon, :223-262); its projections are dgemv calls on the
 * host arrays Q, Qm, Y and an LU factorisation of H.  Here
 *     op = correq.operator(optype, A, M, K, kmax)    A, M: matrices, K: preconditioner of this
 *                                                    package (M, K may be None)
 *     op.set_columns(which, col0, block)             upload columns of Q (0), M Q (1), Y (2)
 *     op.update(k, theta[, hinv])                    columns in use, shift, inv(H[:k, :k])
 *     op.shape, op.matvec(x, y), op.precon(x, y)     the duck-typed protocol of ItSolvers_Solve
 *     op._psp_mat, op._psp_prec                      capsules: krylov.* keeps the solve on the GPU
 * forwards to psp_correq_* (include/psp_b200.h).  A TypeError from the constructor means "one
 * of the operands is not a device object": jdsym then uses its host-side operator instead.
 */
#define PY_ARRAY_UNIQUE_SYMBOL psp_correq_ARRAY_API
#include "psp_host.h"

static void **SpMatrix_API;

typedef struct {
  PyObject_HEAD
  int n, kmax;
  PyObject *amat, *mmat, *prec;    /* strong references: the device objects are borrowed by dev */
  psp_correq *dev;
} CorrEqObject;

static void CorrEq_dealloc(CorrEqObject *self)
{
  if (self->dev) psp_correq_destroy(self->dev);
  Py_XDECREF(self->amat); Py_XDECREF(self->mmat); Py_XDECREF(self->prec);
  PyObject_Del(self);
}

/* pointer inside capsule attribute `attr` of obj; NULL with TypeError when obj has none */
static void *handle_of(PyObject *obj, const char *attr, const char *capsule, const char *what)
{
  PyObject *cap = PyObject_GetAttrString(obj, attr);
  if (!cap) {
    if (PyErr_ExceptionMatches(PyExc_AttributeError)) {
      PyErr_Clear();
      PyErr_Format(PyExc_TypeError, "%s is not a device-resident pysparse_b200 object", what);
    }
    return NULL;
  }
  void *h = PyCapsule_IsValid(cap, capsule) ? PyCapsule_GetPointer(cap, capsule) : NULL;
  Py_DECREF(cap);
  if (!h && !PyErr_Occurred()) PyErr_Format(PyExc_TypeError, "%s has no device handle", what);
  return h;
}

static PyTypeObject CorrEqType;

static PyObject *correq_new(PyObject *mod, PyObject *args)
{
  int optype, kmax;
  PyObject *amat, *mmat, *prec;
  if (!PyArg_ParseTuple(args, "iOOOi", &optype, &amat, &mmat, &prec, &kmax)) return NULL;
  psp_mat *ha = (psp_mat *)handle_of(amat, "_psp_mat", PSP_MAT_CAPSULE, "A");
  if (!ha) return NULL;
  psp_mat *hm = NULL;
  psp_prec *hk = NULL;
  if (mmat != Py_None && !(hm = (psp_mat *)handle_of(mmat, "_psp_mat", PSP_MAT_CAPSULE, "M"))) return NULL;
  if (prec != Py_None && !(hk = (psp_prec *)handle_of(prec, "_psp_prec", PSP_PREC_CAPSULE, "K"))) return NULL;
  psp_correq *dev = NULL;
  int st = psp_correq_create(optype, ha, hm, hk, kmax, &dev);
  if (st != PSP_OK) return psp_raise(st);
  CorrEqObject *op = PyObject_New(CorrEqObject, &CorrEqType);
  if (!op) { psp_correq_destroy(dev); return PyErr_NoMemory(); }
  int m = 0;
  psp_mat_info(ha, &m, NULL, NULL, NULL);
  op->n = m; op->kmax = kmax; op->dev = dev;
  Py_INCREF(amat); op->amat = amat;
  op->mmat = NULL; op->prec = NULL;
  if (mmat != Py_None) { Py_INCREF(mmat); op->mmat = mmat; }
  if (prec != Py_None) { Py_INCREF(prec); op->prec = prec; }
  return (PyObject *)op;
}

static PyObject *CorrEq_set_columns(CorrEqObject *self, PyObject *args)
{
  int which, col0;
  PyObject *blk;
  if (!PyArg_ParseTuple(args, "iiO", &which, &col0, &blk)) return NULL;
  PyArrayObject *a = (PyArrayObject *)PyArray_FROM_OTF(blk, NPY_DOUBLE, NPY_ARRAY_F_CONTIGUOUS | NPY_ARRAY_ALIGNED);
  if (!a) return NULL;
  int nd = PyArray_NDIM(a);
  if ((nd != 1 && nd != 2) || PyArray_DIM(a, 0) != self->n) {
    Py_DECREF(a);
    PyErr_SetString(PyExc_ValueError, "set_columns: block must have n rows");
    return NULL;
  }
  int ncols = nd == 1 ? 1 : (int)PyArray_DIM(a, 1);
  int st = psp_correq_set_columns(self->dev, which, col0, ncols, (const double *)PyArray_DATA(a), self->n);
  Py_DECREF(a);
  if (st != PSP_OK) return psp_raise(st);
  Py_RETURN_NONE;
}

static PyObject *CorrEq_update(CorrEqObject *self, PyObject *args)
{
  int k; double theta;
  PyObject *ho = Py_None;
  if (!PyArg_ParseTuple(args, "id|O", &k, &theta, &ho)) return NULL;
  PyArrayObject *h = NULL;
  if (ho != Py_None) {
    h = (PyArrayObject *)PyArray_FROM_OTF(ho, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
    if (!h) return NULL;
    if (PyArray_NDIM(h) != 2 || PyArray_DIM(h, 0) != k || PyArray_DIM(h, 1) != k) {
      Py_DECREF(h);
      PyErr_SetString(PyExc_ValueError, "update: hinv must be a k x k array");
      return NULL;
    }
  }
  int st = psp_correq_update(self->dev, k, theta, h ? (const double *)PyArray_DATA(h) : NULL);
  Py_XDECREF(h);
  if (st != PSP_OK) return psp_raise(st);
  Py_RETURN_NONE;
}

static PyObject *CorrEq_matvec(CorrEqObject *self, PyObject *args)
{
  double *x, *y;
  if (((int (*)(PyObject *, double **, double **, int))SpMatrix_API[PSP_API_ParseVecOpArgs])(args, &x, &y, self->n)) return NULL;
  int st = psp_matvec(psp_correq_mat(self->dev), 0, x, 1, y, 1);
  if (st != PSP_OK) return psp_raise(st);
  Py_RETURN_NONE;
}

static PyObject *CorrEq_precon(CorrEqObject *self, PyObject *args)
{
  double *x, *y;
  if (((int (*)(PyObject *, double **, double **, int))SpMatrix_API[PSP_API_ParseVecOpArgs])(args, &x, &y, self->n)) return NULL;
  int st = psp_prec_apply(psp_correq_prec(self->dev), x, y);
  if (st != PSP_OK) return psp_raise(st);
  Py_RETURN_NONE;
}

static void cap_noop(PyObject *c) { (void)c; }
static PyObject *CorrEq_get_shape(CorrEqObject *self, void *c) { return Py_BuildValue("(i,i)", self->n, self->n); }
static PyObject *CorrEq_get_mat(CorrEqObject *self, void *c) { return PyCapsule_New(psp_correq_mat(self->dev), PSP_MAT_CAPSULE, cap_noop); }
static PyObject *CorrEq_get_prec(CorrEqObject *self, void *c) { return PyCapsule_New(psp_correq_prec(self->dev), PSP_PREC_CAPSULE, cap_noop); }

static PyMethodDef CorrEq_methods[] = {
  {"set_columns", (PyCFunction)CorrEq_set_columns, METH_VARARGS, "self.set_columns(which, col0, block)\n\nupload columns col0.. of Q (which=0), M Q (1) or Y (2)."},
  {"update", (PyCFunction)CorrEq_update, METH_VARARGS, "self.update(k, theta, hinv=None)\n\nselect the number of columns in use, the shift and inv(H[:k, :k])."},
  {"matvec", (PyCFunction)CorrEq_matvec, METH_VARARGS, "self.matvec(x, y)\n\ny = projected (A - theta M) x."},
  {"precon", (PyCFunction)CorrEq_precon, METH_VARARGS, "self.precon(x, y)\n\ny = projected K^-1 x."},
  {NULL, NULL}
};
static PyGetSetDef CorrEq_getset[] = {
  {"shape", (getter)CorrEq_get_shape, NULL, NULL, NULL},
  {"_psp_mat", (getter)CorrEq_get_mat, NULL, "device handle of the operator as a matrix (PyCapsule)", NULL},
  {"_psp_prec", (getter)CorrEq_get_prec, NULL, "device handle of the operator as a preconditioner (PyCapsule)", NULL},
  {NULL}
};

static PyMethodDef correq_methods[] = {
  {"operator", correq_new, METH_VARARGS, "operator(optype, A, M, K, kmax)\n\nreturn a device-resident correction-equation operator."},
  {NULL, NULL}
};

static struct PyModuleDef correq_def = {PyModuleDef_HEAD_INIT, "correq",
  "pysparse_b200.eigen.correq -- the JDSYM correction-equation operator on the GPU", -1, correq_methods};

PyMODINIT_FUNC PyInit_correq(void)
{
  import_array();
  PyObject *sp = PyImport_ImportModule("pysparse_b200.sparse.spmatrix");
  if (!sp) return NULL;
  Py_DECREF(sp);
  SpMatrix_API = (void **)PyCapsule_Import(PSP_CAPI_NAME, 0);
  if (!SpMatrix_API) return NULL;
  static PyTypeObject zero = {PyVarObject_HEAD_INIT(NULL, 0)};
  CorrEqType = zero;
  CorrEqType.tp_name = "pysparse_b200.eigen.correq.operator";
  CorrEqType.tp_basicsize = sizeof(CorrEqObject);
  CorrEqType.tp_dealloc = (destructor)CorrEq_dealloc;
  CorrEqType.tp_flags = Py_TPFLAGS_DEFAULT;
  CorrEqType.tp_methods = CorrEq_methods;
  CorrEqType.tp_getset = CorrEq_getset;
  if (PyType_Ready(&CorrEqType) < 0) return NULL;
  PyObject *m = PyModule_Create(&correq_def);
  if (!m) return NULL;
  Py_INCREF(&CorrEqType); PyModule_AddObject(m, "OperatorType", (PyObject *)&CorrEqType);
  return m;
}
