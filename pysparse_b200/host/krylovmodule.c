/* krylovmodule.c -- module `krylov` of pysparse_b200: pcg, minres, qmrs,
 * bicgstab, cgs, gmres with the reference's call signatures.
 *
 * Mirrors pysparse/itsolvers/src/itsolversmodule.c:32-670:
 *     info, iter, relres = solver(A, b, x, tol, maxit[, K])      "OOOdi|O"
 *     info, iter, relres = gmres(A, b, x, tol, maxit[, K[, dim]]) "OOOdi|Oi"
 * A needs `.shape` and `.matvec(x, y)`, K `.shape` and `.precon(x, y)`; x is the
 * initial guess and is updated in place when it is a C-contiguous float64
 * array (otherwise, exactly like PyArray_AsCArray in the reference, a converted
 * copy is iterated on and the result is lost).
 *
 * The iteration itself runs on the GPU (psp_solve, include/psp_b200.h).  Native
 * matrices / preconditioners are recognised through their `_psp_mat` /
 * `_psp_prec` capsule attributes and keep the whole solve device-resident;
 * any other object is called back through its Python method with the operand
 * vectors staged in host memory (the reference's SpMatrix_Matvec / _Precon
 * convention, spmatrixmodule.c:169-248).
 */
#define PY_ARRAY_UNIQUE_SYMBOL psp_krylov_ARRAY_API
#include "psp_host.h"

static void **SpMatrix_API;

typedef struct {
  PyObject *obj;
  const char *method;
} py_operator;

/* psp_callback trampoline: wrap the raw buffers as ndarrays, call obj.method(x, y) */
static int call_python(void *ctx, int n, const double *x, double *y)
{
  py_operator *op = (py_operator *)ctx;
  npy_intp d = n;
  PyObject *xa = PyArray_SimpleNewFromData(1, &d, NPY_DOUBLE, (void *)x);
  PyObject *ya = xa ? PyArray_SimpleNewFromData(1, &d, NPY_DOUBLE, (void *)y) : NULL;
  PyObject *res = ya ? PyObject_CallMethod(op->obj, op->method, "OO", xa, ya) : NULL;
  Py_XDECREF(xa); Py_XDECREF(ya);
  if (!res) return -1;
  Py_DECREF(res);
  return 0;
}

/* resolve a Python object into a psp_operator; returns -1 with an exception set */
static int resolve(PyObject *obj, const char *attr, const char *capsule, const char *method, py_operator *pyop,
                   psp_operator *out)
{
  memset(out, 0, sizeof *out);
  PyObject *cap = PyObject_GetAttrString(obj, attr);
  if (cap) {
    void *h = PyCapsule_IsValid(cap, capsule) ? PyCapsule_GetPointer(cap, capsule) : NULL;
    Py_DECREF(cap);
    if (h) {
      if (!strcmp(capsule, PSP_MAT_CAPSULE)) out->mat = (psp_mat *)h;
      else out->prec = (psp_prec *)h;
      return 0;
    }
    PyErr_Clear();
  } else {
    if (!PyErr_ExceptionMatches(PyExc_AttributeError)) return -1;   /* e.g. no GPU: RuntimeError */
    PyErr_Clear();
  }
  pyop->obj = obj; pyop->method = method;
  out->cb = call_python; out->ctx = pyop;
  return 0;
}

static int get_shape(PyObject *op, int dim[2])
{
  return ((int (*)(PyObject *, int[]))SpMatrix_API[PSP_API_GetShape])(op, dim);
}

/* which: psp_solver id.  need_positive: minres/gmres/qmrs/cgs reject dim[0] <= 0
 * with "invalid matrix shape" (itsolversmodule.c:250,:346,:440,:533); pcg and
 * bicgstab use SpMatrix_GetOrder (:65,:157). */
static PyObject *solve_common(PyObject *args, int which, int need_positive)
{
  PyObject *amat, *bo, *xo, *precon = Py_None;
  double tol; int maxit, dim_gmres = 20;
  if (which == PSP_GMRES) {
    if (!PyArg_ParseTuple(args, "OOOdi|Oi", &amat, &bo, &xo, &tol, &maxit, &precon, &dim_gmres)) return NULL;
  } else {
    if (!PyArg_ParseTuple(args, "OOOdi|O", &amat, &bo, &xo, &tol, &maxit, &precon)) return NULL;
  }
  int dim[2], n;
  if (get_shape(amat, dim)) return NULL;
  if (need_positive) {
    if (dim[0] != dim[1] || dim[0] <= 0) { PyErr_SetString(PyExc_ValueError, "invalid matrix shape"); return NULL; }
  } else if (dim[0] != dim[1]) { PyErr_SetString(PyExc_ValueError, "matrix is not square"); return NULL; }
  n = dim[0];

  /* x: in place iff already a C-contiguous double array (PyArray_AsCArray semantics) */
  PyArrayObject *xa = (PyArrayObject *)PyArray_FROM_OTF(xo, NPY_DOUBLE, NPY_ARRAY_CARRAY);
  if (!xa) { PyErr_SetString(PyExc_ValueError, "Unable to convert x to double array"); return NULL; }
  PyArrayObject *ba = (PyArrayObject *)PyArray_FROM_OTF(bo, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!ba) { Py_DECREF(xa); PyErr_SetString(PyExc_ValueError, "Unable to convert b to double array"); return NULL; }
  if (PyArray_NDIM(xa) != 1 || PyArray_NDIM(ba) != 1 || PyArray_DIM(xa, 0) != PyArray_DIM(ba, 0) || PyArray_DIM(xa, 0) != n) {
    Py_DECREF(xa); Py_DECREF(ba);
    PyErr_SetString(PyExc_ValueError, "incompatible operand shapes");
    return NULL;
  }

  py_operator pyA, pyK;
  psp_operator A, K;
  PyObject *ret = NULL;
  if (resolve(amat, "_psp_mat", PSP_MAT_CAPSULE, "matvec", &pyA, &A)) goto done;
  if (precon != Py_None && resolve(precon, "_psp_prec", PSP_PREC_CAPSULE, "precon", &pyK, &K)) goto done;

  int info = 0, iter = 0; double relres = 0.0;
  int st = psp_solve(which, &A, precon != Py_None ? &K : NULL, n, (const double *)PyArray_DATA(ba),
                     (double *)PyArray_DATA(xa), tol, maxit, dim_gmres, &info, &iter, &relres);
  if (st != PSP_OK) { psp_raise(st); goto done; }
  if (PyErr_Occurred()) goto done;
  ret = Py_BuildValue("iid", info, iter, relres);
done:
  Py_DECREF(xa); Py_DECREF(ba);
  return ret;
}

static PyObject *ItSolvers_pcg(PyObject *self, PyObject *args) { return solve_common(args, PSP_PCG, 0); }
static PyObject *ItSolvers_bicgstab(PyObject *self, PyObject *args) { return solve_common(args, PSP_BICGSTAB, 0); }
static PyObject *ItSolvers_minres(PyObject *self, PyObject *args) { return solve_common(args, PSP_MINRES, 1); }
static PyObject *ItSolvers_gmres(PyObject *self, PyObject *args) { return solve_common(args, PSP_GMRES, 1); }
static PyObject *ItSolvers_qmrs(PyObject *self, PyObject *args) { return solve_common(args, PSP_QMRS, 1); }
static PyObject *ItSolvers_cgs(PyObject *self, PyObject *args) { return solve_common(args, PSP_CGS, 1); }

static PyMethodDef krylov_methods[] = {
  {"pcg", ItSolvers_pcg, METH_VARARGS, "info, iter, relres = pcg(A, b, x, tol, maxit)\n\nPreconditioned Conjugate Gradient method."},
  {"bicgstab", ItSolvers_bicgstab, METH_VARARGS, "info, iter, relres = bicgstab(A, b, x, tol, maxit)\n\nPreconditioned Stabilized BiConjugate Gradient method."},
  {"minres", ItSolvers_minres, METH_VARARGS, "info, iter, relres = minres(A, b, x, tol, maxit)\n\nMinimal Residual method."},
  {"gmres", ItSolvers_gmres, METH_VARARGS, "info, iter, relres = gmres(A, b, x, tol, maxit, K, dim)\n\nGeneral Minimal Residual method GMRES(m) (m=dim) of Saad and Schultz."},
  {"qmrs", ItSolvers_qmrs, METH_VARARGS, "info, iter, relres = qmrs(A, b, x, tol, maxit)\n\nQuasi Minimal Residual method for symmetric matrices."},
  {"cgs", ItSolvers_cgs, METH_VARARGS, "info, iter, relres = cgs(A, b, x, tol, maxit)\n\nConjugate Gradient Squared method."},
  {NULL, NULL}
};

static struct PyModuleDef krylov_def = {PyModuleDef_HEAD_INIT, "krylov",
  "pysparse_b200.itsolvers.krylov -- Krylov solvers running on the GPU", -1, krylov_methods};

PyMODINIT_FUNC PyInit_krylov(void)
{
  import_array();
  PyObject *sp = PyImport_ImportModule("pysparse_b200.sparse.spmatrix");
  if (!sp) return NULL;
  Py_DECREF(sp);
  SpMatrix_API = (void **)PyCapsule_Import(PSP_CAPI_NAME, 0);
  if (!SpMatrix_API) return NULL;
  return PyModule_Create(&krylov_def);
}
