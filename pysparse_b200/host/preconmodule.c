/* preconmodule.c -- module `precon` of pysparse_b200: jacobi and ssor
 * preconditioner objects with the reference's interface.
 *
 * Mirrors pysparse/precon/src/preconmodule.c:
 *     K = precon.jacobi(A, omega=1.0, steps=1)     ("O|di",  :470-485)
 *     K = precon.ssor(S, omega=1.0, steps=1)       ("O!|di", :499-514; S must be an sss_mat)
 *     K.shape, K.precon(x, y)                      (x, y contiguous float64 of length n)
 * The sweeps themselves run on the GPU (psp_jacobi_create / psp_ssor_create /
 * psp_prec_apply).  The objects also expose `_psp_prec`, the capsule through
 * which krylov keeps a solve device-resident.
 */
#define PY_ARRAY_UNIQUE_SYMBOL psp_precon_ARRAY_API
#include "psp_host.h"

static void **SpMatrix_API;
#define SSSMatTypePtr ((PyTypeObject *)SpMatrix_API[PSP_API_SSSMatType])

typedef struct {
  PyObject_HEAD
  int n;
  PyObject *matrix;      /* strong reference, like the reference objects (:404, :451) */
  psp_prec *dev;
  double omega;
  int steps;
} PreconObject;

static PyTypeObject JacobiType, SSORType;

static void Precon_dealloc(PreconObject *self)
{
  if (self->dev) psp_prec_destroy(self->dev);
  Py_XDECREF(self->matrix);
  PyObject_Del(self);
}

static psp_mat *matrix_handle(PyObject *matrix, int *failed);

/* A multi-step Jacobi preconditioner multiplies with its matrix.  The device copy of an
 * ll_mat is rebuilt (and the old handle destroyed) whenever the matrix has been modified, so
 * the handle is looked up again before every use instead of being cached at construction
 * (the reference, too, calls the live matrix: preconmodule.c:45-52). */
static int refresh_matrix(PreconObject *self)
{
  if (Py_TYPE(self) != &JacobiType || self->steps <= 1) return 0;
  int failed;
  psp_mat *h = matrix_handle(self->matrix, &failed);
  if (!h) {
    if (!failed) PyErr_SetString(PyExc_RuntimeError, "jacobi: the matrix has no device handle any more");
    return -1;
  }
  int st = psp_prec_set_matrix(self->dev, h);
  if (st != PSP_OK) { psp_raise(st); return -1; }
  return 0;
}

static PyObject *Precon_precon(PreconObject *self, PyObject *args)
{
  double *x, *y;
  if (((int (*)(PyObject *, double **, double **, int))SpMatrix_API[PSP_API_ParseVecOpArgs])(args, &x, &y, self->n)) return NULL;
  if (refresh_matrix(self)) return NULL;
  int st = psp_prec_apply(self->dev, x, y);
  if (st != PSP_OK) return psp_raise(st);
  Py_RETURN_NONE;
}

static void cap_noop(PyObject *c) { (void)c; }
static PyObject *Precon_get_shape(PreconObject *self, void *c) { return Py_BuildValue("(i,i)", self->n, self->n); }
static PyObject *Precon_get_dev(PreconObject *self, void *c)
{
  if (refresh_matrix(self)) return NULL;        /* krylov reads this attribute at the start of every solve */
  return PyCapsule_New(self->dev, PSP_PREC_CAPSULE, cap_noop);
}
static PyObject *Precon_get_levels(PreconObject *self, void *c)
{
  int nf = 0, nb = 0;
  if (psp_ssor_levels(self->dev, &nf, &nb, NULL, NULL) != PSP_OK) return psp_raise(PSP_ERR_ARG);
  return Py_BuildValue("(ii)", nf, nb);
}

static PyMethodDef Precon_methods[] = {
  {"precon", (PyCFunction)Precon_precon, METH_VARARGS, "self.precon(x, y)\n\napply preconditioner self on x, store result in y. x is unchanged."},
  {NULL, NULL}
};
static PyGetSetDef Jacobi_getset[] = {
  {"shape", (getter)Precon_get_shape, NULL, NULL, NULL},
  {"_psp_prec", (getter)Precon_get_dev, NULL, "device handle (PyCapsule)", NULL},
  {NULL}
};
static PyGetSetDef SSOR_getset[] = {
  {"shape", (getter)Precon_get_shape, NULL, NULL, NULL},
  {"_psp_prec", (getter)Precon_get_dev, NULL, "device handle (PyCapsule)", NULL},
  {"levels", (getter)Precon_get_levels, NULL, "(forward, backward) number of dependency levels of the sweeps", NULL},
  {NULL}
};

/* device handle of a native matrix or NULL (no exception) */
static psp_mat *matrix_handle(PyObject *matrix, int *failed)
{
  *failed = 0;
  PyObject *cap = PyObject_GetAttrString(matrix, "_psp_mat");
  if (!cap) {
    if (PyErr_ExceptionMatches(PyExc_AttributeError)) { PyErr_Clear(); return NULL; }
    *failed = 1;
    return NULL;
  }
  psp_mat *h = PyCapsule_IsValid(cap, PSP_MAT_CAPSULE) ? (psp_mat *)PyCapsule_GetPointer(cap, PSP_MAT_CAPSULE) : NULL;
  Py_DECREF(cap);
  return h;
}

/* jacobi(A, omega=1.0, steps=1)   (newJacobiObject, preconmodule.c:352-412) */
static PyObject *jacobi_prec(PyObject *self, PyObject *args)
{
  PyObject *matrix; double omega = 1.0; int steps = 1, n;
  if (!PyArg_ParseTuple(args, "O|di", &matrix, &omega, &steps)) return NULL;
  if (((int (*)(PyObject *, int *))SpMatrix_API[PSP_API_GetOrder])(matrix, &n)) return NULL;
  double *diag = PyMem_New(double, n > 0 ? n : 1);
  if (!diag) return PyErr_NoMemory();
  double (*get)(PyObject *, int, int) = (double (*)(PyObject *, int, int))SpMatrix_API[PSP_API_GetItem];
  for (int i = 0; i < n; i++) {          /* A[i,i]; native types are read directly */
    diag[i] = get(matrix, i, i);
    if (PyErr_Occurred()) { PyMem_Free(diag); return NULL; }
  }
  psp_mat *h = NULL;
  if (steps > 1) {
    int failed;
    h = matrix_handle(matrix, &failed);
    if (failed) { PyMem_Free(diag); return NULL; }
    if (!h) {
      PyMem_Free(diag);
      PyErr_SetString(PyExc_NotImplementedError, "jacobi with steps > 1 needs a pysparse_b200 matrix (ll_mat, csr_mat or sss_mat)");
      return NULL;
    }
  }
  psp_prec *dev = NULL;
  int st = psp_jacobi_create(h, n, diag, omega, steps, &dev);
  PyMem_Free(diag);
  if (st != PSP_OK) return psp_raise(st);   /* "diagonal element close to zero" -> ValueError (:395) */
  PreconObject *op = PyObject_New(PreconObject, &JacobiType);
  if (!op) { psp_prec_destroy(dev); return PyErr_NoMemory(); }
  op->n = n; op->omega = omega; op->steps = steps; op->dev = dev;
  Py_INCREF(matrix); op->matrix = matrix;
  return (PyObject *)op;
}

/* ssor(S, omega=1.0, steps=1)   (newSSORObject, preconmodule.c:414-459) */
static PyObject *ssor_prec(PyObject *self, PyObject *args)
{
  PyObject *matrix; double omega = 1.0; int steps = 1;
  if (!PyArg_ParseTuple(args, "O!|di", SSSMatTypePtr, &matrix, &omega, &steps)) return NULL;
  int failed;
  psp_mat *h = matrix_handle(matrix, &failed);
  if (!h) { if (!failed) PyErr_SetString(PyExc_RuntimeError, "sss_mat has no device handle"); return NULL; }
  psp_prec *dev = NULL;
  int st = psp_ssor_create(h, omega, steps, &dev);
  if (st != PSP_OK) return psp_raise(st);
  PreconObject *op = PyObject_New(PreconObject, &SSORType);
  if (!op) { psp_prec_destroy(dev); return PyErr_NoMemory(); }
  op->n = ((SSSMatObject *)matrix)->n; op->omega = omega; op->steps = steps; op->dev = dev;
  Py_INCREF(matrix); op->matrix = matrix;
  return (PyObject *)op;
}

static PyMethodDef precon_methods[] = {
  {"jacobi", jacobi_prec, METH_VARARGS, "jacobi(A, omega=1.0, steps=1)\n\nreturn Jacobi preconditioner object."},
  {"ssor", ssor_prec, METH_VARARGS, "ssor(A, omega, steps) -- return SSOR preconditioner object\n\n"
   "This preconditioner executes 'steps' SSOR steps with a zero initial guess.\n\nA      'sss_mat' object, symmetric sparse matrix\n"
   "omega  relaxation parameter (default value: 1.0)\nsteps  number of SSOR steps"},
  {NULL, NULL}
};

static struct PyModuleDef precon_def = {PyModuleDef_HEAD_INIT, "precon",
  "pysparse_b200.precon.precon -- Jacobi and SSOR preconditioners running on the GPU", -1, precon_methods};

static int init_type(PyTypeObject *t, const char *name, PyGetSetDef *gs)
{
  t->tp_name = name;
  t->tp_basicsize = sizeof(PreconObject);
  t->tp_dealloc = (destructor)Precon_dealloc;
  t->tp_flags = Py_TPFLAGS_DEFAULT;
  t->tp_methods = Precon_methods;
  t->tp_getset = gs;
  return PyType_Ready(t);
}

PyMODINIT_FUNC PyInit_precon(void)
{
  import_array();
  PyObject *sp = PyImport_ImportModule("pysparse_b200.sparse.spmatrix");
  if (!sp) return NULL;
  Py_DECREF(sp);
  SpMatrix_API = (void **)PyCapsule_Import(PSP_CAPI_NAME, 0);
  if (!SpMatrix_API) return NULL;
  static PyTypeObject zero = {PyVarObject_HEAD_INIT(NULL, 0)};
  JacobiType = zero; SSORType = zero;
  if (init_type(&JacobiType, "pysparse_b200.precon.precon.jacobi", Jacobi_getset) < 0 ||
      init_type(&SSORType, "pysparse_b200.precon.precon.ssor", SSOR_getset) < 0)
    return NULL;
  PyObject *m = PyModule_Create(&precon_def);
  if (!m) return NULL;
  Py_INCREF(&JacobiType); PyModule_AddObject(m, "JacobiType", (PyObject *)&JacobiType);
  Py_INCREF(&SSORType); PyModule_AddObject(m, "SSORType", (PyObject *)&SSORType);
  return m;
}
