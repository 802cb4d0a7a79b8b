/* psp_host.h -- shared declarations of the CPython host layer (modules
 * spmatrix, krylov, precon).  The host layer mirrors the reference's extension
 * modules (same names, call signatures, error behaviour) and forwards all
 * numerical work to the C-ABI in include/psp_b200.h.  It contains no numerical
 * kernels of its own: without the CUDA library / a GPU every compute call
 * raises RuntimeError.
 */
#ifndef PSP_HOST_H
#define PSP_HOST_H

#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <numpy/arrayobject.h>
#include "psp_b200.h"

/* ---- object layouts ------------------------------------------------------
 * The leading fields of csr_mat / sss_mat are the reference's
 * (pysparse/include/csr_mat.h:6-13, sss_mat.h:6-14) so C-level consumers of
 * the `_C_API` table that read them directly keep working; the device handle
 * is appended. */
typedef struct {
  PyObject_VAR_HEAD
  int dim[2];
  int nnz;
  double *val;
  int *col;
  int *ind;
  /* --- pysparse_b200 additions --- */
  psp_mat *dev;            /* device-resident copy, created on first use */
  PyObject *base[3];       /* arrays owning val/col/ind when built from NumPy arrays */
} CSRMatObject;

typedef struct {
  PyObject_VAR_HEAD
  int n;
  int nnz;
  double *val;
  double *diag;
  int *col;
  int *ind;
  /* --- pysparse_b200 additions --- */
  psp_mat *dev;
  PyObject *base[4];
} SSSMatObject;

/* ll_mat: host-side assembly container.  Unlike the reference's linked list
 * (pysparse/include/ll_mat.h) every row is a small sorted array, which keeps
 * insertion O(row length) but makes conversion and traversal contiguous. */
typedef struct {
  int len, cap;
  int *col;
  double *val;
} LLRow;

typedef struct {
  PyObject_VAR_HEAD
  int dim[2];
  int issym;
  int storeZeros;
  int nnz;
  LLRow *rows;
  /* device cache */
  unsigned long version, dev_version;
  psp_mat *dev;
} LLMatObject;

/* ---- the `_C_API` table: slot numbers of pysparse/include/spmatrix_api.h:10-72 */
enum {
  PSP_API_LLMatType = 0, PSP_API_CSRMatType = 1, PSP_API_SSSMatType = 2,
  PSP_API_ParseVecOpArgs = 3, PSP_API_GetShape = 4, PSP_API_GetOrder = 5, PSP_API_GetItem = 6,
  PSP_API_Matvec = 7, PSP_API_Precon = 8, PSP_API_NewLLMatObject = 9, PSP_API_LLMatGetItem = 10,
  PSP_API_LLMatSetItem = 11, PSP_API_LLMatUpdateItemAdd = 12, PSP_API_LLMatBuildColIndex = 13,
  PSP_API_LLMatDestroyColIndex = 14, PSP_API_ItSolvers_Solve = 15, PSP_API_pointers = 16
};
#define PSP_CAPI_NAME "pysparse_b200.sparse.spmatrix._C_API"

/* capsule names of the device handles exposed as attributes `_psp_mat` / `_psp_prec` */
#define PSP_MAT_CAPSULE "psp_mat"
#define PSP_PREC_CAPSULE "psp_prec"

/* map a psp_status to a Python exception (returns NULL for convenience) */
static PyObject *psp_raise(int status)
{
  const char *msg = psp_last_error();
  if (PyErr_Occurred()) return NULL;            /* e.g. raised inside a callback */
  switch (status) {
  case PSP_ERR_NOMEM: PyErr_SetString(PyExc_MemoryError, msg); break;
  case PSP_ERR_ARG:
  case PSP_ERR_STRUCTURE: PyErr_SetString(PyExc_ValueError, msg); break;
  default: PyErr_SetString(PyExc_RuntimeError, msg); break;
  }
  return NULL;
}

#endif
