"""ctypes view of the C-ABI in ``include/psp_b200.h`` (``libpsp_b200.so``).

The drop-in Python API (``pysparse_b200.sparse.spmatrix`` etc.) is implemented
by C extension modules that link the library directly; this module exists for
the benchmarks and parity tests, which also need the device-resident entry
points (``psp_solve_dev``, ``psp_matvec_dev``) that have no NumPy equivalent.

There is no CPU fallback: if the library is missing this raises ImportError, and
without a GPU every compute call raises ``PspError`` (PSP_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
#: PSP_B200_LIB selects another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("PSP_B200_LIB") or os.path.join(_PKG, "libpsp_b200.so")

#: every symbol include/psp_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "psp_init", "psp_device_count", "psp_current_device", "psp_last_error", "psp_version",
    "psp_launch_count", "psp_synchronize", "psp_host_alloc", "psp_host_free",
    "psp_csr_create", "psp_sss_create", "psp_mat_destroy", "psp_mat_info",
    "psp_matvec", "psp_matvec_dev", "psp_runs_analyze", "psp_runs_free",
    "psp_vec_alloc", "psp_vec_free", "psp_vec_upload", "psp_vec_download", "psp_vec_fill",
    "psp_jacobi_create", "psp_ssor_create", "psp_prec_destroy", "psp_prec_apply",
    "psp_prec_apply_dev", "psp_ssor_levels", "psp_prec_set_matrix",
    "psp_correq_create", "psp_correq_destroy", "psp_correq_set_columns", "psp_correq_update",
    "psp_correq_mat", "psp_correq_prec",
    "psp_solve", "psp_solve_dev", "psp_last_device_ms", "psp_set_option", "psp_get_option",
    "psp_event_record", "psp_event_elapsed_ms", "psp_spmv_timing",
    "psp_dist_plan", "psp_dist_init", "psp_dist_connect", "psp_dist_shutdown", "psp_dist_info",
    "psp_dist_csr_create", "psp_dist_mat_destroy", "psp_dist_mat_info", "psp_dist_solve_dev",
    "psp_dist_set_sendmap",
]

SOLVERS = {"pcg": 0, "minres": 1, "qmrs": 2, "bicgstab": 3, "cgs": 4, "gmres": 5}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _dp, _dp)


class Operator(C.Structure):
    _fields_ = [("mat", C.c_void_p), ("prec", C.c_void_p), ("cb", CALLBACK), ("ctx", C.c_void_p)]


class PspError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"psp_b200 error {status}: {msg}")
        self.status = status


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pysparse_b200.build` "
            "(pysparse_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.psp_last_error.restype = C.c_char_p
    L.psp_version.restype = C.c_char_p
    L.psp_launch_count.restype = C.c_longlong
    L.psp_launch_count.argtypes = [C.c_int]
    L.psp_last_device_ms.restype = C.c_double
    L.psp_get_option.restype = C.c_longlong
    L.psp_get_option.argtypes = [C.c_char_p]
    L.psp_set_option.argtypes = [C.c_char_p, C.c_longlong]
    L.psp_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.psp_host_free.argtypes = [C.c_void_p]
    L.psp_csr_create.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, C.POINTER(C.c_void_p)]
    L.psp_sss_create.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, C.POINTER(C.c_void_p)]
    L.psp_mat_destroy.argtypes = [C.c_void_p]
    L.psp_mat_destroy.restype = None
    L.psp_mat_info.argtypes = [C.c_void_p, _ip, _ip, C.POINTER(C.c_longlong), _ip]
    L.psp_matvec.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
    L.psp_matvec_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.psp_vec_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.psp_vec_free.argtypes = [C.c_void_p]
    L.psp_vec_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_long]
    L.psp_vec_download.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_long]
    L.psp_vec_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_double]
    L.psp_jacobi_create.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_int, C.POINTER(C.c_void_p)]
    L.psp_ssor_create.argtypes = [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_void_p)]
    L.psp_prec_destroy.argtypes = [C.c_void_p]
    L.psp_prec_destroy.restype = None
    L.psp_prec_apply.argtypes = [C.c_void_p, _dp, _dp]
    L.psp_prec_apply_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.psp_ssor_levels.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip]
    L.psp_event_record.argtypes = [C.c_int]
    L.psp_event_elapsed_ms.argtypes = [C.c_int, C.c_int, _dp]
    L.psp_spmv_timing.argtypes = [C.c_int, _ip, _dp]
    for name in ("psp_solve", "psp_solve_dev"):
        getattr(L, name).argtypes = [C.c_int, C.POINTER(Operator), C.POINTER(Operator), C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int,
                                     _ip, _ip, _dp]
    return L


lib = _load()


def check(status: int) -> None:
    if status != 0:
        raise PspError(status, lib.psp_last_error().decode(errors="replace"))


def device_count() -> int:
    return lib.psp_device_count()


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pinned_empty(n: int) -> np.ndarray:
    """float64 array of length n backed by page-locked host memory."""
    p = C.c_void_p()
    check(lib.psp_host_alloc(max(1, n) * 8, C.byref(p)))
    buf = (C.c_double * max(1, n)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.float64, count=n)
    _PINNED[arr.ctypes.data] = p.value
    return arr


_PINNED: dict[int, int] = {}


class DeviceVector:
    def __init__(self, n: int, host=None):
        self.n = n
        p = C.c_void_p()
        check(lib.psp_vec_alloc(n, C.byref(p)))
        self.ptr = p.value
        if host is not None:
            self.upload(host)

    def upload(self, host) -> None:
        host = _f64(host)
        assert host.size == self.n
        check(lib.psp_vec_upload(self.ptr, host.ctypes.data, self.n, 1))

    def download(self, out=None) -> np.ndarray:
        out = np.empty(self.n) if out is None else out
        check(lib.psp_vec_download(self.ptr, out.ctypes.data, self.n, 1))
        return out

    def fill(self, v: float) -> None:
        check(lib.psp_vec_fill(self.ptr, self.n, v))

    def __del__(self):
        try:
            if self.ptr:
                lib.psp_vec_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Matrix:
    """Device-resident CSR or SSS matrix (C-ABI handle)."""

    def __init__(self, ptr, kind, shape, nnz):
        self.ptr, self.kind, self.shape, self.nnz = ptr, kind, shape, nnz

    @classmethod
    def csr(cls, m, n, val, col, ind):
        val, col, ind = _f64(val), _i32(col), _i32(ind)
        p = C.c_void_p()
        check(lib.psp_csr_create(m, n, len(val), ind.ctypes.data_as(_ip), col.ctypes.data_as(_ip),
                                 val.ctypes.data_as(_dp), C.byref(p)))
        return cls(p.value, "csr", (m, n), len(val))

    @classmethod
    def sss(cls, n, val, diag, col, ind):
        val, diag, col, ind = _f64(val), _f64(diag), _i32(col), _i32(ind)
        p = C.c_void_p()
        check(lib.psp_sss_create(n, len(val), ind.ctypes.data_as(_ip), col.ctypes.data_as(_ip),
                                 val.ctypes.data_as(_dp), diag.ctypes.data_as(_dp), C.byref(p)))
        return cls(p.value, "sss", (n, n), len(val))

    def matvec(self, x, y, transp=False) -> None:
        """NumPy in / NumPy out (strided views allowed), like CSRMat_matvec."""
        assert x.dtype == np.float64 and y.dtype == np.float64 and x.ndim == 1 and y.ndim == 1
        incx = x.strides[0] // 8 if x.size > 1 else 1
        incy = y.strides[0] // 8 if y.size > 1 else 1
        check(lib.psp_matvec(self.ptr, int(transp), x.ctypes.data, incx, y.ctypes.data, incy))

    def matvec_dev(self, dx: DeviceVector, dy: DeviceVector, transp=False) -> None:
        check(lib.psp_matvec_dev(self.ptr, int(transp), dx.ptr, dy.ptr))

    def __del__(self):
        try:
            if self.ptr:
                lib.psp_mat_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Precon:
    def __init__(self, ptr, kind, n, keep=None):
        self.ptr, self.kind, self.n, self._keep = ptr, kind, n, keep
        self.shape = (n, n)

    @classmethod
    def jacobi(cls, mat: Matrix | None, diag, omega=1.0, steps=1):
        diag = _f64(diag)
        p = C.c_void_p()
        check(lib.psp_jacobi_create(mat.ptr if mat is not None else None, len(diag),
                                    diag.ctypes.data_as(_dp), omega, steps, C.byref(p)))
        return cls(p.value, "jacobi", len(diag), mat)

    @classmethod
    def ssor(cls, sss: Matrix, omega=1.0, steps=1):
        p = C.c_void_p()
        check(lib.psp_ssor_create(sss.ptr, omega, steps, C.byref(p)))
        return cls(p.value, "ssor", sss.shape[0], sss)

    def precon(self, x, y) -> None:
        x = _f64(x)
        assert y.dtype == np.float64 and y.flags.c_contiguous and y.size == self.n
        check(lib.psp_prec_apply(self.ptr, x.ctypes.data_as(_dp), y.ctypes.data_as(_dp)))

    def levels(self):
        nf, nb = C.c_int(), C.c_int()
        lf = np.empty(self.n, dtype=np.int32)
        lb = np.empty(self.n, dtype=np.int32)
        check(lib.psp_ssor_levels(self.ptr, C.byref(nf), C.byref(nb), lf.ctypes.data_as(_ip),
                                  lb.ctypes.data_as(_ip)))
        return nf.value, nb.value, lf, lb

    def __del__(self):
        try:
            if self.ptr:
                lib.psp_prec_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _operator(obj, n, keep):
    """Build a psp_operator from a Matrix / Precon / duck-typed Python object."""
    op = Operator()
    if obj is None:
        return None
    if isinstance(obj, Matrix):
        op.mat = obj.ptr
        return op
    if isinstance(obj, Precon):
        op.prec = obj.ptr
        return op
    meth = getattr(obj, "matvec", None) or getattr(obj, "precon")

    def tramp(_ctx, nn, xp, yp):
        try:
            x = np.ctypeslib.as_array(xp, shape=(nn,))
            y = np.ctypeslib.as_array(yp, shape=(nn,))
            meth(x, y)
            return 0
        except Exception as e:  # noqa: BLE001
            keep.append(e)
            return -1

    cb = CALLBACK(tramp)
    keep.append(cb)
    op.cb = cb
    return op


def solve(solver: str, A, b, x, tol, maxit, K=None, dim=20, dev=False):
    """(info, iter, relres) = solver(A, b, x, tol, maxit[, K[, dim]]).

    dev=False: b, x are NumPy arrays (x updated in place) -> psp_solve.
    dev=True : b, x are DeviceVector                     -> psp_solve_dev.
    """
    keep: list = []
    n = A.shape[0]
    opA = _operator(A, n, keep)
    opK = _operator(K, n, keep)
    info, it, rr = C.c_int(0), C.c_int(0), C.c_double(0.0)
    if dev:
        bp, xp = b.ptr, x.ptr
        fn = lib.psp_solve_dev
    else:
        b = _f64(b)
        assert x.dtype == np.float64 and x.flags.c_contiguous
        bp, xp = b.ctypes.data, x.ctypes.data
        fn = lib.psp_solve
    st = fn(SOLVERS[solver], C.byref(opA), C.byref(opK) if opK is not None else None, n, bp, xp,
            tol, maxit, dim, C.byref(info), C.byref(it), C.byref(rr))
    for k in keep:
        if isinstance(k, Exception):
            raise k
    check(st)
    return info.value, it.value, rr.value


def last_device_ms() -> float:
    return lib.psp_last_device_ms()


def launch_count(reset=False) -> int:
    return lib.psp_launch_count(int(reset))


def event_record(idx: int) -> None:
    check(lib.psp_event_record(idx))


def event_elapsed_ms(start: int, stop: int) -> float:
    ms = C.c_double()
    check(lib.psp_event_elapsed_ms(start, stop, C.byref(ms)))
    return ms.value


def spmv_timing(sample_every: int = 0):
    """(count, total_ms) of the sampled SpMV launches since the last call; re-arms sampling."""
    cnt, ms = C.c_int(), C.c_double()
    check(lib.psp_spmv_timing(sample_every, C.byref(cnt), C.byref(ms)))
    return cnt.value, ms.value


def set_option(name: str, value: int) -> None:
    check(lib.psp_set_option(name.encode(), value))


def get_option(name: str) -> int:
    return int(lib.psp_get_option(name.encode()))


def synchronize() -> None:
    check(lib.psp_synchronize())
