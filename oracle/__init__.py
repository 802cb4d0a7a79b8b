"""oracle -- CPU checkers for the PySparse hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this package; nothing under ``pysparse_b200/`` does.

Two interchangeable back ends behind one ctypes facade:

* ``Oracle("port")`` -- ``oracle/libpsp_oracle.so``, this repo's plain-C
  restatement (``psp_oracle.c``; each routine cites the reference file:line).
* ``Oracle("ref")``  -- ``oracle/_ref/libpsp_ref.so``, the reference's OWN C
  kernels compiled from ``/root/reference`` (``oracle/Makefile``,
  ``ref_glue.c``).  Built in the build container; travels to the GPU box as a
  prebuilt file (``/root/reference`` does not exist there).

``tests/test_oracle.py`` pins "port" bit-exactly against "ref" and both against
``tests/golden/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libpsp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpsp_ref.so")

SOLVERS = {"pcg": 0, "minres": 1, "qmrs": 2, "bicgstab": 3, "cgs": 4, "gmres": 5}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _dp, _dp)


def build(ref: bool = True) -> None:
    """Compile the checkers (idempotent).  ``ref`` is skipped silently when
    ``/root/reference`` is absent (GPU box)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


class Handle:
    """Owns a C operator handle plus the NumPy arrays it borrows."""

    def __init__(self, lib, ptr, kind, keep, shape):
        self._lib, self.ptr, self.kind, self._keep, self.shape = lib, ptr, kind, keep, shape

    def __del__(self):
        try:
            if self.ptr:
                self._lib._free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Oracle:
    def __init__(self, which: str = "port"):
        assert which in ("port", "ref")
        self.which = which
        path = PORT_SO if which == "port" else REF_SO
        if not os.path.exists(path):
            build(ref=(which == "ref"))
        if not os.path.exists(path):
            raise RuntimeError(f"oracle library {path} is missing and could not be built")
        self.lib = L = C.CDLL(path)
        px = "orc_" if which == "port" else "ref_"
        self._px = px
        self._free = getattr(L, px + "op_free")
        self._free.argtypes = [C.c_void_p]
        self._free.restype = None
        for name in ("csr_new", "sss_new", "jacobi_new", "ssor_new", "callback_new"):
            getattr(L, px + name).restype = C.c_void_p
        getattr(L, px + "describe").restype = C.c_char_p
        self._solve = getattr(L, px + "solve")
        self._solve.restype = C.c_int
        self._solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, _dp, _dp,
                                C.c_double, C.c_int, C.c_int, _ip, _ip, _dp]
        self._precon = getattr(L, px + "precon")
        self._precon.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        self._precon.restype = C.c_int

    def describe(self) -> str:
        return getattr(self.lib, self._px + "describe")().decode()

    # -- operators ------------------------------------------------------
    def csr(self, m, n, val, col, ind) -> Handle:
        val, col, ind = _f64(val), _i32(col), _i32(ind)
        if self.which == "port":
            f = self.lib.orc_csr_new
            f.argtypes = [C.c_int, C.c_int, _dp, _ip, _ip]
            ptr = f(m, n, _p(val), _pi(col), _pi(ind))
        else:
            f = self.lib.ref_csr_new
            f.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _ip]
            ptr = f(m, n, len(val), _p(val), _pi(col), _pi(ind))
        return Handle(self, ptr, "csr", (val, col, ind), (m, n))

    def sss(self, n, val, diag, col, ind) -> Handle:
        val, diag, col, ind = _f64(val), _f64(diag), _i32(col), _i32(ind)
        if self.which == "port":
            f = self.lib.orc_sss_new
            f.argtypes = [C.c_int, _dp, _dp, _ip, _ip]
            ptr = f(n, _p(val), _p(diag), _pi(col), _pi(ind))
        else:
            f = self.lib.ref_sss_new
            f.argtypes = [C.c_int, C.c_int, _dp, _dp, _ip, _ip]
            ptr = f(n, len(val), _p(val), _p(diag), _pi(col), _pi(ind))
        return Handle(self, ptr, "sss", (val, diag, col, ind), (n, n))

    def callback(self, n, fn) -> Handle:
        """Duck-typed operator: fn(x: ndarray, y: ndarray) fills y."""
        def tramp(_ctx, nn, xp, yp):
            try:
                x = np.ctypeslib.as_array(xp, shape=(nn,))
                y = np.ctypeslib.as_array(yp, shape=(nn,))
                fn(x, y)
                return 0
            except Exception:
                return -1
        cb = CALLBACK(tramp)
        f = getattr(self.lib, self._px + "callback_new")
        f.argtypes = [C.c_int, CALLBACK, C.c_void_p]
        ptr = f(n, cb, None)
        return Handle(self, ptr, "callback", (cb, fn), (n, n))

    def jacobi(self, mat: Handle | None, diag, omega=1.0, steps=1) -> Handle:
        diag = _f64(diag)
        n = len(diag)
        f = getattr(self.lib, self._px + "jacobi_new")
        f.argtypes = [C.c_void_p, C.c_int, _dp, C.c_double, C.c_int]
        ptr = f(mat.ptr if mat is not None else None, n, _p(diag), omega, steps)
        if not ptr:
            raise ValueError("diagonal element close to zero")
        return Handle(self, ptr, "jacobi", (mat, diag), (n, n))

    def ssor(self, sss: Handle, omega=1.0, steps=1) -> Handle:
        assert sss.kind == "sss"
        f = getattr(self.lib, self._px + "ssor_new")
        f.argtypes = [C.c_void_p, C.c_double, C.c_int]
        ptr = f(sss.ptr, omega, steps)
        return Handle(self, ptr, "ssor", (sss,), sss.shape)

    # -- kernels ----------------------------------------------------------
    @staticmethod
    def _strided(a):
        assert a.dtype == np.float64 and a.ndim == 1
        inc = a.strides[0] // 8 if a.size > 1 else 1
        assert inc >= 1
        return a.ctypes.data_as(_dp), inc

    def matvec(self, A: Handle, x, y, transp=False) -> None:
        """y := A x  (or A^T x); x, y are 1-D float64 arrays, possibly strided
        views, y written in place -- CSRMat_matvec / SSSMat_matvec semantics."""
        xp, incx = self._strided(x)
        yp, incy = self._strided(y)
        v = A._keep
        if A.kind == "sss":
            if self.which == "port":
                f = self.lib.orc_sss_matvec
                f.argtypes = [C.c_int, _dp, _dp, _ip, _ip, _dp, C.c_int, _dp, C.c_int]
                f.restype = None
                f(A.shape[0], _p(v[0]), _p(v[1]), _pi(v[2]), _pi(v[3]), xp, incx, yp, incy)
            else:
                f = self.lib.ref_sss_matvec
                f.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int]
                f.restype = None
                f(A.ptr, xp, incx, yp, incy)
            return
        assert A.kind == "csr"
        m, n = A.shape
        if self.which == "port":
            if transp:
                f = self.lib.orc_csr_matvec_transp
                f.argtypes = [C.c_int, C.c_int, _dp, _ip, _ip, _dp, C.c_int, _dp, C.c_int]
                f.restype = None
                f(m, n, _p(v[0]), _pi(v[1]), _pi(v[2]), xp, incx, yp, incy)
            else:
                f = self.lib.orc_csr_matvec
                f.argtypes = [C.c_int, _dp, _ip, _ip, _dp, C.c_int, _dp, C.c_int]
                f.restype = None
                f(m, _p(v[0]), _pi(v[1]), _pi(v[2]), xp, incx, yp, incy)
        else:
            f = self.lib.ref_csr_matvec_transp if transp else self.lib.ref_csr_matvec
            f.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int]
            f.restype = None
            f(A.ptr, xp, incx, yp, incy)

    def precon(self, K: Handle, x, y) -> None:
        x = _f64(x)
        assert y.dtype == np.float64 and y.flags.c_contiguous
        rc = self._precon(K.ptr, len(x), _p(x), _p(y))
        if rc:
            raise RuntimeError("precon failed")

    def solve(self, solver: str, A: Handle, b, x, tol, maxit, K: Handle | None = None, dim=20):
        """Returns (info, iter, relres) exactly as krylov.<solver> would; x is
        updated in place (must be C-contiguous float64)."""
        b = _f64(b)
        assert x.dtype == np.float64 and x.flags.c_contiguous
        info, it, rr = C.c_int(0), C.c_int(0), C.c_double(0.0)
        rc = self._solve(SOLVERS[solver], A.ptr, K.ptr if K is not None else None, len(b),
                         _p(b), _p(x), tol, maxit, dim, C.byref(info), C.byref(it), C.byref(rr))
        if rc == -100:
            raise ValueError(solver)
        return info.value, it.value, rr.value


class RefCorrEq:
    """The reference's own JDSYM correction-equation operator (``pysparse/eigen/src/correq.c``,
    compiled into ``oracle/_ref``): ``update(k, theta, H)``, ``right(r)`` (in place),
    ``matvec(x, y)``, ``precon(x, y)``.  ``A``, ``M``, ``K`` are operator handles of the same
    ``Oracle("ref")``; ``Q``, ``Qm``, ``Y`` are n x kmax Fortran-ordered arrays that the operator
    borrows (the caller keeps writing into them, as jdsym.c does); ``H`` is kmax x kmax."""

    def __init__(self, ora: Oracle, optype, n, A: Handle, M: Handle | None, K: Handle | None, Q, Qm, Y, kmax):
        assert ora.which == "ref", "only the compiled reference has a correction equation"
        L = ora.lib
        if not hasattr(L, "ref_correq_new"):
            raise RuntimeError("oracle/_ref was built without correq.c (rebuild: make -C oracle ref)")
        for a in (Q, Qm, Y):
            assert a is None or (a.dtype == np.float64 and a.flags.f_contiguous and a.shape == (n, kmax))
        self._lib, self.n, self.kmax = L, n, kmax
        self._keep = (A, M, K, Q, Qm, Y)
        L.ref_correq_new.restype = C.c_void_p
        L.ref_correq_new.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _dp, _dp, _dp, C.c_int]
        L.ref_correq_update.argtypes = [C.c_void_p, C.c_int, C.c_double, _dp]
        L.ref_correq_update.restype = C.c_int
        for name in ("ref_correq_matvec", "ref_correq_precon"):
            getattr(L, name).argtypes = [C.c_void_p, _dp, _dp]
            getattr(L, name).restype = None
        L.ref_correq_right.argtypes = [C.c_void_p, _dp]
        L.ref_correq_right.restype = None
        L.ref_correq_free.argtypes = [C.c_void_p]
        L.ref_correq_free.restype = None
        self.ptr = L.ref_correq_new(optype, n, A.ptr, M.ptr if M is not None else None, K.ptr if K is not None else None,
                                    _p(Q), _p(Qm) if Qm is not None else None, _p(Y) if Y is not None else None, kmax)

    def update(self, k, theta, H=None):
        Hf = None
        if H is not None:
            Hf = np.asfortranarray(H, dtype=np.float64)
            assert Hf.shape == (self.kmax, self.kmax)
        info = self._lib.ref_correq_update(self.ptr, k, theta, _p(Hf) if Hf is not None else None)
        if info:
            raise ValueError("H is singular")

    def right(self, r):
        assert r.dtype == np.float64 and r.flags.c_contiguous
        self._lib.ref_correq_right(self.ptr, _p(r))

    def matvec(self, x, y):
        x = _f64(x)
        assert y.dtype == np.float64 and y.flags.c_contiguous
        self._lib.ref_correq_matvec(self.ptr, _p(x), _p(y))

    def precon(self, x, y):
        x = _f64(x)
        assert y.dtype == np.float64 and y.flags.c_contiguous
        self._lib.ref_correq_precon(self.ptr, _p(x), _p(y))

    def __del__(self):
        try:
            if self.ptr:
                self._lib.ref_correq_free(self.ptr)
                self.ptr = None
        except Exception:
            pass
