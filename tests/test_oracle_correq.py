"""Pins the NumPy restatement of the JDSYM correction equation (pysparse_b200/eigen/jdsym.py
CorrEq with device=False -- the checker of the device-resident operator in
tests/test_gpu_correq.py) against the reference's OWN operator: pysparse/eigen/src/correq.c,
compiled into oracle/_ref together with the kernels it calls (oracle/Makefile, ref_glue.c).

Both sides multiply with the same matrices through the same reference kernels; what differs is the
order of the dense dot products (dgemv there, NumPy here) and LU solve vs. NumPy solve: agreement to
1e-12 relative, every (optype, M, K) combination, matvec / precon / right.
"""
import numpy as np
import pytest

import oracle
from pysparse_b200.eigen import jdsym
from pysparse_b200.tools import gallery


class Op:
    """duck-typed operand for the NumPy operator, backed by the same reference kernels"""

    def __init__(self, ora, handle, what):
        self.ora, self.h, self.what = ora, handle, what
        self.shape = handle.shape

    def matvec(self, x, y):
        self.ora.matvec(self.h, np.ascontiguousarray(x), y)

    def precon(self, x, y):
        self.ora.precon(self.h, np.ascontiguousarray(x), y)


@pytest.mark.parametrize("optype", [jdsym.JDSYM_OP_UNSYM, jdsym.JDSYM_OP_SYM])
@pytest.mark.parametrize("with_m", [False, True])
@pytest.mark.parametrize("with_k", [False, True])
def test_numpy_operator_equals_the_reference_operator(ref, optype, with_m, with_k):
    rng = np.random.default_rng(7 + 4 * optype + 2 * with_m + with_k)
    mat = gallery.poisson2d(17)
    n, kmax = mat["n"], 6
    val, diag, col, ind = mat["sss"]
    A = ref.sss(n, val, diag, col, ind)
    mm = gallery.random_sym(n, 4, seed=5)
    M = ref.sss(n, *mm["sss"]) if with_m else None
    K = ref.jacobi(A, diag) if with_k else None
    Q = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, kmax)))[0])
    Qm = Y = H = None
    if with_m:
        Qm = np.zeros((n, kmax), order="F")
        for j in range(kmax):
            t = np.zeros(n); ref.matvec(M, np.ascontiguousarray(Q[:, j]), t); Qm[:, j] = t
    if with_k:
        Qb = Qm if with_m else Q
        Y = np.zeros((n, kmax), order="F")
        for j in range(kmax):
            t = np.zeros(n); ref.precon(K, np.ascontiguousarray(Qb[:, j]), t); Y[:, j] = t
        H = Qb.T @ Y
    theirs = oracle.RefCorrEq(ref, optype, n, A, M, K, Q, Qm, Y, kmax)
    ours = jdsym.CorrEq(optype, n, Op(ref, A, "A"), Op(ref, M, "M") if with_m else None, Op(ref, K, "K") if with_k else None,
                        Q, Qm, Y, H, device=False)
    x = rng.standard_normal(n)
    for k, theta in ((0, 0.4), (1, 0.0), (4, -0.3), (kmax, 1.7)):
        theirs.update(k, theta, H)
        ours.update(k, theta)
        for what in ("matvec", "precon"):
            yt, yo = np.full(n, np.nan), np.full(n, np.nan)
            getattr(theirs, what)(x, yt)
            getattr(ours, what)(x, yo)
            assert np.linalg.norm(yo - yt) <= 1e-12 * max(1.0, np.linalg.norm(yt)), (what, k, theta)
        rt, ro = x.copy(), x.copy()
        theirs.right(rt)
        ours.right(ro)
        assert np.linalg.norm(ro - rt) <= 1e-12 * max(1.0, np.linalg.norm(rt)), ("right", k, theta)


def test_inner_solve_with_the_reference_operator_and_the_reference_solver(ref):
    """jdsym.c:570: the correction equation solved by the reference's QMRS through the reference's
    operator, against the same solve through the NumPy operator (callbacks): same iteration count,
    same solution -- the inner-solve semantics the device path is tested against on the GPU."""
    rng = np.random.default_rng(3)
    mat = gallery.poisson2d(15)
    n, kmax = mat["n"], 4
    val, diag, col, ind = mat["sss"]
    A = ref.sss(n, val, diag, col, ind)
    K = ref.jacobi(A, diag)
    Q = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, kmax)))[0])
    Y = np.zeros((n, kmax), order="F")
    for j in range(kmax):
        t = np.zeros(n); ref.precon(K, np.ascontiguousarray(Q[:, j]), t); Y[:, j] = t
    H = Q.T @ Y
    theirs = oracle.RefCorrEq(ref, jdsym.JDSYM_OP_SYM, n, A, None, K, Q, None, Y, kmax)
    ours = jdsym.CorrEq(jdsym.JDSYM_OP_SYM, n, Op(ref, A, "A"), None, Op(ref, K, "K"), Q, None, Y, H, device=False)
    theirs.update(kmax, -0.5, H); ours.update(kmax, -0.5)
    r = rng.standard_normal(n)
    rt, ro = r.copy(), r.copy()
    theirs.right(rt); ours.right(ro)
    At = ref.callback(n, theirs.matvec); Kt = ref.callback(n, theirs.precon)
    Ao = ref.callback(n, ours.matvec); Ko = ref.callback(n, ours.precon)
    xt, xo = np.zeros(n), np.zeros(n)
    res_t = ref.solve("qmrs", At, rt, xt, 1e-9, 300, Kt)
    res_o = ref.solve("qmrs", Ao, ro, xo, 1e-9, 300, Ko)
    assert res_t[0] == res_o[0] == 0 and abs(res_t[1] - res_o[1]) <= 1
    assert np.linalg.norm(xt - xo) <= 1e-7 * np.linalg.norm(xt)
