"""The device-resident Jacobi-Davidson correction equation (csrc/psp_correq.cu, eigen/correq)
against its host-side restatement (eigen/jdsym.py CorrEq, which follows
pysparse/eigen/src/correq.c:117-262 with NumPy) -- operator by operator, then through the
inner solvers, then through jdsym itself.

The projections are dense dot products (dgemv in the reference); their summation order is not
the host's, so the comparison is to rounding (1e-12 relative), not bit for bit.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(native):
    if native.device_count() == 0:
        pytest.fail("no CUDA device")
    from pysparse.sparse import spmatrix
    from pysparse.itsolvers import krylov
    from pysparse.precon import precon
    from pysparse.tools import poisson
    from pysparse.eigen import jdsym
    return spmatrix, krylov, precon, poisson, jdsym


def _problem(mods, m=24, kmax=7, seed=3):
    spmatrix, krylov, precon, poisson, jdsym = mods
    L = poisson.poisson2d_sym(m)
    n = m * m
    A = L.to_sss()
    Ml = spmatrix.ll_mat_sym(n)
    d = np.arange(n)
    Ml.put(1.0 + 0.5 * np.cos(d), d, d)
    Ml.put(0.05 * np.ones(n - 1), d[1:], d[:-1])
    Mm = Ml.to_sss()
    K = precon.jacobi(L)
    rng = np.random.default_rng(seed)
    Q = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, kmax)))[0])
    return n, A, Mm, K, Q, rng


def _blocks(n, A, M, K, Q, kmax):
    Qm = Y = H = None
    if M is not None:
        Qm = np.zeros((n, kmax), order="F")
        for j in range(kmax):
            t = np.empty(n); M.matvec(np.ascontiguousarray(Q[:, j]), t); Qm[:, j] = t
    if K is not None:
        Qb = Qm if M is not None else Q
        Y = np.zeros((n, kmax), order="F")
        for j in range(kmax):
            t = np.empty(n); K.precon(np.ascontiguousarray(Qb[:, j]), t); Y[:, j] = t
        H = Qb.T @ Y
    return Qm, Y, H


@pytest.mark.parametrize("optype", [1, 2])
@pytest.mark.parametrize("with_m", [False, True])
@pytest.mark.parametrize("with_k", [False, True])
def test_operator_matches_the_host_restatement(mods, optype, with_m, with_k):
    spmatrix, krylov, precon, poisson, jdsym = mods
    kmax = 7
    n, A, Mm, K, Q, rng = _problem(mods, kmax=kmax)
    M = Mm if with_m else None
    K = K if with_k else None
    Qm, Y, H = _blocks(n, A, M, K, Q, kmax)
    dev = jdsym.CorrEq(optype, n, A, M, K, Q, Qm, Y, H)
    host = jdsym.CorrEq(optype, n, A, M, K, Q, Qm, Y, H, device=False)
    assert dev._dev is not None and host._dev is None
    _ = dev._psp_mat, dev._psp_prec
    with pytest.raises(AttributeError):
        host._psp_mat
    x = rng.standard_normal(n)
    for k, theta in ((0, 0.3), (3, 0.0), (3, 0.7), (kmax, -1.25)):
        dev.update(k, theta)
        host.update(k, theta)
        for what in ("matvec", "precon"):
            yd, yh = np.full(n, np.nan), np.full(n, np.nan)
            getattr(dev, what)(x, yd)
            getattr(host, what)(x, yh)
            scale = max(np.linalg.norm(yh), 1.0)
            assert np.linalg.norm(yd - yh) <= 1e-12 * scale, (what, k, theta)
    # only the columns at and above `stable` are uploaded again: the device keeps its copy of the
    # first two although the host array handed in afterwards has them zeroed
    dev.update(kmax, 0.1, stable=2)
    Q2 = Q.copy(order="F"); Q2[:, :2] = 0.0; Q2[:, 4] = Q[:, 4]
    dev.Q = Q2
    dev.update(kmax, 0.1, stable=2)
    host.update(kmax, 0.1)
    yd, yh = np.empty(n), np.empty(n)
    dev.matvec(x, yd); host.matvec(x, yh)
    assert np.linalg.norm(yd - yh) <= 1e-12 * max(np.linalg.norm(yh), 1.0)


@pytest.mark.parametrize("solver", ["qmrs", "pcg", "minres", "bicgstab", "cgs", "gmres"])
def test_inner_solve_stays_on_the_device_and_agrees(mods, solver, monkeypatch):
    """jdsym.c:570 -- linsolver(correq, r, v, tol, linitmax, correq): with the device operator the
    solver must not call back into Python at all, and it must find what the callback path finds."""
    spmatrix, krylov, precon, poisson, jdsym = mods
    kmax = 5
    n, A, Mm, K, Q, rng = _problem(mods, m=20, kmax=kmax, seed=11)
    optype = 2 if solver in ("qmrs", "pcg", "minres") else 1
    Qm, Y, H = _blocks(n, A, None, K, Q, kmax)
    dev = jdsym.CorrEq(optype, n, A, None, K, Q, Qm, Y, H)
    host = jdsym.CorrEq(optype, n, A, None, K, Q, Qm, Y, H, device=False)
    theta = -0.4                                      # below the spectrum: the projected operator is definite
    dev.update(kmax, theta); host.update(kmax, theta)
    r = rng.standard_normal(n)
    rd, rh = r.copy(), r.copy()
    host.right(rh); dev.right(rd)
    assert np.array_equal(rd, rh)                     # right() is host arithmetic on both
    fn = getattr(krylov, solver)
    xh = np.zeros(n)
    res_h = fn(host, rh, xh, 1e-8, 300, host)
    calls = []
    monkeypatch.setattr(jdsym.CorrEq, "matvec", lambda self, x, y: calls.append("matvec"))
    monkeypatch.setattr(jdsym.CorrEq, "precon", lambda self, x, y: calls.append("precon"))
    xd = np.zeros(n)
    res_d = fn(dev, rd, xd, 1e-8, 300, dev)
    assert calls == []
    slack = 2 if solver in ("qmrs", "pcg", "minres", "gmres") else max(2, res_h[1] // 10)    # BiCGSTAB / CGS amplify rounding
    assert res_d[0] == res_h[0] and abs(res_d[1] - res_h[1]) <= slack, (res_d, res_h)
    assert np.linalg.norm(xd - xh) <= 1e-6 * np.linalg.norm(xh)


@pytest.mark.parametrize("case", ["standard", "generalised", "preconditioned", "unsym_preconditioned", "block"])
def test_jdsym_device_and_callback_operator_agree(mods, case, monkeypatch):
    spmatrix, krylov, precon, poisson, jdsym = mods
    m = 16
    L = poisson.poisson2d_sym(m)
    n = m * m
    # a graded diagonal separates the eigenvalues (a multiple eigenvalue's further copies are found
    # through rounding noise only, which the two operators do not share)
    L.put(4.0 + 0.5 * np.arange(n) / n, np.arange(n), np.arange(n))
    A = L.to_sss()
    Ml = spmatrix.ll_mat_sym(n)
    d = np.arange(n)
    Ml.put(1.0 + 0.25 * np.sin(d), d, d)
    M = Ml.to_sss() if case == "generalised" else None
    K = precon.jacobi(L) if "preconditioned" in case else None
    kw = dict(jmin=6, jmax=14)
    if case == "unsym_preconditioned":
        kw["optype"] = jdsym.JDSYM_OP_UNSYM
    if case == "block":
        kw.update(blksize=2)
    solver = krylov.gmres if case == "unsym_preconditioned" else krylov.qmrs
    out = {}
    for device in (False, True):
        monkeypatch.setattr(jdsym, "DEVICE_CORREQ", device)
        out[device] = jdsym.jdsym(A, M, K, 4, 0.0, 1e-9, 300, solver, **kw)
    (kh, lh, Qh, ith, inh), (kd, ld, Qd, itd, ind) = out[False], out[True]
    assert kh == kd == 4
    assert np.allclose(np.sort(ld), np.sort(lh), rtol=0, atol=1e-8)
    # the eigenvalues are right, not just equal to each other
    if M is None:
        T = 2 * np.eye(m) - np.eye(m, k=1) - np.eye(m, k=-1)
        dense = np.kron(np.eye(m), T) + np.kron(T, np.eye(m))
        dense[np.arange(n), np.arange(n)] = 4.0 + 0.5 * np.arange(n) / n
        w = np.linalg.eigvalsh(dense)
        assert np.allclose(np.sort(ld), w[:4], rtol=0, atol=1e-7)
    for j in range(kd):
        q = np.ascontiguousarray(Qd[:, j]); aq = np.empty(n); A.matvec(q, aq)
        mq = q
        if M is not None:
            mq = np.empty(n); M.matvec(q, mq)
        assert np.linalg.norm(aq - ld[j] * mq) <= 1e-8
    # same outer iteration path up to rounding: the counts may differ by a few, not by a factor
    assert abs(itd - ith) <= max(3, ith // 5), (itd, ith)


def test_operator_argument_errors(mods):
    spmatrix, krylov, precon, poisson, jdsym = mods
    from pysparse_b200.eigen import correq
    L = poisson.poisson2d_sym(6)
    A = L.to_sss()

    class Foreign:
        shape = (36, 36)

        def matvec(self, x, y):
            y[:] = x

    with pytest.raises(TypeError):
        correq.operator(2, Foreign(), None, None, 4)
    with pytest.raises(ValueError):
        correq.operator(2, A, None, None, 65)
    with pytest.raises(ValueError):
        correq.operator(3, A, None, None, 4)
    op = correq.operator(2, A, None, precon.jacobi(L), 4)
    assert op.shape == (36, 36)
    with pytest.raises(ValueError):
        op.set_columns(0, 3, np.zeros((36, 2), order="F"))     # past kmax
    with pytest.raises(ValueError):
        op.set_columns(1, 0, np.zeros((36, 1), order="F"))     # no mass matrix
    with pytest.raises(ValueError):
        op.update(2, 0.0)                                       # preconditioned: inv(H) is required
    with pytest.raises(ValueError):
        op.update(2, 0.0, np.eye(3))
    op.update(0, 0.5)
    x = np.arange(36.0); y = np.empty(36); ya = np.empty(36)
    op.matvec(x, y); A.matvec(x, ya)
    assert np.allclose(y, ya - 0.5 * x, rtol=0, atol=1e-13)
    # the foreign-operand fallback of jdsym's own operator
    ce = jdsym.CorrEq(2, 36, Foreign(), None, None, np.zeros((36, 4), order="F"), None, None, None)
    assert ce._dev is None
