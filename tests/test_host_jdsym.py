"""CPU tests of the JDSYM outer iteration with NumPy stand-ins
for the GPU pieces: operators with ``shape`` / ``matvec`` / ``precon`` and an inner solver with
the ``krylov`` signature (right-preconditioned GMRES, started from zero so that its iterates
stay in the range of the projected operator).  The cases are the reference's own:
test/test_jdsym.py:10-16 (3 x 3 diagonal matrix) and examples/jdsym_test.py:36-141 (diagonal
pencil: standard, generalised, preconditioned, with start vectors), plus a 2-D Poisson
spectrum against LAPACK.  The same cases run on the GPU path in tests/test_gpu_dropin.py."""
import numpy as np
import pytest


class NpMat:
    def __init__(self, D):
        self.D = np.asarray(D, dtype=float)
        self.shape = self.D.shape
        self.calls = 0

    def matvec(self, x, y):
        self.calls += 1
        y[:] = self.D @ x

    matvec_transp = matvec


class NpDiagPrec:
    def __init__(self, d):
        self.dinv = 1.0 / np.asarray(d, dtype=float)
        self.shape = (len(self.dinv), len(self.dinv))

    def precon(self, x, y):
        y[:] = x * self.dinv


def np_gmres(A, b, x, tol, maxit, K=None):
    """Right-preconditioned full GMRES with the krylov calling convention: x updated in place,
    returns (info, iter, relres)."""
    n = len(b)

    def op(z):
        t = np.empty(n)
        if K is not None:
            K.precon(z, t)
        else:
            t[:] = z
        y = np.empty(n); A.matvec(t, y)
        return y

    beta = np.linalg.norm(b)
    if beta == 0.0:
        x[:] = 0.0
        return 0, 0, 0.0
    m = min(maxit, n)
    Vk = np.zeros((n, m + 1)); Hk = np.zeros((m + 1, m))
    Vk[:, 0] = b / beta
    relres, it = 1.0, 0
    yk = np.zeros(0)
    for i in range(m):
        w = op(Vk[:, i])
        for l in range(i + 1):
            Hk[l, i] = Vk[:, l] @ w
            w -= Hk[l, i] * Vk[:, l]
        Hk[i + 1, i] = np.linalg.norm(w)
        it = i + 1
        e1 = np.zeros(i + 2); e1[0] = beta
        yk, *_ = np.linalg.lstsq(Hk[:i + 2, :i + 1], e1, rcond=None)
        relres = np.linalg.norm(Hk[:i + 2, :i + 1] @ yk - e1) / beta
        if relres <= tol or Hk[i + 1, i] < 1e-14 * beta:
            break
        Vk[:, i + 1] = w / Hk[i + 1, i]
    z = Vk[:, :len(yk)] @ yk
    if K is not None:
        t = np.empty(n); K.precon(z, t); z = t
    x[:] = z
    return (0 if relres <= tol else -1), it, relres


@pytest.fixture(scope="module")
def jd(native):
    from pysparse.eigen import jdsym
    return jdsym


def residuals(A, M, lam, Q):
    out = []
    for k in range(len(lam)):
        q = Q[:, k].copy()
        r = A.D @ q - lam[k] * (M.D @ q if M is not None else q)
        out.append(np.linalg.norm(r))
    return np.array(out)


def test_reference_unit_test_diagonal_3x3(jd):
    a = NpMat(np.diag([1.0, 2.0, 3.0]))
    r = jd.jdsym(a, None, None, 3, 1, 1e-9, 100, np_gmres)
    assert len(r) == 5 and r[0] == 3
    assert np.allclose(r[1], [1, 2, 3])
    assert np.allclose(np.abs(r[2]), np.identity(3))
    assert r[2].shape == (3, 3) and r[2].flags.c_contiguous


def test_reference_example_diagonal_pencil(jd):
    """examples/jdsym_test.py with n = 300: M = K = None; K = None; general; start vectors."""
    n, ncv, tol = 300, 5, 1e-6
    dA = np.arange(1.0, n + 1.0); dM = float(n // 2) + np.arange(n)
    A, M = NpMat(np.diag(dA)), NpMat(np.diag(dM))
    K = NpDiagPrec(dA - 0.006 * dM)
    normM = dM[-1]
    kw = dict(jmin=5, jmax=10, eps_tr=1e-4)
    kconv, lam, Q, it, it_inner = jd.jdsym(A, None, None, ncv, 0.0, tol, 150, np_gmres, **kw)
    assert kconv == ncv and it <= 150 and it_inner > 0
    assert np.all(residuals(A, None, lam, Q) <= tol) and np.allclose(lam, dA[:ncv], rtol=tol * tol, atol=1e-9)
    assert np.allclose(Q.T @ Q, np.eye(ncv), atol=1e-8)
    exact = (dA / dM)[:ncv]
    for Kc, extra in ((None, {}), (K, {}), (None, "V0")):
        if extra == "V0":
            X0 = np.random.default_rng(0).random((n, ncv))
            X0[np.arange(ncv), np.arange(ncv)] = 10000.0
            extra = dict(V0=X0)
        kconv, lam, Q, it, it_inner = jd.jdsym(A, M, Kc, ncv, 0.0, tol, 150, np_gmres, **kw, **extra)
        assert kconv == ncv
        assert np.all(residuals(A, M, lam, Q) <= normM * tol)
        assert np.allclose(np.sort(lam), exact, rtol=0, atol=normM * tol * tol + 1e-9)
        assert np.allclose(Q.T @ M.D @ Q, np.eye(ncv), atol=1e-7)         # M-orthonormal eigenvectors


def poisson2d_dense(n):
    T = 2 * np.eye(n) - np.eye(n, k=1) - np.eye(n, k=-1)
    return np.kron(np.eye(n), T) + np.kron(T, np.eye(n))


@pytest.mark.parametrize("kw", [dict(), dict(blksize=2, jmin=6, jmax=12), dict(blksize=2, blkwise=1, jmin=6, jmax=12),
                                dict(optype=1), dict(jmin=4, jmax=8)],
                         ids=["default", "blk2", "blk2_blockwise", "unsym_operator", "small_space"])
def test_poisson_spectrum_against_lapack(jd, kw):
    n = 10
    D = poisson2d_dense(n)
    A = NpMat(D)
    K = NpDiagPrec(np.diag(D))
    w = np.linalg.eigvalsh(D)
    kconv, lam, Q, it, it_inner = jd.jdsym(A, None, K, 4, 0.0, 1e-8, 200, np_gmres, **kw)
    assert kconv == 4
    assert np.allclose(np.sort(lam), w[:4], rtol=0, atol=1e-8)
    assert np.all(residuals(A, None, lam, Q) < 1e-8)
    # interior target: the eigenvalues closest to tau
    tau = 3.9
    kconv, lam, Q, it, _ = jd.jdsym(A, None, None, 3, tau, 1e-8, 300, np_gmres, **kw)
    assert kconv == 3 and np.all(residuals(A, None, lam, Q) < 1e-8)
    # true eigenvalues from the cluster around tau (lambda = 4 alone has multiplicity 10 here,
    # so "the three closest" is not well defined)
    assert all(np.min(np.abs(w - l)) < 1e-7 for l in lam) and np.all(np.abs(lam - tau) < 0.6)


def test_strategy_projector_and_argument_checks(jd):
    n = 60
    d = np.concatenate([np.zeros(5), np.arange(1.0, n - 4.0)])           # 5-dimensional null space
    A = NpMat(np.diag(d))
    # strategy 1: never converge to eigenvalues below tau (here: skip the null space)
    kconv, lam, Q, it, _ = jd.jdsym(A, None, None, 3, 0.5, 1e-8, 200, np_gmres, strategy=1, jmin=5, jmax=10)
    assert kconv == 3 and np.allclose(np.sort(lam), [1.0, 2.0, 3.0], atol=1e-7)

    class DropFirst:                                                      # keeps everything in {x : x[:5] = 0}
        shape = (n, n)

        def project(self, v):
            v[:5] = 0.0

    kconv, lam, Q, it, _ = jd.jdsym(A, None, None, 2, 0.0, 1e-8, 200, np_gmres, projector=DropFirst(), jmin=5, jmax=10)
    assert kconv == 2 and np.allclose(np.sort(lam), [1.0, 2.0], atol=1e-7) and np.allclose(Q[:5], 0.0, atol=1e-12)
    with pytest.raises(ValueError, match="shapes differ"):
        jd.jdsym(A, NpMat(np.eye(n + 1)), None, 1, 0.0, 1e-8, 10, np_gmres)
    with pytest.raises(ValueError, match="V0"):
        jd.jdsym(A, None, None, 1, 0.0, 1e-8, 10, np_gmres, V0=np.ones(n + 1))
    with pytest.raises(ValueError):
        jd.jdsym(A, None, None, n + 1, 0.0, 1e-8, 10, np_gmres)
    with pytest.raises(ValueError):
        jd.jdsym(A, None, None, 2, 0.0, 1e-8, 10, np_gmres, blksize=3)
    kconv, lam, Q, it, it_inner = jd.jdsym(A, None, None, 2, 0.0, 1e-8, 0, np_gmres)
    assert (kconv, it, it_inner) == (0, 0, 0) and Q.shape == (n, 0) and lam.shape == (0,)


def test_verbose_output_format(jd, capsys):
    A = NpMat(np.diag(np.arange(1.0, 21.0)))
    jd.jdsym(A, None, None, 2, 0.0, 1e-8, 50, np_gmres, clvl=1, jmin=4, jmax=8)
    out = capsys.readouterr().out
    assert "JDSYM     Solving  A*x = lambda*x  without preconditioning." in out
    assert "  IT   K   J       RES CGIT RITZVALS(1:5)" in out
    assert "JDSYM execution statistics" in out and "LAMBDA(I)" in out


def test_device_operator_module_rejects_foreign_operands_without_touching_the_gpu():
    """eigen/correq (the device-resident correction equation) only accepts device objects; jdsym's
    own operator then stays on the callback protocol (no GPU is needed to find that out)."""
    from pysparse_b200.eigen import correq, jdsym as J

    class Foreign:
        shape = (5, 5)

        def matvec(self, x, y):
            y[:] = 2.0 * x

    with pytest.raises(TypeError):
        correq.operator(2, Foreign(), None, None, 3)
    with pytest.raises(TypeError):
        correq.operator(2, object(), None, None, 3)
    ce = J.CorrEq(2, 5, Foreign(), None, None, np.zeros((5, 3), order="F"), None, None, None)
    assert ce._dev is None
    with pytest.raises(AttributeError):
        ce._psp_mat
    ce.update(0, 0.5)
    y = np.empty(5)
    ce.matvec(np.ones(5), y)
    assert np.allclose(y, 1.5)
