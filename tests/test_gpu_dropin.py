"""GPU tests of the drop-in Python API: the reference's example scripts' call
patterns (examples/poisson_test.py, ssor_test.py, poisson_gmres.py, tendigit.py,
itsolvers_util.py __main__) run unchanged on pysparse_b200 and agree with the oracle."""
import numpy as np
import pytest

from cases import build_matrix
from helpers import load_kat

pytestmark = pytest.mark.gpu
KAT = {r["id"]: r for r in load_kat()}


@pytest.fixture(scope="module")
def mods(native):
    if native.device_count() == 0:
        pytest.fail("no CUDA device")
    from pysparse.sparse import spmatrix
    from pysparse.itsolvers import krylov
    from pysparse.precon import precon
    from pysparse.tools import poisson
    return spmatrix, krylov, precon, poisson


def test_poisson_test_example_flow(mods, checker):
    """examples/poisson_test.py:62-135: SSS / CSR / LL matrices and SSOR give the same solve."""
    spmatrix, krylov, precon, poisson = mods
    n, tol = 100, 1e-8
    L = poisson.poisson2d_sym(n)
    A, S = L.to_csr(), L.to_sss()
    b = np.ones(n * n)
    g = KAT["p2d100_sss_pcg"]
    sols = []
    for M in (S, A, L):
        x = np.zeros(n * n)
        info, it, relres = krylov.pcg(M, b, x, tol, 2000)
        assert (info, it) == (g["info"], g["iter"]) and abs(relres - g["relres"]) <= 1e-10
        r = np.empty(n * n); M.matvec(x, r)
        assert np.linalg.norm(b - r) <= 1.01 * tol * np.linalg.norm(b)
        sols.append(x)
    assert np.array_equal(sols[0], sols[1]) and np.array_equal(sols[0], sols[2])
    x = np.zeros(n * n)
    info, it, relres = krylov.pcg(S, b, x, tol, 2000, precon.ssor(S, 1.9))
    g = KAT["p2d100_sss_pcg_ssor19"]
    assert (info, it) == (g["info"], g["iter"]) and abs(relres - g["relres"]) <= 1e-10
    K = precon.jacobi(L)              # the documented idiom: jacobi on the ll_mat, solve on the sss_mat
    assert K.shape == (n * n, n * n)
    x = np.zeros(n * n)
    assert krylov.pcg(S, b, x, tol, 2000, K)[:2] == (0, 187)


def test_config1_pcg_jacobi_n300(mods):
    """BASELINE.json configs[0]: 2-D Poisson n=300 sss_mat, pcg + jacobi, tol 1e-10."""
    spmatrix, krylov, precon, poisson = mods
    g = build_matrix(["poisson2d", 300])
    S = spmatrix.sss_from_arrays(g["n"], g["sss"][3], g["sss"][2], g["sss"][0], g["sss"][1])
    x = np.zeros(g["n"])
    info, it, relres = krylov.pcg(S, np.ones(g["n"]), x, 1e-10, 2000, precon.jacobi(S))
    k = KAT["p2d300_sss_pcg_jacobi"]
    assert info == 0 and abs(it - k["iter"]) <= 0.02 * k["iter"] and abs(relres - k["relres"]) <= 1e-10
    assert abs(np.linalg.norm(x) - k["xnorm"]) <= 1e-8 * k["xnorm"]


def test_all_solvers_oo_wrappers(mods):
    """pysparse/itsolvers/itsolvers_util.py:154-178 (__main__ smoke) with asserts"""
    spmatrix, krylov, precon, poisson = mods
    import pysparse.itsolvers as its
    L = poisson.poisson2d_sym(50)
    A, S = L.to_csr(), L.to_sss()
    n = 2500
    b = np.ones(n)
    for cls in (its.Pcg, its.Minres, its.Qmrs, its.Cgs, its.Bicgstab, its.Gmres):
        for K in (None, precon.ssor(S), precon.jacobi(L, 1.0, 1)):
            s = cls(A, precon=K)
            x = np.full(n, 123.0)           # solve() zeroes x first (itsolvers_util.py:41)
            info, it, relres = s.solve(b, x, 1e-8, 1000)
            assert info == 0 and s.converged and s.itn == it and s.nofCalled == 1
            r = np.empty(n); A.matvec(x, r)
            assert np.linalg.norm(b - r) <= 2e-7 * np.linalg.norm(b), (cls.name, K)
    # info < 0 raises, with (message, info, iter, relres) as arguments (itsolvers_util.py:47-48)
    from pysparse.sparse import PysparseMatrix
    s = its.Pcg(PysparseMatrix(matrix=L))          # a PysparseMatrix is unwrapped (:19-22)
    with pytest.raises(RuntimeError) as ei:
        s.solve(b, np.zeros(n), 1e-12, 5)
    assert ei.value.args[1] == -1 and s.lastInfo == -1 and s.nofCalled == 1 and not s.converged
    assert "nofCalled: 1" in str(s)


def test_x_updated_in_place_only_when_contiguous_float64(mods):
    """itsolversmodule.c:70-82 (PyArray_AsCArray): a non-contiguous x is copied and the result lost."""
    spmatrix, krylov, precon, poisson = mods
    S = poisson.poisson2d_sym(10).to_sss()
    b = np.ones(100)
    x = np.zeros(100)
    assert krylov.pcg(S, b, x, 1e-8, 100)[0] == 0 and np.any(x != 0)
    xs = np.zeros(200)[::2]
    assert krylov.pcg(S, b, xs, 1e-8, 100)[0] == 0 and not np.any(xs != 0)
    assert krylov.pcg(S, [1.0] * 100, x, 1e-8, 100)[0] == 0       # b may be any sequence


def test_duck_typed_matrix_and_preconditioner(mods):
    """doc/pysparse/source/precon.rst:15-24: any object with shape + matvec / precon"""
    spmatrix, krylov, precon, poisson = mods
    L = poisson.poisson2d_sym(20)
    S = L.to_sss()

    class Shifted:
        shape = S.shape

        def matvec(self, x, y):
            S.matvec(x, y)
            y += 0.5 * x

    class DiagPrec:
        shape = S.shape

        def precon(self, x, y):
            y[:] = x / 4.5

    x = np.zeros(400)
    info, it, relres = krylov.pcg(Shifted(), np.ones(400), x, 1e-9, 500, DiagPrec())
    assert info == 0
    r = np.empty(400); Shifted().matvec(x, r)
    assert np.linalg.norm(1.0 - r) <= 1e-8 * 20

    class Bad:
        shape = S.shape

        def matvec(self, x, y):
            raise KeyError("nope")

    with pytest.raises(KeyError):
        krylov.qmrs(Bad(), np.ones(400), np.zeros(400), 1e-8, 10)


def test_precon_objects(mods, checker):
    spmatrix, krylov, precon, poisson = mods
    g = build_matrix(["random_sym", 1500, 8, 2])
    n = g["n"]
    S = spmatrix.sss_from_arrays(n, g["sss"][3], g["sss"][2], g["sss"][0], g["sss"][1])
    So = checker.sss(n, *g["sss"])
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n)
    for om, st in ((1.0, 1), (1.4, 2)):
        K = precon.ssor(S, om, st)
        y = np.empty(n); yo = np.zeros(n)
        K.precon(x, y); checker.precon(checker.ssor(So, om, st), x, yo)
        assert np.array_equal(y, yo) and K.shape == (n, n) and K.levels[0] >= 1
        with pytest.raises(ValueError, match="contiguous"):
            K.precon(x[::2], y)
    J = precon.jacobi(S, 0.8, 3)
    y = np.empty(n); yo = np.zeros(n)
    J.precon(x, y); checker.precon(checker.jacobi(So, g["sss"][1], 0.8, 3), x, yo)
    assert np.max(np.abs(y - yo)) <= 1e-13 * np.max(np.abs(yo))
    assert np.array_equal(x, x)      # x unchanged


def test_tendigit_known_answer(mods):
    """examples/tendigit.py:26-49 -- Trefethen's 100-digit challenge, problem 7:
    x[0] of A x = e_0 is 0.7250783462684011..."""
    spmatrix, krylov, precon, poisson = mods
    g = build_matrix(["tendigit", 20000])
    n = g["n"]
    S = spmatrix.sss_from_arrays(n, g["sss"][3], g["sss"][2], g["sss"][0], g["sss"][1])
    b = np.zeros(n); b[0] = 1.0
    x = np.zeros(n)
    info, it, relres = krylov.minres(S, b, x, 1e-16, n, precon.ssor(S))
    assert abs(x[0] - 0.7250783462684011) <= 1e-14


def test_llmat_device_cache_tracks_modifications(mods):
    spmatrix, krylov, precon, poisson = mods
    L = poisson.poisson2d(6)
    x = np.arange(36, dtype=float); y1 = np.empty(36); y2 = np.empty(36)
    L.matvec(x, y1)
    L[0, 0] = 10.0
    L.matvec(x, y2)
    assert y2[0] == y1[0] + 6.0 * x[0] and np.array_equal(y1[1:], y2[1:])
    yt = np.empty(36); L.matvec_transp(x, yt)
    assert np.allclose(yt, y2)


def _dense(A):
    D = np.zeros(A.shape)
    v, i, j = A.find()
    D[i, j] = v
    if A.issym:
        D[j, i] = v
    return D


def test_edited_llmat_feeds_the_gpu_path(mods):
    """Matrices produced by the assembly operations of SURVEY 8(f) -- fancy indexing, block
    assignment, masks, scaling, row / column deletion -- multiply on the GPU like any other:
    y = A x against the dense model (relative max-norm 1e-13), after EVERY modification (the
    device copy must follow the container)."""
    spmatrix, krylov, precon, poisson = mods
    rng = np.random.default_rng(4)
    A = poisson.poisson2d(8)                                  # 64 x 64 general

    def check(M):
        D = _dense(M)
        m, n = M.shape
        x = rng.standard_normal(n); y = np.full(m, np.nan)
        M.matvec(x, y)
        assert np.max(np.abs(y - D @ x)) <= 1e-13 * max(1.0, np.max(np.abs(D @ x)))
        if not M.issym:
            xt = rng.standard_normal(m); yt = np.full(n, np.nan)
            M.matvec_transp(xt, yt)
            assert np.max(np.abs(yt - D.T @ xt)) <= 1e-13 * max(1.0, np.max(np.abs(D.T @ xt)))

    check(A)
    B = A[::2, 5:40]                                          # 32 x 35 rectangular sub-matrix
    check(B)
    B.col_scale(rng.standard_normal(35)); check(B)
    B.row_scale(rng.standard_normal(32)); check(B)
    mask = np.ones(35, dtype=int); mask[3:9] = 0
    B.delete_cols(mask); check(B)
    mask = np.ones(32, dtype=int); mask[::5] = 0
    B.delete_rows(mask); check(B)
    A[8:16, 8:16] = A[0:8, 0:8]; check(A)
    A[[1, 3, 5], :] = 0.0; check(A)
    A.update_add_mask(rng.standard_normal((3, 3)), [0, 9, 63], [2, 9, 40], [1, 1, 1], [1, 0, 1]); check(A)
    S = poisson.poisson2d_sym(8)
    check(S)
    keep = np.ones(64, dtype=int); keep[10:20] = 0
    S.delete_rowcols(keep); check(S)
    El = rng.standard_normal((3, 3)); El = El + El.T
    S.update_add_mask_sym(El, [0, 7, 30], [1, 1, 1]); check(S)
    check(S[:, :])                                            # general copy of the symmetric matrix


def test_matrixmultiply_identity_of_the_reference_test(mods):
    """test/test_spmatrix.py:204-221: (A B) x == A (B x) within eps * n * m * k."""
    spmatrix = mods[0]
    import random
    rnd = random.Random(1)
    eps = 2.2204460492503131e-16
    n, m, k = 30, 60, 30

    def ll_mat_rand(r, c, density):
        M = spmatrix.ll_mat(r, c, int(density * r * c))
        for _ in range(int(density * r * c)):
            M[rnd.randrange(r), rnd.randrange(c)] = rnd.random()
        return M

    rng = np.random.default_rng(2)
    for _ in range(10):
        A = ll_mat_rand(n, k, 0.9); B = ll_mat_rand(k, m, 0.4)
        C = spmatrix.matrixmultiply(A, B)
        t = np.zeros(k); y1 = np.zeros(n); y2 = np.zeros(n)
        for _ in range(3):
            x = rng.random(m)
            C.matvec(x, y1); B.matvec(x, t); A.matvec(t, y2)
            assert np.sqrt(np.dot(y1 - y2, y1 - y2)) < eps * n * m * k


def test_pysparse_matrix_operators(mods):
    """PysparseMatrix: A * x, x * A (numpy defers to __rmul__), (A * B) * x, solver wrappers."""
    from pysparse.sparse import PysparseMatrix, PysparseIdentityMatrix
    from pysparse.itsolvers import krylov
    L1 = PysparseMatrix(size=3)
    L1.put([3., 10., np.pi, 2.5], [0, 0, 1, 2], [2, 1, 1, 0])
    assert np.allclose(L1 * np.array((1, 2, 3), "d"), np.array((29., 6.28318531, 2.5)))      # pysparseMatrix.py:246-248
    assert np.allclose(np.array((1, 2, 3), "d") * L1, np.array((7.5, 16.28318531, 3.)))      # :252-254
    assert np.allclose(L1.matvec(np.ones(3)), L1.getNumpyArray() @ np.ones(3))
    n = 12
    P = PysparseMatrix(matrix=mods[3].poisson2d_sym(n))
    x = np.random.default_rng(0).standard_normal(n * n)
    y = P * x
    y2 = (P + PysparseIdentityMatrix(size=n * n)) * x
    assert np.allclose(y2, y + x, rtol=0, atol=1e-13)
    assert np.allclose((2 * P) * x, 2 * y, rtol=0, atol=1e-12)
    xs = np.zeros(n * n)
    info, it, relres = krylov.pcg(P.getMatrix().to_sss(), y, xs, 1e-12, 1000)
    assert info == 0 and np.max(np.abs(xs - x)) <= 1e-8


def test_jdsym_reference_cases_on_the_gpu_path(mods):
    """test/test_jdsym.py:10-16 and examples/jdsym_test.py:36-117 with the real pieces: sss_mat
    products, krylov.qmrs as the inner solver and the correction equation passed to it as a
    duck-typed operator (whose matvec calls back into the GPU products)."""
    spmatrix, krylov, precon, poisson = mods
    from pysparse.eigen import jdsym
    a = spmatrix.ll_mat(3, 3)
    a.put([1, 2, 3])
    r = jdsym.jdsym(a, None, None, 3, 1, 1e-9, 100, krylov.qmrs)
    assert r[0] == 3 and np.allclose(r[1], [1, 2, 3]) and np.allclose(np.abs(r[2]), np.identity(3))

    n, ncv, tol = 1000, 5, 1e-6
    A = spmatrix.ll_mat_sym(n); M = spmatrix.ll_mat_sym(n)
    d = np.arange(n)
    A.put(d + 1.0, d, d); M.put(float(n // 2) + d, d, d)
    As, Ms = A.to_sss(), M.to_sss()
    normM = M[n - 1, n - 1]

    class DiagPrecShifted:                       # examples/jdsym_test.py:7-15
        shape = (n, n)
        dinv = 1.0 / ((d + 1.0) - 0.006 * (float(n // 2) + d))

        def precon(self, x, y):
            np.multiply(x, self.dinv, y)

    def residuals(Aop, Mop, lam, Q):
        out = np.zeros(len(lam))
        for k in range(len(lam)):
            u = Q[:, k].copy(); rr = np.zeros(n); t = np.zeros(n)
            Aop.matvec(u, rr)
            if Mop is not None:
                Mop.matvec(u, t)
            else:
                t = u
            out[k] = np.linalg.norm(rr - lam[k] * t)
        return out

    kw = dict(jmin=5, jmax=10, eps_tr=1e-4)
    kconv, lam, Q, it, it_inner = jdsym.jdsym(As, None, None, ncv, 0.0, tol, 150, krylov.qmrs, **kw)
    assert kconv == ncv and np.all(residuals(As, None, lam, Q) <= tol)
    assert np.allclose(lam, d[:ncv] + 1.0, rtol=tol * tol, atol=1e-9)
    exact = ((d + 1.0) / (float(n // 2) + d))[:ncv]
    for K in (None, DiagPrecShifted()):
        kconv, lam, Q, it, it_inner = jdsym.jdsym(As, Ms, K, ncv, 0.0, tol, 150, krylov.qmrs, **kw)
        assert kconv == ncv and np.all(residuals(As, Ms, lam, Q) <= normM * tol)
        assert np.allclose(np.sort(lam), exact, rtol=0, atol=normM * tol * tol + 1e-9)
    # a native preconditioner object and a 2-D Poisson spectrum
    L = poisson.poisson2d_sym(12)
    S = L.to_sss()
    kconv, lam, Q, it, _ = jdsym.jdsym(S, None, precon.jacobi(L), 4, 0.0, 1e-8, 200, krylov.qmrs)
    T = 2 * np.eye(12) - np.eye(12, k=1) - np.eye(12, k=-1)
    w = np.linalg.eigvalsh(np.kron(np.eye(12), T) + np.kron(T, np.eye(12)))
    # (the second and third eigenvalue coincide: which copy of a multiple eigenvalue a Krylov
    # method sees beyond the first is decided by rounding, so only membership is asserted)
    assert kconv == 4 and abs(np.min(lam) - w[0]) <= 1e-7
    assert all(np.min(np.abs(w - v)) <= 1e-7 for v in lam)
    for j in range(4):
        q = np.ascontiguousarray(Q[:, j]); aq = np.empty(144); S.matvec(q, aq)
        assert np.linalg.norm(aq - lam[j] * q) <= 1e-7


def test_inner_outer_iteration_through_a_python_preconditioner(mods):
    """A preconditioner object whose precon() itself calls krylov.* (inner-outer iteration): the
    reference's solvers are re-entrant C functions (itsolversmodule.c:32-117 allocate per call);
    here the nested solve must get its own device state and leave the outer solve untouched."""
    spmatrix, krylov, precon, poisson = mods
    L = poisson.poisson2d_sym_vec(40)
    S = L.to_sss()
    n = 1600
    b = np.ones(n)

    class InnerSolve:
        shape = (n, n)
        calls = 0

        def precon(self, x, y):
            y[:] = 0.0
            info, it, rr = krylov.pcg(S, x.copy(), y, 1e-3, 50)      # nested solve on the same device context
            assert info in (0, -1)
            InnerSolve.calls += 1

    x = np.zeros(n)
    info, it, relres = krylov.gmres(S, b, x, 1e-9, 200, InnerSolve(), 20)
    assert info == 0 and InnerSolve.calls >= it >= 1
    r = np.empty(n); S.matvec(x, r)
    assert np.linalg.norm(b - r) <= 1e-8 * np.linalg.norm(b)
    # and a plain solve afterwards is unaffected by the frames that were swapped in and out
    x2 = np.zeros(n)
    k = krylov.pcg(S, b, x2, 1e-10, 2000)
    x3 = np.zeros(n)
    assert krylov.pcg(S, b, x3, 1e-10, 2000) == k and np.array_equal(x2, x3)


def test_multistep_jacobi_follows_an_edited_llmat(mods, checker):
    """precon.jacobi(A, omega, steps > 1) multiplies with A at every application
    (preconmodule.c:45-52).  Editing the ll_mat afterwards rebuilds its device copy; the
    preconditioner must use the new copy, not a dangling handle to the old one."""
    spmatrix, krylov, precon, poisson = mods
    n = 30
    L = poisson.poisson2d_vec(n)
    K = precon.jacobi(L, 0.9, 3)                  # dinv is frozen here (newJacobiObject, :389-401)
    x = np.linspace(-1.0, 1.0, n * n)
    y0 = np.empty(n * n); K.precon(x, y0)
    L[5, 5] = 7.5                                 # invalidates the cached device matrix
    L[0, n * n - 1] = -0.25
    y1 = np.empty(n * n); K.precon(x, y1)
    # model: damped Jacobi sweeps with the OLD diagonal and the NEW matrix
    dinv = 0.9 / 4.0
    A = L.to_csr()
    ym = x * dinv
    t = np.empty(n * n)
    for _ in range(2):
        A.matvec(ym, t)
        ym = (x - t) * dinv + ym
    assert np.max(np.abs(y1 - ym)) <= 1e-13 * np.max(np.abs(ym))
    assert not np.array_equal(y0, y1)
    xs = np.zeros(n * n)
    info, it, rr = krylov.pcg(L, np.ones(n * n), xs, 1e-8, 500, K)     # krylov re-reads K's handle too
    assert info in (0, -1)
