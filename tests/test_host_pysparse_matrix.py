"""CPU tests of the PysparseMatrix wrapper layer: the worked
examples of the reference's docstrings (pysparse/sparse/pysparseMatrix.py:167-456, rendered in
doc/pysparse/source/pysparsematrix.rst) as explicit checks, the sparse x sparse products of the
container against dense NumPy, and the reference's test/test_spmatrix.py:204-221 product
identity (the matvec legs of that test need the GPU and live in the GPU suite)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def sp(native):
    from pysparse.sparse import spmatrix
    return spmatrix


def rows(text):
    "printed matrix as a list of whitespace-normalised rows"
    return [" ".join(line.split()) for line in str(text).splitlines()]


def test_reference_docstring_examples(native):
    """The worked examples of the reference's class docstrings (pysparse/sparse/pysparseMatrix.py:
    __add__ :168-186, __mul__ :226-255, find :313-325, put :333-345, putDiagonal :372-386,
    addAt :434-441, PysparseIdentityMatrix :486-489, PysparseSpDiagsMatrix :506-513), checked
    against the printed text / values they document."""
    from pysparse.sparse import PysparseMatrix, PysparseIdentityMatrix, PysparseSpDiagsMatrix
    L = PysparseMatrix(size=3)
    L.put([3., 10., np.pi, 2.5], [0, 0, 1, 2], [2, 1, 1, 0])
    assert rows(L) == ["--- 10.000000 3.000000", "--- 3.141593 ---", "2.500000 --- ---"]
    assert rows(L + PysparseIdentityMatrix(size=3)) == ["1.000000 10.000000 3.000000", "--- 4.141593 ---", "2.500000 --- 1.000000"]
    assert rows(L + 0) == rows(L)
    assert rows(L + 3) == ["--- 13.000000 6.000000", "--- 6.141593 ---", "5.500000 --- ---"]
    val, irow, jcol = L.find()
    assert np.allclose(val, [10., 3., np.pi, 2.5]) and irow.tolist() == [0, 0, 1, 2] and jcol.tolist() == [1, 2, 1, 0]
    L.put(2 * np.pi, range(3), range(3))
    assert rows(L) == ["6.283185 10.000000 3.000000", "--- 6.283185 ---", "2.500000 --- 6.283185"]
    # products
    L1 = PysparseMatrix(size=3)
    L1.put([3., 10., np.pi, 2.5], [0, 0, 1, 2], [2, 1, 1, 0])
    L2 = PysparseMatrix(size=3)
    L2.put(np.ones(3), np.arange(3), np.arange(3))
    L2.put([4.38, 12357.2, 1.1], [2, 1, 0], [1, 0, 2])
    expect = np.array(((1.23572000e+05, 2.31400000e+01, 3.00000000e+00),
                       (3.88212887e+04, 3.14159265e+00, 0.00000000e+00),
                       (2.50000000e+00, 0.00000000e+00, 2.75000000e+00)))
    assert np.allclose((L1 * L2).getNumpyArray(), expect)
    # diagonals
    D = PysparseMatrix(size=3)
    D.putDiagonal([3., 10., np.pi])
    assert rows(D) == ["3.000000 --- ---", "--- 10.000000 ---", "--- --- 3.141593"]
    D.putDiagonal([10., 3.])
    assert rows(D) == ["10.000000 --- ---", "--- 3.000000 ---", "--- --- 3.141593"]
    D.putDiagonal(2.7182)
    assert rows(D) == ["2.718200 --- ---", "--- 2.718200 ---", "--- --- 2.718200"]
    A = PysparseMatrix(size=3)
    A.put([3., 10., np.pi, 2.5], [0, 0, 1, 2], [2, 1, 1, 0])
    A.addAt((1.73, 2.2, 8.4, 3.9, 1.23), (1, 2, 0, 0, 1), (2, 2, 0, 0, 2))
    assert rows(A) == ["12.300000 10.000000 3.000000", "--- 3.141593 2.960000", "2.500000 --- 2.200000"]
    assert rows(PysparseIdentityMatrix(size=3)) == ["1.000000 --- ---", "--- 1.000000 ---", "--- --- 1.000000"]
    e = np.ones(5)
    assert rows(PysparseSpDiagsMatrix(size=5, vals=(-2 * e, e, 2 * e), pos=(-1, 0, 1))) == [
        "1.000000 2.000000 --- --- ---", "-2.000000 1.000000 2.000000 --- ---", "--- -2.000000 1.000000 2.000000 ---",
        "--- --- -2.000000 1.000000 2.000000", "--- --- --- -2.000000 1.000000"]
    # cell layout: 11 characters per cell, '---' centred
    assert str(PysparseIdentityMatrix(size=2)).splitlines()[0] == " 1.000000  " + "---".center(11)
    big = PysparseMatrix(size=1); big[0, 0] = 123456.0
    assert str(big).strip() == "1.23e+05"


def test_wrapper_surface_and_arithmetic(native):
    from pysparse.sparse import PysparseMatrix, PysparseIdentityMatrix, PysparseSpDiagsMatrix
    from pysparse.sparse.pysparseMatrix import PysparseMatrix as P2
    assert P2 is PysparseMatrix
    L = PysparseMatrix(size=3)
    L.put([3., 10., np.pi, 2.5], [0, 0, 1, 2], [2, 1, 1, 0])
    D = np.array([[0, 10, 3], [0, np.pi, 0], [2.5, 0, 0]])
    assert L.shape == (3, 3) and L.nnz == 4 and not L.isSymmetric()
    assert np.array_equal(L.getNumpyArray(), D)
    I = PysparseIdentityMatrix(size=3)
    assert I.isSymmetric() and np.array_equal((L + I).getNumpyArray(), D + np.eye(3))
    assert np.array_equal((I + L).getNumpyArray(), D + np.eye(3))          # symmetric + general generalizes
    assert np.array_equal((L - I).getNumpyArray(), D - np.eye(3))
    assert (L + 0) is L and np.array_equal((L + 3).getNumpyArray(), np.where(D != 0, D + 3, 0))
    assert np.array_equal((L * 2).getNumpyArray(), 2 * D) and np.array_equal((2 * L).getNumpyArray(), 2 * D)
    assert np.array_equal((-L).getNumpyArray(), -D)
    M = L.copy(); M += I; M -= L
    assert np.array_equal(M.getNumpyArray(), np.eye(3))
    M *= 4
    assert np.array_equal(M.takeDiagonal(), 4 * np.ones(3))
    with pytest.raises(TypeError):
        L + PysparseMatrix(size=4)
    with pytest.raises(TypeError):
        L * PysparseMatrix(size=4)
    with pytest.raises(TypeError):
        M *= L
    with pytest.raises(ValueError):
        PysparseMatrix(size=3, nrow=4, ncol=3)
    sub = L[:2, 1:]
    assert isinstance(sub, PysparseMatrix) and np.array_equal(sub.getNumpyArray(), D[:2, 1:]) and L[0, 1] == 10.0
    L[1:, 1:] = PysparseIdentityMatrix(size=2)
    D[1:, 1:] = np.eye(2)
    assert np.array_equal(L.getNumpyArray(), D)
    L.addAtDiagonal(1.5); D += 1.5 * np.eye(3)
    L.addAtDiagonal([1.0, 2.0]); D[0, 0] += 1.0; D[1, 1] += 2.0
    assert np.array_equal(L.getNumpyArray(), D)
    e = np.ones(5)
    T = PysparseSpDiagsMatrix(size=5, vals=(-2 * e, e, 2 * e), pos=(-1, 0, 1))
    assert np.array_equal(T.getNumpyArray(), np.diag(e) + np.diag(2 * e[:4], 1) + np.diag(-2 * e[:4], -1))
    Ts = PysparseSpDiagsMatrix(size=5, vals=(-2 * e, e, 2 * e), pos=(-1, 0, 1), symmetric=True)
    assert Ts.isSymmetric() and Ts.nnz == 9 and Ts[0, 1] == -2.0
    R = PysparseMatrix(nrow=2, ncol=4)
    assert R.shape == (2, 4)


@pytest.mark.parametrize("seed", range(4))
def test_container_products_vs_dense(sp, seed):
    rng = np.random.default_rng(seed)
    n, k, m = int(rng.integers(3, 12)), int(rng.integers(3, 12)), int(rng.integers(3, 12))

    def rand(r, c, dens):
        D = rng.standard_normal((r, c)) * (rng.random((r, c)) < dens)
        A = sp.ll_mat(r, c)
        for i in range(r):
            for j in range(c):
                if D[i, j]:
                    A[i, j] = D[i, j]
        return A, D

    def dense(A):
        D = np.zeros(A.shape); v, i, j = A.find(); D[i, j] = v
        if A.issym:
            D[j, i] = v
        return D

    A, Da = rand(n, k, 0.6); B, Db = rand(k, m, 0.5)
    C = sp.matrixmultiply(A, B)
    assert C.shape == (n, m) and not C.issym and np.allclose(dense(C), Da @ Db, rtol=0, atol=1e-13)
    # bit-exact against the reference's accumulation order: terms added in ascending k
    exp = np.zeros((n, m))
    for i in range(n):
        for kk in range(k):
            if Da[i, kk]:
                for j in range(m):
                    if Db[kk, j]:
                        exp[i, j] += Da[i, kk] * Db[kk, j]
    assert np.array_equal(dense(C), exp)
    B2, Db2 = rand(n, m, 0.5)
    assert np.allclose(dense(sp.dot(A, B2)), Da.T @ Db2, rtol=0, atol=1e-13)
    S = sp.symdot(A)
    assert S.issym and S.shape == (k, k) and np.allclose(dense(S), Da.T @ Da, rtol=0, atol=1e-13)
    d = rng.standard_normal(n)
    assert np.allclose(dense(sp.symdot(A, d)), Da.T @ np.diag(d) @ Da, rtol=0, atol=1e-13)
    # symmetric x general
    Ds = rng.standard_normal((k, k)) * (rng.random((k, k)) < 0.5); Ds = np.tril(Ds) + np.tril(Ds, -1).T
    Ssym = sp.ll_mat_sym(k)
    for i in range(k):
        for j in range(i + 1):
            if Ds[i, j]:
                Ssym[i, j] = Ds[i, j]
    assert np.allclose(dense(sp.matrixmultiply(Ssym, B)), Ds @ Db, rtol=0, atol=1e-13)
    with pytest.raises(NotImplementedError):
        sp.matrixmultiply(A, Ssym)                       # general x symmetric: as the reference
    with pytest.raises(NotImplementedError):
        sp.matrixmultiply(Ssym, Ssym)
    with pytest.raises(NotImplementedError):
        sp.dot(Ssym, Ssym)
    with pytest.raises(NotImplementedError):
        sp.symdot(Ssym)
    with pytest.raises(ValueError):
        sp.matrixmultiply(A, sp.ll_mat(k + 1, m))
    with pytest.raises(ValueError):
        sp.dot(A, sp.ll_mat(n + 1, m))
    with pytest.raises(sp.error):
        sp.symdot(A, np.ones(n + 1))


def test_tools_and_solver_wrapper_surface(native, capsys):
    """tools.poisson_vec builds the same matrices as tools.poisson (reference:
    examples/poisson_test_vec.py compares them the same way); spmatrix_util helpers; the
    itsolvers_util classes and pass-through functions exist with the reference's protocol."""
    from pysparse.sparse import spmatrix
    from pysparse.tools import poisson, poisson_vec, spmatrix_util, sptime
    import pysparse.itsolvers.itsolvers_util as iu
    from pysparse.itsolvers import krylov

    def same(a, b):
        z = a.copy(); z.shift(-1.0, b)
        return z.nnz == 0 and a.shape == b.shape and a.issym == b.issym and a.nnz == b.nnz

    n = 7
    assert same(poisson.poisson1d(n), poisson_vec.poisson1d_vec(n))
    assert same(poisson.poisson1d_sym(n), poisson_vec.poisson1d_sym_vec(n))
    assert same(poisson.poisson2d(n), poisson_vec.poisson2d_vec(n)) and same(poisson.poisson2d(n), poisson_vec.poisson2d_vec2(n))
    assert same(poisson.poisson2d_sym(n), poisson_vec.poisson2d_sym_vec(n))
    assert same(poisson.poisson2d_sym_blk(n), poisson_vec.poisson2d_sym_blk_vec(n))
    assert same(poisson.poisson2d_sym(n), poisson_vec.poisson2d_sym_blk_vec(n))
    assert spmatrix_util.bytesToString(100) == "100 Bytes" and spmatrix_util.bytesToString(2048) == "2.0 Kbytes"
    assert spmatrix_util.bytesToString(3 * 1024 ** 2) == "3.0 Mbytes" and spmatrix_util.bytesToString(5 * 1024 ** 3) == "5.0 Gbytes"
    L = poisson.poisson2d_sym(n)
    spmatrix_util.printInfo(L, "L"); spmatrix_util.printInfo(L.to_sss(), "S"); spmatrix_util.printInfo(L.to_csr(), "A")
    out = capsys.readouterr().out
    assert "LL symmetric" in out and "SSS" in out and "CSR" in out and "%dx%d" % (n * n, n * n) in out
    # storage formulas of SURVEY section 8: CSR = 4 N + 12 nnz, SSS = 12 N + 12 nnz_lower
    A = L.to_csr(); S = L.to_sss()
    assert spmatrix_util.storage_bytes(A) == 4 * n * n + 12 * A.nnz
    assert spmatrix_util.storage_bytes(S) == 12 * n * n + 12 * (S.nnz - n * n)
    R = spmatrix_util.ll_mat_rand(10, 12, 0.2)
    assert R.shape == (10, 12) and 0 < R.nnz <= 24
    assert sptime.cputime() >= 0.0
    for cls, fn in ((iu.Pcg, krylov.pcg), (iu.Minres, krylov.minres), (iu.Qmrs, krylov.qmrs), (iu.Cgs, krylov.cgs),
                    (iu.Bicgstab, krylov.bicgstab), (iu.Gmres, krylov.gmres)):
        s = cls(L)
        assert s.itsolver is fn and s.A is L and s.nofCalled == 0 and not s.debug and callable(s)
    with pytest.raises(NotImplementedError):
        iu.ItSolver(L, False).solve(np.ones(n * n), np.zeros(n * n), 1e-8, 10)
    for name in ("pcg", "minres", "qmrs", "cgs", "bicgstab", "gmres"):
        assert callable(getattr(iu, name))
