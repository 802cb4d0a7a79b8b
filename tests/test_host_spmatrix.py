"""CPU tests of the host layer (C extension modules): ll_mat container semantics,
bit-exact structure conversion, argument checking and error behaviour.  The
assertions on ll_mat follow the reference's test/test_spmatrix.py where it has
any (:17-65 create/entry/negative index, :77-78 Poisson nnz counts, :84-120
sub-matrices, :180-202 norms)."""
import math
import os

import numpy as np
import pytest

from pysparse_b200.tools import gallery


@pytest.fixture(scope="module")
def mods(native):
    from pysparse.sparse import spmatrix
    from pysparse.itsolvers import krylov
    from pysparse.precon import precon
    from pysparse.tools import poisson
    return spmatrix, krylov, precon, poisson


def test_import_paths_and_module_surface(mods):
    spmatrix, krylov, precon, _ = mods
    for name in ("ll_mat", "ll_mat_sym", "ll_mat_from_mtx", "matrixmultiply", "dot", "symdot",
                 "LLMatType", "CSRMatType", "SSSMatType", "error", "_C_API"):
        assert hasattr(spmatrix, name), name
    for name in ("pcg", "minres", "qmrs", "bicgstab", "cgs", "gmres"):
        assert callable(getattr(krylov, name))
    for name in ("jacobi", "ssor", "JacobiType", "SSORType"):
        assert hasattr(precon, name)
    import pysparse.itsolvers as its
    assert its.Pcg and its.Gmres


def test_llmat_create_entry_negative_index(mods):
    spmatrix = mods[0]
    A = spmatrix.ll_mat(5, 7)
    assert A.shape == (5, 7) and A.nnz == 0 and A.issym == 0
    A[1, 2] = 3.5
    A[4, 6] = -1.0
    assert A[1, 2] == 3.5 and A[-1, -1] == -1.0 and A[0, 0] == 0.0 and A.nnz == 2
    A[1, 2] = 0.0            # zero removes the entry unless storeZeros
    assert A.nnz == 1
    with pytest.raises(IndexError):
        A[5, 0]
    with pytest.raises(IndexError):
        A[0, 7] = 1.0
    Z = spmatrix.ll_mat(3, 3, 10, 1)
    Z[0, 0] = 0.0
    assert Z.nnz == 1
    S = spmatrix.ll_mat_sym(4)
    S[2, 1] = 7.0
    assert S[1, 2] == 7.0 and S.issym == 1
    with pytest.raises(IndexError, match="upper triangle"):
        S[1, 2] = 1.0


def test_poisson_nnz_counts(mods):
    """test/test_spmatrix.py:77-78"""
    poisson = mods[3]
    n = 20
    assert poisson.poisson2d_sym(n).nnz == n * (3 * n - 2)
    assert poisson.poisson2d(n).nnz == n * (5 * n - 4)
    assert poisson.poisson2d_sym_blk(n).nnz == n * (3 * n - 2)
    assert poisson.poisson2d_vec(n).nnz == n * (5 * n - 4)
    assert poisson.poisson2d_sym_vec(n).nnz == n * (3 * n - 2)


def test_submatrix_get_set_and_bulk_ops(mods):
    spmatrix = mods[0]
    A = spmatrix.ll_mat(6, 6)
    for i in range(6):
        for j in range(6):
            if (i + 2 * j) % 3 == 0:
                A[i, j] = 10 * i + j + 1
    B = A[1:4, 2:6]
    assert B.shape == (3, 4)
    for i in range(3):
        for j in range(4):
            assert B[i, j] == A[1 + i, 2 + j]
    C = spmatrix.ll_mat(6, 6)
    C[1:4, 2:6] = B
    for i in range(6):
        for j in range(6):
            assert C[i, j] == (A[i, j] if 1 <= i < 4 and 2 <= j < 6 else 0.0)
    val, irow, jcol = A.find()
    assert len(val) == A.nnz and all(A[int(i), int(j)] == v for v, i, j in zip(val, irow, jcol))
    assert sorted(A.keys()) == sorted(zip(irow.tolist(), jcol.tolist()))
    D = spmatrix.ll_mat(6, 6)
    D.put(val, irow, jcol)
    out = np.zeros(len(val))
    D.take(out, irow, jcol)
    assert np.array_equal(out, val)
    D.update_add_at(val, irow, jcol)
    D.take(out, irow, jcol)
    assert np.array_equal(out, 2 * val)
    E = A.copy(); E.scale(2.0); E.shift(-2.0, A)
    assert E.nnz == 0
    dense = np.array([[A[i, j] for j in range(6)] for i in range(6)])
    assert math.isclose(A.norm("fro"), np.linalg.norm(dense))
    assert A.norm("1") == np.abs(dense).sum(0).max() and A.norm("inf") == np.abs(dense).sum(1).max()


def _dict_model(L):
    val, ir, jc = L.find()
    return {(int(i), int(j)): float(v) for v, i, j in zip(val, ir, jc)}


@pytest.mark.parametrize("n", [1, 2, 7, 23])
def test_conversions_bit_exact_vs_gallery(mods, n):
    """to_csr / to_sss must reproduce ll_mat.c:1577-1708 exactly: ascending columns,
    symmetric expansion order, strict lower + dense diagonal."""
    spmatrix, _, _, poisson = mods
    g = gallery.poisson2d(n)
    Ls, Lg = poisson.poisson2d_sym(n), poisson.poisson2d(n)
    for L in (Ls, Lg):
        v, c, i = L.to_csr().arrays()
        assert np.array_equal(v, g["csr"][0]) and np.array_equal(c, g["csr"][1]) and np.array_equal(i, g["csr"][2])
        v, d, c, i = L.to_sss().arrays()
        assert np.array_equal(v, g["sss"][0]) and np.array_equal(d, g["sss"][1])
        assert np.array_equal(c, g["sss"][2]) and np.array_equal(i, g["sss"][3])
    A, S = Ls.to_csr(), Ls.to_sss()
    assert A.nnz == n * (5 * n - 4) and S.nnz == n * (3 * n - 2) and A.shape == S.shape == (n * n, n * n)


def test_conversions_random_vs_python_model(mods):
    spmatrix = mods[0]
    rng = np.random.default_rng(5)
    n = 60
    L = spmatrix.ll_mat_sym(n)
    G = spmatrix.ll_mat(n, n + 3)
    for _ in range(400):
        i, j = int(rng.integers(n)), int(rng.integers(n))
        v = float(rng.standard_normal())
        L[max(i, j), min(i, j)] = v
        G[i, int(rng.integers(n + 3))] = v
    # general: rows in ascending column order
    m = _dict_model(G)
    v, c, ind = G.to_csr().arrays()
    exp = sorted(m.items())
    assert [(int(r), int(cc)) for r in range(n) for cc in c[ind[r]:ind[r + 1]]] == [k for k, _ in exp]
    assert np.array_equal(v, np.array([x for _, x in exp]))
    # symmetric: full expansion, ascending columns; sss: strict lower + diag
    m = _dict_model(L)
    full = dict(m)
    full.update({(j, i): x for (i, j), x in m.items()})
    v, c, ind = L.to_csr().arrays()
    exp = sorted(full.items())
    assert [(int(r), int(cc)) for r in range(n) for cc in c[ind[r]:ind[r + 1]]] == [k for k, _ in exp]
    assert np.array_equal(v, np.array([x for _, x in exp]))
    sv, sd, sc, sind = L.to_sss().arrays()
    low = sorted((k, x) for k, x in m.items() if k[0] > k[1])
    assert [(int(r), int(cc)) for r in range(n) for cc in sc[sind[r]:sind[r + 1]]] == [k for k, _ in low]
    assert np.array_equal(sv, np.array([x for _, x in low]))
    assert np.array_equal(sd, np.array([m.get((i, i), 0.0) for i in range(n)]))
    # a general matrix silently loses its upper triangle in to_sss (ll_mat.c:1692-1705)
    Gs = spmatrix.ll_mat(3, 3); Gs[0, 2] = 5.0; Gs[2, 0] = 4.0; Gs[1, 1] = 3.0
    sv, sd, sc, sind = Gs.to_sss().arrays()
    assert sv.tolist() == [4.0] and sd.tolist() == [0.0, 3.0, 0.0]
    with pytest.raises(ValueError, match="square"):
        G.to_sss()
    # gallery's vectorised expansion agrees with the ll_mat route
    gv, gc, gi = gallery.sss_to_csr(n, *[np.array(a) for a in L.to_sss().arrays()])
    fv, fc, fi = L.to_csr().arrays()
    keep = gv != 0.0
    assert set(zip(np.repeat(np.arange(n), np.diff(gi))[keep].tolist(), gc[keep].tolist())) == set(full)


def test_mtx_roundtrip(mods, tmp_path):
    spmatrix, _, _, poisson = mods
    L = poisson.poisson2d_sym(6)
    f = str(tmp_path / "p.mtx")
    L.export_mtx(f)
    M = spmatrix.ll_mat_from_mtx(f)
    assert M.issym == 1 and _dict_model(M) == _dict_model(L)


def test_argument_checking_matches_reference_messages(mods):
    spmatrix, krylov, precon, poisson = mods
    L = poisson.poisson2d_sym(4)
    A, S = L.to_csr(), L.to_sss()
    x = np.ones(16)
    for M in (A, S, L):
        with pytest.raises(ValueError, match="arg 1 must be a 1-dimensional double array of appropriate size"):
            M.matvec(np.ones(15), x)
        with pytest.raises(ValueError, match="arg 2 must be a 1-dimensional double array of appropriate size"):
            M.matvec(x, np.ones(16, dtype=np.float32))
        with pytest.raises(TypeError):
            M.matvec([1.0] * 16, x)
    with pytest.raises(TypeError):
        precon.ssor(A)
    with pytest.raises(TypeError):
        precon.ssor(L)
    with pytest.raises(ValueError, match="incompatible operand shapes"):
        krylov.pcg(S, np.ones(16), np.zeros(15), 1e-8, 10)
    with pytest.raises(ValueError, match="matrix is not square"):
        krylov.pcg(spmatrix.ll_mat(3, 4), np.ones(3), np.zeros(3), 1e-8, 10)
    with pytest.raises(ValueError, match="invalid matrix shape"):
        krylov.minres(spmatrix.ll_mat(3, 4), np.ones(3), np.zeros(3), 1e-8, 10)
    with pytest.raises(TypeError):
        krylov.pcg(S, np.ones(16), np.zeros(16), 1e-8)          # positional-only signature "OOOdi|O"
    Z = spmatrix.ll_mat_sym(3); Z[0, 0] = 1.0; Z[2, 2] = 2.0
    with pytest.raises(ValueError, match="diagonal element close to zero"):
        precon.jacobi(Z)
    with pytest.raises(NotImplementedError):
        spmatrix.matrixmultiply(L, L)
    assert S[2, 1] == S[1, 2] == L[2, 1] and A[0, 0] == 4.0 and S.nnz == L.nnz


def test_no_gpu_means_loud_failure_not_cpu_fallback(mods, native):
    if native.device_count() > 0:
        pytest.skip("a GPU is present")
    spmatrix, krylov, precon, poisson = mods
    L = poisson.poisson2d_sym(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.to_sss().matvec(np.ones(16), np.zeros(16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        krylov.pcg(L.to_csr(), np.ones(16), np.zeros(16), 1e-8, 10)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        precon.jacobi(L)


def test_from_arrays_accepts_wide_indices_validates_and_copies():
    """csr_from_arrays / sss_from_arrays: NumPy's default int64 index arrays are accepted after a
    range check, malformed structure is refused at construction, and the arrays are copied."""
    from pysparse_b200.sparse import spmatrix
    ind = np.array([0, 2, 3], dtype=np.int64)
    col = np.array([0, 1, 1], dtype=np.int64)
    val = np.array([1.0, 2.0, 3.0])
    A = spmatrix.csr_from_arrays(2, 2, ind, col, val)
    val[0] = 99.0; col[0] = 1                      # later edits of the caller's arrays do not reach A
    assert A[0, 0] == 1.0 and A[0, 1] == 2.0 and A[1, 1] == 3.0 and A.nnz == 3
    with pytest.raises(ValueError):
        spmatrix.csr_from_arrays(2, 2, np.array([0, 2, 3]), np.array([0, 2**31 + 1, 1]), val)
    with pytest.raises(ValueError):
        spmatrix.csr_from_arrays(2, 2, np.array([0, 2, 3]), np.array([0, 5, 1]), val)       # column out of range
    with pytest.raises(ValueError):
        spmatrix.csr_from_arrays(2, 2, np.array([0, 3, 2]), np.array([0, 1, 1]), val)       # not monotone
    with pytest.raises(ValueError):
        spmatrix.sss_from_arrays(2, np.array([0, 0, 1]), np.array([1]), np.array([1.0]), np.ones(2))   # not strictly lower
    S = spmatrix.sss_from_arrays(2, np.array([0, 0, 1]), np.array([0]), np.array([-1.0]), np.full(2, 4.0))
    assert S[1, 0] == -1.0 and S[0, 1] == -1.0 and S[1, 1] == 4.0


def test_index_arrays_beyond_int32_raise_instead_of_wrapping():
    from pysparse_b200.sparse import spmatrix
    A = spmatrix.ll_mat(3, 3)
    A[1, 0] = 5.0
    out = np.zeros(1)
    with pytest.raises(IndexError):
        A.take(out, np.array([2**32 + 1]), np.array([0]))
    with pytest.raises(IndexError):
        A.put(np.array([1.0]), np.array([2**32 + 1]), np.array([0]))
    with pytest.raises(IndexError):
        A.update_add_at(np.array([1.0]), np.array([0]), np.array([2**32 + 2]))
    assert A[1, 0] == 5.0 and A.nnz == 1
