"""CPU tests of the complete ``ll_mat`` assembly container:
fancy indexing, element-matrix assembly, masks, scaling, row / column deletion, printing.

Three sources of truth:
  * the reference's own test file, test/test_spmatrix.py (create / entry :17-65, Poisson
    counts and sub-matrices :69-123, delete_rowcols and compress :125-178, norms :180-202),
    ported assertion by assertion;
  * the worked examples of the reference's documentation
    (doc/pysparse/source/spmatrix.rst:106-250), compared as printed text;
  * a dense NumPy model of every operation on random inputs (structure work: exact).
"""
import random

import numpy as np
import pytest


@pytest.fixture(scope="module")
def sp(native):
    from pysparse.sparse import spmatrix
    return spmatrix


@pytest.fixture(scope="module")
def poisson(native):
    from pysparse.tools import poisson
    return poisson


def dense(A):
    """Dense copy through find(); a symmetric matrix is expanded."""
    m, n = A.shape
    D = np.zeros((m, n))
    val, irow, jcol = A.find()
    D[irow, jcol] = val
    if A.issym:
        D[jcol, irow] = val
    return D


def from_dense(sp, D, sym=False):
    m, n = D.shape
    A = sp.ll_mat_sym(m) if sym else sp.ll_mat(m, n)
    for i in range(m):
        for j in range(n):
            if D[i, j] != 0.0 and (not sym or i >= j):
                A[i, j] = D[i, j]
    return A


def llmat_isEqual(a, b):                      # test/test_spmatrix.py:10-15
    if a.issym and not b.issym:
        a, b = b, a
    z = a.copy()
    z.shift(-1.0, b)
    return z.nnz == 0


def ll_mat_rand(sp, n, m, density, rnd):      # pysparse/tools/spmatrix_util.py:44-56
    A = sp.ll_mat(n, m, int(density * n * m))
    for _ in range(int(density * n * m)):
        A[rnd.randrange(n), rnd.randrange(m)] = rnd.random()
    return A


# ------------------------------------------------------------------ the reference's tests
def test_ref_create_and_entry(sp):
    n = 10
    A = sp.ll_mat(n, n); S = sp.ll_mat_sym(n)
    assert A.shape == (n, n) and A.nnz == 0 and not A.issym
    assert S.shape == (n, n) and S.nnz == 0 and S.issym
    A[0, 0] = 1.0; S[0, 0] = 1.0
    assert A[0, 0] == 1.0 and A.nnz == 1 and S[0, 0] == 1.0 and S.nnz == 1
    with pytest.raises(IndexError):
        S[0, 1] = 1.0
    A[0, 0] += 1.0
    assert A[0, 0] == 2.0 and A.nnz == 1
    A[0, 0] -= 2.0
    assert A[0, 0] == 0.0 and A.nnz == 0
    for idx in ((10, 0), (0, -11), (0, 10)):
        with pytest.raises(IndexError):
            S[idx] = 1.0
    I = sp.ll_mat(10, 10, 100)
    for i in range(10):
        for j in range(10):
            I[i, j] = 10 * i + j
    for i in range(-10, 0):
        for j in range(-10, 0):
            assert I[i, j] == I[10 + i, 10 + j]


def test_ref_poisson_basic_and_submatrices(sp, poisson):
    n = 20
    A = poisson.poisson2d(n); S = poisson.poisson2d_sym(n); B = poisson.poisson2d_sym_blk(n)
    assert S.nnz == n * (3 * n - 2) and A.nnz == n * (5 * n - 4)
    assert llmat_isEqual(A, A) and llmat_isEqual(S, S) and llmat_isEqual(A, S) and llmat_isEqual(A, B)
    Psym = poisson.poisson1d_sym(n); P = poisson.poisson1d(n)
    for i in range(n):
        P[i, i] = 4.0; Psym[i, i] = 4.0
    for i in range(n):
        blk = slice(n * i, n * (i + 1))
        assert llmat_isEqual(A[blk, blk], P) and llmat_isEqual(A[blk, blk], Psym)
        assert llmat_isEqual(S[blk, blk], P)           # commented out in the reference: works here
    R = ll_mat_rand(sp, n * n, n * n, 0.01, random.Random(7))
    for i in range(n):
        blk = slice(n * i, n * (i + 1))
        R[blk, blk] = P
        assert llmat_isEqual(R[blk, blk], P)
        R[blk, blk] = Psym
        assert llmat_isEqual(R[blk, blk], Psym)
    for i in range(n - 1):
        b0, b1 = slice(n * i, n * (i + 1)), slice(n * (i + 1), n * (i + 2))
        R[b0, b1] = P
        assert llmat_isEqual(R[b0, b1], P)
        R[b0, b1] = Psym
        assert llmat_isEqual(R[b0, b1], Psym)
    R = sp.ll_mat_sym(n * n)
    for i in range(n):
        blk = slice(n * i, n * (i + 1))
        R[blk, blk] = Psym
        assert llmat_isEqual(R[blk, blk], Psym)
    for i in range(n - 1):
        b0, b1 = slice(n * i, n * (i + 1)), slice(n * (i + 1), n * (i + 2))
        R[b1, b0] = P
        assert llmat_isEqual(R[b1, b0], P)
        R[b1, b0] = Psym
        assert llmat_isEqual(R[b1, b0], Psym)


@pytest.fixture()
def delete_case(sp, poisson):
    n = 30
    P = poisson.poisson1d(n)
    for i in range(n):
        P[i, i] = 4.0
    I = sp.ll_mat_sym(n)
    for i in range(n):
        I[i, i] = -1.0
    mask = np.zeros(n ** 2, "l"); mask[n // 2 * n:(n // 2 + 1) * n] = 1
    mask1 = np.zeros(n ** 2, "l"); mask1[(n // 2 + 1) * n:(n // 2 + 2) * n] = 1
    return dict(n=n, P=P, I=I, mask=mask, mask1=mask1, A=poisson.poisson2d(n), S=poisson.poisson2d_sym(n))


def test_ref_delete_rowcols(delete_case):
    c = delete_case
    c["S"].delete_rowcols(c["mask"])
    assert c["S"].issym and c["S"].shape == (c["n"], c["n"]) and llmat_isEqual(c["S"], c["P"])
    A = c["A"].copy(); A.delete_rowcols(c["mask"])
    assert llmat_isEqual(A, c["P"])
    A = c["A"].copy(); A.delete_rows(c["mask"]); A.delete_cols(c["mask"])
    assert A.shape == (c["n"], c["n"]) and llmat_isEqual(A, c["P"])
    A = c["A"].copy(); A.delete_rows(c["mask"]); A.delete_cols(c["mask1"])
    assert llmat_isEqual(A, c["I"])
    n1 = A.norm("fro"); A.compress(); assert A.norm("fro") == n1


def test_ref_compress_stress(sp):
    n = 20
    rnd = random.Random(3)
    A = sp.ll_mat(n, n)
    D = np.zeros((n, n))
    for _ in range(20):
        for _ in range(n * n // 2):
            i, j = rnd.randrange(n), rnd.randrange(n); A[i, j] = 1.0; D[i, j] = 1.0
        for _ in range(n * n // 2):
            i, j = rnd.randrange(n), rnd.randrange(n); A[i, j] = 0.0; D[i, j] = 0.0
        assert A.compress() >= 0
        assert np.array_equal(dense(A), D) and A.nnz == np.count_nonzero(D)


def test_ref_norms(sp, poisson):
    A = poisson.poisson2d(30)
    assert A.norm("1") == 8 and A.norm("inf") == 8 and poisson.poisson1d(3).norm("fro") == 4
    S = sp.ll_mat_sym(4)
    S[0, 0] = 1; S[1, 1] = 2; S[2, 2] = 3; S[3, 3] = 4; S[1, 0] = 3; S[2, 0] = 2; S[3, 0] = 2
    assert S.norm("fro") == 8
    for p in ("1", "inf"):
        with pytest.raises(NotImplementedError):
            poisson.poisson2d_sym(30).norm(p)


# ------------------------------------------------------------------ the documentation's examples
DOC_POISSON = """ll_mat(general, [5,5]):
 2.000000 -1.000000  --------  --------  --------
-1.000000  2.000000 -1.000000  --------  --------
 -------- -1.000000  2.000000 -1.000000  --------
 --------  -------- -1.000000  2.000000 -1.000000
 --------  --------  -------- -1.000000  2.000000
"""
DOC_REV = """ll_mat(general, [3,3]):
 --------  -------- -1.000000
 -------- -1.000000  2.000000
-1.000000  2.000000 -1.000000
"""
DOC_PERM = """ll_mat(general, [4,3]):
-1.000000 -1.000000  --------
 --------  --------  2.000000
 --------  2.000000  --------
 2.000000  --------  --------
"""
DOC_SEVEN = """ll_mat(general, [5,5]):
 7.000000  7.000000  7.000000  --------  --------
 7.000000  7.000000  7.000000  --------  --------
 7.000000  7.000000  7.000000 -1.000000  --------
 --------  -------- -1.000000  2.000000 -1.000000
 --------  --------  -------- -1.000000  2.000000
"""


def lines(text):
    """printed matrix without the trailing blank each cell carries (LLMatType_print pads every cell)"""
    return [l.rstrip() for l in str(text).splitlines()]


def test_doc_examples_printed_text(sp, poisson):
    n = 5
    A = poisson.poisson1d(n)
    assert str(A).splitlines()[1] == " 2.000000 -1.000000  --------  --------  -------- "
    assert lines(A) == lines(DOC_POISSON)                                    # spmatrix.rst:156-161
    assert lines(A[n - 1:1:-1, 1:n - 1]) == lines(DOC_REV)                   # :163-168
    assert lines(A[::-1, ::-1]) == lines(DOC_POISSON)                        # :186-192
    assert str(A[::-1, :]) == str(A[:, ::-1])                       # :170-184
    assert str(A[1:3, 3:]) == "ll_mat(general, [2,2]):\n --------  -------- \n-1.000000  -------- \n"
    assert str(A[::2, ::2]) == ("ll_mat(general, [3,3]):\n 2.000000  --------  -------- \n"
                                " --------  2.000000  -------- \n --------  --------  2.000000 \n")
    assert lines(A[[1, 4, 2, 0], ::2]) == lines(DOC_PERM)                    # :213-218
    assert lines(A[[1, 4, 2, 0], [0, 2, 4]]) == lines(DOC_PERM)              # :219-226
    assert lines(A[np.array([1, 4, 2, 0]), np.array([0, 2, 4])]) == lines(DOC_PERM)   # :230-237
    A[:3, :3] = 7
    assert lines(A) == lines(DOC_SEVEN)                                      # :242-249
    # entry-list form for matrices wider than 20 columns (LLMatType_print, ll_mat.c:3138-3155)
    L = sp.ll_mat(30, 30); L[3, 4] = 1.5; L[29, 0] = 1e-7
    assert str(L) == "ll_mat(general, [30,30], [(3,4): 1.5, (29,0): 1e-07])"
    assert str(sp.ll_mat_sym(25)) == "ll_mat(symmetric, [25,25])"
    # first example of the chapter (:106-142)
    B = sp.ll_mat(5, 5)
    for i in range(5):
        B[i, i] = i + 1
    B[2, 1] = B[0, 0]
    assert B.items() == [((0, 0), 1.0), ((1, 1), 2.0), ((2, 1), 1.0), ((2, 2), 3.0), ((3, 3), 4.0), ((4, 4), 5.0)]
    assert B[:2, :].items() == [((0, 0), 1.0), ((1, 1), 2.0)]
    assert B[:, 2:5].items() == [((2, 0), 3.0), ((3, 1), 4.0), ((4, 2), 5.0)]
    assert B[1:3, 2:5].items() == [((1, 0), 3.0)]
    C = sp.ll_mat(2, 2); C[0, 0] = -1; C[1, 1] = -1
    B[:2, :2] = C
    assert B.items() == [((0, 0), -1.0), ((1, 1), -1.0), ((2, 1), 1.0), ((2, 2), 3.0), ((3, 3), 4.0), ((4, 4), 5.0)]
    assert len(B) == 25 and B.storeZeros == 0


# ------------------------------------------------------------------ dense-model property tests
def random_index(rng, dim):
    kind = rng.integers(0, 4)
    if kind == 0:
        start = int(rng.integers(-dim, dim)); stop = int(rng.integers(-dim, dim + 1))
        step = int(rng.choice([1, 1, 2, 3, -1, -2]))
        return slice(start, stop, step)
    if kind == 1:
        return slice(None, None, int(rng.choice([1, -1, 2])))
    idx = rng.integers(-dim, dim, size=int(rng.integers(0, dim + 2)))
    return idx.tolist() if kind == 2 else idx.astype(np.int64)


def np_index(ix, dim):
    return np.arange(dim)[ix] if isinstance(ix, slice) else np.asarray(ix, dtype=np.int64) % dim if len(ix) else np.zeros(0, dtype=np.int64)


@pytest.mark.parametrize("seed", range(6))
def test_fancy_get_matches_dense_model(sp, seed):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(1, 14)), int(rng.integers(1, 14))
    D = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.4)
    A = from_dense(sp, D)
    for _ in range(60):
        ri, ci = random_index(rng, m), random_index(rng, n)
        sub = A[ri, ci]
        exp = D[np.ix_(np_index(ri, m), np_index(ci, n))]
        assert sub.shape == exp.shape and not sub.issym
        assert np.array_equal(dense(sub), exp), (ri, ci)
        assert sub.nnz == np.count_nonzero(exp)
    # symmetric storage: the sub-matrix of the FULL symmetric matrix
    k = int(rng.integers(2, 12))
    Ds = rng.standard_normal((k, k)) * (rng.random((k, k)) < 0.5); Ds = np.tril(Ds) + np.tril(Ds, -1).T
    S = from_dense(sp, Ds, sym=True)
    for _ in range(30):
        ri, ci = random_index(rng, k), random_index(rng, k)
        assert np.array_equal(dense(S[ri, ci]), Ds[np.ix_(np_index(ri, k), np_index(ci, k))])
    assert np.array_equal(dense(S[:, :]), Ds) and not S[:, :].issym


@pytest.mark.parametrize("seed", range(6))
def test_fancy_set_matches_dense_model(sp, seed):
    rng = np.random.default_rng(100 + seed)
    m, n = int(rng.integers(2, 12)), int(rng.integers(2, 12))
    D = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.4)
    A = from_dense(sp, D)
    for _ in range(40):
        ri, ci = random_index(rng, m), random_index(rng, n)
        rows, cols = np_index(ri, m), np_index(ci, n)
        if rng.random() < 0.4:
            v = float(rng.choice([0.0, 2.5, -1.0]))
            A[ri, ci] = v
            D[np.ix_(rows, cols)] = v
        else:
            M = rng.standard_normal((len(rows), len(cols))) * (rng.random((len(rows), len(cols))) < 0.5)
            A[ri, ci] = from_dense(sp, M)
            for a, r in enumerate(rows):              # sequential semantics for repeated indices
                for b, c in enumerate(cols):
                    D[r, c] = 0.0
            for a, r in enumerate(rows):
                for b, c in enumerate(cols):
                    if M[a, b] != 0.0:
                        D[r, c] = M[a, b]
        assert np.array_equal(dense(A), D), (ri, ci)
        assert A.nnz == np.count_nonzero(D)
    with pytest.raises(ValueError, match="shapes"):
        A[0:2, 0:2] = sp.ll_mat(3, 3)
    A[:, :] = A                                      # self-assignment reads a snapshot
    assert np.array_equal(dense(A), D)


def test_fancy_set_symmetric_rules(sp):
    rng = np.random.default_rng(5)
    k = 8
    Ds = rng.standard_normal((4, 4)); Ds = np.tril(Ds) + np.tril(Ds, -1).T
    Msym = from_dense(sp, Ds, sym=True)
    S = sp.ll_mat_sym(k)
    S[2:6, 2:6] = Msym                               # diagonal block, symmetric source
    exp = np.zeros((k, k)); exp[2:6, 2:6] = Ds
    assert np.array_equal(dense(S), exp) and S.issym
    S[5:1:-1, 5:1:-1] = Msym                         # reversed: the mirrored entries land below the diagonal
    exp[2:6, 2:6] = Ds[::-1, ::-1]
    assert np.array_equal(dense(S), exp)
    G = sp.ll_mat(k, k)
    G[0:4, 4:8] = Msym                               # a symmetric source is the FULL matrix, anywhere
    assert np.array_equal(dense(G)[0:4, 4:8], Ds)
    Mgen = from_dense(sp, Ds)
    before = dense(S).copy()
    with pytest.raises(IndexError, match="upper triangle"):
        S[2:6, 2:6] = Mgen                           # a general block reaching above the diagonal
    with pytest.raises(IndexError, match="upper triangle"):
        S[0:3, 0:3] = 1.0
    assert np.array_equal(dense(S), before)          # nothing was written
    S[4:8, 0:4] = Mgen                               # strictly below the diagonal: fine
    assert np.array_equal(dense(S)[4:8, 0:4], Ds)
    S[6, 0:3] = 0.0
    assert np.all(dense(S)[6, 0:3] == 0.0)


def test_index_errors(sp):
    A = sp.ll_mat(4, 5)
    for bad in ((4, 0), (0, 5), ([0, 4], 0), (0, [5]), (np.array([-5]), 0)):
        with pytest.raises(IndexError):
            A[bad]
        with pytest.raises(IndexError):
            A[bad] = 1.0
    with pytest.raises(IndexError):
        A[1]
    with pytest.raises(IndexError):
        A[1, 2, 3]
    with pytest.raises(IndexError):
        A[1.5, 0]
    with pytest.raises(IndexError):
        A[[0.5], 0]
    with pytest.raises(ValueError):
        A[0, 0] = "x"
    with pytest.raises(ValueError):
        A[0, 0] = sp.ll_mat(1, 1)
    assert A[np.int64(1), np.int32(2)] == 0.0


@pytest.mark.parametrize("seed", range(4))
def test_update_add_mask_and_sym(sp, seed):
    rng = np.random.default_rng(200 + seed)
    n = 9
    A = sp.ll_mat(n, n + 2); D = np.zeros((n, n + 2))
    for _ in range(20):
        l0, l1 = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        B = rng.standard_normal((l0, l1))
        i0 = rng.integers(-n, n, l0); i1 = rng.integers(-(n + 2), n + 2, l1)
        m0 = rng.integers(0, 2, l0); m1 = rng.integers(0, 2, l1)
        A.update_add_mask(B, i0, i1, m0, m1)
        for i in range(l0):                           # spmatrix.rst:366-371
            for j in range(l1):
                if m0[i] and m1[j]:
                    D[i0[i], i1[j]] += B[i, j]
        assert np.array_equal(dense(A), D)
    with pytest.raises(sp.error):
        sp.ll_mat_sym(3).update_add_mask(np.zeros((1, 1)), [0], [0], [1], [1])
    with pytest.raises(ValueError):
        A.update_add_mask(np.zeros((2, 2)), [0, 1], [0, 1], [1], [1, 1])
    with pytest.raises(ValueError):
        A.update_add_mask(np.zeros((2, 3)), [0, 1], [0, 1], [1, 1], [1, 1])
    with pytest.raises(IndexError):
        A.update_add_mask(np.ones((1, 1)), [n], [0], [1], [1])
    # symmetric element matrices into a symmetric and a general matrix: same full matrix
    S = sp.ll_mat_sym(n); G = sp.ll_mat(n, n); Dm = np.zeros((n, n))
    for _ in range(20):
        l = int(rng.integers(1, 5))
        B = rng.standard_normal((l, l)); B = np.tril(B) + np.tril(B, -1).T
        ind = rng.choice(n, l, replace=False); mask = rng.integers(0, 2, l)
        S.update_add_mask_sym(B, ind, mask); G.update_add_mask_sym(B, ind, mask)
        for i in range(l):                            # ll_mat.c:2345-2377
            for j in range(l):
                if mask[i] and mask[j]:
                    Dm[ind[i], ind[j]] += B[i, j]
        assert np.allclose(dense(S), Dm, rtol=0, atol=1e-15) and np.allclose(dense(G), Dm, rtol=0, atol=1e-15)
    assert S.issym and not G.issym


def test_row_col_scale_and_masks(sp):
    rng = np.random.default_rng(9)
    m, n = 7, 5
    D = rng.standard_normal((m, n)) * (rng.random((m, n)) < 0.6)
    A = from_dense(sp, D)
    r, c = rng.standard_normal(m), rng.standard_normal(n)
    A.row_scale(r); D *= r[:, None]
    assert np.array_equal(dense(A), D)
    A.col_scale(c); D *= c[None, :]
    assert np.array_equal(dense(A), D)
    with pytest.raises(sp.error):
        A.col_scale(np.ones(m))
    with pytest.raises(sp.error):
        A.row_scale(np.ones(n))
    mr = np.array([1, 0, 1, 1, 0, 0, 1]); mc = np.array([0, 1, 1, 0, 1])
    B = A.copy(); B.delete_rows(mr)
    assert B.shape == (4, n) and np.array_equal(dense(B), D[mr != 0])
    B.delete_cols(mc)
    assert B.shape == (4, 3) and np.array_equal(dense(B), D[mr != 0][:, mc != 0]) and B.nnz == np.count_nonzero(dense(B))
    B[3, 2] = 5.0                                     # still a working container after shrinking
    assert B[3, 2] == 5.0
    for bad in (np.ones(m + 1, dtype=int), np.ones((m, 1), dtype=int), np.ones(m)):
        with pytest.raises(ValueError):
            A.delete_rows(bad)
    S = sp.ll_mat_sym(4)
    with pytest.raises(sp.error):
        S.delete_rows(np.ones(4, dtype=int))
    with pytest.raises(sp.error):
        S.delete_cols(np.ones(4, dtype=int))
    with pytest.raises(ValueError):
        sp.ll_mat(3, 4).delete_rowcols(np.ones(3, dtype=int))
    Ds = rng.standard_normal((6, 6)); Ds = np.tril(Ds) + np.tril(Ds, -1).T
    S = from_dense(sp, Ds, sym=True)
    keep = np.array([1, 1, 0, 1, 0, 1])
    S.delete_rowcols(keep)
    assert S.issym and np.array_equal(dense(S), Ds[keep != 0][:, keep != 0])


def test_conversions_after_editing(sp, poisson):
    """The containers edited with the operations above still feed the hot path: to_csr / to_sss
    of a sliced, masked, scaled matrix are bit-exact against the dense model."""
    A = poisson.poisson2d(6)
    sub = A[::2, :]
    sub.col_scale(np.arange(1.0, 37.0))
    mask = np.ones(36, dtype=int); mask[5:9] = 0
    sub.delete_cols(mask)
    D = dense(sub)
    val, col, ind = sub.to_csr().arrays()
    assert sub.shape == (18, 32) and ind[-1] == sub.nnz
    R = np.zeros_like(D)
    for i in range(18):
        R[i, col[ind[i]:ind[i + 1]]] = val[ind[i]:ind[i + 1]]
        assert np.all(np.diff(col[ind[i]:ind[i + 1]]) > 0)
    assert np.array_equal(R, D)
