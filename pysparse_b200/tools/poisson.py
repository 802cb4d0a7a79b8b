"""Poisson matrix generators through the ``ll_mat`` API.

Python-3 versions of ``pysparse/tools/poisson.py:4-66`` and
``pysparse/tools/poisson_vec.py:5-98`` (same orderings and values); for large
grids use ``pysparse_b200.tools.gallery`` + ``spmatrix.csr_from_arrays``.
"""
import numpy as np

from ..sparse import spmatrix


def _laplacian_triples(shape, lower_only):
    """Yield (row, col, value) of the (2d+1)-point Laplacian on a lexicographically numbered
    grid of extents ``shape`` (first index fastest): the diagonal 2d first, then the couplings to
    the lower / upper neighbour along each axis, row after row."""
    dims = len(shape)
    strides = [int(np.prod(shape[:a], dtype=np.int64)) for a in range(dims)]
    for idx in np.ndindex(*reversed(shape)):
        pos = idx[::-1]                                   # pos[a]: coordinate along axis a
        row = sum(c * s for c, s in zip(pos, strides))
        yield row, row, 2.0 * dims
        for a in range(dims):
            if pos[a] > 0:
                yield row, row - strides[a], -1.0
            if not lower_only and pos[a] < shape[a] - 1:
                yield row, row + strides[a], -1.0


def _assign_elementwise(L, triples):
    "one ``L[i, j] = v`` per entry: the slow path the reference's generators exercise"
    for i, j, v in triples:
        L[i, j] = v
    return L


def poisson1d(n):
    return _assign_elementwise(spmatrix.ll_mat(n, n, 3 * n - 2), _laplacian_triples((n,), False))


def poisson1d_sym(n):
    return _assign_elementwise(spmatrix.ll_mat_sym(n, 2 * n - 1), _laplacian_triples((n,), True))


def poisson2d(n):
    return _assign_elementwise(spmatrix.ll_mat(n * n, n * n, 5 * n * n - 4 * n), _laplacian_triples((n, n), False))


def poisson2d_sym(n):
    return _assign_elementwise(spmatrix.ll_mat_sym(n * n, 3 * n * n - 2 * n), _laplacian_triples((n, n), True))


def poisson2d_sym_blk(n):
    "block-wise assembly through slice assignment: tridiagonal blocks T, coupling blocks -I"
    L = spmatrix.ll_mat_sym(n * n, 3 * n * n - 2 * n)
    T = poisson1d_sym(n)
    T.shift(2.0, _identity_sym(n))                        # 1-D stencil (2, -1) -> (4, -1)
    minus_I = _identity_sym(n)
    minus_I.scale(-1.0)
    for blk in range(n):
        lo, hi = blk * n, (blk + 1) * n
        L[lo:hi, lo:hi] = T
        if blk:
            L[lo:hi, lo - n:lo] = minus_I
    return L


def _identity_sym(n):
    eye = spmatrix.ll_mat_sym(n, n)
    d = np.arange(n, dtype=np.int64)
    eye.put(np.ones(n), d, d)
    return eye


def poisson1d_vec(n):
    L = spmatrix.ll_mat(n, n, 3 * n - 2)
    e = np.ones(n)
    d = np.arange(n, dtype=np.int64)
    L.put(2 * e, d, d)
    L.put(-e[1:], d[1:], d[:-1])
    L.put(-e[1:], d[:-1], d[1:])
    return L


def poisson2d_vec(n):
    n2 = n * n
    L = spmatrix.ll_mat(n2, n2, 5 * n2 - 4 * n)
    e = np.ones(n2)
    d = np.arange(n2, dtype=np.int64)
    L.put(4 * e, d, d)
    inner = d[(d % n) != 0]
    L.put(-np.ones(len(inner)), inner, inner - 1)
    L.put(-np.ones(len(inner)), inner - 1, inner)
    L.put(-e[n:], d[n:], d[:-n])
    L.put(-e[n:], d[:-n], d[n:])
    return L


def poisson2d_sym_vec(n):
    n2 = n * n
    L = spmatrix.ll_mat_sym(n2, 3 * n2 - 2 * n)
    e = np.ones(n2)
    d = np.arange(n2, dtype=np.int64)
    L.put(4 * e, d, d)
    inner = d[(d % n) != 0]
    L.put(-np.ones(len(inner)), inner, inner - 1)
    L.put(-e[n:], d[n:], d[:-n])
    return L


def poisson1d_sym_vec(n):
    L = spmatrix.ll_mat_sym(n, 2 * n - 1)
    e = np.ones(n)
    d = np.arange(n, dtype=np.int64)
    L.put(2 * e, d, d)
    L.put(-e[1:], d[1:], d[:-1])
    return L


def poisson2d_vec2(n):
    "Same matrix as :func:`poisson2d_vec` (the reference keeps two vectorised variants)."
    return poisson2d_vec(n)


def poisson2d_sym_blk_vec(n):
    "Block assembly with vectorised blocks: diagonal blocks D, sub-diagonal blocks -I."
    n2 = n * n
    L = spmatrix.ll_mat_sym(n2, 3 * n2 - 2 * n)
    D = poisson1d_sym_vec(n)
    d = np.arange(n, dtype=np.int64)
    D.put(4 * np.ones(n), d, d)
    P = spmatrix.ll_mat(n, n, n)
    P.put(-np.ones(n), d, d)
    for i in range(n - 1):
        L[i * n:(i + 1) * n, i * n:(i + 1) * n] = D
        L[(i + 1) * n:(i + 2) * n, i * n:(i + 1) * n] = P
    L[n2 - n:n2, n2 - n:n2] = D
    return L
