"""Poisson matrix generators through the ``ll_mat`` API.

Python-3 versions of ``pysparse/tools/poisson.py:4-66`` and
``pysparse/tools/poisson_vec.py:5-98`` (same orderings and values); for large
grids use ``pysparse_b200.tools.gallery`` + ``spmatrix.csr_from_arrays``.
"""
import numpy as np

from ..sparse import spmatrix


def poisson1d(n):
    L = spmatrix.ll_mat(n, n, 3 * n - 2)
    for i in range(n):
        L[i, i] = 2
        if i > 0:
            L[i, i - 1] = -1
        if i < n - 1:
            L[i, i + 1] = -1
    return L


def poisson1d_sym(n):
    L = spmatrix.ll_mat_sym(n, 2 * n - 1)
    for i in range(n):
        L[i, i] = 2
        if i > 0:
            L[i, i - 1] = -1
    return L


def poisson2d(n):
    n2 = n * n
    L = spmatrix.ll_mat(n2, n2, 5 * n2 - 4 * n)
    for i in range(n):
        for j in range(n):
            k = i + n * j
            L[k, k] = 4
            if i > 0:
                L[k, k - 1] = -1
            if i < n - 1:
                L[k, k + 1] = -1
            if j > 0:
                L[k, k - n] = -1
            if j < n - 1:
                L[k, k + n] = -1
    return L


def poisson2d_sym(n):
    n2 = n * n
    L = spmatrix.ll_mat_sym(n2, 3 * n2 - 2 * n)
    for i in range(n):
        for j in range(n):
            k = i + n * j
            L[k, k] = 4
            if i > 0:
                L[k, k - 1] = -1
            if j > 0:
                L[k, k - n] = -1
    return L


def poisson2d_sym_blk(n):
    n2 = n * n
    L = spmatrix.ll_mat_sym(n2, 3 * n2 - 2 * n)
    I = spmatrix.ll_mat_sym(n, n)
    for i in range(n):
        I[i, i] = -1
    P = spmatrix.ll_mat_sym(n, 2 * n - 1)
    for i in range(n):
        P[i, i] = 4
        if i > 0:
            P[i, i - 1] = -1
    for i in range(0, n * n, n):
        L[i:i + n, i:i + n] = P
        if i > 0:
            L[i:i + n, i - n:i] = I
    return L


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
