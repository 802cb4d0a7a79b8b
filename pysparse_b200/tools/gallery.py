"""Synthetic matrix gallery for benchmarks and parity tests.

Vectorised NumPy builders that return the *final* CSR / SSS arrays directly
(``ll_mat`` element-wise assembly is far too slow at 16.7 M rows).  Ordering
and values coincide with the reference's generators where one exists:

* ``poisson2d``  -- ``pysparse/tools/poisson.py:22-50`` (``k = i + n*j``, diag 4, off -1)
* ``poisson1d``  -- ``pysparse/tools/poisson.py:4-20``
* ``poisson3d``, ``convdiff3d``, ``powerlaw`` -- not in the reference; defined in
  SURVEY.md section 8(d) (``k = i + n*j + n*n*l``).

Every builder returns a ``Stencil``/dict with ``csr = (val, col, ind)`` in the
column order ``ll_mat.to_csr()`` produces (ascending columns) and, for the
symmetric ones, ``sss = (val, diag, col, ind)`` as ``ll_mat.to_sss()`` would
(strict lower triangle, ascending columns, dense diagonal).
"""
from __future__ import annotations

import numpy as np


def _stencil(dims, offsets_lower, w_lower, w_diag, w_upper, want_csr=True, want_sss=True,
             chunk_rows=1 << 22, rows=None):
    """Generic structured-grid builder.  ``dims`` = (n0, n1, ...) with index
    k = i0 + n0*i1 + n0*n1*i2...; neighbour d along axis a exists iff the
    coordinate stays inside the grid.  ``offsets_lower`` is implied: axis a has
    linear stride s_a; lower neighbour k - s_a (weight w_lower[a]), upper
    k + s_a (weight w_upper[a])."""
    dims = tuple(int(d) for d in dims)
    nd = len(dims)
    strides = [1]
    for d in dims[:-1]:
        strides.append(strides[-1] * d)
    N = int(np.prod(dims))
    out = {"n": N, "dims": dims}

    def coords(k):
        cs = []
        for a in range(nd):
            cs.append((k // strides[a]) % dims[a])
        return cs

    # `rows = (a, b)`: only rows a..b-1 are generated (a row slab for one GPU); column
    # indices stay GLOBAL and the row pointer is relative to the slab.
    ra, rb = (0, N) if rows is None else (int(rows[0]), int(rows[1]))
    nloc = rb - ra
    out["rows"] = (ra, rb)
    csr_parts, sss_parts = [], []
    csr_cnt = np.empty(nloc, dtype=np.int32) if want_csr else None
    sss_cnt = np.empty(nloc, dtype=np.int32) if want_sss else None
    for r0 in range(ra, rb, chunk_rows):
        r1 = min(rb, r0 + chunk_rows)
        k = np.arange(r0, r1, dtype=np.int64)
        cs = coords(k)
        # column slots in ascending column order: lower axes descending, diag, upper ascending
        cols, vals, masks = [], [], []
        for a in reversed(range(nd)):
            masks.append(cs[a] > 0)
            cols.append(k - strides[a])
            vals.append(w_lower[a])
        nlow = len(cols)
        masks.append(np.ones(r1 - r0, dtype=bool)); cols.append(k); vals.append(w_diag)
        for a in range(nd):
            masks.append(cs[a] < dims[a] - 1)
            cols.append(k + strides[a])
            vals.append(w_upper[a])
        M = np.stack(masks, axis=1)
        Cc = np.stack(cols, axis=1)
        Vv = np.broadcast_to(np.asarray(vals, dtype=np.float64), M.shape)
        if want_csr:
            csr_cnt[r0 - ra:r1 - ra] = M.sum(axis=1)
            csr_parts.append((Vv[M], Cc[M].astype(np.int32)))
        if want_sss:
            Ml = M[:, :nlow]
            sss_cnt[r0 - ra:r1 - ra] = Ml.sum(axis=1)
            sss_parts.append((Vv[:, :nlow][Ml], Cc[:, :nlow][Ml].astype(np.int32)))
    if want_csr:
        ind = np.zeros(nloc + 1, dtype=np.int64)
        np.cumsum(csr_cnt, out=ind[1:])
        assert ind[-1] < 2**31
        out["csr"] = (np.concatenate([p[0] for p in csr_parts]) if csr_parts else np.zeros(0),
                      np.concatenate([p[1] for p in csr_parts]) if csr_parts else np.zeros(0, np.int32),
                      ind.astype(np.int32))
    if want_sss:
        ind = np.zeros(nloc + 1, dtype=np.int64)
        np.cumsum(sss_cnt, out=ind[1:])
        out["sss"] = (np.concatenate([p[0] for p in sss_parts]) if sss_parts else np.zeros(0),
                      np.full(nloc, float(w_diag)),
                      np.concatenate([p[1] for p in sss_parts]) if sss_parts else np.zeros(0, np.int32),
                      ind.astype(np.int32))
    return out


def poisson1d(n, **kw):
    return _stencil((n,), None, [-1.0], 2.0, [-1.0], **kw)


def poisson2d(n, **kw):
    """2-D 5-point Laplacian on an n x n grid (N = n^2)."""
    return _stencil((n, n), None, [-1.0, -1.0], 4.0, [-1.0, -1.0], **kw)


def poisson3d(n, **kw):
    """3-D 7-point Laplacian on an n^3 grid (N = n^3), diag 6, off -1."""
    return _stencil((n, n, n), None, [-1.0] * 3, 6.0, [-1.0] * 3, **kw)


def convdiff3d(n, gamma=0.5, **kw):
    """Non-symmetric 3-D convection-diffusion: diag 6,
    lower neighbours -1-gamma, upper neighbours -1+gamma."""
    kw["want_sss"] = False
    return _stencil((n, n, n), None, [-1.0 - gamma] * 3, 6.0, [-1.0 + gamma] * 3, **kw)


def powerlaw(nrows, ncols=None, mean_nnz=30, seed=12345, max_len=65536):
    """Irregular CSR (SURVEY.md 8(d) #5b): Pareto(alpha=2) row lengths rescaled
    to ~mean_nnz, columns sampled uniformly, sorted ascending and unique per
    row, values U(-1, 1)."""
    ncols = nrows if ncols is None else ncols
    rng = np.random.default_rng(seed)
    raw = (rng.pareto(2.0, nrows) + 1.0)
    lens = np.clip(np.rint(raw * (mean_nnz / raw.mean())), 1, min(max_len, ncols)).astype(np.int64)
    ind = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(lens, out=ind[1:])
    nnz = int(ind[-1])
    assert nnz < 2**31
    rows = np.repeat(np.arange(nrows, dtype=np.int64), lens)
    cols = rng.integers(0, ncols, nnz, dtype=np.int64)
    # sort by (row, col) and drop duplicates inside a row
    key = rows * ncols + cols
    key = np.unique(key)
    rows = key // ncols
    cols = (key % ncols).astype(np.int32)
    cnt = np.bincount(rows, minlength=nrows)
    ind = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(cnt, out=ind[1:])
    val = rng.uniform(-1.0, 1.0, len(cols))
    return {"n": nrows, "shape": (nrows, ncols), "csr": (val, cols, ind.astype(np.int32))}


def random_sym(n, mean_nnz=6, seed=0, diag_shift=None):
    """Random symmetric, diagonally dominant (SPD) matrix; returns csr + sss."""
    rng = np.random.default_rng(seed)
    nlow = max(0, int(n * mean_nnz // 2))
    i = rng.integers(0, n, nlow)
    j = rng.integers(0, n, nlow)
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    keep = lo != hi
    key = np.unique(hi[keep].astype(np.int64) * n + lo[keep])
    r = (key // n).astype(np.int64)
    c = (key % n).astype(np.int32)
    v = rng.uniform(-1.0, 1.0, len(key))
    rowsum = np.bincount(r, weights=np.abs(v), minlength=n) + np.bincount(c, weights=np.abs(v), minlength=n)
    diag = rowsum + (1.0 if diag_shift is None else diag_shift) + rng.uniform(0, 1, n)
    ind = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=ind[1:])
    sss = (v, diag, c, ind.astype(np.int32))
    return {"n": n, "sss": sss, "csr": sss_to_csr(n, *sss)}


def sss_to_csr(n, val, diag, col, ind):
    """Expand SSS to full CSR in ``ll_mat.to_csr()`` order for a symmetric
    matrix (``ll_mat.c:1601-1623``): row i = its lower entries (ascending col),
    the diagonal, then column-i entries of later rows (ascending row).
    Zero diagonals are kept (structurally present)."""
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ind))
    col = col.astype(np.int64)
    r_all = np.concatenate([rows, np.arange(n, dtype=np.int64), col])
    c_all = np.concatenate([col, np.arange(n, dtype=np.int64), rows])
    v_all = np.concatenate([val, diag, val])
    order = np.lexsort((c_all, r_all))
    cnt = np.bincount(r_all, minlength=n)
    oind = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=oind[1:])
    return (v_all[order], c_all[order].astype(np.int32), oind.astype(np.int32))
