"""Shared case tables for the golden generator and the parity tests."""
from __future__ import annotations

import numpy as np

from pysparse_b200.tools import gallery as g


def build_matrix(spec):
    kind = spec[0]
    if kind == "poisson1d":
        return g.poisson1d(spec[1])
    if kind == "poisson2d":
        return g.poisson2d(spec[1])
    if kind == "poisson3d":
        return g.poisson3d(spec[1])
    if kind == "convdiff3d":
        return g.convdiff3d(spec[1], spec[2] if len(spec) > 2 else 0.5)
    if kind == "powerlaw":
        return g.powerlaw(spec[1], mean_nnz=spec[2], seed=spec[3], max_len=spec[4])
    if kind == "random_sym":
        return g.random_sym(spec[1], mean_nnz=spec[2], seed=spec[3])
    if kind == "rect":
        m, n, seed = spec[1], spec[2], spec[3]
        return g.powerlaw(m, ncols=n, mean_nnz=5, seed=seed, max_len=n)
    if kind == "diag":      # diagonal-only matrix, examples/jdsym_test.py style
        n = spec[1]
        d = np.arange(1, n + 1, dtype=np.float64)
        z = np.zeros(n + 1, dtype=np.int32)
        return {"n": n, "sss": (np.zeros(0), d, np.zeros(0, dtype=np.int32), z),
                "csr": (d.copy(), np.arange(n, dtype=np.int32), np.arange(n + 1, dtype=np.int32))}
    if kind == "tendigit":  # examples/tendigit.py (Trefethen problem 7), scaled down
        n = spec[1]
        return tendigit(n)
    raise ValueError(spec)


def primes(n):
    out, c = [], 2
    while len(out) < n:
        if all(c % p for p in out if p * p <= c):
            out.append(c)
        c += 1
    return np.array(out, dtype=np.float64)


def tendigit(n):
    """A[i,i] = i-th prime; A[i,j] = 1 for |i-j| in {1,2,4,8,...}  (examples/tendigit.py:26-38)"""
    rows, cols = [], []
    for i in range(n):
        k = 1
        js = []
        while i - k >= 0:
            js.append(i - k)
            k *= 2
        for j in sorted(js):
            rows.append(i); cols.append(j)
    rows = np.array(rows, dtype=np.int64); cols = np.array(cols, dtype=np.int32)
    ind = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=ind[1:])
    sss = (np.ones(len(cols)), primes(n), cols, ind.astype(np.int32))
    return {"n": n, "sss": sss, "csr": g.sss_to_csr(n, *sss)}


def diag_of(mat, fmt):
    n = mat["n"]
    if fmt == "sss":
        return mat["sss"][1]
    val, col, ind = mat["csr"]
    rows = np.repeat(np.arange(n), np.diff(ind))
    d = np.zeros(n)
    m = col == rows
    d[rows[m]] = val[m]
    return d


def make_ops(lib, mat, case):
    """Build (A, K) with an oracle-like factory `lib` (Oracle or the GPU shim in tests)."""
    n = mat["n"]
    fmt = case["fmt"]
    A = lib.csr(n, n, *mat["csr"]) if fmt == "csr" else lib.sss(n, *mat["sss"])
    prec = case.get("prec")
    K = None
    if prec == "jacobi":
        K = lib.jacobi(A, diag_of(mat, fmt), case.get("omega", 1.0), case.get("steps", 1))
    elif prec == "ssor":
        K = lib.ssor(A, case.get("omega", 1.0), case.get("steps", 1))
    return A, K


def _c(id, matrix, fmt, solver, tol, maxit, prec=None, **kw):
    d = dict(id=id, matrix=list(matrix), fmt=fmt, solver=solver, tol=tol, maxit=maxit, prec=prec)
    d.update(kw)
    return d


# `stable`: iteration count is insensitive to rounding (checked with an FMA
# build of the oracle), so the +-2 % iteration contract is asserted; the
# others (CGS / BiCGSTAB on Poisson) are rounding-chaotic even on the CPU
# (see DESIGN.md "Parity") and are held to flag + relres < tol + solution.
SOLVER_CASES = [
    _c("p2d100_sss_pcg", ("poisson2d", 100), "sss", "pcg", 1e-8, 2000, stable=True),
    _c("p2d100_sss_pcg_1e-12", ("poisson2d", 100), "sss", "pcg", 1e-12, 2000, stable=True),
    _c("p2d100_csr_pcg_jacobi", ("poisson2d", 100), "csr", "pcg", 1e-8, 2000, "jacobi", stable=True),
    _c("p2d300_sss_pcg_jacobi", ("poisson2d", 300), "sss", "pcg", 1e-10, 2000, "jacobi", stable=True),   # BASELINE config 1
    _c("p2d100_sss_pcg_maxit", ("poisson2d", 100), "sss", "pcg", 1e-8, 50, stable=True),                  # iter = maxit+1 quirk
    _c("p2d60_sss_pcg_x0", ("poisson2d", 60), "sss", "pcg", 1e-9, 2000, x0="half", stable=True),
    _c("p2d60_sss_pcg_bzero", ("poisson2d", 60), "sss", "pcg", 1e-9, 2000, b="zero", x0="half", stable=True),
    _c("p2d100_csr_minres_jacobi", ("poisson2d", 100), "csr", "minres", 1e-8, 2000, "jacobi", stable=True),
    _c("p2d100_sss_minres", ("poisson2d", 100), "sss", "minres", 1e-8, 2000, stable=True),
    _c("p2d100_csr_minres_maxit", ("poisson2d", 100), "csr", "minres", 1e-8, 40, "jacobi", stable=True),
    _c("p2d100_csr_qmrs_jacobi", ("poisson2d", 100), "csr", "qmrs", 1e-8, 2000, "jacobi", stable=True),
    _c("p2d100_sss_qmrs", ("poisson2d", 100), "sss", "qmrs", 1e-8, 2000, stable=True),
    _c("p2d60_sss_qmrs_x0_ignored", ("poisson2d", 60), "sss", "qmrs", 1e-8, 2000, x0="half", stable=True),
    _c("p2d100_csr_cgs_jacobi", ("poisson2d", 100), "csr", "cgs", 1e-8, 2000, "jacobi", stable=False),
    _c("p2d100_csr_bicgstab_jacobi", ("poisson2d", 100), "csr", "bicgstab", 1e-8, 2000, "jacobi", stable=False),
    _c("p2d100_csr_gmres30_jacobi", ("poisson2d", 100), "csr", "gmres", 1e-8, 2000, "jacobi", dim=30, stable=True),
    _c("p2d100_csr_gmres20", ("poisson2d", 100), "csr", "gmres", 1e-6, 300, dim=20, stable=True),
    _c("p2d100_sss_pcg_ssor", ("poisson2d", 100), "sss", "pcg", 1e-8, 2000, "ssor", stable=True),
    _c("p2d100_sss_pcg_ssor19", ("poisson2d", 100), "sss", "pcg", 1e-8, 2000, "ssor", omega=1.9, stable=True),
    _c("p2d200_sss_pcg_ssor15x2", ("poisson2d", 200), "sss", "pcg", 1e-8, 2000, "ssor", omega=1.5, steps=2, stable=True),
    _c("p2d300_sss_pcg_ssor", ("poisson2d", 300), "sss", "pcg", 1e-10, 2000, "ssor", stable=True),
    _c("p2d100_sss_minres_ssor", ("poisson2d", 100), "sss", "minres", 1e-8, 2000, "ssor", stable=True),
    _c("p3d32_csr_pcg_jacobi", ("poisson3d", 32), "csr", "pcg", 1e-10, 2000, "jacobi", stable=True),
    _c("p3d32_csr_pcg_jacobi2", ("poisson3d", 32), "csr", "pcg", 1e-10, 2000, "jacobi", steps=2, stable=True),
    _c("p3d32_csr_minres_jacobi", ("poisson3d", 32), "csr", "minres", 1e-10, 2000, "jacobi", stable=True),
    _c("p3d32_csr_qmrs_jacobi", ("poisson3d", 32), "csr", "qmrs", 1e-10, 2000, "jacobi", stable=True),
    _c("p3d32_csr_bicgstab_jacobi", ("poisson3d", 32), "csr", "bicgstab", 1e-10, 2000, "jacobi", stable=False),
    _c("p3d32_csr_cgs_jacobi", ("poisson3d", 32), "csr", "cgs", 1e-8, 2000, "jacobi", stable=False),
    _c("p3d32_csr_gmres30_jacobi", ("poisson3d", 32), "csr", "gmres", 1e-10, 2000, "jacobi", dim=30, stable=True),
    _c("p3d64_sss_pcg_jacobi", ("poisson3d", 64), "sss", "pcg", 1e-10, 2000, "jacobi", stable=True),
    _c("cd32_csr_bicgstab_jacobi", ("convdiff3d", 32), "csr", "bicgstab", 1e-8, 2000, "jacobi", stable=True),
    _c("cd32_csr_gmres30_jacobi", ("convdiff3d", 32), "csr", "gmres", 1e-8, 2000, "jacobi", dim=30, stable=True),
    _c("cd32_csr_gmres5_maxit", ("convdiff3d", 32), "csr", "gmres", 1e-12, 37, "jacobi", dim=5, stable=True),
    _c("cd32_csr_cgs_maxit", ("convdiff3d", 24), "csr", "cgs", 1e-8, 25, "jacobi", stable=True),
    _c("tendigit2000_minres_ssor", ("tendigit", 2000), "sss", "minres", 1e-14, 2000, "ssor", stable=True),
    _c("diag1000_qmrs", ("diag", 1000), "sss", "qmrs", 1e-10, 2000, "jacobi", stable=True),
    _c("rsym5000_pcg_ssor", ("random_sym", 5000, 8, 7), "sss", "pcg", 1e-10, 500, "ssor", omega=1.2, stable=True),
    _c("rsym5000_bicgstab", ("random_sym", 5000, 8, 7), "csr", "bicgstab", 1e-10, 500, "jacobi", stable=True),
    # CGS cases that CONVERGE and whose iteration count does not move when the reference's own C
    # is recompiled with FMA contraction (stock / FMA build: 60 / 60, 31 / 31, 33 / 33, 9 / 9) --
    # held to the +-2 % contract
    _c("p2d40_csr_cgs_jacobi", ("poisson2d", 40), "csr", "cgs", 1e-8, 2000, "jacobi", stable=True),
    _c("p3d16_csr_cgs", ("poisson3d", 16), "csr", "cgs", 1e-8, 2000, stable=True),
    _c("cd12_csr_cgs_jacobi", ("convdiff3d", 12), "csr", "cgs", 1e-8, 2000, "jacobi", stable=True),
    _c("rsym5000_cgs_jacobi", ("random_sym", 5000, 8, 7), "csr", "cgs", 1e-10, 500, "jacobi", stable=True),
]

KERNEL_CASES = [
    dict(id="p2d_37", matrix=["poisson2d", 37], seed=1, ssor=[(1.0, 1), (1.0, 2), (1.6, 1), (0.8, 3)], jacobi=[(1.0, 1), (0.9, 3)]),
    dict(id="p3d_13", matrix=["poisson3d", 13], seed=2, ssor=[(1.0, 1), (1.3, 2)], jacobi=[(1.0, 2)]),
    dict(id="cd3d_11", matrix=["convdiff3d", 11], seed=3),
    dict(id="pow_3000", matrix=["powerlaw", 3000, 30, 11, 2500], seed=4),
    dict(id="rect_700x1900", matrix=["rect", 700, 1900, 5], seed=5),
    dict(id="rsym_2500", matrix=["random_sym", 2500, 9, 6], seed=6, ssor=[(1.0, 1), (1.5, 2)], jacobi=[(1.0, 1)]),
    dict(id="tendigit_500", matrix=["tendigit", 500], seed=7, ssor=[(1.0, 1)]),
    dict(id="one_row", matrix=["poisson1d", 1], seed=8),
]
