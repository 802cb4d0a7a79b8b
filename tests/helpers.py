"""Test helpers: an Oracle-shaped facade over the CUDA C-ABI, comparison utilities."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relmax(a, b):
    d = float(np.max(np.abs(a - b))) if a.size else 0.0
    s = float(np.max(np.abs(b))) if b.size else 0.0
    return d / s if s > 0 else d


def load_kat():
    with open(os.path.join(GOLDEN, "solver_kat.json")) as f:
        recs = json.load(f)
    for r in recs:
        r["relres"] = float("nan") if r["relres"] == "nan" else float.fromhex(r["relres"])
        for k in ("xnorm", "x0", "xmid"):
            r[k] = float.fromhex(r[k])
    return recs


def load_kernels():
    return np.load(os.path.join(GOLDEN, "kernels.npz"))


class GpuLib:
    """Same factory surface as oracle.Oracle, backed by libpsp_b200.so (C-ABI)."""

    def __init__(self, abi):
        self.P = abi

    def csr(self, m, n, val, col, ind):
        return self.P.Matrix.csr(m, n, val, col, ind)

    def sss(self, n, val, diag, col, ind):
        return self.P.Matrix.sss(n, val, diag, col, ind)

    def jacobi(self, A, diag, omega=1.0, steps=1):
        return self.P.Precon.jacobi(A, diag, omega, steps)

    def ssor(self, S, omega=1.0, steps=1):
        return self.P.Precon.ssor(S, omega, steps)

    def matvec(self, A, x, y, transp=False):
        A.matvec(x, y, transp)

    def precon(self, K, x, y):
        K.precon(x, y)

    def solve(self, solver, A, b, x, tol, maxit, K=None, dim=20):
        return self.P.solve(solver, A, b, x, tol, maxit, K, dim)


def start_vectors(case, n):
    b = np.zeros(n) if case.get("b") == "zero" else np.ones(n)
    x = np.zeros(n) if case.get("x0", "zeros") == "zeros" else np.full(n, 0.5)
    return b, x
