"""2-D Poisson problem solved with PCG on the three matrix formats, with and without SSOR
(the flow of the reference's examples/poisson_test.py:62-135)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # run from the source tree

import sys
import time

import numpy as np

from pysparse.itsolvers import krylov
from pysparse.precon import precon
from pysparse.tools import poisson

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
tol, maxit = 1e-8, 4 * n

t0 = time.perf_counter()
L = poisson.poisson2d_sym_blk(n)
print("assembly (ll_mat_sym, %d x %d, %d entries): %.2f s" % (L.shape + (L.nnz, time.perf_counter() - t0)))
A, S = L.to_csr(), L.to_sss()
b = np.ones(n * n)

for name, mat in (("sss_mat", S), ("csr_mat", A), ("ll_mat", L)):
    x = np.zeros(n * n)
    t0 = time.perf_counter()
    info, it, relres = krylov.pcg(mat, b, x, tol, maxit)
    dt = time.perf_counter() - t0
    r = np.empty(n * n); mat.matvec(x, r)
    print("%-8s pcg          info=%d iter=%4d relres=%.3e  ||b-Ax||/||b||=%.3e  %.3f s" %
          (name, info, it, relres, np.linalg.norm(b - r) / np.linalg.norm(b), dt))

for omega in (1.0, 1.9):
    x = np.zeros(n * n)
    t0 = time.perf_counter()
    info, it, relres = krylov.pcg(S, b, x, tol, maxit, precon.ssor(S, omega))
    print("sss_mat  pcg+ssor(%.1f) info=%d iter=%4d relres=%.3e  %.3f s" % (omega, info, it, relres, time.perf_counter() - t0))

x = np.zeros(n * n)
info, it, relres = krylov.pcg(S, b, x, tol, maxit, precon.jacobi(L))
print("sss_mat  pcg+jacobi   info=%d iter=%4d relres=%.3e" % (info, it, relres))
