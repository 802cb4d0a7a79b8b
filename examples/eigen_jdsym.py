"""Five smallest eigenpairs of the 2-D Poisson matrix with JDSYM, QMRS as the inner solver and a
Jacobi preconditioner (the call pattern of the reference's doc/pysparse/source/jdsym.rst example
and examples/jdsym_test.py)."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # run from the source tree

import numpy as np

from pysparse.eigen import jdsym
from pysparse.itsolvers import krylov
from pysparse.precon import precon
from pysparse.tools import poisson

n = 40
L = poisson.poisson2d_sym(n)
S = L.to_sss()
kconv, lam, Q, it, it_inner = jdsym.jdsym(S, None, precon.jacobi(L), 5, 0.0, 1e-9, 200, krylov.qmrs,
                                          jmin=8, jmax=16, clvl=1)
k = np.arange(1, n + 1)
ev = 2 - 2 * np.cos(k * np.pi / (n + 1))
exact = np.sort((ev[:, None] + ev[None, :]).ravel())[:5]
print("computed:", np.sort(lam))
print("exact:   ", exact)
for i in range(kconv):
    r = np.empty(n * n); S.matvec(np.ascontiguousarray(Q[:, i]), r)
    print("  residual %d: %.2e" % (i, np.linalg.norm(r - lam[i] * Q[:, i])))
