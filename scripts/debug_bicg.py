import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
import oracle
ora = oracle.Oracle("ref")
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mat = g.convdiff3d(nn); N = mat["n"]
A = P.Matrix.csr(N, N, *mat["csr"]); K = P.Precon.jacobi(None, np.full(N, 6.0))
Ao = ora.csr(N, N, *mat["csr"]); Ko = ora.jacobi(Ao, np.full(N, 6.0))
for maxit in (1, 2, 3, 4, 6, 10, 20, 40, 60, 70):
    x = np.zeros(N); xo = np.zeros(N)
    r = P.solve("bicgstab", A, np.ones(N), x, 1e-8, maxit, K)
    ro = ora.solve("bicgstab", Ao, np.ones(N), xo, 1e-8, maxit, Ko)
    print(maxit, r, ro, "xdiff", np.max(np.abs(x - xo)) / np.max(np.abs(xo)), flush=True)
