"""JDSYM with the correction equation on the device vs through the host-callback protocol.
    python scripts/jdsym_ab.py [grid=96] [kmax=5]
3-D Poisson (sss_mat), Jacobi preconditioner, krylov.qmrs as the inner solver."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200.tools import gallery as g
from pysparse.sparse import spmatrix
from pysparse.itsolvers import krylov
from pysparse.precon import precon
from pysparse.eigen import jdsym

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 96
kmax = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mat = g.poisson3d(grid)
n = mat["n"]
val, diag, col, ind = mat["sss"]
A = spmatrix.sss_from_arrays(n, ind, col, val, diag)
K = precon.jacobi(A)
for device in (True, False, True):
    jdsym.DEVICE_CORREQ = device
    t0 = time.perf_counter()
    kconv, lam, Q, it, it_inner = jdsym.jdsym(A, None, K, kmax, 0.0, 1e-8, 200, krylov.qmrs, jmin=8, jmax=16, linitmax=100)
    dt = time.perf_counter() - t0
    print(f"poisson3d {grid}^3 (n={n}) kmax={kmax} correction equation on the {'device' if device else 'host callbacks'}: "
          f"kconv={kconv} outer={it} inner={it_inner} {dt:.2f} s  ({1e3 * dt / max(it_inner, 1):.2f} ms per inner iteration)  "
          f"lambda={np.array2string(np.sort(lam), precision=8)}", flush=True)
