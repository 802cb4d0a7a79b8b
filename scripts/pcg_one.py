"""usage: pcg_one.py N [persist] [reps] [maxit] -- one PCG+Jacobi solve on 3-D Poisson N^3 (for ncu captures)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
nn = int(sys.argv[1]); persist = int(sys.argv[2]) if len(sys.argv) > 2 else -1; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
maxit = int(sys.argv[4]) if len(sys.argv) > 4 else 5000
P.set_option("pcg_persist", persist)
mat = g.poisson3d(nn, want_sss=False); n = mat["n"]
A = P.Matrix.csr(n, n, *mat["csr"]); K = P.Precon.jacobi(None, np.full(n, 6.0))
db = P.DeviceVector(n, np.ones(n)); dx = P.DeviceVector(n)
for _ in range(reps):
    dx.fill(0.0)
    r = P.solve("pcg", A, db, dx, 1e-10, maxit, K, dev=True)
    print(nn, "persist", persist, r, f"{P.last_device_ms():.2f} ms {P.last_device_ms()*1e3/r[1]:.1f} us/iter", flush=True)
