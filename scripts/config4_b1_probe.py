"""BASELINE config 4 as specified (b = 1): what the GPU returns, next to tests/golden/config4_b1.json."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 200
gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "config4_b1.json")))["results"]
mat = g.convdiff3d(nn); n = mat["n"]
K = P.Precon.jacobi(None, np.full(n, 6.0))
db = P.DeviceVector(n, np.ones(n))
for fmt in (-1, 0):
    P.set_option("spmv_format", fmt)
    A = P.Matrix.csr(n, n, *mat["csr"])
    for solver, persist in (("bicgstab", 0), ("bicgstab", 1), ("gmres", 0)):
        P.set_option("bicg_persist", persist)
        dx = P.DeviceVector(n); dx.fill(0.0)
        r = P.solve(solver, A, db, dx, 1e-8, 5000, K, 30, dev=True)
        ms = P.last_device_ms()
        x = dx.download(); y = np.empty(n); A.matvec(x, y)
        print(f"format {fmt} {solver} persist {persist}: {r} {ms:.0f} ms true relres {np.linalg.norm(1 - y) / np.sqrt(n):.3e}  reference: {gold[solver]['info'], gold[solver]['iter'], gold[solver]['relres_dec']}", flush=True)
    P.set_option("bicg_persist", 0)
