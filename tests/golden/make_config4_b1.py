"""BASELINE.json config 4 as specified (conv-diff 200^3, b = 1, x0 = 0, tol 1e-8, maxit 5000,
Jacobi): what the REFERENCE's own kernels (oracle/_ref, built from /root/reference) return.
Written to tests/golden/config4_b1.json; the GPU parity test compares flags with it.
Run in the build container (minutes of CPU):  python tests/golden/make_config4_b1.py [n]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from pysparse_b200.tools import gallery as g

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
oracle.build(ref=True)
ora = oracle.Oracle("ref")
mat = g.convdiff3d(n); N = mat["n"]
A = ora.csr(N, N, *mat["csr"])
K = ora.jacobi(A, np.full(N, 6.0))
out = {"matrix": ["convdiff3d", n], "b": "ones", "x0": "zeros", "tol": 1e-8, "maxit": 5000, "precon": "jacobi",
       "checker": ora.describe(), "results": {}}
for solver, dim in (("bicgstab", 0), ("gmres", 30)):
    x = np.zeros(N)
    t0 = time.time()
    info, it, rr = ora.solve(solver, A, np.ones(N), x, 1e-8, 5000, K, dim or 20)
    dt = time.time() - t0
    r = np.empty(N); ora.matvec(A, x, r)
    true_rr = float(np.linalg.norm(np.ones(N) - r) / np.sqrt(N))
    out["results"][solver] = {"info": info, "iter": it, "relres": float(rr).hex() if np.isfinite(rr) else "nan",
                              "relres_dec": float(rr) if np.isfinite(rr) else None,
                              "true_relres": true_rr, "xnorm": float(np.linalg.norm(x)).hex(), "cpu_seconds": round(dt, 1)}
    print(solver, info, it, rr, true_rr, f"{dt:.1f}s", flush=True)
name = "config4_b1.json" if n == 200 else f"config4_b1_{n}.json"
with open(os.path.join(ROOT, "tests", "golden", name), "w") as f:
    json.dump(out, f, indent=1)
