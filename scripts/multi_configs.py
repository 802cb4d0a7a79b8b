"""BASELINE configs 3 and 4 on N GPUs driven by ONE process through the plug-in API
(krylov.* with PSP_GPUS=N): python scripts/multi_configs.py N [tag]  -> gpurun_out/multi_configs_<tag>.jsonl"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
N = int(sys.argv[1]); tag = sys.argv[2] if len(sys.argv) > 2 else "r2"
os.environ["PSP_GPUS"] = str(N)
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
from pysparse_b200.sparse import spmatrix
from pysparse_b200.itsolvers import krylov
from pysparse_b200.precon import precon
P.set_option("gpus", N)
out = open(os.path.join("gpurun_out", f"multi_configs_{tag}.jsonl"), "a")


def run(name, solver, A, K, b, tol, maxit, extra=(), opts=()):
    for k, v in opts:
        P.set_option(k, v)
    best = None
    for rep in range(3):
        x = np.zeros(len(b))
        t0 = time.perf_counter()
        r = getattr(krylov, solver)(A, b, x, tol, maxit, K, *extra)
        dt = time.perf_counter() - t0
        ms = P.last_device_ms()
        if best is None or ms < best[1]:
            best = (r, ms, dt)
    for k, v in opts:
        P.set_option(k, -1)
    r, ms, dt = best
    y = np.empty(len(b)); A.matvec(x, y)
    rec = {"config": name, "gpus": N, "solver": solver, "options": dict(opts), "info": r[0], "iter": r[1], "relres": r[2],
           "device_ms": ms, "e2e_ms": dt * 1e3, "iters_per_s": r[1] / ms * 1e3 if ms > 0 else None, "us_per_iter": ms * 1e3 / max(r[1], 1),
           "true_relres": float(np.linalg.norm(b - y) / np.linalg.norm(b))}
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n"); out.flush()


mat = g.convdiff3d(200); n = mat["n"]
val, col, ind = mat["csr"]
A = spmatrix.csr_from_arrays(n, n, ind, col, val); K = precon.jacobi(A)
brand = np.random.default_rng(1).standard_normal(n)
run("config 4 (conv-diff 200^3, b = randn seed 1)", "bicgstab", A, K, brand, 1e-8, 5000, opts=(("bicg_persist", 0),))
run("config 4 (conv-diff 200^3, b = randn seed 1)", "bicgstab", A, K, brand, 1e-8, 5000, opts=(("bicg_persist", 1),))
run("config 4 (conv-diff 200^3, b = 1)", "gmres", A, K, np.ones(n), 1e-8, 5000, extra=(30,), opts=(("gmres_persist", 0),))
run("config 4 (conv-diff 200^3, b = 1)", "gmres", A, K, np.ones(n), 1e-8, 5000, extra=(30,), opts=(("gmres_persist", 1),))
del A, K
mat = g.poisson3d(256); n = mat["n"]
val, diag, col, ind = mat["sss"]
S = spmatrix.sss_from_arrays(n, ind, col, val, diag); K = precon.jacobi(S)
run("config 3 (Poisson 256^3, sss_mat)", "pcg", S, K, np.ones(n), 1e-10, 5000)
