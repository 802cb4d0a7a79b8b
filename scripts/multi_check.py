"""N GPUs behind the unchanged single-process API (PSP_GPUS / option "gpus"): parity of the
sharded krylov.pcg / minres / qmrs / bicgstab / cgs / gmres against the one-GPU solve and the oracle, then timings.
    python scripts/multi_check.py N [--time] [grid]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
from pysparse_b200.sparse import spmatrix
from pysparse_b200.itsolvers import krylov
from pysparse_b200.precon import precon
import oracle

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ora = oracle.Oracle("ref") if oracle.have_ref() else oracle.Oracle("port")
ok = True


def objs(gpus, mat, fmt):
    P.set_option("gpus", gpus)
    n = mat["n"]
    if fmt == "csr":
        val, col, ind = mat["csr"]
        A = spmatrix.csr_from_arrays(n, n, ind, col, val)
    else:
        val, diag, col, ind = mat["sss"]
        A = spmatrix.sss_from_arrays(n, ind, col, val, diag)
    _ = A.shape
    y = np.empty(n); A.matvec(np.ones(n), y)          # forces the device upload while "gpus" has this value
    K = precon.jacobi(A)
    P.set_option("gpus", 1)
    return A, K


cases = [("poisson3d 20", g.poisson3d(20), "csr", 1e-10), ("poisson3d 20", g.poisson3d(20), "sss", 1e-10),
         ("poisson2d 90", g.poisson2d(90), "sss", 1e-10), ("convdiff3d 24", g.convdiff3d(24), "csr", 1e-8),
         ("random_sym 4000", g.random_sym(4000, 6, seed=4), "csr", 1e-10)]
for name, mat, fmt, tol in cases:
    n = mat["n"]
    b = np.random.default_rng(1).standard_normal(n)
    solvers = ["pcg", "minres", "qmrs", "bicgstab", "cgs", "gmres"] if "convdiff" not in name else ["bicgstab", "gmres", "cgs"]
    for solver in solvers:
        for withK in (True, False):
            res = []
            for gpus in (N, 1):
                A, K = objs(gpus, mat, fmt)
                P.set_option("gpus", gpus)
                x = np.zeros(n)
                args = (A, b, x, tol, 600) + ((K,) if withK else ()) + ((None, 25) if solver == "gmres" and not withK else ((25,) if solver == "gmres" else ()))
                r = getattr(krylov, solver)(*args)
                P.set_option("gpus", 1)
                res.append((r, x))
            (rm, xm), (r1, x1) = res
            tol_it = 0.1 if (solver in ("bicgstab", "cgs") and "poisson" in name) else 0.02   # BiCGSTAB / CGS on Poisson: rounding-chaotic
            good = rm[0] == r1[0] and abs(rm[1] - r1[1]) <= max(1, tol_it * r1[1]) and np.max(np.abs(xm - x1)) <= 1e-6 * max(1.0, np.max(np.abs(x1)))
            ok = ok and good
            print(f"{name:16s} {fmt} {solver:8s} K={withK}: {N} GPUs {rm} | 1 GPU {r1} | max|dx| {np.max(np.abs(xm - x1)):.2e} {'OK' if good else 'FAIL'}", flush=True)
print("MULTI CHECK", "PASSED" if ok else "FAILED", flush=True)

if "--time" in sys.argv:
    nn = int(sys.argv[-1]) if sys.argv[-1].isdigit() and sys.argv[-1] != sys.argv[1] else 256
    mat = g.poisson3d(nn); n = mat["n"]
    b = np.ones(n)
    for fmt in ("csr", "sss"):
        for gpus in (N, 1):
            A, K = objs(gpus, mat, fmt)
            P.set_option("gpus", gpus)
            for rep in range(3):
                x = np.zeros(n)
                t0 = time.perf_counter()
                r = krylov.pcg(A, b, x, 1e-10, 5000, K)
                dt = time.perf_counter() - t0
            P.set_option("gpus", 1)
            print(f"krylov.pcg {fmt} {nn}^3 on {gpus} GPU(s): {r} e2e {dt*1e3:.1f} ms = {r[1]/dt:.0f} it/s; device {P.last_device_ms():.1f} ms = {r[1]/P.last_device_ms()*1e3:.0f} it/s", flush=True)
            del A, K
sys.exit(0 if ok else 1)
