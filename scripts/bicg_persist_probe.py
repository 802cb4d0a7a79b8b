"""Round-2 bring-up of the single-launch BiCGSTAB kernel (option bicg_persist=1, OFF by default,
not yet run on a GPU): parity against the per-phase kernel sequence and the oracle on small
cases, then the per-iteration time at the config-4 size.  Run on one GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
import oracle

ora = oracle.Oracle("ref") if oracle.have_ref() else oracle.Oracle("port")


def run(mat, tol, maxit, jac, b, persist):
    n = mat["n"]
    P.set_option("bicg_persist", persist)
    try:
        A = P.Matrix.csr(n, n, *mat["csr"])
        K = P.Precon.jacobi(None, jac) if jac is not None else None
        db = P.DeviceVector(n, b); dx = P.DeviceVector(n, np.zeros(n))
        P.launch_count(reset=True)
        r = P.solve("bicgstab", A, db, dx, tol, maxit, K, dev=True)
        return r, P.last_device_ms(), dx.download(), P.launch_count()
    finally:
        P.set_option("bicg_persist", 0)


ok = True
cases = [("convdiff3d 16", g.convdiff3d(16), 1e-8, 500), ("convdiff3d 32", g.convdiff3d(32), 1e-8, 500),
         ("poisson3d 24", g.poisson3d(24), 1e-10, 500), ("convdiff3d 32 maxit", g.convdiff3d(32), 1e-12, 7)]
for name, mat, tol, maxit in cases:
    n = mat["n"]
    rng = np.random.default_rng(1)
    for bname, b in (("randn", rng.standard_normal(n)), ("zeros", np.zeros(n))):
        for jac in (None, np.full(n, 6.0)):
            r1, ms1, x1, l1 = run(mat, tol, maxit, jac, b, 1)
            r0, ms0, x0, l0 = run(mat, tol, maxit, jac, b, 0)
            Ao = ora.csr(n, n, *mat["csr"]); Ko = ora.jacobi(Ao, jac) if jac is not None else None
            xo = np.zeros(n); ro = ora.solve("bicgstab", Ao, b, xo, tol, maxit, Ko)
            good = (r1 == r0 or (np.isnan(r1[2]) and r1[:2] == r0[:2])) and np.array_equal(x1, x0) and (name.startswith("poisson") or r1[0] == ro[0] and abs(r1[1] - ro[1]) <= max(1, 0.02 * ro[1]) and l1 == 1 and
                    (not np.isfinite(ro[2]) or abs(r1[2] - ro[2]) <= max(1e-10, 1e-6 * abs(ro[2]))) and
                    np.max(np.abs(x1 - xo)) <= 1e-6 * max(1.0, np.max(np.abs(xo))))
            ok = ok and good
            print(f"{name:22s} b={bname:5s} jac={jac is not None}: persist={r1} {ms1:.2f} ms ({l1} launch) "
                  f"per-phase={r0} {ms0:.2f} ms oracle={ro} {'OK' if good else 'FAIL'}", flush=True)
print("BICG PERSIST", "PASSED" if ok else "FAILED", flush=True)
if ok and "--time" in sys.argv:
    mat = g.convdiff3d(200); n = mat["n"]
    b = np.random.default_rng(1).standard_normal(n)
    for persist in (1, 0):
        r, ms, _, l = run(mat, 1e-8, 2000, np.full(n, 6.0), b, persist)
        print(f"conv-diff 200^3 bicgstab+jacobi persist={persist}: {r} {ms:.1f} ms {ms * 1e3 / r[1]:.1f} us/iter ({l} launches)", flush=True)
sys.exit(0 if ok else 1)
