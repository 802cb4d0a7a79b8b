import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
import oracle
ora = oracle.Oracle("ref") if oracle.have_ref() else oracle.Oracle("port")
def run(mat, fmt, tol, maxit, jac, b=None, x0=None, persist=1):
    n = mat["n"]
    P.set_option("pcg_persist", persist)
    A = P.Matrix.csr(n, n, *mat["csr"]) if fmt == "csr" else P.Matrix.sss(n, *mat["sss"])
    K = P.Precon.jacobi(None, np.full(n, 4.0 if "2d" in mat.get("kind", "") else 6.0) if jac is True else jac) if jac is not None else None
    db = P.DeviceVector(n, np.ones(n) if b is None else b); dx = P.DeviceVector(n, np.zeros(n) if x0 is None else x0)
    best = None
    for _ in range(3):
        dx.upload(np.zeros(n) if x0 is None else x0)
        r = P.solve("pcg", A, db, dx, tol, maxit, K, dev=True)
        ms = P.last_device_ms()
        best = ms if best is None else min(best, ms)
    return r, best, dx.download()
for name, mat, fmt, tol, maxit in [("p2d100", g.poisson2d(100), "sss", 1e-8, 2000), ("p2d300", g.poisson2d(300), "sss", 1e-10, 2000),
                                   ("p3d64", g.poisson3d(64), "csr", 1e-10, 2000), ("p2d100 maxit", g.poisson2d(100), "sss", 1e-8, 50),
                                   ("p3d128", g.poisson3d(128), "csr", 1e-10, 2000), ("rsym", g.random_sym(50000, 9, 3), "sss", 1e-12, 500)]:
    n = mat["n"]
    d = mat["sss"][1] if "sss" in mat else np.full(n, 6.0)
    for jac in (None, d):
        r1, ms1, x1 = run(mat, fmt, tol, maxit, jac, persist=1)
        r0, ms0, x0 = run(mat, fmt, tol, maxit, jac, persist=0)
        Ao = ora.csr(n, n, *mat["csr"]) if fmt == "csr" else ora.sss(n, *mat["sss"])
        Ko = ora.jacobi(Ao, d) if jac is not None else None
        xo = np.zeros(n); ro = ora.solve("pcg", Ao, np.ones(n), xo, tol, maxit, Ko)
        ok = r1[:2] == ro[:2] and abs(r1[2] - ro[2]) <= 1e-10 * max(1, abs(ro[2])) and np.max(np.abs(x1 - xo)) <= 1e-7 * np.max(np.abs(xo))
        print(f"{name:12s} jac={jac is not None} persist={r1} {ms1:.2f}ms ({ms1*1e3/max(1,r1[1]):.1f} us/it) multi={r0} {ms0:.2f}ms ({ms0*1e3/max(1,r0[1]):.1f} us/it) ref={ro} {'OK' if ok else 'FAIL'}", flush=True)
# zero rhs and good x0
mat = g.poisson2d(50); n = mat["n"]
r, _, x = run(mat, "sss", 1e-8, 100, None, b=np.zeros(n), x0=np.full(n, 0.5)); print("zero rhs:", r, float(np.abs(x).max()))
xs = np.zeros(n); ora.solve("pcg", ora.sss(n, *mat["sss"]), np.ones(n), xs, 1e-12, 1000)
r, _, x = run(mat, "sss", 1e-6, 100, None, x0=xs); print("good x0:", r)
