"""Developer probe (run under gpurun): parity of the CUDA path against the CPU
oracle on small cases plus a few timings.  Not part of the test-suite; the
real parity tests live in tests/."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import Oracle  # noqa: E402
from pysparse_b200 import _abi as P  # noqa: E402
from pysparse_b200.tools import gallery as g  # noqa: E402

ora = Oracle("ref") if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libpsp_ref.so")) else Oracle("port")
print("oracle:", ora.describe())
print("lib:", P.lib.psp_version().decode(), "devices:", P.device_count())
fails = 0


def relmax(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def diag_of_csr(n, val, col, ind):
    rows = np.repeat(np.arange(n), np.diff(ind))
    d = np.zeros(n)
    m = col == rows
    d[rows[m]] = val[m]
    return d


# ---------------------------------------------------------------- SpMV parity
rng = np.random.default_rng(0)
for name, mat in [("p2d_37", g.poisson2d(37)), ("p3d_24", g.poisson3d(24)), ("cd3d_20", g.convdiff3d(20)),
                  ("pow_5000", g.powerlaw(5000, mean_nnz=30, max_len=4000)), ("rsym_3000", g.random_sym(3000))]:
    n = mat["n"]
    shape = mat.get("shape", (n, n))
    val, col, ind = mat["csr"]
    for variant in (1, 2, 3, 4):
        P.set_option("spmv_variant", variant)
        A = P.Matrix.csr(shape[0], shape[1], val, col, ind)
        Ao = ora.csr(shape[0], shape[1], val, col, ind)
        x = rng.standard_normal(shape[1]); y = np.full(shape[0], np.nan); yo = np.empty(shape[0])
        A.matvec(x, y); ora.matvec(Ao, x, yo)
        e = relmax(y, yo); bit = np.array_equal(y, yo)
        xt = rng.standard_normal(shape[0]); yt = np.full(shape[1], np.nan); yto = np.empty(shape[1])
        A.matvec(xt, yt, transp=True); ora.matvec(Ao, xt, yto, transp=True)
        et = relmax(yt, yto); bitt = np.array_equal(yt, yto)
        ok = e <= 1e-13 and et <= 1e-13
        fails += not ok
        print(f"spmv csr {name} v{variant}: rel={e:.2e} bit={bit}  transp rel={et:.2e} bit={bitt} {'OK' if ok else 'FAIL'}")
    if "sss" in mat:
        S = P.Matrix.sss(n, *mat["sss"]); So = ora.sss(n, *mat["sss"])
        x = rng.standard_normal(n); y = np.full(n, np.nan); yo = np.zeros(n)
        S.matvec(x, y); ora.matvec(So, x, yo)
        e = relmax(y, yo)
        ok = e <= 1e-13
        fails += not ok
        print(f"spmv sss {name}: rel={e:.2e} bit={np.array_equal(y, yo)} {'OK' if ok else 'FAIL'}")
        # strided views
        xs = np.zeros(3 * n)[::3]; xs[:] = x; ys = np.zeros(2 * n)[::2]
        S.matvec(xs, ys)
        print("   strided ok:", np.array_equal(ys, y))
P.set_option("spmv_variant", 0)

# ---------------------------------------------------------------- solver parity
cases = [
    ("p2d100", g.poisson2d(100), "sss", "pcg", 1e-8, 2000, None, {}),
    ("p2d100", g.poisson2d(100), "sss", "pcg", 1e-8, 2000, "jacobi", {}),
    ("p2d100", g.poisson2d(100), "csr", "minres", 1e-8, 2000, "jacobi", {}),
    ("p2d100", g.poisson2d(100), "csr", "minres", 1e-8, 2000, None, {}),
    ("p2d100", g.poisson2d(100), "csr", "qmrs", 1e-8, 2000, "jacobi", {}),
    ("p2d100", g.poisson2d(100), "csr", "qmrs", 1e-8, 2000, None, {}),
    ("p2d100", g.poisson2d(100), "csr", "cgs", 1e-8, 2000, "jacobi", {}),
    ("p2d100", g.poisson2d(100), "csr", "cgs", 1e-8, 2000, None, {}),
    ("p2d100", g.poisson2d(100), "csr", "bicgstab", 1e-8, 2000, "jacobi", {}),
    ("p2d100", g.poisson2d(100), "csr", "bicgstab", 1e-8, 2000, None, {}),
    ("p2d100", g.poisson2d(100), "csr", "gmres", 1e-8, 2000, "jacobi", {}),
    ("p2d100", g.poisson2d(100), "csr", "gmres", 1e-8, 2000, None, {}),
    ("p2d100", g.poisson2d(100), "sss", "pcg", 1e-8, 2000, "ssor", {}),
    ("p2d100", g.poisson2d(100), "sss", "pcg", 1e-8, 2000, "ssor", dict(omega=1.9)),
    ("p2d200", g.poisson2d(200), "sss", "pcg", 1e-8, 2000, "ssor", dict(omega=1.5, steps=2)),
    ("p2d300", g.poisson2d(300), "sss", "pcg", 1e-10, 2000, "jacobi", {}),
    ("p3d32", g.poisson3d(32), "csr", "pcg", 1e-10, 2000, "jacobi", {}),
    ("cd32", g.convdiff3d(32), "csr", "bicgstab", 1e-8, 2000, "jacobi", {}),
    ("cd32", g.convdiff3d(32), "csr", "gmres", 1e-8, 2000, "jacobi", {}),
    ("p3d32", g.poisson3d(32), "csr", "pcg", 1e-10, 2000, "jacobi2", dict(steps=2)),
    ("p2d100", g.poisson2d(100), "sss", "pcg", 1e-8, 50, None, {}),     # maxit exhausted
]
for name, mat, fmt, solver, tol, maxit, prec, pk in cases:
    n = mat["n"]
    dim = 30 if solver == "gmres" else 20
    if fmt == "csr":
        A = P.Matrix.csr(n, n, *mat["csr"]); Ao = ora.csr(n, n, *mat["csr"]); d = diag_of_csr(n, *mat["csr"])
    else:
        A = P.Matrix.sss(n, *mat["sss"]); Ao = ora.sss(n, *mat["sss"]); d = mat["sss"][1]
    K = Ko = None
    if prec in ("jacobi", "jacobi2"):
        K = P.Precon.jacobi(A, d, **pk); Ko = ora.jacobi(Ao, d, **pk)
    elif prec == "ssor":
        K = P.Precon.ssor(A, **pk); Ko = ora.ssor(Ao, **pk)
    b = np.ones(n)
    x = np.zeros(n); xo = np.zeros(n)
    t0 = time.time(); r = P.solve(solver, A, b, x, tol, maxit, K, dim); dt = time.time() - t0
    ro = ora.solve(solver, Ao, b, xo, tol, maxit, Ko, dim)
    ok = (r[0] == ro[0]) and abs(r[1] - ro[1]) <= max(1, 0.02 * ro[1]) and \
        (abs(r[2] - ro[2]) <= 1e-10 or (np.isnan(r[2]) and np.isnan(ro[2])))
    ex = relmax(x, xo)
    fails += not ok
    print(f"{name:7s} {fmt} {solver:8s} K={str(prec):8s}{str(pk):28s} gpu={r} ref={ro} xerr={ex:.1e} "
          f"{dt*1e3:.1f}ms dev={P.last_device_ms():.2f}ms {'OK' if ok else 'FAIL'}")

# precon apply parity
for om, st in [(1.0, 1), (1.0, 3), (1.3, 1), (1.7, 2)]:
    mat = g.random_sym(2000, seed=3)
    n = mat["n"]
    S = P.Matrix.sss(n, *mat["sss"]); So = ora.sss(n, *mat["sss"])
    K = P.Precon.ssor(S, om, st); Ko = ora.ssor(So, om, st)
    x = rng.standard_normal(n); y = np.zeros(n); yo = np.zeros(n)
    K.precon(x, y); ora.precon(Ko, x, yo)
    print(f"ssor apply omega={om} steps={st}: rel={relmax(y, yo):.2e} bit={np.array_equal(y, yo)} levels={K.levels()[:2]}")
    fails += not relmax(y, yo) <= 1e-13

# ---------------------------------------------------------------- timings
if "--time" in sys.argv:
    for nn in (128, 256):
        mat = g.poisson3d(nn)
        n = mat["n"]
        val, col, ind = mat["csr"]
        bytes_csr = 12 * len(val) + 4 * (n + 1) + 16 * n
        for variant in (1, 2):
            P.set_option("spmv_variant", variant)
            A = P.Matrix.csr(n, n, val, col, ind)
            dx = P.DeviceVector(n, rng.standard_normal(n)); dy = P.DeviceVector(n)
            for _ in range(5):
                A.matvec_dev(dx, dy)
            P.synchronize()
            t0 = time.perf_counter()
            reps = 50
            for _ in range(reps):
                A.matvec_dev(dx, dy)
            P.synchronize()
            dt = (time.perf_counter() - t0) / reps
            print(f"spmv csr {nn}^3 variant {variant}: {dt*1e6:.1f} us  {bytes_csr/dt/1e9:.0f} GB/s (algorithmic)")
            del A
        P.set_option("spmv_variant", 0)
        S = P.Matrix.sss(n, *mat["sss"])
        K = P.Precon.jacobi(S, mat["sss"][1])
        db = P.DeviceVector(n, np.ones(n)); dxx = P.DeviceVector(n, np.zeros(n))
        for rep in range(2):
            dxx.fill(0.0)
            t0 = time.perf_counter()
            r = P.solve("pcg", S, db, dxx, 1e-10, 5000, K, dev=True)
            dt = time.perf_counter() - t0
            print(f"pcg+jacobi sss {nn}^3: {r} wall={dt*1e3:.1f} ms dev={P.last_device_ms():.1f} ms -> {r[1]/(P.last_device_ms()*1e-3):.0f} it/s")
print("FAILS:", fails)
sys.exit(1 if fails else 0)
