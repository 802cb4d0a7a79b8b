"""Bring-up of the diagonal-run matrix format (spmv_format): parity of the run kernels against
the CSR row-block kernels and the oracle, then timings.  usage: runs_probe.py [--time] [sizes...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
import oracle

ora = oracle.Oracle("ref") if oracle.have_ref() else oracle.Oracle("port")
ok = True


def mats(fmt, mat):
    P.set_option("spmv_format", fmt)
    n = mat["n"]
    A = P.Matrix.csr(n, n, *mat["csr"])
    S = P.Matrix.sss(n, *mat["sss"]) if "sss" in mat else None
    P.set_option("spmv_format", -1)
    return A, S


cases = [("poisson1d 1", g.poisson1d(1)), ("poisson1d 700", g.poisson1d(700)), ("poisson2d 37", g.poisson2d(37)),
         ("poisson2d 300", g.poisson2d(300)), ("poisson3d 13", g.poisson3d(13)), ("poisson3d 40", g.poisson3d(40)),
         ("convdiff3d 23", g.convdiff3d(23)), ("random_sym 3000", g.random_sym(3000, 6, seed=4))]
for name, mat in cases:
    n = mat["n"]
    x = np.random.default_rng(7).standard_normal(n)
    Ao = ora.csr(n, n, *mat["csr"]); yo = np.empty(n); ora.matvec(Ao, x, yo)
    yto = np.zeros(n); ora.matvec(Ao, x, yto, True)
    res = []
    for fmt in (1, 0):
        A, S = mats(fmt, mat)
        y = np.empty(n); A.matvec(x, y); res.append(("csr", fmt, np.array_equal(y, yo), np.max(np.abs(y - yo))))
        y = np.empty(n); A.matvec(x, y, True); res.append(("csrT", fmt, np.array_equal(y, yto), np.max(np.abs(y - yto))))
        if S is not None:
            So = ora.sss(n, *mat["sss"]); ys = np.zeros(n); ora.matvec(So, x, ys)
            y = np.empty(n); S.matvec(x, y); res.append(("sss", fmt, np.array_equal(y, ys), np.max(np.abs(y - ys))))
            xs = np.zeros(n); r = P.solve("pcg", S, np.ones(n), xs, 1e-10, 3000, P.Precon.jacobi(None, mat["sss"][1]))
            xo = np.zeros(n); ro = ora.solve("pcg", So, np.ones(n), xo, 1e-10, 3000, ora.jacobi(So, mat["sss"][1]))
            res.append((f"pcg {r} vs {ro}", fmt, r[:2] == ro[:2], np.max(np.abs(xs - xo))))
    bad = [r for r in res if not (r[2] or r[3] <= 1e-13 * max(1.0, np.max(np.abs(yo))))]
    ok = ok and not bad
    print(name, "OK" if not bad else "FAIL", [(r[0], r[1], "bit-identical" if r[2] else f"{r[3]:.2e}") for r in res], flush=True)
print("RUNS PROBE", "PASSED" if ok else "FAILED", flush=True)

if "--time" in sys.argv:
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [128, 256]
    for nn in sizes:
        mat = g.poisson3d(nn); n = mat["n"]
        nnz, nl = len(mat["csr"][0]), len(mat["sss"][0])
        bcsr = 12 * nnz + 4 * (n + 1) + 16 * n
        bsss = 12 * nl + 4 * (n + 1) + 24 * n
        dx = P.DeviceVector(n, np.random.default_rng(0).standard_normal(n)); dy = P.DeviceVector(n)
        for fmt in (1, 0):
            A, S = mats(fmt, mat)
            for nm, M, nb, tr in (("csr", A, bcsr, False), ("csrT", A, bcsr, True), ("sss", S, bsss, False)):
                for _ in range(5):
                    M.matvec_dev(dx, dy, tr)
                P.event_record(0)
                for _ in range(50):
                    M.matvec_dev(dx, dy, tr)
                P.event_record(1)
                us = P.event_elapsed_ms(0, 1) * 1e3 / 50
                print(f"spmv {nm:4s} {nn}^3 format {fmt}: {us:8.1f} us  {nb / us / 1e3:7.0f} GB/s algorithmic  ({nb / us / 1e3 / 6549.4:.2f} of measured peak)", flush=True)
            db = P.DeviceVector(n, np.ones(n)); K = P.Precon.jacobi(None, np.full(n, 6.0))
            for nm, M in (("csr", A), ("sss", S)):
                for persist in (-1, 0):
                    P.set_option("pcg_persist", persist)
                    dxx = P.DeviceVector(n); dxx.fill(0.0)
                    r = P.solve("pcg", M, db, dxx, 1e-10, 5000, K, dev=True)
                    ms = P.last_device_ms()
                    print(f"pcg+jacobi {nm} {nn}^3 format {fmt} persist {persist}: {r} {ms:.1f} ms {ms * 1e3 / r[1]:.1f} us/iter {r[1] / ms * 1e3:.0f} it/s", flush=True)
                P.set_option("pcg_persist", -1)
            del A, S
sys.exit(0 if ok else 1)
