"""Run the five BASELINE.json configurations on one GPU (run under gpurun) and write one JSON
record per measurement to gpurun_out/configs_<tag>.jsonl.   usage: run_configs.py [tag] [cfg,cfg,...]

cfg names: c1 (2-D n=300 PCG+Jacobi), c2 (SpMV sweep 128^3..512^3 csr/sss), c3 (256^3 PCG+Jacobi, 1 GPU leg),
c4 (conv-diff 200^3 gmres(30) / bicgstab), c5a (2-D 4096^2 PCG+SSOR), c5b (power-law CSR SpMV)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P  # noqa: E402
from pysparse_b200.tools import gallery as g  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
want = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c1", "c2", "c3", "c4", "c5a", "c5b"]
big = "--big" in sys.argv
PEAK = 6549.4
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
out = open(f"gpurun_out/configs_{tag}.jsonl", "a")


def emit(**kw):
    print(json.dumps(kw), flush=True)
    out.write(json.dumps(kw) + "\n")
    out.flush()


def time_spmv(A, nx, ny, reps=50, transp=False):
    rng = np.random.default_rng(0)
    dx = P.DeviceVector(nx, rng.standard_normal(nx)); dy = P.DeviceVector(ny)
    for _ in range(5):
        A.matvec_dev(dx, dy, transp)
    P.synchronize()
    P.event_record(2)
    for _ in range(reps):
        A.matvec_dev(dx, dy, transp)
    P.event_record(3)
    return P.event_elapsed_ms(2, 3) * 1e-3 / reps


def solve_timed(solver, A, n, tol, maxit, K=None, dim=20, reps=2, b=None):
    db = P.DeviceVector(n, np.ones(n) if b is None else b); dx = P.DeviceVector(n)
    best = None
    for _ in range(reps):
        dx.fill(0.0)
        P.synchronize()
        t0 = time.perf_counter()
        r = P.solve(solver, A, db, dx, tol, maxit, K, dim, dev=True)
        wall = time.perf_counter() - t0
        ms = P.last_device_ms()
        if best is None or ms < best[1]:
            best = (r, ms, wall)
    return best


if "c1" in want:
    mat = g.poisson2d(300)
    n = mat["n"]
    S = P.Matrix.sss(n, *mat["sss"]); K = P.Precon.jacobi(None, mat["sss"][1])
    r, ms, wall = solve_timed("pcg", S, n, 1e-10, 2000, K, reps=5)
    emit(config="c1", what="2-D Poisson n=300 sss pcg+jacobi tol 1e-10", info=r[0], iter=r[1], relres=r[2],
         device_ms=ms, wall_ms=wall * 1e3, iters_per_s=r[1] / (ms * 1e-3), us_per_iter=ms * 1e3 / r[1])
    hb = np.ones(n); hx = np.zeros(n)
    t0 = time.perf_counter(); r = P.solve("pcg", S, hb, hx, 1e-10, 2000, K); e2e = time.perf_counter() - t0
    emit(config="c1", what="same, host buffers (psp_solve)", info=r[0], iter=r[1], e2e_ms=e2e * 1e3)

if "c2" in want:
    sizes = [128, 192, 256, 384] + ([512] if big else [])
    for nn in sizes:
        t0 = time.time()
        mat = g.poisson3d(nn)
        N = mat["n"]
        gen_s = time.time() - t0
        val, col, ind = mat["csr"]
        nb = 12 * len(val) + 4 * (N + 1) + 16 * N
        A = P.Matrix.csr(N, N, val, col, ind)
        dt = time_spmv(A, N, N)
        emit(config="c2", n=nn, format="csr", us=dt * 1e6, GBps=nb / dt / 1e9, frac=nb / dt / 1e9 / PEAK,
             GFLOPs=2 * len(val) / dt / 1e9, gen_s=gen_s)
        dtt = time_spmv(A, N, N, transp=True)
        emit(config="c2", n=nn, format="csr_transp", us=dtt * 1e6, GBps=nb / dtt / 1e9, frac=nb / dtt / 1e9 / PEAK)
        del A
        sval, sdiag, scol, sind = mat["sss"]
        del mat, val, col, ind
        S = P.Matrix.sss(N, sval, sdiag, scol, sind)
        nbs = 12 * len(sval) + 4 * (N + 1) + 24 * N
        dts = time_spmv(S, N, N)
        emit(config="c2", n=nn, format="sss", us=dts * 1e6, GBps=nbs / dts / 1e9, frac=nbs / dts / 1e9 / PEAK,
             GFLOPs=(4 * len(sval) + 2 * N) / dts / 1e9)
        del S, sval, sdiag, scol, sind

if "c3" in want:
    mat = g.poisson3d(256)
    N = mat["n"]
    for fmt in ("csr", "sss"):
        A = P.Matrix.csr(N, N, *mat["csr"]) if fmt == "csr" else P.Matrix.sss(N, *mat["sss"])
        K = P.Precon.jacobi(None, np.full(N, 6.0))
        r, ms, wall = solve_timed("pcg", A, N, 1e-10, 5000, K)
        emit(config="c3", format=fmt, what="3-D Poisson 256^3 pcg+jacobi tol 1e-10", info=r[0], iter=r[1], relres=r[2],
             device_ms=ms, iters_per_s=r[1] / (ms * 1e-3))
        del A, K
    del mat

if "c4" in want:
    nn = 200
    mat = g.convdiff3d(nn)
    N = mat["n"]
    val, col, ind = mat["csr"]
    A = P.Matrix.csr(N, N, val, col, ind)
    K = P.Precon.jacobi(None, np.full(N, 6.0))
    nb = 12 * len(val) + 4 * (N + 1) + 16 * N
    # b = standard normal (seed 1): with b = ones the BiCGSTAB trajectory of the REFERENCE is
    # itself ill-behaved at >= 128^3, b = ones is still reported for gmres
    brand = np.random.default_rng(1).standard_normal(N)
    for solver, dim, b, bname in (("bicgstab", 20, brand, "randn(seed 1)"), ("gmres", 30, brand, "randn(seed 1)"),
                                  ("gmres", 30, None, "ones")):
        r, ms, wall = solve_timed(solver, A, N, 1e-8, 5000, K, dim, reps=2, b=b)
        emit(config="c4", what=f"conv-diff 200^3 {solver} + jacobi tol 1e-8, b={bname}" + (" (restart 30)" if solver == "gmres" else ""),
             info=r[0], iter=r[1], relres=r[2], device_ms=ms, iters_per_s=r[1] / (ms * 1e-3))
    del A, K, mat

if "c5a" in want:
    nn = 4096
    mat = g.poisson2d(nn, want_csr=False)
    N = mat["n"]
    S = P.Matrix.sss(N, *mat["sss"])
    t0 = time.time(); K = P.Precon.ssor(S, 1.0, 1); ana = time.time() - t0
    nf, nb_, _, _ = K.levels()
    rng = np.random.default_rng(0)
    dxv = P.DeviceVector(N, rng.standard_normal(N)); dyv = P.DeviceVector(N)
    P.check(P.lib.psp_prec_apply_dev(K.ptr, dxv.ptr, dyv.ptr)); P.synchronize()
    P.event_record(2)
    for _ in range(3):
        P.check(P.lib.psp_prec_apply_dev(K.ptr, dxv.ptr, dyv.ptr))
    P.event_record(3)
    apply_ms = P.event_elapsed_ms(2, 3) / 3
    nbytes = 2 * (12 * len(mat["sss"][0]) + 4 * (N + 1)) + 13 * 8 * N
    emit(config="c5a", what="2-D 4096^2 SSOR(1,1) apply", levels=[nf, nb_], analysis_s=ana, apply_ms=apply_ms,
         GBps=nbytes / (apply_ms * 1e-3) / 1e9)
    maxit = 10000 if big else 60
    r, ms, wall = solve_timed("pcg", S, N, 1e-10, maxit, K, reps=1)
    emit(config="c5a", what=f"2-D 4096^2 pcg+ssor(1,1) tol 1e-10 maxit {maxit}", info=r[0], iter=r[1], relres=r[2],
         device_ms=ms, iters_per_s=min(r[1], maxit) / (ms * 1e-3))
    del S, K, mat

if "c5b" in want:
    mat = g.powerlaw(1 << 22, mean_nnz=30, seed=12345)
    N = mat["n"]
    val, col, ind = mat["csr"]
    lens = np.diff(ind)
    nb = 12 * len(val) + 4 * (N + 1) + 16 * N
    A = P.Matrix.csr(N, N, val, col, ind)
    dt = time_spmv(A, N, N, reps=20)
    emit(config="c5b", what="power-law CSR 2^22 rows", nnz=int(len(val)), max_row=int(lens.max()), mean_row=float(lens.mean()),
         us=dt * 1e6, GBps=nb / dt / 1e9, frac=nb / dt / 1e9 / PEAK, GFLOPs=2 * len(val) / dt / 1e9)
    dtt = time_spmv(A, N, N, reps=20, transp=True)
    emit(config="c5b", what="power-law CSR transposed", us=dtt * 1e6, GBps=nb / dtt / 1e9, frac=nb / dtt / 1e9 / PEAK)
