#!/usr/bin/env python
"""bench.py -- headline benchmark of pysparse_b200 (contract: see the task brief).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
3-D 7-point Poisson 256^3 (16.7 M rows, 117 M non-zeros, fp64), PCG + Jacobi,
b = 1, x0 = 0, tol = 1e-10.  One "step" = one complete solve.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid 256]

* value   : PCG iterations per second, vectors and matrix resident in HBM
            (psp_solve_dev), timed with CUDA events on the library stream;
* e2e     : the same solve through the host-buffer API call (krylov.pcg /
            psp_solve: b, x0 copied host->device from pinned memory, x copied
            back) inside the timed region;
* roofline: the fused SpMV+dot kernel (dominant kernel), sampled live with CUDA
            event pairs inside the solver loop; algorithmic bytes
            12 nnz + 4 (N+1) + 16 N per launch;
* cpu_baseline: the reference's own C kernels (oracle/_ref) on one host core,
            first CPU_ITERS iterations of the same solve.

N > 1 (torchrun): the same system row-sharded over N GPUs (strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CPU_ITERS = 25          # bounded CPU sample: ~12 s of serial reference PCG at 256^3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="grid points per dimension (256 = headline config)")
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--maxit", type=int, default=5000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sweep", action="store_true", help="also report the SpMV sweep (csr/sss, 128^3..)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.path = index, None, f"/tmp/psp_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(n):
    from pysparse_b200.tools import gallery as g
    mat = g.poisson3d(n, want_csr=True, want_sss=False)
    val, col, ind = mat["csr"]
    N = mat["n"]
    diag = np.full(N, 6.0)
    return N, val, col, ind, diag


def alg_bytes_spmv(N, nnz):
    return 12 * nnz + 4 * (N + 1) + 8 * N + 8 * N


def alg_bytes_pcg_iter(N, nnz):
    return 12 * nnz + 4 * (N + 1) + 104 * N


# --------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref when built, else the
    plain-C restatement) on the host cores: serial, as the reference is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    kind = "reference" if oracle.have_ref() else "port"
    ora = oracle.Oracle("ref" if kind == "reference" else "port")
    N, val, col, ind, diag = build_problem(args.grid)
    A = ora.csr(N, N, val, col, ind)
    K = ora.jacobi(A, diag)
    b = np.ones(N)
    times, iters = [], 0
    for step in range(args.warmup + args.steps):
        x = np.zeros(N)
        t0 = time.perf_counter()
        info, it, rr = ora.solve("pcg", A, b, x, args.tol, CPU_ITERS, K)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt); iters += CPU_ITERS
        if args.warmup + args.steps > 2 and step == 0 and dt > 60:
            break
    if not times:
        times, iters = [dt], CPU_ITERS
    value = iters / sum(times)
    sample = f"first {CPU_ITERS} PCG iterations (maxit={CPU_ITERS}) of the same {args.grid}^3 solve per step"
    line = {
        "impl": "reference", "metric": "pcg_iters_per_s", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3-D 7-point Poisson {args.grid}^3, PCG + Jacobi, tol {args.tol:g}, b=1, x0=0 (CPU: capped sample)",
                   "rows": N, "nnz": int(len(val)), "format": "csr"},
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm, single GPU
def run_single(args):
    from pysparse_b200 import _abi as P
    if P.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (pysparse_b200 has no CPU fallback)")
    peak, peak_src = peaks()
    N, val, col, ind, diag = build_problem(args.grid)
    nnz = int(len(val))
    A = P.Matrix.csr(N, N, val, col, ind)
    K = P.Precon.jacobi(A, diag)
    db = P.DeviceVector(N, np.ones(N))
    dx = P.DeviceVector(N)
    hb = P.pinned_empty(N); hb[:] = 1.0
    hx = P.pinned_empty(N)

    def dev_step():
        dx.fill(0.0)
        return P.solve("pcg", A, db, dx, args.tol, args.maxit, K, dev=True)

    for _ in range(args.warmup):
        res = dev_step()
    P.synchronize()

    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    clocks.start()
    P.spmv_timing(8)                     # sample every 8th SpMV launch from here on
    P.launch_count(reset=True)
    P.synchronize()
    P.event_record(0)
    iters = 0
    for _ in range(args.steps):
        res = dev_step()
        iters += res[1]
    P.event_record(1)
    P.synchronize()
    ms = P.event_elapsed_ms(0, 1)
    launches = P.launch_count()
    n_spmv, spmv_ms = P.spmv_timing(0)
    clk = clocks.stop()

    # end to end through the host-buffer call (pinned host arrays)
    e2e_step_fn = lambda: (hx.__setitem__(slice(None), 0.0), P.solve("pcg", A, hb, hx, args.tol, args.maxit, K))[1]
    api = "psp_solve (C-ABI, host buffers)"
    try:
        from pysparse_b200.itsolvers import krylov
        from pysparse_b200.sparse import spmatrix
        from pysparse_b200.precon import precon
        Am = spmatrix.csr_from_arrays(N, N, ind, col, val)
        Km = precon.jacobi(Am)
        e2e_step_fn = lambda: (hx.__setitem__(slice(None), 0.0), krylov.pcg(Am, hb, hx, args.tol, args.maxit, Km))[1]
        api = "pysparse_b200.itsolvers.krylov.pcg(A, b, x, tol, maxit, K) with NumPy arrays"
    except (ImportError, AttributeError):
        pass
    e2e_step_fn()
    P.synchronize()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        r2 = e2e_step_fn()
        e2e_iters += r2[1]
    P.synchronize()
    e2e_s = time.perf_counter() - t0

    line = {
        "metric": "pcg_iters_per_s", "value": iters / (ms * 1e-3), "unit": "iter/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3-D 7-point Poisson {args.grid}^3, PCG + Jacobi, tol {args.tol:g}, b=1, x0=0",
                   "rows": N, "nnz": nnz, "format": "csr", "parallelism": "1 GPU",
                   "l2_policy": "inputs larger than L2 (matrix 1.47 GB + 6 vectors of 134 MB vs 126 MB L2)",
                   "result": {"info": res[0], "iter": res[1], "relres": res[2]}},
        "e2e": {"value": e2e_iters / e2e_s, "unit": "iter/s", "h2d_bytes_per_step": 16 * N,
                "d2h_bytes_per_step": 8 * N + 16, "api": api, "seconds_per_step": e2e_s / args.steps},
        "gpu_launches": launches,
        "clocks": clk,
        "iters_per_step": iters / args.steps,
        "pcg_iter_algorithmic_GBps": alg_bytes_pcg_iter(N, nnz) * iters / (ms * 1e-3) / 1e9,
    }
    def traffic_of(key):
        tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
        if os.path.exists(tp) and args.grid == 256:
            return json.load(open(tp)).get(key)
        return None

    def spmv_roofline(n_spmv, spmv_ms):
        per = spmv_ms / n_spmv * 1e-3
        ach = alg_bytes_spmv(N, nnz) / per / 1e9
        return {"bound": "hbm", "kernel": "k_spmv_tma<2,true> (CSR SpMV + fused p.q dot)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic_of("dram_bytes_per_launch"), "peak_source": peak_src, "launches_sampled": n_spmv,
                "avg_launch_us": per * 1e6, "algorithmic_bytes_per_launch": alg_bytes_spmv(N, nnz),
                "gflops": 2 * nnz / per / 1e9}

    if n_spmv:
        line["roofline"] = spmv_roofline(n_spmv, spmv_ms)
    else:
        # The solve ran as ONE persistent cooperative kernel (k_pcg_persist): that launch is
        # the dominant kernel.  Algorithmic bytes per launch = iterations x the per-iteration
        # figure of DESIGN.md (matrix once + 13 vector streams); duration = the CUDA-event
        # time of the timed region / launches (it contains nothing else but a memset).
        per = ms * 1e-3 / args.steps
        nb = alg_bytes_pcg_iter(N, nnz) * iters / args.steps
        line["roofline"] = {"bound": "hbm", "kernel": "k_pcg_persist (whole PCG + Jacobi solve in one cooperative launch)",
                            "achieved": nb / per / 1e9, "peak": peak, "unit": "GB/s", "frac": nb / per / 1e9 / peak,
                            "traffic": traffic_of("pcg_persist_dram_bytes_per_launch"), "peak_source": peak_src,
                            "launches_sampled": args.steps, "avg_launch_us": per * 1e6,
                            "algorithmic_bytes_per_launch": nb,
                            "algorithmic_bytes_per_iteration": alg_bytes_pcg_iter(N, nnz)}
        # for continuity: the stand-alone SpMV + dot kernel, sampled in a short solve on the
        # per-phase path (not part of the timed region)
        P.set_option("pcg_persist", 0)
        dx.fill(0.0)
        P.solve("pcg", A, db, dx, args.tol, 32, K, dev=True)      # settle on the per-phase path first
        P.spmv_timing(8)
        dx.fill(0.0)
        P.solve("pcg", A, db, dx, args.tol, 256, K, dev=True)
        n2, ms2 = P.spmv_timing(0)
        P.set_option("pcg_persist", -1)
        if n2:
            line["roofline_spmv"] = spmv_roofline(n2, ms2)
    if args.sweep:
        line["spmv_sweep"] = spmv_sweep(P)
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(args, N, val, col, ind, diag)
    print(json.dumps(line), flush=True)


def spmv_sweep(P, sizes=(128, 192, 256)):
    from pysparse_b200.tools import gallery as g
    out = []
    rng = np.random.default_rng(0)
    peak, _ = peaks()
    for n in sizes:
        mat = g.poisson3d(n)
        N = mat["n"]
        for fmt in ("csr", "sss"):
            if fmt == "csr":
                val, col, ind = mat["csr"]
                A = P.Matrix.csr(N, N, val, col, ind)
                nb = alg_bytes_spmv(N, len(val)); flops = 2 * len(val)
            else:
                val, diag, col, ind = mat["sss"]
                A = P.Matrix.sss(N, val, diag, col, ind)
                nb = 12 * len(val) + 4 * (N + 1) + 24 * N; flops = 4 * len(val) + 2 * N
            dx = P.DeviceVector(N, rng.standard_normal(N)); dy = P.DeviceVector(N)
            for _ in range(5):
                A.matvec_dev(dx, dy)
            P.synchronize()
            P.event_record(2)
            reps = 50
            for _ in range(reps):
                A.matvec_dev(dx, dy)
            P.event_record(3)
            dt = P.event_elapsed_ms(2, 3) * 1e-3 / reps
            out.append({"n": n, "format": fmt, "us": dt * 1e6, "GBps_algorithmic": nb / dt / 1e9,
                        "frac_of_peak": nb / dt / 1e9 / peak, "GFLOPs": flops / dt / 1e9})
            del A, dx, dy
    return out


def cpu_baseline(args, N, val, col, ind, diag):
    import oracle
    kind = "reference" if oracle.have_ref() else "port"
    ora = oracle.Oracle("ref" if kind == "reference" else "port")
    A = ora.csr(N, N, val, col, ind)
    K = ora.jacobi(A, diag)
    x = np.zeros(N)
    t0 = time.perf_counter()
    ora.solve("pcg", A, np.ones(N), x, args.tol, CPU_ITERS, K)
    dt = time.perf_counter() - t0
    return {"value": CPU_ITERS / dt, "unit": "iter/s", "cores": 1, "kind": kind,
            "sample": f"first {CPU_ITERS} PCG iterations (maxit={CPU_ITERS}) of the same {args.grid}^3 solve, {dt:.1f} s",
            "host_cores_available": os.cpu_count()}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from pysparse_b200 import dist_bench
        return dist_bench.run(args, sys.modules[__name__])
    return run_single(args)


if __name__ == "__main__":
    main()
