#!/usr/bin/env python
"""bench.py -- headline benchmark of pysparse_b200 (contract: see the task brief).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
3-D 7-point Poisson 256^3 (16.7 M rows, 117 M non-zeros, fp64), PCG + Jacobi,
b = 1, x0 = 0, tol = 1e-10.  One "step" = one complete solve.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid 256]

* value   : PCG iterations per second with matrix and vectors resident in HBM (psp_solve_dev),
            timed with CUDA events on the library stream;
* e2e     : the same solve through the plug-in call a PySparse user makes,
            krylov.pcg(A, b, x, tol, maxit, K) with ordinary pageable NumPy arrays: b, x0 copied
            host->device and x copied back inside the timed region;
* roofline: the persistent solver kernel (one cooperative launch per solve and GPU), algorithmic
            bytes per iteration 12 nnz + 4 (N+1) + 104 N;
* pcg_sss : the same solve on the sss_mat form; spmv: the SpMV half of the metric (BASELINE
            config 2 and 5b: csr / csr transposed / sss on 128^3..384^3 and the power-law matrix,
            device-resident and end to end through A.matvec(x, y));
* cpu_baseline: the reference's own C kernels (oracle/_ref) on one host core,
            first CPU_ITERS iterations of the same solve.

N > 1: ONE process drives N GPUs through the same plug-in call (option "gpus" / PSP_GPUS,
psp_multi.cu): the system is cut into N row slabs (strong scaling).  Under torchrun rank 0 does
that, the other ranks only take part in the barriers and the max-over-ranks reduction.
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
    ap.add_argument("--no-sweep", action="store_true", help="skip the SpMV half of the metric (config 2 / 5b)")
    ap.add_argument("--sweep-sizes", default="128,256,384", help="3-D Poisson grids of the SpMV sweep (512 fits too: about 2 more minutes)")
    ap.add_argument("--powerlaw-rows", type=int, default=1 << 21, help="rows of the power-law SpMV matrix (config 5b: 4194304); 0: skip")
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


# --------------------------------------------------------------------------- B200 arm
def run_b200(args, world=1):
    """One Python process drives `world` GPUs through the unchanged PySparse API (option "gpus",
    psp_multi.cu).  Under torchrun only rank 0 gets here; the other ranks take part in the
    barriers of main()."""
    from pysparse_b200 import _abi as P
    if P.device_count() < world:
        raise SystemExit(f"bench.py: {world} CUDA device(s) needed, {P.device_count()} visible (pysparse_b200 has no CPU fallback)")
    from pysparse_b200.itsolvers import krylov
    from pysparse_b200.sparse import spmatrix
    from pysparse_b200.precon import precon
    peak, peak_src = peaks()
    P.set_option("gpus", world)
    N, val, col, ind, diag = build_problem(args.grid)
    nnz = int(len(val))
    A = P.Matrix.csr(N, N, val, col, ind)
    K = P.Precon.jacobi(A, diag)
    db = P.DeviceVector(N, np.ones(N))
    dx = P.DeviceVector(N)

    def dev_step():
        dx.fill(0.0)
        return P.solve("pcg", A, db, dx, args.tol, args.maxit, K, dev=True)

    for _ in range(args.warmup):
        res = dev_step()
    P.synchronize()

    clocks = ClockSampler(0)
    clocks.start()
    P.launch_count(reset=True)
    P.synchronize()
    P.event_record(0)
    iters, dev_ms = 0, 0.0
    for _ in range(args.steps):
        res = dev_step()
        iters += res[1]
        dev_ms += P.last_device_ms()
    P.event_record(1)
    P.synchronize()
    ms = P.event_elapsed_ms(0, 1)
    launches = P.launch_count()
    clk = clocks.stop()

    # end to end: the plug-in call itself, with ordinary (pageable) NumPy arrays, as a PySparse
    # user makes it -- b and x0 go host -> device(s), x comes back, all inside the timed region
    Am = spmatrix.csr_from_arrays(N, N, ind, col, val)
    Km = precon.jacobi(Am)
    hb = np.ones(N)
    hx = np.zeros(N)
    api = "pysparse_b200.itsolvers.krylov.pcg(A, b, x, tol, maxit, K) with pageable NumPy arrays"

    def e2e_step_fn():
        hx[:] = 0.0                                   # a fresh start vector (not part of the call being timed)
        t0 = time.perf_counter()
        r = krylov.pcg(Am, hb, hx, args.tol, args.maxit, Km)     # returns after x has been copied back
        return r, time.perf_counter() - t0

    e2e_step_fn()
    e2e_iters, e2e_s = 0, 0.0
    for _ in range(args.steps):
        r2, dt = e2e_step_fn()
        e2e_iters += r2[1]
        e2e_s += dt
    del Am, Km

    nl = N // world                                   # rows of one rank's slab
    par = "1 GPU" if world == 1 else (f"{world} GPUs driven by one process: contiguous row slabs, boundary values pushed over "
                                      f"NVLink peer memory inside the solver kernel, in-kernel all-reduce (no collective library)")
    line = {
        "metric": "pcg_iters_per_s", "value": iters / (ms * 1e-3), "unit": "iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3-D 7-point Poisson {args.grid}^3, PCG + Jacobi, tol {args.tol:g}, b=1, x0=0",
                   "rows": N, "nnz": nnz, "format": "csr", "parallelism": par,
                   "device_layout": "diagonal runs (values only; no column indices on the device)",
                   "l2_policy": "inputs larger than L2 (matrix values 0.94 GB + 6 vectors of 134 MB vs 126 MB L2)" if N * 8 * 6 / world > 126e6
                                else "per-GPU working set may fit in L2 at this N (strong scaling)",
                   "result": {"info": res[0], "iter": res[1], "relres": res[2]}},
        "e2e": {"value": e2e_iters / e2e_s, "unit": "iter/s", "h2d_bytes_per_step": 16 * N,
                "d2h_bytes_per_step": 8 * N, "api": api, "host_memory": "pageable", "seconds_per_step": e2e_s / args.steps},
        "gpu_launches": launches,
        "clocks": clk,
        "iters_per_step": iters / args.steps,
        "solver_kernel_ms_per_step": dev_ms / args.steps,
    }

    def traffic_of(key):
        tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
        if os.path.exists(tp) and args.grid == 256 and world == 1:
            return json.load(open(tp)).get(key)
        return None

    # The solve is ONE persistent cooperative kernel per GPU (k_pcg_persist): that launch is the
    # dominant kernel.  Algorithmic bytes per launch = iterations x the per-iteration figure of
    # DESIGN.md / SURVEY 8(d) (CSR matrix once + 13 vector streams) for one rank's slab; duration =
    # the CUDA-event time of that launch (psp_last_device_ms: max over the ranks), against ONE
    # GPU's measured peak.  The kernel moves FEWER bytes than the algorithmic figure (no column
    # indices: 8 instead of 12 bytes per entry), which is how frac can approach or pass 1.
    per = dev_ms * 1e-3 / args.steps
    nb = alg_bytes_pcg_iter(nl, nnz // world) * iters / args.steps
    line["roofline"] = {"bound": "hbm", "kernel": "k_pcg_persist<1> (whole PCG + Jacobi solve in one cooperative launch per GPU, diagonal-run matrix format)",
                        "achieved": nb / per / 1e9, "peak": peak, "unit": "GB/s", "frac": nb / per / 1e9 / peak,
                        "traffic": traffic_of("pcg_persist_runs_dram_bytes_per_launch"), "peak_source": peak_src,
                        "launches_sampled": args.steps, "avg_launch_us": per * 1e6,
                        "algorithmic_bytes_per_launch": nb,
                        "algorithmic_bytes_per_iteration": alg_bytes_pcg_iter(nl, nnz // world),
                        "moved_bytes_per_iteration_by_design": 8 * (nnz // world) + 104 * nl}
    if world == 1:
        line["pcg_sss"] = pcg_sss(P, args, peak)
        if not args.no_sweep:
            line["spmv"] = spmv_sweep(P, args)
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args, N, val, col, ind, diag)
    P.set_option("gpus", 1)
    return line


def pcg_sss(P, args, peak):
    """The same solve on the sss_mat form of the matrix (the reference's idiom for symmetric
    systems: examples/poisson_test.py solves on L.to_sss())."""
    from pysparse_b200.tools import gallery as g
    mat = g.poisson3d(args.grid, want_csr=False, want_sss=True)
    N = mat["n"]
    val, diag, col, ind = mat["sss"]
    S = P.Matrix.sss(N, val, diag, col, ind)
    K = P.Precon.jacobi(None, diag)
    db = P.DeviceVector(N, np.ones(N)); dx = P.DeviceVector(N)
    ms, iters = 0.0, 0
    for rep in range(1 + args.steps):
        dx.fill(0.0)
        r = P.solve("pcg", S, db, dx, args.tol, args.maxit, K, dev=True)
        if rep:
            ms += P.last_device_ms(); iters += r[1]
    nb = 12 * len(val) + 4 * (N + 1) + 8 * N + 104 * N          # SSS matrix once + 13 vector streams
    return {"value": iters / (ms * 1e-3), "unit": "iter/s", "result": {"info": r[0], "iter": r[1], "relres": r[2]},
            "algorithmic_bytes_per_iteration": nb, "achieved_GBps": nb * iters / (ms * 1e-3) / 1e9,
            "frac": nb * iters / (ms * 1e-3) / 1e9 / peak}


def spmv_sweep(P, args):
    """BASELINE config 2 (+ 5b): device-resident csr_mat / sss_mat matvec and csr matvec_transp on
    3-D Poisson grids, the power-law matrix, and the same product end to end through
    A.matvec(x, y) with pageable NumPy arrays."""
    from pysparse_b200.tools import gallery as g
    from pysparse_b200.sparse import spmatrix
    out = []
    rng = np.random.default_rng(0)
    peak, _ = peaks()

    def timed(A, dx, dy, transp=False, reps=30):
        for _ in range(4):
            A.matvec_dev(dx, dy, transp)
        P.event_record(2)
        for _ in range(reps):
            A.matvec_dev(dx, dy, transp)
        P.event_record(3)
        return P.event_elapsed_ms(2, 3) * 1e-3 / reps

    def e2e(Am, m, n, transp=False, reps=3):
        x = rng.standard_normal(m if transp else n); y = np.empty(n if transp else m)
        f = Am.matvec_transp if transp else Am.matvec
        f(x, y)
        t0 = time.perf_counter()
        for _ in range(reps):
            f(x, y)
        return (time.perf_counter() - t0) / reps

    sizes = [int(v) for v in args.sweep_sizes.split(",") if v]
    for n in sizes:
        mat = g.poisson3d(n)
        N = mat["n"]
        dx = P.DeviceVector(N, rng.standard_normal(N)); dy = P.DeviceVector(N)
        val, col, ind = mat["csr"]
        A = P.Matrix.csr(N, N, val, col, ind)
        nb = alg_bytes_spmv(N, len(val)); flops = 2 * len(val)
        for what, tr in (("csr", False), ("csr_transp", True)):
            dt = timed(A, dx, dy, tr)
            rec = {"matrix": f"poisson3d {n}^3", "op": what, "us": dt * 1e6, "GBps_algorithmic": nb / dt / 1e9,
                   "frac": nb / dt / 1e9 / peak, "GFLOPs": flops / dt / 1e9}
            if n <= 256:
                Am = spmatrix.csr_from_arrays(N, N, ind, col, val)
                rec["e2e_ms_pageable_numpy"] = e2e(Am, N, N, tr) * 1e3
                del Am
            out.append(rec)
        del A
        val, diag, col, ind = mat["sss"]
        S = P.Matrix.sss(N, val, diag, col, ind)
        nb = 12 * len(val) + 4 * (N + 1) + 24 * N; flops = 4 * len(val) + 2 * N
        dt = timed(S, dx, dy)
        rec = {"matrix": f"poisson3d {n}^3", "op": "sss", "us": dt * 1e6, "GBps_algorithmic": nb / dt / 1e9,
               "frac": nb / dt / 1e9 / peak, "GFLOPs": flops / dt / 1e9}
        if n <= 256:
            Sm = spmatrix.sss_from_arrays(N, ind, col, val, diag)
            rec["e2e_ms_pageable_numpy"] = e2e(Sm, N, N) * 1e3
            del Sm
        out.append(rec)
        del S, dx, dy, mat
    if args.powerlaw_rows:
        mat = g.powerlaw(args.powerlaw_rows)
        m, n = mat["shape"] if "shape" in mat else (mat["n"], mat["n"])
        val, col, ind = mat["csr"]
        A = P.Matrix.csr(m, n, val, col, ind)
        dx = P.DeviceVector(n, rng.standard_normal(n)); dy = P.DeviceVector(m)
        nb = 12 * len(val) + 4 * (m + 1) + 8 * n + 8 * m
        for what, tr in (("csr", False), ("csr_transp", True)):
            dt = timed(A, dx, dy, tr, reps=10)
            out.append({"matrix": f"power-law rows {m} (mean {len(val) / m:.1f}, max {int(np.diff(ind).max())} entries per row)",
                        "op": what, "us": dt * 1e6, "GBps_algorithmic": nb / dt / 1e9, "frac": nb / dt / 1e9 / peak,
                        "GFLOPs": 2 * len(val) / dt / 1e9})
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
    rank = int(os.environ.get("RANK", "0"))
    n_gpus = max(args.gpus, world)
    dist = ctl = None
    if world > 1:
        # torchrun launches one process per GPU; ONE process (rank 0) drives all the GPUs through
        # the plug-in API, the others hold their place in the launch protocol: barrier before
        # and after, and the max-over-ranks reduction of the timed region.  They wait on a CPU
        # (gloo) barrier: an NCCL barrier would park a spinning kernel on every GPU, and the
        # persistent solver kernels are cooperative launches that need all SMs of their GPU.
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        ctl = dist.new_group(backend="gloo")
        dist.barrier()
        torch.cuda.synchronize()
        dist.barrier(group=ctl)
    line = run_b200(args, n_gpus) if rank == 0 else None
    if world > 1:
        import torch
        dist.barrier(group=ctl)
        torch.cuda.synchronize()
        t = torch.tensor([line["ms_per_step"] if line else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if line:
            line["ms_per_step"] = float(t.item())
        dist.destroy_process_group()
    if line:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
