"""N > 1 arm of bench.py: the headline solve row-sharded over N GPUs of one node
(strong scaling; one process per GPU under torchrun)."""
from __future__ import annotations

import json
import os
import time

import numpy as np


def setup(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    os.environ.setdefault("PSP_DEVICE", str(local))
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    ctl = dist.new_group(backend="gloo")       # control plane for Python objects (IPC handle blobs)
    return torch, dist, ctl, rank, world, local


def build_slab(args, rank, world):
    from pysparse_b200 import dist as D
    from pysparse_b200.tools import gallery as g

    n = args.grid
    N = n ** 3
    rowptr = D.row_partition(N, world, align=n * n)       # whole l-planes per rank
    r0, r1 = int(rowptr[rank]), int(rowptr[rank + 1])
    slab = g.poisson3d(n, want_sss=False, rows=(r0, r1))
    return N, rowptr, slab["csr"]


def run(args, bench):
    torch, dist, ctl, rank, world, local = setup(args)
    from pysparse_b200 import _abi as P
    from pysparse_b200 import dist as D

    peak, peak_src = bench.peaks()
    N, rowptr, (val, col, ind) = build_slab(args, rank, world)
    nl = len(ind) - 1
    nnz_local = int(len(val))
    plane = args.grid * args.grid
    slot = nl + 2 * plane + 64
    D.init(rank, world, slot, nslots=4, group=ctl)
    A = D.ShardedMatrix(rowptr, ind, col, val, group=ctl)
    K = P.Precon.jacobi(None, np.full(nl, 6.0))
    db = P.DeviceVector(nl, np.ones(nl))
    dx = P.DeviceVector(nl)

    def step():
        dx.fill(0.0)
        return D.pcg(A, db, dx, args.tol, args.maxit, K)

    def barrier():
        P.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step()
    clocks = bench.ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    P.spmv_timing(8)
    P.launch_count(reset=True)
    barrier()
    P.event_record(0)
    t0 = time.perf_counter()
    iters = 0
    for _ in range(args.steps):
        res = step()
        iters += res[1]
    P.event_record(1)
    P.synchronize()
    ms = P.event_elapsed_ms(0, 1)
    barrier()
    wall = time.perf_counter() - t0
    launches = P.launch_count()
    n_spmv, spmv_ms = P.spmv_timing(0)
    clk = clocks.stop() if rank == 0 else None
    persistent = n_spmv == 0
    if persistent:
        # the sharded solve ran as ONE persistent kernel (no separate SpMV launches to
        # sample): time the SpMV kernel in one extra solve on the per-phase path
        P.set_option("pcg_persist", 0)
        P.spmv_timing(4)
        step()
        n_spmv, spmv_ms = P.spmv_timing(0)
        P.set_option("pcg_persist", -1)

    # max over ranks of the device-timed region
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    nnz_t = torch.tensor([nnz_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())

    # end to end: host slices in pinned memory -> device -> sharded solve -> host
    hb = P.pinned_empty(nl); hb[:] = 1.0
    hx = P.pinned_empty(nl)

    def e2e_step():
        hx[:] = 0.0
        db.upload(hb); dx.upload(hx)
        r = D.pcg(A, db, dx, args.tol, args.maxit, K)
        dx.download(hx)
        return r

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        e2e_iters += e2e_step()[1]
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    # residual check of the sharded solution (true residual, all-reduced)
    if rank == 0:
        line = {
            "metric": "pcg_iters_per_s", "value": iters / (ms_max * 1e-3), "unit": "iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3-D 7-point Poisson {args.grid}^3, PCG + Jacobi, tol {args.tol:g}, b=1, x0=0",
                       "rows": N, "nnz": nnz, "format": "csr",
                       "parallelism": f"{world} GPUs, contiguous row slabs (whole grid planes), boundary entries pushed over NVLink "
                                      f"peer memory, in-kernel all-reduce",
                       "l2_policy": "inputs larger than L2 per GPU" if nnz_local * 12 > 126e6 else
                                    "per-GPU working set may fit in L2 at this N (strong scaling)",
                       "result": {"info": res[0], "iter": res[1], "relres": res[2]}},
            "e2e": {"value": e2e_iters / e2e_s, "unit": "iter/s", "h2d_bytes_per_step": 16 * N,
                    "d2h_bytes_per_step": 8 * N, "api": "pysparse_b200.dist.pcg with pinned host slices",
                    "seconds_per_step": e2e_s / args.steps},
            "gpu_launches": launches * world,
            "clocks": clk,
            "iters_per_step": iters / args.steps,
            "wall_s_per_step": wall / args.steps,
        }
        spmv_line = None
        if n_spmv:
            per = spmv_ms / n_spmv * 1e-3
            nb = 12 * nnz_local + 4 * (nl + 1) + 8 * (nl + A.nhalo) + 8 * nl
            spmv_line = {"bound": "hbm", "kernel": "k_spmv_tma<2,true> on rank 0's slab",
                         "achieved": nb / per / 1e9, "peak": peak, "unit": "GB/s", "frac": nb / per / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src, "launches_sampled": n_spmv,
                         "avg_launch_us": per * 1e6, "algorithmic_bytes_per_launch": nb}
        if persistent:
            # dominant kernel = the persistent solve kernel of each rank (one launch per solve):
            # rank 0's algorithmic bytes (its slab once + 13 vector streams per iteration)
            # over the max-over-ranks launch time, against ONE GPU's peak
            per = ms_max * 1e-3 / args.steps
            nb = (12 * nnz_local + 4 * (nl + 1) + 104 * nl) * iters / args.steps
            line["roofline"] = {"bound": "hbm", "kernel": "k_pcg_persist on rank 0's slab (whole sharded solve, one launch per rank)",
                                "achieved": nb / per / 1e9, "peak": peak, "unit": "GB/s", "frac": nb / per / 1e9 / peak,
                                "traffic": None, "peak_source": peak_src, "launches_sampled": args.steps,
                                "avg_launch_us": per * 1e6, "algorithmic_bytes_per_launch": nb}
            if spmv_line:
                line["roofline_spmv"] = spmv_line
        elif spmv_line:
            line["roofline"] = spmv_line
        print(json.dumps(line), flush=True)
    barrier()
    del A
    D.shutdown()
    dist.destroy_process_group()
