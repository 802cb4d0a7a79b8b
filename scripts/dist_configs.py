"""BASELINE configs 3 and 4 on N GPUs (run under torchrun): sharded PCG + Jacobi on 3-D Poisson 256^3
and sharded BiCGSTAB / GMRES(30) + Jacobi on the 3-D convection-diffusion matrix 200^3 (b = standard
normal, seed 1: see DESIGN.md on why not b = ones).  One JSON line per measurement on rank 0, appended
to gpurun_out/dist_configs_<tag>.jsonl.   usage: dist_configs.py [tag] [cfgs]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ["PSP_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctl = dist.new_group(backend="gloo")
from pysparse_b200 import _abi as P
from pysparse_b200 import dist as D
from pysparse_b200.tools import gallery as g

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
want = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c3", "c4"]
DIM = 30


def emit(**kw):
    if rank == 0:
        kw["n_gpus"] = world
        print(json.dumps(kw), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/dist_configs_{tag}.jsonl", "a") as f:
            f.write(json.dumps(kw) + "\n")


def barrier():
    P.synchronize(); dist.barrier(); torch.cuda.synchronize()


def timed(solver, A, db, dx, tol, maxit, K, dim, reps=2):
    best = None
    for _ in range(reps + 1):
        dx.fill(0.0)
        barrier()
        r = D.solve(solver, A, db, dx, tol, maxit, K, dim)
        t = torch.tensor([P.last_device_ms()], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if best is None or ms < best[1]:
            best = (r, ms)
    return best


n_max = 256 ** 3 // world + 2 * 256 * 256 + 64
D.init(rank, world, n_max, nslots=2 * DIM + 2 if "c4" in want else 8, group=ctl)

if "c3" in want:
    n = 256
    N = n ** 3
    rowptr = D.row_partition(N, world, align=n * n)
    r0, r1 = int(rowptr[rank]), int(rowptr[rank + 1])
    slab = g.poisson3d(n, want_sss=False, rows=(r0, r1))
    A = D.ShardedMatrix(rowptr, slab["csr"][2], slab["csr"][1], slab["csr"][0], group=ctl)
    nl = r1 - r0
    K = P.Precon.jacobi(None, np.full(nl, 6.0))
    db = P.DeviceVector(nl, np.ones(nl)); dx = P.DeviceVector(nl)
    r, ms = timed("pcg", A, db, dx, 1e-10, 5000, K, DIM)
    emit(config="c3", what="3-D Poisson 256^3 pcg+jacobi tol 1e-10", info=r[0], iter=r[1], relres=r[2], device_ms=ms,
         iters_per_s=r[1] / (ms * 1e-3))
    del A, K, db, dx, slab

if "c4" in want:
    n = 200
    N = n ** 3
    rowptr = D.row_partition(N, world, align=n * n)
    r0, r1 = int(rowptr[rank]), int(rowptr[rank + 1])
    slab = g.convdiff3d(n, rows=(r0, r1))
    A = D.ShardedMatrix(rowptr, slab["csr"][2], slab["csr"][1], slab["csr"][0], group=ctl)
    nl = r1 - r0
    K = P.Precon.jacobi(None, np.full(nl, 6.0))
    bfull = np.random.default_rng(1).standard_normal(N)
    db = P.DeviceVector(nl, bfull[r0:r1]); dx = P.DeviceVector(nl)
    for solver in ("bicgstab", "gmres"):
        r, ms = timed(solver, A, db, dx, 1e-8, 5000, K, DIM, reps=1)
        emit(config="c4", what=f"conv-diff 200^3 {solver}+jacobi tol 1e-8 b=randn(seed 1)" + (" restart 30" if solver == "gmres" else ""),
             info=r[0], iter=r[1], relres=r[2], device_ms=ms, iters_per_s=r[1] / (ms * 1e-3),
             reference_cpu={"bicgstab": [0, 504, 7.2719e-09, 173.5], "gmres": [0, 745, 9.7272e-09, 437.5]}[solver])
    del A
barrier()
D.shutdown()
dist.destroy_process_group()
