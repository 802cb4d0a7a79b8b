"""Multi-GPU parity check (run under torchrun on a >= 2-GPU box): sharded PCG / BiCGSTAB / GMRES
vs the single-GPU solve and the CPU oracle; repeated solves on the same handles."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["PSP_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctl = dist.new_group(backend="gloo")
from pysparse_b200 import _abi as P
from pysparse_b200 import dist as D
from pysparse_b200.tools import gallery as g
import oracle

DIM = 10
NMAX = 48
D.init(rank, world, NMAX ** 3 // world + 2 * NMAX * NMAX + 64, nslots=2 * DIM + 2, group=ctl)
ok = True
cases = [("poisson3d", 24, "pcg", "ones", True), ("poisson3d", 48, "pcg", "ones", True),
         ("convdiff3d", 32, "bicgstab", "randn", True), ("convdiff3d", 32, "gmres", "randn", True),
         ("convdiff3d", 40, "gmres", "ones", False), ("convdiff3d", 24, "bicgstab", "randn", False)]
only = os.environ.get("DIST_CASES")               # e.g. DIST_CASES=pcg: a quick subset
if only:
    cases = [c for c in cases if c[2] in only.split(",")]
for kind, n, solver, rhs, jac in cases:
    N = n ** 3
    tol = 1e-10 if kind == "poisson3d" else 1e-8
    rowptr = D.row_partition(N, world, align=n * n)
    r0, r1 = int(rowptr[rank]), int(rowptr[rank + 1])
    gen = getattr(g, kind)
    slab = gen(n, want_sss=False, rows=(r0, r1)) if kind == "poisson3d" else gen(n, rows=(r0, r1))
    val, col, ind = slab["csr"]
    nl = r1 - r0
    bfull = np.ones(N) if rhs == "ones" else np.random.default_rng(1).standard_normal(N)
    A = D.ShardedMatrix(rowptr, ind, col, val, group=ctl)
    K = P.Precon.jacobi(None, np.full(nl, 6.0)) if jac else None
    db = P.DeviceVector(nl, bfull[r0:r1]); dx = P.DeviceVector(nl, np.zeros(nl))
    res = D.solve(solver, A, db, dx, tol, 2000, K, DIM)
    xl = dx.download()
    parts = [None] * world
    dist.all_gather_object(parts, xl, group=ctl)
    x = np.concatenate(parts)
    if rank == 0:
        full = gen(n)
        ora = oracle.Oracle("ref") if oracle.have_ref() else oracle.Oracle("port")
        Ao = ora.csr(N, N, *full["csr"]); Ko = ora.jacobi(Ao, np.full(N, 6.0)) if jac else None
        xo = np.zeros(N)
        ro = ora.solve(solver, Ao, bfull, xo, tol, 2000, Ko, DIM)
        A1 = P.Matrix.csr(N, N, *full["csr"]); K1 = P.Precon.jacobi(None, np.full(N, 6.0)) if jac else None
        x1 = np.zeros(N)
        r1g = P.solve(solver, A1, bfull, x1, tol, 2000, K1, DIM)
        good = (res[0] == ro[0] == 0 and abs(res[1] - ro[1]) <= max(1, 0.02 * ro[1]) and abs(res[2] - ro[2]) <= 1e-10
                and np.max(np.abs(x - xo)) <= 1e-6 * np.max(np.abs(xo)) and abs(res[1] - r1g[1]) <= max(1, 0.02 * r1g[1]))
        print(f"{kind} n={n} {solver} b={rhs} jacobi={jac} world={world}: sharded={res} single_gpu={r1g} oracle={ro} "
              f"max|x-x_oracle|={np.max(np.abs(x - xo)):.2e} {'OK' if good else 'FAIL'}", flush=True)
        ok = ok and good
    # a second solve on the same handles (sequence numbers keep counting)
    dx.fill(0.0)
    res2 = D.solve(solver, A, db, dx, tol, 2000, K, DIM)
    if res2 != res:
        print(f"rank {rank}: repeated solve differs: {res} vs {res2}", flush=True)
        ok = False
    # pull-model halo exchange on the same handles, then push again
    for push in (0, 1):
        P.set_option("dist_push", push)
        dx.fill(0.0)
        res3 = D.solve(solver, A, db, dx, tol, 2000, K, DIM)
        if res3 != res:
            print(f"rank {rank}: dist_push={push} solve differs: {res} vs {res3}", flush=True)
            ok = False
    del A
flags = [None] * world
dist.all_gather_object(flags, ok, group=ctl)
D.shutdown()
dist.destroy_process_group()
if rank == 0:
    print("DIST CHECK", "PASSED" if all(flags) else "FAILED", flush=True)
sys.exit(0 if all(flags) else 1)
