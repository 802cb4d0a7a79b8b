"""Timing diagnostics of the sharded persistent PCG (results are INVALID in debug modes):
per-iteration time with the halo pull and / or the cross-GPU all-reduce removed."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); os.environ["PSP_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local)); ctl = dist.new_group(backend="gloo")
from pysparse_b200 import _abi as P, dist as D
from pysparse_b200.tools import gallery as g
n = 256; N = n ** 3
rowptr = D.row_partition(N, world, align=n * n); r0, r1 = int(rowptr[rank]), int(rowptr[rank + 1]); nl = r1 - r0
slab = g.poisson3d(n, want_sss=False, rows=(r0, r1))
D.init(rank, world, nl + 2 * n * n + 64, nslots=4, group=ctl)
A = D.ShardedMatrix(rowptr, slab["csr"][2], slab["csr"][1], slab["csr"][0], group=ctl)
K = P.Precon.jacobi(None, np.full(nl, 6.0)); db = P.DeviceVector(nl, np.ones(nl)); dx = P.DeviceVector(nl)
IT = 300
MODES = [tuple(int(c) for c in m) for m in os.environ.get("DIAG_MODES", "10 00 11 01 12 02 10 00").split()]
for push, mode in MODES:
    P.set_option("dist_push", push)
    P.set_option("dist_debug", mode)
    best = 1e9
    for rep in range(3):
        dx.fill(0.0); P.synchronize(); dist.barrier()
        r = D.pcg(A, db, dx, 1e-300, IT, K)
        t = torch.tensor([P.last_device_ms()], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    if rank == 0:
        print(f"world={world} dist_push={push} dist_debug={mode}: {best*1e3/IT:.1f} us/iter (iter={r[1]})", flush=True)
P.set_option("dist_debug", 0)
P.set_option("dist_push", 1)
dist.barrier(); del A; D.shutdown(); dist.destroy_process_group()
