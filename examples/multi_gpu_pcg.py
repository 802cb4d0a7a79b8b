"""Row-sharded PCG + Jacobi on a 3-D Poisson matrix, one process per GPU:

    torchrun --nproc-per-node N --master-addr 127.0.0.1 examples/multi_gpu_pcg.py [n]

Every rank assembles only its own slab of rows (whole grid planes), the halo exchange and the
scalar all-reduces are done by the solver kernel over NVLink peer memory; torch.distributed is
used for the rendezvous only."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # run from the source tree

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
os.environ["PSP_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctl = dist.new_group(backend="gloo")                      # Python objects (IPC handles, halo lists)

from pysparse_b200 import _abi as P                        # noqa: E402
from pysparse_b200 import dist as D                        # noqa: E402
from pysparse_b200.tools import gallery                    # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = n ** 3
rowptr = D.row_partition(N, world, align=n * n)
r0, r1 = int(rowptr[rank]), int(rowptr[rank + 1])
slab = gallery.poisson3d(n, want_sss=False, rows=(r0, r1))
val, col, ind = slab["csr"]
nl = r1 - r0

D.init(rank, world, nl + 2 * n * n + 64, nslots=4, group=ctl)
A = D.ShardedMatrix(rowptr, ind, col, val, group=ctl)
K = P.Precon.jacobi(None, np.full(nl, 6.0))
b = P.DeviceVector(nl, np.ones(nl))
x = P.DeviceVector(nl)
info, it, relres = D.pcg(A, b, x, 1e-10, 5000, K)
ms = P.last_device_ms()
if rank == 0:
    print("%d GPUs, %d^3 unknowns: info=%d iter=%d relres=%.3e  %.1f ms  (%.0f iterations/s)" %
          (world, n, info, it, relres, ms, it / (ms * 1e-3)))
del A
D.shutdown()
dist.destroy_process_group()
