"""Per-iteration time of PCG + Jacobi on one GPU, single-launch (persistent) kernel vs the
multi-kernel path, at the slab sizes of the 256^3 strong-scaling run (256^3 / 1, 2, 4, 8)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g

IT = 300
sizes = [int(s) for s in sys.argv[1:]] or [8, 4, 2, 1]
for parts in sizes:
    n = 256
    rows = (0, n ** 3 // parts)
    slab = g.poisson3d(n, want_sss=False, rows=rows)
    val, col, ind = slab["csr"]
    nl = rows[1]
    keep = col < nl                      # drop the couplings to the next slab: a square local problem
    cnt = np.add.reduceat(keep.astype(np.int64), ind[:-1]) if parts > 1 else np.diff(ind)
    ind2 = np.zeros(nl + 1, dtype=np.int32); np.cumsum(cnt, out=ind2[1:])
    A = P.Matrix.csr(nl, nl, val[keep], col[keep], ind2)
    K = P.Precon.jacobi(None, np.full(nl, 6.0))
    db = P.DeviceVector(nl, np.ones(nl)); dx = P.DeviceVector(nl)
    out = []
    for persist in (1, 0):
        P.set_option("pcg_persist", persist)
        best = 1e9
        for _ in range(3):
            dx.fill(0.0)
            r = P.solve("pcg", A, db, dx, 1e-300, IT, K, dev=True)
            best = min(best, P.last_device_ms())
        out.append(best * 1e3 / IT)
    print(f"rows=256^3/{parts} ({nl}): persist {out[0]:.1f} us/iter, multi-kernel {out[1]:.1f} us/iter", flush=True)
    del A, K, db, dx
