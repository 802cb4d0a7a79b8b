"""Config 5b: CSR SpMV on the power-law matrix (Pareto row lengths, uniformly random columns).
    python scripts/powerlaw_probe.py [log2 rows ...]   (default 21 22)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g

peak = 6549.0
for lg in [int(a) for a in sys.argv[1:]] or [21, 22]:
    cache = f"/tmp/powerlaw_{lg}.npz"                 # the generator takes longer than the measurement
    if os.path.exists(cache):
        z = np.load(cache); mat = {"shape": tuple(z["shape"]), "csr": (z["val"], z["col"], z["ind"])}
    else:
        mat = g.powerlaw(1 << lg)
        np.savez(cache, shape=np.array(mat["shape"]), val=mat["csr"][0], col=mat["csr"][1], ind=mat["csr"][2])
    m, n = mat["shape"]
    val, col, ind = mat["csr"]
    A = P.Matrix.csr(m, n, val, col, ind)
    rng = np.random.default_rng(0)
    dx = P.DeviceVector(n, rng.standard_normal(n)); dy = P.DeviceVector(m)
    nb = 12 * len(val) + 4 * (m + 1) + 8 * n + 8 * m
    for what, tr in (("csr", False), ("csr_transp", True)):
        for _ in range(4):
            A.matvec_dev(dx, dy, tr)
        P.event_record(2)
        for _ in range(10):
            A.matvec_dev(dx, dy, tr)
        P.event_record(3)
        dt = P.event_elapsed_ms(2, 3) * 1e-3 / 10
        print(f"power-law 2^{lg} rows, {len(val)} entries (max row {int(np.diff(ind).max())}) {what}: {dt * 1e6:.0f} us, "
              f"{len(val) / dt / 1e9:.1f} G entries/s, {nb / dt / 1e9:.0f} GB/s algorithmic = {nb / dt / 1e9 / peak:.3f} of peak", flush=True)
