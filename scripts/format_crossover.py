"""PCG + Jacobi and SpMV, diagonal-run format vs CSR row blocks, over problem sizes (which format
should the automatic choice take for small matrices?)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
cases = [("poisson2d", 100), ("poisson2d", 300), ("poisson2d", 1000), ("poisson3d", 48), ("poisson3d", 64), ("poisson3d", 96),
         ("poisson3d", 128), ("poisson3d", 160), ("poisson3d", 192)]
for kind, nn in cases:
    mat = getattr(g, kind)(nn); n = mat["n"]
    dg = mat["sss"][1]
    out = []
    for fmt in (1, 0):
        P.set_option("spmv_format", fmt)
        A = P.Matrix.csr(n, n, *mat["csr"]); S = P.Matrix.sss(n, *mat["sss"])
        P.set_option("spmv_format", -1)
        dx = P.DeviceVector(n, np.ones(n)); dy = P.DeviceVector(n); db = P.DeviceVector(n, np.ones(n))
        K = P.Precon.jacobi(None, dg)
        for nm, M in (("csr", A), ("sss", S)):
            for _ in range(5): M.matvec_dev(dx, dy)
            P.event_record(0)
            for _ in range(50): M.matvec_dev(dx, dy)
            P.event_record(1)
            us = P.event_elapsed_ms(0, 1) * 20
            dxx = P.DeviceVector(n); 
            for rep in range(2):
                dxx.fill(0.0)
                r = P.solve("pcg", M, db, dxx, 1e-10, 5000, K, dev=True)
            out.append(f"{nm} fmt{fmt}: spmv {us:7.1f} us, pcg {P.last_device_ms()*1e3/r[1]:7.1f} us/it")
    print(f"{kind} {nn} (n={n}): " + " | ".join(out), flush=True)
