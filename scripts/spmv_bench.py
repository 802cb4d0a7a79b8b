"""Developer tool (run under gpurun, optionally under ncu): time the device-resident
SpMV on a 3-D Poisson matrix.  usage: spmv_bench.py N [variant] [reps] [fmt]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P  # noqa: E402
from pysparse_b200.tools import gallery as g  # noqa: E402

nn = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"]   # "variant[:ctas_per_sm]"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
fmt = sys.argv[4] if len(sys.argv) > 4 else "csr"
kind = sys.argv[5] if len(sys.argv) > 5 else "poisson3d"

if kind == "poisson3d":
    mat = g.poisson3d(nn, want_csr=(fmt == "csr"), want_sss=(fmt == "sss"))
    n = mat["n"]; shape = (n, n)
elif kind == "powerlaw":
    mat = g.powerlaw(nn)
    n = mat["n"]; shape = mat["shape"]
elif kind == "poisson2d":
    mat = g.poisson2d(nn, want_csr=(fmt == "csr"), want_sss=(fmt == "sss"))
    n = mat["n"]; shape = (n, n)
rng = np.random.default_rng(0)
if fmt == "csr":
    val, col, ind = mat["csr"]
    nbytes = 12 * len(val) + 4 * (shape[0] + 1) + 8 * shape[1] + 8 * shape[0]
else:
    val, diag, col, ind = mat["sss"]
    nbytes = 12 * len(val) + 4 * (n + 1) + 24 * n
for variant in variants:
    vv = variant.split(":")
    P.set_option("spmv_variant", int(vv[0]))
    P.set_option("spmv_ctas", int(vv[1]) if len(vv) > 1 else 0)
    A = P.Matrix.csr(shape[0], shape[1], val, col, ind) if fmt == "csr" else P.Matrix.sss(n, val, diag, col, ind)
    dx = P.DeviceVector(shape[1], rng.standard_normal(shape[1])); dy = P.DeviceVector(shape[0])
    for _ in range(3):
        A.matvec_dev(dx, dy)
    P.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        A.matvec_dev(dx, dy)
    P.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"spmv {fmt} {kind} {nn} variant {variant}: {dt*1e6:.1f} us  {nbytes/dt/1e9:.0f} GB/s (algorithmic, {nbytes/1e6:.1f} MB)", flush=True)
    del A, dx, dy
