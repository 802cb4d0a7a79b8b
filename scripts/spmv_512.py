"""BASELINE config 2, upper end: SSS (and optionally CSR) SpMV on the 3-D Poisson matrix 512^3
(134 217 728 rows).  The lower triangle is generated directly (the general gallery generator
takes longer than the GPU box time this measurement is worth).
    python scripts/spmv_512.py [grid=512] [--csr]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

g = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 512
t0 = time.perf_counter()
n = g ** 3
i = np.arange(n, dtype=np.int32)
hx = (i % g) > 0
hy = ((i // g) % g) > 0
hz = i >= g * g
cnt = hz.astype(np.int32) + hy + hx
ind = np.zeros(n + 1, dtype=np.int32)
np.cumsum(cnt, out=ind[1:])
nnz = int(ind[-1])
col = np.empty(nnz, dtype=np.int32)
p = ind[:-1].copy()
col[p[hz]] = i[hz] - g * g
p += hz
col[p[hy]] = i[hy] - g
p += hy
col[p[hx]] = i[hx] - 1
del p, hx, hy, hz, cnt
val = np.full(nnz, -1.0)
diag = np.full(n, 6.0)
print(f"generated {g}^3: n={n} lower nnz={nnz} in {time.perf_counter() - t0:.1f} s", flush=True)
if "--dry" in sys.argv:
    sys.exit(0)

from pysparse_b200 import _abi as P
peak = 6549.4
t0 = time.perf_counter()
S = P.Matrix.sss(n, val, diag, col, ind)
dx = P.DeviceVector(n, np.random.default_rng(0).standard_normal(n)); dy = P.DeviceVector(n)
S.matvec_dev(dx, dy); P.synchronize()
print(f"device matrix ready in {time.perf_counter() - t0:.1f} s", flush=True)
for _ in range(3):
    S.matvec_dev(dx, dy)
P.event_record(2)
for _ in range(10):
    S.matvec_dev(dx, dy)
P.event_record(3)
dt = P.event_elapsed_ms(2, 3) * 1e-4
nb = 12 * nnz + 4 * (n + 1) + 24 * n
print(f"spmv sss poisson3d {g}^3: {dt * 1e6:.0f} us, {nb / dt / 1e9:.0f} GB/s algorithmic (B_sss = {nb / 1e9:.2f} GB) = {nb / dt / 1e9 / peak:.3f} of peak, "
      f"{2 * (2 * nnz + n) / dt / 1e9:.0f} GFLOP/s", flush=True)
# spot check against the stencil itself: y = A x at a few interior rows
y = dy.download(); x = dx.download()
for z, yy, xx in ((5, 7, 9), (g // 2, g // 2, 3), (g - 2, g - 2, g - 2)):          # interior nodes
    r = (z * g + yy) * g + xx
    ref = 6.0 * x[r] - x[r - 1] - x[r + 1] - x[r - g] - x[r + g] - x[r - g * g] - x[r + g * g]
    assert abs(y[r] - ref) <= 1e-12 * max(1.0, abs(ref)), (r, y[r], ref)
print("spot check ok", flush=True)
