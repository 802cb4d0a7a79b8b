"""A/B timing of library builds on N GPUs driven by one process (psp_multi.cu):
    python scripts/multi_ab.py N grid lib1 lib2 ...      ("default" = the in-tree library)"""
import os, subprocess, sys
n, grid = sys.argv[1], sys.argv[2]
for lib in sys.argv[3:]:
    env = dict(os.environ)
    if lib != "default":
        env["PSP_B200_LIB"] = lib
    code = f"""
import numpy as np, sys
sys.path.insert(0, '.')
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
P.set_option('gpus', {n})
mat = g.poisson3d({grid}, want_sss=False); n = mat['n']
A = P.Matrix.csr(n, n, *mat['csr']); K = P.Precon.jacobi(A, np.full(n, 6.0))
db = P.DeviceVector(n, np.ones(n)); dx = P.DeviceVector(n)
best = 1e30
for rep in range(5):
    dx.fill(0.0)
    r = P.solve('pcg', A, db, dx, 1e-10, 5000, K, dev=True)
    best = min(best, P.last_device_ms())
print('{lib}', {n}, 'GPUs', {grid}, r, f'{{best:.1f}} ms {{best * 1e3 / r[1]:.1f}} us/iter {{r[1] / best * 1e3:.0f}} it/s', flush=True)
"""
    subprocess.run([sys.executable, "-c", code], env=env)
