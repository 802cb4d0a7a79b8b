"""A/B timing of persistent-PCG builds on one box: persist_ab.py N lib1 lib2 ... (each in a fresh process)."""
import os, subprocess, sys
n = sys.argv[1]
for lib in sys.argv[2:]:
    env = dict(os.environ, PSP_B200_LIB=lib if lib != "default" else "")
    if lib == "default":
        env.pop("PSP_B200_LIB")
    code = f"""
import numpy as np, sys
sys.path.insert(0, '.')
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
mat = g.poisson3d({n}); n = mat['n']
db = P.DeviceVector(n, np.ones(n)); K = P.Precon.jacobi(None, np.full(n, 6.0))
for nm in ('csr', 'sss'):
    M = P.Matrix.csr(n, n, *mat['csr']) if nm == 'csr' else P.Matrix.sss(n, *mat['sss'])
    for rep in range(2):
        dx = P.DeviceVector(n); dx.fill(0.0)
        r = P.solve('pcg', M, db, dx, 1e-10, 5000, K, dev=True)
        ms = P.last_device_ms()
    print('{lib}', nm, {n}, r, f'{{ms:.1f}} ms {{ms * 1e3 / r[1]:.1f}} us/iter {{r[1] / ms * 1e3:.0f}} it/s', flush=True)
"""
    subprocess.run([sys.executable, "-c", code], env=env)
