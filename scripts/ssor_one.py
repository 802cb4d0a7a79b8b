"""usage: ssor_one.py N [engine] [reps] -- SGS applications on 2-D Poisson N^2 (ncu target)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
nn = int(sys.argv[1]); eng = int(sys.argv[2]) if len(sys.argv) > 2 else 0; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
P.set_option("ssor_engine", eng)
mat = g.poisson2d(nn, want_csr=False); N = mat["n"]
S = P.Matrix.sss(N, *mat["sss"]); K = P.Precon.ssor(S, 1.0, 1)
dx = P.DeviceVector(N, np.random.default_rng(0).standard_normal(N)); dy = P.DeviceVector(N)
for _ in range(reps):
    P.check(P.lib.psp_prec_apply_dev(K.ptr, dx.ptr, dy.ptr))
P.synchronize()
print("done")
