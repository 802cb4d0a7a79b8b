import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
nn = int(sys.argv[1]); eng = int(sys.argv[2]); cl = int(sys.argv[3]) if len(sys.argv) > 3 else 8
P.set_option("ssor_engine", eng); P.set_option("ssor_cluster", cl)
mat = g.poisson2d(nn, want_csr=False); N = mat["n"]
S = P.Matrix.sss(N, *mat["sss"]); K = P.Precon.ssor(S, 1.0, 1)
dx = P.DeviceVector(N, np.random.default_rng(0).standard_normal(N)); dy = P.DeviceVector(N)
for _ in range(3):
    P.check(P.lib.psp_prec_apply_dev(K.ptr, dx.ptr, dy.ptr))
P.synchronize()
P.event_record(2); P.check(P.lib.psp_prec_apply_dev(K.ptr, dx.ptr, dy.ptr)); P.event_record(3)
ms = P.event_elapsed_ms(2, 3)
print(f"{nn}^2 engine {eng} cluster {cl}: {ms:.3f} ms, {ms*1e3/(2*(2*nn-1)):.3f} us/level")
