import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
import oracle
ora = oracle.Oracle("ref") if oracle.have_ref() else oracle.Oracle("port")
P.set_option("ssor_engine", 4)
for nn in [int(a) for a in sys.argv[1:]]:
    mat = g.poisson2d(nn, want_csr=False); N = mat["n"]
    S = P.Matrix.sss(N, *mat["sss"]); K = P.Precon.ssor(S, 1.0, 1)
    x = np.random.default_rng(0).standard_normal(N); y = np.full(N, np.nan)
    t0 = time.time()
    try:
        K.precon(x, y)
        err = None
    except Exception as e:
        err = str(e)
    dt = time.time() - t0
    So = ora.sss(N, *mat["sss"]); yo = np.zeros(N); ora.precon(ora.ssor(So, 1.0, 1), x, yo)
    print(nn, "error" if err else "ok", err, f"{dt*1e3:.1f} ms (host) device {P.last_device_ms():.2f} ms", "bit-identical" if np.array_equal(y, yo) else f"MISMATCH nan={np.isnan(y).sum()}", flush=True)
