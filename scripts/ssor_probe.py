import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g
import oracle
ora = oracle.Oracle("ref") if oracle.have_ref() else oracle.Oracle("port")
rng = np.random.default_rng(0)
# parity on small cases, both engines
for name, mat in [("p2d64", g.poisson2d(64)), ("p3d20", g.poisson3d(20)), ("rsym", g.random_sym(5000, 8, 3))]:
    n = mat["n"]; S = P.Matrix.sss(n, *mat["sss"]); So = ora.sss(n, *mat["sss"])
    x = rng.standard_normal(n)
    for om, st in ((1.0, 1), (1.0, 2), (1.5, 1), (0.7, 3)):
        yo = np.zeros(n); ora.precon(ora.ssor(So, om, st), x, yo)
        for eng, cl in ((1, 8), (2, 8), (3, 8), (4, 8)):
            P.set_option("ssor_engine", eng); P.set_option("ssor_cluster", cl)
            K = P.Precon.ssor(S, om, st); y = np.full(n, np.nan); K.precon(x, y)
            if not np.array_equal(y, yo):
                print("MISMATCH", name, om, st, eng, cl, np.max(np.abs(y - yo)))
print("parity done")
nn = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
mat = g.poisson2d(nn, want_csr=False); N = mat["n"]
S = P.Matrix.sss(N, *mat["sss"])
dx = P.DeviceVector(N, rng.standard_normal(N)); dy = P.DeviceVector(N)
engines = ((4, 8), (2, 8)) if "--quick" in sys.argv else ((4, 8), (3, 8), (2, 8), (1, 8))
for eng, cl in engines:
    P.set_option("ssor_engine", eng); P.set_option("ssor_cluster", cl)
    K = P.Precon.ssor(S, 1.0, 1)
    P.check(P.lib.psp_prec_apply_dev(K.ptr, dx.ptr, dy.ptr)); P.synchronize()
    P.event_record(2)
    reps = 3
    for _ in range(reps):
        P.check(P.lib.psp_prec_apply_dev(K.ptr, dx.ptr, dy.ptr))
    P.event_record(3)
    ms = P.event_elapsed_ms(2, 3) / reps
    nl = len(mat["sss"][0])
    bsgs = 2 * (12 * nl + 4 * (N + 1)) + 104 * N
    print(f"2-D {nn}^2 SGS apply engine={eng} cluster={cl}: {ms:.2f} ms  ({ms*1e3/(2*(2*nn-1)):.3f} us/level, {bsgs/ms/1e6:.0f} GB/s algorithmic)", flush=True)
P.set_option("ssor_engine", 0); P.set_option("ssor_cluster", 8)
K = P.Precon.ssor(S, 1.0, 1)
db = P.DeviceVector(N, np.ones(N)); dxx = P.DeviceVector(N, np.zeros(N))
r = P.solve("pcg", S, db, dxx, 1e-10, 200, K, dev=True)
print("pcg+ssor 200 its:", r, f"{P.last_device_ms():.0f} ms -> {200/(P.last_device_ms()*1e-3):.1f} it/s")
