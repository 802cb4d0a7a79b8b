"""Small runs of every kernel with hand-rolled synchronisation, as the target of
compute-sanitizer (memcheck / racecheck / synccheck): scripts/sanitize.sh.
Sizes are tiny on purpose: the sanitizer slows kernels down by orders of magnitude."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200 import _abi as P
from pysparse_b200.tools import gallery as g

which = sys.argv[1:] or ["spmv", "pcg", "ssor", "phase"]
mat = g.poisson3d(12); n = mat["n"]
A = P.Matrix.csr(n, n, *mat["csr"]); S = P.Matrix.sss(n, *mat["sss"])
x = np.random.default_rng(0).standard_normal(n); y = np.empty(n)
if "spmv" in which:
    A.matvec(x, y); A.matvec(x, y, True); S.matvec(x, y)            # diagonal-run kernels (automatic format)
    P.set_option("spmv_format", 0)                                  # CSR row blocks through the TMA ring
    A0 = P.Matrix.csr(n, n, *mat["csr"]); S0 = P.Matrix.sss(n, *mat["sss"])
    P.set_option("spmv_format", -1)
    A0.matvec(x, y); A0.matvec(x, y, True); S0.matvec(x, y)
    irr = g.powerlaw(2000, mean_nnz=30, seed=3, max_len=4096)
    if irr is not None:
        B = P.Matrix.csr(irr["n"], irr["n"], *irr["csr"])
        yy = np.empty(irr["n"]); B.matvec(np.ones(irr["n"]), yy)
    print("spmv done", flush=True)
if "pcg" in which:
    for fmt in (-1, 0):
        P.set_option("spmv_format", fmt)
        Sf = P.Matrix.sss(n, *mat["sss"])
        P.set_option("spmv_format", -1)
        for persist in (1, 0):
            P.set_option("pcg_persist", persist)
            xs = np.zeros(n)
            print("pcg format", fmt, "persist", persist, P.solve("pcg", Sf, np.ones(n), xs, 1e-8, 60, P.Precon.jacobi(None, mat["sss"][1])), flush=True)
    P.set_option("bicg_persist", 1)
    xs = np.zeros(n)
    print("bicgstab single launch", P.solve("bicgstab", A, np.ones(n), xs, 1e-8, 40, P.Precon.jacobi(None, np.full(n, 6.0))), flush=True)
    P.set_option("bicg_persist", -1)
    P.set_option("pcg_persist", -1)
if "ssor" in which:
    m2 = g.poisson2d(24)
    S2 = P.Matrix.sss(m2["n"], *m2["sss"])
    for eng in (1, 2, 4):
        P.set_option("ssor_engine", eng)
        for om, st in ((1.0, 1), (1.3, 2)):
            K = P.Precon.ssor(S2, om, st)
            yy = np.empty(m2["n"]); K.precon(np.ones(m2["n"]), yy)
    P.set_option("ssor_engine", 0)
    print("ssor done", flush=True)
if "phase" in which:
    for solver in ("minres", "qmrs", "bicgstab", "cgs", "gmres"):
        xs = np.zeros(n)
        print(solver, P.solve(solver, A, np.ones(n), xs, 1e-8, 40, P.Precon.jacobi(None, np.full(n, 6.0)), 10), flush=True)
print("sanitize target finished", flush=True)
