"""How much BASELINE config 4 with b = 1 (conv-diff 200^3, BiCGSTAB + Jacobi, tol 1e-8) depends on
rounding in the REFERENCE ITSELF.  Two runs of the reference's own kernels (oracle/_ref sources):
  perturb : the stock build, right-hand side b = 1 + 1e-13 * standard_normal(seed 0)
  fma     : b = 1 exactly, the same sources compiled with -mfma -ffp-contract=fast
Results are merged into tests/golden/config4_b1.json under "sensitivity".
Run in the build container (4-5 minutes of CPU each):
    python tests/golden/make_config4_sensitivity.py perturb|fma [n]"""
import importlib.util, json, os, shutil, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pysparse_b200.tools import gallery as g

variant = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
if variant == "fma":
    work = "/tmp/oracle_fma"
    shutil.rmtree(work, ignore_errors=True)
    shutil.copytree(os.path.join(ROOT, "oracle"), work, ignore=shutil.ignore_patterns("_ref", "*.so", "__pycache__"))
    flags = "-O2 -fPIC -mfma -ffp-contract=fast -fno-math-errno -w"
    subprocess.run(["make", "-s", "-C", work, "ref", "oracle", "CFLAGS=" + flags], check=True)
    spec = importlib.util.spec_from_file_location("oracle_fma", os.path.join(work, "__init__.py"))
    oracle = importlib.util.module_from_spec(spec); spec.loader.exec_module(oracle)
else:
    import oracle
    oracle.build(ref=True)
ora = oracle.Oracle("ref")
mat = g.convdiff3d(n); N = mat["n"]
A = ora.csr(N, N, *mat["csr"]); K = ora.jacobi(A, np.full(N, 6.0))
b = np.ones(N)
if variant == "perturb":
    b = b + 1e-13 * np.random.default_rng(0).standard_normal(N)
x = np.zeros(N); t0 = time.time()
info, it, rr = ora.solve("bicgstab", A, b, x, 1e-8, 5000, K)
dt = time.time() - t0
r = np.empty(N); ora.matvec(A, x, r)
rec = {"info": info, "iter": it, "relres_dec": float(rr) if np.isfinite(rr) else None,
       "true_relres": float(np.linalg.norm(b - r) / np.linalg.norm(b)), "cpu_seconds": round(dt, 1),
       "what": {"perturb": "stock reference build, b = 1 + 1e-13 * standard_normal(seed 0)",
                "fma": "reference sources compiled with -mfma -ffp-contract=fast, b = 1"}[variant]}
print(variant, rec, flush=True)
name = os.path.join(ROOT, "tests", "golden", "config4_b1.json" if n == 200 else f"config4_b1_{n}.json")
doc = json.load(open(name)) if os.path.exists(name) else {}
doc.setdefault("sensitivity", {})[variant] = rec
json.dump(doc, open(name, "w"), indent=1)
