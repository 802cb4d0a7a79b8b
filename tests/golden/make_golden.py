"""Generate the golden fixtures in tests/golden/ from the REFERENCE's own C
kernels (oracle/_ref/libpsp_ref.so, compiled from /root/reference by
oracle/Makefile).  Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Outputs (committed):
  solver_kat.json  -- (info, iter, relres, |x|_2, x[0], x[n//2]) for every
                      solver / preconditioner / matrix case of CASES
  kernels.npz      -- SpMV / transposed SpMV / SSS SpMV / Jacobi / SSOR
                      outputs on small seeded inputs (bit patterns)
The reference's own test-suite holds no vectors for this path (SURVEY.md
section 4), so these files are the pin: tests/test_oracle.py checks the
plain-C restatement against them, and the GPU tests check the CUDA path
against the restatement and against these numbers.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import Oracle  # noqa: E402
from pysparse_b200.tools import gallery as g  # noqa: E402
from cases import SOLVER_CASES, KERNEL_CASES, build_matrix, diag_of, make_ops  # noqa: E402


def main():
    ref = Oracle("ref")
    print(ref.describe())
    out = []
    for case in SOLVER_CASES:
        mat = build_matrix(case["matrix"])
        n = mat["n"]
        A, K = make_ops(ref, mat, case)
        b = np.ones(n)
        x = np.zeros(n) if case.get("x0", "zeros") == "zeros" else np.full(n, 0.5)
        if case.get("b") == "zero":
            b = np.zeros(n)
        info, it, rr = ref.solve(case["solver"], A, b, x, case["tol"], case["maxit"], K, case.get("dim", 20))
        rec = dict(case)
        rec.update(info=info, iter=it, relres=float(rr).hex() if rr == rr else "nan",
                   xnorm=float(np.linalg.norm(x)).hex(), x0=float(x[0]).hex(), xmid=float(x[n // 2]).hex())
        out.append(rec)
        print(case["id"], info, it, rr)
    with open(os.path.join(HERE, "solver_kat.json"), "w") as f:
        json.dump(out, f, indent=1)

    arrays = {}
    for kc in KERNEL_CASES:
        mat = build_matrix(kc["matrix"])
        n = mat["n"]
        shape = mat.get("shape", (n, n))
        rng = np.random.default_rng(kc["seed"])
        x = rng.standard_normal(shape[1])
        xt = rng.standard_normal(shape[0])
        key = kc["id"]
        A = ref.csr(shape[0], shape[1], *mat["csr"])
        y = np.empty(shape[0]); ref.matvec(A, x, y); arrays[key + "/csr_y"] = y
        yt = np.empty(shape[1]); ref.matvec(A, xt, yt, transp=True); arrays[key + "/csr_yt"] = yt
        if "sss" in mat:
            S = ref.sss(n, *mat["sss"])
            ys = np.zeros(n); ref.matvec(S, x, ys); arrays[key + "/sss_y"] = ys
            for om, st in kc.get("ssor", []):
                K = ref.ssor(S, om, st)
                z = np.zeros(n); ref.precon(K, x, z); arrays[f"{key}/ssor_{om}_{st}"] = z
            for om, st in kc.get("jacobi", []):
                K = ref.jacobi(S, mat["sss"][1], om, st)
                z = np.zeros(n); ref.precon(K, x, z); arrays[f"{key}/jacobi_{om}_{st}"] = z
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **arrays)
    print("wrote", len(out), "solver cases and", len(arrays), "kernel vectors")


if __name__ == "__main__":
    main()
