"""CPU tests: the oracle is pinned.

* the plain-C restatement (oracle/psp_oracle.c) must agree BIT-EXACTLY with the
  reference's own kernels (oracle/_ref, compiled from /root/reference) -- runs
  only where that library exists (the build container);
* the restatement must reproduce the committed golden fixtures (generated from
  the reference kernels by tests/golden/make_golden.py) -- runs everywhere.
"""
import numpy as np
import pytest

from cases import KERNEL_CASES, SOLVER_CASES, build_matrix, make_ops
from helpers import load_kat, load_kernels, start_vectors

KAT = {r["id"]: r for r in load_kat()}


def same_float(a, b):
    return (a == b) or (np.isnan(a) and np.isnan(b))


@pytest.mark.parametrize("case", SOLVER_CASES, ids=lambda c: c["id"])
def test_port_reproduces_golden_solver_table(port, case):
    mat = build_matrix(case["matrix"])
    n = mat["n"]
    A, K = make_ops(port, mat, case)
    b, x = start_vectors(case, n)
    info, it, rr = port.solve(case["solver"], A, b, x, case["tol"], case["maxit"], K, case.get("dim", 20))
    g = KAT[case["id"]]
    assert (info, it) == (g["info"], g["iter"])
    assert same_float(rr, g["relres"])
    assert float(np.linalg.norm(x)) == g["xnorm"]
    assert x[0] == g["x0"] and x[n // 2] == g["xmid"]


@pytest.mark.parametrize("kc", KERNEL_CASES, ids=lambda c: c["id"])
def test_port_reproduces_golden_kernels(port, kc):
    G = load_kernels()
    mat = build_matrix(kc["matrix"])
    n = mat["n"]
    shape = mat.get("shape", (n, n))
    rng = np.random.default_rng(kc["seed"])
    x = rng.standard_normal(shape[1])
    xt = rng.standard_normal(shape[0])
    key = kc["id"]
    A = port.csr(shape[0], shape[1], *mat["csr"])
    y = np.empty(shape[0]); port.matvec(A, x, y)
    assert np.array_equal(y, G[key + "/csr_y"])
    yt = np.empty(shape[1]); port.matvec(A, xt, yt, transp=True)
    assert np.array_equal(yt, G[key + "/csr_yt"])
    if "sss" in mat:
        S = port.sss(n, *mat["sss"])
        ys = np.zeros(n); port.matvec(S, x, ys)
        assert np.array_equal(ys, G[key + "/sss_y"])
        for om, st in kc.get("ssor", []):
            z = np.zeros(n); port.precon(port.ssor(S, om, st), x, z)
            assert np.array_equal(z, G[f"{key}/ssor_{om}_{st}"])
        for om, st in kc.get("jacobi", []):
            z = np.zeros(n); port.precon(port.jacobi(S, mat["sss"][1], om, st), x, z)
            assert np.array_equal(z, G[f"{key}/jacobi_{om}_{st}"])


@pytest.mark.parametrize("case", SOLVER_CASES[::3], ids=lambda c: c["id"])
def test_port_bit_identical_to_reference_kernels(port, ref, case):
    mat = build_matrix(case["matrix"])
    n = mat["n"]
    out = []
    for lib in (port, ref):
        A, K = make_ops(lib, mat, case)
        b, x = start_vectors(case, n)
        r = lib.solve(case["solver"], A, b, x, case["tol"], case["maxit"], K, case.get("dim", 20))
        out.append((r, x))
    (r1, x1), (r2, x2) = out
    assert r1[:2] == r2[:2] and same_float(r1[2], r2[2])
    assert np.array_equal(x1, x2)


def test_strided_and_sss_scatter_semantics(port, ref):
    mat = build_matrix(["random_sym", 400, 6, 1])
    n = mat["n"]
    rng = np.random.default_rng(0)
    for lib in (port, ref):
        S = lib.sss(n, *mat["sss"]); A = lib.csr(n, n, *mat["csr"])
        x = rng.standard_normal(3 * n)[::3]
        y1 = np.zeros(2 * n)[::2]; y2 = np.zeros(n)
        lib.matvec(S, x, y1); lib.matvec(A, np.ascontiguousarray(x), y2)
        # SSS and the expanded CSR accumulate in the same order
        assert np.array_equal(np.ascontiguousarray(y1), y2)


def test_duck_typed_operator_callback(port):
    """solvers accept any object with matvec/precon (spmatrixmodule.c:169-248)"""
    mat = build_matrix(["poisson2d", 20])
    n = mat["n"]
    A = port.csr(n, n, *mat["csr"])
    calls = []

    def mv(x, y):
        calls.append(1)
        port.matvec(A, x, y)

    Acb = port.callback(n, mv)
    x1 = np.zeros(n); x2 = np.zeros(n)
    r1 = port.solve("pcg", A, np.ones(n), x1, 1e-8, 500)
    r2 = port.solve("pcg", Acb, np.ones(n), x2, 1e-8, 500)
    assert r1 == r2 and np.array_equal(x1, x2) and len(calls) == r1[1] + 1


def test_jacobi_rejects_tiny_diagonal(port):
    with pytest.raises(ValueError):
        port.jacobi(None, np.array([1.0, 1e-300, 2.0]))
