"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C-ABI,
against the CPU oracle and the committed golden fixtures.

Contract (BASELINE.json north_star):
  * SpMV: relative max-norm error <= 1e-13 (short-row paths are in fact
    bit-identical to the reference because they keep its accumulation order);
  * solvers: same flag, iteration count within +-2 %, relres within 1e-10;
  * structure work (level sets, transposes): bit-exact.
"""
import numpy as np
import pytest

from cases import KERNEL_CASES, SOLVER_CASES, build_matrix, make_ops
from helpers import GpuLib, load_kat, load_kernels, relmax, start_vectors

pytestmark = pytest.mark.gpu

SPMV_TOL = 1e-13       # north_star: relative max-norm error of SpMV
RELRES_TOL = 1e-10     # north_star: final relative residual within 1e-10 of the reference
ITER_TOL = 0.02        # north_star: iteration count within +-2 %
KAT = {r["id"]: r for r in load_kat()}


@pytest.fixture(scope="module")
def gpu(native):
    if native.device_count() == 0:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return GpuLib(native)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, "runs", "auto"])
@pytest.mark.parametrize("kc", KERNEL_CASES, ids=lambda c: c["id"])
def test_spmv_and_precon_kernels_vs_golden(gpu, native, kc, variant):
    """variant 1-4: the CSR row-block kernels (register-streamed / TMA ring, device format 0);
    "runs": the diagonal-run format forced on every matrix, irregular ones included (their blocks
    overflow the descriptor cache: slow path); "auto": what the analysis picks."""
    native.set_option("spmv_variant", variant if isinstance(variant, int) else 0)
    native.set_option("spmv_format", 0 if isinstance(variant, int) else (1 if variant == "runs" else -1))
    try:
        G = load_kernels()
        mat = build_matrix(kc["matrix"])
        n = mat["n"]
        shape = mat.get("shape", (n, n))
        rng = np.random.default_rng(kc["seed"])
        x = rng.standard_normal(shape[1])
        xt = rng.standard_normal(shape[0])
        key = kc["id"]
        A = gpu.csr(shape[0], shape[1], *mat["csr"])
        y = np.full(shape[0], np.nan); gpu.matvec(A, x, y)
        assert relmax(y, G[key + "/csr_y"]) <= SPMV_TOL
        yt = np.full(shape[1], np.nan); gpu.matvec(A, xt, yt, transp=True)
        assert relmax(yt, G[key + "/csr_yt"]) <= SPMV_TOL
        # row blocks averaging <= 16 entries per row keep the reference's summation order
        if key not in ("pow_3000", "tendigit_500"):
            assert np.array_equal(y, G[key + "/csr_y"]) and np.array_equal(yt, G[key + "/csr_yt"])
        elif variant in (3, 4, "auto"):
            # irregular blocks (TMA-ring kernels): products entry-parallel, then every row of at most
            # 48 entries is added by one thread in row order -- bit-identical; only longer rows are
            # summed by a warp (psp_tma_ring.cuh)
            val, col, ind = mat["csr"]
            short = np.diff(ind) <= 48
            assert short.any() and np.array_equal(y[short], G[key + "/csr_y"][short])
            short_t = np.bincount(col, minlength=shape[1]) <= 48
            assert np.array_equal(yt[short_t], G[key + "/csr_yt"][short_t])
        if "sss" in mat:
            S = gpu.sss(n, *mat["sss"])
            ys = np.full(n, np.nan); gpu.matvec(S, x, ys)
            assert relmax(ys, G[key + "/sss_y"]) <= SPMV_TOL
            yst = np.full(n, np.nan); gpu.matvec(S, x, yst, transp=True)   # matvec_transp aliased (sss_mat.c:108)
            assert np.array_equal(ys, yst)
            for om, st in kc.get("ssor", []):
                z = np.full(n, np.nan); gpu.precon(gpu.ssor(S, om, st), x, z)
                assert relmax(z, G[f"{key}/ssor_{om}_{st}"]) <= SPMV_TOL
            for om, st in kc.get("jacobi", []):
                z = np.full(n, np.nan); gpu.precon(gpu.jacobi(S, mat["sss"][1], om, st), x, z)
                assert relmax(z, G[f"{key}/jacobi_{om}_{st}"]) <= SPMV_TOL
    finally:
        native.set_option("spmv_variant", 0)
        native.set_option("spmv_format", -1)


def test_spmv_strided_views(gpu, checker):
    mat = build_matrix(["convdiff3d", 9])
    n = mat["n"]
    A = gpu.csr(n, n, *mat["csr"]); Ao = checker.csr(n, n, *mat["csr"])
    rng = np.random.default_rng(3)
    xs = rng.standard_normal(3 * n)[::3]
    ys = np.zeros(2 * n)[::2]; yo = np.zeros(2 * n)[::2]
    gpu.matvec(A, xs, ys); checker.matvec(Ao, xs, yo)
    assert np.array_equal(ys, yo)
    gpu.matvec(A, xs, ys, transp=True); checker.matvec(Ao, xs, yo, transp=True)
    assert np.array_equal(ys, yo)


@pytest.fixture(params=["runs", "rowblocks"])
def device_format(request, native):
    """Both device layouts of a matrix, forced: diagonal runs (the automatic choice for matrices
    of a gigabyte and more) and the CSR row-block stream (the automatic choice below that)."""
    native.set_option("spmv_format", 1 if request.param == "runs" else 0)
    yield request.param
    native.set_option("spmv_format", -1)


@pytest.mark.parametrize("case", SOLVER_CASES, ids=lambda c: c["id"])
def test_solvers_vs_golden_and_oracle(gpu, checker, case, device_format):
    mat = build_matrix(case["matrix"])
    n = mat["n"]
    A, K = make_ops(gpu, mat, case)
    b, x = start_vectors(case, n)
    info, it, rr = gpu.solve(case["solver"], A, b, x, case["tol"], case["maxit"], K, case.get("dim", 20))
    g = KAT[case["id"]]
    assert info == g["info"], (info, it, rr, g)
    if np.isnan(g["relres"]):
        assert np.isnan(rr)
    elif g["info"] == 0 or case.get("stable", False):
        if case.get("stable", False):
            assert abs(it - g["iter"]) <= max(1, int(ITER_TOL * g["iter"])), (it, g["iter"])
            if g["info"] == 0 or g["relres"] < 1e3:
                assert abs(rr - g["relres"]) <= max(RELRES_TOL, 1e-6 * abs(g["relres"])), (rr, g["relres"])
        else:
            # rounding-chaotic case: converged below tol like the reference
            assert rr < case["tol"] and abs(it - g["iter"]) <= 0.35 * g["iter"]
    if g["info"] == 0:
        # same solution as the reference to solver accuracy
        Ao, Ko = make_ops(checker, mat, case)
        bo, xo = start_vectors(case, n)
        checker.solve(case["solver"], Ao, bo, xo, case["tol"], case["maxit"], Ko, case.get("dim", 20))
        scale = max(1.0, float(np.max(np.abs(xo))))
        assert float(np.max(np.abs(x - xo))) <= 50 * case["tol"] * scale * max(1.0, n ** 0.5) + 1e-9 * scale


def test_ssor_level_sets_bit_exact(gpu):
    """Level sets of L / L^T against a direct Python restatement of the recurrence."""
    mat = build_matrix(["random_sym", 3000, 7, 4])
    n = mat["n"]
    val, diag, col, ind = mat["sss"]
    S = gpu.sss(n, val, diag, col, ind)
    K = gpu.ssor(S, 1.0, 1)
    nf, nb, lf, lb = K.levels()
    ef = np.zeros(n, dtype=np.int32)
    for i in range(n):
        c = col[ind[i]:ind[i + 1]]
        ef[i] = ef[c].max() + 1 if len(c) else 0
    eb = np.zeros(n, dtype=np.int32)
    for i in range(n - 1, -1, -1):
        for j in col[ind[i]:ind[i + 1]]:
            eb[j] = max(eb[j], eb[i] + 1)
    assert np.array_equal(lf, ef) and np.array_equal(lb, eb)
    assert nf == ef.max() + 1 and nb == eb.max() + 1
    # 2-D lexicographic grid: 2n-1 wavefronts
    m2 = build_matrix(["poisson2d", 23])
    K2 = gpu.ssor(gpu.sss(m2["n"], *m2["sss"]))
    assert K2.levels()[:2] == (45, 45)


def test_foreign_operator_callbacks(gpu, native, checker):
    """A and K given as duck-typed Python objects (host-staged callback path)."""
    mat = build_matrix(["poisson2d", 40])
    n = mat["n"]
    A = gpu.csr(n, n, *mat["csr"])
    Ao = checker.csr(n, n, *mat["csr"])

    class PyMat:
        shape = (n, n)
        calls = 0

        def matvec(self, x, y):
            PyMat.calls += 1
            checker.matvec(Ao, x, y)

    class PyPrec:
        shape = (n, n)

        def precon(self, x, y):
            y[:] = x / 4.0

    for solver in ("pcg", "minres", "qmrs", "bicgstab", "cgs", "gmres"):
        x1 = np.zeros(n); x2 = np.zeros(n)
        r1 = native.solve(solver, A, np.ones(n), x1, 1e-8, 500, native.Precon.jacobi(A, np.full(n, 4.0)))
        r2 = native.solve(solver, PyMat(), np.ones(n), x2, 1e-8, 500, PyPrec())
        assert r1[0] == r2[0] == 0
        assert abs(r1[1] - r2[1]) <= max(1, 0.1 * r1[1])
        assert np.max(np.abs(x1 - x2)) <= 1e-6 * np.max(np.abs(x1))

    class Boom:
        shape = (n, n)

        def matvec(self, x, y):
            raise ValueError("boom")

    with pytest.raises(ValueError):
        native.solve("pcg", Boom(), np.ones(n), np.zeros(n), 1e-8, 10)


def test_device_resident_solve_and_launch_counter(native):
    mat = build_matrix(["poisson3d", 20])
    n = mat["n"]
    S = native.Matrix.sss(n, *mat["sss"])
    K = native.Precon.jacobi(S, mat["sss"][1])
    db = native.DeviceVector(n, np.ones(n)); dx = native.DeviceVector(n, np.zeros(n))
    native.set_option("pcg_persist", 0)          # per-phase kernels: >= 3 launches per iteration
    native.launch_count(reset=True)
    r = native.solve("pcg", S, db, dx, 1e-10, 1000, K, dev=True)
    assert r[0] == 0 and native.launch_count() >= 3 * r[1]
    native.set_option("pcg_persist", 1)          # single-launch persistent kernel: same result
    dx.fill(0.0)
    native.launch_count(reset=True)
    r1 = native.solve("pcg", S, db, dx, 1e-10, 1000, K, dev=True)
    native.set_option("pcg_persist", -1)
    assert native.launch_count() == 1 and r1[:2] == r[:2] and abs(r1[2] - r[2]) <= 1e-12
    x = dx.download()
    y = np.zeros(n); S.matvec(x, y)
    assert np.linalg.norm(1.0 - y) / np.sqrt(n) <= 1e-9


@pytest.mark.parametrize("nn", [96])
def test_full_size_properties(native, nn):
    """Size-independent properties at a larger size: linearity of the SpMV, symmetry
    x'(Ay) == y'(Ax), SSS == CSR, transposed == direct for a symmetric matrix."""
    mat = build_matrix(["poisson3d", nn])
    n = mat["n"]
    A = native.Matrix.csr(n, n, *mat["csr"]); S = native.Matrix.sss(n, *mat["sss"])
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n); z = rng.standard_normal(n)
    ax = np.empty(n); az = np.empty(n); axz = np.empty(n); sx = np.empty(n); atx = np.empty(n)
    A.matvec(x, ax); A.matvec(z, az); A.matvec(x + 2.0 * z, axz); S.matvec(x, sx); A.matvec(x, atx, transp=True)
    assert relmax(axz, ax + 2.0 * az) <= 1e-13
    assert np.array_equal(ax, sx) and np.array_equal(ax, atx)
    assert abs(z @ ax - x @ az) <= 1e-10 * abs(z @ ax)


PCG_CASES = [c for c in SOLVER_CASES if c["solver"] == "pcg"]


@pytest.mark.parametrize("case", PCG_CASES, ids=lambda c: c["id"])
def test_pcg_single_launch_kernel_vs_per_phase_kernels(gpu, native, case, device_format):
    """The persistent single-launch PCG kernel (default) and the per-phase kernel sequence
    are two schedules of the same algorithm: same flag, same iteration count, same
    residual to rounding, on every PCG case of the suite (zero rhs, maxit hit, long rows,
    SSS and CSR, identity / Jacobi / SSOR -- the last one is not eligible and must simply
    agree with itself)."""
    mat = build_matrix(case["matrix"])
    n = mat["n"]
    out = []
    for persist in (0, 1, -1):
        native.set_option("pcg_persist", persist)
        try:
            A, K = make_ops(gpu, mat, case)
            b, x = start_vectors(case, n)
            native.launch_count(reset=True)
            r = gpu.solve("pcg", A, b, x, case["tol"], case["maxit"], K)
            out.append((r, x.copy(), native.launch_count()))
        finally:
            native.set_option("pcg_persist", -1)
    (r0, x0, _), (r1, x1, _), (r2, x2, _) = out
    assert r1 == r2 and np.array_equal(x1, x2)                       # default == forced on
    assert r0[0] == r1[0] and abs(r0[1] - r1[1]) <= max(1, int(ITER_TOL * r0[1])), (r0, r1)
    if np.isfinite(r0[2]):
        assert abs(r0[2] - r1[2]) <= max(RELRES_TOL, 1e-6 * abs(r0[2])), (r0, r1)
    scale = max(1.0, float(np.max(np.abs(x0))))
    assert float(np.max(np.abs(x0 - x1))) <= 50 * case["tol"] * scale * max(1.0, n ** 0.5) + 1e-9 * scale


BICG_CASES = [c for c in SOLVER_CASES if c["solver"] == "bicgstab" and c.get("prec") in (None, "jacobi") and c.get("steps", 1) == 1]


@pytest.mark.parametrize("case", BICG_CASES, ids=lambda c: c["id"])
def test_bicgstab_single_launch_kernel_vs_per_phase_kernels(gpu, native, case, device_format):
    """k_bicgstab_persist (the default; bicg_persist = 0 selects the per-phase kernel sequence)
    runs the same algorithm as the per-phase kernels: same flag, iteration count within the
    contract, same residual and iterate to solver accuracy (the order of the partial sums of the
    dot products differs between the two, so the last bits do)."""
    mat = build_matrix(case["matrix"])
    n = mat["n"]
    out = []
    for persist in (0, 1):
        native.set_option("bicg_persist", persist)
        try:
            A, K = make_ops(gpu, mat, case)
            b, x = start_vectors(case, n)
            native.launch_count(reset=True)
            r = gpu.solve("bicgstab", A, b, x, case["tol"], case["maxit"], K)
            out.append((r, x.copy(), native.launch_count()))
        finally:
            native.set_option("bicg_persist", -1)
    (r0, x0, l0), (r1, x1, l1) = out
    assert l1 == 1 and l0 > 1
    tol_it = ITER_TOL if case.get("stable", False) else 0.35     # Poisson cases: rounding-chaotic (parity)
    assert r0[0] == r1[0] and abs(r0[1] - r1[1]) <= max(1, int(tol_it * r0[1])), (r0, r1)
    if np.isfinite(r0[2]) and r0[0] == 0:
        assert r1[2] < case["tol"]
    scale = max(1.0, float(np.max(np.abs(x0))))
    assert float(np.max(np.abs(x0 - x1))) <= 50 * case["tol"] * scale * max(1.0, n ** 0.5) + 1e-9 * scale


GMRES_CASES = [c for c in SOLVER_CASES if c["solver"] == "gmres" and c.get("prec") in (None, "jacobi") and c.get("steps", 1) == 1]


@pytest.mark.parametrize("case", GMRES_CASES, ids=lambda c: c["id"])
def test_gmres_single_launch_kernel_vs_per_phase_kernels(gpu, native, case, device_format):
    """k_gmres_persist (gmres_persist = 1; chosen automatically from 4 GPUs on) against the
    per-phase kernel sequence: same flag and restart-cycle structure, iteration count within
    the contract, same true-residual relres and iterate."""
    mat = build_matrix(case["matrix"])
    n = mat["n"]
    out = []
    for persist in (0, 1):
        native.set_option("gmres_persist", persist)
        try:
            A, K = make_ops(gpu, mat, case)
            b, x = start_vectors(case, n)
            native.launch_count(reset=True)
            r = gpu.solve("gmres", A, b, x, case["tol"], case["maxit"], K, case.get("dim", 20))
            out.append((r, x.copy(), native.launch_count()))
        finally:
            native.set_option("gmres_persist", -1)
    (r0, x0, l0), (r1, x1, l1) = out
    assert l1 == 1 and l0 > 1
    assert r0[0] == r1[0] and abs(r0[1] - r1[1]) <= max(1, int(ITER_TOL * r0[1])), (r0, r1)
    assert abs(r0[2] - r1[2]) <= max(RELRES_TOL, 1e-6 * abs(r0[2])), (r0, r1)
    scale = max(1.0, float(np.max(np.abs(x0))))
    assert float(np.max(np.abs(x0 - x1))) <= 50 * case["tol"] * scale * max(1.0, n ** 0.5) + 1e-9 * scale


@pytest.mark.parametrize("matrix", [["poisson2d", 40], ["random_sym", 3000, 7, 4], ["poisson3d", 12]],
                         ids=["poisson2d_40", "random_sym_3000", "poisson3d_12"])
def test_ssor_sweep_engines_agree_bit_for_bit(gpu, native, checker, matrix):
    """Grid-cooperative, thread-block-cluster and single-CTA sweep engines walk the same
    level schedule and add each row's terms in the same (reference) order."""
    mat = build_matrix(matrix)
    n = mat["n"]
    x = np.random.default_rng(11).standard_normal(n)
    So = checker.sss(n, *mat["sss"])
    zo = np.zeros(n); checker.precon(checker.ssor(So, 1.3, 2), x, zo)
    ref = None
    for engine in (0, 1, 2, 3, 4):
        native.set_option("ssor_engine", engine)
        try:
            S = gpu.sss(n, *mat["sss"])
            z = np.full(n, np.nan); gpu.precon(gpu.ssor(S, 1.3, 2), x, z)
        finally:
            native.set_option("ssor_engine", 0)
        assert relmax(z, zo) <= SPMV_TOL
        if ref is None:
            ref = z
        assert np.array_equal(z, ref), engine


def test_options_round_trip(native):
    for name, value in (("pcg_persist", 0), ("ssor_engine", 2), ("l2_stream", 1), ("dist_push", 0), ("ssor_cluster", 4)):
        old = native.get_option(name)
        native.set_option(name, value)
        assert native.get_option(name) == value
        native.set_option(name, old)
    with pytest.raises(Exception):
        native.set_option("no_such_option", 1)


def test_plugin_api_drives_two_gpus_from_one_process(native):
    """PSP_GPUS=2: krylov.pcg / bicgstab / gmres on sss_mat / csr_mat objects run sharded over two
    GPUs from this one process and agree with the one-GPU solve (scripts/multi_check.py).
    Needs two GPUs."""
    import os, subprocess, sys
    if native.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "scripts", "multi_check.py"), "2"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MULTI CHECK PASSED" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]


def test_sharded_solvers_two_gpus(native):
    """Multi-GPU path (row slabs, push-model halo exchange, in-kernel all-reduce): sharded PCG /
    BiCGSTAB / GMRES against the single-GPU solve and the oracle.  Needs two GPUs."""
    import os, socket, subprocess, sys
    if native.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "scripts", "dist_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DIST CHECK PASSED" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]


@pytest.mark.parametrize("grid", [96, 384, 200])
def test_run_format_bridged_gaps_on_unaligned_grids_vs_oracle(gpu, native, checker, grid):
    """Grids whose lines do not line up with the 256-row blocks (2-D Poisson, lines of 96 / 384 / 200
    rows), run format forced.  The -1 / +1 runs bridge the missing entry at every line start with a
    marker value; the mirrored slices of the symmetric form are cut at the target blocks' boundaries
    and may therefore START or END on such a marker (lines of 384: row 768 = 3 * 256) -- every block in
    reach of one must skip it.  Products, transposed products and a PCG solve against the reference's
    kernels, bit for bit."""
    from pysparse_b200.tools import gallery
    native.set_option("spmv_format", 1)
    try:
        mat = gallery.poisson2d(grid)
        n = mat["n"]
        x = np.random.default_rng(grid).standard_normal(n)
        A = gpu.csr(n, n, *mat["csr"]); S = gpu.sss(n, *mat["sss"])
        Ao = checker.csr(n, n, *mat["csr"]); So = checker.sss(n, *mat["sss"])
        y = np.full(n, np.nan); gpu.matvec(A, x, y)
        yo = np.zeros(n); checker.matvec(Ao, x, yo)
        assert np.array_equal(y, yo)
        yt = np.full(n, np.nan); gpu.matvec(A, x, yt, transp=True)
        yto = np.zeros(n); checker.matvec(Ao, x, yto, transp=True)
        assert np.array_equal(yt, yto)
        ys = np.full(n, np.nan); gpu.matvec(S, x, ys)
        yso = np.zeros(n); checker.matvec(So, x, yso)
        assert np.array_equal(ys, yso)
        b = np.ones(n); xs = np.zeros(n); xso = np.zeros(n)
        K = gpu.jacobi(S, mat["sss"][1], 1.0, 1); Ko = checker.jacobi(So, mat["sss"][1])
        r = gpu.solve("pcg", S, b, xs, 1e-8, 60, K)
        ro = checker.solve("pcg", So, b, xso, 1e-8, 60, Ko)
        assert r[:2] == ro[:2] and abs(r[2] - ro[2]) <= 1e-6 * abs(ro[2])
        assert np.max(np.abs(xs - xso)) <= 1e-9 * np.max(np.abs(xso))
    finally:
        native.set_option("spmv_format", -1)
