"""CPU check of the SEQUENCING of the single-launch solver kernels (no GPU needed).

tests/harness/persist_flow.cpp transcribes the control flow of k_pcg_persist,
k_bicgstab_persist and k_gmres_persist phase by phase and drives the product's own
run_epilogue() / State (compiled from pysparse_b200/csrc/psp_epilogue.cuh) -- here it is
run against the oracle: same flag, iterations within 2 %, residual and solution to tolerance.
The PCG flow belongs to a kernel that is validated on the GPU, so it also validates the
method; the BiCGSTAB and GMRES flows belong to kernels written at the end of round 1 that
are checked on the GPU as well (tests/test_gpu_parity.py); this is the CPU-only check of their control flow."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pysparse_b200.tools import gallery

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def flow(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("flow") / "libpersist_flow.so")
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I" + cuda_inc,
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "pysparse_b200", "csrc"),
           os.path.join(ROOT, "tests", "harness", "persist_flow.cpp"), "-o", out]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
    lib = C.CDLL(out)
    common = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_double, C.c_int]
    lib.flow_pcg.argtypes = common + [_ip, _ip, _dp]
    lib.flow_bicgstab.argtypes = common + [_ip, _ip, _dp]
    lib.flow_gmres.argtypes = common + [C.c_int, _ip, _ip, _dp]
    return lib


def run_flow(lib, solver, mat, b, tol, maxit, jac, dim=20):
    n = mat["n"]
    val, col, ind = (np.ascontiguousarray(a) for a in mat["csr"])
    val = val.astype(np.float64); col = col.astype(np.int32); ind = ind.astype(np.int32)
    x = np.zeros(n)
    b = np.ascontiguousarray(b, dtype=np.float64)
    dinv = None if jac is None else np.ascontiguousarray(1.0 / jac)
    info, it, rr = C.c_int(), C.c_int(), C.c_double()
    args = [n, ind.ctypes.data_as(_ip), col.ctypes.data_as(_ip), val.ctypes.data_as(_dp),
            dinv.ctypes.data_as(_dp) if dinv is not None else None, b.ctypes.data_as(_dp), x.ctypes.data_as(_dp), tol, maxit]
    if solver == "gmres":
        args.append(dim)
    getattr(lib, "flow_" + solver)(*args, C.byref(info), C.byref(it), C.byref(rr))
    return (info.value, it.value, rr.value), x


def run_oracle(checker, solver, mat, b, tol, maxit, jac, dim=20):
    n = mat["n"]
    Ao = checker.csr(n, n, *mat["csr"])
    Ko = checker.jacobi(Ao, jac) if jac is not None else None
    xo = np.zeros(n)
    return checker.solve(solver, Ao, b, xo, tol, maxit, Ko, dim), xo


def agree(rf, xf, ro, xo, tol):
    assert rf[0] == ro[0], (rf, ro)
    assert abs(rf[1] - ro[1]) <= max(1, int(0.02 * ro[1])), (rf, ro)
    if np.isfinite(ro[2]):
        assert abs(rf[2] - ro[2]) <= max(1e-10, 1e-6 * abs(ro[2])), (rf, ro)
    else:
        assert not np.isfinite(rf[2])
    if ro[0] == 0:
        scale = max(1.0, float(np.max(np.abs(xo))))
        assert float(np.max(np.abs(xf - xo))) <= 50 * tol * scale * max(1.0, len(xo) ** 0.5) + 1e-9 * scale


MATS = {
    "poisson2d_30": lambda: gallery.poisson2d(30),
    "poisson3d_12": lambda: gallery.poisson3d(12),
    "convdiff3d_12": lambda: gallery.convdiff3d(12),
}


@pytest.mark.parametrize("jac", [False, True], ids=["identity", "jacobi"])
@pytest.mark.parametrize("name", ["poisson2d_30", "poisson3d_12"])
def test_pcg_flow_matches_the_oracle(flow, checker, name, jac):
    mat = MATS[name](); n = mat["n"]
    d = np.full(n, 4.0 if "2d" in name else 6.0)
    for b, tol, maxit in ((np.ones(n), 1e-10, 2000), (np.zeros(n), 1e-8, 50), (np.ones(n), 1e-12, 7)):
        rf, xf = run_flow(flow, "pcg", mat, b, tol, maxit, d if jac else None)
        ro, xo = run_oracle(checker, "pcg", mat, b, tol, maxit, d if jac else None)
        agree(rf, xf, ro, xo, tol)


@pytest.mark.parametrize("jac", [False, True], ids=["identity", "jacobi"])
@pytest.mark.parametrize("name", ["convdiff3d_12", "poisson2d_30"])
def test_bicgstab_flow_matches_the_oracle(flow, checker, name, jac):
    mat = MATS[name](); n = mat["n"]
    d = np.full(n, 4.0 if "2d" in name else 6.0)
    rng = np.random.default_rng(1)
    for b, tol, maxit in ((rng.standard_normal(n), 1e-8, 2000), (np.zeros(n), 1e-8, 50), (rng.standard_normal(n), 1e-13, 5)):
        rf, xf = run_flow(flow, "bicgstab", mat, b, tol, maxit, d if jac else None)
        ro, xo = run_oracle(checker, "bicgstab", mat, b, tol, maxit, d if jac else None)
        if name == "poisson2d_30" and ro[0] == 0 and ro[1] > 20:
            # rounding-chaotic on the Poisson matrices: flag and solution only
            assert rf[0] == ro[0] and abs(rf[1] - ro[1]) <= 0.35 * ro[1]
            assert np.max(np.abs(xf - xo)) <= 1e-5 * np.max(np.abs(xo))
        else:
            agree(rf, xf, ro, xo, tol)


@pytest.mark.parametrize("jac", [False, True], ids=["identity", "jacobi"])
@pytest.mark.parametrize("name,dim", [("convdiff3d_12", 10), ("convdiff3d_12", 30), ("poisson2d_30", 20)])
def test_gmres_flow_matches_the_oracle(flow, checker, name, dim, jac):
    mat = MATS[name](); n = mat["n"]
    d = np.full(n, 4.0 if "2d" in name else 6.0)
    rng = np.random.default_rng(2)
    for b, tol, maxit in ((rng.standard_normal(n), 1e-8, 3000), (np.zeros(n), 1e-8, 50), (rng.standard_normal(n), 1e-13, 17),
                          (np.ones(n), 1e-6, 3000)):
        rf, xf = run_flow(flow, "gmres", mat, b, tol, maxit, d if jac else None, dim)
        ro, xo = run_oracle(checker, "gmres", mat, b, tol, maxit, d if jac else None, dim)
        agree(rf, xf, ro, xo, tol)


# ---------------------------------------------------------------------------------------------
# Sharded flows: W emulated ranks, each with its row slab ([own | halo] column numbering from
# psp_dist_plan), the send maps of dist.build_sendmap, a landing zone per rank with
# (exchange number, value) words, and ONE scalar state (all ranks compute identical scalars).
# Checks what the single-rank flows cannot: that the boundary entries a rank recomputes and
# pushes at the START of a vector phase are exactly the values their owner stores in that
# phase, that they land in the halo of the right (double-buffered) vector, and that each
# product reads the buffer that was just exchanged.
class Ranks:
    def __init__(self, native, mat, world):
        from pysparse_b200 import dist as D
        val, col, ind = mat["csr"]
        self.n, self.world = mat["n"], world
        self.rowptr = D.row_partition(self.n, world)
        self.slabs, lists = [], []
        for r in range(world):
            r0, r1 = int(self.rowptr[r]), int(self.rowptr[r + 1])
            k0, k1 = int(ind[r0]), int(ind[r1])
            lind = (ind[r0:r1 + 1] - k0).astype(np.int32)
            cl, hg, ho, hr = D.plan(world, r, self.rowptr, lind, col[k0:k1])
            self.slabs.append(dict(nl=r1 - r0, nh=len(hg), ind=lind, col=cl, val=val[k0:k1], r0=r0, r1=r1))
            lists.append((r1 - r0, ho, hr))
        self.sendmaps = [D.build_sendmap(r, lists) for r in range(world)]
        self.zone = [np.full((s["nl"] + s["nh"], 2), -1.0) for s in self.slabs]      # (tag, value) per element
        self.pubs = 0

    def vec(self):
        return [np.zeros(s["nl"] + s["nh"]) for s in self.slabs]

    def spmv(self, r, x):
        s = self.slabs[r]
        prod = s["val"] * x[s["col"]]
        return np.add.reduceat(prod, s["ind"][:-1]) if len(prod) else np.zeros(s["nl"])

    def push(self, r, values_of_rows):
        rows, dst_rank, dst_off = self.sendmaps[r]
        vals = values_of_rows(rows)
        for k in range(len(rows)):
            self.zone[dst_rank[k]][dst_off[k]] = (self.pubs + 1, vals[k])

    def unpack(self, vecs):
        self.pubs += 1
        for r, s in enumerate(self.slabs):
            z = self.zone[r][s["nl"]:]
            assert np.all(z[:, 0] == self.pubs), "a halo entry of this exchange never arrived"
            vecs[r][s["nl"]:] = z[:, 1]


class Scalars:
    def __init__(self, lib, tol, maxit):
        lib.st_new.restype = C.c_void_p; lib.st_new.argtypes = [C.c_double, C.c_int, C.c_int]
        lib.st_epi.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        lib.st_v.restype = C.c_double; lib.st_v.argtypes = [C.c_void_p, C.c_int]
        lib.st_relres.restype = C.c_double; lib.st_relres.argtypes = [C.c_void_p]
        lib.st_int.argtypes = [C.c_void_p, C.c_int]; lib.st_free.argtypes = [C.c_void_p]
        lib.epi_id.argtypes = [C.c_char_p]
        self.lib, self.S = lib, lib.st_new(tol, maxit, 0)

    def id(self, name):
        v = self.lib.epi_id(name.encode()); assert v >= 0, name
        return v

    def epi(self, name, r0=0.0, r1=0.0, mx=0.0):
        self.lib.st_epi(self.S, self.id(name), r0, r1, mx)

    def v(self, name): return self.lib.st_v(self.S, self.id(name))
    done = property(lambda self: self.lib.st_int(self.S, 0))
    it = property(lambda self: self.lib.st_int(self.S, 3))
    phase = property(lambda self: self.lib.st_int(self.S, 4))
    stag = property(lambda self: self.lib.st_int(self.S, 5))

    def result(self):
        return self.lib.st_int(self.S, 1), self.lib.st_int(self.S, 2), self.lib.st_relres(self.S)


def sharded_bicgstab(native, lib, mat, b, tol, maxit, jac, world):
    """k_bicgstab_persist, push model, rank by rank."""
    R = Ranks(native, mat, world); S = Scalars(lib, tol, maxit)
    W = range(world); nl = [s["nl"] for s in R.slabs]
    sl = [slice(s["r0"], s["r1"]) for s in R.slabs]
    bl = [b[s] for s in sl]
    dinv = [None if jac is None else 1.0 / jac[s] for s in sl]
    ident = jac is None
    x = [np.zeros(k) for k in nl]; r = [None] * world; rhat = [None] * world
    v = [np.zeros(k) for k in nl]; t = [None] * world
    pa, pb, phat, s_, shat = R.vec(), R.vec(), R.vec(), R.vec(), R.vec()
    S.epi("EPI_BICG_NORMB", sum(float(bl[q] @ bl[q]) for q in W))
    p_cur, p_new = pa, pb
    for q in W:                                                   # x0 travels in the p buffer
        R.push(q, lambda rows, q=q: x[q][rows])
        p_cur[q][:nl[q]] = x[q]
    R.unpack(p_cur)
    for q in W:
        r[q] = R.spmv(q, p_cur[q])
    if S.phase == 1:
        x = [np.zeros(k) for k in nl]
    else:
        for q in W:
            r[q] = bl[q] + -r[q]; rhat[q] = r[q].copy()
    S.epi("EPI_BICG_RES0", sum(float(r[q] @ r[q]) for q in W) if S.phase != 1 else 0.0)
    while not S.done:
        x1 = p_new if ident else phat
        x2 = s_ if ident else shat
        beta, omega, first = S.v("B_BETA"), S.v("B_OMEGA"), S.it == 1
        for q in W:
            def pn_of(rows, q=q):
                pn = r[q][rows] if first else r[q][rows] + beta * (p_cur[q][rows] - omega * v[q][rows])
                return pn if ident else pn * dinv[q][rows]
            R.push(q, pn_of)                                      # from the OLD p, before the owner updates
            own = slice(0, nl[q])
            pn = r[q] if first else r[q] + beta * (p_cur[q][own] - omega * v[q])
            p_new[q][own] = pn
            if not ident:
                phat[q][own] = pn * dinv[q]
        R.unpack(x1)
        # the pushed entries are bit-identical to what their owners stored
        full = np.concatenate([x1[q][:nl[q]] for q in W])
        for q, sb in enumerate(R.slabs):
            rows, dr, do = R.sendmaps[q]
            for k in range(len(rows)):
                assert x1[dr[k]][do[k]] == full[sb["r0"] + rows[k]]
        for q in W:
            v[q] = R.spmv(q, x1[q])
        S.epi("EPI_BICG_ALPHA", sum(float(rhat[q] @ v[q]) for q in W))
        alpha = S.v("B_ALPHA")
        for q in W:
            def sn_of(rows, q=q):
                sn = r[q][rows] + (-alpha) * v[q][rows]
                return sn if ident else sn * dinv[q][rows]
            R.push(q, sn_of)
            sn = r[q] + (-alpha) * v[q]
            s_[q][:nl[q]] = sn
            if not ident:
                shat[q][:nl[q]] = sn * dinv[q]
        R.unpack(x2)
        for q in W:
            t[q] = R.spmv(q, x2[q])
        S.epi("EPI_BICG_OMEGA", sum(float(s_[q][:nl[q]] @ t[q]) for q in W), sum(float(t[q] @ t[q]) for q in W))
        alpha, omega = S.v("B_ALPHA"), S.v("B_OMEGA")
        for q in W:
            own = slice(0, nl[q])
            x[q] = x[q] + alpha * x1[q][own] + omega * x2[q][own]
            r[q] = s_[q][own] - omega * t[q]
        S.epi("EPI_BICG_RES", sum(float(r[q] @ r[q]) for q in W), sum(float(rhat[q] @ r[q]) for q in W))
        p_cur, p_new = p_new, p_cur
    return S.result(), np.concatenate(x)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("jac", [False, True], ids=["identity", "jacobi"])
def test_sharded_bicgstab_flow_matches_the_oracle(native, flow, checker, world, jac):
    mat = gallery.convdiff3d(10); n = mat["n"]
    d = np.full(n, 6.0)
    rng = np.random.default_rng(3)
    for b, tol, maxit in ((rng.standard_normal(n), 1e-8, 500), (np.zeros(n), 1e-8, 20), (rng.standard_normal(n), 1e-13, 4)):
        rf, xf = sharded_bicgstab(native, flow, mat, b, tol, maxit, d if jac else None, world)
        ro, xo = run_oracle(checker, "bicgstab", mat, b, tol, maxit, d if jac else None)
        agree(rf, xf, ro, xo, tol)
