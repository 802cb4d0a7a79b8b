"""CPU tests of the host-side analysis behind the diagonal-run device layout (psp_runs_analyze):
structure work must be bit-exact -- the matrix rebuilt from the run descriptors is the input,
entry for entry and in the reference's order of accumulation."""
import ctypes as C

import numpy as np
import pytest

from cases import build_matrix

_ip = C.POINTER(C.c_int)
GAP_FLAG = 1 << 30
_dp = C.POINTER(C.c_double)


def analyze(native, sss, m, ind, col, val, diag=None):
    lib = native.lib
    lib.psp_runs_analyze.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _ip, _ip, C.POINTER(C.c_longlong),
                                     C.POINTER(_ip), C.POINTER(_ip), C.POINTER(_dp)]
    lib.psp_runs_free.argtypes = [C.c_void_p]
    lib.psp_runs_free.restype = None
    ind = np.ascontiguousarray(ind, np.int32); col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    dg = np.ascontiguousarray(diag, np.float64) if diag is not None else np.zeros(1)
    nblk, nseg, nval = C.c_int(), C.c_int(), C.c_longlong()
    blk, seg, rval = _ip(), _ip(), _dp()
    native.check(lib.psp_runs_analyze(int(sss), m, ind.ctypes.data_as(_ip), col.ctypes.data_as(_ip), val.ctypes.data_as(_dp),
                                      dg.ctypes.data_as(_dp), C.byref(nblk), C.byref(nseg), C.byref(nval),
                                      C.byref(blk), C.byref(seg), C.byref(rval)))
    if nblk.value == 0:
        return None
    out = (np.ctypeslib.as_array(blk, (nblk.value, 4)).copy(), np.ctypeslib.as_array(seg, (nseg.value, 4)).copy(),
           np.ctypeslib.as_array(rval, (nval.value,)).copy())
    # bit 30 of a block's row count: some run of the block covers a bridged gap (only those blocks test
    # the loaded values for the marker) -- must be set exactly where a marker is in reach
    skip = out[2].view(np.uint64) == 0xFFF8C0DE5EED0002
    for b in range(nblk.value):
        reach = any(skip[k:k + ln].any() for k, _, _, ln in out[1][out[0][b, 2]:out[0][b, 3]])
        assert bool(out[0][b, 1] & GAP_FLAG) == reach, b
    out[0][:, 1] &= ~GAP_FLAG
    for p in (blk, seg, rval):
        lib.psp_runs_free(C.cast(p, C.c_void_p))
    return out


SKIP = np.array([0xFFF8C0DE5EED0002], dtype=np.uint64).view(np.float64)[0]     # kRunSkip of psp_runs.cuh


def is_skip(v):
    return np.array([v], dtype=np.float64).view(np.uint64)[0] == 0xFFF8C0DE5EED0002


def n_skips(rval):
    return int((np.asarray(rval).view(np.uint64) == 0xFFF8C0DE5EED0002).sum())


def rebuild_rows(m, blk, seg, rval):
    """rows[r] = [(col, val), ...] in the order the kernel adds them (the marker in a bridged gap
    of a run is what the kernel skips)"""
    rows = [[] for _ in range(m)]
    skip = np.asarray(rval).view(np.uint64) == 0xFFF8C0DE5EED0002
    for r0, nrows, s0, s1 in blk:
        for kpos, x0, t0, ln in seg[s0:s1]:
            assert 0 <= t0 and t0 + ln <= nrows and ln > 0
            for u in range(ln):
                if not skip[kpos + u]:
                    rows[r0 + t0 + u].append((x0 + u, rval[kpos + u]))
    return rows


@pytest.mark.parametrize("spec", [["poisson1d", 700], ["poisson2d", 37], ["poisson3d", 13], ["convdiff3d", 11],
                                  ["random_sym", 600, 5, 3], ["powerlaw", 500, 9, 5, 200]], ids=lambda s: "_".join(map(str, s)))
def test_csr_runs_rebuild_the_matrix_in_row_order(native, spec):
    mat = build_matrix(spec)
    val, col, ind = mat["csr"]
    m = len(ind) - 1
    blk, seg, rval = analyze(native, False, m, ind, col, val)
    assert np.array_equal(blk[:, 0], np.arange(0, m, 256)) and blk[:, 1].sum() == m
    assert len(rval) - n_skips(rval) == len(val)                    # every value stored exactly once
    rows = rebuild_rows(m, blk, seg, rval)
    for r in range(m):
        want = list(zip(col[ind[r]:ind[r + 1]].tolist(), val[ind[r]:ind[r + 1]].tolist()))
        assert rows[r] == want, r                                   # same entries, same (ascending-column) order
    if spec[0] in ("poisson2d", "poisson3d", "convdiff3d"):
        assert len(seg) * 8 <= len(val)                             # stencil matrices: long runs


@pytest.mark.parametrize("spec", [["poisson1d", 700], ["poisson2d", 37], ["poisson3d", 13], ["random_sym", 600, 5, 3],
                                  ["diag", 300]], ids=lambda s: "_".join(map(str, s)))
def test_sss_runs_store_each_value_once_and_keep_the_sss_order(native, spec):
    """sss_matvec (sss_mat.c:40-56) builds y[i] from the lower entries (ascending), the diagonal,
    then the upper entries (ascending): the rebuilt rows must list the symmetric expansion in
    exactly that order, while only nnz + n values are stored."""
    mat = build_matrix(spec)
    val, diag, col, ind = mat["sss"]
    n = len(diag)
    blk, seg, rval = analyze(native, True, n, ind, col, val, diag)
    assert len(rval) - n_skips(rval) == len(val) + n
    rows = rebuild_rows(n, blk, seg, rval)
    upper = [[] for _ in range(n)]
    for i in range(n):
        for k in range(ind[i], ind[i + 1]):
            upper[col[k]].append((i, val[k]))
    for r in range(n):
        want = list(zip(col[ind[r]:ind[r + 1]].tolist(), val[ind[r]:ind[r + 1]].tolist())) + [(r, diag[r])] + upper[r]
        assert rows[r] == want, r


def test_runs_bridge_the_one_row_gaps_of_grid_lines(native):
    """The -1 / +1 diagonals of a grid matrix have no entry at the first / last node of every grid
    line.  The analysis bridges such a gap with a marker value instead of ending the run there, so a
    block needs one descriptor per diagonal (one batch of loads in the kernel) however the lines
    fall into the 256-row blocks; a value that looks like the marker switches bridging off."""
    from pysparse_b200.tools import gallery as g
    mat = g.poisson2d(96)                                            # lines of 96 rows: 2 2/3 lines per block
    val, col, ind = mat["csr"]
    m = mat["n"]
    blk, seg, rval = analyze(native, False, m, ind, col, val)
    per_block = blk[:, 3] - blk[:, 2]
    assert n_skips(rval) > 0 and per_block.max() <= 7               # 5 diagonals; +-96 start / stop inside edge blocks
    interior = per_block[2:-2]
    assert (interior == 5).all()
    rows = rebuild_rows(m, blk, seg, rval)
    for r in range(0, m, 37):
        assert rows[r] == list(zip(col[ind[r]:ind[r + 1]].tolist(), val[ind[r]:ind[r + 1]].tolist()))
    # the same structure with a stored value equal to the marker: no bridging, nothing is skipped
    val2 = val.copy(); val2[5] = SKIP
    blk2, seg2, rval2 = analyze(native, False, m, ind, col, val2)
    assert len(rval2) == len(val2) and len(seg2) > len(seg)
    # symmetric form: lower runs bridge, the mirrored runs inherit the gaps
    sval, sdiag, scol, sind = mat["sss"]
    blk3, seg3, rval3 = analyze(native, True, m, sind, scol, sval, sdiag)
    assert n_skips(rval3) > 0 and (blk3[:, 3] - blk3[:, 2] <= 8).all()       # one batch of descriptors per block
    # 3-D: block-major storage would need 3 lower + diagonal + up to 6 mirrored slices per block; the
    # analysis then stores the values global-run-major and the interior blocks are back to 7
    mat = g.poisson3d(96, want_csr=False)
    sval, sdiag, scol, sind = mat["sss"]
    blk4, seg4, rval4 = analyze(native, True, mat["n"], sind, scol, sval, sdiag)
    per = blk4[:, 3] - blk4[:, 2]
    assert np.median(per) == 7 and per.max() <= 8
    assert len(rval4) - n_skips(rval4) == len(sval) + mat["n"]


@pytest.mark.parametrize("seed", range(6))
def test_runs_fuzz_banded_matrices_with_holes(native, seed):
    """Banded matrices with randomly missing entries (isolated holes, double holes, long holes, holes at
    block boundaries): whatever the analysis bridges or splits, and whichever value layout it picks for
    the symmetric form, the rebuilt rows must be the input rows, entry for entry, in row order."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(700, 1500))
    offs = sorted(set(int(o) for o in rng.choice(np.arange(1, 300), size=int(rng.integers(2, 6)), replace=False)))
    rows_lower = [[] for _ in range(n)]
    for o in offs:
        present = rng.random(n) > rng.choice([0.002, 0.02, 0.2])          # sparse to frequent holes
        for start in rng.integers(0, n, 3):                                # a few long holes
            present[start:start + int(rng.integers(3, 40))] = False
        present[::256] &= rng.random(len(present[::256])) > 0.5             # holes on block boundaries
        for i in range(o, n):
            if present[i]:
                rows_lower[i].append(i - o)
    ind = np.zeros(n + 1, dtype=np.int32)
    for i in range(n):
        rows_lower[i].sort()
        ind[i + 1] = ind[i] + len(rows_lower[i])
    col = np.array([c for r in rows_lower for c in r], dtype=np.int32)
    val = rng.standard_normal(len(col))
    diag = rng.standard_normal(n)
    # symmetric form
    blk, seg, rval = analyze(native, True, n, ind, col, val, diag)
    assert len(rval) - n_skips(rval) == len(val) + n
    rows = rebuild_rows(n, blk, seg, rval)
    upper = [[] for _ in range(n)]
    for i in range(n):
        for k in range(ind[i], ind[i + 1]):
            upper[col[k]].append((i, val[k]))
    for r in range(n):
        want = list(zip(col[ind[r]:ind[r + 1]].tolist(), val[ind[r]:ind[r + 1]].tolist())) + [(r, diag[r])] + upper[r]
        assert rows[r] == want, r
    # the same structure as a general (CSR) matrix: lower part only is enough for the scan
    blk2, seg2, rval2 = analyze(native, False, n, ind, col, val)
    if blk2 is not None:
        assert len(rval2) - n_skips(rval2) == len(val)
        rows2 = rebuild_rows(n, blk2, seg2, rval2)
        for r in range(n):
            assert rows2[r] == list(zip(col[ind[r]:ind[r + 1]].tolist(), val[ind[r]:ind[r + 1]].tolist())), r
