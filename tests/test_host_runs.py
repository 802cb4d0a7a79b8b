"""CPU tests of the host-side analysis behind the diagonal-run device layout (psp_runs_analyze):
structure work must be bit-exact -- the matrix rebuilt from the run descriptors is the input,
entry for entry and in the reference's order of accumulation."""
import ctypes as C

import numpy as np
import pytest

from cases import build_matrix

_ip = C.POINTER(C.c_int)
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
    for p in (blk, seg, rval):
        lib.psp_runs_free(C.cast(p, C.c_void_p))
    return out


def rebuild_rows(m, blk, seg, rval):
    """rows[r] = [(col, val), ...] in the order the kernel adds them"""
    rows = [[] for _ in range(m)]
    for r0, nrows, s0, s1 in blk:
        for kpos, x0, t0, ln in seg[s0:s1]:
            assert 0 <= t0 and t0 + ln <= nrows and ln > 0
            for u in range(ln):
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
    assert len(rval) == len(val)                                    # every value stored exactly once
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
    assert len(rval) == len(val) + n
    rows = rebuild_rows(n, blk, seg, rval)
    upper = [[] for _ in range(n)]
    for i in range(n):
        for k in range(ind[i], ind[i + 1]):
            upper[col[k]].append((i, val[k]))
    for r in range(n):
        want = list(zip(col[ind[r]:ind[r + 1]].tolist(), val[ind[r]:ind[r + 1]].tolist())) + [(r, diag[r])] + upper[r]
        assert rows[r] == want, r
