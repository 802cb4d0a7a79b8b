"""Host-side timings of the ll_mat assembly container (no GPU needed): the paths a user script
takes before the matrix reaches the GPU.  usage: llmat_bench.py [n]   (3-D Poisson n^3, default 128)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pysparse_b200.sparse import spmatrix
from pysparse_b200.tools import gallery as g, poisson

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mat = g.poisson3d(n, want_sss=False)
val, col, ind = mat["csr"]; N = mat["n"]
irow = np.repeat(np.arange(N, dtype=np.int64), np.diff(ind)); jcol = col.astype(np.int64)


def timed(label, fn, per=None):
    t = time.perf_counter(); out = fn(); dt = time.perf_counter() - t
    extra = f"  ({dt / per * 1e9:.0f} ns per entry)" if per else ""
    print(f"{label:58s} {dt:8.3f} s{extra}", flush=True)
    return out


print(f"3-D Poisson {n}^3: {N} rows, {len(val)} entries")
A = spmatrix.ll_mat(N, N, len(val))
timed("put(val, irow, jcol), row order", lambda: A.put(val, irow, jcol), len(val))
p = np.random.default_rng(0).permutation(len(val))
B = spmatrix.ll_mat(N, N, len(val))
timed("put(val, irow, jcol), random order", lambda: B.put(val[p], irow[p], jcol[p]), len(val))
C = timed("to_csr()", A.to_csr)
v2, c2, i2 = C.arrays()
assert np.array_equal(v2, val) and np.array_equal(c2, col) and np.array_equal(i2, ind)
S = spmatrix.ll_mat_sym(N, len(val))
low = jcol <= irow
timed("ll_mat_sym.put (lower triangle)", lambda: S.put(val[low], irow[low], jcol[low]), int(low.sum()))
timed("to_sss()", S.to_sss)
timed("find()", A.find, len(val))
timed("copy()", A.copy, len(val))
half = N // 2
sub = timed(f"A[:{half}, :{half}]  (slice, slice)", lambda: A[:half, :half], None)
idx = np.arange(0, N, 2)
timed("A[idx, idx]  (integer arrays, every other row/col)", lambda: A[idx, idx], None)
timed("A[0:half, 0:half] = sub  (block assignment)", lambda: A.__setitem__((slice(0, half), slice(0, half)), sub), None)
mask = np.ones(N, dtype=np.int64); mask[::3] = 0
D = A.copy()
timed("delete_rowcols(mask), every third removed", lambda: D.delete_rowcols(mask), None)
timed("row_scale + col_scale", lambda: (A.row_scale(np.ones(N)), A.col_scale(np.ones(N))), 2 * len(val))
m2 = min(n, 64)
P2 = g.poisson3d(m2, want_sss=False); pv, pc, pi = P2["csr"]; M2 = P2["n"]
E = spmatrix.ll_mat(M2, M2, len(pv)); E.put(pv, np.repeat(np.arange(M2, dtype=np.int64), np.diff(pi)), pc.astype(np.int64))
F = timed(f"matrixmultiply(A, A) for {m2}^3 ({len(pv)} entries)", lambda: spmatrix.matrixmultiply(E, E), None)
print(f"  product has {F.nnz} entries")
t = time.perf_counter(); L = poisson.poisson2d(300); dt = time.perf_counter() - t
print(f"{'poisson2d(300) element by element (A[i,j] = v from Python)':58s} {dt:8.3f} s  ({dt / L.nnz * 1e9:.0f} ns per entry)")
