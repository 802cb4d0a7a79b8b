"""CPU tests of the multi-GPU host logic (no GPU needed): row partition and halo
maps (bit-exact integer work), checked single-process and with a world_size-2
gloo job that emulates the sharded SpMV with the planned maps."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from pysparse_b200.tools import gallery

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition(native):
    from pysparse_b200 import dist as D
    rp = D.row_partition(10 ** 3, 4, align=100)
    assert rp.tolist() == [0, 300, 600, 800, 1000] and rp.dtype == np.int32
    assert D.row_partition(7, 3).tolist() == [0, 3, 5, 7]
    assert D.row_partition(256 ** 3, 8, align=256 ** 2).tolist() == [i * 32 * 65536 for i in range(9)]


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_plan_vs_numpy_model(native, world):
    from pysparse_b200 import dist as D
    mat = gallery.powerlaw(400, mean_nnz=7, seed=3, max_len=60)
    val, col, ind = mat["csr"]
    n = mat["n"]
    rowptr = D.row_partition(n, world)
    x = np.random.default_rng(0).standard_normal(n)
    y_ref = np.zeros(n)
    for i in range(n):
        y_ref[i] = sum(val[k] * x[col[k]] for k in range(ind[i], ind[i + 1]))
    for r in range(world):
        r0, r1 = rowptr[r], rowptr[r + 1]
        k0, k1 = ind[r0], ind[r1]
        lind = (ind[r0:r1 + 1] - k0).astype(np.int32)
        cl, hg, ho, hr = D.plan(world, r, rowptr, lind, col[k0:k1])
        # model: halo = sorted unique non-local columns, owners by the partition
        gc = col[k0:k1]
        exp_halo = np.unique(gc[(gc < r0) | (gc >= r1)])
        assert np.array_equal(hg, exp_halo)
        exp_owner = np.searchsorted(rowptr, exp_halo, side="right") - 1
        assert np.array_equal(ho, exp_owner) and np.array_equal(hr, exp_halo - rowptr[exp_owner])
        assert not np.any(ho == r)
        # local numbering: own columns first, halo columns after, ascending global order
        xl = np.concatenate([x[r0:r1], x[hg]])
        assert np.array_equal(xl[cl], x[gc])
        nl = r1 - r0
        own = (gc >= r0) & (gc < r1)
        assert np.array_equal(cl[own], gc[own] - r0) and np.all(cl[~own] >= nl)
        y = np.array([sum(val[k0 + k] * xl[cl[k]] for k in range(lind[i], lind[i + 1])) for i in range(nl)])
        assert np.array_equal(y, y_ref[r0:r1])


def test_plan_poisson3d_halo_is_one_plane_per_neighbour(native):
    from pysparse_b200 import dist as D
    n, world = 12, 4
    rowptr = D.row_partition(n ** 3, world, align=n * n)
    for r in range(world):
        slab = gallery.poisson3d(n, want_sss=False, rows=(rowptr[r], rowptr[r + 1]))
        _, hg, ho, _ = D.plan(world, r, rowptr, slab["csr"][2], slab["csr"][1])
        nb = [p for p in (r - 1, r + 1) if 0 <= p < world]
        assert len(hg) == len(nb) * n * n and sorted(set(ho.tolist())) == nb


@pytest.mark.parametrize("world", [2, 3, 5])
def test_sendmap_push_fills_every_halo_entry(native, world):
    """The send map (push model) is the exact inverse of the halo lists: pushing with it
    writes every rank's halo, each entry once, with the value the pull model would fetch."""
    from pysparse_b200 import dist as D
    mat = gallery.powerlaw(300, mean_nnz=6, seed=5, max_len=40)
    val, col, ind = mat["csr"]
    n = mat["n"]
    rowptr = D.row_partition(n, world)
    x = np.random.default_rng(1).standard_normal(n)
    lists, halos = [], []
    for r in range(world):
        r0, r1 = rowptr[r], rowptr[r + 1]
        k0, k1 = ind[r0], ind[r1]
        _, hg, ho, hr = D.plan(world, r, rowptr, (ind[r0:r1 + 1] - k0).astype(np.int32), col[k0:k1])
        lists.append((int(r1 - r0), ho, hr))
        halos.append(hg)
    vecs = [np.concatenate([x[rowptr[r]:rowptr[r + 1]], np.full(len(halos[r]), np.nan)]) for r in range(world)]
    hits = [np.zeros(len(v), dtype=int) for v in vecs]
    for r in range(world):
        rows, dst_rank, dst_off = D.build_sendmap(r, lists)
        assert rows.dtype == dst_rank.dtype == dst_off.dtype == np.int32
        assert not np.any(dst_rank == r)
        for s in range(len(rows)):
            vecs[dst_rank[s]][dst_off[s]] = vecs[r][rows[s]]
            hits[dst_rank[s]][dst_off[s]] += 1
    for r in range(world):
        nl = lists[r][0]
        assert np.all(hits[r][:nl] == 0) and np.all(hits[r][nl:] == 1)
        assert np.array_equal(vecs[r][nl:], x[halos[r]])


WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["PSP_ROOT"])
from pysparse_b200 import dist as D
from pysparse_b200.tools import gallery
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 10
N = n ** 3
rowptr = D.row_partition(N, world, align=n * n)
r0, r1 = int(rowptr[rank]), int(rowptr[rank + 1])
slab = gallery.convdiff3d(n, rows=(r0, r1))
val, col, ind = slab["csr"]
cl, hg, ho, hr = D.plan(world, rank, rowptr, ind, col)
x = np.random.default_rng(42).standard_normal(N)      # same on every rank
xl = x[r0:r1].copy()
# halo exchange emulated over gloo: tell each owner which of its entries we need, receive them
need = [None] * world
dist.all_gather_object(need, {int(o): hr[ho == o].tolist() for o in set(ho.tolist())})
halo = np.empty(len(hg))
reqs = []
for o in sorted(set(ho.tolist())):
    buf = torch.empty(int((ho == o).sum()), dtype=torch.float64)
    reqs.append((o, buf, dist.irecv(buf, src=o)))
for peer, wants in enumerate(need):
    if rank in wants:
        dist.send(torch.from_numpy(xl[np.array(wants[rank], dtype=np.int64)]), dst=peer)
for o, buf, req in reqs:
    req.wait()
    halo[ho == o] = buf.numpy()
xfull = np.concatenate([xl, halo])
nl = r1 - r0
y = np.zeros(nl)
for i in range(nl):
    s = 0.0
    for k in range(ind[i], ind[i + 1]):
        s += val[k] * xfull[cl[k]]
    y[i] = s
# reference: the same rows of the global product, same summation order
full = gallery.convdiff3d(n)
fv, fc, fi = full["csr"]
yr = np.zeros(nl)
for i in range(nl):
    s = 0.0
    for k in range(fi[r0 + i], fi[r0 + i + 1]):
        s += fv[k] * x[fc[k]]
    yr[i] = s
ok = bool(np.array_equal(y, yr))
# push model: invert the all-gathered halo lists into send maps (what ShardedMatrix does) and
# deliver every boundary entry to element dst_off of the neighbour's vector
lists = [None] * world
dist.all_gather_object(lists, (nl, ho, hr))
rows, dst_rank, dst_off = D.build_sendmap(rank, lists)
xpush = np.concatenate([xl, np.full(len(hg), np.nan)])
recv = []
for src in range(world):
    if src == rank:
        continue
    cnt = int(np.sum(np.asarray(lists[rank][1]) == src))          # my halo entries owned by src
    if cnt:
        buf = torch.empty(2 * cnt, dtype=torch.float64)
        recv.append((buf, dist.irecv(buf, src=src)))
for dst in range(world):
    sel = dst_rank == dst
    if dst != rank and sel.any():
        payload = np.concatenate([dst_off[sel].astype(np.float64), xl[rows[sel]]])   # (offsets, values)
        dist.send(torch.from_numpy(payload), dst=dst)
for buf, req in recv:
    req.wait()
    a = buf.numpy(); c = len(a) // 2
    xpush[a[:c].astype(np.int64)] = a[c:]
ok = ok and bool(np.array_equal(xpush, xfull))
flags = [None] * world
dist.all_gather_object(flags, ok)
dist.destroy_process_group()
sys.exit(0 if all(flags) else 1)
'''


def test_world_size_2_gloo_sharded_spmv(native, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, PSP_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
