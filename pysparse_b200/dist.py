"""Multi-GPU front end (one process per GPU, launched by torchrun).

Thin Python over the ``psp_dist_*`` C-ABI: row partition, host-side halo
planning, CUDA-IPC symmetric heap set-up (the 128-byte handle blobs are
all-gathered with ``torch.distributed``), sharded matrices and the sharded PCG.
``torch.distributed`` is plumbing only -- no tensor ever goes through it on the
data path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as P

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)

P.lib.psp_dist_plan.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _ip, _ip, _ip, _ip, _ip, _ip, _ip]
P.lib.psp_dist_init.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p]
P.lib.psp_dist_connect.argtypes = [C.c_void_p]
P.lib.psp_dist_info.argtypes = [_ip, _ip, C.POINTER(C.c_size_t), _ip]
P.lib.psp_dist_csr_create.argtypes = [_ip, C.c_int, _ip, _ip, _dp, C.POINTER(C.c_void_p)]
P.lib.psp_dist_set_sendmap.argtypes = [C.c_void_p, C.c_int, _ip, _ip, _ip]
P.lib.psp_dist_mat_destroy.argtypes = [C.c_void_p]
P.lib.psp_dist_mat_destroy.restype = None
P.lib.psp_dist_mat_info.argtypes = [C.c_void_p, _ip, _ip, C.POINTER(C.c_uint)]
P.lib.psp_dist_solve_dev.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                     C.c_int, C.c_int, _ip, _ip, _dp]

HANDLE_BYTES = 128


def row_partition(n_rows: int, world: int, align: int = 1) -> np.ndarray:
    """Contiguous near-equal row blocks; block boundaries are multiples of ``align``
    (e.g. one grid plane) whenever ``n_rows`` allows.  Returns rowptr[world+1] (int32)."""
    units = n_rows // align if n_rows % align == 0 else n_rows
    a = align if n_rows % align == 0 else 1
    base, extra = divmod(units, world)
    counts = np.array([base + (1 if r < extra else 0) for r in range(world)], dtype=np.int64) * a
    rowptr = np.zeros(world + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    assert rowptr[-1] == n_rows
    return rowptr.astype(np.int32)


def plan(world, rank, rowptr, ind, col_global):
    """Host-only halo planning of one slab (psp_dist_plan).  Returns
    (col_local, halo_gcol, halo_owner, halo_ridx)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    ind = np.ascontiguousarray(ind, dtype=np.int32)
    col_global = np.ascontiguousarray(col_global, dtype=np.int32)
    nl = len(ind) - 1
    nnz = int(ind[-1])
    col_local = np.empty(max(nnz, 1), dtype=np.int32)
    hg = np.empty(max(nnz, 1), dtype=np.int32)
    ho = np.empty(max(nnz, 1), dtype=np.int32)
    hr = np.empty(max(nnz, 1), dtype=np.int32)
    nh = C.c_int()
    P.check(P.lib.psp_dist_plan(world, rank, rowptr.ctypes.data_as(_ip), nl, ind.ctypes.data_as(_ip),
                                col_global.ctypes.data_as(_ip), col_local.ctypes.data_as(_ip),
                                hg.ctypes.data_as(_ip), ho.ctypes.data_as(_ip), hr.ctypes.data_as(_ip),
                                C.byref(nh)))
    k = nh.value
    return col_local[:nnz], hg[:k].copy(), ho[:k].copy(), hr[:k].copy()


def build_sendmap(rank: int, halo_lists):
    """Invert the halo lists: ``halo_lists[w] = (nl_w, halo_owner_w, halo_ridx_w)`` as
    returned by :func:`plan` on rank ``w``.  Returns this rank's send map
    (rows, dst_rank, dst_off): local row ``rows[s]`` goes to element ``dst_off[s]``
    (= nl_w + halo position) of the vector on rank ``dst_rank[s]``."""
    rows, dst_rank, dst_off = [], [], []
    for w, (nl_w, ho, hr) in enumerate(halo_lists):
        if w == rank:
            continue
        ho = np.asarray(ho)
        hr = np.asarray(hr)
        h = np.nonzero(ho == rank)[0]
        rows.append(hr[h].astype(np.int32))
        dst_rank.append(np.full(len(h), w, dtype=np.int32))
        dst_off.append((nl_w + h).astype(np.int32))
    if not rows:
        z = np.zeros(0, dtype=np.int32)
        return z, z.copy(), z.copy()
    return np.concatenate(rows), np.concatenate(dst_rank), np.concatenate(dst_off)


def init(rank: int, world: int, slot_elems: int, nslots: int = 6, group=None) -> None:
    """Create the symmetric heap and connect all ranks of the node (collective)."""
    import torch
    import torch.distributed as dist

    blob = (C.c_ubyte * HANDLE_BYTES)()
    P.check(P.lib.psp_dist_init(rank, world, slot_elems, nslots, C.cast(blob, C.c_void_p)))
    mine = bytes(blob)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=group)
    else:
        gathered = [mine]
    allb = b"".join(gathered)
    buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
    P.check(P.lib.psp_dist_connect(C.cast(buf, C.c_void_p)))
    if world > 1:
        dist.barrier(group=group)


def shutdown() -> None:
    P.check(P.lib.psp_dist_shutdown())


class ShardedMatrix:
    """This rank's row slab of a CSR matrix (global column indices)."""

    def __init__(self, rowptr, ind, col_global, val, group=None, sendmap=True):
        """Collective when ``sendmap`` is true and the world has more than one rank: the
        ranks exchange their halo lists (host metadata, through ``group``) so that each
        one learns which of its rows the others need -- the send map of the push-model
        halo exchange (psp_dist_set_sendmap)."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        ind = np.ascontiguousarray(ind, dtype=np.int32)
        col_global = np.ascontiguousarray(col_global, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        p = C.c_void_p()
        P.check(P.lib.psp_dist_csr_create(rowptr.ctypes.data_as(_ip), len(val), ind.ctypes.data_as(_ip),
                                          col_global.ctypes.data_as(_ip), val.ctypes.data_as(_dp), C.byref(p)))
        self.ptr = p.value
        nl, nh, mask = C.c_int(), C.c_int(), C.c_uint()
        P.check(P.lib.psp_dist_mat_info(self.ptr, C.byref(nl), C.byref(nh), C.byref(mask)))
        self.nl, self.nhalo, self.neighbour_mask = nl.value, nh.value, mask.value
        self.nsend = 0
        rank, world = C.c_int(), C.c_int()
        P.check(P.lib.psp_dist_info(C.byref(rank), C.byref(world), None, None))
        if sendmap and world.value > 1:
            import torch.distributed as dist

            _, _, ho, hr = plan(world.value, rank.value, rowptr, ind, col_global)
            gathered = [None] * world.value
            dist.all_gather_object(gathered, (self.nl, ho, hr), group=group)
            rows, dst_rank, dst_off = build_sendmap(rank.value, gathered)
            self.set_sendmap(rows, dst_rank, dst_off)

    def set_sendmap(self, rows, dst_rank, dst_off):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        dst_rank = np.ascontiguousarray(dst_rank, dtype=np.int32)
        dst_off = np.ascontiguousarray(dst_off, dtype=np.int32)
        P.check(P.lib.psp_dist_set_sendmap(self.ptr, len(rows), rows.ctypes.data_as(_ip),
                                           dst_rank.ctypes.data_as(_ip), dst_off.ctypes.data_as(_ip)))
        self.nsend = len(rows)

    def __del__(self):
        try:
            if self.ptr:
                P.lib.psp_dist_mat_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def solve(solver: str, A: ShardedMatrix, db: P.DeviceVector, dx: P.DeviceVector, tol: float, maxit: int,
          K: P.Precon | None = None, dim: int = 20):
    """Sharded solve ("pcg", "bicgstab", "gmres") on device-resident local slices; collective.
    Returns (info, iter, relres)."""
    info, it, rr = C.c_int(), C.c_int(), C.c_double()
    P.check(P.lib.psp_dist_solve_dev(P.SOLVERS[solver], A.ptr, K.ptr if K is not None else None, db.ptr, dx.ptr,
                                     tol, maxit, dim, C.byref(info), C.byref(it), C.byref(rr)))
    return info.value, it.value, rr.value


def pcg(A, db, dx, tol, maxit, K=None):
    return solve("pcg", A, db, dx, tol, maxit, K)
