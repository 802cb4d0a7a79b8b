// psp_dist.cu -- multi-GPU layer: row-slab partition, halo maps, peer-memory
// (CUDA IPC over NVLink) symmetric heap, halo pull kernel.
//
// Not in the reference (PySparse is single-process, serial): SURVEY.md 8(e).
// One process per GPU (torchrun).  Every rank owns a contiguous block of rows
// of A and the matching slices of all vectors.  Work vectors live in a
// "symmetric heap" that every peer maps through CUDA IPC, so
//   * the x-halo is PULLED by the consumer straight out of the owner's vector
//     (k_halo_pull: NVLink loads, one small kernel, no packing, no NCCL),
//     after the owner has published the vector (release store of a counter
//     that the consumer polls with acquire loads);
//   * dot products / norms are all-reduced inside the reduction kernels through
//     per-rank mailboxes (dist_allreduce in psp_common.cuh).
// The partition / halo maps are integer work done on the host (psp_dist_plan,
// testable without a GPU) and must be bit-exact.
#include <algorithm>
#include "psp_mat.cuh"
#include "psp_vec.cuh"
#include "psp_dist.cuh"

namespace psp {

static DistHost g_dist;
static thread_local DistHost *t_dist = nullptr;    // set in the rank threads of the in-process multi-GPU layer
DistHost &dist() { return t_dist ? *t_dist : g_dist; }
void set_thread_ctx(Ctx *c);
void bind_thread(Ctx *c, DistHost *d) { set_thread_ctx(c); t_dist = d; }

__global__ void __launch_bounds__(kThreads) k_halo_pull(State *S, size_t slot_off, const int *owner,
                                                        const int *ridx, int nhalo, int nl, double *xloc,
                                                        unsigned nbmask, unsigned *ticket, int gate) {
  if (gated_out(S, gate)) return;
  DistDev *dd = S->dist;
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    const unsigned long long want = dd->pulls_done + 1;
    for (int w = 0; w < dd->world; w++)
      if ((nbmask >> w) & 1u)
        while (ld_acquire_sys(&dd->ctl[w]->pub) < want) { }
  }
  __syncthreads();
  const int stride = gridDim.x * blockDim.x;
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < nhalo; h += stride)
    xloc[nl + h] = ld_relaxed_sys_f64(dd->heap[owner[h]] + slot_off + ridx[h]);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) { *ticket = 0u; dd->pulls_done += 1; }
  }
}

int dist_pull(DistMat *dm, double *xloc, int gate) {
  Ctx &c = ctx();
  DistHost &d = dist();
  int grid = (dm->nhalo + kThreads - 1) / kThreads;
  if (grid < 1) grid = 1;
  if (grid > 2 * c.sms) grid = 2 * c.sms;
  size_t slot_off = (size_t)(xloc - d.heap);
  k_halo_pull<<<grid, kThreads, 0, c.stream>>>(c.state, slot_off, dm->d_owner, dm->d_ridx, dm->nhalo, dm->nl,
                                              xloc, dm->nbmask, c.tickets, gate);
  PSP_LAUNCH_CHECK();
  return PSP_OK;
}

}  // namespace psp

using namespace psp;

// ------------------------------------------------------------------ planning (host only)
extern "C" int psp_dist_plan(int world, int rank, const int *rowptr, int nl, const int *ind,
                             const int *col_global, int *col_local, int *halo_gcol, int *halo_owner,
                             int *halo_ridx, int *nhalo_out) {
  if (world < 1 || rank < 0 || rank >= world || !rowptr || !ind || !nhalo_out) {
    set_error("psp_dist_plan: bad arguments");
    return PSP_ERR_ARG;
  }
  const int r0 = rowptr[rank], r1 = rowptr[rank + 1];
  if (r1 - r0 != nl) { set_error("psp_dist_plan: nl does not match the row partition"); return PSP_ERR_ARG; }
  const int nglobal = rowptr[world];
  const long long nnz = ind[nl];
  std::vector<int> halo;
  for (long long k = 0; k < nnz; k++) {
    const int c = col_global[k];
    if (c < 0 || c >= nglobal) { set_error("psp_dist_plan: column %d out of range", c); return PSP_ERR_STRUCTURE; }
    if (c < r0 || c >= r1) halo.push_back(c);
  }
  std::sort(halo.begin(), halo.end());
  halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
  const int nh = (int)halo.size();
  *nhalo_out = nh;
  for (int h = 0; h < nh; h++) {
    const int c = halo[h];
    int o = (int)(std::upper_bound(rowptr, rowptr + world + 1, c) - rowptr) - 1;
    if (halo_gcol) halo_gcol[h] = c;
    if (halo_owner) halo_owner[h] = o;
    if (halo_ridx) halo_ridx[h] = c - rowptr[o];
  }
  if (col_local) {
    for (long long k = 0; k < nnz; k++) {
      const int c = col_global[k];
      if (c >= r0 && c < r1) col_local[k] = c - r0;
      else col_local[k] = nl + (int)(std::lower_bound(halo.begin(), halo.end(), c) - halo.begin());
    }
  }
  return PSP_OK;
}

// ------------------------------------------------------------------ symmetric heap / IPC
// First element of the halo landing zone: after the work vectors, rounded up to an even
// element so that every (2e, 2e+1) word pair is 16-byte aligned (it is read with one
// 16-byte load).
static size_t ll_offset(size_t slot_elems, int nslots) {
  return (slot_elems * (size_t)nslots + 1) & ~(size_t)1;
}

extern "C" int psp_dist_init(int rank, int world, size_t slot_elems, int nslots, void *handles_out) {
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !handles_out) {
    set_error("psp_dist_init: bad arguments (world must be 1..%d)", kMaxRanks);
    return PSP_ERR_ARG;
  }
  PSP_TRY(ensure_init());
  DistHost &d = dist();
  if (d.ready) { set_error("psp_dist_init: already initialised"); return PSP_ERR_DIST; }
  d.rank = rank; d.world = world; d.slot_elems = slot_elems; d.nslots = nslots;
  // nslots work vectors, then the flag-in-data halo landing zone of the persistent kernels:
  // two 8-byte words per vector element (see DistDev::ll_off)
  const size_t heap_bytes = sizeof(double) * (ll_offset(slot_elems, nslots) + 2 * slot_elems);
  PSP_CUDA(cudaMalloc((void **)&d.heap, heap_bytes ? heap_bytes : 8));
  PSP_CUDA(cudaMemset(d.heap, 0, heap_bytes ? heap_bytes : 8));
  PSP_CUDA(cudaMalloc((void **)&d.ctl, sizeof(DistCtl)));
  PSP_CUDA(cudaMemset(d.ctl, 0, sizeof(DistCtl)));
  cudaIpcMemHandle_t h[2];
  PSP_CUDA(cudaIpcGetMemHandle(&h[0], d.heap));
  PSP_CUDA(cudaIpcGetMemHandle(&h[1], d.ctl));
  memcpy(handles_out, h, sizeof h);
  return PSP_OK;
}

extern "C" int psp_dist_connect(const void *all_handles) {
  DistHost &d = dist();
  if (!d.heap || !all_handles) { set_error("psp_dist_connect: call psp_dist_init first"); return PSP_ERR_DIST; }
  const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)all_handles;
  DistDev dd;
  memset(&dd, 0, sizeof dd);
  dd.rank = d.rank; dd.world = d.world;
  dd.ll_off = ll_offset(d.slot_elems, d.nslots);
  for (int w = 0; w < d.world; w++) {
    if (w == d.rank) { d.peer_heap[w] = d.heap; d.peer_ctl[w] = d.ctl; }
    else {
      cudaError_t e = cudaIpcOpenMemHandle((void **)&d.peer_heap[w], h[2 * w], cudaIpcMemLazyEnablePeerAccess);
      if (e == cudaSuccess) e = cudaIpcOpenMemHandle((void **)&d.peer_ctl[w], h[2 * w + 1], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("psp_dist_connect: cannot map rank %d's memory over CUDA IPC: %s", w, cudaGetErrorString(e));
        cudaGetLastError();
        return PSP_ERR_DIST;
      }
    }
    dd.heap[w] = d.peer_heap[w];
    dd.ctl[w] = d.peer_ctl[w];
  }
  PSP_CUDA(cudaMalloc((void **)&d.dev, sizeof(DistDev)));
  PSP_CUDA(cudaMemcpy(d.dev, &dd, sizeof dd, cudaMemcpyHostToDevice));
  d.ready = true;
  return PSP_OK;
}

extern "C" int psp_dist_shutdown(void) {
  DistHost &d = dist();
  if (!d.heap) return PSP_OK;
  cudaDeviceSynchronize();
  for (int w = 0; w < d.world && !d.in_process; w++)
    if (w != d.rank) {
      if (d.peer_heap[w]) cudaIpcCloseMemHandle(d.peer_heap[w]);
      if (d.peer_ctl[w]) cudaIpcCloseMemHandle(d.peer_ctl[w]);
    }
  cudaFree(d.heap); cudaFree(d.ctl); cudaFree(d.dev);
  d = DistHost();
  return PSP_OK;
}

// ---- in-process variants (psp_multi.cu: one thread per GPU in ONE process; peer access instead
// of CUDA IPC -- with unified addressing the peers' allocations are directly addressable)
namespace psp {

int dist_local_alloc(int rank, int world, size_t slot_elems, int nslots) {
  DistHost &d = dist();
  if (d.heap) { set_error("dist_local_alloc: heap already exists"); return PSP_ERR_DIST; }
  d.rank = rank; d.world = world; d.slot_elems = slot_elems; d.nslots = nslots; d.in_process = true;
  const size_t heap_bytes = sizeof(double) * (ll_offset(slot_elems, nslots) + 2 * slot_elems);
  if (cudaMalloc((void **)&d.heap, heap_bytes ? heap_bytes : 8) != cudaSuccess) {
    cudaGetLastError();
    set_error("multi-GPU: cannot allocate the %zu-byte symmetric heap on device %d", heap_bytes, ctx().device);
    return PSP_ERR_NOMEM;
  }
  PSP_CUDA(cudaMemset(d.heap, 0, heap_bytes ? heap_bytes : 8));
  PSP_CUDA(cudaMalloc((void **)&d.ctl, sizeof(DistCtl)));
  PSP_CUDA(cudaMemset(d.ctl, 0, sizeof(DistCtl)));
  return PSP_OK;
}

// all[w]: the DistHost of rank w (its heap / ctl pointers are valid on every device)
int dist_local_connect(DistHost *const *all) {
  DistHost &d = dist();
  DistDev dd;
  memset(&dd, 0, sizeof dd);
  dd.rank = d.rank; dd.world = d.world;
  dd.ll_off = ll_offset(d.slot_elems, d.nslots);
  for (int w = 0; w < d.world; w++) {
    d.peer_heap[w] = all[w]->heap; d.peer_ctl[w] = all[w]->ctl;
    dd.heap[w] = d.peer_heap[w]; dd.ctl[w] = d.peer_ctl[w];
  }
  PSP_CUDA(cudaMalloc((void **)&d.dev, sizeof(DistDev)));
  PSP_CUDA(cudaMemcpy(d.dev, &dd, sizeof dd, cudaMemcpyHostToDevice));
  d.ready = true;
  return PSP_OK;
}

int dist_local_free() {
  DistHost &d = dist();
  if (!d.heap) return PSP_OK;
  cudaDeviceSynchronize();
  cudaFree(d.heap); cudaFree(d.ctl); cudaFree(d.dev);
  d = DistHost();
  return PSP_OK;
}

}  // namespace psp

extern "C" int psp_dist_info(int *rank, int *world, size_t *slot_elems, int *nslots) {
  DistHost &d = dist();
  if (rank) *rank = d.rank;
  if (world) *world = d.world;
  if (slot_elems) *slot_elems = d.slot_elems;
  if (nslots) *nslots = d.nslots;
  return d.ready ? PSP_OK : PSP_ERR_DIST;
}

// ------------------------------------------------------------------ distributed matrix
namespace psp {
int csr_create(int, int, long long, const int *, const int *, const double *, psp_mat **, bool, int, const int *, int);
}

namespace psp {
int dist_csr_finish(DistMat *dm, const std::vector<int> &owner, const std::vector<int> &ridx, int nhalo);
// Slab of rank `rank`: local column numbering, device upload, halo tables.  Independent of the
// symmetric heap (the caller checks that nl + nhalo fits a slot).  owner_out / ridx_out: the
// halo lists (host copies) for building the send maps.
int dist_csr_build(int world, int rank, const int *rowptr, int nnz, const int *ind, const int *col_global,
                   const double *val, DistMat **out, std::vector<int> *owner_out, std::vector<int> *ridx_out) {
  const int nl = rowptr[rank + 1] - rowptr[rank];
  if (ind[nl] != nnz) { set_error("psp_dist_csr_create: ind[nl] != nnz"); return PSP_ERR_ARG; }
  std::vector<int> col_local((size_t)(nnz > 0 ? nnz : 1)), owner((size_t)(nnz > 0 ? nnz : 1)), ridx((size_t)(nnz > 0 ? nnz : 1));
  int nhalo = 0;
  PSP_TRY(psp_dist_plan(world, rank, rowptr, nl, ind, col_global, col_local.data(), nullptr, owner.data(),
                        ridx.data(), &nhalo));
  DistMat *dm = new DistMat();
  dm->nl = nl; dm->nhalo = nhalo; dm->nglobal = rowptr[world];
  int st = csr_create(nl, nl + nhalo, nnz, ind, col_local.data(), val, &dm->local, false, nl, col_global, rowptr[rank]);
  if (st != PSP_OK) { delete dm; return st; }
  st = dist_csr_finish(dm, owner, ridx, nhalo);
  if (st != PSP_OK) return st;
  if (owner_out) owner_out->assign(owner.begin(), owner.begin() + nhalo);
  if (ridx_out) ridx_out->assign(ridx.begin(), ridx.begin() + nhalo);
  *out = dm;
  return PSP_OK;
}
}  // namespace psp

extern "C" int psp_dist_csr_create(const int *rowptr, int nnz, const int *ind, const int *col_global,
                                   const double *val, psp_dmat **out) {
  DistHost &d = dist();
  if (!d.ready) { set_error("psp_dist_csr_create: distributed layer not initialised"); return PSP_ERR_DIST; }
  if (!rowptr || !ind || !out) { set_error("psp_dist_csr_create: NULL argument"); return PSP_ERR_ARG; }
  DistMat *dm = nullptr;
  PSP_TRY(dist_csr_build(d.world, d.rank, rowptr, nnz, ind, col_global, val, &dm, nullptr, nullptr));
  if ((size_t)dm->nl + (size_t)dm->nhalo > d.slot_elems) {
    set_error("psp_dist_csr_create: local rows + halo (%d + %d) exceed the heap slot size %zu", dm->nl, dm->nhalo, d.slot_elems);
    psp_dist_mat_destroy((psp_dmat *)dm);
    return PSP_ERR_ARG;
  }
  *out = (psp_dmat *)dm;
  return PSP_OK;
}

namespace psp {
int dist_csr_finish(DistMat *dm, const std::vector<int> &owner, const std::vector<int> &ridx, int nhalo) {
  dm->nbmask = 0;
  for (int h = 0; h < nhalo; h++) dm->nbmask |= 1u << owner[h];
  const size_t hb = sizeof(int) * (size_t)(nhalo > 0 ? nhalo : 1);
  if (cudaMalloc((void **)&dm->d_owner, hb) != cudaSuccess || cudaMalloc((void **)&dm->d_ridx, hb) != cudaSuccess) {
    cudaGetLastError();
    psp_dist_mat_destroy((psp_dmat *)dm);
    set_error("psp_dist_csr_create: out of device memory");
    return PSP_ERR_NOMEM;
  }
  if (nhalo) {
    PSP_CUDA(cudaMemcpy(dm->d_owner, owner.data(), sizeof(int) * (size_t)nhalo, cudaMemcpyHostToDevice));
    PSP_CUDA(cudaMemcpy(dm->d_ridx, ridx.data(), sizeof(int) * (size_t)nhalo, cudaMemcpyHostToDevice));
  }
  return PSP_OK;
}
}  // namespace psp

// Send map for the push model: entry s says "local row rows[s] is halo column number
// off[s] - nl(dst) of rank dst_rank[s]", i.e. it has to be stored at element off[s] of the
// multiplied vector in that rank's heap.  Built by the launcher from the all-gathered
// halo lists (pysparse_b200/dist.py); integer work, verified against psp_dist_plan.
extern "C" int psp_dist_set_sendmap(psp_dmat *m, int nsend, const int *rows, const int *dst_rank,
                                    const int *dst_off) {
  DistMat *dm = (DistMat *)m;
  DistHost &d = dist();
  if (!dm || nsend < 0 || (nsend > 0 && (!rows || !dst_rank || !dst_off))) {
    set_error("psp_dist_set_sendmap: bad arguments");
    return PSP_ERR_ARG;
  }
  for (int s = 0; s < nsend; s++) {
    if (rows[s] < 0 || rows[s] >= dm->nl || dst_rank[s] < 0 || dst_rank[s] >= d.world || dst_rank[s] == d.rank ||
        dst_off[s] < 0 || (size_t)dst_off[s] >= d.slot_elems) {
      set_error("psp_dist_set_sendmap: entry %d out of range", s);
      return PSP_ERR_ARG;
    }
  }
  if (!d.ready) { set_error("psp_dist_set_sendmap: distributed layer not connected"); return PSP_ERR_DIST; }
  cudaFree(dm->d_send_row); cudaFree(dm->d_send_dst);
  dm->d_send_row = nullptr; dm->d_send_dst = nullptr;
  const size_t cnt = (size_t)(nsend > 0 ? nsend : 1);
  if (cudaMalloc((void **)&dm->d_send_row, sizeof(int) * cnt) != cudaSuccess ||
      cudaMalloc((void **)&dm->d_send_dst, sizeof(unsigned long long *) * cnt) != cudaSuccess) {
    cudaGetLastError();
    set_error("psp_dist_set_sendmap: out of device memory");
    return PSP_ERR_NOMEM;
  }
  if (nsend) {
    // element e of the neighbour's vector <-> words 2e, 2e+1 of its landing zone
    std::vector<unsigned long long *> dst((size_t)nsend);
    const size_t ll_off = ll_offset(d.slot_elems, d.nslots);
    for (int s = 0; s < nsend; s++)
      dst[s] = reinterpret_cast<unsigned long long *>(d.peer_heap[dst_rank[s]] + ll_off) + 2 * (size_t)dst_off[s];
    PSP_CUDA(cudaMemcpy(dm->d_send_row, rows, sizeof(int) * (size_t)nsend, cudaMemcpyHostToDevice));
    PSP_CUDA(cudaMemcpy(dm->d_send_dst, dst.data(), sizeof(unsigned long long *) * (size_t)nsend, cudaMemcpyHostToDevice));
  }
  dm->nsend = nsend;
  dm->has_sendmap = true;
  return PSP_OK;
}

extern "C" void psp_dist_mat_destroy(psp_dmat *m) {
  DistMat *dm = (DistMat *)m;
  if (!dm) return;
  if (dm->local) psp_mat_destroy(dm->local);
  cudaFree(dm->d_owner); cudaFree(dm->d_ridx);
  cudaFree(dm->d_send_row); cudaFree(dm->d_send_dst);
  delete dm;
}

extern "C" int psp_dist_mat_info(const psp_dmat *m, int *nl, int *nhalo, unsigned *neighbour_mask) {
  const DistMat *dm = (const DistMat *)m;
  if (!dm) { set_error("psp_dist_mat_info: NULL"); return PSP_ERR_ARG; }
  if (nl) *nl = dm->nl;
  if (nhalo) *nhalo = dm->nhalo;
  if (neighbour_mask) *neighbour_mask = dm->nbmask;
  return PSP_OK;
}
