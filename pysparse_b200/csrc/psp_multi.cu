// psp_multi.cu -- several GPUs behind the unchanged single-process API.
//
// PySparse is a single Python process; its solvers are called as
//     info, iter, relres = krylov.pcg(A, b, x, tol, maxit, K)
// (pysparse/itsolvers/src/itsolversmodule.c:32-117).  With PSP_GPUS=N in the environment (or
// psp_set_option("gpus", N)) that very call runs on N GPUs of the node:
//   * psp_csr_create / psp_sss_create additionally cut the matrix into N contiguous row slabs
//     and upload slab r to device r (psp_dist.cu: local column numbering, halo tables, send
//     maps -- the same structures the one-process-per-GPU launcher builds);
//   * psp_solve / psp_solve_dev scatter b and x to the devices, run the sharded solver
//     (psp_solvers.cu: dist_solve_dev; all six solvers, with no or single-step Jacobi
//     preconditioning) on all of them at once and gather x.
// One host thread per GPU ("rank thread") owns that device's context, stream and slab and
// launches its persistent solver kernel; the kernels talk to each other directly over NVLink
// (peer access inside one process instead of CUDA IPC): halo values are pushed into the
// neighbours' landing zones and the scalar all-reduces go through the peers' mailboxes,
// exactly as in the multi-process path.  Everything the sharded path cannot do (other
// solvers, SSOR, foreign operators, matvec) keeps running on device 0 from the full copy.
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include "psp_mat.cuh"
#include "psp_dist.cuh"
#include "psp_multi.cuh"

namespace psp {

namespace {

struct Rank {
  int id = 0;
  Ctx ctx;
  DistHost dist;
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = false, quit = false;
  int status = PSP_OK;
  char err[512] = "";
};

struct Multi {
  int world = 0;
  std::vector<Rank *> ranks;
  size_t slot_elems = 0;
  int nslots = 0;
  unsigned long long heap_gen = 0;        // bumped whenever the symmetric heaps are re-created
  std::vector<MultiMat *> live;
} g_multi;

void rank_main(Rank *r) {
  cudaSetDevice(r->id);
  bind_thread(&r->ctx, &r->dist);
  for (;;) {
    std::unique_lock<std::mutex> lk(r->m);
    r->cv.wait(lk, [&] { return r->has_job || r->quit; });
    if (r->quit) return;
    std::function<int()> job = r->job;
    lk.unlock();
    int st = job();
    lk.lock();
    r->status = st;
    if (st != PSP_OK) { strncpy(r->err, psp_last_error(), sizeof r->err - 1); r->err[sizeof r->err - 1] = 0; }
    r->has_job = false;
    r->done = true;
    r->cv.notify_all();
  }
}

// run fn(rank) on every rank thread at the same time; first failure is returned with its message
int run_all(const std::function<int(int)> &fn) {
  Multi &M = g_multi;
  for (Rank *r : M.ranks) {
    std::lock_guard<std::mutex> lk(r->m);
    const int id = r->id;
    r->job = [fn, id]() { return fn(id); };
    r->has_job = true; r->done = false;
    r->cv.notify_all();
  }
  int st = PSP_OK;
  for (Rank *r : M.ranks) {
    std::unique_lock<std::mutex> lk(r->m);
    r->cv.wait(lk, [&] { return r->done; });
    if (r->status != PSP_OK && st == PSP_OK) { st = r->status; set_error("GPU %d: %s", r->id, r->err); }
  }
  return st;
}

void copy_options(Ctx &dst, const Ctx &src) {
  dst.spmv_variant = src.spmv_variant; dst.spmv_format = src.spmv_format; dst.spmv_ctas = src.spmv_ctas;
  dst.l2_stream = src.l2_stream; dst.pcg_persist = src.pcg_persist; dst.gmres_persist = src.gmres_persist;
  dst.bicg_persist = src.bicg_persist; dst.dist_push = src.dist_push; dst.dist_debug = src.dist_debug;
  dst.poll_chunk = src.poll_chunk;
}

int multi_start(int world) {
  Multi &M = g_multi;
  if (M.world == world) return PSP_OK;
  if (M.world != 0) { set_error("the number of GPUs cannot change once matrices have been sharded (%d -> %d)", M.world, world); return PSP_ERR_ARG; }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < world) {
    cudaGetLastError();
    set_error("gpus = %d requested but only %d CUDA device(s) are visible", world, count);
    return PSP_ERR_NO_DEVICE;
  }
  if (world > kMaxRanks) { set_error("gpus = %d exceeds the supported %d", world, kMaxRanks); return PSP_ERR_ARG; }
  int cur = 0;
  cudaGetDevice(&cur);
  for (int i = 0; i < world; i++) {
    cudaSetDevice(i);
    for (int j = 0; j < world; j++) {
      if (i == j) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, i, j);
      if (!can) { cudaSetDevice(cur); set_error("GPU %d cannot access GPU %d's memory (no NVLink / P2P)", i, j); return PSP_ERR_DIST; }
      cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        cudaSetDevice(cur);
        set_error("cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", i, j, cudaGetErrorString(e));
        return PSP_ERR_DIST;
      }
      cudaGetLastError();
    }
  }
  cudaSetDevice(cur);
  for (int i = 0; i < world; i++) {
    Rank *r = new Rank();
    r->id = i;
    M.ranks.push_back(r);
    r->th = std::thread(rank_main, r);
  }
  M.world = world;
  int st = run_all([](int rank) -> int { return init_ctx(g_multi.ranks[(size_t)rank]->ctx, rank); });
  return st;
}

// symmetric heaps of at least (slot_elems, nslots); re-created (and every live matrix re-bound)
// when they have to grow
int ensure_heap(size_t slot_elems, int nslots) {
  Multi &M = g_multi;
  if (M.heap_gen && M.slot_elems >= slot_elems && M.nslots >= nslots) return PSP_OK;
  const size_t se = std::max(M.slot_elems, slot_elems);
  const int ns = std::max(M.nslots, nslots);
  PSP_TRY(run_all([](int) -> int { return dist_local_free(); }));
  PSP_TRY(run_all([&](int rank) -> int { return dist_local_alloc(rank, g_multi.world, se, ns); }));
  std::vector<DistHost *> all;
  for (Rank *r : M.ranks) all.push_back(&r->dist);
  PSP_TRY(run_all([&](int) -> int { return dist_local_connect(all.data()); }));
  M.slot_elems = se; M.nslots = ns;
  M.heap_gen++;
  return PSP_OK;
}

int bind_sendmaps(MultiMat *mm) {
  if (mm->heap_gen == g_multi.heap_gen) return PSP_OK;
  PSP_TRY(run_all([&](int r) -> int {
    return psp_dist_set_sendmap((psp_dmat *)mm->slab[(size_t)r], (int)mm->send_rows[(size_t)r].size(), mm->send_rows[(size_t)r].data(),
                                mm->send_rank[(size_t)r].data(), mm->send_off[(size_t)r].data());
  }));
  mm->heap_gen = g_multi.heap_gen;
  return PSP_OK;
}

}  // namespace

int multi_gpus() {
  // the main context's option; rank threads never shard again
  return (int)ctx().gpus;
}

// Cut a square CSR matrix (host arrays, global column indices) into `world` row slabs of
// (nearly) equal size and upload slab r to device r.
int multi_attach_csr(psp_mat *A, const int *ind, const int *col, const double *val) {
  const int world = multi_gpus();
  if (world <= 1 || A->m != A->n || A->m < world) return PSP_OK;
  PSP_TRY(multi_start(world));
  Multi &M = g_multi;
  for (Rank *r : M.ranks) copy_options(r->ctx, ctx());
  MultiMat *mm = new MultiMat();
  mm->world = world;
  mm->rowptr.resize((size_t)world + 1);
  const int n = A->m;
  for (int r = 0; r <= world; r++) mm->rowptr[(size_t)r] = (int)((long long)n * r / world);
  mm->slab.assign((size_t)world, nullptr);
  std::vector<std::vector<int>> owner((size_t)world), ridx((size_t)world);
  int st = run_all([&](int r) -> int {
    const int r0 = mm->rowptr[(size_t)r], r1 = mm->rowptr[(size_t)r + 1];
    std::vector<int> lind((size_t)(r1 - r0) + 1);
    for (int i = r0; i <= r1; i++) lind[(size_t)(i - r0)] = ind[i] - ind[r0];
    return dist_csr_build(world, r, mm->rowptr.data(), ind[r1] - ind[r0], lind.data(), col + ind[r0], val + ind[r0],
                          &mm->slab[(size_t)r], &owner[(size_t)r], &ridx[(size_t)r]);
  });
  if (st != PSP_OK) { multi_detach(mm); return st; }
  // send maps: rank o pushes its local row ridx to position nl_w + h of rank w's vector for
  // every halo entry h of rank w that o owns
  mm->send_rows.assign((size_t)world, {}); mm->send_rank.assign((size_t)world, {}); mm->send_off.assign((size_t)world, {});
  for (int w = 0; w < world; w++) {
    const int nlw = mm->rowptr[(size_t)w + 1] - mm->rowptr[(size_t)w];
    for (size_t h = 0; h < owner[(size_t)w].size(); h++) {
      const int o = owner[(size_t)w][h];
      mm->send_rows[(size_t)o].push_back(ridx[(size_t)w][h]);
      mm->send_rank[(size_t)o].push_back(w);
      mm->send_off[(size_t)o].push_back(nlw + (int)h);
    }
    mm->max_slot = std::max(mm->max_slot, nlw + (int)owner[(size_t)w].size());
  }
  M.live.push_back(mm);
  A->multi = mm;
  return PSP_OK;
}

void multi_detach(MultiMat *mm) {
  if (!mm) return;
  Multi &M = g_multi;
  run_all([&](int r) -> int {
    if (mm->slab[(size_t)r]) psp_dist_mat_destroy((psp_dmat *)mm->slab[(size_t)r]);
    return PSP_OK;
  });
  M.live.erase(std::remove(M.live.begin(), M.live.end(), mm), M.live.end());
  delete mm;
}

void multi_prec_release(psp_prec *K) {
  if (K->multi_prec.empty() || g_multi.world == 0) return;
  run_all([&](int r) -> int {
    if ((size_t)r < K->multi_prec.size() && K->multi_prec[(size_t)r]) psp_prec_destroy(K->multi_prec[(size_t)r]);
    return PSP_OK;
  });
  K->multi_prec.clear();
}

bool multi_eligible(int solver, const psp_mat *A, const psp_prec *K, bool K_present) {
  if (!A || !A->multi) return false;
  if (solver < PSP_PCG || solver > PSP_GMRES) return false;
  if (!K_present) return true;
  return K && K->kind == 1 && K->steps == 1 && (int)K->h_dinv.size() == A->n;
}

int multi_solve(int solver, psp_mat *A, psp_prec *K, const double *b, double *x, bool on_device, double tol, int maxit,
                int gmres_dim, int *info, int *iter, double *relres) {
  Multi &M = g_multi;
  MultiMat *mm = A->multi;
  const int world = mm->world;
  Ctx &main = ctx();
  if (solver == PSP_GMRES && gmres_dim < 1) { set_error("gmres: dim must be >= 1"); return PSP_ERR_ARG; }
  const int slots = solver == PSP_PCG ? 4 : solver == PSP_BICGSTAB ? 9 : solver == PSP_MINRES ? 7 : solver == PSP_QMRS ? 6
                  : solver == PSP_CGS ? 9 : (K ? 2 * gmres_dim + 2 : gmres_dim + 2);
  PSP_TRY(ensure_heap((size_t)mm->max_slot + 64, slots));
  PSP_TRY(bind_sendmaps(mm));
  for (Rank *r : M.ranks) copy_options(r->ctx, main);
  // the Jacobi preconditioner's slice of every rank (dinv = omega / diag was formed at creation)
  if (K && (K->multi_prec.empty() || K->multi_rowptr != mm->rowptr)) {
    multi_prec_release(K);
    K->multi_prec.assign((size_t)world, nullptr);
    K->multi_rowptr = mm->rowptr;
    PSP_TRY(run_all([&](int r) -> int {
      const int r0 = mm->rowptr[(size_t)r], nl = mm->rowptr[(size_t)r + 1] - r0;
      return jacobi_from_dinv(nl, K->h_dinv.data() + r0, &K->multi_prec[(size_t)r]);
    }));
  }
  std::vector<int> rinfo((size_t)world, 0), riter((size_t)world, 0);
  std::vector<double> rrel((size_t)world, 0.0), rms((size_t)world, 0.0);
  std::vector<long long> rl((size_t)world, 0);
  int src_dev = 0;
  if (on_device) src_dev = main.device;
  PSP_TRY(run_all([&](int r) -> int {
    Ctx &c = ctx();
    const int r0 = mm->rowptr[(size_t)r], nl = mm->rowptr[(size_t)r + 1] - r0;
    double *db = nullptr, *dx = nullptr;
    PSP_TRY(pool_get(0, (size_t)nl, &db));
    PSP_TRY(pool_get(1, (size_t)nl, &dx));
    const size_t bytes = sizeof(double) * (size_t)nl;
    if (on_device) {
      PSP_CUDA(cudaMemcpyPeerAsync(db, c.device, b + r0, src_dev, bytes, c.stream));
      PSP_CUDA(cudaMemcpyPeerAsync(dx, c.device, x + r0, src_dev, bytes, c.stream));
    } else {
      PSP_CUDA(cudaMemcpyAsync(db, b + r0, bytes, cudaMemcpyHostToDevice, c.stream));
      PSP_CUDA(cudaMemcpyAsync(dx, x + r0, bytes, cudaMemcpyHostToDevice, c.stream));
    }
    const long long l0 = c.launches;
    PSP_TRY(dist_solve_dev(solver, mm->slab[(size_t)r], K ? K->multi_prec[(size_t)r] : nullptr, db, dx, tol, maxit, gmres_dim,
                           &rinfo[(size_t)r], &riter[(size_t)r], &rrel[(size_t)r]));
    rms[(size_t)r] = c.last_ms;
    rl[(size_t)r] = c.launches - l0;
    if (on_device) PSP_CUDA(cudaMemcpyPeerAsync(x + r0, src_dev, dx, c.device, bytes, c.stream));
    else PSP_CUDA(cudaMemcpyAsync(x + r0, dx, bytes, cudaMemcpyDeviceToHost, c.stream));
    PSP_CUDA(cudaStreamSynchronize(c.stream));
    return PSP_OK;
  }));
  for (int r = 1; r < world; r++)
    if (rinfo[(size_t)r] != rinfo[0] || riter[(size_t)r] != riter[0]) {
      set_error("multi-GPU solve: ranks disagree (rank 0: info %d iter %d, rank %d: info %d iter %d)", rinfo[0], riter[0], r,
                rinfo[(size_t)r], riter[(size_t)r]);
      return PSP_ERR_DIST;
    }
  if (info) *info = rinfo[0];
  if (iter) *iter = riter[0];
  if (relres) *relres = rrel[0];
  main.last_ms = *std::max_element(rms.begin(), rms.end());
  for (long long l : rl) main.launches += l;
  return PSP_OK;
}

void multi_shutdown() {
  Multi &M = g_multi;
  if (M.world == 0) return;
  for (Rank *r : M.ranks) {
    { std::lock_guard<std::mutex> lk(r->m); r->quit = true; r->cv.notify_all(); }
    if (r->th.joinable()) r->th.join();
  }
}

}  // namespace psp
