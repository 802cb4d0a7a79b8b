// psp_solvers.cu -- the six Krylov solvers, device-resident.
//
// Replaces Itsolvers_{pcg,minres,qmrs,bicgstab,cgs,gmres}_kernel
// (pysparse/itsolvers/src/*.c) and their Python wrappers' workspace handling
// (itsolversmodule.c:32-586).  Vectors stay in HBM for the whole solve; every
// iteration is a short, fixed sequence of fused kernels (psp_vec.cuh,
// psp_spmv.cu) whose scalar work runs in device-side epilogues
// (psp_epilogue.cuh).  The host enqueues iterations ahead of the GPU and only
// polls the `done` flag asynchronously, so there is no host synchronisation
// per step.  Foreign (Python) operators are supported by staging the operand
// vectors through pinned host memory around each callback.
#include "psp_mat.cuh"
#include "psp_vec.cuh"
#include "psp_correq.cuh"
#include "psp_dist.cuh"

namespace psp {

int pcg_persist_launch(const psp_mat *M, int n, const double *b, double *x, double *r, double *p, double *p2, double *q,
                       const double *dinv, DistMat *dm);
int gmres_persist_launch(const psp_mat *M, int n, const double *b, double *x, double *V, double *W, double *xs,
                         size_t N, int m, const double *dinv, const double *g, DistMat *dm);
int bicg_persist_launch(const psp_mat *M, int n, const double *b, double *x, double *r, double *rhat, double *p,
                        double *p2, double *phat, double *v, double *s, double *shat, double *t, const double *dinv,
                        DistMat *dm);                        // psp_persist.cu

// ------------------------------------------------------------------ generic functors
struct FNormSq {            // red0 = sum a_i^2
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *a;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *acc, double &) { double t = a[i]; acc[0] += t * t; }
};
struct FDot {               // red0 = sum a_i b_i
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *a, *b;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *acc, double &) { acc[0] += a[i] * b[i]; }
};
struct FDots2 {             // red0 = sum w1_i y_i ; red1 = sum w2_i y_i (w2 may be y)
  static constexpr int NS = 2; static constexpr bool HASMAX = false, TICKET = false;
  const double *w1, *w2, *y;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *acc, double &) {
    double t = y[i];
    if (w1) acc[0] += w1[i] * t;
    if (w2) acc[1] += w2[i] * t;
  }
};
struct FCopy {
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *a; double *b;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *, double &) { b[i] = a[i]; }
};

struct FScaleBy {           // y = x .* dinv
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *x, *dinv; double *y;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *, double &) { y[i] = x[i] * dinv[i]; }
};

// ------------------------------------------------------------------ operators
struct Op {
  psp_mat *mat = nullptr;
  psp_prec *prec = nullptr;
  psp_callback cb = nullptr;
  void *cbctx = nullptr;
  psp_correq *ceq = nullptr;   // device-resident correction-equation operator (psp_correq.cu), handed in as a mat / prec facade
  bool present = false;
  bool native() const { return mat || prec || ceq; }
};

static Op make_op(const psp_operator *o) {
  Op r;
  if (!o) return r;
  r.mat = o->mat; r.prec = o->prec; r.cb = o->cb; r.cbctx = o->ctx;
  if (r.mat && r.mat->ceq) { r.ceq = r.mat->ceq; r.mat = nullptr; }
  if (r.prec && r.prec->ceq) { r.ceq = r.prec->ceq; r.prec = nullptr; }
  r.present = (r.mat || r.prec || r.cb || r.ceq);
  return r;
}

struct Solve {
  int n = 0;
  Op A, K;
  State *S = nullptr;
  double *hx = nullptr, *hy = nullptr;   // pinned staging for callbacks
  bool cb_failed = false;
  // fused identity / Jacobi(steps == 1) preconditioner: z = r .* dinv (dinv == nullptr: z = r)
  bool fusedK = false;
  const double *dinv = nullptr;
  // multi-GPU: A is this rank's row slab; work vectors live in the symmetric heap
  DistMat *dm = nullptr;
  long vstride = 0;          // distance between work vectors (0: n)
};

// no operator of this solve leaves the device: the iteration can be enqueued ahead of the GPU
static bool device_resident(const Solve &s) {
  return (s.A.mat || s.dm || s.A.ceq) && (s.fusedK || s.K.prec || s.K.ceq);
}

static int read_done(Solve &s, int *done) {
  Ctx &c = ctx();
  PSP_CUDA(cudaMemcpyAsync(c.h_state, c.state, sizeof(State), cudaMemcpyDeviceToHost, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  *done = c.h_state->done;
  return PSP_OK;
}

// y := Op(x) through a host callback (vectors staged over PCIe)
static int apply_callback(Solve &s, const Op &op, const double *dx, double *dy, bool check_done = true) {
  Ctx &c = ctx();
  if (check_done) {
    int done = 0;
    PSP_TRY(read_done(s, &done));
    if (done) return PSP_OK;
  }
  size_t bytes = sizeof(double) * (size_t)s.n;
  PSP_CUDA(cudaMemcpyAsync(s.hx, dx, bytes, cudaMemcpyDeviceToHost, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  if (op.cb(op.cbctx, s.n, s.hx, s.hy) != 0) {
    s.cb_failed = true;
    set_error("operator callback failed");
    return PSP_ERR_CALLBACK;
  }
  PSP_CUDA(cudaMemcpyAsync(dy, s.hy, bytes, cudaMemcpyHostToDevice, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  return PSP_OK;
}

// y := A x, optionally fused with dots against y and a scalar epilogue
static int apply_A(Solve &s, const double *dx, double *dy, const double *w1 = nullptr,
                   const double *w2 = nullptr, bool w2_is_y = false, int epilogue = EPI_NONE,
                   int gate = 0) {
  if (s.dm) {
    // dx is a heap vector its producer has published: fetch the halo entries from the
    // neighbours' copies (NVLink), then multiply with the local slab
    PSP_TRY(dist_pull(s.dm, const_cast<double *>(dx), gate));
    SpmvDots d;
    d.w1 = w1; d.w2 = w2; d.w2_is_y = w2_is_y ? 1 : 0; d.epilogue = epilogue; d.S = s.S;
    d.gate = gate;
    return spmv_launch(s.dm->local, dx, dy, d);
  }
  if (s.A.mat) {
    SpmvDots d;
    d.w1 = w1; d.w2 = w2; d.w2_is_y = w2_is_y ? 1 : 0; d.epilogue = epilogue; d.S = s.S;
    d.gate = gate;
    return spmv_launch(s.A.mat, dx, dy, d);
  }
  if (s.A.ceq) PSP_TRY(correq_apply(s.A.ceq, 0, dx, dy, s.S, gate));
  else PSP_TRY(apply_callback(s, s.A, dx, dy));
  if (w1 || w2 || w2_is_y || epilogue != EPI_NONE) {
    FDots2 f{w1, w2_is_y ? dy : w2, dy};
    PSP_TRY(launch_map(f, s.n, epilogue, gate));
  }
  return PSP_OK;
}

// y := K^-1 x for a preconditioner that is present and not fused
static int apply_K(Solve &s, const double *dx, double *dy) {
  if (s.K.prec) return prec_apply_dev(s.K.prec, dx, dy, s.S);
  if (s.K.ceq) return correq_apply(s.K.ceq, 1, dx, dy, s.S, 0);
  return apply_callback(s, s.K, dx, dy);
}

// ------------------------------------------------------------------ polling driver
// Enqueue `body(it)` repeatedly, staying up to two polling chunks ahead of the
// GPU; stops once the device reports done.  `max_rounds` bounds the number of
// bodies ever enqueued (the device stops by itself through maxit).
template <class Body>
static int run_until_done(Solve &s, long long max_rounds, bool sync_each, Body body) {
  Ctx &c = ctx();
  long long chunk = c.poll_chunk > 0 ? c.poll_chunk : 8;
  int slot = 0;
  bool pending[2] = {false, false};
  long long rounds = 0;
  for (;;) {
    if (pending[slot]) {
      bool was_idle = (cudaEventQuery(c.evpoll[slot]) == cudaSuccess);
      PSP_CUDA(cudaEventSynchronize(c.evpoll[slot]));
      pending[slot] = false;
      if (c.h_state[slot].done) return PSP_OK;
      if (was_idle && c.poll_chunk <= 0 && chunk < 128) chunk *= 2;
    }
    if (rounds >= max_rounds) break;
    for (long long j = 0; j < chunk && rounds < max_rounds; j++, rounds++) {
      PSP_TRY(body(rounds));
      if (sync_each) {
        int done = 0;
        PSP_TRY(read_done(s, &done));
        if (done) return PSP_OK;
      }
    }
    PSP_CUDA(cudaMemcpyAsync(&c.h_state[slot], c.state, sizeof(State), cudaMemcpyDeviceToHost, c.stream));
    PSP_CUDA(cudaEventRecord(c.evpoll[slot], c.stream));
    pending[slot] = true;
    slot ^= 1;
  }
  for (int k = 0; k < 2; k++)
    if (pending[k]) PSP_CUDA(cudaEventSynchronize(c.evpoll[k]));
  return PSP_OK;
}

// ====================================================================== PCG
struct FPcgResid {          // r = b - r (r holds A x); sums r.r and r.(K^-1 r); zero rhs: x = 0
  static constexpr int NS = 2; static constexpr bool HASMAX = false, TICKET = false;
  const double *b; double *r, *x; const double *dinv; int zero_rhs;
  __device__ void prepare(const State *S) { zero_rhs = (S->phase == 1); }
  __device__ void elem(long i, double *acc, double &) {
    if (zero_rhs) { x[i] = 0.0; return; }
    double t = b[i] - r[i];
    r[i] = t;
    acc[0] += t * t;
    double z = dinv ? t * dinv[i] : t;
    acc[1] += t * z;
  }
};
struct FPcgP {              // p = z + beta p   (first iteration: p = z), pcg.c:105-115
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *__restrict__ z, *__restrict__ dinv; double *__restrict__ p; double beta; int first;
  __device__ void prepare(const State *S) { beta = S->v[V_BETA]; first = (S->it == 1); }
  __device__ void elem(long i, double *, double &) {
    double zi = dinv ? z[i] * dinv[i] : z[i];
    p[i] = first ? zi : zi + beta * p[i];
  }
};
struct FPcgUpdate {         // stagnation scan, x += alpha p, r -= alpha q, |r|^2 (and r.K^-1 r)
  static constexpr int NS = 2; static constexpr bool HASMAX = true, TICKET = false;
  const double *__restrict__ p, *__restrict__ q, *__restrict__ dinv; double *__restrict__ x, *__restrict__ r;
  double alpha; int scan;
  __device__ void prepare(const State *S) { alpha = S->v[V_ALPHA]; scan = (S->stag == 0); }
  __device__ void elem(long i, double *acc, double &mx) {
    double xi = x[i], pi = p[i];
    if (scan) {             // pcg.c:128-139
      if (xi != 0.0) mx = fmax(mx, fabs(alpha * pi / xi));
      else if (pi != 0.0) mx = fmax(mx, 1.0);
    }
    x[i] = xi + alpha * pi;
    double t = r[i] + (-alpha) * q[i];
    r[i] = t;
    acc[0] += t * t;
    double z = dinv ? t * dinv[i] : t;
    acc[1] += t * z;
  }
};

static int solve_pcg(Solve &s, const double *b, double *x, double *work) {
  const long n = s.n;
  const long vs = s.vstride ? s.vstride : n;
  double *r = work, *z = work + vs, *p = work + 2 * vs, *q = work + 3 * vs;
  const int pub = s.dm ? kPublish : 0;     // multi-GPU: p is read by the neighbours
  {
    // single-launch path: the whole solve in one persistent cooperative kernel
    // (psp_persist.cu).  pcg_persist: 0 never, anything else (default -1) whenever the
    // operator and the preconditioner are native (it wins at every size measured:
    // 256^3 563 vs 581 us/iteration, 128^3 77 vs 86, 100^2 16 vs 23).
    Ctx &c = ctx();
    const psp_mat *M = s.dm ? s.dm->local : s.A.mat;
    const bool eligible = M && s.fusedK;
    const bool choose = c.pcg_persist != 0;
    if (eligible && choose) return pcg_persist_launch(M, (int)n, b, x, r, p, z, q, s.dinv, s.dm);   // z: second p buffer
  }
  PSP_TRY(launch_map(FNormSq{b}, n, EPI_PCG_NORMB));
  if (s.dm) {
    PSP_TRY(launch_map(FCopy{x, p}, n, EPI_NONE, pub));   // x0 into the shared heap, published
    PSP_TRY(apply_A(s, p, r));
  } else {
    PSP_TRY(apply_A(s, x, r));
  }
  PSP_TRY(launch_map(FPcgResid{b, r, x, s.dinv, 0}, n, s.fusedK ? EPI_PCG_RESID0_J : EPI_PCG_RESID0));
  const bool sync_each = !device_resident(s);
  auto body = [&](long long) -> int {
    if (s.fusedK) {
      PSP_TRY(launch_map(FPcgP{r, s.dinv, p, 0.0, 0}, n, EPI_NONE, pub));
    } else {
      PSP_TRY(apply_K(s, r, z));
      PSP_TRY(launch_map(FDot{r, z}, n, EPI_PCG_RHO));
      PSP_TRY(launch_map(FPcgP{z, nullptr, p, 0.0, 0}, n, EPI_NONE, pub));
    }
    PSP_TRY(apply_A(s, p, q, p, nullptr, false, EPI_PCG_PQ));
    PSP_TRY(launch_map(FPcgUpdate{p, q, s.fusedK ? s.dinv : nullptr, x, r, 0.0, 0}, n,
                       s.fusedK ? EPI_PCG_UPDATE_J : EPI_PCG_UPDATE));
    return PSP_OK;
  };
  return run_until_done(s, (long long)ctx().h_state[0].maxit + 1, sync_each, body);
}

// ====================================================================== MINRES
struct FMinresInit {        // vh = b - vh (vh holds A x0); |vh|^2; clears vh_old, w, w_old
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *b; double *vh, *vh_old, *w, *w_old;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *acc, double &) {
    double t = b[i] - vh[i];
    vh[i] = t; vh_old[i] = 0.0; w[i] = 0.0; w_old[i] = 0.0;
    acc[0] += t * t;
  }
};
struct FDotK {              // red0 = sum a_i * (K^-1 a)_i with K identity / diagonal; writes y if given
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *a, *dinv; double *y;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *acc, double &) {
    double t = a[i];
    double z = dinv ? t * dinv[i] : t;
    if (y) y[i] = z;
    acc[0] += t * z;
  }
};
struct FMinresV {           // v = y / beta   (minres.c:124-125)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *y; double *v; double beta;
  __device__ void prepare(const State *S) { beta = S->v[M_BETA]; }
  __device__ void elem(long i, double *, double &) { v[i] = y[i] / beta; }
};
template <bool FUSED>
struct FMinresLanczos {     // vh_new = av - k1 vh - k2 vh_old; vh_old = vh  (minres.c:131-135)
  static constexpr int NS = FUSED ? 1 : 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *av, *dinv; double *vh, *vh_old, *y; double k1, k2;
  __device__ void prepare(const State *S) { k1 = S->v[M_K1]; k2 = S->v[M_K2]; }
  __device__ void elem(long i, double *acc, double &) {
    double old = vh[i];
    double t = av[i] - k1 * old - k2 * vh_old[i];
    vh_old[i] = old;
    vh[i] = t;
    if (FUSED) {
      double z = dinv ? t * dinv[i] : t;
      if (dinv) y[i] = z;
      acc[0] += t * z;
    }
  }
};
struct FMinresUpdate {      // w, w_old, x update (minres.c:176-184) + loop test epilogue
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = true;
  const double *v; double *w, *w_old, *x; double r1, r2, r3, step;
  __device__ void prepare(const State *S) {
    r1 = S->v[M_R1]; r2 = S->v[M_R2]; r3 = S->v[M_R3]; step = S->v[M_STEP];
  }
  __device__ void elem(long i, double *, double &) {
    double keep = w[i];
    double wn = (v[i] - r3 * w_old[i] - r2 * keep) / r1;
    w[i] = wn; w_old[i] = keep;
    x[i] += step * wn;
  }
};

static int solve_minres(Solve &s, const double *b, double *x, double *work) {
  const long n = s.n;
  const long vs = s.vstride ? s.vstride : n;
  double *vh_old = work, *vh = work + vs, *ybuf = work + 2 * vs, *w = work + 3 * vs,
         *w_old = work + 4 * vs, *v = work + 5 * vs, *av = work + 6 * vs;
  const int pub = s.dm ? kPublish : 0;     // multi-GPU: v is read by the neighbours
  // identity preconditioner: y is vh itself (never written separately)
  double *y = (s.fusedK && !s.dinv) ? vh : ybuf;
  if (s.dm) {
    PSP_TRY(launch_map(FCopy{x, v}, n, EPI_NONE, pub));   // x0 into the shared heap, published
    PSP_TRY(apply_A(s, v, vh));
  } else
  PSP_TRY(apply_A(s, x, vh));
  PSP_TRY(launch_map(FMinresInit{b, vh, vh_old, w, w_old}, n, EPI_MINRES_R0));
  if (s.fusedK) {
    PSP_TRY(launch_map(FDotK{vh, s.dinv, s.dinv ? y : nullptr}, n, EPI_MINRES_BETA0));
  } else {
    PSP_TRY(apply_K(s, vh, y));
    PSP_TRY(launch_map(FDot{vh, y}, n, EPI_MINRES_BETA0));
  }
  const bool sync_each = !device_resident(s);
  auto body = [&](long long) -> int {
    PSP_TRY(launch_map(FMinresV{y, v, 0.0}, n, EPI_NONE, pub));
    PSP_TRY(apply_A(s, v, av, v, nullptr, false, EPI_MINRES_ALPHA));
    if (s.fusedK) {
      PSP_TRY(launch_map(FMinresLanczos<true>{av, s.dinv, vh, vh_old, y, 0.0, 0.0}, n, EPI_MINRES_BETA));
    } else {
      PSP_TRY(launch_map(FMinresLanczos<false>{av, nullptr, vh, vh_old, y, 0.0, 0.0}, n));
      PSP_TRY(apply_K(s, vh, y));
      PSP_TRY(launch_map(FDot{vh, y}, n, EPI_MINRES_BETA));
    }
    PSP_TRY(launch_map(FMinresUpdate{v, w, w_old, x, 0, 0, 0, 0}, n, EPI_MINRES_LOOP));
    return PSP_OK;
  };
  return run_until_done(s, (long long)ctx().h_state[0].maxit + 1, sync_each, body);
}

// ====================================================================== QMRS
struct FQmrsInit0 {         // v = b ; |b|^2            (qmrs.c:68-69)
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *b; double *v;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *acc, double &) { double t = b[i]; v[i] = t; acc[0] += t * t; }
};
struct FQmrsInit1 {         // v /= rho0 ; p = g = d = x = 0   (qmrs.c:71-77) + first loop test
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = true;
  double *v, *p, *g, *d, *x; double rho0;
  __device__ void prepare(const State *S) { rho0 = S->v[Q_RHO0]; }
  __device__ void elem(long i, double *, double &) {
    v[i] = v[i] / rho0; p[i] = 0.0; g[i] = 0.0; d[i] = 0.0; x[i] = 0.0;
  }
};
struct FQmrsPG {            // p = v - p cc ; g = t0 - g cc   (qmrs.c:101-105)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *v, *t0, *dinv; double *p, *g; double cc;
  __device__ void prepare(const State *S) { cc = S->v[Q_CC]; }
  __device__ void elem(long i, double *, double &) {
    double vi = v[i];
    double ti = t0 ? t0[i] : (dinv ? vi * dinv[i] : vi);
    p[i] = vi - p[i] * cc;
    g[i] = ti - g[i] * cc;
  }
};
struct FQmrsV {             // v = t - v beta ; |v|^2   (qmrs.c:108-111)
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *t; double *v; double beta;
  __device__ void prepare(const State *S) { beta = S->v[Q_BETA]; }
  __device__ void elem(long i, double *acc, double &) {
    double nv = t[i] - v[i] * beta;
    v[i] = nv;
    acc[0] += nv * nv;
  }
};
struct FQmrsUpdate {        // d = p eta + d cc ; x += d ; v *= 1/rho1   (qmrs.c:131-135)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = true;
  const double *p; double *d, *x, *v; double eta, cc, inv;
  __device__ void prepare(const State *S) { eta = S->v[Q_ETA]; cc = S->v[Q_CC2]; inv = S->v[Q_INV]; }
  __device__ void elem(long i, double *, double &) {
    double di = p[i] * eta + d[i] * cc;
    d[i] = di;
    x[i] += di;
    v[i] *= inv;
  }
};

static int solve_qmrs(Solve &s, const double *b, double *x, double *work) {
  const long n = s.n;
  const long vs = s.vstride ? s.vstride : n;
  double *t0 = work, *p = work + vs, *d = work + 2 * vs, *v = work + 3 * vs, *t = work + 4 * vs,
         *g = work + 5 * vs;
  const int pub = s.dm ? kPublish : 0;     // multi-GPU: g is read by the neighbours
  PSP_TRY(launch_map(FQmrsInit0{b, v}, n, EPI_QMRS_RHO0));
  PSP_TRY(launch_map(FQmrsInit1{v, p, g, d, x, 0.0}, n, EPI_QMRS_LOOP));
  const bool sync_each = !device_resident(s);
  auto body = [&](long long) -> int {
    if (s.fusedK) {
      PSP_TRY(launch_map(FDotK{v, s.dinv, nullptr}, n, EPI_QMRS_DELTA));
      PSP_TRY(launch_map(FQmrsPG{v, nullptr, s.dinv, p, g, 0.0}, n, EPI_NONE, pub));
    } else {
      PSP_TRY(apply_K(s, v, t0));
      PSP_TRY(launch_map(FDot{t0, v}, n, EPI_QMRS_DELTA));
      PSP_TRY(launch_map(FQmrsPG{v, t0, nullptr, p, g, 0.0}, n, EPI_NONE, pub));
    }
    PSP_TRY(apply_A(s, g, t, g, nullptr, false, EPI_QMRS_EPS));
    PSP_TRY(launch_map(FQmrsV{t, v, 0.0}, n, EPI_QMRS_RHO1));
    PSP_TRY(launch_map(FQmrsUpdate{p, d, x, v, 0, 0, 0}, n, EPI_QMRS_LOOP));
    return PSP_OK;
  };
  PSP_TRY(run_until_done(s, (long long)ctx().h_state[0].maxit + 1, sync_each, body));
  // final x <- K^-1 x after a regular loop exit (qmrs.c:146-149)
  Ctx &c = ctx();
  State hs;
  PSP_CUDA(cudaMemcpyAsync(c.h_state, c.state, sizeof(State), cudaMemcpyDeviceToHost, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  hs = *c.h_state;
  if (hs.phase == 2 && s.K.present) {
    if (s.fusedK) {
      if (s.dinv) {
        PSP_TRY(launch_map(FScaleBy{x, s.dinv, t0}, n, EPI_NONE, /*gate=*/3));
        PSP_CUDA(cudaMemcpyAsync(x, t0, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
      }
    } else {
      if (s.K.prec) PSP_TRY(prec_apply_dev(s.K.prec, x, t0, nullptr));
      else if (s.K.ceq) PSP_TRY(correq_apply(s.K.ceq, 1, x, t0, nullptr, 3));
      else PSP_TRY(apply_callback(s, s.K, x, t0, /*check_done=*/false));
      PSP_CUDA(cudaMemcpyAsync(x, t0, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, c.stream));
    }
  }
  return PSP_OK;
}

// ====================================================================== BiCGSTAB
struct FBicgInit {          // r = b + -r ; rhat = r ; |r|^2 ; zero rhs: x = 0   (bicgstab.c:207-241)
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *b; double *r, *rhat, *x; int zero_rhs;
  __device__ void prepare(const State *S) { zero_rhs = (S->phase == 1); }
  __device__ void elem(long i, double *acc, double &) {
    if (zero_rhs) { x[i] = 0.0; return; }
    double t = b[i] + -r[i];
    r[i] = t; rhat[i] = t;
    acc[0] += t * t;
  }
};
struct FBicgP {             // p = r + beta (p - omega v)  (first: p = r); phat = K^-1 p when fused
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *r, *v, *dinv; double *p, *phat; double beta, omega; int first;
  __device__ void prepare(const State *S) { beta = S->v[B_BETA]; omega = S->v[B_OMEGA]; first = (S->it == 1); }
  __device__ void elem(long i, double *, double &) {
    double pn = first ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pn;
    if (dinv) phat[i] = pn * dinv[i];
  }
};
struct FBicgS {             // s = r + (-alpha) v ; shat = K^-1 s when fused   (bicgstab.c:265)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *r, *v, *dinv; double *s, *shat; double alpha;
  __device__ void prepare(const State *S) { alpha = S->v[B_ALPHA]; }
  __device__ void elem(long i, double *, double &) {
    double sn = r[i] + (-alpha) * v[i];
    s[i] = sn;
    if (dinv) shat[i] = sn * dinv[i];
  }
};
struct FBicgUpdate {        // x += alpha phat + omega shat ; r = s - omega t ; |r|^2 ; rhat.r
  static constexpr int NS = 2; static constexpr bool HASMAX = false, TICKET = false;
  const double *phat, *shat, *s, *t, *rhat; double *x, *r; double alpha, omega;
  __device__ void prepare(const State *S) { alpha = S->v[B_ALPHA]; omega = S->v[B_OMEGA]; }
  __device__ void elem(long i, double *acc, double &) {
    x[i] = x[i] + alpha * phat[i] + omega * shat[i];
    double rn = s[i] - omega * t[i];
    r[i] = rn;
    acc[0] += rn * rn;
    acc[1] += rhat[i] * rn;
  }
};

static int solve_bicgstab(Solve &s, const double *b, double *x, double *work) {
  const long n = s.n;
  const long vs = s.vstride ? s.vstride : n;
  double *r = work, *rhat = work + vs, *p = work + 2 * vs, *phatb = work + 3 * vs, *v = work + 4 * vs,
         *sv = work + 5 * vs, *shatb = work + 6 * vs, *t = work + 7 * vs;
  const bool ident = s.fusedK && !s.dinv;
  double *phat = ident ? p : phatb, *shat = ident ? sv : shatb;
  const int pub = s.dm ? kPublish : 0;     // multi-GPU: phat / shat are read by the neighbours
  {
    // single-launch path (psp_persist_bicg.cu): the default whenever operator and preconditioner are native
    Ctx &c = ctx();
    const psp_mat *M = s.dm ? s.dm->local : s.A.mat;
    if (M && s.fusedK && c.bicg_persist != 0) {
      double *p2 = (s.dm && dist().nslots >= 9) ? work + 8 * vs : nullptr;   // second p buffer (push model)
      return bicg_persist_launch(M, (int)n, b, x, r, rhat, p, p2, phat, v, sv, shat, t, s.dinv, s.dm);
    }
  }
  PSP_TRY(launch_map(FNormSq{b}, n, EPI_BICG_NORMB));
  if (s.dm) {
    PSP_TRY(launch_map(FCopy{x, t}, n, EPI_NONE, pub));
    PSP_TRY(apply_A(s, t, r));
  } else {
    PSP_TRY(apply_A(s, x, r));
  }
  PSP_TRY(launch_map(FBicgInit{b, r, rhat, x, 0}, n, EPI_BICG_RES0));
  const bool sync_each = !device_resident(s);
  auto body = [&](long long) -> int {
    PSP_TRY(launch_map(FBicgP{r, v, s.fusedK ? s.dinv : nullptr, p, phat, 0, 0, 0}, n, EPI_NONE, pub));
    if (!s.fusedK) PSP_TRY(apply_K(s, p, phat));
    PSP_TRY(apply_A(s, phat, v, rhat, nullptr, false, EPI_BICG_ALPHA));
    PSP_TRY(launch_map(FBicgS{r, v, s.fusedK ? s.dinv : nullptr, sv, shat, 0}, n, EPI_NONE, pub));
    if (!s.fusedK) PSP_TRY(apply_K(s, sv, shat));
    PSP_TRY(apply_A(s, shat, t, sv, nullptr, true, EPI_BICG_OMEGA));
    PSP_TRY(launch_map(FBicgUpdate{phat, shat, sv, t, rhat, x, r, 0, 0}, n, EPI_BICG_RES));
    return PSP_OK;
  };
  return run_until_done(s, (long long)ctx().h_state[0].maxit + 1, sync_each, body);
}

// ====================================================================== CGS
struct FCgsInit {           // r0 = b - A x0 ; r = u = p = r0 ; r0.r0 ; b.b   (cgs.c:43-56)
  static constexpr int NS = 2; static constexpr bool HASMAX = false, TICKET = false;
  const double *b, *ax; double *r0, *r, *u, *p;
  __device__ void prepare(const State *) {}
  __device__ void elem(long i, double *acc, double &) {
    double bi = b[i];
    double t = bi + (-1.0) * ax[i];
    r0[i] = t; r[i] = t; u[i] = t; p[i] = t;
    acc[0] += t * t;
    acc[1] += bi * bi;
  }
};
struct FCgsQ {              // q = u - alpha v ; ta = u + q ; tb = K^-1 ta when fused   (cgs.c:74-79)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  const double *u, *v, *dinv; double *q, *ta, *tb; double alpha;
  __device__ void prepare(const State *S) { alpha = S->v[C_ALPHA]; }
  __device__ void elem(long i, double *, double &) {
    double ui = u[i];
    double qi = ui + (-alpha) * v[i];
    q[i] = qi;
    double ti = ui + 1.0 * qi;
    ta[i] = ti;
    if (dinv) tb[i] = ti * dinv[i];
  }
};
struct FCgsUpdate {         // x += alpha tb ; r -= alpha tc ; r.r ; r.r0   (cgs.c:84-95)
  static constexpr int NS = 2; static constexpr bool HASMAX = false, TICKET = false;
  const double *tb, *tc, *r0; double *x, *r; double alpha;
  __device__ void prepare(const State *S) { alpha = S->v[C_ALPHA]; }
  __device__ void elem(long i, double *acc, double &) {
    x[i] = x[i] + alpha * tb[i];
    double rn = r[i] + (-alpha) * tc[i];
    r[i] = rn;
    acc[0] += rn * rn;
    acc[1] += rn * r0[i];
  }
};
struct FCgsUP {             // u = r + beta q ; p = u + beta (q + beta p)   (cgs.c:100-105)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = true;
  const double *r, *q; double *u, *p; double beta;
  __device__ void prepare(const State *S) { beta = S->v[C_BETA]; }
  __device__ void elem(long i, double *, double &) {
    double qi = q[i];
    double un = r[i] + beta * qi;
    u[i] = un;
    double tt = qi + beta * p[i];
    p[i] = un + beta * tt;
  }
};

static int solve_cgs(Solve &s, const double *b, double *x, double *work) {
  const long n = s.n;
  const long vs = s.vstride ? s.vstride : n;
  double *r0 = work, *r = work + vs, *p = work + 2 * vs, *q = work + 3 * vs, *u = work + 4 * vs,
         *v = work + 5 * vs, *ta = work + 6 * vs, *tbb = work + 7 * vs, *tc = work + 8 * vs;
  const bool ident = s.fusedK && !s.dinv;
  // multi-GPU: the vectors A is applied to (K^-1 p and K^-1 (u + q)) are read by the neighbours
  const int pub = s.dm ? kPublish : 0;
  if (s.dm) {
    PSP_TRY(launch_map(FCopy{x, tc}, n, EPI_NONE, pub));   // x0 into the shared heap, published
    PSP_TRY(apply_A(s, tc, ta));
  } else
  PSP_TRY(apply_A(s, x, ta));
  PSP_TRY(launch_map(FCgsInit{b, ta, r0, r, u, p}, n, EPI_CGS_INIT, ident ? pub : 0));   // identity K: p itself is multiplied next
  const bool sync_each = !device_resident(s);
  auto body = [&](long long) -> int {
    const double *kp = p;                       // K^-1 p
    if (s.fusedK) {
      if (s.dinv) { PSP_TRY(launch_map(FScaleBy{p, s.dinv, ta}, n, EPI_NONE, pub)); kp = ta; }
    } else {
      PSP_TRY(apply_K(s, p, ta)); kp = ta;
    }
    PSP_TRY(apply_A(s, kp, v, r0, nullptr, false, EPI_CGS_ALPHA));
    double *tb = ident ? ta : tbb;
    PSP_TRY(launch_map(FCgsQ{u, v, s.fusedK ? s.dinv : nullptr, q, ta, tb, 0}, n, EPI_NONE, pub));
    if (!s.fusedK) PSP_TRY(apply_K(s, ta, tb));
    PSP_TRY(apply_A(s, tb, tc));
    PSP_TRY(launch_map(FCgsUpdate{tb, tc, r0, x, r, 0}, n, EPI_CGS_RES));
    PSP_TRY(launch_map(FCgsUP{r, q, u, p, 0}, n, EPI_CGS_LOOP, ident ? pub : 0));
    return PSP_OK;
  };
  return run_until_done(s, (long long)ctx().h_state[0].maxit + 1, sync_each, body);
}

// ====================================================================== GMRES
struct FGmresResid {        // V0 = V0 - b (V0 holds A x) ; |V0|^2 ; zero rhs: x = 0   (gmres.c:102-116)
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *b; double *v0, *x; int zero_rhs;
  __device__ void prepare(const State *S) { zero_rhs = (S->phase == 1); }
  __device__ void elem(long i, double *acc, double &) {
    if (zero_rhs) { x[i] = 0.0; return; }
    double t = v0[i] + (-1.0) * b[i];
    v0[i] = t;
    acc[0] += t * t;
  }
};
struct FGmresScale {        // a *= scalar slot (dscal)   (gmres.c:117,:150)
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = false;
  double *a; int slot; double f;
  __device__ void prepare(const State *S) { f = S->v[slot]; }
  __device__ void elem(long i, double *, double &) { a[i] = f * a[i]; }
};
struct FGmresMgs {          // vn -= H(k,i) vk ; red0 = <vn, vnext>   (gmres.c:141-148)
  static constexpr int NS = 1; static constexpr bool HASMAX = false, TICKET = false;
  const double *vk, *vnext; double *vn; double h;
  __device__ void prepare(const State *S) { h = S->v[G_HK]; }
  __device__ void elem(long i, double *acc, double &) {
    double t = vn[i] + (-h) * vk[i];
    vn[i] = t;
    acc[0] += t * (vnext ? vnext[i] : t);
  }
};
struct FGmresX {            // x += sum_{j<=last} g[j] W(j)   (gmres.c:177-181) + outer loop test
  static constexpr int NS = 0; static constexpr bool HASMAX = false, TICKET = true;
  const double *W; double *x; const double *g; long n; int last;
  __device__ void prepare(const State *S) { last = S->last_i; }
  __device__ void elem(long i, double *, double &) {
    double xi = x[i];
    for (int j = 0; j <= last; j++) xi = xi + g[j] * W[(size_t)j * n + i];
    x[i] = xi;
  }
};

static int solve_gmres(Solve &s, const double *b, double *x, int m) {
  Ctx &c = ctx();
  const long n = s.n;
  const size_t N = (size_t)(s.vstride ? s.vstride : n);     // distance between basis vectors
  double *V = nullptr, *W = nullptr, *small = nullptr, *xs = nullptr;
  const bool ident = s.fusedK && !s.dinv;
  const int pub = s.dm ? kPublish : 0;
  auto cleanup = [&]() { if (!s.dm) { cudaFree(V); if (!ident) cudaFree(W); } cudaFree(small); };
  if (s.dm) {
    // multi-GPU: the basis lives in the symmetric heap (every W(i) is multiplied by A)
    DistHost &d = dist();
    const int need = (ident ? m + 1 : 2 * m + 1) + 1;
    if (d.nslots < need) { set_error("gmres: symmetric heap needs %d slots, has %d", need, d.nslots); return PSP_ERR_ARG; }
    V = d.heap;
    W = ident ? V : V + (size_t)(m + 1) * N;
    xs = d.heap + (size_t)(need - 1) * N;
  }
  if ((!s.dm && (cudaMalloc((void **)&V, sizeof(double) * N * (size_t)(m + 1)) != cudaSuccess ||
                 (!ident && cudaMalloc((void **)&W, sizeof(double) * N * (size_t)m) != cudaSuccess))) ||
      cudaMalloc((void **)&small, sizeof(double) * ((size_t)(m + 1) * m + (m + 1) + 2 * (size_t)m)) != cudaSuccess) {
    cudaGetLastError();
    cleanup();
    set_error("gmres: out of device memory for the Krylov basis (n=%ld, m=%d)", n, m);
    return PSP_ERR_NOMEM;
  }
  if (ident) W = V;
  GmresSmall G;
  G.H = small; G.g = small + (size_t)(m + 1) * m; G.cs = G.g + (m + 1); G.sn = G.cs + m;
  cudaMemcpyAsync((void *)(c.state + 1), &G, sizeof G, cudaMemcpyHostToDevice, c.stream);
  int st = PSP_OK;
  auto run = [&]() -> int {
    PSP_TRY(launch_map(FNormSq{b}, n, EPI_GMRES_NORMB));
    auto cycle = [&](long long) -> int {
      if (s.dm) {
        PSP_TRY(launch_map(FCopy{x, xs}, n, EPI_NONE, pub));
        PSP_TRY(apply_A(s, xs, V));
      } else {
        PSP_TRY(apply_A(s, x, V));
      }
      PSP_TRY(launch_map(FGmresResid{b, V, x, 0}, n, EPI_GMRES_BETA));
      PSP_TRY(launch_map(FGmresScale{V, G_SCALE0, 0}, n, EPI_NONE, ident ? pub : 0));
      for (int i = 0; i < m; i++) {
        double *vi = V + (size_t)i * N, *vn = V + (size_t)(i + 1) * N, *wi = W + (size_t)i * N;
        if (s.fusedK) {
          if (s.dinv) PSP_TRY(launch_map(FScaleBy{vi, s.dinv, wi}, n, EPI_NONE, 1 | pub));
        } else {
          if (s.K.prec) PSP_TRY(prec_apply_dev(s.K.prec, vi, wi, s.S));
          else if (s.K.ceq) PSP_TRY(correq_apply(s.K.ceq, 1, vi, wi, s.S, 1));
          else {
            // callback: must not run once the cycle is over
            PSP_CUDA(cudaMemcpyAsync(c.h_state, c.state, sizeof(State), cudaMemcpyDeviceToHost, c.stream));
            PSP_CUDA(cudaStreamSynchronize(c.stream));
            if (c.h_state->done || c.h_state->cyc_done) break;
            PSP_TRY(apply_callback(s, s.K, vi, wi, false));
          }
        }
        if (!s.A.mat && !s.dm && !s.A.ceq) {
          PSP_CUDA(cudaMemcpyAsync(c.h_state, c.state, sizeof(State), cudaMemcpyDeviceToHost, c.stream));
          PSP_CUDA(cudaStreamSynchronize(c.stream));
          if (c.h_state->done || c.h_state->cyc_done) break;
        }
        PSP_TRY(apply_A(s, wi, vn, V, nullptr, false, EPI_GMRES_H, 1));
        for (int k = 0; k <= i; k++) {
          const double *vk = V + (size_t)k * N;
          const double *vnext = (k < i) ? V + (size_t)(k + 1) * N : nullptr;
          PSP_TRY(launch_map(FGmresMgs{vk, vnext, vn, 0}, n, k < i ? EPI_GMRES_H : EPI_GMRES_HLAST, 1));
        }
        // V(i+1) is the next SpMV input when the preconditioner is the identity
        PSP_TRY(launch_map(FGmresScale{vn, G_INVH, 0}, n, EPI_NONE, 1 | (ident ? pub : 0)));
      }
      PSP_TRY(launch_map(FGmresX{W, x, G.g, (long)N, 0}, n, EPI_GMRES_CYCLE));
      // poll once per restart cycle
      int done = 0;
      PSP_TRY(read_done(s, &done));
      return PSP_OK;
    };
    // every cycle performs >= 1 iteration; +2 covers the final true-residual pass
    long long max_cycles = (long long)c.h_state[0].maxit + 2;
    for (long long cyc = 0; cyc < max_cycles; cyc++) {
      PSP_TRY(cycle(cyc));
      if (c.h_state->done) break;
    }
    return PSP_OK;
  };
  {
    // single-launch path (psp_persist_gmres.cu): pays off when the launches and all-reduces of a
    // restart cycle cross several GPUs (automatic from 4 ranks on)
    const psp_mat *M = s.dm ? s.dm->local : s.A.mat;
    const bool want = c.gmres_persist == 1 || (c.gmres_persist < 0 && s.dm && dist().world >= 4);
    if (M && s.fusedK && want && m <= 256) {
      st = gmres_persist_launch(M, (int)n, b, x, V, W, s.dm ? xs : x, N, m, s.dinv, G.g, s.dm);
      cudaStreamSynchronize(c.stream);
      cleanup();
      return st;
    }
  }
  st = run();
  cudaStreamSynchronize(c.stream);
  cleanup();
  return st;
}

// ====================================================================== entry
static size_t work_vectors(int solver) {
  switch (solver) {
  case PSP_PCG: return 4;        // itsolversmodule.c:90
  case PSP_MINRES: return 7;     // :278
  case PSP_QMRS: return 6;       // :467
  case PSP_BICGSTAB: return 8;   // :182
  case PSP_CGS: return 9;        // :560 (8) + one extra so the SpMV is never in place
  default: return 0;             // gmres allocates V, W itself (gmres.c:84-96)
  }
}

// Per-solve resources of the context (device State + the small block behind it, its pinned
// mirror, the timing / polling events).  The outermost solve uses the ones created by
// psp_init; a solve started from inside a Python operator callback of a running solve (inner-
// outer iterations: a preconditioner that itself calls krylov.*) gets fresh ones, swapped into
// the context for its duration, so that it cannot overwrite the outer solve's state -- the
// reference's solvers are re-entrant C functions and support this.
struct SolveFrame {
  State *state = nullptr, *h_state = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evpoll[2] = {nullptr, nullptr};
  void take(Ctx &c) { state = c.state; h_state = c.h_state; ev0 = c.ev0; ev1 = c.ev1; evpoll[0] = c.evpoll[0]; evpoll[1] = c.evpoll[1]; }
  void put(Ctx &c) const { c.state = state; c.h_state = h_state; c.ev0 = ev0; c.ev1 = ev1; c.evpoll[0] = evpoll[0]; c.evpoll[1] = evpoll[1]; }
  int create() {
    PSP_CUDA(cudaMalloc((void **)&state, sizeof(State) + 256));
    PSP_CUDA(cudaMemset(state, 0, sizeof(State) + 256));
    PSP_CUDA(cudaHostAlloc((void **)&h_state, 2 * sizeof(State), cudaHostAllocDefault));
    PSP_CUDA(cudaEventCreate(&ev0));
    PSP_CUDA(cudaEventCreate(&ev1));
    PSP_CUDA(cudaEventCreateWithFlags(&evpoll[0], cudaEventDisableTiming));
    PSP_CUDA(cudaEventCreateWithFlags(&evpoll[1], cudaEventDisableTiming));
    return PSP_OK;
  }
  void destroy() {
    cudaFree(state); cudaFreeHost(h_state);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (evpoll[0]) cudaEventDestroy(evpoll[0]);
    if (evpoll[1]) cudaEventDestroy(evpoll[1]);
  }
};

static int solve_dev_frame(int solver, const psp_operator *Aop, const psp_operator *Kop, int n, const double *db,
                           double *dx, double tol, int maxit, int gmres_dim, int *info, int *iter, double *relres);

// gpus > 1 and the operands allow it: the sharded solve on all GPUs (psp_multi.cu)
static bool wants_multi(int solver, const psp_operator *Aop, const psp_operator *Kop, int n) {
  if (!Aop || !Aop->mat || ctx().solve_depth != 0 || Aop->mat->n != n) return false;
  const bool kp = Kop && (Kop->mat || Kop->prec || Kop->cb);
  return multi_eligible(solver, Aop->mat, Kop ? Kop->prec : nullptr, kp);
}

int solve_dev(int solver, const psp_operator *Aop, const psp_operator *Kop, int n, const double *db,
              double *dx, double tol, int maxit, int gmres_dim, int *info, int *iter, double *relres) {
  PSP_TRY(ensure_init());
  Ctx &c = ctx();
  if (wants_multi(solver, Aop, Kop, n)) {
    PSP_CUDA(cudaStreamSynchronize(c.stream));      // b and x are complete before the other devices read them
    return multi_solve(solver, Aop->mat, Kop ? Kop->prec : nullptr, db, dx, true, tol, maxit, gmres_dim, info, iter, relres);
  }
  if (c.solve_depth == 0) {
    c.solve_depth = 1;
    const int st = solve_dev_frame(solver, Aop, Kop, n, db, dx, tol, maxit, gmres_dim, info, iter, relres);
    c.solve_depth = 0;
    return st;
  }
  // nested solve: private frame
  SolveFrame outer, inner;
  outer.take(c);
  int st = inner.create();
  if (st == PSP_OK) {
    const double outer_ms = c.last_ms;
    inner.put(c);
    c.solve_depth++;
    st = solve_dev_frame(solver, Aop, Kop, n, db, dx, tol, maxit, gmres_dim, info, iter, relres);
    c.solve_depth--;
    cudaStreamSynchronize(c.stream);
    outer.put(c);
    c.last_ms = outer_ms;
  }
  inner.destroy();
  return st;
}

static int solve_dev_frame(int solver, const psp_operator *Aop, const psp_operator *Kop, int n, const double *db,
                           double *dx, double tol, int maxit, int gmres_dim, int *info, int *iter, double *relres) {
  Ctx &c = ctx();
  if (solver < 0 || solver > PSP_GMRES) { set_error("unknown solver id %d", solver); return PSP_ERR_ARG; }
  if (n < 0 || !Aop) { set_error("psp_solve: bad arguments"); return PSP_ERR_ARG; }
  Solve s;
  s.n = n;
  s.A = make_op(Aop);
  s.K = make_op(Kop);
  s.S = c.state;
  if (!s.A.present) { set_error("psp_solve: operator A is empty"); return PSP_ERR_ARG; }
  if (s.A.mat && (s.A.mat->m != n || s.A.mat->n != n)) { set_error("psp_solve: matrix is not n x n"); return PSP_ERR_ARG; }
  if (s.K.prec && s.K.prec->n != n) { set_error("psp_solve: preconditioner order mismatch"); return PSP_ERR_ARG; }
  if (solver == PSP_GMRES && gmres_dim < 1) { set_error("gmres: dim must be >= 1"); return PSP_ERR_ARG; }
  // identity and single-step Jacobi preconditioners are fused into the vector kernels
  if (!s.K.present) { s.fusedK = true; s.dinv = nullptr; }
  else if (s.K.prec && s.K.prec->kind == 1 && s.K.prec->steps == 1) { s.fusedK = true; s.dinv = s.K.prec->d_dinv; }

  double *work = nullptr;
  size_t nw = work_vectors(solver);
  const bool pooled = (c.solve_depth == 1);       // nested solves must not share the outer solve's scratch
  if (nw) {
    const size_t elems = nw * (size_t)(n > 0 ? n : 1);
    if (pooled) PSP_TRY(pool_get(2, elems, &work));
    else if (cudaMalloc((void **)&work, sizeof(double) * elems) != cudaSuccess) {
      cudaGetLastError();
      set_error("psp_solve: out of device memory for %zu work vectors of length %d", nw, n);
      return PSP_ERR_NOMEM;
    }
  }
  if (s.A.cb || s.K.cb) {
    if (cudaHostAlloc((void **)&s.hx, sizeof(double) * (size_t)(n > 0 ? n : 1), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&s.hy, sizeof(double) * (size_t)(n > 0 ? n : 1), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      if (!pooled) cudaFree(work);
      cudaFreeHost(s.hx); cudaFreeHost(s.hy);
      set_error("psp_solve: out of pinned host memory");
      return PSP_ERR_NOMEM;
    }
  }
  State init;
  memset(&init, 0, sizeof init);
  init.maxit = maxit; init.tol = tol; init.mdim = gmres_dim;
  init.iter = -1; init.relres = nan("");
  c.h_state[0] = init;
  c.h_state[1] = init;
  int st = PSP_OK;
  auto go = [&]() -> int {
    PSP_CUDA(cudaMemcpyAsync(c.state, &c.h_state[0], sizeof(State), cudaMemcpyHostToDevice, c.stream));
    PSP_CUDA(cudaStreamSynchronize(c.stream));
    c.h_state[0].done = 0; c.h_state[1].done = 0;
    PSP_CUDA(cudaEventRecord(c.ev0, c.stream));
    switch (solver) {
    case PSP_PCG: PSP_TRY(solve_pcg(s, db, dx, work)); break;
    case PSP_MINRES: PSP_TRY(solve_minres(s, db, dx, work)); break;
    case PSP_QMRS: PSP_TRY(solve_qmrs(s, db, dx, work)); break;
    case PSP_BICGSTAB: PSP_TRY(solve_bicgstab(s, db, dx, work)); break;
    case PSP_CGS: PSP_TRY(solve_cgs(s, db, dx, work)); break;
    case PSP_GMRES: PSP_TRY(solve_gmres(s, db, dx, gmres_dim)); break;
    }
    PSP_CUDA(cudaEventRecord(c.ev1, c.stream));
    PSP_CUDA(cudaMemcpyAsync(&c.h_state[0], c.state, sizeof(State), cudaMemcpyDeviceToHost, c.stream));
    PSP_CUDA(cudaStreamSynchronize(c.stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c.ev0, c.ev1);
    c.last_ms = ms;
    return PSP_OK;
  };
  st = go();
  if (st != PSP_OK) cudaStreamSynchronize(c.stream);
  if (!pooled) cudaFree(work);
  if (s.hx) cudaFreeHost(s.hx);
  if (s.hy) cudaFreeHost(s.hy);
  if (st != PSP_OK) return st;
  const State &R = c.h_state[0];
  if (info) *info = R.info;
  if (iter) *iter = R.iter;
  if (relres) *relres = R.relres;
  return PSP_OK;
}

// Distributed solve: this rank's row slab, vectors in the symmetric heap.  The
// kernels are the single-GPU ones; reductions are all-reduced in-kernel and
// the multiplied vector's halo is pulled from the neighbours (psp_dist.cu).
int dist_solve_dev(int solver, DistMat *dm, psp_prec *K, const double *db, double *dx, double tol,
                   int maxit, int gmres_dim, int *info, int *iter, double *relres) {
  PSP_TRY(ensure_init());
  Ctx &c = ctx();
  DistHost &d = dist();
  if (!d.ready || !dm) { set_error("psp_dist_solve_dev: distributed layer not initialised"); return PSP_ERR_DIST; }
  if (solver < PSP_PCG || solver > PSP_GMRES) { set_error("psp_dist_solve_dev: unknown solver id %d", solver); return PSP_ERR_ARG; }
  if (solver == PSP_GMRES && gmres_dim < 1) { set_error("gmres: dim must be >= 1"); return PSP_ERR_ARG; }
  if (K && !(K->kind == 1 && K->steps == 1 && K->n == dm->nl)) {
    set_error("psp_dist_solve_dev: preconditioner must be None or single-step Jacobi on the local rows");
    return PSP_ERR_ARG;
  }
  const int need_slots = solver == PSP_PCG ? 4 : solver == PSP_BICGSTAB ? 8 : solver == PSP_MINRES ? 7 : solver == PSP_QMRS ? 6
                       : solver == PSP_CGS ? 9 : 0;
  if (d.nslots < need_slots) {
    set_error("psp_dist_solve_dev: symmetric heap has %d slots, the solver needs %d", d.nslots, need_slots);
    return PSP_ERR_ARG;
  }
  Solve s;
  s.n = dm->nl;
  s.S = c.state;
  s.dm = dm;
  s.vstride = (long)d.slot_elems;
  s.fusedK = true;
  s.dinv = K ? K->d_dinv : nullptr;
  s.K.present = (K != nullptr);
  s.K.prec = K;
  State init;
  memset(&init, 0, sizeof init);
  init.dist = d.dev;
  init.maxit = maxit; init.tol = tol; init.iter = -1; init.relres = nan(""); init.mdim = gmres_dim;
  c.h_state[0] = init;
  c.h_state[1] = init;
  PSP_CUDA(cudaMemcpyAsync(c.state, &c.h_state[0], sizeof(State), cudaMemcpyHostToDevice, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  PSP_CUDA(cudaEventRecord(c.ev0, c.stream));
  int st = solver == PSP_PCG ? solve_pcg(s, db, dx, d.heap)
         : solver == PSP_BICGSTAB ? solve_bicgstab(s, db, dx, d.heap)
         : solver == PSP_MINRES ? solve_minres(s, db, dx, d.heap)
         : solver == PSP_QMRS ? solve_qmrs(s, db, dx, d.heap)
         : solver == PSP_CGS ? solve_cgs(s, db, dx, d.heap) : solve_gmres(s, db, dx, gmres_dim);
  if (st != PSP_OK) { cudaStreamSynchronize(c.stream); return st; }
  PSP_CUDA(cudaEventRecord(c.ev1, c.stream));
  PSP_CUDA(cudaMemcpyAsync(&c.h_state[0], c.state, sizeof(State), cudaMemcpyDeviceToHost, c.stream));
  PSP_CUDA(cudaStreamSynchronize(c.stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c.ev0, c.ev1);
  c.last_ms = ms;
  const State &R = c.h_state[0];
  if (info) *info = R.info;
  if (iter) *iter = R.iter;
  if (relres) *relres = R.relres;
  return PSP_OK;
}

}  // namespace psp

using namespace psp;

extern "C" int psp_dist_solve_dev(int solver, psp_dmat *A, psp_prec *K, const double *db, double *dx,
                                  double tol, int maxit, int gmres_dim, int *info, int *iter, double *relres) {
  return dist_solve_dev(solver, (DistMat *)A, K, db, dx, tol, maxit, gmres_dim, info, iter, relres);
}

extern "C" int psp_solve_dev(int solver, const psp_operator *A, const psp_operator *K, int n,
                             const double *db, double *dx, double tol, int maxit, int gmres_dim,
                             int *info, int *iter, double *relres) {
  return solve_dev(solver, A, K, n, db, dx, tol, maxit, gmres_dim, info, iter, relres);
}

extern "C" int psp_solve(int solver, const psp_operator *A, const psp_operator *K, int n,
                         const double *b, double *x, double tol, int maxit, int gmres_dim,
                         int *info, int *iter, double *relres) {
  if (n < 0 || (n > 0 && (!b || !x))) { set_error("psp_solve: NULL vector"); return PSP_ERR_ARG; }
  PSP_TRY(ensure_init());
  if (wants_multi(solver, A, K, n))                 // every GPU takes its slices straight from the host arrays
    return multi_solve(solver, A->mat, K ? K->prec : nullptr, b, x, false, tol, maxit, gmres_dim, info, iter, relres);
  double *db = nullptr, *dx = nullptr;
  const bool pooled = (ctx().solve_depth == 0);
  if (pooled) {
    PSP_TRY(pool_get(0, (size_t)n, &db));
    PSP_TRY(pool_get(1, (size_t)n, &dx));
  } else {
    PSP_TRY(psp_vec_alloc((size_t)n, &db));
    int sa = psp_vec_alloc((size_t)n, &dx);
    if (sa != PSP_OK) { psp_vec_free(db); return sa; }
  }
  int st = psp_vec_upload(db, b, (size_t)n, 1);
  if (st == PSP_OK) st = psp_vec_upload(dx, x, (size_t)n, 1);
  if (st == PSP_OK) st = solve_dev(solver, A, K, n, db, dx, tol, maxit, gmres_dim, info, iter, relres);
  if (st == PSP_OK) st = psp_vec_download(dx, x, (size_t)n, 1);
  if (!pooled) { psp_vec_free(db); psp_vec_free(dx); }
  return st;
}
