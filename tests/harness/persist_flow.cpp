// persist_flow.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Sequential host transcriptions of the control flow of the single-launch solver kernels
// (pysparse_b200/csrc/psp_persist.cu, psp_persist_bicg.cu, psp_persist_gmres.cu): the same
// phases in the same order, the same operands, the same scalar epilogues -- run_epilogue()
// and the State struct are compiled straight from the product headers -- with every
// "vector phase + grid all-reduce" replaced by a plain loop.  tests/test_persist_flow.py runs
// them against the CPU oracle: that checks the SEQUENCING of the kernels (which epilogue
// follows which phase, which scalar feeds which update, loop exits, restart logic) without a
// GPU.  What it cannot check is the parallel machinery (barriers, halo exchange, memory
// model); that is what the GPU parity tests are for.
//
// Build: g++ -O1 -std=c++17 -shared -fPIC -ffp-contract=off -I<cuda>/include -Iinclude
//        -Ipysparse_b200/csrc tests/harness/persist_flow.cpp -o <out>.so
#include <cmath>
#include <cstring>
#include <vector>
#define __device__
#define __forceinline__ inline
#include "psp_epilogue.cuh"

using namespace psp;

namespace {

struct Csr { int n; const int *ind, *col; const double *val; };

void spmv(const Csr &A, const double *x, double *y) {
  for (int i = 0; i < A.n; i++) {
    double s = 0.0;
    for (int k = A.ind[i]; k < A.ind[i + 1]; k++) s += A.val[k] * x[A.col[k]];
    y[i] = s;
  }
}

State make_state(double tol, int maxit, int mdim) {   // psp_solvers.cu: State init of solve_dev
  State S;
  std::memset(&S, 0, sizeof S);
  S.maxit = maxit; S.tol = tol; S.mdim = mdim;
  S.iter = -1; S.relres = std::nan("");
  return S;
}

void finish(const State &S, int *info, int *iter, double *relres) { *info = S.info; *iter = S.iter; *relres = S.relres; }

}  // namespace

// ---------------------------------------------------------------- k_pcg_persist (validated kernel: checks the method)
extern "C" int flow_pcg(int n, const int *ind, const int *col, const double *val, const double *dinv, const double *b,
                        double *x, double tol, int maxit, int *info, int *iter, double *relres) {
  Csr A{n, ind, col, val};
  State S = make_state(tol, maxit, 0);
  std::vector<double> r(n), p(n), q(n);
  double acc0 = 0.0, acc1 = 0.0;
  for (int i = 0; i < n; i++) acc0 += b[i] * b[i];
  S.red[0] = acc0; run_epilogue(EPI_PCG_NORMB, &S, nullptr);
  spmv(A, x, r.data());
  acc0 = acc1 = 0.0;
  for (int i = 0; i < n; i++) {
    if (S.phase == 1) { x[i] = 0.0; continue; }
    double t = b[i] - r[i];
    r[i] = t; acc0 += t * t;
    double z = dinv ? t * dinv[i] : t;
    acc1 += t * z;
  }
  S.red[0] = acc0; S.red[1] = acc1; run_epilogue(EPI_PCG_RESID0_J, &S, nullptr);
  while (!S.done) {
    const double beta = S.v[V_BETA];
    const bool first = (S.it == 1);
    for (int i = 0; i < n; i++) {
      double zi = dinv ? r[i] * dinv[i] : r[i];
      p[i] = first ? zi : zi + beta * p[i];
    }
    spmv(A, p.data(), q.data());
    acc0 = 0.0;
    for (int i = 0; i < n; i++) acc0 += p[i] * q[i];
    S.red[0] = acc0; run_epilogue(EPI_PCG_PQ, &S, nullptr);
    if (S.done) break;
    const double alpha = S.v[V_ALPHA];
    const bool scan = (S.stag == 0);
    double mx = 0.0;
    acc0 = acc1 = 0.0;
    for (int i = 0; i < n; i++) {
      double xi = x[i], pi = p[i];
      if (scan) {
        if (xi != 0.0) mx = std::fmax(mx, std::fabs(alpha * pi / xi));
        else if (pi != 0.0) mx = std::fmax(mx, 1.0);
      }
      x[i] = xi + alpha * pi;
      double t = r[i] + (-alpha) * q[i];
      r[i] = t; acc0 += t * t;
      double z = dinv ? t * dinv[i] : t;
      acc1 += t * z;
    }
    S.red[0] = acc0; S.red[1] = acc1; S.redmax = mx; run_epilogue(EPI_PCG_UPDATE_J, &S, nullptr);
  }
  finish(S, info, iter, relres);
  return 0;
}

// ---------------------------------------------------------------- k_bicgstab_persist
extern "C" int flow_bicgstab(int n, const int *ind, const int *col, const double *val, const double *dinv,
                             const double *b, double *x, double tol, int maxit, int *info, int *iter, double *relres) {
  Csr A{n, ind, col, val};
  State S = make_state(tol, maxit, 0);
  const bool ident = (dinv == nullptr);
  std::vector<double> r(n), rhat(n), p(n), phatb(n), v(n), s(n), shatb(n), t(n);
  double *x1 = ident ? p.data() : phatb.data();      // phat (== p when K = I; single GPU: p updated in place)
  double *x2 = ident ? s.data() : shatb.data();      // shat
  double acc0 = 0.0, acc1 = 0.0;
  for (int i = 0; i < n; i++) acc0 += b[i] * b[i];
  S.red[0] = acc0; run_epilogue(EPI_BICG_NORMB, &S, nullptr);
  spmv(A, x, r.data());
  acc0 = 0.0;
  for (int i = 0; i < n; i++) {
    if (S.phase == 1) { x[i] = 0.0; continue; }
    double tt = b[i] + -r[i];
    r[i] = tt; rhat[i] = tt; acc0 += tt * tt;
  }
  S.red[0] = acc0; run_epilogue(EPI_BICG_RES0, &S, nullptr);
  while (!S.done) {
    {
      const double beta = S.v[B_BETA], omega = S.v[B_OMEGA];
      const bool first = (S.it == 1);
      for (int i = 0; i < n; i++) {
        const double pn = first ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
        p[i] = pn;
        if (!ident) phatb[i] = pn * dinv[i];
      }
    }
    spmv(A, x1, v.data());
    acc0 = 0.0;
    for (int i = 0; i < n; i++) acc0 += rhat[i] * v[i];
    S.red[0] = acc0; run_epilogue(EPI_BICG_ALPHA, &S, nullptr);
    {
      const double alpha = S.v[B_ALPHA];
      for (int i = 0; i < n; i++) {
        const double sn = r[i] + (-alpha) * v[i];
        s[i] = sn;
        if (!ident) shatb[i] = sn * dinv[i];
      }
    }
    spmv(A, x2, t.data());
    acc0 = acc1 = 0.0;
    for (int i = 0; i < n; i++) { acc0 += s[i] * t[i]; acc1 += t[i] * t[i]; }
    S.red[0] = acc0; S.red[1] = acc1; run_epilogue(EPI_BICG_OMEGA, &S, nullptr);
    {
      const double alpha = S.v[B_ALPHA], omega = S.v[B_OMEGA];
      acc0 = acc1 = 0.0;
      for (int i = 0; i < n; i++) {
        x[i] = x[i] + alpha * x1[i] + omega * x2[i];
        const double rn = s[i] - omega * t[i];
        r[i] = rn; acc0 += rn * rn; acc1 += rhat[i] * rn;
      }
      S.red[0] = acc0; S.red[1] = acc1; run_epilogue(EPI_BICG_RES, &S, nullptr);
    }
  }
  finish(S, info, iter, relres);
  return 0;
}

// ---------------------------------------------------------------- k_gmres_persist
extern "C" int flow_gmres(int n, const int *ind, const int *col, const double *val, const double *dinv, const double *b,
                          double *x, double tol, int maxit, int m, int *info, int *iter, double *relres) {
  Csr A{n, ind, col, val};
  State S = make_state(tol, maxit, m);
  const bool ident = (dinv == nullptr);
  const size_t N = (size_t)n;
  std::vector<double> V(N * (m + 1)), Wb(ident ? 0 : N * m), H((size_t)(m + 1) * m), g(m + 1), cs(m), sn(m);
  double *W = ident ? V.data() : Wb.data();
  GmresSmall G{H.data(), g.data(), cs.data(), sn.data()};
  void *aux = &G;
  auto reduce1 = [&](double vv, int epi) { S.red[0] = vv; run_epilogue(epi, &S, aux); };
  double acc = 0.0;
  for (int i = 0; i < n; i++) acc += b[i] * b[i];
  reduce1(acc, EPI_GMRES_NORMB);
  for (;;) {
    spmv(A, x, V.data());
    acc = 0.0;
    {
      const bool zero_rhs = (S.phase == 1);
      for (int i = 0; i < n; i++) {
        if (zero_rhs) { x[i] = 0.0; continue; }
        const double tt = V[i] + (-1.0) * b[i];
        V[i] = tt; acc += tt * tt;
      }
    }
    reduce1(acc, EPI_GMRES_BETA);
    if (S.done) break;
    for (int i = 0; i < m; i++) {
      if (S.cyc_done) break;
      double *vi = V.data() + (size_t)i * N, *vn = V.data() + (size_t)(i + 1) * N;
      double *wi = ident ? vi : W + (size_t)i * N;
      {
        const double f = S.v[i == 0 ? G_SCALE0 : G_INVH];
        for (int e = 0; e < n; e++) {
          const double tt = f * vi[e];
          vi[e] = tt;
          if (!ident) wi[e] = tt * dinv[e];
        }
      }
      spmv(A, wi, vn);
      acc = 0.0;
      for (int e = 0; e < n; e++) acc += vn[e] * V[e];
      reduce1(acc, EPI_GMRES_H);
      for (int k = 0; k <= i; k++) {
        const double h = S.v[G_HK];
        const double *vk = V.data() + (size_t)k * N;
        const double *vnext = (k < i) ? V.data() + (size_t)(k + 1) * N : nullptr;
        acc = 0.0;
        for (int e = 0; e < n; e++) {
          const double tt = vn[e] + (-h) * vk[e];
          vn[e] = tt;
          acc += tt * (vnext ? vnext[e] : tt);
        }
        reduce1(acc, k < i ? EPI_GMRES_H : EPI_GMRES_HLAST);
      }
    }
    {
      const int last = S.last_i;
      for (int e = 0; e < n; e++) {
        double xi = x[e];
        for (int j = 0; j <= last; j++) xi = xi + g[j] * W[(size_t)j * N + e];
        x[e] = xi;
      }
      run_epilogue(EPI_GMRES_CYCLE, &S, aux);
    }
  }
  finish(S, info, iter, relres);
  return 0;
}

// ---------------------------------------------------------------- scalar state for Python-driven flows
// (tests/test_persist_flow.py emulates the SHARDED kernels rank by rank in NumPy and needs the
// product's scalar logic: one State shared by all emulated ranks, as on the GPUs where every
// rank computes identical scalars)
extern "C" void *st_new(double tol, int maxit, int mdim) { return new State(make_state(tol, maxit, mdim)); }
extern "C" void st_free(void *s) { delete (State *)s; }
extern "C" void st_epi(void *s, int epi, double r0, double r1, double mx) {
  State *S = (State *)s;
  S->red[0] = r0; S->red[1] = r1; S->redmax = mx;
  run_epilogue(epi, S, nullptr);
}
extern "C" double st_v(void *s, int idx) { return ((State *)s)->v[idx]; }
extern "C" double st_relres(void *s) { return ((State *)s)->relres; }
extern "C" int st_int(void *s, int which) {
  const State *S = (const State *)s;
  switch (which) {
    case 0: return S->done;
    case 1: return S->info;
    case 2: return S->iter;
    case 3: return S->it;
    case 4: return S->phase;
    case 5: return S->stag;
    default: return -1;
  }
}
extern "C" int epi_id(const char *name) {
#define PSP_EPI(x) if (!std::strcmp(name, #x)) return (int)x;
  PSP_EPI(EPI_PCG_NORMB) PSP_EPI(EPI_PCG_RESID0_J) PSP_EPI(EPI_PCG_PQ) PSP_EPI(EPI_PCG_UPDATE_J)
  PSP_EPI(EPI_BICG_NORMB) PSP_EPI(EPI_BICG_RES0) PSP_EPI(EPI_BICG_ALPHA) PSP_EPI(EPI_BICG_OMEGA) PSP_EPI(EPI_BICG_RES)
  PSP_EPI(V_ALPHA) PSP_EPI(V_BETA) PSP_EPI(B_ALPHA) PSP_EPI(B_BETA) PSP_EPI(B_OMEGA)
#undef PSP_EPI
  return -1;
}
