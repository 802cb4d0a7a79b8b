// psp_epilogue.cuh -- the scalar recurrences of the Krylov solvers.
//
// Every reduction kernel (fused BLAS-1 or SpMV+dot) ends in ONE thread of the
// last-arriving block holding the global sums in S->red[] / S->redmax.  That
// thread then runs the epilogue selected by the launcher: the handful of
// scalar operations (alpha = rho/pq, Givens rotations, convergence and
// breakdown tests ...) the reference performs on the host between BLAS calls.
// Keeping them here means a whole solve needs no host round trip: the host
// only enqueues kernels and polls S->done.
//
// Each case cites the reference lines it restates.
#pragma once
#include "psp_common.cuh"

namespace psp {

enum Epilogue {
  EPI_NONE = 0,
  // PCG (pysparse/itsolvers/src/pcg.c)
  EPI_PCG_NORMB,      // :57-67
  EPI_PCG_RESID0,     // :72-89
  EPI_PCG_RESID0_J,   // same, identity/Jacobi preconditioner fused: also rho of iteration 1
  EPI_PCG_RHO,        // :99-111
  EPI_PCG_PQ,         // :117-125
  EPI_PCG_UPDATE,     // :126-165
  EPI_PCG_UPDATE_J,   // same, Jacobi fused: also the next iteration's rho test
  // MINRES (minres.c)
  EPI_MINRES_R0, EPI_MINRES_BETA0, EPI_MINRES_ALPHA, EPI_MINRES_BETA, EPI_MINRES_LOOP,
  // QMRS (qmrs.c)
  EPI_QMRS_RHO0, EPI_QMRS_DELTA, EPI_QMRS_EPS, EPI_QMRS_RHO1, EPI_QMRS_LOOP,
  // BiCGSTAB (bicgstab.c)
  EPI_BICG_NORMB, EPI_BICG_RES0, EPI_BICG_ALPHA, EPI_BICG_OMEGA, EPI_BICG_RES,
  // CGS (cgs.c)
  EPI_CGS_INIT, EPI_CGS_ALPHA, EPI_CGS_RES, EPI_CGS_LOOP,
  // GMRES (gmres.c)
  EPI_GMRES_NORMB, EPI_GMRES_BETA, EPI_GMRES_H, EPI_GMRES_HLAST, EPI_GMRES_CYCLE,
};

// scalar slots in State::v
enum {
  // shared
  V_N2B = 0, V_TOLB = 1, V_NORMR = 2,
  // PCG
  V_RHO = 3, V_RHO1 = 4, V_ALPHA = 5, V_BETA = 6, V_PQ = 7,
  // MINRES
  M_R0 = 0, M_BETA = 1, M_BETA_OLD = 2, M_C = 3, M_C_OLD = 4, M_S = 5, M_S_OLD = 6, M_ETA = 7,
  M_RMR = 8, M_ALPHA = 9, M_K1 = 10, M_K2 = 11, M_R1 = 12, M_R2 = 13, M_R3 = 14, M_STEP = 15,
  M_INVBETA = 16,
  // QMRS
  Q_RHO0 = 0, Q_TAU = 1, Q_C0 = 2, Q_EPS = 3, Q_XI = 4, Q_THETA0 = 5, Q_ETA = 6, Q_RES0 = 7,
  Q_ERR = 8, Q_DELTA = 9, Q_CC = 10, Q_BETA = 11, Q_INV = 12, Q_CC2 = 13, Q_RHO1 = 14,
  // BiCGSTAB
  B_N2B = 0, B_RES0 = 1, B_RHO1 = 2, B_RHO2 = 3, B_ALPHA = 4, B_OMEGA = 5, B_BETA = 6, B_RES = 7,
  // CGS
  C_RHO = 0, C_B2 = 1, C_TOL2 = 2, C_ALPHA = 3, C_BETA = 4, C_RES = 5,
  // GMRES
  G_N2B = 0, G_RES0 = 1, G_BETA = 2, G_REL = 3, G_HK = 4, G_INVH = 5, G_SCALE0 = 6,
};

__device__ __forceinline__ void pcg_finish(State *S) {
  S->done = 1;
  S->iter = S->it;
  S->relres = S->v[V_NORMR] / S->v[V_N2B];
}

// rho test + beta of pcg.c:99-111, shared by EPI_PCG_RHO and the fused variant
__device__ __forceinline__ void pcg_rho_step(State *S, double rho_new) {
  S->v[V_RHO1] = S->v[V_RHO];
  S->v[V_RHO] = rho_new;
  if (rho_new == 0.0) { S->info = -2; pcg_finish(S); return; }
  if (S->it == 1) {
    S->v[V_BETA] = 0.0;
  } else {
    double beta = rho_new / S->v[V_RHO1];
    S->v[V_BETA] = beta;
    if (beta == 0.0) { S->info = -6; pcg_finish(S); }
  }
}

__device__ __forceinline__ void pcg_update_step(State *S, bool fused_jacobi) {
  // pcg.c:126-165
  double normr = sqrt(S->red[0]);
  S->v[V_NORMR] = normr;
  if (!S->stag) S->stag = (1.0 + S->redmax == 1.0);
  if (normr <= S->v[V_TOLB]) { S->info = 0; pcg_finish(S); return; }
  if (S->stag) { S->info = -5; pcg_finish(S); return; }
  S->it += 1;
  if (S->it > S->maxit) { pcg_finish(S); return; }   // info stays -1, iter = maxit+1
  if (fused_jacobi) pcg_rho_step(S, S->red[1]);
}

// GMRES plane rotations (gmres.c:47-71)
__device__ __forceinline__ void rot_make(double a, double b, double *c, double *s) {
  if (b == 0.0) { *c = 1.0; *s = 0.0; }
  else if (fabs(b) > fabs(a)) { double t = a / b; *s = 1.0 / sqrt(1.0 + t * t); *c = t * *s; }
  else { double t = b / a; *c = 1.0 / sqrt(1.0 + t * t); *s = t * *c; }
}
__device__ __forceinline__ void rot_apply(double *a, double *b, double c, double s) {
  double t = c * *a + s * *b;
  *b = -s * *a + c * *b;
  *a = t;
}

// GMRES small dense state (Hessenberg matrix, rotations, rhs): device buffers
// whose addresses sit in a GmresSmall placed right behind the State.
struct GmresSmall {
  double *H;    // (m+1) x m, column-major, ld = m+1   (gmres.c:44 H(i,j))
  double *g;    // m+1   (reference name: s)
  double *cs;   // m
  double *sn;   // m
};

__device__ inline void run_epilogue(int id, State *S, void *aux) {
  double *v = S->v;
  const double *red = S->red;
  switch (id) {
  case EPI_NONE: break;

  // ------------------------------------------------------------- PCG
  case EPI_PCG_NORMB: {              // pcg.c:57-67  n2b = ||b||
    double n2b = sqrt(red[0]);
    v[V_N2B] = n2b;
    S->phase = (n2b == 0.0) ? 1 : 0; // 1: zero rhs -> x := 0 in the next kernel
    break;
  }
  case EPI_PCG_RESID0:
  case EPI_PCG_RESID0_J: {           // pcg.c:70-89
    if (S->phase == 1) { S->info = 0; S->relres = 0.0; S->iter = 0; S->done = 1; break; }
    S->info = -1;
    v[V_TOLB] = S->tol * v[V_N2B];
    double normr = sqrt(red[0]);
    v[V_NORMR] = normr;
    if (normr <= v[V_TOLB]) {
      S->info = 0; S->relres = normr / v[V_N2B]; S->iter = 0; S->done = 1; break;
    }
    v[V_RHO] = 1.0;
    S->stag = 0;
    S->it = 1;
    if (S->maxit < 1) { pcg_finish(S); break; }      // loop body never runs: iter = 1 = maxit+1
    if (id == EPI_PCG_RESID0_J) pcg_rho_step(S, red[1]);
    break;
  }
  case EPI_PCG_RHO: pcg_rho_step(S, red[0]); break;
  case EPI_PCG_PQ: {                 // pcg.c:117-125
    double pq = red[0];
    v[V_PQ] = pq;
    if (pq == 0.0) { S->info = -6; pcg_finish(S); break; }
    double alpha = v[V_RHO] / pq;
    v[V_ALPHA] = alpha;
    if (alpha == 0.0) S->stag = 1;
    break;
  }
  case EPI_PCG_UPDATE:   pcg_update_step(S, false); break;
  case EPI_PCG_UPDATE_J: pcg_update_step(S, true); break;

  // ------------------------------------------------------------- MINRES
  case EPI_MINRES_R0: {              // minres.c:71  norm_r0 = ||b - A x0||
    v[M_R0] = sqrt(red[0]);
    v[M_RMR] = v[M_R0];
    break;
  }
  case EPI_MINRES_BETA0: {           // minres.c:78-92 and the first loop test (:114)
    double beta = red[0];
    S->it = 0;
    if (beta < 0.0) { S->info = -3; S->done = 1; S->iter = 0; S->relres = nan(""); break; }
    beta = sqrt(beta);
    v[M_BETA] = beta; v[M_BETA_OLD] = 1.0;
    v[M_C] = 1.0; v[M_C_OLD] = 1.0; v[M_S] = 0.0; v[M_S_OLD] = 0.0;
    v[M_ETA] = beta;
    if (S->it >= S->maxit || v[M_RMR] < S->tol * v[M_R0]) {
      S->relres = v[M_RMR] / v[M_R0];
      S->info = (v[M_RMR] < S->tol * v[M_R0]) ? 0 : -1;
      S->iter = S->it; S->done = 1; break;
    }
    S->it = 1;
    break;
  }
  case EPI_MINRES_ALPHA: {           // minres.c:129-131
    double alpha = red[0];
    v[M_ALPHA] = alpha;
    v[M_K1] = alpha / v[M_BETA];
    v[M_K2] = v[M_BETA] / v[M_BETA_OLD];
    break;
  }
  case EPI_MINRES_BETA: {            // minres.c:141-192 (scalars; the w/x update kernel follows)
    double beta_old = v[M_BETA];
    double beta = red[0];
    if (beta < 0.0) { S->info = -3; S->done = 1; S->iter = S->it; S->relres = nan(""); break; }
    beta = sqrt(beta);
    v[M_BETA_OLD] = beta_old; v[M_BETA] = beta;
    double c_oold = v[M_C_OLD], s_oold = v[M_S_OLD];
    double c_old = v[M_C], s_old = v[M_S];
    v[M_C_OLD] = c_old; v[M_S_OLD] = s_old;
    double alpha = v[M_ALPHA];
    double r1_hat = c_old * alpha - c_oold * s_old * beta_old;
    double r1 = sqrt(r1_hat * r1_hat + beta * beta);
    double r2 = s_old * alpha + c_oold * c_old * beta_old;
    double r3 = s_oold * beta_old;
    if (r1 == 0.0) { S->info = -6; S->done = 1; S->iter = S->it; S->relres = nan(""); break; }
    double c = r1_hat / r1, s = beta / r1;
    v[M_C] = c; v[M_S] = s;
    v[M_R1] = r1; v[M_R2] = r2; v[M_R3] = r3;
    v[M_STEP] = c * v[M_ETA];
    v[M_ETA] = -s * v[M_ETA];
    v[M_RMR] *= fabs(s);
    break;
  }
  case EPI_MINRES_LOOP: {            // minres.c:114 loop test, :196-199 exit
    if (S->it >= S->maxit || v[M_RMR] < S->tol * v[M_R0]) {
      S->relres = v[M_RMR] / v[M_R0];
      S->info = (v[M_RMR] < S->tol * v[M_R0]) ? 0 : -1;
      S->iter = S->it; S->done = 1; break;
    }
    S->it += 1;
    break;
  }

  // ------------------------------------------------------------- QMRS
  case EPI_QMRS_RHO0: {              // qmrs.c:68-90
    double rho0 = sqrt(red[0]);
    v[Q_RHO0] = rho0; v[Q_TAU] = rho0; v[Q_RES0] = rho0;
    v[Q_C0] = 1.0; v[Q_EPS] = 1.0; v[Q_XI] = 1.0; v[Q_THETA0] = 0.0; v[Q_ETA] = -1.0;
    v[Q_ERR] = 1.0;
    S->it = 0;
    S->phase = 0;
    break;                           // first loop test: EPI_QMRS_LOOP after the init kernel
  }
  case EPI_QMRS_DELTA: {             // qmrs.c:95-99
    double delta = red[0];
    v[Q_DELTA] = delta;
    if (delta == 0.0) { S->info = -2; S->iter = S->it; S->relres = v[Q_ERR]; S->done = 1; break; }
    v[Q_CC] = v[Q_XI] * (delta / v[Q_EPS]);
    break;
  }
  case EPI_QMRS_EPS: {               // qmrs.c:106-107
    v[Q_EPS] = red[0];
    v[Q_BETA] = v[Q_EPS] / v[Q_DELTA];
    break;
  }
  case EPI_QMRS_RHO1: {              // qmrs.c:111-141 (scalars; the d/x/v update kernel follows)
    double rho1 = sqrt(red[0]);
    double beta = v[Q_BETA], c0 = v[Q_C0];
    v[Q_RHO1] = rho1;
    v[Q_XI] = rho1;
    bool bad = (c0 * fabs(beta) == 0.0);
    double theta = 0.0, c1 = 0.0;
    if (!bad) {
      theta = rho1 / (c0 * fabs(beta));
      c1 = 1.0 / sqrt(theta * theta + 1.0);
      bad = (beta * (c0 * c0) == 0.0);
    }
    if (!bad) {
      v[Q_ETA] = -v[Q_ETA] * v[Q_RHO0] * (c1 * c1) / (beta * (c0 * c0));
      v[Q_TAU] = v[Q_TAU] * theta * c1;
      bad = (rho1 == 0.0);
    }
    if (bad) { S->info = -6; S->iter = S->it; S->relres = v[Q_ERR]; S->done = 1; break; }
    double tc = v[Q_THETA0] * c1;
    v[Q_CC2] = tc * tc;
    v[Q_INV] = 1.0 / rho1;
    v[Q_RHO0] = rho1;
    v[Q_ERR] = v[Q_TAU] / v[Q_RES0];
    v[Q_C0] = c1;
    v[Q_THETA0] = theta;
    break;
  }
  case EPI_QMRS_LOOP: {              // qmrs.c:86-90 loop test, ++iter, eps0 test
    if (!(v[Q_ERR] > S->tol && S->it < S->maxit)) {
      S->phase = 2;                  // regular loop exit: host applies the final K (:146-153)
      S->iter = S->it; S->relres = v[Q_ERR];
      S->info = (v[Q_ERR] < S->tol) ? 0 : -1;
      S->done = 1; break;
    }
    S->it += 1;
    if (v[Q_EPS] == 0.0) { S->info = -6; S->iter = S->it; S->relres = v[Q_ERR]; S->done = 1; }
    break;
  }

  // ------------------------------------------------------------- BiCGSTAB
  case EPI_BICG_NORMB: {             // bicgstab.c:203-216
    double n2b = sqrt(red[0]);
    v[B_N2B] = n2b;
    S->info = -6;
    S->phase = (n2b == 0.0) ? 1 : 0;
    break;
  }
  case EPI_BICG_RES0: {              // bicgstab.c:235-248
    if (S->phase == 1) { S->info = 0; S->relres = 0.0; S->iter = 0; S->done = 1; break; }
    v[B_RES0] = sqrt(red[0]);
    v[B_OMEGA] = 0.0; v[B_BETA] = 0.0; v[B_ALPHA] = 0.0; v[B_RHO2] = 0.0;
    S->it = 1;
    v[B_RHO1] = red[0];              // rhat.r with rhat == r (:247)
    if (v[B_RHO1] == 0.0) { S->iter = 1; S->relres = nan(""); S->done = 1; }
    break;
  }
  case EPI_BICG_ALPHA: {             // bicgstab.c:264
    v[B_ALPHA] = v[B_RHO1] / red[0];
    break;
  }
  case EPI_BICG_OMEGA: {             // bicgstab.c:274
    v[B_OMEGA] = red[0] / red[1];
    break;
  }
  case EPI_BICG_RES: {               // bicgstab.c:278-293, and :247-253 of the next pass
    double res = sqrt(red[0]);
    v[B_RES] = res;
    if (v[B_OMEGA] == 0.0) { S->iter = S->it; S->relres = nan(""); S->done = 1; break; }
    v[B_RHO2] = v[B_RHO1];
    double rel = res / v[B_RES0];
    if (rel > S->tol && S->it < S->maxit) {
      S->it += 1;
      double rho1 = red[1];          // rhat.r accumulated in the same pass
      v[B_RHO1] = rho1;
      if (rho1 == 0.0) { S->iter = S->it; S->relres = nan(""); S->done = 1; break; }
      v[B_BETA] = (rho1 / v[B_RHO2]) * (v[B_ALPHA] / v[B_OMEGA]);
    } else {
      S->relres = rel;
      S->info = (rel >= S->tol) ? -1 : 0;
      S->iter = S->it;
      S->done = 1;
    }
    break;
  }

  // ------------------------------------------------------------- CGS
  case EPI_CGS_INIT: {               // cgs.c:54-64   red[0] = r0.r0, red[1] = b.b
    v[C_RHO] = red[0];
    v[C_B2] = red[1];
    v[C_TOL2] = S->tol * S->tol;
    S->it = 0;
    S->info = -1;
    if (v[C_RHO] < v[C_B2] * v[C_TOL2]) {
      S->relres = sqrt(v[C_RHO] / v[C_B2]); S->info = 0; S->iter = 0; S->done = 1; break;
    }
    if (!(S->it < S->maxit)) {       // loop never entered: *res is never assigned (cgs.c:109)
      S->relres = nan(""); S->iter = 0; S->done = 1;
    }
    break;
  }
  case EPI_CGS_ALPHA: {              // cgs.c:72
    v[C_ALPHA] = v[C_RHO] / red[0];
    break;
  }
  case EPI_CGS_RES: {                // cgs.c:89-98   red[0] = r.r, red[1] = r.r0
    double res = red[0];
    v[C_RES] = res;
    if (res < v[C_B2] * v[C_TOL2]) {
      S->relres = sqrt(res / v[C_B2]); S->info = 0; S->iter = S->it; S->done = 1; break;
    }
    double rho_new = red[1];
    v[C_BETA] = rho_new / v[C_RHO];
    v[C_RHO] = rho_new;
    break;
  }
  case EPI_CGS_LOOP: {               // cgs.c:64 loop increment / test, :108-110 exit
    S->it += 1;
    if (!(S->it < S->maxit)) {
      S->relres = sqrt(v[C_RES] / v[C_B2]); S->info = -1; S->iter = S->it; S->done = 1;
    }
    break;
  }

  // ------------------------------------------------------------- GMRES
  case EPI_GMRES_NORMB: {            // gmres.c:100-108
    double n2b = sqrt(red[0]);
    v[G_N2B] = n2b;
    S->phase = (n2b == 0.0) ? 1 : 0;
    S->it = 0;
    S->info = 0;
    break;
  }
  case EPI_GMRES_BETA: {             // gmres.c:114-125 (cycle start) / :187-193 (exit)
    if (S->phase == 1) { S->relres = 0.0; S->iter = 0; S->done = 1; break; }
    double beta = sqrt(red[0]);
    if (S->phase == 3) {             // true residual at exit
      S->iter = S->it; S->relres = beta / v[G_RES0]; S->done = 1; break;
    }
    GmresSmall *G = (GmresSmall *)aux;
    v[G_BETA] = beta;
    v[G_SCALE0] = -1.0 / beta;
    if (S->it == 0) v[G_RES0] = beta;
    for (int j = 1; j < S->mdim + 1; j++) G->g[j] = 0.0;
    G->g[0] = beta;
    S->inner = 0;
    S->cyc_done = 0;
    S->it += 1;                      // iter++ of the first inner step (:131)
    break;
  }
  case EPI_GMRES_H: {                // gmres.c:143  H(k,i) = <V(i+1), V(k)>
    GmresSmall *G = (GmresSmall *)aux;
    int k = S->phase, i = S->inner, ld = S->mdim + 1;
    G->H[(size_t)i * ld + k] = red[0];
    v[G_HK] = red[0];
    S->phase = k + 1;
    break;
  }
  case EPI_GMRES_HLAST: {            // gmres.c:148-174
    GmresSmall *G = (GmresSmall *)aux;
    int i = S->inner, ld = S->mdim + 1;
    double *Hc = G->H + (size_t)i * ld;
    Hc[i + 1] = sqrt(red[0]);
    v[G_INVH] = 1.0 / Hc[i + 1];
    for (int k = 0; k < i; k++) rot_apply(&Hc[k], &Hc[k + 1], G->cs[k], G->sn[k]);
    rot_make(Hc[i], Hc[i + 1], &G->cs[i], &G->sn[i]);
    rot_apply(&Hc[i], &Hc[i + 1], G->cs[i], G->sn[i]);
    rot_apply(&G->g[i], &G->g[i + 1], G->cs[i], G->sn[i]);
    double rel = fabs(G->g[i + 1]) / v[G_RES0];
    v[G_REL] = rel;
    S->phase = 0;
    bool stop = (rel <= S->tol) || !(i + 1 < S->mdim && S->it + 1 <= S->maxit);
    if (stop) {
      for (int j = i; j >= 0; j--) {          // back substitution, gmres.c:169-174
        G->g[j] /= G->H[(size_t)j * ld + j];
        for (int k = j - 1; k >= 0; k--) G->g[k] -= G->H[(size_t)j * ld + k] * G->g[j];
      }
      S->last_i = i;
      S->cyc_done = 1;
    } else {
      S->inner = i + 1;
      S->it += 1;
    }
    break;
  }
  case EPI_GMRES_CYCLE: {            // gmres.c:183 outer loop test (after the x update)
    if (!(v[G_REL] > S->tol && S->it + 1 <= S->maxit)) S->phase = 3;
    S->cyc_done = 0;
    break;
  }
  default: break;
  }
}

}  // namespace psp
