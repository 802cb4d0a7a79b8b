// psp_correq.cuh -- device-resident Jacobi-Davidson correction-equation operator (psp_correq.cu).
#pragma once
#include "psp_mat.cuh"

struct psp_correq {
  int optype = 2;           // 1 UNSYM, 2 SYM (JDSYM_OP_*, pysparse/eigen/src/jdsym.h)
  int n = 0, kmax = 0;
  int k = 0;                // locked + active columns in use
  double theta = 0.0;
  psp_mat *A = nullptr, *M = nullptr;     // borrowed
  psp_prec *K = nullptr;                  // borrowed
  double *dQ = nullptr, *dQm = nullptr, *dY = nullptr;    // n x kmax, column-major
  double *dG = nullptr;     // inv(H), row-major with leading dimension kCeqMaxK
  double *d_w = nullptr, *d_t = nullptr, *d_t2 = nullptr, *d_part = nullptr;
  int nblk = 0;             // rows of d_part (upper bound of the projection kernels' grid)
  psp_mat *as_mat = nullptr;              // the handles the solvers are given (their ceq points back here)
  psp_prec *as_prec = nullptr;
};

namespace psp {

constexpr int kCeqMaxK = 64;

// role 0: y = matvec(x), role 1: y = precon(x); S == nullptr: not inside a solve (never gated)
int correq_apply(psp_correq *E, int role, const double *dx, double *dy, State *S, int gate);
int prec_apply_dev(psp_prec *K, const double *dx, double *dy, State *gateS);   // psp_precon.cu

}  // namespace psp
