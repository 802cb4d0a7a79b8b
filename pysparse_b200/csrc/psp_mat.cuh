// psp_mat.cuh -- device-resident matrix / preconditioner objects.
#pragma once
#include "psp_common.cuh"

struct psp_mat {
  int kind = 0;             // 1 CSR, 2 SSS
  int m = 0, n = 0;
  long long nnz = 0;        // stored entries as the reference counts them (SSS: strict lower)
  // CSR arrays on the device.  For SSS this is the symmetric expansion in
  // ll_mat.to_csr() order: [lower row i ascending | diag | column-i entries of
  // later rows ascending]  (ll_mat.c:1601-1623) -- the same accumulation order
  // sss_matvec (sss_mat.c:40-56) produces for y[i].
  long long fnnz = 0;
  int *d_ind = nullptr, *d_col = nullptr;
  double *d_val = nullptr;
  // row-block schedule of the streaming SpMV kernel
  int nb = 0;
  int nint = 0;             // leading blocks that read no halo column (== nb unless sharded)
  int *d_blk = nullptr;
  int *d_blk_halo = nullptr; // sharded slabs: per block, [first, end) of the halo positions it reads (int2)
  // SSS originals (strict lower CSR + diagonal), kept for the SSOR sweeps
  int *d_lind = nullptr, *d_lcol = nullptr;
  double *d_lval = nullptr, *d_diag = nullptr;
  // host copy of the SSS structure (level analysis) / CSR structure (transpose)
  std::vector<int> h_ind, h_col;
  std::vector<double> h_val, h_diag;
  // lazily built transpose (CSR of A^T), for matvec_transp
  psp_mat *transp = nullptr;
  // scratch vectors for the host-buffer entry points
  double *d_x = nullptr, *d_y = nullptr;
  size_t x_cap = 0, y_cap = 0;
};

struct psp_prec {
  int kind = 0;             // 1 Jacobi, 2 SSOR
  int n = 0;
  double omega = 1.0;
  int steps = 1;
  psp_mat *mat = nullptr;   // borrowed
  double *d_dinv = nullptr; // Jacobi
  double *d_t1 = nullptr, *d_t2 = nullptr;
  // SSOR: level schedules
  int nlev_f = 0, nlev_b = 0;
  int *d_rows_f = nullptr, *d_lptr_f = nullptr;   // rows sorted by forward level, level pointers
  int *d_rows_b = nullptr, *d_lptr_b = nullptr;
  // level-major copies of the sweep operands: position p of the forward (backward)
  // schedule is row rows_f[p] (rows_b[p]); its entries are f_col/f_val[f_ind[p]..f_ind[p+1])
  // = row of L in ascending column order (b_*: row of L^T in DESCENDING original-row
  // order, the order the reference's backward loop scatters in).  Column indices
  // stay in the original numbering (vectors are not permuted).
  int *d_find = nullptr, *d_fcol = nullptr, *d_bind = nullptr, *d_bcol = nullptr;
  int *d_floc = nullptr, *d_bloc = nullptr;       // per entry: position in the previous level, or -1
  double *d_fval = nullptr, *d_bval = nullptr;
  int maxw_f = 0, maxw_b = 0;                     // widest level of each schedule
  std::vector<int> h_level_f, h_level_b;
  unsigned *d_bar = nullptr;                      // grid barrier counter
  double *d_x = nullptr, *d_y = nullptr;
};

namespace psp {

// epilogue selector for the fused SpMV reductions (see psp_epilogue.cuh)
struct SpmvDots {
  const double *w1 = nullptr;   // red[0] = sum w1[i] * y[i]   (nullptr: none)
  const double *w2 = nullptr;   // red[1] = sum w2[i] * y[i]
  int w2_is_y = 0;              // red[1] = sum y[i] * y[i]
  int epilogue = 0;             // scalar epilogue run by the last block
  int gate = 0;                 // 1: also skip when S->cyc_done (GMRES inner loop)
  State *S = nullptr;           // nullptr: plain SpMV (no done-check, no reduction)
};

int spmv_launch(const psp_mat *a, const double *dx, double *dy, const SpmvDots &dots);
int mat_transpose(psp_mat *a);     // builds a->transp lazily

}  // namespace psp
