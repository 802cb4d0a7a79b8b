// psp_mat.cuh -- device-resident matrix / preconditioner objects.
#pragma once
#include "psp_common.cuh"
#include "psp_multi.cuh"

struct psp_correq;          // psp_correq.cuh

struct psp_mat {
  int kind = 0;             // 1 CSR, 2 SSS, 3 operator facade (ceq != nullptr: no arrays at all)
  psp_correq *ceq = nullptr;
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
  bool irregular = false;   // mean row longer than 16 entries: the blocks take the entry-parallel path (psp_tma_ring.cuh)
  int nint = 0;             // leading blocks that read no halo column (== nb unless sharded)
  int *d_blk = nullptr;
  int *d_blk_halo = nullptr; // sharded slabs: per block, [first, end) of the halo positions it reads (int2)
  // "diagonal run" format (psp_runs.cuh), chosen by the analysis at creation when the runs are
  // long: fmt == 1, the CSR arrays above are then NOT on the device (d_ind/d_col/d_val stay
  // nullptr) and nb / nint describe the run blocks.  SSS: d_diag points into d_rval.
  int fmt = 0;
  double *d_rval = nullptr;
  int *d_rseg = nullptr, *d_rblk = nullptr;
  long long rval_len = 0;   // doubles in d_rval
  int nseg = 0;
  // SSS: the diagonal on the device (the SSOR sweeps read it; their level-major copies of the
  // lower triangle are built from the host arrays below)
  double *d_diag = nullptr;
  // host copy of the SSS structure (level analysis) / CSR structure (transpose)
  std::vector<int> h_ind, h_col;
  std::vector<double> h_val, h_diag;
  // lazily built transpose (CSR of A^T), for matvec_transp
  psp_mat *transp = nullptr;
  // scratch vectors for the host-buffer entry points
  double *d_x = nullptr, *d_y = nullptr;
  size_t x_cap = 0, y_cap = 0;
  // gpus > 1: the same matrix cut into row slabs, one per GPU (psp_multi.cu)
  psp::MultiMat *multi = nullptr;
};

struct psp_prec {
  int kind = 0;             // 1 Jacobi, 2 SSOR, 3 operator facade (ceq)
  psp_correq *ceq = nullptr;
  int n = 0;
  double omega = 1.0;
  int steps = 1;
  psp_mat *mat = nullptr;   // borrowed
  double *d_dinv = nullptr; // Jacobi
  std::vector<double> h_dinv;               // host copy (the multi-GPU path uploads one slice per device)
  std::vector<psp_prec *> multi_prec;       // per-rank slices, built on first use
  std::vector<int> multi_rowptr;            // the row partition they were built for
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
  // point-to-point ("wavefront strip") sweep engine, psp_sweep_p2p.cuh: per direction the chain
  // schedule in warp-slot-lane order; p2p_ok == false: not applicable to this matrix
  struct P2PSched {
    int nwarps = 0, D = 0;
    long long nslots = 0;
    int *d_wbase = nullptr;
    void *d_rec = nullptr;        // packed records (row, diagonal, dependencies) per (slot, lane)
  } p2p[2];                                        // [0] forward, [1] backward
  bool p2p_ok = false;
  int *d_abort = nullptr;
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
int runs_launch(const psp_mat *a, const double *dx, double *dy, const SpmvDots &dots);   // psp_runs.cu
// blocks per chunk of the run pass (a chunk is the unit of work a CTA takes)
inline int run_chunk(int nb, int resident_ctas) {
  int c = resident_ctas > 0 ? nb / (8 * resident_ctas) : 1;       // >= 8 rounds: the last one costs little
  return c < 1 ? 1 : (c > 8 ? 8 : c);
}
// analysis + upload of the run format; *accepted = false leaves A untouched (psp_runs.cu).
// gcol / grow0: global column of every entry and global index of row 0 (row slabs of sharded
// matrices: the runs are ordered by GLOBAL diagonal so that a slab adds in the same order as
// the unsharded matrix); own_cols >= 0: columns >= own_cols are halo columns.
int runs_try_csr(psp_mat *A, const int *ind, const int *col, const double *val, const int *gcol, int grow0,
                 int own_cols, bool *accepted);
int runs_try_sss(psp_mat *A, const int *ind, const int *col, const double *val, const double *diag, bool *accepted);
// inverse of the analysis (CSR matrices): expand the device runs into host CSR arrays
int runs_to_csr(const psp_mat *A, std::vector<int> &ind, std::vector<int> &col, std::vector<double> &val);
int mat_transpose(psp_mat *a);     // builds a->transp lazily

}  // namespace psp
