/* psp_b200.h -- C-ABI of libpsp_b200.so, the B200 (sm_100a) implementation of
 * PySparse's data-parallel hot path: fp64 CSR / SSS SpMV, the six Krylov
 * solvers and the Jacobi / SSOR preconditioners.
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes
 * plain pointers and sizes, and replaces one interface of the reference
 * (cited as file:line under /root/reference/).  The CPython glue in
 * pysparse_b200/host/ (modules spmatrix, krylov, precon) is a thin caller of
 * these functions; INTEGRATION.md shows the binding a PySparse maintainer
 * would add to the reference's own C sources.
 *
 * Conventions
 *  - indices are 32-bit C int, 0-based; values are IEEE fp64;
 *  - every function returns PSP_OK (0) or a negative psp_status; the message
 *    of the last failure is available from psp_last_error();
 *  - "host" pointers are ordinary (pageable or pinned) CPU memory, "dev"
 *    pointers come from psp_vec_alloc();
 *  - there is NO CPU fallback: without a CUDA device every compute call
 *    fails with PSP_ERR_NO_DEVICE.
 */
#ifndef PSP_B200_H
#define PSP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum psp_status {
  PSP_OK = 0,
  PSP_ERR_NO_DEVICE = -1,   /* no usable CUDA device / driver            */
  PSP_ERR_CUDA = -2,        /* a CUDA runtime call failed                 */
  PSP_ERR_ARG = -3,         /* invalid argument (shape, NULL, ...)        */
  PSP_ERR_NOMEM = -4,       /* host or device allocation failed           */
  PSP_ERR_CALLBACK = -5,    /* a foreign operator callback reported -1    */
  PSP_ERR_STRUCTURE = -6,   /* matrix structure invalid (unsorted, range) */
  PSP_ERR_DIST = -7         /* multi-GPU layer (NCCL / peer access) error */
} psp_status;

typedef struct psp_mat psp_mat;     /* device-resident sparse matrix        */
typedef struct psp_prec psp_prec;   /* device-resident preconditioner       */

/* ---- runtime ---------------------------------------------------------- */

/* Select the CUDA device and create the context (streams, scratch).  Called
 * implicitly (device 0, or $PSP_DEVICE) by the first compute call. */
int psp_init(int device);
int psp_device_count(void);             /* 0 when there is no GPU/driver     */
int psp_current_device(void);           /* -1 before psp_init                */
const char *psp_last_error(void);
const char *psp_version(void);
/* number of kernels this library has launched since the last reset (the
 * `gpu_launches` figure of bench.py) */
long long psp_launch_count(int reset);
int psp_synchronize(void);

/* pinned host memory for fast host<->device staging (optional) */
int psp_host_alloc(size_t bytes, void **out);
int psp_host_free(void *p);

/* ---- matrices --------------------------------------------------------- */

/* CSR matrix: replaces CSRMatObject (pysparse/include/csr_mat.h:6-13) as
 * produced by ll_mat.to_csr() (pysparse/sparse/src/ll_mat.c:1577-1648).
 * ind[m+1], col[nnz], val[nnz]; arrays are copied to the device. */
int psp_csr_create(int m, int n, int nnz, const int *ind, const int *col,
                   const double *val, psp_mat **out);

/* SSS matrix (strict lower triangle in CSR + dense diagonal): replaces
 * SSSMatObject (pysparse/include/sss_mat.h:6-14) as produced by
 * ll_mat.to_sss() (ll_mat.c:1654-1708).  The one-time transpose index that
 * makes the upper-triangle part an atomic-free gather is built here. */
int psp_sss_create(int n, int nnz, const int *ind, const int *col,
                   const double *val, const double *diag, psp_mat **out);

void psp_mat_destroy(psp_mat *a);
/* kind: 1 = CSR, 2 = SSS */
int psp_mat_info(const psp_mat *a, int *m, int *n, long long *nnz, int *kind);

/* y := A x (transp = 0) or y := A^T x (transp != 0) with HOST vectors and
 * element strides incx / incy: replaces csr_matvec_kernel(_stride) /
 * csr_matvec_transp_kernel(_stride) (csr_mat.c:17-106) and sss_matvec(_stride)
 * (sss_mat.c:40-76) behind CSRMat_matvec / SSSMat_matvec. */
int psp_matvec(psp_mat *a, int transp, const double *x, long incx, double *y, long incy);

/* Host-only analysis behind the "diagonal run" device layout of a matrix (csrc/psp_runs.cuh): the
 * rows of csr_matvec_kernel / sss_matvec re-expressed as runs of consecutive rows on one diagonal.
 * sss = 0: CSR arrays (diag ignored); sss = 1: strict lower triangle + diagonal, every stored
 * value kept once and referenced by two run descriptors.  Outputs (malloc'ed, release with
 * psp_runs_free): blk[4 nblk] = (first row, rows | 1<<30 when a run of the block covers a bridged gap,
 * first run, end run), seg[4 nseg] = (position of
 * the first value, first column, first local row, length), rval[nval].  Bit-exact structure
 * work; needs no GPU (the CPU tests rebuild the matrix from it). */
int psp_runs_analyze(int sss, int m, const int *ind, const int *col, const double *val, const double *diag,
                     int *nblk, int *nseg, long long *nval, int **blk, int **seg, double **rval);
void psp_runs_free(void *p);

/* same with DEVICE vectors (contiguous); asynchronous on the library stream */
int psp_matvec_dev(psp_mat *a, int transp, const double *dx, double *dy);

/* ---- device vectors (used by benchmarks and the callback path) -------- */
int psp_vec_alloc(size_t n, double **dptr);
int psp_vec_free(double *dptr);
int psp_vec_upload(double *dptr, const double *host, size_t n, long inc);
int psp_vec_download(const double *dptr, double *host, size_t n, long inc);
int psp_vec_fill(double *dptr, size_t n, double value);

/* ---- preconditioners --------------------------------------------------- */

/* Jacobi: replaces newJacobiObject + jacobi() (pysparse/precon/src/
 * preconmodule.c:352-412, :35-54).  diag[n] are the matrix diagonal entries
 * A[i,i]; dinv = omega/diag is formed here; PSP_ERR_ARG with the reference's
 * message "diagonal element close to zero" when 1.0 + d == 1.0.  `a` is only
 * needed (and may otherwise be NULL) for steps > 1. */
int psp_jacobi_create(psp_mat *a, int n, const double *diag, double omega, int steps,
                      psp_prec **out);

/* SSOR / symmetric Gauss-Seidel: replaces newSSORObject + ssor_kernel /
 * symgs_kernel (preconmodule.c:414-459, :95-193).  `sss` must be an SSS
 * matrix.  Runs the level-set analysis of L (forward sweep) and L^T (backward
 * sweep); the sweeps keep the reference's lexicographic dependency order. */
int psp_ssor_create(psp_mat *sss, double omega, int steps, psp_prec **out);

/* (Re)bind the matrix a multi-step Jacobi preconditioner multiplies with (the host layer calls
 * it before every use: the device copy of a mutable ll_mat may have been rebuilt). */
int psp_prec_set_matrix(psp_prec *k, psp_mat *a);

void psp_prec_destroy(psp_prec *k);

/* y := K^-1 x with HOST vectors (contiguous): Jacobi_precon / SSOR_precon
 * (preconmodule.c:60-80, :199-223) */
int psp_prec_apply(psp_prec *k, const double *x, double *y);
int psp_prec_apply_dev(psp_prec *k, const double *dx, double *dy);

/* level-set structure of an SSOR preconditioner (bit-exact integer work;
 * exposed for the parity tests): number of forward / backward levels and,
 * optionally, the level of every row (arrays of n ints, may be NULL). */
int psp_ssor_levels(const psp_prec *k, int *nlev_fwd, int *nlev_bwd,
                    int *level_fwd, int *level_bwd);

/* ---- Krylov solvers ---------------------------------------------------- */

typedef enum psp_solver {
  PSP_PCG = 0,       /* Itsolvers_pcg_kernel      pysparse/itsolvers/src/pcg.c:22-171      */
  PSP_MINRES = 1,    /* Itsolvers_minres_kernel   .../minres.c:43-200                      */
  PSP_QMRS = 2,      /* Itsolvers_qmrs_kernel     .../qmrs.c:34-154                        */
  PSP_BICGSTAB = 3,  /* Itsolvers_bicgstab_kernel .../bicgstab.c:182-300                   */
  PSP_CGS = 4,       /* Itsolvers_cgs_kernel      .../cgs.c:13-111                         */
  PSP_GMRES = 5      /* Itsolvers_gmres_kernel    .../gmres.c:73-206                       */
} psp_solver;

/* A foreign (duck-typed) operator: y := Op(x) on HOST buffers of length n.
 * Return 0 on success, non-zero on failure (mirrors SpMatrix_Matvec /
 * SpMatrix_Precon, pysparse/sparse/src/spmatrixmodule.c:169-248). */
typedef int (*psp_callback)(void *ctx, int n, const double *x, double *y);

/* An operator is EITHER a native handle (solve stays on the device) OR a
 * callback (vectors are staged through host memory around each call). */
typedef struct psp_operator {
  psp_mat *mat;        /* native matrix, or NULL                      */
  psp_prec *prec;      /* native preconditioner, or NULL              */
  psp_callback cb;     /* foreign operator, used when both are NULL   */
  void *ctx;
} psp_operator;

/* Solve A x = b.  Replaces the six wrappers ItSolvers_<solver>
 * (pysparse/itsolvers/src/itsolversmodule.c:32-586) together with the
 * kernels they call: same start vector convention (x is the initial guess,
 * except QMRS which ignores it), same stopping rules, same
 * (info, iter, relres) results including the reference's quirks
 * (PCG iter = maxit+1 on non-convergence, 0-based CGS count, GMRES info
 * always 0 ...).  b and x are HOST arrays of length n; x is overwritten.
 * K may be NULL (no preconditioner).  gmres_dim is only read for PSP_GMRES.
 * Where the reference leaves relres uninitialised (MINRES / BiCGSTAB
 * breakdown) NaN is returned. */
int psp_solve(int solver, const psp_operator *A, const psp_operator *K, int n,
              const double *b, double *x, double tol, int maxit, int gmres_dim,
              int *info, int *iter, double *relres);

/* Same, with b and x resident on the device (no host<->device vector
 * traffic); used to time the solve with inputs already in HBM. */
int psp_solve_dev(int solver, const psp_operator *A, const psp_operator *K, int n,
                  const double *db, double *dx, double tol, int maxit, int gmres_dim,
                  int *info, int *iter, double *relres);

/* ---- Jacobi-Davidson correction equation (consumer of the path) --------- */

/* The operator JDSYM hands its inner Krylov solver as both A and K:
 * replaces CorrEq_New / CorrEq_matvec / CorrEq_precon (pysparse/eigen/src/correq.c:28-262;
 * called from jdsym.c:570) with the search-space blocks Q, M Q and Y = K^-1 (M) Q
 * resident on the device, so that an inner solve never crosses PCIe.
 * optype: 1 = JDSYM_OP_UNSYM, 2 = JDSYM_OP_SYM; m and k may be NULL; kmax <= 64.
 * The matrices and the preconditioner are borrowed and must outlive the operator. */
typedef struct psp_correq psp_correq;
int psp_correq_create(int optype, psp_mat *a, psp_mat *m, psp_prec *k, int kmax, psp_correq **out);
void psp_correq_destroy(psp_correq *e);

/* Upload columns [col0, col0 + ncols) of Q (which = 0), M Q (1) or Y (2) from a
 * column-major HOST array with leading dimension ld >= n. */
int psp_correq_set_columns(psp_correq *e, int which, int col0, int ncols, const double *host, long ld);

/* Select the number of columns in use and the shift (CorrEq_update, correq.c:94-110).
 * hinv: inverse of the leading k x k block of H = Qb' Y, row-major with leading
 * dimension k (the reference keeps an LU factorisation instead); required when the
 * operator has a preconditioner and k > 0, ignored otherwise. */
int psp_correq_update(psp_correq *e, int k, double theta, const double *hinv);

/* The handles to put into a psp_operator (mat for A, prec for K), or to call
 * psp_matvec / psp_prec_apply with.  They belong to the operator: destroying
 * them is a no-op. */
psp_mat *psp_correq_mat(psp_correq *e);
psp_prec *psp_correq_prec(psp_correq *e);

/* Time of the device part of the most recent psp_solve / psp_solve_dev /
 * psp_matvec call, measured with CUDA events on the library stream (ms). */
double psp_last_device_ms(void);

/* Timing helpers on the library stream (the stream every kernel of this
 * library is launched on): record one of 8 CUDA events / elapsed ms between
 * two recorded events (synchronises on `stop`). */
int psp_event_record(int idx);
int psp_event_elapsed_ms(int start, int stop, double *ms);

/* Sampled SpMV kernel timing: returns the number of SpMV launches bracketed by
 * event pairs since the last call and their summed duration, then re-arms
 * sampling to every `sample_every`-th SpMV launch (0 = off, the default). */
int psp_spmv_timing(int sample_every, int *count, double *total_ms);

/* ---- multi-GPU (one process per GPU; not in the reference) --
 *
 * Rows are sharded in contiguous blocks: rank r owns rows
 * [rowptr[r], rowptr[r+1]) of A and the same slices of every vector.  Work
 * vectors live in a symmetric heap that all ranks of the node map through
 * CUDA IPC; halo values are pulled from the owners over NVLink and scalar
 * all-reduces happen inside the reduction kernels through peer mailboxes.
 * The launcher (torchrun + torch.distributed in bench.py / tests) only
 * all-gathers the 128-byte IPC handle blobs. */
typedef struct psp_dmat psp_dmat;   /* this rank's row slab of a sharded CSR matrix */

/* Host-only planning of one rank's slab (bit-exact integer work, no GPU
 * needed): given the row partition and the slab's CSR with GLOBAL column
 * indices, produce the local column numbering (own columns first, then the
 * halo columns in ascending global order), and for every halo column its
 * global index, owner rank and index inside the owner's slab.  The output
 * arrays may be NULL; halo arrays need room for up to nnz entries. */
int psp_dist_plan(int world, int rank, const int *rowptr, int nl, const int *ind,
                  const int *col_global, int *col_local, int *halo_gcol, int *halo_owner,
                  int *halo_ridx, int *nhalo_out);

/* Allocate this rank's symmetric heap (nslots vectors of slot_elems doubles)
 * and control block; writes two CUDA IPC handles (128 bytes) to handles_out. */
int psp_dist_init(int rank, int world, size_t slot_elems, int nslots, void *handles_out);
/* all_handles: the world x 128-byte blobs of all ranks, in rank order */
int psp_dist_connect(const void *all_handles);
int psp_dist_shutdown(void);
int psp_dist_info(int *rank, int *world, size_t *slot_elems, int *nslots);

/* This rank's slab: ind[nl+1], col_global[nnz], val[nnz] */
int psp_dist_csr_create(const int *rowptr, int nnz, const int *ind, const int *col_global,
                        const double *val, psp_dmat **out);
/* Optional send map (push model of the single-launch PCG kernel): entry s says that
 * local row rows[s] must be stored at element dst_off[s] of the multiplied vector in
 * rank dst_rank[s]'s heap (dst_off = that rank's local row count + its halo position).
 * Without it the halo is pulled by the consumer.  The maps of all ranks together must
 * cover every halo entry of every rank exactly once (the receiver waits for each of its
 * halo entries); pysparse_b200/dist.py builds them by inverting the all-gathered halo
 * lists of psp_dist_plan.  Call after psp_dist_connect. */
int psp_dist_set_sendmap(psp_dmat *a, int nsend, const int *rows, const int *dst_rank,
                         const int *dst_off);
void psp_dist_mat_destroy(psp_dmat *a);
int psp_dist_mat_info(const psp_dmat *a, int *nl, int *nhalo, unsigned *neighbour_mask);

/* Sharded solve (PSP_PCG, PSP_BICGSTAB, PSP_GMRES; K = NULL or a single-step
 * Jacobi built from the local diagonal).  db / dx are this rank's slices,
 * device-resident.  Collective: every rank must call it with the same
 * arguments.  Heap slots needed: 4 (pcg), 8 (bicgstab), 2*dim+2 (gmres with
 * Jacobi) or dim+2 (gmres without preconditioner). */
int psp_dist_solve_dev(int solver, psp_dmat *A, psp_prec *K, const double *db, double *dx,
                       double tol, int maxit, int gmres_dim, int *info, int *iter, double *relres);

/* ---- tuning knobs (default to reference semantics) --------------------- */
/* name: "spmv_variant" (0 auto, 1 scalar-stream, 2 vector-stream ...),
 *       "poll_chunk" (iterations enqueued between convergence polls), ...
 * the full list is in DESIGN.md section 1; unknown names return PSP_ERR_ARG    */
int psp_set_option(const char *name, long long value);
long long psp_get_option(const char *name);

#ifdef __cplusplus
}
#endif
#endif /* PSP_B200_H */
