// psp_multi.cuh -- several GPUs behind the single-process API (psp_multi.cu).
#pragma once
#include <vector>
#include "psp_common.cuh"

struct psp_mat;
struct psp_prec;

namespace psp {

struct DistMat;

struct MultiMat {                         // the row slabs of one matrix, slab r on device r
  int world = 0;
  std::vector<int> rowptr;                // rows [rowptr[r], rowptr[r+1]) belong to rank r
  std::vector<DistMat *> slab;
  std::vector<std::vector<int>> send_rows, send_rank, send_off;   // push-model send maps (host copies)
  int max_slot = 0;                       // largest local rows + halo over the ranks
  unsigned long long heap_gen = 0;        // generation of the symmetric heaps the send maps are bound to
};

int multi_gpus();                         // the "gpus" option of the calling context
int multi_attach_csr(psp_mat *A, const int *ind, const int *col, const double *val);
void multi_detach(MultiMat *mm);
void multi_prec_release(psp_prec *K);
bool multi_eligible(int solver, const psp_mat *A, const psp_prec *K, bool K_present);
// b, x: host arrays, or (on_device) vectors on the main context's device
int multi_solve(int solver, psp_mat *A, psp_prec *K, const double *b, double *x, bool on_device, double tol, int maxit,
                int gmres_dim, int *info, int *iter, double *relres);
int jacobi_from_dinv(int n, const double *dinv, psp_prec **out);    // psp_precon.cu

}  // namespace psp
