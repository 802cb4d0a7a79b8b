// psp_dist.cuh -- host-side structures of the multi-GPU layer.
#pragma once
#include "psp_mat.cuh"

namespace psp {

struct DistHost {
  bool ready = false;
  int rank = 0, world = 1;
  size_t slot_elems = 0;
  int nslots = 0;
  double *heap = nullptr;                 // local symmetric heap: nslots x slot_elems doubles
  DistCtl *ctl = nullptr;                 // local control block (publication counter + mailboxes)
  double *peer_heap[kMaxRanks] = {};
  DistCtl *peer_ctl[kMaxRanks] = {};
  DistDev *dev = nullptr;                 // device copy of the descriptor
  bool in_process = false;                // peers are threads of this process (peer access, no IPC mappings)
};

struct DistMat {
  psp_mat *local = nullptr;               // nl x (nl + nhalo) CSR, halo columns appended
  int nl = 0, nhalo = 0, nglobal = 0;
  int *d_owner = nullptr, *d_ridx = nullptr;
  unsigned nbmask = 0;                    // ranks this rank pulls halo values from
  // push model (persistent kernels): local rows other ranks need, and where they go
  int nsend = 0;
  int *d_send_row = nullptr;
  unsigned long long **d_send_dst = nullptr;   // landing-zone address of each entry in the neighbour's heap
  bool has_sendmap = false;
};

DistHost &dist();
int dist_pull(DistMat *dm, double *xloc, int gate = 0);
int dist_local_alloc(int rank, int world, size_t slot_elems, int nslots);
int dist_local_connect(DistHost *const *all);
int dist_local_free();
int dist_csr_build(int world, int rank, const int *rowptr, int nnz, const int *ind, const int *col_global,
                   const double *val, DistMat **out, std::vector<int> *owner_out, std::vector<int> *ridx_out);
int dist_solve_dev(int solver, DistMat *dm, psp_prec *K, const double *db, double *dx, double tol,
                   int maxit, int gmres_dim, int *info, int *iter, double *relres);

}  // namespace psp
