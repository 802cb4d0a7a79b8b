// micro-benchmark: cost per iteration of cluster barrier variants (B200)
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
__device__ __forceinline__ void arrive_rel() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void arrive_rlx() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_acq() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_rlx() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
template <int MODE>
__global__ void k(double *x, int iters, int n) {
  cg::cluster_group cl = cg::this_cluster();
  int t = cl.block_rank() * blockDim.x + threadIdx.x;
  int tot = cl.num_blocks() * blockDim.x;
  double acc = 0;
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) { arrive_rel(); wait_acq(); }                         // barrier only
    if (MODE == 1) { arrive_rlx(); wait_rlx(); }                         // relaxed barrier only
    if (MODE == 2) { x[(t + i) % n] = acc; arrive_rel(); wait_acq(); }   // store + release barrier
    if (MODE == 3) { acc += __ldcg(x + (t + 1 + i) % n); x[(t + i) % n] = acc; arrive_rel(); wait_acq(); } // gather+store
    if (MODE == 4) { __syncthreads(); }                                  // CTA barrier
    if (MODE == 5) { acc += __ldcg(x + (t + 1 + i) % n); x[(t + i) % n] = acc; __syncthreads(); }
    if (MODE == 6) { acc += __ldcg(x + (t + 1 + i) % n); x[(t + i) % n] = acc; __threadfence(); arrive_rlx(); wait_rlx(); }
    if (MODE == 7) { acc += __ldcg(x + (t + 1 + i) % n); x[(t + i) % n] = acc / 3.0; arrive_rel(); wait_acq(); }
  }
  if (acc == 12345.678) x[0] = acc;
}
template <int MODE> void run(const char *name, int csize, int threads) {
  double *x; int n = 1 << 20; cudaMalloc(&x, n * 8); cudaMemset(x, 0, n * 8);
  int iters = 4000;
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(csize); cfg.blockDim = dim3(threads);
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, k<MODE>, x, 10, n);
  cudaEventRecord(e0); cudaLaunchKernelEx(&cfg, k<MODE>, x, iters, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-34s cluster=%2d threads=%4d : %.3f us/iter  (%s)\n", name, csize, threads, ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(x);
}
int main() {
  for (int cs : {1, 2, 8, 16}) for (int th : {128, 512}) {
    run<0>("barrier.cluster rel/acq", cs, th); run<1>("barrier.cluster relaxed", cs, th);
    run<2>("store + rel/acq", cs, th); run<3>("ldcg + store + rel/acq", cs, th);
    run<6>("ldcg+store+threadfence+relaxed", cs, th); run<7>("ldcg + div + store + rel/acq", cs, th);
  }
  run<4>("__syncthreads", 1, 512); run<5>("ldcg + store + __syncthreads", 1, 512); run<5>("ldcg + store + __syncthreads", 1, 1024);
  return 0;
}
