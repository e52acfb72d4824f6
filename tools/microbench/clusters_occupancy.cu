// cluster_probe.cu -- how many thread-block clusters of size 2/4/8/16 are co-resident on this GPU for a CTA of the
// fused cluster kernel's shape (576 threads, ~220 KB dynamic shared memory), and which SMs they land on.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void probe_kernel(int* smid, int* rank) {
  extern __shared__ unsigned char sm[];
  cg::cluster_group cl = cg::this_cluster();
  if (threadIdx.x == 0) {
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    smid[blockIdx.x] = (int)id;
    rank[blockIdx.x] = (int)cl.block_rank();
    sm[0] = 1;
  }
  cl.sync();
}

int main() {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  printf("%s SMs=%d smem/block optin=%zu\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerBlockOptin);
  const int smems[] = {220 * 1024, 110 * 1024, 64 * 1024};
  const int threads[] = {576, 384, 256};
  for (int si = 0; si < 3; ++si) {
    const int smem = smems[si], nt = threads[si];
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs = 1; cs <= 16; cs *= 2) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
      printf("smem=%dKB threads=%d cluster=%2d: max active clusters=%d (%d SMs) %s\n", smem / 1024, nt, cs, n, n * cs,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
      if (e == cudaSuccess && n > 0 && si == 0 && cs == 8) {
        int *smid, *rank; cudaMalloc(&smid, 4 * n * cs); cudaMalloc(&rank, 4 * n * cs);
        cfg.gridDim = dim3(n * cs);
        e = cudaLaunchKernelEx(&cfg, probe_kernel, smid, rank);
        cudaError_t e2 = cudaDeviceSynchronize();
        printf("  launch %d CTAs: %s / %s\n", n * cs, cudaGetErrorString(e), cudaGetErrorString(e2));
        int h[4096];
        cudaMemcpy(h, smid, 4 * n * cs, cudaMemcpyDeviceToHost);
        for (int c = 0; c < n; ++c) { printf("  cluster %2d: SMs", c); for (int r = 0; r < cs; ++r) printf(" %3d", h[c * cs + r]); printf("\n"); }
      }
    }
  }
  return 0;
}
