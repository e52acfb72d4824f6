// l1_bw.cu -- does the L1/shared carveout limit the gather bandwidth?  The column-pass gather (8 B of every
// 32 B sector) through (a) register loads and (b) LDGSTS.ca into shared memory, with 0 / 100 / 200 KB of dynamic
// shared memory per SM taken away from L1.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int H = 640, W = 368;
__device__ __forceinline__ float2 ld8(const float2* p) {
  float2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
}
__device__ __forceinline__ void cpa8(float2* s, const float2* g) {
  unsigned d = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(g) : "memory");
}
template <int MODE>   // 0: register loads, 1: LDGSTS.ca 8 B, one producer warp per CTA style (all threads issue)
__global__ void __launch_bounds__(256) k(const float2* __restrict__ ksp, int n_frames, float* sink) {
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x, groups = W / 32, items = n_frames * groups;
  const int kk = tid & 7, hs = tid >> 3;
  float acc = 0.f;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int f = item / groups, g = item % groups;
    const float2* src = ksp + (size_t)f * H * W + (size_t)hs * W + g * 32 + 4 * kk;
    if (MODE == 0) {
#pragma unroll 1
      for (int it = 0; it < 20; it += 10) {
        float2 v[10];
#pragma unroll
        for (int u = 0; u < 10; ++u) v[u] = ld8(src + (size_t)(it + u) * 32 * W);
#pragma unroll
        for (int u = 0; u < 10; ++u) acc += v[u].x + v[u].y;
      }
    } else {
#pragma unroll
      for (int it = 0; it < 20; ++it) cpa8(sm + it * 256 + tid, src + (size_t)it * 32 * W);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      acc += sm[tid].x;
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}
template <int MODE> void run(const char* name, const float2* d, int frames, float* sink, int per_sm, int smem_kb_per_cta) {
  const int smem = smem_kb_per_cta * 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * per_sm;
  for (int i = 0; i < 2; ++i) k<MODE><<<grid, 256, smem>>>(d, frames, sink);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k<MODE><<<grid, 256, smem>>>(d, frames, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
  const double cov = (double)frames * H * (W / 32 * 32) * 8;
  printf("%-22s %d CTA/SM x %3d KB smem: %.3f ms  %5.0f GB/s (%s)\n", name, per_sm, smem_kb_per_cta, ms, cov / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
template <int MODE> void run2(const float2* d, int frames, float* sink, int per_sm, int smem_kb_per_cta, int carve) {
  const int smem = smem_kb_per_cta * 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 0 ? smem : 1024);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * per_sm;
  for (int i = 0; i < 2; ++i) k<MODE><<<grid, 256, smem>>>(d, frames, sink);
  cudaEventRecord(a);
  for (int i = 0; i < 4; ++i) k<MODE><<<grid, 256, smem>>>(d, frames, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 4;
  const double cov = (double)frames * H * (W / 32 * 32) * 8;
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, 256, smem);
  printf("mode %d grid/SM %d smem/CTA %3d KB carve %3d occ %d : %5.0f GB/s\n", MODE, per_sm, smem_kb_per_cta, carve, occ, cov / ms / 1e6);
}
int main() {
  const int frames = 960;
  float2* d; float* sink;
  cudaMalloc(&d, (size_t)frames * H * W * 8); cudaMalloc(&sink, 4); cudaMemset(d, 0, (size_t)frames * H * W * 8);
  for (int carve : {-1, 0, 25, 50, 75, 100})
    for (int per : {1, 2, 4, 8})
      for (int kb : {0, 20, 46, 55}) {
        if (kb * per > 224) continue;
        run2<0>(d, frames, sink, per, kb, carve);
      }
  for (int carve : {-1, 50, 100}) for (int per : {2, 4}) for (int kb : {20, 46, 55}) run2<1>(d, frames, sink, per, kb, carve);
  return 0;
}
