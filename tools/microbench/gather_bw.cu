// gather_bw.cu -- what HBM bandwidth can the column-pass access pattern reach at all?
// Reads a [frames][640][368] complex64 buffer (1.81 GB for 960 frames) with different traversals and reports
// GB/s of k-space covered.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bw gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int H = 640, W = 368;

__device__ __forceinline__ float2 ld8(const float2* p) {
  float2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v;
}
__device__ __forceinline__ float4 ld16(const float4* p) {
  float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v;
}

// mode 0: every 4th complex of a 32-column strip (8-byte loads), lanes = 8 cols x 4 rows (the column pass)
// mode 1: dense strip of SW columns (16-byte loads), lanes sweep the strip row by row
// mode 2: dense full rows, RB rows per item
template <int MODE, int SW>
__global__ void __launch_bounds__(256) k_gather(const float2* __restrict__ ksp, int n_frames, float* sink) {
  const int tid = threadIdx.x;
  float acc = 0.f;
  if (MODE == 0) {
    const int groups = W / 32;                     // 11 full strips (the tail is ignored here)
    const int items = n_frames * groups;
    const int k = tid & 7, hs = tid >> 3;          // 32 rows per sweep with 256 threads
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int f = item / groups, g = item % groups;
      const float2* src = ksp + (size_t)f * H * W + (size_t)hs * W + g * 32 + 4 * k;
#pragma unroll 1
      for (int it = 0; it < H / 32; it += 10) {
        float2 v[10];
#pragma unroll
        for (int u = 0; u < 10; ++u) v[u] = ld8(src + (size_t)(it + u) * 32 * W);
#pragma unroll
        for (int u = 0; u < 10; ++u) acc += v[u].x + v[u].y;
      }
    }
  } else if (MODE == 1) {
    const int strips = W / SW;
    const int items = n_frames * strips;
    constexpr int TPR = SW / 2;                    // threads per row (16 B each)
    constexpr int RPS = 256 / TPR;                 // rows per sweep
    const int c2 = tid % TPR, hs = tid / TPR;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int f = item / strips, g = item % strips;
      const float4* src = reinterpret_cast<const float4*>(ksp + (size_t)f * H * W + (size_t)hs * W + g * SW) + c2;
      constexpr int SWEEPS = H / RPS;
#pragma unroll 1
      for (int it = 0; it < SWEEPS; it += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) if (it + u < SWEEPS) v[u] = ld16(src + (size_t)(it + u) * RPS * (W / 2)); else v[u] = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
      }
    }
  } else {
    // contiguous: 46 KB chunks (20 rows)
    constexpr int CH = 20 * W / 2;                 // float4 per item
    const size_t total4 = (size_t)n_frames * H * W / 2;
    const size_t items = total4 / CH;
    for (size_t item = blockIdx.x; item < items; item += gridDim.x) {
      const float4* src = reinterpret_cast<const float4*>(ksp) + item * CH;
#pragma unroll 1
      for (int i = tid; i < CH; i += 256 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + u * 256 < CH) ? ld16(src + i + u * 256) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].w;
      }
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

template <int MODE, int SW> void run(const char* name, const float2* d, int frames, float* sink, int grid, double covered_bytes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) k_gather<MODE, SW><<<grid, 256>>>(d, frames, sink);
  cudaEventRecord(a);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) k_gather<MODE, SW><<<grid, 256>>>(d, frames, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
  printf("%-28s grid %5d  %.3f ms  %.0f GB/s  (%s)\n", name, grid, ms, covered_bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int frames = 960;
  const size_t bytes = (size_t)frames * H * W * 8;
  float2* d; float* sink;
  cudaMalloc(&d, bytes); cudaMalloc(&sink, 4); cudaMemset(d, 0, bytes);
  const double strip_cov = (double)frames * H * (W / 32 * 32) * 8;    // 352 of 368 columns
  for (int per_sm : {2, 4, 8}) {
    const int grid = 148 * per_sm;
    run<0, 32>("gather 8B every 4th, 256B strip", d, frames, sink, grid, strip_cov);
    run<1, 32>("dense 256B strip", d, frames, sink, grid, strip_cov);
    run<1, 64>("dense 512B strip", d, frames, sink, grid, (double)frames * H * (W / 64 * 64) * 8);
    run<1, 128>("dense 1KB strip", d, frames, sink, grid, (double)frames * H * (W / 128 * 128) * 8);
    run<2, 32>("contiguous 46KB chunks", d, frames, sink, grid, (double)bytes);
  }
  return 0;
}
