// gather_ring.cu -- what does the k-space gather of a persistent, warp-specialised CTA sustain per SM?
// Models the producer side of the fused cluster kernel: P producer warps per CTA stream (frame, column group) items with
// 8-byte cp.async into a ring of S shared-memory slots (G columns x 640 rows each), completion is signalled by an
// mbarrier the copies complete themselves; one consumer warp per producer waits for the slot, optionally burns
// `delay` cycles (the transform), and hands the slot back.  Reports GB/s of k-space covered for grids of 128 and
// 148 CTAs (one CTA per SM) so that the cluster kernel (16 clusters x 8 = 128 SMs) can be sized before it exists.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_ring gather_ring.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int H = 640, W = 368;
constexpr int PITCH = 722;

__device__ __forceinline__ void cp8(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_cp_arrive(unsigned long long* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, int parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(b);
  for (long long spin = 0; spin < (1ll << 24); ++spin) {      // bounded: a protocol bug must not hang the box
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
  }
  if ((threadIdx.x & 31) == 0) printf("mbar_wait timeout: block %d thread %d parity %d\n", blockIdx.x, threadIdx.x, parity);
  __trap();
}

struct Params {
  const float2* ksp; const int* act_w; int n_act, n_groups, n_frames, n_slots, n_prod, delay; float* sink;
};

template <int G>
__global__ void __launch_bounds__(512, 1) ring_kernel(Params p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* sm = reinterpret_cast<float2*>(smraw);
  __shared__ unsigned long long full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) for (int s = 0; s < p.n_slots; ++s) { mbar_init(&full[s], 32); mbar_init(&empty[s], 1); }
  __syncthreads();
  const int n_items = p.n_frames * p.n_groups;
  const int my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items of this CTA
  constexpr int RPI = 32 / G;              // rows per instruction
  if (warp < p.n_prod) {
    const int k = lane % G, hs = lane / G;
    for (int i = warp; i < my; i += p.n_prod) {
      const int slot = i % p.n_slots, use = i / p.n_slots;
      if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
      const int item = blockIdx.x + i * gridDim.x;
      const int f = item / p.n_groups, g = item - f * p.n_groups;
      const int j0 = g * G;
      if (j0 + k < p.n_act) {
        const float2* src = p.ksp + ((size_t)f * H + hs) * W + p.act_w[j0 + k];
        float2* dst = sm + (size_t)slot * G * PITCH + k * PITCH + hs;
#pragma unroll 1
        for (int blk = 0; blk < 8; ++blk) {
#pragma unroll
          for (int q = 0; q < 80 / RPI; ++q) { cp8(dst + blk * 90 + q * RPI, src); src += (size_t)RPI * W; }
        }
      }
      mbar_cp_arrive(&full[slot]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp < 2 * p.n_prod) {
    const int w = warp - p.n_prod;
    float acc = 0.f;
    for (int i = w; i < my; i += p.n_prod) {
      const int slot = i % p.n_slots, use = i / p.n_slots;
      mbar_wait(&full[slot], use & 1);
      acc += sm[(size_t)slot * G * PITCH + lane].x;
      if (p.delay) { const long long t0 = clock64(); while (clock64() - t0 < p.delay) {} }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 123.456f) p.sink[0] = acc;
  }
}


// ---- variant: register-staged gather.  L loader warps share every item: each lane issues batches of 8-byte loads that
// bypass L1 (ld.global.cg / .nc.L1::no_allocate) into registers, two batches in flight, and stores them with STS.64.
// The in-flight data lives in registers, not in L1 lines, so the gather does not depend on the L1 carve-out.
template <int OP> __device__ __forceinline__ float2 ldg8(const float2* p) {
  float2 v;
  if (OP == 0) asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  else if (OP == 1) asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  else asm volatile("ld.global.ca.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

template <int OP, int L, int B>
__global__ void __launch_bounds__(512, 1) ring_ldg_kernel(Params p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  float2* sm = reinterpret_cast<float2*>(smraw);
  __shared__ unsigned long long full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) for (int s = 0; s < p.n_slots; ++s) { mbar_init(&full[s], L * 32); mbar_init(&empty[s], 1); }
  __syncthreads();
  constexpr int G = 8;
  const int n_items = p.n_frames * p.n_groups;
  const int my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr int IPW = 160 / L;             // instructions (4-row sweeps) per warp per item
  constexpr int NB = IPW / B;              // batches per item
  static_assert(IPW % B == 0 && NB >= 1, "batching");
  if (warp < L) {
    const int k = lane & 7, hs = lane >> 3;
    float2 r[2][B];
    auto src_of = [&](int i) {
      const int item = blockIdx.x + i * gridDim.x;
      const int f = item / p.n_groups, g = item - f * p.n_groups;
      const int j = min(g * G + k, p.n_act - 1);
      return p.ksp + ((size_t)f * H + hs + 4 * IPW * warp) * W + p.act_w[j];
    };
    auto load = [&](const float2* src, int b, int which) {
#pragma unroll
      for (int u = 0; u < B; ++u) r[which][u] = ldg8<OP>(src + (size_t)(b * B + u) * 4 * W);
    };
    if (my > 0) load(src_of(0), 0, 0);
    int n = 0;                               // running batch counter
    for (int i = 0; i < my; ++i) {
      const int slot = i % p.n_slots, use = i / p.n_slots;
      const float2* src = src_of(i);
      float2* dst = sm + (size_t)slot * G * PITCH + k * PITCH + hs + 4 * IPW * warp;
#pragma unroll
      for (int b = 0; b < NB; ++b, ++n) {
        // issue the next batch (possibly of the next item) before storing this one
        if (b + 1 < NB) load(src, b + 1, (n + 1) & 1);
        else if (i + 1 < my) load(src_of(i + 1), 0, (n + 1) & 1);
        if (b == 0 && use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
#pragma unroll
        for (int u = 0; u < B; ++u) dst[(b * B + u) * 4] = r[n & 1][u];
      }
      mbar_arrive(&full[slot]);
    }
  } else if (warp == L) {
    float acc = 0.f;
    for (int i = 0; i < my; ++i) {
      const int slot = i % p.n_slots, use = i / p.n_slots;
      mbar_wait(&full[slot], use & 1);
      acc += sm[(size_t)slot * G * PITCH + lane].x;
      if (p.delay) { const long long t0 = clock64(); while (clock64() - t0 < p.delay) {} }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 123.456f) p.sink[0] = acc;
  }
}

template <int OP, int L, int B>
void run_ldg(const char* name, Params p, int grid, size_t smem, size_t bytes) {
  cudaFuncSetAttribute(ring_ldg_kernel<OP, L, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) ring_ldg_kernel<OP, L, B><<<grid, 32 * (L + 1), smem>>>(p);
  cudaEventRecord(a);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) ring_ldg_kernel<OP, L, B><<<grid, 32 * (L + 1), smem>>>(p);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
  printf("grid %3d  LDG %-10s loaders=%d batch=%d slots=%d (CTA %3zu KB) delay=%5d : %.3f ms  %.0f GB/s  (%s)\n", grid, name, L, B,
         p.n_slots, smem / 1024, p.delay, ms, (double)bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  fflush(stdout);
}

int main(int argc, char** argv) {
  const int frames = 960;
  const size_t bytes = (size_t)frames * H * W * 8;
  float2* d; float* sink;
  cudaMalloc(&d, bytes); cudaMalloc(&sink, 4); cudaMemset(d, 0, bytes);
  // knee mask: every 4th column + 29 ACS columns from 170
  std::vector<int> act;
  for (int w = 0; w < W; ++w) if (w % 4 == 0 || (w >= 170 && w < 199)) act.push_back(w);
  int* dact; cudaMalloc(&dact, act.size() * 4); cudaMemcpy(dact, act.data(), act.size() * 4, cudaMemcpyHostToDevice);
  printf("active columns %zu\n", act.size());
  cudaFuncSetAttribute(ring_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(ring_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  struct Cfg { int G, prod, slots, delay, pad_kb; };
  // pad_kb: total dynamic shared memory of the CTA (the rest of the fused kernel's buffers): what is left of the 256 KB
  // unified array is L1, which the in-flight .ca copies need
  const Cfg cfgs[] = {{8, 1, 2, 0, 0}, {8, 1, 3, 0, 0}, {8, 1, 3, 0, 160}, {8, 2, 2, 0, 160}, {8, 1, 2, 0, 190}, {8, 2, 2, 0, 190},
                      {8, 2, 4, 0, 190}, {8, 1, 3, 4000, 160}, {8, 2, 2, 4000, 160}, {8, 2, 2, 4000, 190}};
  for (int grid : {128, 148}) {
    for (const Cfg& c : cfgs) {
      Params p{d, dact, (int)act.size(), ((int)act.size() + c.G - 1) / c.G, frames, c.slots, c.prod, c.delay, sink};
      const size_t smem = std::max((size_t)c.slots * c.G * PITCH * 8, (size_t)c.pad_kb * 1024);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      auto launch = [&] {
        if (c.G == 8) ring_kernel<8><<<grid, 64 * c.prod, smem>>>(p); else ring_kernel<4><<<grid, 64 * c.prod, smem>>>(p);
      };
      for (int i = 0; i < 2; ++i) launch();
      cudaEventRecord(a);
      const int reps = 5;
      for (int i = 0; i < reps; ++i) launch();
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
      printf("grid %3d  G=%d producers=%d slots=%d (ring %3zu KB, CTA %3zu KB) delay=%5d : %.3f ms  %.0f GB/s  (%s)\n", grid, c.G, c.prod, c.slots,
             (size_t)c.slots * c.G * PITCH * 8 / 1024, smem / 1024, c.delay, ms, (double)bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      fflush(stdout);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("device error, stopping\n"); return 1; }
    }
    for (int slots : {2, 3}) {
      for (int pad : {0, 220}) {
        for (int delay : {0, 4000}) {
          Params p{d, dact, (int)act.size(), ((int)act.size() + 7) / 8, frames, slots, 0, delay, sink};
          const size_t smem = std::max((size_t)slots * 8 * PITCH * 8, (size_t)pad * 1024);
          run_ldg<0, 4, 20>("cg", p, grid, smem, bytes);
          run_ldg<1, 4, 20>("nc.noalloc", p, grid, smem, bytes);
          if (delay == 0) {
            run_ldg<2, 4, 20>("ca", p, grid, smem, bytes);
            run_ldg<0, 2, 20>("cg", p, grid, smem, bytes);
            run_ldg<0, 8, 10>("cg", p, grid, smem, bytes);
            run_ldg<0, 8, 20>("cg", p, grid, smem, bytes);
            run_ldg<0, 4, 40>("cg", p, grid, smem, bytes);
          }
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("device error, stopping\n"); return 1; }
        }
      }
    }
  }
  return 0;
}
