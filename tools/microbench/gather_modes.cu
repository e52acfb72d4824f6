// gather_modes.cu -- can the k-space gather run at HBM speed WITHOUT the L1?  (round-2 follow-up of gather_ring.cu)
// The product column pass gathers sampled columns with 8-byte cp.async.ca, whose in-flight lines are staged in L1; with a
// large shared-memory CTA beside it (any co-resident / cluster design) L1 shrinks to 28 KB and the gather halves.  This
// probe measures, for a persistent CTA per SM whose dynamic shared memory is padded to `pad` KB:
//   mode ca8   : 8-byte  cp.async.ca of the sampled column                    (slot = G cols x 640 rows x 8 B)
//   mode cg16  : 16-byte cp.async.cg of the aligned pair holding the column   (slot = G cols x 640 rows x 16 B, L1 bypassed)
//   mode tma   : cp.async.bulk.tensor.2d of the whole band of BW raw columns  (slot = BW cols x 640 rows x 8 B, no L1, no LSU)
// One consumer warp per CTA waits for the slot, touches it, optionally burns `delay` cycles and hands it back.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_modes gather_modes.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int H = 640, W = 368;

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_cp_arrive(unsigned long long* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, int parity) {
  const unsigned a = s32(b);
  for (long long spin = 0; spin < (1ll << 24); ++spin) {      // bounded: a protocol bug must not hang the box
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
  }
  if ((threadIdx.x & 31) == 0) printf("mbar_wait timeout: block %d thread %d parity %d\n", blockIdx.x, threadIdx.x, parity);
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(dst)), "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}

struct Params {
  const float2* ksp; const int* act_w; int n_act, n_groups, n_frames, n_slots, delay; float* sink;
};

// MODE 0: ca8, MODE 1: cg16.  G sampled columns per item, P producer warps.
template <int MODE, int G>
__global__ void __launch_bounds__(256, 1) lsu_kernel(Params p, int n_prod) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int ES = MODE == 0 ? 8 : 16;       // bytes per staged element
  constexpr int PITCH = 722;
  __shared__ unsigned long long full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) for (int s = 0; s < p.n_slots; ++s) { mbar_init(&full[s], 32 * n_prod); mbar_init(&empty[s], 1); }
  __syncthreads();
  const int n_items = p.n_frames * p.n_groups;
  const int my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr int RPI = 32 / G;
  if (warp < n_prod) {
    // producers split the 640 rows of every item
    const int k = lane % G, hs = lane / G;
    const int rows_per = 640 / n_prod;
    for (int i = 0; i < my; ++i) {
      const int slot = i % p.n_slots, use = i / p.n_slots;
      if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
      const int item = blockIdx.x + i * gridDim.x;
      const int f = item / p.n_groups, g = item - f * p.n_groups;
      const int j = min(g * G + k, p.n_act - 1);
      const int w = p.act_w[j];
      const int h0 = warp * rows_per + hs;
      unsigned char* dst = smraw + ((size_t)slot * G * PITCH + k * PITCH + h0) * ES;
      if (MODE == 0) {
        const float2* src = p.ksp + ((size_t)f * H + h0) * W + w;
#pragma unroll 8
        for (int q = 0; q < rows_per / RPI; ++q) { cp8(dst + (size_t)q * RPI * ES, src); src += (size_t)RPI * W; }
      } else {
        const float2* src = p.ksp + ((size_t)f * H + h0) * W + (w & ~1);
#pragma unroll 8
        for (int q = 0; q < rows_per / RPI; ++q) { cp16(dst + (size_t)q * RPI * ES, src); src += (size_t)RPI * W; }
      }
      mbar_cp_arrive(&full[slot]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == n_prod) {
    float acc = 0.f;
    for (int i = 0; i < my; ++i) {
      const int slot = i % p.n_slots, use = i / p.n_slots;
      mbar_wait(&full[slot], use & 1);
      acc += reinterpret_cast<float*>(smraw + (size_t)slot * G * PITCH * ES)[lane];
      if (p.delay) { const long long t0 = clock64(); while (clock64() - t0 < p.delay) {} }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 123.456f) p.sink[0] = acc;
  }
}

// MODE tma: item = (frame, band of BW raw columns); ROWS rows per box, 640/ROWS boxes per item.
template <int BW, int ROWS>
__global__ void __launch_bounds__(64, 1) tma_kernel(const __grid_constant__ CUtensorMap map, Params p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ unsigned long long full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.n_slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr int NB = W / BW;                     // bands per frame
  constexpr unsigned SLOT_BYTES = BW * 8 * H;
  unsigned char* base = smraw + ((128 - (s32(smraw) & 127)) & 127);
  const int n_items = p.n_frames * NB;
  const int my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < my; ++i) {
        const int slot = i % p.n_slots, use = i / p.n_slots;
        if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
        const int item = blockIdx.x + i * gridDim.x;
        const int f = item / NB, b = item - f * NB;
        mbar_expect_tx(&full[slot], SLOT_BYTES);
#pragma unroll
        for (int q = 0; q < H / ROWS; ++q)
          tma_load_2d(base + (size_t)slot * SLOT_BYTES + (size_t)q * ROWS * BW * 8, &map, b * BW * 2, f * H + q * ROWS, &full[slot]);
      }
    }
  } else {
    float acc = 0.f;
    for (int i = 0; i < my; ++i) {
      const int slot = i % p.n_slots, use = i / p.n_slots;
      mbar_wait(&full[slot], use & 1);
      acc += reinterpret_cast<float*>(base + (size_t)slot * SLOT_BYTES)[lane];
      if (p.delay) { const long long t0 = clock64(); while (clock64() - t0 < p.delay) {} }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc == 123.456f) p.sink[0] = acc;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


template <typename F>
static void report(const char* what, F launch, size_t bytes) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) launch();
  cudaEventRecord(a);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
  printf("%s : %.3f ms  %.0f GB/s  (%s)\n", what, ms, (double)bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  fflush(stdout);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("device error, stopping\n"); exit(1); }
}

int main(int argc, char** argv) {
  const int frames = 960;
  const size_t bytes = (size_t)frames * H * W * 8;
  float2* d; float* sink;
  cudaMalloc(&d, bytes); cudaMalloc(&sink, 4); cudaMemset(d, 0, bytes);
  std::vector<int> act;
  for (int w = 0; w < W; ++w) if (w % 4 == 0 || (w >= 170 && w < 199)) act.push_back(w);
  int* dact; cudaMalloc(&dact, act.size() * 4); cudaMemcpy(dact, act.data(), act.size() * 4, cudaMemcpyHostToDevice);
  printf("active columns %zu, k-space %.2f GB\n", act.size(), bytes / 1e9);

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }

  char what[256];
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  const bool tma_only = argc > 2;
  printf("grid %d\n", grid);
  for (int pad : {0, 200}) {
    for (int delay : {0, 3000}) {
      // ---- LSU modes, 4-column items
      for (int mode = 0; mode < (tma_only ? 0 : 2); ++mode) {
        for (int n_prod : {1, 2}) {
          for (int slots : {2, 4}) {
            const int G = 4, ES = mode ? 16 : 8;
            const size_t ring = (size_t)slots * G * 722 * ES;
            const size_t smem = std::max(ring, (size_t)pad * 1024);
            if (smem > 220 * 1024) continue;
            Params p{d, dact, (int)act.size(), ((int)act.size() + G - 1) / G, frames, slots, delay, sink};
            snprintf(what, sizeof what, "%-5s G=4 producers=%d slots=%d ring %3zu KB CTA %3zu KB delay %4d", mode ? "cg16" : "ca8",
                     n_prod, slots, ring / 1024, smem / 1024, delay);
            if (mode == 0) {
              cudaFuncSetAttribute(lsu_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
              report(what, [&] { lsu_kernel<0, 4><<<grid, 32 * (n_prod + 1), smem>>>(p, n_prod); }, bytes);
            } else {
              cudaFuncSetAttribute(lsu_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
              report(what, [&] { lsu_kernel<1, 4><<<grid, 32 * (n_prod + 1), smem>>>(p, n_prod); }, bytes);
            }
          }
        }
      }
      // ---- TMA bands
      auto run_tma = [&](auto kern, int BW, int ROWS, int slots) {
        CUtensorMap map;
        const cuuint64_t dims[2] = {(cuuint64_t)W * 2, (cuuint64_t)H * frames};
        const cuuint64_t strides[1] = {(cuuint64_t)W * 8};
        const cuuint32_t box[2] = {(cuuint32_t)BW * 2, (cuuint32_t)ROWS};
        const cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
        const size_t ring = (size_t)slots * BW * 8 * H + 128;
        const size_t smem = std::max(ring, (size_t)pad * 1024);
        if (smem > 220 * 1024) return;
        Params p{d, dact, (int)act.size(), 0, frames, slots, delay, sink};
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        snprintf(what, sizeof what, "tma   band=%2d cols box %3d rows slots=%d ring %3zu KB CTA %3zu KB delay %4d", BW, ROWS, slots,
                 ring / 1024, smem / 1024, delay);
        report(what, [&] { kern<<<grid, 64, smem>>>(map, p); }, bytes);
      };
      run_tma(tma_kernel<16, 128>, 16, 128, 2);
      run_tma(tma_kernel<8, 128>, 8, 128, 2);
      run_tma(tma_kernel<8, 128>, 8, 128, 3);
      run_tma(tma_kernel<8, 128>, 8, 128, 4);
      run_tma(tma_kernel<4, 128>, 4, 128, 4);
      run_tma(tma_kernel<4, 128>, 4, 128, 8);
      run_tma(tma_kernel<8, 64>, 8, 64, 4);
      run_tma(tma_kernel<16, 64>, 16, 64, 2);
    }
  }
  return 0;
}
