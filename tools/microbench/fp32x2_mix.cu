// fp32x2_mix.cu -- does a packed FFMA2 leave an issue slot free for another pipe?  Times (cudaEvent, whole grid)
// loops of FP instructions mixed with integer (alu pipe) and shared-memory (lsu) instructions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_mix fp32x2_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int IT = 4096;
// MODE 0: 16 FFMA(imm)            1: 8 FFMA2(imm)
//      2: 16 FFMA + 8 LOP3/IADD   3: 8 FFMA2 + 8 LOP3/IADD
//      4: 16 FFMA + 8 LDS.64      5: 8 FFMA2 + 8 LDS.64
//      6: 8 LOP3/IADD only        7: 8 LDS.64 only
//      8: 16 FFMA + 8 ALU + 8 LDS 9: 8 FFMA2 + 8 ALU + 8 LDS
template <int MODE>
__global__ void __launch_bounds__(512) k(float2* p, int* q) {
  __shared__ float2 sm[1024];
  float2 v[8];
  unsigned u[8];
  for (int i = 0; i < 8; ++i) { v[i] = p[threadIdx.x + 32 * i]; u[i] = q[threadIdx.x + 32 * i]; }
  sm[threadIdx.x] = v[0]; sm[threadIdx.x + 512] = v[1];
  __syncthreads();
  const float2 w = p[4000 + threadIdx.x];
  const float2* sp = sm + (threadIdx.x & 31);
  float2 ld = make_float2(0.f, 0.f);
  constexpr bool FP1 = MODE == 0 || MODE == 2 || MODE == 4 || MODE == 8;
  constexpr bool FP2 = MODE == 1 || MODE == 3 || MODE == 5 || MODE == 9;
  constexpr bool ALU = MODE == 2 || MODE == 3 || MODE == 6 || MODE == 8 || MODE == 9;
  constexpr bool LDS = MODE == 4 || MODE == 5 || MODE == 7 || MODE == 8 || MODE == 9;
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (FP1) { v[i].x = fmaf(v[i].x, 0.999f, w.x); v[i].y = fmaf(v[i].y, 0.999f, w.y); }
      if (FP2) { v[i] = __ffma2_rn(v[i], make_float2(0.999f, 0.999f), w); }
      if (ALU) { u[i] = (u[i] ^ (unsigned)it) + 0x9e3779b9u; }                   // LOP3 + IADD (2 alu instr)... counted as 2
      if (LDS) { float2 t = sp[32 * ((i + it) & 31)]; ld.x += t.x; }              // LDS.64 + FADD
    }
  }
  float2 acc = ld;
  unsigned uu = 0;
  for (int i = 0; i < 8; ++i) { acc.x += v[i].x; acc.y += v[i].y; uu += u[i]; }
  p[8000 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
  q[8000 + blockIdx.x * blockDim.x + threadIdx.x] = uu;
}
template <int MODE> void run(const char* name, float2* d, int* q) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {8, 16}) {
    k<MODE><<<148, warps * 32>>>(d, q);
    cudaEventRecord(e0);
    k<MODE><<<148, warps * 32>>>(d, q);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per loop iteration per SMSP-warp at 1.965 GHz (assumes max clock)
    double cyc = ms * 1e-3 * 1.965e9 / IT / (warps / 4.0);
    printf("%-34s warps/SM=%2d  %.3f ms  cycles/iter/warp(SMSP-serial)=%.2f\n", name, warps, ms, cyc);
  }
}
int main() {
  float2* d; int* q;
  cudaMalloc(&d, 1 << 24); cudaMemset(d, 0, 1 << 24); cudaMalloc(&q, 1 << 24); cudaMemset(q, 0, 1 << 24);
  run<0>("16 FFMA", d, q);
  run<1>("8 FFMA2", d, q);
  run<6>("8x(LOP3+IADD)", d, q);
  run<7>("8x(LDS.64+FADD)", d, q);
  run<2>("16 FFMA + 8x(LOP3+IADD)", d, q);
  run<3>("8 FFMA2 + 8x(LOP3+IADD)", d, q);
  run<4>("16 FFMA + 8x(LDS.64+FADD)", d, q);
  run<5>("8 FFMA2 + 8x(LDS.64+FADD)", d, q);
  run<8>("16 FFMA + 8xALU2 + 8xLDS", d, q);
  run<9>("8 FFMA2 + 8xALU2 + 8xLDS", d, q);
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
