// fp32x2_tput.cu -- issue/pipe throughput of scalar vs packed fp32 on sm_100a (FFMA/FADD vs FFMA2/FADD2),
// register, immediate and swapped-operand forms.  Prints warp-instructions per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_tput fp32x2_tput.cu
#include <cstdio>
#include <cuda_runtime.h>
#define PK(d, a, b, c, OP) asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; " OP " rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}" \
   : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y))
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c){ float2 d; PK(d,a,b,c,"fma.rn.f32x2"); return d; }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b){
  float2 d;
  asm volatile("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
   : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
constexpr int CH = 8;      // independent chains per thread
constexpr int IT = 512;
template <int MODE>
__global__ void k(float2* p, long long* cyc, float s) {
  float2 v[CH];
  for (int i = 0; i < CH; ++i) v[i] = p[threadIdx.x + 32 * i];
  float2 w = p[1000 + threadIdx.x];
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) { v[i].x = fmaf(v[i].x, w.x, w.y); v[i].y = fmaf(v[i].y, w.x, w.y); }          // 2 FFMA reg
      if (MODE == 1) { v[i].x = fmaf(v[i].x, 0.999f, w.y); v[i].y = fmaf(v[i].y, 0.999f, w.y); }      // 2 FFMA imm
      if (MODE == 2) { v[i].x = v[i].x + w.x; v[i].y = v[i].y + w.y; }                                // 2 FADD
      if (MODE == 3) { v[i] = ffma2(v[i], w, w); }                                                    // FFMA2 reg
      if (MODE == 4) { v[i] = ffma2(v[i], make_float2(0.999f, 0.999f), w); }                          // FFMA2 imm
      if (MODE == 5) { v[i] = fadd2(v[i], w); }                                                       // FADD2
      if (MODE == 6) { v[i] = ffma2(make_float2(-w.y, w.x), make_float2(0.999f, 0.999f), v[i]); }     // FFMA2 swap.NP + imm
      if (MODE == 7) { v[i] = ffma2(make_float2(w.x, w.x), v[i], w); }                                // FFMA2 scalar-broadcast
      if (MODE == 8) { v[i] = fadd2(v[i], make_float2(-v[(i + 1) % CH].y, v[(i + 1) % CH].x)); }      // FADD2 with swapped variable operand
    }
  }
  long long t1 = clock64();
  float2 acc = v[0];
  for (int i = 1; i < CH; ++i) acc = fadd2(acc, v[i]);
  p[2000 + blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int per_iter_instr, float2* d, long long* dc) {
  for (int warps : {4, 8, 16, 32}) {
    k<MODE><<<148, warps * 32>>>(d, dc, 1.f);
    k<MODE><<<148, warps * 32>>>(d, dc, 1.f);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    double instr = (double)IT * CH * per_iter_instr * warps;
    printf("%-28s warps/SM=%2d  cycles=%8lld  warp-instr/clk/SM=%.3f  flop/clk/SM=%.1f\n", name, warps, c, instr / c,
           instr / c * 32 * (per_iter_instr == 2 ? 1 : 2) * (MODE == 2 || MODE == 5 || MODE == 8 ? 1 : 2));
  }
}
int main() {
  float2* d; long long* dc;
  cudaMalloc(&d, 1 << 24); cudaMemset(d, 0, 1 << 24); cudaMalloc(&dc, 8);
  run<0>("FFMA reg (x2)", 2, d, dc);
  run<1>("FFMA imm (x2)", 2, d, dc);
  run<2>("FADD (x2)", 2, d, dc);
  run<3>("FFMA2 reg", 1, d, dc);
  run<4>("FFMA2 imm", 1, d, dc);
  run<5>("FADD2", 1, d, dc);
  run<6>("FFMA2 swapNP+imm", 1, d, dc);
  run<7>("FFMA2 bcast", 1, d, dc);
  run<8>("FADD2 swapNP var", 1, d, dc);
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
