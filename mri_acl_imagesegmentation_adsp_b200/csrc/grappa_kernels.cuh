// grappa_kernels.cuh -- GRAPPA weight application (ZIP!/fastmri_prostate/reconstruction/grappa.py:173-222): for every
// hole (an unsampled k-space position) of kernel geometry g, recon[x, y, :] = W_g @ S, where S collects the sampled
// neighbours of the 5 x 5 window around the hole over all coils (window position major, coil minor) and W_g is that
// geometry's (n_coils x n_sources) weight matrix of the slice; the result is recon + kspace.
//
// This is the one step of the prostate chain that IS a dense contraction (per geometry: [holes x K] times [K x coils],
// K = sampled window positions x coils ~ 160-400), compute-bound in fp32: ~26 kFLOP per hole, ~4.5 GFLOP per
// (average, slice) of the 640 x 451 x 16 files.  It stays on the fp32 pipe (complex64 in, complex64 out, rel-L2 1e-5
// against numpy's complex64 matmul rules out TF32) and is organised so that the FMA pipe is the limiter:
//   * one CTA item = (slice, geometry, 256 consecutive holes of that geometry); the slice's W_g is staged once in
//     shared memory, transposed to [k][16 outputs] so that a thread reads its 16 weights of one source as four
//     128-bit broadcast loads;
//   * every thread owns TWO holes (tid and tid + 128 of the item) and all 16 output coils of both: 64 accumulator
//     registers, each weight load feeds two complex FMAs, each source load sixteen (FFMA2 : LDS = 8 : 1);
//   * source values are gathered straight from global memory through L1: the host plan orders the holes of a geometry so
//     that the 32 lanes of a warp are neighbours along the axis with the SMALLER memory stride -- a warp's load then falls
//     into ~4 cache lines instead of 32 (measured: 3.3 ms -> 1.9 ms; with lanes along the strided axis the L1 hit rate is
//     2 % and 67 % of the stall samples are the scoreboard of these loads).
// Coils beyond 16 outputs run as further output groups (grid.z); any coil count works for the sources.
#pragma once
#include "common.cuh"

namespace mriacl {

constexpr int GR_T = 128;         // threads per CTA
constexpr int GR_HPT = 2;         // holes per thread
constexpr int GR_OUT = 16;        // output coils per CTA pass

struct GrappaParams {
  const cf* ksp;                  // element (slice, x, y, c) at slice*ss + x*sx + y*sy + c*sc (complex elements)
  cf* out;                        // same addressing; must already hold a copy of ksp (non-holes keep their value)
  long long ss, sx, sy, sc;
  int nc;
  const int* hole_xy;             // [n_holes] x * Y + y, grouped by geometry, ascending inside a group
  int Y;
  const int* item_geom;           // [n_items] geometry of item i
  const int* item_first;          // [n_items] first hole of item i
  const int* item_count;          // [n_items] holes of item i (<= GR_T * GR_HPT)
  const int* geom_src_start;      // [n_geom + 1] into src_off
  const int* src_off;             // [sum n_s] window offsets (di + 2) * 8 + (dj + 2)  (di, dj in -2..2 for a 5 x 5 kernel)
  const long long* geom_w_start;  // [n_geom] offset of W_g inside one slice's weight block (complex elements)
  const cf* weights;              // [n_slices][w_per_slice]: W_g as (nc, n_s * nc) row-major, the reference's layout
  long long w_per_slice;
  int kx2, ky2;
};

__global__ void __launch_bounds__(GR_T) grappa_apply_kernel(GrappaParams p) {
  MRIACL_DYN_SMEM(cf, wsm);                      // [K][GR_OUT]
  __shared__ int s_off[64];                      // window offsets of this geometry (<= 49 for kernels up to 7 x 7)
  const int item = blockIdx.x, slice = blockIdx.y, og = blockIdx.z;
  const int g = p.item_geom[item];
  const int s0 = p.geom_src_start[g], n_s = p.geom_src_start[g + 1] - s0;
  const int K = n_s * p.nc;
  const int tid = threadIdx.x;
  const int o0 = og * GR_OUT;
  const int n_out = min(GR_OUT, p.nc - o0);

  // stage W_g[o0 .. o0+16)[:] transposed: wsm[k][o]; outputs beyond n_out are zero
  const cf* W = p.weights + (long long)slice * p.w_per_slice + p.geom_w_start[g];
  for (int i = tid; i < K * GR_OUT; i += GR_T) {
    const int o = i / K, k = i - o * K;          // consecutive threads read consecutive k of one output row
    wsm[k * GR_OUT + o] = o < n_out ? W[(long long)(o0 + o) * K + k] : cf_make(0.f, 0.f);
  }
  if (tid < n_s) s_off[tid] = p.src_off[s0 + tid];
  __syncthreads();

  const int first = p.item_first[item], count = p.item_count[item];
  const cf* base = p.ksp + (long long)slice * p.ss;
  long long pos[GR_HPT];
  bool live[GR_HPT];
#pragma unroll
  for (int h = 0; h < GR_HPT; ++h) {
    const int idx = tid + h * GR_T;
    live[h] = idx < count;
    const int xy = p.hole_xy[first + (live[h] ? idx : 0)];
    const int x = xy / p.Y, y = xy - x * p.Y;
    pos[h] = (long long)x * p.sx + (long long)y * p.sy;
  }
  cf acc[GR_HPT][GR_OUT];
#pragma unroll
  for (int h = 0; h < GR_HPT; ++h)
#pragma unroll
    for (int o = 0; o < GR_OUT; ++o) acc[h][o] = cf_make(0.f, 0.f);

  for (int s = 0; s < n_s; ++s) {
    const int code = s_off[s];
    const long long d = (long long)((code >> 3) - p.kx2) * p.sx + (long long)((code & 7) - p.ky2) * p.sy;
    const float4* wrow = reinterpret_cast<const float4*>(wsm + (size_t)s * p.nc * GR_OUT);
#pragma unroll 2
    for (int c = 0; c < p.nc; ++c) {
      cf v[GR_HPT];
#pragma unroll
      for (int h = 0; h < GR_HPT; ++h) v[h] = base[pos[h] + d + (long long)c * p.sc];
#pragma unroll
      for (int q = 0; q < GR_OUT / 2; ++q) {
        const float4 w2 = wrow[c * (GR_OUT / 2) + q];
        const cf w0 = cf_make(w2.x, w2.y), w1 = cf_make(w2.z, w2.w);
#pragma unroll
        for (int h = 0; h < GR_HPT; ++h) {
          // acc += w * v  =  (v.x, v.x) * w + (v.y, v.y) * (i w)
          acc[h][2 * q] = pk_fma(mul_i<true>(w0), bc(v[h].y), pk_fma(w0, bc(v[h].x), acc[h][2 * q]));
          acc[h][2 * q + 1] = pk_fma(mul_i<true>(w1), bc(v[h].y), pk_fma(w1, bc(v[h].x), acc[h][2 * q + 1]));
        }
      }
    }
  }
  cf* obase = p.out + (long long)slice * p.ss;
#pragma unroll
  for (int h = 0; h < GR_HPT; ++h) {
    if (!live[h]) continue;
#pragma unroll
    for (int o = 0; o < GR_OUT; ++o) {
      if (o < n_out) {
        const long long a = pos[h] + (long long)(o0 + o) * p.sc;
        obase[a] = cadd(base[a], acc[h][o]);       // recon + kspace (grappa.py:221)
      }
    }
  }
}

// ---- SENSE-style coil combine: |sum_c img_c * conj(sens_c)| (ZIP!/fastmri_prostate/reconstruction/dwi/prostate_dwi_recon.py:
// 106-109) or the complex sum itself (ZIP!/DL_reconstruction/models/varnet.py:199-203, sens_reduce) ------------------------
struct SenseParams {
  const cf* img;       // [B][C][n]
  const cf* sens;      // [Bs][C][n], Bs == B or 1 (shared maps)
  void* out;           // float [B][n] (magnitude) or cf [B][n]
  long long n;
  int B, C, shared_sens, magnitude;
};

__global__ void __launch_bounds__(256) sense_combine_kernel(SenseParams p) {
  const long long total = (long long)p.B * p.n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / p.n, r = i - b * p.n;
    const cf* x = p.img + b * p.C * p.n + r;
    const cf* s = p.sens + (p.shared_sens ? 0 : b * p.C * p.n) + r;
    cf acc = cf_make(0.f, 0.f);
    for (int c = 0; c < p.C; ++c) acc = cadd(acc, cmulc(x[(long long)c * p.n], s[(long long)c * p.n]));
    if (p.magnitude) reinterpret_cast<float*>(p.out)[i] = sqrtf(cnorm2(acc));
    else reinterpret_cast<cf*>(p.out)[i] = acc;
  }
}

}  // namespace mriacl
