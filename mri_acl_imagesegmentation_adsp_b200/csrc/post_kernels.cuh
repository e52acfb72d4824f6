// post_kernels.cuh -- the per-slice steps that follow the reconstruction on the reference's live call path
// (REF/src/preprocess/mri_preprocess.py:59-84): percentile clip (:182-185), bilinear resize (:187-191), in-mask
// z-score (:216-224) and the [0,1] preview (:226-233).  (The Otsu / morphology body mask between clip and resize needs
// scikit-image and stays on the host: the mask is an input here.)
//
// Everything is float32 and follows the reference's own arithmetic step by step, because parity is judged on it:
//  * np.percentile (numpy >= 2.0, method "linear") on a float32 image works entirely in float32: q32 = float32(q) / 100,
//    virtual index v = float32(n - 1) * q32, i = floor(v), g = v - i, result = a + (b - a) * g for g < 0.5, else
//    b - (b - a) * (1 - g), with a, b the i-th and (i+1)-th smallest values -- unfused multiplies and adds.
//    The two order statistics are found EXACTLY by a radix select over the monotone integer image of the floats
//    (four 8-bit passes, one CTA per image, all four ranks of the two percentiles at once).
//  * F.interpolate(mode="bilinear", align_corners=False) on the CPU: scale = in / out (float32), source coordinate
//    s = scale * (d + 0.5) - 0.5 clamped at 0, taps i0 = int(s), i1 = i0 + (i0 < in - 1), weights (1 - l, l) with
//    l = s - i0, value = h0 * (w0 * p00 + w1 * p01) + h1 * (w0 * p10 + w1 * p11), unfused.
//  * z-score: mean and POPULATION standard deviation of the pixels inside the mask (of the whole image when fewer than
//    ten are inside), std <= 1e-6 replaced by 1; preview: (x - lo) / float32(hi - lo + 1e-6) with lo / hi the extrema
//    inside the mask (of the whole image when the mask is empty).  Statistics are accumulated in double.
//
// Both one-image kernels are bound by a single CTA's sweep of its image, and a batch of 64 images fills 64 of 148 SMs, so
// they run as thread-block CLUSTERS: CL CTAs (1, 2 or 4, chosen so that B x CL fills the machine) share one image, each
// sweeps 1/CL of it, and the per-CTA results (histograms, partial sums, extrema) are combined through distributed shared
// memory -- every CTA reads its peers' shared memory (cluster.map_shared_rank) in rank order, so all of them continue from
// bit-identical totals without a trip through global memory.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace mriacl {
namespace cg = cooperative_groups;

constexpr int POST_T = 1024;          // threads of the one-CTA-per-image kernels
constexpr int POST_HIST_BYTES = 4 * 256 * 32 * 4;   // dynamic shared memory of percentile_clip_kernel
constexpr int POST_ILP = 4;           // 128-bit loads a thread keeps in flight per sweep iteration

__device__ __forceinline__ unsigned post_key(float x) {
  const unsigned u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);       // monotone: key(a) < key(b)  <=>  a < b
}
__device__ __forceinline__ float post_unkey(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// numpy's float32 lerp between the order statistics a <= b
__device__ __forceinline__ float post_lerp(float a, float b, float g) {
  const float d = __fsub_rn(b, a);
  return g < 0.5f ? __fadd_rn(a, __fmul_rn(d, g)) : __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g)));
}

struct PercentileParams {
  const float* in;      // [B][n]
  float* out;           // [B][n] clipped image, or nullptr
  float* lo_hi;         // [B][2] or nullptr
  long long n;
  float pmin, pmax;     // percent, 0..100
};

// one cluster of CL CTAs per image (grid = B * CL)
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(POST_T) percentile_clip_kernel(PercentileParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = CL > 1 ? (int)cluster.block_rank() : 0;
  const int img = (int)blockIdx.x / CL;
  MRIACL_DYN_SMEM(unsigned, hist);          // [4 ranks][256 bins][32 lanes]: a lane owns its column, so the lanes of a warp never
                                            // meet on an address or a bank however the values cluster (POST_HIST_BYTES)
  __shared__ unsigned folded[4][256];       // this CTA's share of the image
  __shared__ unsigned merged[4][256];       // the whole image (sum over the cluster, same values in every CTA)
  __shared__ unsigned prefix[4];
  __shared__ unsigned long long rank[4];
  __shared__ int alias[4];
  __shared__ float s_lo_hi[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const float* x = p.in + (long long)img * p.n;
  const long long n = p.n;

  // virtual indices in float32, exactly as numpy computes them
  const float nm1 = (float)(n - 1);
  const float v_lo = __fmul_rn(nm1, __fdiv_rn(p.pmin, 100.0f)), v_hi = __fmul_rn(nm1, __fdiv_rn(p.pmax, 100.0f));
  const float f_lo = floorf(v_lo), f_hi = floorf(v_hi);
  if (tid == 0) {
    const long long i_lo = (long long)f_lo, i_hi = (long long)f_hi;
    rank[0] = (unsigned long long)min(max(i_lo, 0LL), n - 1);
    rank[1] = (unsigned long long)min(max(i_lo + 1, 0LL), n - 1);
    rank[2] = (unsigned long long)min(max(i_hi, 0LL), n - 1);
    rank[3] = (unsigned long long)min(max(i_hi + 1, 0LL), n - 1);
    for (int t = 0; t < 4; ++t) prefix[t] = 0u;
  }
  __syncthreads();

  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid == 0)
      for (int t = 0; t < 4; ++t) {          // ranks that still share a prefix share a histogram
        alias[t] = t;
        for (int u = 0; u < t; ++u) if (prefix[u] == prefix[t]) { alias[t] = u; break; }
      }
    for (int i = tid; i < 4 * 256 * 32; i += POST_T) hist[i] = 0u;
    __syncthreads();
    const unsigned pf0 = prefix[0], pf1 = prefix[1], pf2 = prefix[2], pf3 = prefix[3];
    const bool own1 = alias[1] == 1, own2 = alias[2] == 2, own3 = alias[3] == 3;
    // An image's values share a handful of exponent bytes, so in the first passes most lanes of a warp hit the same bin:
    // with one histogram per CTA those shared-memory atomics serialise 32 ways (and a match_any vote to merge them costs as
    // much); with a column per lane they never collide inside a warp.
    auto count = [&](unsigned k, bool in_range) {
      const unsigned hi_bits = pass == 0 ? 0u : (k >> (shift + 8));
      const unsigned slot = ((k >> shift) & 255u) * 32u + lane;
      if (in_range && hi_bits == pf0) atomicAdd(&hist[slot], 1u);
      if (in_range && own1 && hi_bits == pf1) atomicAdd(&hist[8192 + slot], 1u);
      if (in_range && own2 && hi_bits == pf2) atomicAdd(&hist[16384 + slot], 1u);
      if (in_range && own3 && hi_bits == pf3) atomicAdd(&hist[24576 + slot], 1u);
    };
    // each sweep is bound by the latency of its loads, so a thread keeps POST_ILP 128-bit loads in flight per iteration
    const bool vec = (n & 3) == 0 && ((reinterpret_cast<unsigned long long>(x) & 15ull) == 0);
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      const long long n4_all = n >> 2;
      const long long lo4 = n4_all * crank / CL, n4 = n4_all * (crank + 1) / CL;     // this CTA's share [lo4, n4)
      const long long n_iter = (n4 - lo4 + (long long)POST_T * POST_ILP - 1) / ((long long)POST_T * POST_ILP);
      for (long long it = 0; it < n_iter; ++it) {
        float4 v[POST_ILP];
        bool ok[POST_ILP];
#pragma unroll
        for (int u = 0; u < POST_ILP; ++u) {
          const long long i4 = lo4 + (it * POST_ILP + u) * POST_T + tid;
          ok[u] = i4 < n4;
          v[u] = ok[u] ? x4[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < POST_ILP; ++u) {
          count(post_key(v[u].x), ok[u]); count(post_key(v[u].y), ok[u]);
          count(post_key(v[u].z), ok[u]); count(post_key(v[u].w), ok[u]);
        }
      }
    } else {
      const long long lo = n * crank / CL, hi = n * (crank + 1) / CL;
      const long long n_iter = (hi - lo + POST_T - 1) / POST_T;
      for (long long it = 0; it < n_iter; ++it) {
        const long long i = lo + it * POST_T + tid;
        count(i < hi ? post_key(x[i]) : 0u, i < hi);
      }
    }
    __syncthreads();
    for (int i = tid; i < 4 * 256; i += POST_T) {      // fold the lane columns
      unsigned sum = 0;
#pragma unroll 8
      for (int c = 0; c < 32; ++c) sum += hist[i * 32 + ((c + lane) & 31)];
      folded[i >> 8][i & 255] = sum;
    }
    if (CL > 1) {
      cluster.sync();                                    // every CTA's folded histogram is complete
      for (int i = tid; i < 4 * 256; i += POST_T) {
        unsigned sum = 0;
#pragma unroll
        for (int r = 0; r < CL; ++r) sum += cluster.map_shared_rank(&folded[0][0], r)[i];
        merged[i >> 8][i & 255] = sum;
      }
      cluster.sync();                                    // ... and has been read by every peer before the next pass refills it
    } else {
      __syncthreads();
      for (int i = tid; i < 4 * 256; i += POST_T) merged[i >> 8][i & 255] = folded[i >> 8][i & 255];
      __syncthreads();
    }
    // bin of each rank: warp t scans histogram t (8 bins per lane, shuffle prefix sum) -- the bin b with
    // cum(b) <= r < cum(b + 1), the last bin if r is beyond the total
    if (tid < 128) {
      const int t = tid >> 5;
      const unsigned* h = merged[alias[t]];
      unsigned c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = h[8 * lane + j]; sum += c[j]; }
      unsigned inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
      const unsigned long long r = rank[t];
      const unsigned long long exc = (unsigned long long)(inc - sum);
      const bool beyond = r >= (unsigned long long)__shfl_sync(0xffffffffu, inc, 31);
      const bool here = beyond ? lane == 31 : (r >= exc && r < (unsigned long long)inc);
      __syncwarp();
      if (here) {
        unsigned long long rr = r - exc;
        int j = 0;
        for (; j < 7; ++j) { if (rr < c[j]) break; rr -= c[j]; }
        if (beyond) {                                    // (cannot happen with exact counts; same answer as the serial scan)
          rr = r - exc; j = 0;
          for (; j < 7; ++j) rr -= c[j];
          j = 7;
        }
        rank[t] = rr;
        prefix[t] = (prefix[t] << 8) | (unsigned)(8 * lane + j);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const float a0 = post_unkey(prefix[0]), b0 = post_unkey(prefix[1]), a1 = post_unkey(prefix[2]), b1 = post_unkey(prefix[3]);
    s_lo_hi[0] = post_lerp(a0, b0, __fsub_rn(v_lo, f_lo));
    s_lo_hi[1] = post_lerp(a1, b1, __fsub_rn(v_hi, f_hi));
    if (p.lo_hi && crank == 0) { p.lo_hi[2 * img] = s_lo_hi[0]; p.lo_hi[2 * img + 1] = s_lo_hi[1]; }
  }
  __syncthreads();
  if (p.out) {
    const float lo = s_lo_hi[0], hi = s_lo_hi[1];
    float* y = p.out + (long long)img * n;
    const long long i0 = n * crank / CL, i1 = n * (crank + 1) / CL;
    for (long long i = i0 + tid; i < i1; i += POST_T) y[i] = fminf(fmaxf(x[i], lo), hi);      // np.clip = minimum(maximum(x, lo), hi)
  }
}

struct ResizeParams {
  const float* in;            // [B][H][W] float32, or
  const unsigned char* in_u8; // [B][H][W] uint8 (a mask: values are used as 0 / 1 floats), exactly one of the two is set
  float* out;                 // [B][oh][ow] float32, or
  unsigned char* out_u8;      // [B][oh][ow] uint8 = (resized > 0.5)
  const float* lo_hi;         // [B][2] clip applied to every tap before the interpolation, or nullptr
  int B, H, W, oh, ow;
  // optional second plane through the same source coordinates and weights in the same launch (the body mask beside its image)
  const unsigned char* mask_in;   // [B][H][W] uint8 or nullptr
  unsigned char* mask_out;        // [B][oh][ow] uint8 = (resized > 0.5)
};

// thread = output pixel
__global__ void __launch_bounds__(256) resize_bilinear_kernel(ResizeParams p) {
  const long long total = (long long)p.B * p.oh * p.ow;
  const float sh = __fdiv_rn((float)p.H, (float)p.oh), sw = __fdiv_rn((float)p.W, (float)p.ow);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % p.ow);
    const long long t = i / p.ow;
    const int oy = (int)(t % p.oh), b = (int)(t / p.oh);
    float sy = __fsub_rn(__fmul_rn(sh, __fadd_rn((float)oy, 0.5f)), 0.5f), sx = __fsub_rn(__fmul_rn(sw, __fadd_rn((float)ox, 0.5f)), 0.5f);
    sy = sy < 0.f ? 0.f : sy; sx = sx < 0.f ? 0.f : sx;
    int y0 = (int)sy, x0 = (int)sx;
    y0 = min(y0, p.H - 1); x0 = min(x0, p.W - 1);
    const int y1 = y0 + (y0 < p.H - 1 ? 1 : 0), x1 = x0 + (x0 < p.W - 1 ? 1 : 0);
    const float ly = __fsub_rn(sy, (float)y0), lx = __fsub_rn(sx, (float)x0);
    const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
    const long long base = (long long)b * p.H * p.W;
    float p00, p01, p10, p11;
    if (p.in) {
      p00 = p.in[base + (long long)y0 * p.W + x0]; p01 = p.in[base + (long long)y0 * p.W + x1];
      p10 = p.in[base + (long long)y1 * p.W + x0]; p11 = p.in[base + (long long)y1 * p.W + x1];
      if (p.lo_hi) {
        const float lo = p.lo_hi[2 * b], hi = p.lo_hi[2 * b + 1];
        p00 = fminf(fmaxf(p00, lo), hi); p01 = fminf(fmaxf(p01, lo), hi);
        p10 = fminf(fmaxf(p10, lo), hi); p11 = fminf(fmaxf(p11, lo), hi);
      }
    } else {
      p00 = (float)p.in_u8[base + (long long)y0 * p.W + x0]; p01 = (float)p.in_u8[base + (long long)y0 * p.W + x1];
      p10 = (float)p.in_u8[base + (long long)y1 * p.W + x0]; p11 = (float)p.in_u8[base + (long long)y1 * p.W + x1];
    }
    const float top = __fadd_rn(__fmul_rn(hx, p00), __fmul_rn(lx, p01)), bot = __fadd_rn(__fmul_rn(hx, p10), __fmul_rn(lx, p11));
    const float v = __fadd_rn(__fmul_rn(hy, top), __fmul_rn(ly, bot));
    if (p.out) p.out[i] = v; else p.out_u8[i] = v > 0.5f ? 1 : 0;
    if (p.mask_in) {
      const float m00 = (float)p.mask_in[base + (long long)y0 * p.W + x0], m01 = (float)p.mask_in[base + (long long)y0 * p.W + x1];
      const float m10 = (float)p.mask_in[base + (long long)y1 * p.W + x0], m11 = (float)p.mask_in[base + (long long)y1 * p.W + x1];
      const float mt = __fadd_rn(__fmul_rn(hx, m00), __fmul_rn(lx, m01)), mb = __fadd_rn(__fmul_rn(hx, m10), __fmul_rn(lx, m11));
      p.mask_out[i] = __fadd_rn(__fmul_rn(hy, mt), __fmul_rn(ly, mb)) > 0.5f ? 1 : 0;
    }
  }
}

struct ZscoreParams {
  const float* in;            // [B][n] (resized, clipped) image; may alias out_z
  const unsigned char* mask;  // [B][n] or nullptr (= every pixel inside)
  float* out_z;               // [B][n] (x - mean) / std, or nullptr
  float* out_01;              // [B][n] preview, or nullptr
  float* stats;               // [B][6]: mean, std (after the floor), lo, hi, pixels inside the mask, 1 if the mask was used
  long long n;
};

template <class T, class Op> __device__ __forceinline__ T post_block_reduce(T v, T* scratch /* 32 */, Op op, T identity) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = identity;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = op(r, scratch[w]);
  return r;
}

// one cluster of CL CTAs per image (grid = B * CL): three sweeps (count / sum / extrema -> mean; squared deviations -> std;
// outputs), each CTA over its share of the image; the partial results meet through distributed shared memory
template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(POST_T) zscore_preview_kernel(ZscoreParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = CL > 1 ? (int)cluster.block_rank() : 0;
  const int img = (int)blockIdx.x / CL;
  __shared__ double sd[32];
  __shared__ float sf[32];
  __shared__ double xd[4];        // this CTA's partial count / sums, read by the peers
  __shared__ float xf[4];         // ... and extrema
  const int tid = threadIdx.x;
  const long long n = p.n;
  const long long i0 = n * crank / CL, i1 = n * (crank + 1) / CL;      // this CTA's share
  const float* x = p.in + (long long)img * n;
  const unsigned char* m = p.mask ? p.mask + (long long)img * n : nullptr;
  auto add = [](double a, double b) { return a + b; };
  auto fmn = [](float a, float b) { return fminf(a, b); };
  auto fmx = [](float a, float b) { return fmaxf(a, b); };

  double cnt_in = 0.0, sum_in = 0.0, sum_all = 0.0;
  float mn_in = INFINITY, mx_in = -INFINITY, mn_all = INFINITY, mx_all = -INFINITY;
  for (long long i = i0 + tid; i < i1; i += POST_T) {
    const float v = x[i];
    sum_all += (double)v; mn_all = fminf(mn_all, v); mx_all = fmaxf(mx_all, v);
    if (!m || m[i] > 0) { cnt_in += 1.0; sum_in += (double)v; mn_in = fminf(mn_in, v); mx_in = fmaxf(mx_in, v); }
  }
  cnt_in = post_block_reduce(cnt_in, sd, add, 0.0);
  sum_in = post_block_reduce(sum_in, sd, add, 0.0);
  sum_all = post_block_reduce(sum_all, sd, add, 0.0);
  mn_in = post_block_reduce(mn_in, sf, fmn, INFINITY);
  mx_in = post_block_reduce(mx_in, sf, fmx, -INFINITY);
  mn_all = post_block_reduce(mn_all, sf, fmn, INFINITY);
  mx_all = post_block_reduce(mx_all, sf, fmx, -INFINITY);
  if (CL > 1) {                      // combine over the cluster, in rank order (every CTA arrives at the same values)
    if (tid == 0) { xd[0] = cnt_in; xd[1] = sum_in; xd[2] = sum_all; xf[0] = mn_in; xf[1] = mx_in; xf[2] = mn_all; xf[3] = mx_all; }
    cluster.sync();
    cnt_in = sum_in = sum_all = 0.0;
    mn_in = mn_all = INFINITY; mx_in = mx_all = -INFINITY;
#pragma unroll
    for (int r = 0; r < CL; ++r) {
      const double* rd = cluster.map_shared_rank(xd, r);
      const float* rf = cluster.map_shared_rank(xf, r);
      cnt_in += rd[0]; sum_in += rd[1]; sum_all += rd[2];
      mn_in = fminf(mn_in, rf[0]); mx_in = fmaxf(mx_in, rf[1]); mn_all = fminf(mn_all, rf[2]); mx_all = fmaxf(mx_all, rf[3]);
    }
    cluster.sync();                  // the peers have read xd / xf: it may be rewritten below
  }
  const bool use_mask = cnt_in >= 10.0;                     // `vals.size < 10` -> statistics of the whole image
  const double mean_d = use_mask ? sum_in / cnt_in : sum_all / (double)n;
  double m2 = 0.0;
  for (long long i = i0 + tid; i < i1; i += POST_T) {
    if (!use_mask || !m || m[i] > 0) { const double d = (double)x[i] - mean_d; m2 += d * d; }
  }
  m2 = post_block_reduce(m2, sd, add, 0.0);
  if (CL > 1) {
    if (tid == 0) xd[3] = m2;
    cluster.sync();
    m2 = 0.0;
#pragma unroll
    for (int r = 0; r < CL; ++r) m2 += cluster.map_shared_rank(xd, r)[3];
    cluster.sync();                  // no CTA leaves (and frees its shared memory) while a peer still reads it
  }
  const float mean = (float)mean_d;
  float stdv = (float)sqrt(m2 / (use_mask ? cnt_in : (double)n));   // np.std: population
  stdv = stdv > 1e-6f ? stdv : 1.0f;
  const float lo = cnt_in > 0.0 ? mn_in : mn_all, hi = cnt_in > 0.0 ? mx_in : mx_all;
  const float den = (float)((double)hi - (double)lo + 1e-6);
  if (tid == 0 && p.stats && crank == 0) {
    float* s = p.stats + 6 * (long long)img;
    s[0] = mean; s[1] = stdv; s[2] = lo; s[3] = hi; s[4] = (float)cnt_in; s[5] = use_mask ? 1.f : 0.f;
  }
  float* z = p.out_z ? p.out_z + (long long)img * n : nullptr;
  float* q = p.out_01 ? p.out_01 + (long long)img * n : nullptr;
  for (long long i = i0 + tid; i < i1; i += POST_T) {
    const float v = x[i];
    if (q) q[i] = __fdiv_rn(__fsub_rn(v, lo), den);
    if (z) z[i] = __fdiv_rn(__fsub_rn(v, mean), stdv);
  }
}

// 2.5-D stacking and the encoder's input normalisation (REF/src/dataio/datasets.py:90-95,128-131): channel d of output slice s
// is input slice clamp(s + d - k/2, 0, S-1); with `repeat` a single channel is repeated to `k` channels instead (ImageNet
// encoders); mean / std (per output channel) give (x - mean) / std, unfused like torch.
struct StackParams {
  const float* in;     // [S][n]
  float* out;          // [S][k][n]
  const float* mean;   // [k] or nullptr
  const float* stdv;   // [k] or nullptr
  long long n;
  int S, k, repeat;
};

__global__ void __launch_bounds__(256) stack25d_kernel(StackParams p) {
  const long long total = (long long)p.S * p.k * p.n;
  const int half = p.k / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % p.n;
    const long long t = i / p.n;
    const int d = (int)(t % p.k), s = (int)(t / p.k);
    const int src = p.repeat ? s : min(max(s + d - half, 0), p.S - 1);
    float v = p.in[(long long)src * p.n + r];
    if (p.mean) v = __fdiv_rn(__fsub_rn(v, p.mean[d]), p.stdv[d]);
    p.out[i] = v;
  }
}

}  // namespace mriacl
