// generic_kernels.cuh -- any-size kernels: the generic centred 1-D FFT pass (shared-memory
// Stockham, one output element per thread per stage), coil combine + crop, instance
// normalisation and the small elementwise operators of the reference API.
//
// These serve every (H, W) the fused plans do not cover (odd sizes, 372-wide knee files,
// padded prostate 640x640, tiny test shapes) and the complex-output API (fft2c / ifft2c).
// They are correct for any line length up to MRIACL_MAX_LINE, not tuned: cost per line is
// N * sum(radices) complex MACs.
#pragma once
#include "common.cuh"

namespace mriacl {

#define MRIACL_GEN_THREADS 256
#define MRIACL_GEN_MAX_STAGES 16
#define MRIACL_GEN_SMEM_ELEMS 8192   // complex elements per ping-pong buffer

struct GenFftParams {
  const cf* in;
  cf* out;
  const cf* tw;         // w_N^k = exp(-2 pi i k / N), k < N (forward sign; inverse conjugates)
  const float* mask;    // device float[in_len] applied along the line axis, or nullptr
  int N;                // line length (padded)
  int in_len;           // backed elements of each input line; physical p maps to p - in_pad
  int in_pad;
  int lines_per_frame;  // lines in one 2-D frame
  int n_frames;
  int lines_per_block;  // L
  int lines_contig;     // 1: consecutive lines are adjacent in memory (column pass), 0: row pass
  long long in_es, in_ls;     // input element / line strides (complex elements)
  long long in_sb, in_sa, in_sc;  // frame f = (b*A + a)*C + c -> b*in_sb + a*in_sa + c*in_sc
  int A, C;
  long long out_es, out_ls, out_fs;
  int inverse;
  float scale;
  int n_stages;
  int radix[MRIACL_GEN_MAX_STAGES];
};

__global__ void __launch_bounds__(MRIACL_GEN_THREADS) generic_fft_kernel(GenFftParams p) {
  MRIACL_DYN_SMEM(cf, sm);
  const int N = p.N, L = p.lines_per_block;
  cf* buf0 = sm;
  cf* buf1 = sm + L * N;
  const int blocks_per_frame = (p.lines_per_frame + L - 1) / L;
  const int f = blockIdx.x / blocks_per_frame;
  const int line0 = (blockIdx.x % blocks_per_frame) * L;
  const int nl = min(L, p.lines_per_frame - line0);
  const int b = f / (p.A * p.C), a = (f / p.C) % p.A, c = f % p.C;
  const cf* in = p.in + b * p.in_sb + a * p.in_sa + c * p.in_sc;
  cf* out = p.out + f * p.out_fs;
  const int total = nl * N;

  // load: logical index i of the line <- physical (i + N/2) % N, zero outside the backed window
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int l, ph;
    if (p.lines_contig) { ph = e / nl; l = e - ph * nl; } else { l = e / N; ph = e - l * N; }
    const int src = ph - p.in_pad;
    cf v = cf_make(0.f, 0.f);
    if (src >= 0 && src < p.in_len) {
      v = in[(long long)(line0 + l) * p.in_ls + (long long)src * p.in_es];
      if (p.mask) v = cscale(v, p.mask[src]);
    }
    buf0[l * N + logical_of_phys(ph, N)] = v;
  }
  __syncthreads();

  int Ns = 1;
  for (int s = 0; s < p.n_stages; ++s) {
    const int R = p.radix[s];
    const int NR = N / R;
    const int tstep = N / (Ns * R);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
      const int l = e / N, i = e - l * N;
      const int k = i % Ns, t = i / Ns;
      const int q = t % R, jh = t / R;
      const int j = jh * Ns + k;
      int step = (int)(((long long)k * tstep + (long long)q * NR) % N);
      int idx = 0;
      const cf* src = buf0 + l * N + j;
      float re = 0.f, im = 0.f;
      for (int r = 0; r < R; ++r) {
        const cf x = src[r * NR];
        cf w = p.tw[idx];
        if (p.inverse) w.y = -w.y;
        re = fmaf(x.x, w.x, fmaf(-x.y, w.y, re));
        im = fmaf(x.x, w.y, fmaf(x.y, w.x, im));
        idx += step;
        if (idx >= N) idx -= N;
      }
      buf1[e] = cf_make(re, im);
    }
    __syncthreads();
    cf* tswap = buf0; buf0 = buf1; buf1 = tswap;
    Ns *= R;
  }

  // store: logical m -> physical (m + N/2) % N
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int l, ph;
    if (p.lines_contig) { ph = e / nl; l = e - ph * nl; } else { l = e / N; ph = e - l * N; }
    const cf v = buf0[l * N + logical_of_phys(ph, N)];
    out[(long long)(line0 + l) * p.out_ls + (long long)ph * p.out_es] = cscale(v, p.scale);
  }
}

// out[b][r][c] = (1/A) sum_a sqrt( sum_c |img[(b,a,c)][row(r)][col0 + c]|^2 )
struct RssCropParams {
  const cf* img;   // [n_slices*A*C][H][Wp] full centred images
  float* out;      // [n_slices][oh][ow]
  int n_slices, A, C, H, Wp, oh, ow, row0, col0, flip;
};

__global__ void __launch_bounds__(256) rss_crop_kernel(RssCropParams p) {
  const long long total = (long long)p.n_slices * p.oh * p.ow;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int cc = (int)(e % p.ow);
    const int rr = (int)((e / p.ow) % p.oh);
    const int s = (int)(e / ((long long)p.ow * p.oh));
    const int row = p.flip ? p.H - 1 - (p.row0 + rr) : p.row0 + rr;
    const long long pix = (long long)row * p.Wp + p.col0 + cc;
    float avg = 0.f;
    for (int a = 0; a < p.A; ++a) {
      float acc = 0.f;
      for (int c = 0; c < p.C; ++c) {
        const cf v = p.img[((long long)(s * p.A + a) * p.C + c) * p.H * p.Wp + pix];
        acc = cnorm2_acc(v, acc);
      }
      avg += sqrtf(acc);
    }
    p.out[e] = p.A > 1 ? avg / (float)p.A : avg;
  }
}

// ---- block reductions ------------------------------------------------------------
__device__ __forceinline__ double block_sum_double(double v, double* red /* >= 33 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nw ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// Instance normalisation, one block per image: mean, UNBIASED std (two-pass), then
// out = (in - mean) / (std + eps).  `partials` (optional) holds per-tile (count, mean, M2)
// triples written by the fused row pass; when given they are merged (Chan et al.) instead of
// re-reading the image for the statistics.
struct NormParams {
  const float* in;
  float* out;          // may alias in; nullptr = statistics only
  float* mean_std;     // [B][2] or nullptr
  const float* partials;  // [B][n_part][3] or nullptr
  int n_part;
  int n_split;         // blocks per image (1 unless partials are given)
  long long n;         // elements per image
  float eps;
  int normalize;
};

// grid = B * n_split: block (b, part) normalises its 1/n_split share of image b.  With partials the statistics
// are merged by one thread and broadcast (a handful of FMAs); the share is then streamed with 128-bit accesses.
__global__ void __launch_bounds__(256) normalize_instance_kernel(NormParams p) {
  __shared__ double red[33];
  __shared__ float s_mean, s_std;
  const int b = blockIdx.x / p.n_split, part = blockIdx.x - b * p.n_split;
  const float* x = p.in + (long long)b * p.n;
  if (p.partials) {
    // Chan's pairwise merge of the tiles' (count, mean, M2), in parallel: lane i of warp 0 starts from tiles i, i+32, ...
    // and the lanes combine by a shuffle tree (double precision; a handful of operations instead of a serial loop)
    if (threadIdx.x < 32) {
      double cnt = 0.0, mean = 0.0, m2 = 0.0;
      auto merge = [&](double nb, double mb, double sb) {
        if (nb > 0.0) {
          const double d = mb - mean, tot = cnt + nb;
          mean += d * nb / tot;
          m2 += sb + d * d * cnt * nb / tot;
          cnt = tot;
        }
      };
      for (int t = threadIdx.x; t < p.n_part; t += 32) {
        const float* q = p.partials + ((long long)b * p.n_part + t) * 3;
        merge((double)q[0], (double)q[1], (double)q[2]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double nb = __shfl_xor_sync(0xffffffffu, cnt, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                     sb = __shfl_xor_sync(0xffffffffu, m2, o);
        // both partners must arrive at the same value: merge (lower lane's triple) <- (upper lane's triple)
        const bool low = (threadIdx.x & o) == 0;
        const double an = low ? cnt : nb, am = low ? mean : mb, as = low ? m2 : sb;
        const double bn = low ? nb : cnt, bm = low ? mb : mean, bs = low ? sb : m2;
        cnt = an; mean = am; m2 = as;
        merge(bn, bm, bs);
      }
      if (threadIdx.x == 0) {
        s_mean = (float)mean;
        s_std = (float)sqrt(m2 / (double)(p.n - 1));
      }
    }
    __syncthreads();
  } else {
    double s = 0.0;
    for (long long i = threadIdx.x; i < p.n; i += blockDim.x) s += (double)x[i];
    const double mean = block_sum_double(s, red) / (double)p.n;
    double q = 0.0;
    for (long long i = threadIdx.x; i < p.n; i += blockDim.x) { const double d = (double)x[i] - mean; q += d * d; }
    const double m2 = block_sum_double(q, red);
    if (threadIdx.x == 0) { s_mean = (float)mean; s_std = (float)sqrt(m2 / (double)(p.n - 1)); }
    __syncthreads();
  }
  const float fmean = s_mean, fstd = s_std;
  if (threadIdx.x == 0 && part == 0 && p.mean_std) { p.mean_std[2 * b] = fmean; p.mean_std[2 * b + 1] = fstd; }
  if (p.normalize && p.out) {
    float* y = p.out + (long long)b * p.n;
    const float den = fstd + p.eps;
    const long long per = (((p.n + p.n_split - 1) / p.n_split) + 3) & ~3LL;      // shares start on 16-byte boundaries
    const long long lo = part * per < p.n ? part * per : p.n, hi = lo + per < p.n ? lo + per : p.n;
    const bool vec = ((lo | hi) & 3) == 0 && ((((unsigned long long)x) | ((unsigned long long)y)) & 15) == 0;
    const float inv = 1.0f / den;
    (void)inv;
    if (vec) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      float4* y4 = reinterpret_cast<float4*>(y);
      const long long i0 = lo / 4, i1 = hi / 4;
      long long i = i0 + threadIdx.x;
      // four independent 128-bit loads in flight per thread
      for (; i + 3LL * blockDim.x < i1; i += 4LL * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = x4[i + (long long)u * blockDim.x];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u].x = (v[u].x - fmean) / den; v[u].y = (v[u].y - fmean) / den;
          v[u].z = (v[u].z - fmean) / den; v[u].w = (v[u].w - fmean) / den;
          y4[i + (long long)u * blockDim.x] = v[u];
        }
      }
      for (; i < i1; i += blockDim.x) {
        float4 v = x4[i];
        v.x = (v.x - fmean) / den; v.y = (v.y - fmean) / den; v.z = (v.z - fmean) / den; v.w = (v.w - fmean) / den;
        y4[i] = v;
      }
    } else {
      for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) y[i] = (x[i] - fmean) / den;
    }
  }
}

// ---- small elementwise operators of the reference API --------------------------------
__global__ void __launch_bounds__(256) complex_abs_kernel(const cf* in, float* out, long long n, int squared) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const cf v = in[i];
    const float s = fmaf(v.x, v.x, v.y * v.y);
    out[i] = squared ? s : sqrtf(s);
  }
}

__global__ void __launch_bounds__(256) rss_kernel(const float* in, float* out, long long outer, int C, long long inner, int is_complex) {
  const long long total = outer * inner;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long o = e / inner, i = e - o * inner;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
      const long long idx = (o * C + c) * inner + i;
      if (is_complex) acc = cnorm2_acc(reinterpret_cast<const cf*>(in)[idx], acc);
      else { const float v = in[idx]; acc = fmaf(v, v, acc); }
    }
    out[e] = sqrtf(acc);
  }
}

// centre crop or zero-pad the last two axes; element = 4 or 8 bytes (words4 = 1 or 2)
__global__ void __launch_bounds__(256) crop_or_pad_kernel(const float* in, float* out, int B, int H, int W, int oh, int ow, int words4) {
  const int kh = min(H, oh), kw = min(W, ow);
  const int sh = (H - kh) / 2, sw = (W - kw) / 2, dh = (oh - kh) / 2, dw = (ow - kw) / 2;
  const long long total = (long long)B * oh * ow;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % ow), r = (int)((e / ow) % oh);
    const long long b = e / ((long long)ow * oh);
    const int rs = r - dh + sh, cs = c - dw + sw;
    const bool inside = (r >= dh) && (r < dh + kh) && (c >= dw) && (c < dw + kw);
    for (int w = 0; w < words4; ++w)
      out[e * words4 + w] = inside ? in[((b * H + rs) * W + cs) * words4 + w] : 0.f;
  }
}

}  // namespace mriacl
