// rowpass_generic.cuh -- pruned, fused row pass for ANY (padded) phase-encode length, used behind the H = 640 column
// pass when the width has no specialised row kernel (372-, 320-, 400-wide knee files ...).
//
// The any-size path of generic_kernels.cuh transforms every frame completely (all H rows, then all Wp columns), keeps
// the full complex images in the workspace and combines them afterwards.  Here the column pass has already reduced a
// frame to T[j][row] -- sampled columns x kept rows -- so the row pass only touches the `oh` kept rows, scatters the
// sampled values into a zero line in shared memory, runs the Stockham stages of generic_kernels.cuh on L lines at a
// time, and accumulates |X|^2 of the kept columns in shared memory across the coil loop (RSS, mean over averages
// after the RSS, crop, tile statistics: same epilogue as the specialised row passes).  Nothing but the final image
// is written to global memory.
#pragma once
#include "common.cuh"
#include "rowpass.cuh"

namespace mriacl {

constexpr int RG_T = 256;
constexpr int RG_MAX_STAGES = 16;

struct RowGenParams {
  const cf* T;                 // [n_slices * A * C][n_act][ohp]
  int n_act, oh, ohp;
  const int* act_logical;      // [n_act] logical (un-shifted) index of active column j in the padded line
  const cf* tw;                // w_N^k = exp(-2 pi i k / N) (forward sign; conjugated here: inverse transform)
  int N;                       // padded line length
  int L;                       // lines (output rows) per item
  float* out;
  float* partials;             // [n_slices][n_tiles][3] or nullptr
  int ow, col0;
  int A, C;
  float scale;
  int n_slices, n_tiles;       // n_tiles = ceil(oh / L)
  int n_stages;
  int radix[RG_MAX_STAGES];
};

__host__ __device__ inline int rowgen_smem_bytes(int N, int L, int ow, int n_stages) {
  return 2 * L * N * 8 + 2 * L * (ow + 1) * 4 + n_stages * N * 8 + N * 8;
}

__global__ void __launch_bounds__(RG_T, 3) rowpass_generic_kernel(RowGenParams p) {
  MRIACL_DYN_SMEM(cf, sm);
  const int N = p.N, L = p.L;
  float* accsm = reinterpret_cast<float*>(sm + 2 * (size_t)L * N);      // [L][ow + 1] sum over coils of |X|^2
  float* avsm = accsm + L * (p.ow + 1);                                  // [L][ow + 1] sum over averages of the RSS
  int2* stab = reinterpret_cast<int2*>(avsm + L * (p.ow + 1));          // [n_stages][N] (first input, twiddle step) per output
  cf* twsm = reinterpret_cast<cf*>(stab + p.n_stages * N);               // [N] twiddles (shared memory: one load per MAC)
  __shared__ float red[RG_T / 32];
  const int tid = threadIdx.x;
  const int opitch = p.ow + 1;
  for (int i = tid; i < N; i += RG_T) twsm[i] = p.tw[i];
  const int n_frames = p.A * p.C;
  const long long frame_elems = (long long)p.n_act * p.ohp;
  const int n_items = p.n_slices * p.n_tiles;
  // Stockham index maps, once per CTA: output i of stage st reads inputs j, j + N/R, ... with twiddles idx = 0, step, 2 step ...
  {
    int Ns = 1;
    for (int st = 0; st < p.n_stages; ++st) {
      const int R = p.radix[st], NR = N / R, tstep = N / (Ns * R);
      for (int i = tid; i < N; i += RG_T) {
        const int k = i % Ns, t = i / Ns;
        const int q = t % R, jh = t / R;
        stab[st * N + i] = make_int2(jh * Ns + k, (int)(((long long)k * tstep + (long long)q * NR) % N));
      }
      Ns *= R;
    }
  }
  __syncthreads();
  // (line, index) of element e = tid + n RG_T advance without divisions
  const int step_l = RG_T / N, step_i = RG_T - step_l * N;
  const int l_first = tid / N, i_first = tid - l_first * N;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int s = item / p.n_tiles, tile = item - s * p.n_tiles;
    const int nl = min(L, p.oh - tile * L);
    const int total = nl * N;
    const cf* Tit = p.T + (long long)s * n_frames * frame_elems + tile * L;
    for (int i = tid; i < L * opitch; i += RG_T) { accsm[i] = 0.f; avsm[i] = 0.f; }
    int coil = 0;
    for (int f = 0; f < n_frames; ++f) {
      cf* buf0 = sm;
      cf* buf1 = sm + (size_t)L * N;
      // zero line, then the sampled columns at their logical positions
      for (int e = tid; e < total; e += RG_T) buf0[e] = cf_make(0.f, 0.f);
      __syncthreads();
      const cf* Tf = Tit + (long long)f * frame_elems;
      for (int e = tid; e < p.n_act * nl; e += RG_T) {
        const int j = e / nl, l = e - j * nl;
        buf0[l * N + p.act_logical[j]] = Tf[(long long)j * p.ohp + l];
      }
      __syncthreads();
      for (int st = 0; st < p.n_stages; ++st) {
        const int R = p.radix[st];
        const int NR = N / R;
        const int2* tab = stab + st * N;
        int l = l_first, i = i_first;
        for (int e = tid; e < total; e += RG_T) {
          const int2 js = tab[i];
          const int step = js.y;
          int idx = 0;
          const cf* src = buf0 + l * N + js.x;
          cf acc = cf_make(0.f, 0.f);
          for (int r = 0; r < R; ++r) {
            const cf x = src[r * NR];
            const cf w = twsm[idx];                   // forward table: inverse transform uses conj(w)
            acc = pk_fma(mul_i<false>(x), bc(w.y), pk_fma(x, bc(w.x), acc));
            idx += step;
            if (idx >= N) idx -= N;
          }
          buf1[e] = acc;
          l += step_l; i += step_i;
          if (i >= N) { i -= N; ++l; }
        }
        __syncthreads();
        cf* tswap = buf0; buf0 = buf1; buf1 = tswap;
      }
      // |X|^2 of the kept columns (fftshift + crop) into the coil accumulators; element e always belongs to the same thread
      {
        int l = l_first, m = i_first;
        for (int e = tid; e < total; e += RG_T) {
          const int cc = phys_of_logical(m, N) - p.col0;
          if (cc >= 0 && cc < p.ow) accsm[l * opitch + cc] = cnorm2_acc(buf0[e], accsm[l * opitch + cc]);
          l += step_l; m += step_i;
          if (m >= N) { m -= N; ++l; }
        }
      }
      if (++coil == p.C) {
        coil = 0;
        int l = l_first, m = i_first;
        for (int e = tid; e < total; e += RG_T) {
          const int cc = phys_of_logical(m, N) - p.col0;
          if (cc >= 0 && cc < p.ow) {
            avsm[l * opitch + cc] += sqrtf(accsm[l * opitch + cc]) * p.scale;
            accsm[l * opitch + cc] = 0.f;
          }
          l += step_l; m += step_i;
          if (m >= N) { m -= N; ++l; }
        }
      }
      __syncthreads();
    }

    const float inv_a = 1.0f / (float)p.A;
    const int n_here = nl * p.ow;
    float* dst = p.out + ((long long)s * p.oh + tile * L) * p.ow;
    float lsum = 0.f;
    for (int e = tid; e < n_here; e += RG_T) {
      const int rr = e / p.ow, cc = e - rr * p.ow;
      const float v = avsm[rr * opitch + cc] * inv_a;
      dst[e] = v;
      lsum += v;
    }
    if (p.partials) {
      const float mean = rp_block_sum<RG_T / 32>(lsum, red) / (float)n_here;
      float lq = 0.f;
      for (int e = tid; e < n_here; e += RG_T) {
        const int rr = e / p.ow, cc = e - rr * p.ow;
        const float d = avsm[rr * opitch + cc] * inv_a - mean;
        lq = fmaf(d, d, lq);
      }
      const float m2 = rp_block_sum<RG_T / 32>(lq, red);
      if (tid == 0) {
        float* q = p.partials + ((long long)s * p.n_tiles + tile) * 3;
        q[0] = (float)n_here; q[1] = mean; q[2] = m2;
      }
    }
    __syncthreads();
  }
}

}  // namespace mriacl
