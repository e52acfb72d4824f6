// rowpass.cuh -- fused row pass: pruned P x Q inverse DFT along the (padded) phase-encode axis,
// |.|^2 accumulation over coils, sqrt, mean over averages, fftshift + centre crop as the store
// index map, and per-tile statistics for the instance normalisation.
//
// One work item = (slice, tile of 32 output rows); lane = row, so every lane of a warp runs the
// same plan-driven control flow and every shared-memory access is stride-1 across lanes.
// For each coil frame:
//   stage 1  (decimation in time, n = Q n1 + n2):  for each residue n2 the P-point DFT over n1 of
//            the SAMPLED columns only.  The host plan classifies each residue as dense (symmetric
//            direct DFT with immediate constants) or sparse (<= 6 sampled columns: direct
//            accumulation with w_N^{n k1}), and balances the units over the 8 warps.  Inputs come
//            straight from the intermediate T[frame][j][row] (coalesced, lane = row).
//   stage 2  for each k1 the Q-point FFT over n2 in registers -> X[k1 + P k2]; |X|^2 is added to
//            per-thread accumulators that stay in registers across all coils.
// Width 368 = 23 x 16 is the knee case; the template is general in (P odd, Q = 16).
#pragma once
#include "butterflies.cuh"

namespace mriacl {

constexpr int RP_T = 256;          // threads per CTA
constexpr int RP_NW = 8;           // warps
constexpr int RP_ROWS = 32;        // rows per tile (= lanes)
constexpr int RP_MAX_SPARSE = 6;   // a residue with more sampled columns than this is "dense"

struct RowPassParams {
  const cf* T;           // [n_slices*A*C][n_act][oh]
  int n_act, oh;
  const int* sched;      // warp schedule, see plan.h
  const cf* tw;          // w_N^k = exp(+2 pi i k / N)
  float* out;            // [n_slices][oh][ow]  (already offset to the first slice of the launch)
  float* partials;       // [n_slices][n_tiles][3] (count, mean, M2) or nullptr
  int ow, col0;
  int A, C;
  float scale;           // 1 / sqrt(H * N)
  int n_slices, n_tiles;
};

template <int P, int Q> constexpr int rowpass_smem_bytes(int ow, int A) {
  return P * Q * RP_ROWS * 8 + P * Q * 8 + (A > 1 ? RP_ROWS * (ow + 1) * 4 : 0);
}

__device__ __forceinline__ float rp_block_sum(float v, float* red /* 9 floats */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < RP_NW; ++w) t += red[w];
  return t;
}

template <int P, int Q>
__global__ void __launch_bounds__(RP_T, 2) rowpass_kernel(RowPassParams p) {
  static_assert(Q == 16, "stage 2 is the register-level 16-point FFT");
  constexpr int N = P * Q;
  constexpr int KPW = (P + RP_NW - 1) / RP_NW;
  MRIACL_DYN_SMEM(cf, Y);                       // [P][Q][32]
  cf* twsm = Y + N * RP_ROWS;                   // [N]
  float* osm = reinterpret_cast<float*>(Y);     // output tile [32][ow+1], aliases Y
  float* avsm = reinterpret_cast<float*>(twsm + N);  // running sum over averages (A > 1 only)
  __shared__ float red[RP_NW + 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int opitch = p.ow + 1;
  for (int i = tid; i < N; i += RP_T) twsm[i] = p.tw[i];

  const int my_off = p.sched[warp];
  const int n_items = p.n_slices * p.n_tiles;
  const long long frame_elems = (long long)p.n_act * p.oh;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int s = item / p.n_tiles, tile = item - s * p.n_tiles;
    const int row = tile * RP_ROWS + lane;
    const bool rvalid = row < p.oh;
    if (p.A > 1) for (int i = tid; i < RP_ROWS * opitch; i += RP_T) avsm[i] = 0.f;
    __syncthreads();   // twsm / avsm ready; previous item's tile fully consumed

    for (int a = 0; a < p.A; ++a) {
      float acc[KPW][Q];
#pragma unroll
      for (int kk = 0; kk < KPW; ++kk)
#pragma unroll
        for (int k2 = 0; k2 < Q; ++k2) acc[kk][k2] = 0.f;

      for (int c = 0; c < p.C; ++c) {
        const cf* Tf = p.T + ((long long)(s * p.A + a) * p.C + c) * frame_elems + (rvalid ? row : 0);

        // ---------------- stage 1: pruned P-point DFTs of my residues ----------------
        const int n_units = p.sched[my_off];
        int off = my_off + 1;
        for (int u = 0; u < n_units; ++u) {
          const int n2 = p.sched[off], type = p.sched[off + 1], nnz = p.sched[off + 2];
          off += 3;
          cf* ycol = Y + n2 * RP_ROWS + lane;          // + k1 * Q * 32
          if (type == 1) {
            cf x[P];
#pragma unroll
            for (int n1 = 0; n1 < P; ++n1) {
              const int j = p.sched[off + n1];
              x[n1] = (j >= 0 && rvalid) ? Tf[(long long)j * p.oh] : cf_make(0.f, 0.f);
            }
            off += P;
            dft_odd_sym<P, true>(x, [&](auto kc, cf val) {
              constexpr int k1 = decltype(kc)::value;
              if (k1 != 0) val = cmul(val, twsm[(n2 * k1) % N]);
              ycol[k1 * Q * RP_ROWS] = val;
            });
          } else {
            cf xe[RP_MAX_SPARSE];
            int ne[RP_MAX_SPARSE], idx[RP_MAX_SPARSE];
#pragma unroll
            for (int e = 0; e < RP_MAX_SPARSE; ++e) {
              xe[e] = cf_make(0.f, 0.f); ne[e] = 0; idx[e] = 0;
              if (e < nnz) {
                ne[e] = p.sched[off + 2 * e];
                const int j = p.sched[off + 2 * e + 1];
                if (rvalid) xe[e] = Tf[(long long)j * p.oh];
              }
            }
            off += 2 * nnz;
            for (int k1 = 0; k1 < P; ++k1) {
              float re = 0.f, im = 0.f;
#pragma unroll
              for (int e = 0; e < RP_MAX_SPARSE; ++e) {
                if (e < nnz) {
                  const cf w = twsm[idx[e]];
                  re = fmaf(xe[e].x, w.x, fmaf(-xe[e].y, w.y, re));
                  im = fmaf(xe[e].x, w.y, fmaf(xe[e].y, w.x, im));
                  idx[e] += ne[e];
                  if (idx[e] >= N) idx[e] -= N;
                }
              }
              ycol[k1 * Q * RP_ROWS] = cf_make(re, im);
            }
          }
        }
        __syncthreads();

        // ---------------- stage 2: Q-point FFT over n2, accumulate |X|^2 ----------------
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) {
          const int k1 = warp + RP_NW * kk;
          if (k1 < P) {
            cf v[Q];
            const cf* yrow = Y + k1 * Q * RP_ROWS + lane;
#pragma unroll
            for (int n2 = 0; n2 < Q; ++n2) v[n2] = yrow[n2 * RP_ROWS];
            fft16<true>(v);
#pragma unroll
            for (int k2 = 0; k2 < Q; ++k2) acc[kk][k2] = cnorm2_acc(v[k2], acc[kk][k2]);
          }
        }
        __syncthreads();
      }

      // ---------------- per-average epilogue: sqrt, shift + crop into the tile ----------------
#pragma unroll
      for (int kk = 0; kk < KPW; ++kk) {
        const int k1 = warp + RP_NW * kk;
        if (k1 < P) {
#pragma unroll
          for (int k2 = 0; k2 < Q; ++k2) {
            const int cc = phys_of_logical(k1 + P * k2, N) - p.col0;
            if (cc >= 0 && cc < p.ow) {
              const float v = sqrtf(acc[kk][k2]) * p.scale;
              if (p.A > 1) avsm[lane * opitch + cc] += v; else osm[lane * opitch + cc] = v;
            }
          }
        }
      }
      // (A > 1: avsm is private per (lane, cc) owner, no barrier needed between averages)
    }
    __syncthreads();

    // ---------------- write the tile (coalesced) and its statistics ----------------
    const float* tile_sm = p.A > 1 ? avsm : osm;
    const float inv_a = 1.0f / (float)p.A;
    const int rows_here = min(RP_ROWS, p.oh - tile * RP_ROWS);
    const int n_here = rows_here * p.ow;
    float* dst = p.out + ((long long)s * p.oh + tile * RP_ROWS) * p.ow;
    float lsum = 0.f;
    for (int e = tid; e < n_here; e += RP_T) {
      const int r = e / p.ow, cc = e - r * p.ow;
      float v = tile_sm[r * opitch + cc];
      if (p.A > 1) v *= inv_a;
      dst[e] = v;
      lsum += v;
    }
    if (p.partials) {
      const float mean = rp_block_sum(lsum, red) / (float)n_here;
      float lq = 0.f;
      for (int e = tid; e < n_here; e += RP_T) {
        const int r = e / p.ow, cc = e - r * p.ow;
        float v = tile_sm[r * opitch + cc];
        if (p.A > 1) v *= inv_a;
        const float d = v - mean;
        lq = fmaf(d, d, lq);
      }
      const float m2 = rp_block_sum(lq, red);
      if (tid == 0) {
        float* q = p.partials + ((long long)s * p.n_tiles + tile) * 3;
        q[0] = (float)n_here; q[1] = mean; q[2] = m2;
      }
    }
    // the loop-top barrier orders these tile reads before the next item's writes
  }
}

}  // namespace mriacl
