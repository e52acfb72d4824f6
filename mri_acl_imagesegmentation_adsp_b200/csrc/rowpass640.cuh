// rowpass640.cuh -- fused row pass for a (padded) phase-encode length of 640: the prostate-shape plans
// (ZIP!/fastmri_prostate/reconstruction/t2/prostate_t2_recon.py:65-75: PE 451 zero-padded to 640, RSS, flipud,
// mean over averages, crop 320 x 320) and every other 640 x 640 plan.
//
// One work item = (slice, 8 output rows); one CTA of 160 threads = the transform team of colpass640.cuh, and the
// same three-pass 8 x 8 x 10 in-place shared-memory FFT, here along the phase-encode axis of 8 rows at a time:
//   stage   the frame's [n_act][8 rows] block of the intermediate T (what the column pass wrote for the SAMPLED
//           columns only) is copied with cp.async one frame ahead (two tiny buffers);
//   pass 1  "expanding" radix-8: butterfly position pos = n mod 80 gathers its sampled inputs n = 80 n1 + pos from
//           the staged block through the host plan -- all eight (a regular butterfly), a few (direct sum with w8
//           powers) or none: then NOTHING is loaded or stored, because
//   pass 2  (radix-8 over n2) knows from the same plan which of its eight inputs are empty positions and takes
//           them as zero without touching shared memory -- the zero padding and the unsampled columns cost neither
//           a memset nor loads nor stores; pass 3 is the radix-10 over n3 of the column pass;
//   pass 3 does not store: each thread adds |X|^2 of its kept (fftshift + crop) columns to registers that live
//           across the coil loop; at the end of an average sqrt(.) / sqrt(HW) is added to a [8][ow] tile in shared
//           memory (mean over averages AFTER the RSS), and the tile leaves with coalesced stores plus its
//           (n, mean, M2) statistics for the instance normalisation.
// The row flip and the row crop were already applied by the column pass when it wrote T.
#pragma once
#include "colpass640.cuh"
#include "rowpass.cuh"

namespace mriacl {

constexpr int R640_ROWS = CP_G;          // 8 rows per item = the 8 "columns" of the transform buffer
constexpr int R640_T = CP_T;             // 160 threads

struct Row640Params {
  const cf* T;                 // [n_slices * A * C][n_act][ohp]
  int n_act, oh, ohp;
  const int* pos_off;          // [81] entries of butterfly position pos are ent[pos_off[pos] .. pos_off[pos + 1])
  const int* ent;              // per entry: n1 | (j << 3)
  int n_ent;
  const int* perm;             // [160] first-pass position of thread tid (plan.h: kinds of butterfly grouped per warp)
  const int* upos;             // [n_upos] non-empty first-pass positions, most entries first (plan.h)
  int n_upos;                  // 0: fixed position per thread (dense plans); > 0: balanced first pass over (position, line) units
  const cf* tw;                // w640^k = exp(+2 pi i k / 640)
  float* out;                  // [n_slices][oh][ow]
  float* partials;             // [n_slices][n_tiles][3] or nullptr
  int ow, col0;
  int A, C;
  float scale;                 // 1 / sqrt(H * 640)
  int n_slices, n_tiles;       // n_tiles = ceil(oh / 8)
};

__host__ __device__ inline int row640_smem_bytes(int n_act, int n_ent, int ow, int n_upos = 0) {
  return CP_BUF * 8 + 2 * n_act * R640_ROWS * 8 + R640_ROWS * (ow + 1) * 4 + ((81 + n_ent + 3) / 4) * 16 + 8 * 8 +   // (perm is read once from global)
         n_upos * (8 * 8 + 16);      // balanced first pass: per non-empty position 8 twiddles + (pos, e0, e1, -)
}

struct __align__(16) R640Unit { int pos, e0, e1, pad; };     // one non-empty first-pass position: entries ent[e0 .. e1)

template <int NW> __device__ __forceinline__ float r640_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) t += red[w];
  return t;
}

// BAL: balanced first pass over (non-empty position, line) units (p.n_upos > 0); otherwise a fixed position per thread
template <bool BAL>
__global__ void __launch_bounds__(R640_T, 3) rowpass640_kernel(Row640Params p) {
  MRIACL_DYN_SMEM(unsigned char, smem_raw);
  cf* buf = reinterpret_cast<cf*>(smem_raw);                                  // transform buffer, 8 lines
  cf* stage = buf + CP_BUF;                                                   // [2][n_act][8]
  float* avsm = reinterpret_cast<float*>(stage + 2 * (size_t)p.n_act * R640_ROWS);   // [8][ow + 1]
  int* plan = reinterpret_cast<int*>(avsm + R640_ROWS * (p.ow + 1));          // pos_off[81] | ent[n_ent]
  cf* w8sm = reinterpret_cast<cf*>(plan + ((81 + p.n_ent + 3) / 4) * 4);      // w8^k, k = 0..7
  cf* utw = w8sm + 8;                                                         // [n_upos][8]: w640^{pos m}, m = 0..7
  R640Unit* uinfo = reinterpret_cast<R640Unit*>(utw + 8 * (size_t)p.n_upos);  // [n_upos]
  __shared__ float red[R640_T / 32];

  const int tid = threadIdx.x;
  const int sub = tid / 80, pos = tid - sub * 80;
  for (int i = tid; i < 81; i += R640_T) plan[i] = p.pos_off[i];
  for (int i = tid; i < p.n_ent; i += R640_T) plan[81 + i] = p.ent[i];
  if (tid < 8) w8sm[tid] = p.tw[80 * tid];
  for (int i = tid; i < 8 * p.n_upos; i += R640_T) utw[i] = p.tw[(p.upos[i >> 3] * (i & 7)) % CP_N];
  for (int i = tid; i < p.n_upos; i += R640_T) {
    const int ps = p.upos[i];
    uinfo[i] = R640Unit{ps, p.pos_off[ps], p.pos_off[ps + 1], 0};
  }
  const int n_units = p.n_upos * R640_ROWS;
  cf tw1[8], tw2[8];
  const int pos1 = BAL ? 0 : p.perm[tid];                      // first pass: my butterfly position (grouped by kind)
  const int base2 = (pos / 10) * CP_BLK + (pos % 10);          // second pass: (m1, n3) = (pos / 10, pos % 10)
  {
    const int n3 = pos % 10;
#pragma unroll
    for (int m = 1; m < 8; ++m) {
      if (!BAL) tw1[m] = p.tw[(pos1 * m) % CP_N];
      tw2[m] = p.tw[(8 * n3 * m) % CP_N];
    }
  }
  // pass 3: thread r3 < 64 owns (m1, m2) = (r3 % 8, r3 / 8); outputs m = r3 + 64 m3 land on fixed image columns
  const int sub3 = tid / 64, r3 = tid - sub3 * 64;
  const int base3 = (r3 % 8) * CP_BLK + (r3 / 8) * 10;
  int cc3[10];
#pragma unroll
  for (int m3 = 0; m3 < 10; ++m3) {
    const int cc = phys_of_logical(r3 + 64 * m3, CP_N) - p.col0;
    cc3[m3] = (cc >= 0 && cc < p.ow) ? cc : -1;
  }
  __syncthreads();
  const int e0 = plan[pos1], e1 = plan[pos1 + 1], cnt = e1 - e0;
  // pass 2 reads positions 10 n2 + (pos % 10), n2 = 0..7: bit n2 of present2 = that position has sampled inputs
  int present2 = 0;
#pragma unroll
  for (int n2 = 0; n2 < 8; ++n2) present2 |= (plan[10 * n2 + pos % 10 + 1] > plan[10 * n2 + pos % 10]) ? (1 << n2) : 0;
  const int* ent = plan + 81;
  const int opitch = p.ow + 1;
  const int n_frames = p.A * p.C;
  const long long frame_elems = (long long)p.n_act * p.ohp;
  const int tile_elems = p.n_act * R640_ROWS;
  const int n_copies = p.n_act * (R640_ROWS / 2);        // 16-byte pieces: 4 per column
  const int n_items = p.n_slices * p.n_tiles;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int s = item / p.n_tiles, tile = item - s * p.n_tiles;
    const cf* Tit = p.T + (long long)s * n_frames * frame_elems + tile * R640_ROWS;
    auto prefetch = [&](int f, int b) {
      const cf* src = Tit + (long long)f * frame_elems;
      cf* dst = stage + (size_t)b * tile_elems;
      for (int i = tid; i < n_copies; i += R640_T) {
        const int j = i >> 2, part = i & 3;
        cp_async16(dst + j * R640_ROWS + 2 * part, src + (long long)j * p.ohp + 2 * part);
      }
      cp_async_commit();
    };
    for (int i = tid; i < R640_ROWS * opitch; i += R640_T) avsm[i] = 0.f;
    prefetch(0, 0);

    float acc[4][10];
    int coil = 0;
    for (int f = 0; f < n_frames; ++f) {
      if (coil == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int m3 = 0; m3 < 10; ++m3) acc[c][m3] = 0.f;
      }
      if (f + 1 < n_frames) { prefetch(f + 1, (f + 1) & 1); cp_async_wait<1>(); } else cp_async_wait<0>();
      __syncthreads();       // block f visible; pass 3 of frame f-1 has finished reading the transform buffer
      const cf* st = stage + (size_t)(f & 1) * tile_elems;

      // ---- pass 1: expanding radix-8 over n1 (stride 90), twiddle w640^{pos * m1} ----
      if (BAL) {
        // balanced: (non-empty position, line) units dealt to all threads; the unit's twiddles come from shared memory
        for (int u = tid; u < n_units; u += R640_T) {
          const int q = u >> 3, kc = u & 7;
          const R640Unit inf = uinfo[q];
          const cf* twq = utw + 8 * q;
          cf* col = buf + kc * CP_PITCH + inf.pos;
          if (inf.e1 - inf.e0 == 8) {
            cf v[8];
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) v[n1] = st[(ent[inf.e0 + n1] >> 3) * R640_ROWS + kc];
            radix8<true>(v);
            col[0] = v[0];
#pragma unroll
            for (int m1 = 1; m1 < 8; ++m1) col[m1 * CP_BLK] = cmul(v[m1], twq[m1]);
          } else {
            cf y[8];
#pragma unroll
            for (int m1 = 0; m1 < 8; ++m1) y[m1] = cf_make(0.f, 0.f);
            for (int e = inf.e0; e < inf.e1; ++e) {
              const int n1 = ent[e] & 7;
              const cf x = st[(ent[e] >> 3) * R640_ROWS + kc];
              y[0] = cadd(y[0], x);
#pragma unroll
              for (int m1 = 1; m1 < 8; ++m1) {
                const cf w = w8sm[(n1 * m1) & 7];
                y[m1] = pk_fma(mul_i<true>(x), bc(w.y), pk_fma(x, bc(w.x), y[m1]));
              }
            }
            col[0] = y[0];
#pragma unroll
            for (int m1 = 1; m1 < 8; ++m1) col[m1 * CP_BLK] = cmul(y[m1], twq[m1]);
          }
        }
      } else
      for (int kc = sub; kc < R640_ROWS; kc += 2) {
        cf* col = buf + kc * CP_PITCH + pos1;
        if (cnt == 0) {
          // empty position: pass 2 substitutes zeros (present2 below), nothing to do
        } else if (cnt == 8) {
          cf v[8];
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) v[n1] = st[(ent[e0 + n1] >> 3) * R640_ROWS + kc];   // entries are sorted by n1
          radix8<true>(v);
          col[0] = v[0];
#pragma unroll
          for (int m1 = 1; m1 < 8; ++m1) col[m1 * CP_BLK] = cmul(v[m1], tw1[m1]);
        } else {
          cf y[8];
#pragma unroll
          for (int m1 = 0; m1 < 8; ++m1) y[m1] = cf_make(0.f, 0.f);
          for (int e = e0; e < e1; ++e) {
            const int n1 = ent[e] & 7;
            const cf x = st[(ent[e] >> 3) * R640_ROWS + kc];
            y[0] = cadd(y[0], x);
#pragma unroll
            for (int m1 = 1; m1 < 8; ++m1) {
              const cf w = w8sm[(n1 * m1) & 7];
              y[m1] = pk_fma(mul_i<true>(x), bc(w.y), pk_fma(x, bc(w.x), y[m1]));
            }
          }
          col[0] = y[0];
#pragma unroll
          for (int m1 = 1; m1 < 8; ++m1) col[m1 * CP_BLK] = cmul(y[m1], tw1[m1]);
        }
      }
      __syncthreads();

      // ---- pass 2: radix-8 over n2 (stride 10), twiddle w80^{n3 * m2} ----
      for (int kc = sub; kc < R640_ROWS; kc += 4) {
        cf* colA = buf + kc * CP_PITCH + base2;
        cf* colB = colA + 2 * CP_PITCH;
        cf va[8], vb[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) {
          const bool has = (present2 >> n2) & 1;
          va[n2] = has ? colA[n2 * 10] : cf_make(0.f, 0.f);
          vb[n2] = has ? colB[n2 * 10] : cf_make(0.f, 0.f);
        }
        radix8<true>(va);
        radix8<true>(vb);
        colA[0] = va[0];
        colB[0] = vb[0];
#pragma unroll
        for (int m2 = 1; m2 < 8; ++m2) { colA[m2 * 10] = cmul(va[m2], tw2[m2]); colB[m2 * 10] = cmul(vb[m2], tw2[m2]); }
      }
      __syncthreads();

      // ---- pass 3: radix-10 over n3, |X|^2 of the kept columns into the registers ----
      if (tid < 128) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4* col4 = reinterpret_cast<const float4*>(buf + (sub3 + 2 * c) * CP_PITCH + base3);
          cf v[10];
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const float4 t = col4[q];
            v[2 * q] = cf_make(t.x, t.y);
            v[2 * q + 1] = cf_make(t.z, t.w);
          }
          radix10<true>(v);
#pragma unroll
          for (int m3 = 0; m3 < 10; ++m3) acc[c][m3] = cnorm2_acc(v[m3], acc[c][m3]);
        }
      }
      if (++coil == p.C) {      // end of an average: RSS of this average into the running tile (private slots)
        coil = 0;
        if (tid < 128) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int m3 = 0; m3 < 10; ++m3)
              if (cc3[m3] >= 0) avsm[(sub3 + 2 * c) * opitch + cc3[m3]] += sqrtf(acc[c][m3]) * p.scale;
        }
      }
    }
    __syncthreads();

    // ---- write the tile (coalesced) and its statistics ----
    const float inv_a = 1.0f / (float)p.A;
    const int rows_here = min(R640_ROWS, p.oh - tile * R640_ROWS);
    const int n_here = rows_here * p.ow;
    float* dst = p.out + ((long long)s * p.oh + tile * R640_ROWS) * p.ow;
    float lsum = 0.f;
    for (int e = tid; e < n_here; e += R640_T) {
      const int rr = e / p.ow, cc = e - rr * p.ow;
      const float v = avsm[rr * opitch + cc] * inv_a;
      dst[e] = v;
      lsum += v;
    }
    if (p.partials) {
      const float mean = r640_block_sum<R640_T / 32>(lsum, red) / (float)n_here;
      float lq = 0.f;
      for (int e = tid; e < n_here; e += R640_T) {
        const int rr = e / p.ow, cc = e - rr * p.ow;
        const float d = avsm[rr * opitch + cc] * inv_a - mean;
        lq = fmaf(d, d, lq);
      }
      const float m2 = r640_block_sum<R640_T / 32>(lq, red);
      if (tid == 0) {
        float* q = p.partials + ((long long)s * p.n_tiles + tile) * 3;
        q[0] = (float)n_here; q[1] = mean; q[2] = m2;
      }
    }
    __syncthreads();
  }
}

}  // namespace mriacl
