// rowpass16.cuh -- fused row pass on 16-row tiles: the two half-warps of a warp work on two different
// plan units (stage 1) / two different k1 (stage 2) of the SAME 16 rows, lane & 15 = row.
//
// Compared with the 32-row kernel in rowpass.cuh this halves the shared-memory tile (Y: 47 KB), halves
// the |X|^2 accumulators a thread keeps across the coil loop (16 instead of 32-48), doubles the number
// of work items (finer tail) and leaves room on the SM for column-pass CTAs (overlapped schedule).
// Every shared-memory access is still two fully used 128-byte segments per warp: no bank conflicts.
//
// Algorithm per coil frame (identical arithmetic to rowpass.cuh, so results are bit-identical):
//   prefetch [n_act][16 rows] of T with cp.async one frame ahead;
//   stage 1: pruned P-point DFTs of the sampled columns (dense residues: symmetric direct DFT, split in two
//            output halves; sparse residues: tabulated twiddles), two units per warp, paired by the host
//            plan so that both half-warps run the same code path;
//   stage 2: two 16-point register FFTs per warp (k1 = 2 pair, 2 pair + 1), accumulate |X|^2.
#pragma once
#include "rowpass.cuh"

namespace mriacl {

constexpr int RP16_ROWS = 16;

struct RowPass16Params {
  const cf* T;           // [n_slices*A*C][n_act][ohp]
  int n_act, oh, ohp;
  int n_slots;           // slots of the residue-major staged tile (>= n_act; slots without a column stay zero)
  const int* slot_of_j;  // [n_act]
  const int* sched;      // pair schedule, see plan.h (build_pair_schedule)
  int sched_len;
  const cf* sptw;        // sparse twiddle rows followed by the dense residues' rows w_N^{n2 k1} (pitch 24)
  int sptw_len;
  float* out;
  float* partials;       // [n_slices][n_tiles][3]
  int ow, col0;
  int A, C;
  float scale;
  int n_slices, n_tiles; // n_tiles = ceil(oh / 16)
  int n_buf;
  const int* done;
  int done_target;
  int* error_flag;
  // fused instance normalisation: the team that finishes the last tile of a slice normalises the slice
  int* tiles_done;       // [n_slices] zeroed before the launch, or nullptr (statistics / normalisation by a later launch)
  float* mean_std;       // [n_slices][2] or nullptr
  float eps;
  int normalize;         // 1: (x - mean) / (std + eps) in place
  int l2_hints;          // unused by this kernel (kept for the launch code; the hint applies to the column pass only)
  int debug_skip;        // profiling only (results are garbage): 1 = no stage 1, 2 = no stage 2, 4 = no T prefetch
  int reverse;           // 1: take the slices last-to-first: the column pass wrote them first-to-last, so the most
                         //    recently written part of T is still in L2 when the row pass starts (sequential schedule)
  int ring;              // co-resident schedule: T holds `ring` slices, slice s lives in slot s % ring (0: one slot per slice)
};

inline int rowpass16_smem_bytes(int P, int Q, int sptw_len, int sched_len, int n_slots, int n_buf, int ow, int A) {
  return P * (Q + 1) * RP16_ROWS * 8 + rp_round16(sptw_len * 8) + rp_round16(sched_len * 4) + rp_round16(n_slots * 4) +
         n_buf * (n_slots + 1) * RP16_ROWS * 8 + (A > 1 ? RP16_ROWS * (ow + 1) * 4 : 0);
}

// team barrier: BAR = 0 is the CTA barrier; BAR > 0 a named barrier over the NT threads of a sub-CTA team
// BAR = -1: several identical teams of NT threads in one CTA share ONE copy of the code (the row pass is ~100 KB of
// instructions: a copy per team would thrash the instruction cache); team k = threadIdx.x / NT uses named barrier 1 + k
template <int BAR, int NT> __device__ __forceinline__ void rp16_sync() {
  if constexpr (BAR == 0) __syncthreads();
  else if constexpr (BAR < 0) named_bar_sync(1 + (int)threadIdx.x / NT, NT);
  else named_bar_sync(BAR, NT);
}
template <int NW, int BAR> __device__ __forceinline__ float rp16_team_sum(float v, float* red /* NW floats */, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  rp16_sync<BAR, NW * 32>();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  rp16_sync<BAR, NW * 32>();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) t += red[w];
  return t;
}

template <int P, int Q, int NNZ>
__device__ __forceinline__ void rp16_sparse_unit(const int* pay, const cf* tb, const cf* sptw, cf* ycol) {
  constexpr int YS = (Q + 1) * RP16_ROWS;   // Y stride between k1
  constexpr int PITCH = (P + 2) & ~1;       // plan.h sptw_pitch(P)
  cf xe[NNZ > 0 ? NNZ : 1];
  const float4* tw4 = reinterpret_cast<const float4*>(sptw + pay[1]);
#pragma unroll
  for (int e = 0; e < NNZ; ++e) xe[e] = tb[pay[2 + e] * RP16_ROWS];
#pragma unroll
  for (int kp = 0; kp < PITCH / 2; ++kp) {
    cf y0 = cf_make(0.f, 0.f), y1 = cf_make(0.f, 0.f);
#pragma unroll
    for (int e = 0; e < NNZ; ++e) {
      const float4 w = tw4[e * (PITCH / 2) + kp];       // twiddles of k1 = 2 kp and 2 kp + 1
      const cf w0 = cf_make(w.x, w.y), w1 = cf_make(w.z, w.w);
      y0 = pk_fma(mul_i<true>(w0), bc(xe[e].y), pk_fma(w0, bc(xe[e].x), y0));
      y1 = pk_fma(mul_i<true>(w1), bc(xe[e].y), pk_fma(w1, bc(xe[e].x), y1));
    }
    ycol[(2 * kp) * YS] = y0;
    if (2 * kp + 1 < P) ycol[(2 * kp + 1) * YS] = y1;
  }
}

// shared-memory carve-up of one row-pass CTA
template <int P, int Q> struct Rp16Smem {
  cf* Y; cf* sptw; int* sch; int* slot; cf* tbuf; float* av; float* osm;
  __device__ __forceinline__ Rp16Smem(void* base, const RowPass16Params& p) {
    constexpr int YS = (Q + 1) * RP16_ROWS;
    Y = reinterpret_cast<cf*>(base);                                     // [P][Q + 1][16]
    sptw = Y + P * YS;
    sch = reinterpret_cast<int*>(reinterpret_cast<char*>(sptw) + rp_round16(p.sptw_len * 8));
    slot = reinterpret_cast<int*>(reinterpret_cast<char*>(sch) + rp_round16(p.sched_len * 4));  // [n_act] (room for n_slots)
    tbuf = reinterpret_cast<cf*>(reinterpret_cast<char*>(slot) + rp_round16(p.n_slots * 4));    // [n_buf][n_slots + 1][16], residue-major
    av = reinterpret_cast<float*>(tbuf + (size_t)p.n_buf * (p.n_slots + 1) * RP16_ROWS);
    osm = reinterpret_cast<float*>(Y);                                   // output tile [16][ow+1], aliases Y
  }
};

// copy the plan tables into shared memory (all NT threads; followed by a barrier at the caller)
template <int NT> __device__ __forceinline__ void rp16_load_tables(const RowPass16Params& p, cf* sptwsm, int* schsm, cf* tbuf, int tid) {
  for (int i = tid; i < p.sched_len; i += NT) schsm[i] = p.sched[i];
  for (int i = tid; i < p.sptw_len; i += NT) sptwsm[i] = p.sptw[i];
  int* slotsm = reinterpret_cast<int*>(reinterpret_cast<char*>(schsm) + rp_round16(p.sched_len * 4));
  for (int i = tid; i < p.n_act; i += NT) slotsm[i] = p.slot_of_j[i];
  // slots no column fills (unsampled positions of dense residues, the spare column) must read as zero: the
  // prefetch never writes them, so clearing the staging buffers once is enough
  for (int i = tid; i < p.n_buf * (p.n_slots + 1) * RP16_ROWS; i += NT) tbuf[i] = cf_make(0.f, 0.f);
}

// one work item = (slice, 16-row tile); called by all NW*32 threads of the (sub-)CTA; tables already loaded
template <int P, int Q, int NW, int BAR = 0>
__device__ __forceinline__ void rowpass16_item(const RowPass16Params& p, void* smem_base, int item, int tid,
                                               float* red, int* ready_flag) {
  constexpr int N = P * Q;
  constexpr int NT = NW * 32;
  constexpr int NPAIR = (P + 1) / 2;                 // stage-2 pairs of k1
  constexpr int KPW = (NPAIR + NW - 1) / NW;
  constexpr int HP = (P - 1) / 2;
  // dense residues are shared by three warps-halves: output pairs [1, B1) (+ X0), [B1, B2), [B2, HP]
  constexpr int B1 = 1 + HP / 3, B2 = B1 + (HP - HP / 3 + 1) / 2;
  constexpr int YS = (Q + 1) * RP16_ROWS;            // residue column Q is a write-only spare for idle half-warps
  Rp16Smem<P, Q> S(smem_base, p);
  cf* Y = S.Y; cf* sptwsm = S.sptw; int* schsm = S.sch; cf* tbuf = S.tbuf; float* avsm = S.av; float* osm = S.osm;
  int& ready = *ready_flag;

  const int lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, r = lane & 15;
  const int opitch = p.ow + 1;
  const int my_off = schsm[warp];
  const int n_frames = p.A * p.C;
  const long long frame_elems = (long long)p.n_act * p.ohp;
  const int tile_elems = (p.n_slots + 1) * RP16_ROWS;
  const int* slotsm = S.slot;
  const int n_copies = p.n_act * (RP16_ROWS / 2);    // 16-byte copies per block (8 per column)

  {

    const int s_fwd = item / p.n_tiles, tile = item - s_fwd * p.n_tiles;
    const int s = p.reverse ? p.n_slices - 1 - s_fwd : s_fwd;
    const cf* Tit = p.T + (long long)(p.ring ? s % p.ring : s) * n_frames * frame_elems + tile * RP16_ROWS;

    const cf* src0 = Tit + (long long)(tid >> 3) * p.ohp + 2 * (tid & 7);
    cf* dst0 = tbuf + 2 * (tid & 7);
    const long long src_step = (long long)(NT / 8) * p.ohp;
    const int n_iter = tid < n_copies ? (n_copies - tid + NT - 1) / NT : 0;
    auto prefetch = [&](int f, int buf) {
      const cf* src = src0 + (long long)f * frame_elems;
      cf* dst = dst0 + (size_t)buf * tile_elems;
      int j = tid >> 3;
      if (!MRIACL_DBG_SKIP(p, 4)) {
        // (no L2 hint on these copies: ptxas 12.9 encodes a hinted LDGSTS with a uniform-register address offset
        //  that sm_100a rejects as an illegal instruction, and the hints did not pay off anyway)
#pragma unroll 1
        for (int i = 0; i < n_iter; ++i) {
          cp_async16(dst + slotsm[j] * RP16_ROWS, src);
          src += src_step;
          j += NT / 8;
        }
      }
      cp_async_commit();
    };

    if (p.done) {
      if (tid == 0) {
        ready = rp_wait_count(p.done + s, p.done_target, p.error_flag) ? 1 : 0;
        if (!ready && p.error_flag) atomicAdd(p.error_flag, 1);
      }
      rp16_sync<BAR, NT>();
      if (!ready) return;
    }
    if (p.A > 1) for (int i = tid; i < RP16_ROWS * opitch; i += NT) avsm[i] = 0.f;
    prefetch(0, 0);
    if (p.n_buf == 3) { if (n_frames > 1) prefetch(1, 1); else cp_async_commit(); }

    float acc[KPW][Q];
    int coil = 0;
    for (int f = 0; f < n_frames; ++f) {
      const int buf = p.n_buf == 3 ? f % 3 : p.n_buf == 2 ? (f & 1) : 0;
      if (coil == 0) {
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk)
#pragma unroll
          for (int k2 = 0; k2 < Q; ++k2) acc[kk][k2] = 0.f;
      }
      if (p.n_buf == 3) {
        // two frames ahead: the block of frame f+2 goes where frame f-1 was (all its readers passed the stage-1 barrier);
        // one (possibly empty) group is committed per frame, so "at most 2 pending" means frame f has landed
        if (f + 2 < n_frames) prefetch(f + 2, (f + 2) % 3); else cp_async_commit();
        cp_async_wait<2>();
      } else if (p.n_buf == 2) {
        if (f + 1 < n_frames) { prefetch(f + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
      } else {
        cp_async_wait<0>();
      }
      rp16_sync<BAR, NT>();

      // ---------------- stage 1: two units per warp ----------------
      {
        const cf* tb = tbuf + (size_t)buf * tile_elems + r;
        const int n_pairs = MRIACL_DBG_SKIP(p, 1) ? 0 : schsm[my_off];
        int off = my_off + 1;
        for (int u = 0; u < n_pairs; ++u, off += 4) {
          const int type = schsm[off], nnz = schsm[off + 1];
          const int offB = schsm[off + 3];
          const int* pay = schsm + ((half == 0 || offB < 0) ? schsm[off + 2] : offB);
          const int n2 = (half == 1 && offB < 0) ? Q : pay[0];     // idle half: spare column
          cf* ycol = Y + n2 * RP16_ROWS + r;                       // + k1 * YS
          if (type != 0) {
            // operands are read where the streaming DFT consumes them (immediate offsets from the residue's base slot)
            struct { const cf* p; __device__ __forceinline__ cf operator[](int n1) const { return p[n1 * RP16_ROWS]; } } x{tb + pay[2] * RP16_ROWS};
            const cf* dtw = sptwsm + pay[1];
            auto emit = [&](auto kc, cf val) {
              constexpr int k1 = decltype(kc)::value;
              if (k1 != 0) val = cmul(val, dtw[k1]);
              ycol[k1 * YS] = val;
            };
            if (type == 2) dft_odd_sym_part<P, true, 1, B1, true>(x, emit);
            else if (type == 3) dft_odd_sym_part<P, true, B1, B2, false>(x, emit);
            else dft_odd_sym_part<P, true, B2, HP + 1, false>(x, emit);
          } else {
            switch (nnz) {
              case 0: rp16_sparse_unit<P, Q, 0>(pay, tb, sptwsm, ycol); break;
              case 1: rp16_sparse_unit<P, Q, 1>(pay, tb, sptwsm, ycol); break;
              case 2: rp16_sparse_unit<P, Q, 2>(pay, tb, sptwsm, ycol); break;
              case 3: rp16_sparse_unit<P, Q, 3>(pay, tb, sptwsm, ycol); break;
              case 4: rp16_sparse_unit<P, Q, 4>(pay, tb, sptwsm, ycol); break;
              case 5: rp16_sparse_unit<P, Q, 5>(pay, tb, sptwsm, ycol); break;
              default: rp16_sparse_unit<P, Q, 6>(pay, tb, sptwsm, ycol); break;
            }
          }
        }
      }
      rp16_sync<BAR, NT>();
      if (p.n_buf == 1 && f + 1 < n_frames) prefetch(f + 1, 0);

      // ---------------- stage 2: two 16-point FFTs per warp ----------------
#pragma unroll
      for (int kk = 0; kk < KPW; ++kk) {
        const int pair = warp + NW * kk;
        const int k1 = 2 * pair + half;
        if (pair < NPAIR && !MRIACL_DBG_SKIP(p, 2)) {
          const int k1c = k1 < P ? k1 : P - 1;           // odd P: the last pair has one idle half
          cf v[Q];
          const cf* yrow = Y + k1c * YS + r;
#pragma unroll
          for (int n2 = 0; n2 < Q; ++n2) v[n2] = yrow[n2 * RP16_ROWS];
          fft_q<Q, true>(v);
#pragma unroll
          for (int k2 = 0; k2 < Q; ++k2) acc[kk][k2] = cnorm2_acc(v[k2], acc[kk][k2]);
        }
      }

      if (++coil == p.C) {
        coil = 0;
        if (p.A > 1) {
#pragma unroll
          for (int kk = 0; kk < KPW; ++kk) {
            const int pair = warp + NW * kk;
            const int k1 = 2 * pair + half;
            if (pair < NPAIR && k1 < P) {
#pragma unroll
              for (int k2 = 0; k2 < Q; ++k2) {
                const int cc = phys_of_logical(k1 + P * k2, N) - p.col0;
                if (cc >= 0 && cc < p.ow) avsm[r * opitch + cc] += sqrtf(acc[kk][k2]) * p.scale;
              }
            }
          }
        }
      }
    }
    rp16_sync<BAR, NT>();
    if (MRIACL_DBG_SKIP(p, 8)) return;
    if (p.A == 1) {
#pragma unroll
      for (int kk = 0; kk < KPW; ++kk) {
        const int pair = warp + NW * kk;
        const int k1 = 2 * pair + half;
        if (pair < NPAIR && k1 < P) {
#pragma unroll
          for (int k2 = 0; k2 < Q; ++k2) {
            const int cc = phys_of_logical(k1 + P * k2, N) - p.col0;
            if (cc >= 0 && cc < p.ow) osm[r * opitch + cc] = sqrtf(acc[kk][k2]) * p.scale;
          }
        }
      }
      rp16_sync<BAR, NT>();
    }

    const float* tile_sm = p.A > 1 ? avsm : osm;
    const float inv_a = 1.0f / (float)p.A;
    const int rows_here = min(RP16_ROWS, p.oh - tile * RP16_ROWS);
    const int n_here = rows_here * p.ow;
    float* dst = p.out + ((long long)s * p.oh + tile * RP16_ROWS) * p.ow;
    // each thread keeps its share of the tile (elements tid, tid + NT, ...) in registers for the statistics pass;
    // (row, column) advance incrementally: no integer division in the loops
    constexpr int VMAX = (RP16_ROWS * N + NT - 1) / NT;
    float vals[VMAX];
    const int step_r = NT / p.ow, step_c = NT - step_r * p.ow;
    int rr = tid / p.ow, cc = tid - rr * p.ow;
    float lsum = 0.f;
#pragma unroll
    for (int i = 0; i < VMAX; ++i) {
      const int e = tid + i * NT;
      float v = 0.f;
      if (e < n_here) {
        v = tile_sm[rr * opitch + cc];
        if (p.A > 1) v *= inv_a;
        dst[e] = v;
        lsum += v;
      }
      vals[i] = v;
      rr += step_r; cc += step_c;
      if (cc >= p.ow) { cc -= p.ow; ++rr; }
    }
    if (p.partials) {
      const float mean = rp16_team_sum<NW, BAR>(lsum, red, tid) / (float)n_here;
      float lq = 0.f;
#pragma unroll
      for (int i = 0; i < VMAX; ++i) {
        const float d = vals[i] - mean;
        if (tid + i * NT < n_here) lq = fmaf(d, d, lq);
      }
      const float m2 = rp16_team_sum<NW, BAR>(lq, red, tid);
      if (tid == 0) {
        float* q = p.partials + ((long long)s * p.n_tiles + tile) * 3;
        q[0] = (float)n_here; q[1] = mean; q[2] = m2;
      }
    }
    rp16_sync<BAR, NT>();
  }
}

__device__ __forceinline__ float rp16_ld_cg(const float* p) {
#if defined(MRIACL_EMU)
  return *p;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ float4 rp16_ld_cg4(const float4* p) {
#if defined(MRIACL_EMU)
  return *p;
#else
  return __ldcg(p);
#endif
}

// Called by the whole team after rowpass16_item when p.tiles_done is set: publishes the tile, and if it was the
// last tile of its slice, merges the tiles' (n, mean, M2) partials (Chan) into mean / unbiased std
// (ZIP!/DL_reconstruction/data/transforms.py:143-162) and normalises the slice in place.  Tiles of other CTAs are
// read with ld.cg (L2): they were published with a device-wide fence before the counter moved.
template <int NW, int BAR>
__device__ __forceinline__ void rowpass16_finish_slice(const RowPass16Params& r, int item, int t, float* s_stat, int* s_last) {
  constexpr int NT = NW * 32;
  const int s = r.reverse ? r.n_slices - 1 - item / r.n_tiles : item / r.n_tiles;
  __threadfence();
  rp16_sync<BAR, NT>();
  if (t == 0) {
    const int prev = atomicAdd(r.tiles_done + s, 1);
    *s_last = prev == r.n_tiles - 1 ? 1 : 0;
    if (*s_last) __threadfence();
  }
  rp16_sync<BAR, NT>();
  if (*s_last) {
    const long long n = (long long)r.oh * r.ow;
    if (t == 0) {
      double cnt = 0.0, mean = 0.0, m2 = 0.0;
      for (int i = 0; i < r.n_tiles; ++i) {
        const float* q = r.partials + ((long long)s * r.n_tiles + i) * 3;
        const double nb = rp16_ld_cg(q), mb = rp16_ld_cg(q + 1), sb = rp16_ld_cg(q + 2);
        if (nb > 0.0) {
          const double d = mb - mean, tot = cnt + nb;
          mean += d * nb / tot;
          m2 += sb + d * d * cnt * nb / tot;
          cnt = tot;
        }
      }
      s_stat[0] = (float)mean;
      s_stat[1] = (float)sqrt(m2 / (double)(n - 1));
      if (r.mean_std) { r.mean_std[2 * s] = s_stat[0]; r.mean_std[2 * s + 1] = s_stat[1]; }
    }
    rp16_sync<BAR, NT>();
    if (r.normalize) {
      const float fmean = s_stat[0], den = s_stat[1] + r.eps;
      float* y = r.out + (long long)s * n;
      if ((n & 3) == 0 && (((unsigned long long)y) & 15) == 0) {
        float4* y4 = reinterpret_cast<float4*>(y);
        for (long long i = t; i < n / 4; i += NT) {
          float4 v = rp16_ld_cg4(y4 + i);
          v.x = (v.x - fmean) / den; v.y = (v.y - fmean) / den; v.z = (v.z - fmean) / den; v.w = (v.w - fmean) / den;
          y4[i] = v;
        }
      } else {
        for (long long i = t; i < n; i += NT) y[i] = (rp16_ld_cg(y + i) - fmean) / den;
      }
    }
  }
  rp16_sync<BAR, NT>();
}

template <int P, int Q, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) rowpass16_kernel(RowPass16Params p) {
  MRIACL_DYN_SMEM(cf, smem);
  __shared__ float red[NW];
  __shared__ float s_stat[2];
  __shared__ int ready, s_last;
  Rp16Smem<P, Q> S(smem, p);
  rp16_load_tables<NW * 32>(p, S.sptw, S.sch, S.tbuf, threadIdx.x);
  __syncthreads();
  const int n_items = p.n_slices * p.n_tiles;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    rowpass16_item<P, Q, NW>(p, smem, item, threadIdx.x, red, &ready);
    if (p.done && !ready) return;
    if (p.tiles_done) rowpass16_finish_slice<NW, 0>(p, item, threadIdx.x, s_stat, &s_last);
  }
}

}  // namespace mriacl
