// colpass640.cuh -- fused column pass of the k-space -> image stage for H = 640.
//
// One work item = (frame, group of <= G sampled phase-encode columns; G = 4 in the product kernel, see the end of
// this file for why).  The CTA gathers the
// group's columns from all 640 readout rows of complex64 k-space (the only HBM read of the
// path; every 32-byte sector of a row that holds a sampled column is touched exactly once)
// with 8-byte cp.async straight into shared memory, one item ahead of the arithmetic (double
// buffered), multiplies by the mask value, and runs the 640-point inverse DFT of each column in shared
// memory as three in-place decimation-in-frequency passes 8 x 8 x 10:
//
//   n = 80 n1 + 10 n2 + n3   ->   m = m1 + 8 m2 + 64 m3
//   pass 1: radix-8 over n1, twiddle w640^{(10 n2 + n3) m1}
//   pass 2: radix-8 over n2, twiddle w80^{n3 m2}
//   pass 3: radix-10 over n3, results leave in natural order m
//
// The input ifftshift is the load index map (physical row h -> logical n = (h - 320) mod 640),
// the output fftshift + centre crop + optional flipud is the store index map, so neither shift
// moves data.  Only the out_h kept rows are written, to the intermediate T[frame][j][row]
// (row-contiguous with a pitch that is a multiple of 32, so the row pass can bulk-copy 32-row
// blocks with 16-byte cp.async).
//
// Shared-memory layout per column: slot(m1, m2, n3) = 90 m1 + 10 m2 + n3 (each block of 80 is
// padded to 90) and column pitch 722 complex.  With it pass 1 (lanes over 10 n2 + n3), pass 2
// (lanes over 10 m1 + n3) and pass 3's 128-bit loads (lanes over m1, then m2) run at the ideal
// wavefront count (ncu source view, profiles/r02_colpass_smem_wavefronts.md: LDS.64 / STS.64 / LDS.128 excess 0).
// The gather itself is different: an LDGSTS writes shared memory once per 128-byte global line it
// touched, and at 4x undersampling a line holds only four wanted elements, so the 8-byte copies
// cost 3.4x their ideal wavefronts (reported by ncu as "bank conflicts": 19 % of the kernel's
// shared-memory wavefronts).  That is a property of the access pattern, not of the layout.
#pragma once
#include "butterflies.cuh"

namespace mriacl {

constexpr int CP_N = 640;        // transform length
constexpr int CP_T = 160;        // threads per CTA: 2 column subsets x 80 butterfly positions
constexpr int CP_G = 8;          // columns per work item
constexpr int CP_BLK = 90;       // padded size of one n1/m1 block of 80
constexpr int CP_PITCH = 722;    // complex elements between columns in shared memory (== 2 mod 16)
constexpr int CP_BUF = CP_G * CP_PITCH;          // complex elements of one item buffer
constexpr int CP_SMEM_BYTES_DB = 2 * CP_BUF * 8; // persistent, double buffered
constexpr int CP_SMEM_BYTES_SB = CP_BUF * 8;     // one item per CTA, single buffer

struct ColPassParams {
  const cf* ksp;                 // k-space, element (b,a,c,h,w) at b*sb + a*sa + (c*H + h)*W + w
  long long sb, sa;
  int A, C, W;
  const int* act_w;              // [n_act] physical (unpadded) column of active column j
  const float* act_m;            // [n_act] mask value of active column j
  int unit_mask;                 // 1: every mask value is exactly 1 (skip the multiply)
  int n_act;
  int n_groups;                  // ceil(n_act / 8)
  const cf* tw;                  // w640^k = exp(+2 pi i k / 640)
  cf* T;                         // [n_frames][n_act][ohp]
  int oh, ohp, row0, flip;       // ohp = row pitch of T (multiple of 32)
  int frame0;                    // first global frame of this launch (f = (b*A + a)*C + c)
  int n_frames;
  int* done;                     // optional [n_slices]: += 1 per finished item of the slice (row pass waits on it)
  int persist;                   // single-buffer variant: 1 = grid-stride over items, 0 = one item per CTA
  int debug_skip;                // profiling only: 1 = no gather, 2 = no arithmetic, 4 = no stores (results are garbage)
  int l2_hints;                  // bit 0: k-space loads evict_first, bit 1: T stores evict_last
  // co-resident schedule with a ring of T slots: slice s lives in slot s % ring, and its columns may only be written
  // once every row tile of slice s - ring has been consumed (rows_done[s - ring] == rows_target); ring = 0: no ring
  int ring;
  const int* rows_done;
  int rows_target;
};

// 8-byte asynchronous global -> shared copy (LDGSTS): the gather needs one complex64 out of
// every 32-byte sector, so 16-byte copies cannot be used.
__device__ __forceinline__ void cp_async8(cf* smem_dst, const cf* gsrc) {
#if defined(MRIACL_EMU)
  *smem_dst = *gsrc;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8_hint(cf* smem_dst, const cf* gsrc, unsigned long long pol) {
#if defined(MRIACL_EMU)
  *smem_dst = *gsrc;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit_group() {
#if !defined(MRIACL_EMU)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N_PENDING> __device__ __forceinline__ void cp_async_wait_group() {
#if !defined(MRIACL_EMU)
  asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory");
#endif
}

#ifdef MRIACL_EXPERIMENTAL   // the compute warps issue their own gather: superseded by the warp-specialised kernel below
// DB = true: persistent grid, next item's gather in flight during the arithmetic (2 CTAs/SM);
// DB = false: one item per CTA, single buffer (4 CTAs/SM; leaves room for a row-pass CTA on the SM).
template <bool DB>
__global__ void __launch_bounds__(CP_T, DB ? 2 : 4) colpass640_kernel(ColPassParams p) {
  MRIACL_DYN_SMEM(cf, sm);
  const int tid = threadIdx.x;
  const int sub = tid / 80, pos = tid - sub * 80;

  // twiddles depend only on the butterfly position: keep them in registers across items
  cf tw1[8], tw2[8];
  const int base2 = (pos / 10) * CP_BLK + (pos % 10);     // pass 2: (m1, n3) = (pos/10, pos%10)
  {
    const int n3 = pos % 10;
#pragma unroll
    for (int m = 1; m < 8; ++m) {
      tw1[m] = p.tw[(pos * m) % CP_N];
      tw2[m] = p.tw[(8 * n3 * m) % CP_N];
    }
  }
  // pass 3: thread r < 64 owns (m1, m2) = (r % 8, r / 8); its ten outputs m = r + 64 m3 land on
  // fixed rows of T (fftshift + crop + flip), computed once
  const int sub3 = tid / 64, r3 = tid - sub3 * 64;
  const int base3 = (r3 % 8) * CP_BLK + (r3 / 8) * 10;
  int rr3[10];
#pragma unroll
  for (int m3 = 0; m3 < 10; ++m3) {
    const int rfull = phys_of_logical(r3 + 64 * m3, CP_N);
    const int rr = (p.flip ? CP_N - 1 - rfull : rfull) - p.row0;
    rr3[m3] = (rr >= 0 && rr < p.oh) ? rr : -1;
  }

  // gather: thread (k, hs) copies rows h = 80 blk + 20 q + hs of column k; the logical index is
  // (h + 320) mod 640 = 80 ((blk + 4) & 7) + 20 q + hs, i.e. slot 90 ((blk + 4) & 7) + 20 q + hs
  const int k_ld = tid & 7, hs_ld = tid >> 3;
  const int n_items = p.n_frames * p.n_groups;
  const long long row_step = 20LL * p.W;

  auto issue_gather = [&](int item, int buf) {
    const int fl = item / p.n_groups, g = item - fl * p.n_groups;
    const int f = p.frame0 + fl;
    const int b = f / (p.A * p.C), a = (f / p.C) % p.A, c = f % p.C;
    const int j0 = g * CP_G;
    if (j0 + k_ld < p.n_act) {
      const cf* src = p.ksp + b * p.sb + a * p.sa + ((long long)c * CP_N + hs_ld) * p.W + p.act_w[j0 + k_ld];
      cf* dst = sm + buf * CP_BUF + k_ld * CP_PITCH + hs_ld;
#pragma unroll
      for (int blk = 0; blk < 8; ++blk) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          cp_async8(dst + CP_BLK * ((blk + 4) & 7) + 20 * q, src);
          src += row_step;
        }
      }
    }
    cp_async_commit_group();
  };

  int buf = 0;
  if (blockIdx.x < n_items) issue_gather(blockIdx.x, 0);

  for (int item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= (DB ? 1 : 0)) {
    const int fl = item / p.n_groups, g = item - fl * p.n_groups;
    const int j0 = g * CP_G;
    const int ncols = min(CP_G, p.n_act - j0);
    cf* cur = sm + buf * CP_BUF;

    // prefetch the next item into the other buffer (its last readers finished before the barrier
    // that ended the previous iteration), then wait for this item's gather
    const int next = item + gridDim.x;
    if (DB && next < n_items) { issue_gather(next, buf ^ 1); cp_async_wait_group<1>(); } else cp_async_wait_group<0>();
    __syncthreads();

    // ---- pass 1: radix-8 over n1 (stride 90), mask multiply, twiddle w640^{pos * m1} -------
    for (int k = sub; k < ncols; k += 2) {
      cf* col = cur + k * CP_PITCH + pos;
      cf v[8];
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) v[n1] = col[n1 * CP_BLK];
      if (!p.unit_mask) {
        const float mv = p.act_m[j0 + k];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) v[n1] = cscale(v[n1], mv);
      }
      radix8<true>(v);
      col[0] = v[0];
#pragma unroll
      for (int m1 = 1; m1 < 8; ++m1) col[m1 * CP_BLK] = cmul(v[m1], tw1[m1]);
    }
    __syncthreads();

    // ---- pass 2: radix-8 over n2 (stride 10), twiddle w80^{n3 * m2} -----------------------
    for (int k = sub; k < ncols; k += 2) {
      cf* col = cur + k * CP_PITCH + base2;
      cf v[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) v[n2] = col[n2 * 10];
      radix8<true>(v);
      col[0] = v[0];
#pragma unroll
      for (int m2 = 1; m2 < 8; ++m2) col[m2 * 10] = cmul(v[m2], tw2[m2]);
    }
    __syncthreads();

    // ---- pass 3: radix-10 over n3 (contiguous), crop/shift/flip on the way out ------------
    if (tid < 128) {
      for (int k = sub3; k < ncols; k += 2) {
        const float4* col4 = reinterpret_cast<const float4*>(cur + k * CP_PITCH + base3);
        cf v[10];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float4 t = col4[q];
          v[2 * q] = cf_make(t.x, t.y);
          v[2 * q + 1] = cf_make(t.z, t.w);
        }
        radix10<true>(v);
        cf* dst = p.T + ((long long)fl * p.n_act + j0 + k) * p.ohp;
#pragma unroll
        for (int m3 = 0; m3 < 10; ++m3)
          if (rr3[m3] >= 0) dst[rr3[m3]] = v[m3];
      }
    }
    if (p.done) {        // publish: T rows of this item are visible device-wide before the count moves
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicAdd(p.done + fl / (p.A * p.C), 1);
    } else {
      __syncthreads();   // this buffer is free for the gather issued in the next iteration
    }
    if (!DB) {           // single buffer: the next gather can only start now (other CTAs of the SM cover the latency)
      if (!p.persist) break;
      if (next < n_items) issue_gather(next, 0);
    }
  }
}


#endif  // MRIACL_EXPERIMENTAL

// ---------------------------------------------------------------------------------------------
// Warp-specialised column pass: 5 compute warps + 1 producer warp per CTA, two item buffers.
// An LDGSTS whose queue is full stalls the issuing warp, so when the compute warps issue their own
// gather the load phase cannot overlap their arithmetic.  Here the producer warp does nothing but
// gather (it may sit in the memory-queue stall for as long as it likes) and hands buffers to the compute
// warps through named barriers FULL[b] / EMPTY[b]; the compute warps synchronise among themselves on a
// 160-thread barrier, so the gather of item i+1 streams in during passes 1-3 of item i.
// ---------------------------------------------------------------------------------------------
constexpr int CP_WS_T = CP_T + 32;
constexpr int CP_BAR_FULL = 1, CP_BAR_EMPTY = 3, CP_BAR_COMPUTE = 5;   // FULL: 1,2  EMPTY: 3,4

// Runs the items first, first + stride, ... (count of them) through the producer / consumer pipeline.
// Called by CP_WS_T consecutive threads (tid 0 .. CP_WS_T-1) with two item buffers at `sm`.  uses[b] counts
// how often buffer b has been filled since full_init (it selects the mbarrier phase), so the function can be
// called again and again by a persistent CTA.
// G = columns per item (8: the stand-alone kernel; 4: half-size items, so that two column teams fit beside a row team on one
// SM -- p.n_groups must then count groups of 4); BARB = first named barrier of this team minus one (a team uses five).
template <int G = CP_G, int BARB = 0>
__device__ __forceinline__ void colpass_ws_run(const ColPassParams& p, cf* sm, FullBarrier* full_bar, int tid,
                                               int first, int stride, int count, int* uses) {
  static_assert(G == 8 || G == 4 || G == 2, "items of 8, 4 or 2 columns");
  constexpr int BUF = G * CP_PITCH;                      // complex elements of one item buffer
  constexpr int BAR_FULL = BARB + CP_BAR_FULL, BAR_EMPTY = BARB + CP_BAR_EMPTY, BAR_COMPUTE = BARB + CP_BAR_COMPUTE;
  constexpr int RPI = 32 / G;                            // rows per gather instruction
  constexpr int NC = G / 2;                              // columns a transform thread owns on a full item
  if (tid >= CP_T) {
    // ------------------------------ producer warp ------------------------------
    const int lane = tid - CP_T;
    const int k_ld = lane % G, hs = lane / G;           // G columns x (32 / G) rows per instruction
    const long long row_step = (long long)RPI * p.W;
    const unsigned long long pol = l2_policy_evict_first();
    for (int k = 0; k < count; ++k) {
      const int item = first + k * stride;
      const int buf = k & 1;
      if (k >= 2) named_bar_sync(BAR_EMPTY + buf, CP_WS_T);   // buffer released by the compute warps
      const int fl = item / p.n_groups, g = item - fl * p.n_groups;
      const int f = p.frame0 + fl;
      const int b = f / (p.A * p.C), a = (f / p.C) % p.A, c = f % p.C;
      const int j0 = g * G;
      if (j0 + k_ld < p.n_act && !MRIACL_DBG_SKIP(p, 1)) {
        const cf* src = p.ksp + b * p.sb + a * p.sa + ((long long)c * CP_N + hs) * p.W + p.act_w[j0 + k_ld];
        cf* dst = sm + buf * BUF + k_ld * CP_PITCH + hs;
        if (p.l2_hints & 1) {
#pragma unroll 1
          for (int blk = 0; blk < 8; ++blk) {
            cf* d = dst + CP_BLK * ((blk + 4) & 7);
#pragma unroll
            for (int q = 0; q < 80 / RPI; ++q) {
              cp_async8_hint(d + RPI * q, src, pol);
              src += row_step;
            }
          }
        } else {
#pragma unroll 1
          for (int blk = 0; blk < 8; ++blk) {
            cf* d = dst + CP_BLK * ((blk + 4) & 7);
#pragma unroll
            for (int q = 0; q < 80 / RPI; ++q) {
              cp_async8(d + RPI * q, src);
              src += row_step;
            }
          }
        }
      }
      // fires when this lane's copies have landed; the warp moves straight on to the next item's gather
      full_signal_async(&full_bar[buf], BAR_FULL + buf, CP_WS_T);
    }
    cp_async_wait_group<0>();
    return;
  }

  // ------------------------------ compute warps ------------------------------
  const int sub = tid / 80, pos = tid - sub * 80;
  cf tw1[8], tw2[8];
  const int base2 = (pos / 10) * CP_BLK + (pos % 10);
  {
    const int n3 = pos % 10;
#pragma unroll
    for (int m = 1; m < 8; ++m) {
      tw1[m] = p.tw[(pos * m) % CP_N];
      tw2[m] = p.tw[(8 * n3 * m) % CP_N];
    }
  }
  const int sub3 = tid / 64, r3 = tid - sub3 * 64;
  const int base3 = (r3 % 8) * CP_BLK + (r3 / 8) * 10;
  int rr3[10];
#pragma unroll
  for (int m3 = 0; m3 < 10; ++m3) {
    const int rfull = phys_of_logical(r3 + 64 * m3, CP_N);
    const int rr = (p.flip ? CP_N - 1 - rfull : rfull) - p.row0;
    rr3[m3] = (rr >= 0 && rr < p.oh) ? rr : -1;
  }

  const unsigned long long pol_t = l2_policy_evict_last();
  for (int k = 0; k < count; ++k) {
    const int item = first + k * stride;
    const int fl = item / p.n_groups, g = item - fl * p.n_groups;
    const int j0 = g * G;
    const int ncols = min(G, p.n_act - j0);
    const int buf = k & 1;
    cf* cur = sm + buf * BUF;
    full_wait(&full_bar[buf], (uses[buf] + (k >> 1)) & 1, BAR_FULL + buf, CP_WS_T);

    const int ncols_c = MRIACL_DBG_SKIP(p, 2) ? 0 : ncols;
    // Full item: each thread owns columns sub, sub+2, ... (G / 2 of them) and keeps all their butterflies of a pass in
    // flight (instruction-level parallelism between the team barriers); ragged items take the loop below.
    if (ncols_c == G && p.unit_mask) {
      cf* col0 = cur + sub * CP_PITCH;
      {
        cf v[NC][8];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) v[c][n1] = col0[2 * c * CP_PITCH + pos + n1 * CP_BLK];
#pragma unroll
        for (int c = 0; c < NC; ++c) radix8<true>(v[c]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          cf* col = col0 + 2 * c * CP_PITCH + pos;
          col[0] = v[c][0];
#pragma unroll
          for (int m1 = 1; m1 < 8; ++m1) col[m1 * CP_BLK] = cmul(v[c][m1], tw1[m1]);
        }
      }
      named_bar_sync(BAR_COMPUTE, CP_T);
      {
        cf v[NC][8];
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
          for (int n2 = 0; n2 < 8; ++n2) v[c][n2] = col0[2 * c * CP_PITCH + base2 + n2 * 10];
#pragma unroll
        for (int c = 0; c < NC; ++c) radix8<true>(v[c]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          cf* col = col0 + 2 * c * CP_PITCH + base2;
          col[0] = v[c][0];
#pragma unroll
          for (int m2 = 1; m2 < 8; ++m2) col[m2 * 10] = cmul(v[c][m2], tw2[m2]);
        }
      }
      named_bar_sync(BAR_COMPUTE, CP_T);
    } else {
    // ---- pass 1: radix-8 over n1 (stride 90), mask multiply, twiddle w640^{pos * m1}; two columns in flight ----
    for (int kc = sub; kc < ncols_c; kc += 4) {
      const bool two = kc + 2 < ncols;
      cf* colA = cur + kc * CP_PITCH + pos;
      cf* colB = colA + (two ? 2 * CP_PITCH : 0);
      cf va[8], vb[8];
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) { va[n1] = colA[n1 * CP_BLK]; vb[n1] = colB[n1 * CP_BLK]; }
      if (!p.unit_mask) {
        const float ma = p.act_m[j0 + kc], mb = p.act_m[j0 + (two ? kc + 2 : kc)];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) { va[n1] = cscale(va[n1], ma); vb[n1] = cscale(vb[n1], mb); }
      }
      radix8<true>(va);
      radix8<true>(vb);
      colA[0] = va[0];
#pragma unroll
      for (int m1 = 1; m1 < 8; ++m1) colA[m1 * CP_BLK] = cmul(va[m1], tw1[m1]);
      if (two) {
        colB[0] = vb[0];
#pragma unroll
        for (int m1 = 1; m1 < 8; ++m1) colB[m1 * CP_BLK] = cmul(vb[m1], tw1[m1]);
      }
    }
    named_bar_sync(BAR_COMPUTE, CP_T);

    // ---- pass 2: radix-8 over n2 (stride 10), twiddle w80^{n3 * m2} ----
    for (int kc = sub; kc < ncols_c; kc += 4) {
      const bool two = kc + 2 < ncols;
      cf* colA = cur + kc * CP_PITCH + base2;
      cf* colB = colA + (two ? 2 * CP_PITCH : 0);
      cf va[8], vb[8];
#pragma unroll
      for (int n2 = 0; n2 < 8; ++n2) { va[n2] = colA[n2 * 10]; vb[n2] = colB[n2 * 10]; }
      radix8<true>(va);
      radix8<true>(vb);
      colA[0] = va[0];
#pragma unroll
      for (int m2 = 1; m2 < 8; ++m2) colA[m2 * 10] = cmul(va[m2], tw2[m2]);
      if (two) {
        colB[0] = vb[0];
#pragma unroll
        for (int m2 = 1; m2 < 8; ++m2) colB[m2 * 10] = cmul(vb[m2], tw2[m2]);
      }
    }
    named_bar_sync(BAR_COMPUTE, CP_T);

    }
    // ---- pass 3: radix-10 over n3 (contiguous), crop/shift/flip on the way out ----
    int t_frame = fl;
    if (p.ring) {        // ring of T slots: wait until the slot's previous slice has been consumed by the row teams
      const int fpf = p.A * p.C, sl = fl / fpf;
      t_frame = (sl % p.ring) * fpf + (fl - sl * fpf);
      if (sl >= p.ring) {
        if (tid == 0) wait_count_ge(p.rows_done + (sl - p.ring), p.rows_target);
        named_bar_sync(BAR_COMPUTE, CP_T);
      }
    }
    if (tid < 128) {
      for (int kc = sub3; kc < ncols_c; kc += 2) {
        const float4* col4 = reinterpret_cast<const float4*>(cur + kc * CP_PITCH + base3);
        cf v[10];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float4 t = col4[q];
          v[2 * q] = cf_make(t.x, t.y);
          v[2 * q + 1] = cf_make(t.z, t.w);
        }
        radix10<true>(v);
        cf* dst = p.T + ((long long)t_frame * p.n_act + j0 + kc) * p.ohp;
        if (p.l2_hints & 2) {
#pragma unroll
          for (int m3 = 0; m3 < 10; ++m3)
            if (rr3[m3] >= 0) st_global_hint(dst + rr3[m3], v[m3], pol_t);
        } else if (!MRIACL_DBG_SKIP(p, 4)) {
#pragma unroll
          for (int m3 = 0; m3 < 10; ++m3)
            if (rr3[m3] >= 0) dst[rr3[m3]] = v[m3];
        }
      }
    }
    if (p.done) {
      __threadfence();
      named_bar_sync(BAR_COMPUTE, CP_T);
      if (tid == 0) atomicAdd(p.done + fl / (p.A * p.C), 1);
    }
    if (k + 2 < count) named_bar_arrive(BAR_EMPTY + buf, CP_WS_T);   // producer may refill
  }
}

// The product kernel works on FOUR-column items (p.n_groups counts groups of CP_GW), two persistent CTAs per SM.
// The 8-byte LDGSTS gather stages its in-flight lines in L1, and L1 is whatever the CTAs' shared memory leaves of the
// SM's 256 KB: with eight-column items (2 x 92 KB -> the 196 KB configuration, 60 KB of L1) the column pass takes
// 0.381 ms per 64 slices, with four-column items (2 x 46 KB -> 100 KB configuration, 156 KB of L1) 0.315 ms; forcing
// the carveout shows the dependence directly: 28 KB of L1 0.65 ms, 92 KB 0.40 ms, 124 KB 0.33 ms, 156 KB 0.315 ms
// (profiles/r02t_*).  Two-column items (0.33 ms) and three or four CTAs per SM (87 / 80 registers: 0.36 / 0.42 ms) lose
// more in the transform than the larger L1 gives.
constexpr int CP_GW = 4;
constexpr int CP_SMEM_BYTES_WS = 2 * CP_GW * CP_PITCH * 8;

__global__ void __launch_bounds__(CP_WS_T, 2) colpass640_ws_kernel(ColPassParams p) {
  MRIACL_DYN_SMEM(cf, sm);
  __shared__ FullBarrier full_bar[2];
  const int n_items = p.n_frames * p.n_groups;
  if (threadIdx.x == 0) { full_init(&full_bar[0], 32); full_init(&full_bar[1], 32); }
  __syncthreads();
  int uses[2] = {0, 0};
  const int first = blockIdx.x;
  if (first < n_items)
    colpass_ws_run<CP_GW, 0>(p, sm, full_bar, threadIdx.x, first, gridDim.x, (n_items - first + gridDim.x - 1) / gridDim.x, uses);
}

#ifdef MRIACL_EXPERIMENTAL
// stand-alone column pass on G-column items with MINB CTAs per SM (A/B of item size against L1 capacity: the smaller the
// CTAs' shared memory, the larger the L1 left for the gather's in-flight lines)
template <int G, int MINB>
__global__ void __launch_bounds__(CP_WS_T, MINB) colpass640_ws_g_kernel(ColPassParams p) {
  MRIACL_DYN_SMEM(cf, sm);
  __shared__ FullBarrier full_bar[2];
  const int n_items = p.n_frames * p.n_groups;
  if (threadIdx.x == 0) { full_init(&full_bar[0], 32); full_init(&full_bar[1], 32); }
  __syncthreads();
  int uses[2] = {0, 0};
  const int first = blockIdx.x;
  if (first < n_items)
    colpass_ws_run<G, 0>(p, sm, full_bar, threadIdx.x, first, gridDim.x, (n_items - first + gridDim.x - 1) / gridDim.x, uses);
}
#endif

}  // namespace mriacl
