// colpass640_tma.cuh -- column pass of the H = 640 plans with a TMA gather (no L1, no LSU, no issue slots).
//
// Why: the 8-byte LDGSTS.ca gather of colpass640.cuh stages every in-flight line in L1, and L1 is whatever the resident
// CTAs' shared memory leaves of the SM's 256 KB -- with a large CTA (a row team beside the column team) 28 KB of L1 are
// left and the gather halves.  At 4x undersampling every 32-byte sector of k-space holds a sampled
// column anyway, so nothing is saved by gathering single elements: this variant lets the TMA unit stream whole BANDS of
// 8 raw columns x 640 rows (64-byte rows, 40 KB, five 128-row boxes of cp.async.bulk.tensor) into a ring of raw slots.
// Measured (tools/microbench/gather_modes.cu, one thread issuing per SM): 6.5-6.9 TB/s of k-space with a 200 KB CTA,
// against 3.0-3.4 TB/s for the 8-byte cp.async.ca gather in the same CTA.
//
// One item = (frame, band with at least one sampled column).  A transform team (160 threads) takes the item's sampled
// columns two at a time: pass 1 (radix-8 over n1) reads them out of the raw slot (64-byte-swizzled rows: eight
// consecutive rows of one column sit in eight different 16-byte bank groups) and writes the compact transform layout
// of colpass640.cuh into a small work buffer; passes 2 and 3 are the ones of colpass640.cuh on that work buffer.  The
// raw slot goes back to the TMA thread as soon as the item's last pass 1 has read it, so two slots per CTA keep about
// one and a half bands in flight.
#pragma once
#include "colpass640.cuh"

#ifndef MRIACL_EMU
#include <cuda.h>

namespace mriacl {

constexpr int CT_BW = 8;                               // raw columns per band
constexpr int CT_G = 2;                                // sampled columns per transform round
constexpr int CT_BOX_ROWS = 128;                       // rows per TMA box (640 = 5 boxes)
constexpr int CT_SLOT_BYTES = CT_BW * 8 * CP_N;        // 40 960
constexpr int CT_WORK_CF = CT_G * CP_PITCH;            // complex elements of one work buffer
constexpr int CT_TEAM_SMEM = 2 * CT_WORK_CF * 8;       // two alternating work buffers per team: 23 104 B
constexpr int CT_MAX_SLOTS = 4;

struct ColTmaParams {
  ColPassParams cp;          // ksp / sb / sa are unused (the tensor map carries them); n_groups = items per frame
  const int* item_band;      // [items_per_frame]     band index (raw columns 8 band .. 8 band + 7)
  const int* item_j0;        // [items_per_frame + 1] first active column of the item
  int n_slots;               // raw slots in the ring (2 .. CT_MAX_SLOTS)
  int work_bufs;             // work buffers per team: 2 (alternating, no extra barrier) or 1 (one more team barrier per round)
};

// (+ coltma_table_bytes for the plan tables behind the work buffers)
inline int coltma_smem_bytes(int n_slots, int n_teams, int work_bufs = 2) {
  return 1024 + n_slots * CT_SLOT_BYTES + n_teams * work_bufs * CT_WORK_CF * 8;
}

__device__ __forceinline__ unsigned ct_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_mbar_init(unsigned long long* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ct_s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void ct_mbar_arrive(unsigned long long* b) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.shared::cta.b64 t, [%0];\n\t}" ::"r"(ct_s32(b)) : "memory");
}
__device__ __forceinline__ void ct_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("{\n\t.reg .b64 t;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n\t}" ::"r"(ct_s32(b)), "r"(bytes) : "memory");
}
// bounded (about a second): a protocol bug must trap, not hang the device
__device__ __forceinline__ void ct_mbar_wait(unsigned long long* b, int parity) {
  const unsigned a = ct_s32(b);
  for (int spin = 0; spin < (1 << 26); ++spin) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void ct_tma_load_4d(void* dst, const CUtensorMap* map, int x, int y, int z, int w, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(ct_s32(dst)), "l"(map), "r"(ct_s32(bar)), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// (frame, item-in-frame) of a CTA item index, advanced by a fixed step without divisions
struct ColTmaIter {
  int fl, i, dq, dr, ipf;
  __device__ __forceinline__ void init(int item, int step, int ipf_) {
    ipf = ipf_; fl = item / ipf; i = item - fl * ipf; dq = step / ipf; dr = step - dq * ipf;
  }
  __device__ __forceinline__ void next() { fl += dq; i += dr; if (i >= ipf) { i -= ipf; ++fl; } }
};

// plan tables in shared memory: band of item i, first active column of item i (one entry more), raw column of active j
struct ColTmaTables { const int* band; const int* j0; const int* w; };
__host__ __device__ inline int coltma_table_bytes(int n_items_per_frame, int n_act) { return ((2 * n_items_per_frame + 1 + n_act) * 4 + 15) & ~15; }
__device__ __forceinline__ ColTmaTables coltma_load_tables(const ColTmaParams& p, int* dst, int tid, int nt) {
  const int ipf = p.cp.n_groups;
  for (int i = tid; i < ipf; i += nt) dst[i] = p.item_band[i];
  for (int i = tid; i <= ipf; i += nt) dst[ipf + i] = p.item_j0[i];
  for (int i = tid; i < p.cp.n_act; i += nt) dst[2 * ipf + 1 + i] = p.cp.act_w[i];
  return ColTmaTables{dst, dst + ipf, dst + 2 * ipf + 1};
}

// issue the five boxes of (local frame fl, item i) into `dst` (one thread)
__device__ __forceinline__ void coltma_issue(const CUtensorMap* map, const ColPassParams& cp, int band, int fl,
                                             unsigned char* dst, unsigned long long* bar) {
  const int f = cp.frame0 + fl, fpf = cp.A * cp.C;
  const int b = f / fpf, rem = f - b * fpf, a = rem / cp.C, c = rem - a * cp.C;
  ct_mbar_expect_tx(bar, CT_SLOT_BYTES);
#pragma unroll
  for (int q = 0; q < CP_N / CT_BOX_ROWS; ++q)
    ct_tma_load_4d(dst + q * (CT_BOX_ROWS * CT_BW * 8), map, band * (2 * CT_BW), c * CP_N + q * CT_BOX_ROWS, a, b, bar);
}

// One transform team (CP_T threads, tid 0 .. CP_T-1): items k = team, team + n_teams, ... of the CTA's sequence
// first, first + stride, ...  The team owns the raw slots team, team + n_teams, ... (< n_slots) and feeds them itself:
// its thread 0 issues the TMA of a later item into a slot right after the team barrier that ends the slot's last pass 1
// (a separate TMA thread serving every slot of the CTA in turn was the bottleneck of the first version: one thread's
// index arithmetic, at the issue rate a warp gets on a busy SM, capped the CTA at one item per ~1.4 us).
// BAR = the team's named barrier (a run-time value: every team of a CTA runs the same ~30 KB copy of this code).
__device__ __forceinline__ void coltma_team(const int BAR, const CUtensorMap* map, const ColTmaParams& p, const ColTmaTables& tb,
                                         unsigned char* raw, cf* work, unsigned long long* full, int tid, int team,
                                         int n_teams, int first, int stride, int count) {
  const ColPassParams& cp = p.cp;
  const int ns = p.n_slots;
  const int owned = (ns - team + n_teams - 1) / n_teams;           // >= 1: the launch code keeps n_slots >= n_teams
  const int my_count = count > team ? (count - team + n_teams - 1) / n_teams : 0;
  const int sub = tid / 80, pos = tid - sub * 80;
  cf tw1[8], tw2[8];
  const int base2 = (pos / 10) * CP_BLK + (pos % 10);
  {
    const int n3 = pos % 10;
#pragma unroll
    for (int m = 1; m < 8; ++m) {
      tw1[m] = cp.tw[(pos * m) % CP_N];
      tw2[m] = cp.tw[(8 * n3 * m) % CP_N];
    }
  }
  const int sub3 = tid / 64, r3 = tid - sub3 * 64;
  const int base3 = (r3 % 8) * CP_BLK + (r3 / 8) * 10;
  int rr3[10];
#pragma unroll
  for (int m3 = 0; m3 < 10; ++m3) {
    const int rfull = phys_of_logical(r3 + 64 * m3, CP_N);
    const int rr = (cp.flip ? CP_N - 1 - rfull : rfull) - cp.row0;
    rr3[m3] = (rr >= 0 && rr < cp.oh) ? rr : -1;
  }
  // pass-1 source of logical n = 80 n1 + pos: physical row 80 ((n1 + 4) & 7) + pos of the raw slot; 64-byte rows with
  // the TMA 64B swizzle (16-byte chunk index ^= bits 7-8 of the byte offset; 80 rows = 5120 B leave those bits alone)
  const unsigned row_off = (unsigned)pos * (CT_BW * 8);
  const unsigned swz = ((row_off >> 7) & 3u) << 4;
  const int fpf = cp.A * cp.C;

  ColTmaIter it, iss;                     // item being transformed / next item to fetch
  it.init(first + team * stride, n_teams * stride, cp.n_groups);
  iss = it;
  int n_issued = 0;
  if (tid == 0) {
    for (; n_issued < owned && n_issued < my_count; ++n_issued, iss.next()) {
      const int slot = team + n_issued * n_teams;
      coltma_issue(map, cp, tb.band[iss.i], iss.fl, raw + (size_t)slot * CT_SLOT_BYTES, &full[slot]);
    }
  }

  int wsel = 0, oslot = 0, ophase = 0;    // owned-slot cursor and its mbarrier phase
  int pend_slice = -1, pend_count = 0;    // finished items of slice pend_slice whose counter has not moved yet
  bool started = false;
  for (int j = 0; j < my_count; ++j, it.next()) {
    const int fl = it.fl, i = it.i;
    const int band = tb.band[i], j0 = tb.j0[i], ncols = tb.j0[i + 1] - j0;
    const int slot = team + oslot * n_teams;
    ct_mbar_wait(&full[slot], ophase);
    const unsigned char* rs = raw + (size_t)slot * CT_SLOT_BYTES;
    const int sl = fl / fpf;
    int t_frame = fl;
    if (cp.ring) {           // ring of T slots (co-resident schedules): the slot's previous slice must have been consumed
      t_frame = (sl % cp.ring) * fpf + (fl - sl * fpf);
      if (sl >= cp.ring && tid == 0) wait_count_ge(cp.rows_done + (sl - cp.ring), cp.rows_target);
      // (the team barrier after pass 1 orders the other threads' T stores behind this wait)
    }
    const int n_rounds = (ncols + CT_G - 1) / CT_G;
    for (int sg = 0; sg < n_rounds; ++sg) {
      cf* wk = work + wsel * CT_WORK_CF;
      if (p.work_bufs == 2) wsel ^= 1;
      else if (started) named_bar_sync(BAR, CP_T);      // single work buffer: the previous round's pass 3 has read it
      started = true;
      const int kc = CT_G * sg + sub;
      // ---- pass 1: radix-8 over n1 out of the raw slot, mask multiply, twiddle w640^{pos * m1} ----
      if (kc < ncols) {
        const unsigned off = (row_off + (unsigned)(tb.w[j0 + kc] - band * CT_BW) * 8u) ^ swz;
        cf v[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1)
          v[n1] = *reinterpret_cast<const cf*>(rs + off + (unsigned)(((n1 + 4) & 7) * 80 * CT_BW * 8));
        if (!cp.unit_mask) {
          const float mv = cp.act_m[j0 + kc];
#pragma unroll
          for (int n1 = 0; n1 < 8; ++n1) v[n1] = cscale(v[n1], mv);
        }
        radix8<true>(v);
        cf* col = wk + sub * CP_PITCH + pos;
        col[0] = v[0];
#pragma unroll
        for (int m1 = 1; m1 < 8; ++m1) col[m1 * CP_BLK] = cmul(v[m1], tw1[m1]);
      }
      named_bar_sync(BAR, CP_T);
      if (tid == 0) {
        if (sg == n_rounds - 1 && n_issued < my_count) {      // every pass-1 read of the slot is done: refill it
          coltma_issue(map, cp, tb.band[iss.i], iss.fl, raw + (size_t)slot * CT_SLOT_BYTES, &full[slot]);
          ++n_issued; iss.next();
        }
        // deferred publication, once per (team, slice): the T stores of the previous slice's items are behind the barrier
        // just passed and have had a pass 1 to drain, so this one fence (the barrier -> fence -> atomic pattern of a
        // grid-wide sync) is short and stalls one warp only.  A fence by every thread after every item cost 0.2 ms per
        // step, and one atomic per item (all CTAs work on the same slice: 44 k atomics on one address) another 0.06 ms.
        if (pend_count > 0 && pend_slice != sl) {
          __threadfence();
          atomicAdd(cp.done + pend_slice, pend_count);
          pend_count = 0;
        }
      }
      // ---- pass 2: radix-8 over n2 (stride 10), twiddle w80^{n3 * m2} ----
      if (kc < ncols) {
        cf* col = wk + sub * CP_PITCH + base2;
        cf v[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) v[n2] = col[n2 * 10];
        radix8<true>(v);
        col[0] = v[0];
#pragma unroll
        for (int m2 = 1; m2 < 8; ++m2) col[m2 * 10] = cmul(v[m2], tw2[m2]);
      }
      named_bar_sync(BAR, CP_T);
      // ---- pass 3: radix-10 over n3 (contiguous), crop/shift/flip on the way out ----
      const int kc3 = CT_G * sg + sub3;
      if (tid < 128 && kc3 < ncols) {
        const float4* col4 = reinterpret_cast<const float4*>(wk + sub3 * CP_PITCH + base3);
        cf v[10];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const float4 t = col4[q];
          v[2 * q] = cf_make(t.x, t.y);
          v[2 * q + 1] = cf_make(t.z, t.w);
        }
        radix10<true>(v);
        cf* dst = cp.T + ((long long)t_frame * cp.n_act + j0 + kc3) * cp.ohp;
#pragma unroll
        for (int m3 = 0; m3 < 10; ++m3)
          if (rr3[m3] >= 0) dst[rr3[m3]] = v[m3];
      }
      // (two work buffers: no barrier -- the next round writes the OTHER buffer, and the round after that comes behind
      // two team barriers that every thread only passes after its pass 3 of this round)
    }
    if (cp.done) { pend_slice = sl; ++pend_count; }
    if (++oslot == owned) { oslot = 0; ophase ^= 1; }
  }
  if (cp.done) {             // the last slice of this team
    named_bar_sync(BAR, CP_T);
    if (tid == 0 && pend_count > 0) { __threadfence(); atomicAdd(cp.done + pend_slice, pend_count); }
  }
}

// Stand-alone kernel: NT transform teams of CP_T threads, each feeding its own raw slots.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT * CP_T, MINB) colpass640_tma_kernel(const __grid_constant__ CUtensorMap map, ColTmaParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem0);
  __shared__ unsigned long long full[CT_MAX_SLOTS];
  unsigned char* raw = smem0 + ((1024u - (ct_s32(smem0) & 1023u)) & 1023u);
  cf* work = reinterpret_cast<cf*>(raw + (size_t)p.n_slots * CT_SLOT_BYTES);
  int* tabs = reinterpret_cast<int*>(work + (size_t)NT * p.work_bufs * CT_WORK_CF);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < p.n_slots; ++s) ct_mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const ColTmaTables tb = coltma_load_tables(p, tabs, tid, NT * CP_T);
  __syncthreads();
  const int n_items = p.cp.n_frames * p.cp.n_groups;
  const int first = blockIdx.x;
  if (first >= n_items) return;
  const int count = (n_items - first + gridDim.x - 1) / gridDim.x;
  const int team = tid / CP_T, t = tid - team * CP_T;
  cf* wk = work + (size_t)team * p.work_bufs * CT_WORK_CF;
  coltma_team(1 + team, &map, p, tb, raw, wk, full, t, team, NT, first, gridDim.x, count);
}

}  // namespace mriacl
#endif  // !MRIACL_EMU
