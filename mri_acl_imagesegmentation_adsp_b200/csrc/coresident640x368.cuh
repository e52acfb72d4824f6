// coresident640x368.cuh -- the whole knee plan as ONE persistent kernel with both passes resident on every SM.
//
// One CTA per SM, 18 warps:
//   warps 0-5   column team  (colpass640.cuh: producer warp + 5 transform warps, two item buffers) -- streams
//               k-space from HBM, writes the intermediate T and bumps a per-slice counter per finished item;
//   warps 6-17  row team     (rowpass16.cuh, 12 warps on a named barrier) -- takes (slice, 16-row tile) items in
//               slice order, waits until the slice's counter says every column group has landed, reads T back
//               while it is still in L2, and writes the image tile.  The team that finishes the LAST tile of a
//               slice also normalises the slice (mean / unbiased std from the tiles' (n, mean, M2) partials), so
//               the stage needs no further launch.
// Every CTA is resident (grid <= number of SMs, one CTA per SM by shared memory), column teams never wait for
// anybody, and both kinds of items are taken in increasing slice order, so the waits cannot deadlock.  Compared
// with two concurrent launches (the "overlapped" schedule) co-residency is guaranteed rather than hoped for: the
// HBM-bound gather and the issue/latency-bound row transform share each SM's issue slots for the whole step.
#pragma once
#include "colpass640.cuh"
#include "rowpass16.cuh"
#include "rowpair.cuh"
#include "colpass640_tma.cuh"

namespace mriacl {

constexpr int KC_ROW_W = 12;
constexpr int KC_ROW_T = KC_ROW_W * 32;
constexpr int KC_T = CP_WS_T + KC_ROW_T;      // 576
constexpr int KC_BAR_ROW = 6;                 // named barrier of the row team (the column team uses 1-5)

struct CoresParams {
  ColPassParams cp;          // cp.done = per-slice counters (zeroed before the launch)
  RowPass16Params rp;        // rp.done = the same counters; rp.tiles_done / mean_std / eps / normalize: fused normalisation
};

template <int P, int Q, int G = CP_G>
__global__ void __launch_bounds__(KC_T, 1) knee_coresident_kernel(CoresParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem);
  __shared__ FullBarrier full_bar[2];
  __shared__ float red[KC_ROW_W];
  __shared__ float s_stat[2];
  __shared__ int s_ready, s_last;
  const int tid = threadIdx.x;
  if (tid == 0) { full_init(&full_bar[0], 32); full_init(&full_bar[1], 32); }
  __syncthreads();

  if (tid < CP_WS_T) {
    // ------------------------------ column team ------------------------------
    const int n_items = p.cp.n_frames * p.cp.n_groups;
    int uses[2] = {0, 0};
    const int first = blockIdx.x;
    if (first < n_items)
      colpass_ws_run<G, 0>(p.cp, reinterpret_cast<cf*>(smem), full_bar, tid, first, gridDim.x,
                     (n_items - first + gridDim.x - 1) / gridDim.x, uses);
    return;
  }

  // ------------------------------ row team ------------------------------
  const int t = tid - CP_WS_T;
  unsigned char* rsm = smem + 2 * G * CP_PITCH * 8;
  const RowPass16Params& r = p.rp;
  {
    Rp16Smem<P, Q> S(rsm, r);
    rp16_load_tables<KC_ROW_T>(r, S.sptw, S.sch, S.tbuf, t);
  }
  rp16_sync<KC_BAR_ROW, KC_ROW_T>();
  const int n_items = r.n_slices * r.n_tiles;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    rowpass16_item<P, Q, KC_ROW_W, KC_BAR_ROW>(r, rsm, item, t, red, &s_ready);
    if (r.done && !s_ready) return;                  // the wait for the slice timed out (error flag is set)
    if (r.tiles_done) rowpass16_finish_slice<KC_ROW_W, KC_BAR_ROW>(r, item, t, s_stat, &s_last);
  }
}

// ---- the same idea with the pair row pass (rowpair.cuh) as the row team: 6 compute warps + stager, no team-wide
// barrier in the coil loop, so the row team needs fewer warps to keep its pipes busy: 13 warps per CTA in all.
constexpr int KP_T = CP_WS_T + RPP_TT;        // 416
constexpr int KP_BAR0 = 6;                    // row team: FULL 6-8 (emulator only), EMPTY 9-11, compute 12

struct CoresPairParams {
  ColPassParams cp;
  RowPairParams rp;
};

template <int P, int Q, int STEP, int NE>
__global__ void __launch_bounds__(KP_T, 1) knee_coresident_pair_kernel(CoresPairParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem);
  __shared__ FullBarrier full_bar[2];
  __shared__ FullBarrier row_full[3];
  __shared__ float red[RPP_CW];
  const int tid = threadIdx.x;
  if (tid == 0) { full_init(&full_bar[0], 32); full_init(&full_bar[1], 32); }
  if (tid >= 32 && tid < 35) full_init(&row_full[tid - 32], 32);
  unsigned char* rsm = smem + CP_SMEM_BYTES_DB;
  if (tid >= CP_WS_T) rowpair_setup<P, Q, STEP, NE>(p.rp, rsm, tid - CP_WS_T);
  __syncthreads();

  if (tid < CP_WS_T) {
    const int n_items = p.cp.n_frames * p.cp.n_groups;
    int uses[2] = {0, 0};
    const int first = blockIdx.x;
    if (first < n_items)
      colpass_ws_run(p.cp, reinterpret_cast<cf*>(smem), full_bar, tid, first, gridDim.x,
                     (n_items - first + gridDim.x - 1) / gridDim.x, uses);
    return;
  }
  const int t = tid - CP_WS_T;
  const int n_items = p.rp.n_slices * p.rp.n_tiles;
  int k_stage = 0;
  if (t < RPP_CT) {
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
      rowpair_compute_item<P, Q, STEP, NE>(p.rp, rsm, row_full, item, t, KP_BAR0, red, k_stage);
  } else {
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
      rowpair_stage_item<P, Q, STEP, NE>(p.rp, rsm, row_full, item, t - RPP_CT, KP_BAR0, k_stage);
    cp_async_wait<0>();
  }
}

// ---- warp-group register split ------------------------------------------------------------------------------
// Same two teams, 20 warps per CTA laid out on warp-group (4-warp) boundaries so that `setmaxnreg` can move
// registers from the row team to the column team: the kernel is launched with 96 registers per thread (all an
// 18-20-warp CTA can have), the three row warp groups shrink to 80 (what the row pass needs, spill-free) and the two
// column warp groups grow to 120 (the column transform wants 115).
//   threads   0-159  column transform warps      160-191 gather producer      192-255 idle (they only take part in
//   the register hand-over)                      256-639 row team (12 warps, named barrier)
constexpr int KS_T = 640, KS_ROW0 = 256;

template <int REGS> __device__ __forceinline__ void kc_reg_inc() {
#if !defined(MRIACL_EMU)
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
#endif
}
template <int REGS> __device__ __forceinline__ void kc_reg_dec() {
#if !defined(MRIACL_EMU)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
#endif
}

template <int P, int Q>
__global__ void __launch_bounds__(KS_T, 1) knee_coresident_split_kernel(CoresParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem);
  __shared__ FullBarrier full_bar[2];
  __shared__ float red[KC_ROW_W];
  __shared__ float s_stat[2];
  __shared__ int s_ready, s_last;
  const int tid = threadIdx.x;
  if (tid == 0) { full_init(&full_bar[0], 32); full_init(&full_bar[1], 32); }
  __syncthreads();

  if (tid < KS_ROW0) {
    kc_reg_inc<120>();
    if (tid >= CP_WS_T) return;
    const int n_items = p.cp.n_frames * p.cp.n_groups;
    int uses[2] = {0, 0};
    const int first = blockIdx.x;
    if (first < n_items)
      colpass_ws_run(p.cp, reinterpret_cast<cf*>(smem), full_bar, tid, first, gridDim.x,
                     (n_items - first + gridDim.x - 1) / gridDim.x, uses);
    return;
  }
  kc_reg_dec<80>();
  const int t = tid - KS_ROW0;
  unsigned char* rsm = smem + CP_SMEM_BYTES_DB;
  const RowPass16Params& r = p.rp;
  {
    Rp16Smem<P, Q> S(rsm, r);
    rp16_load_tables<KC_ROW_T>(r, S.sptw, S.sch, S.tbuf, t);
  }
  rp16_sync<KC_BAR_ROW, KC_ROW_T>();
  const int n_items = r.n_slices * r.n_tiles;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    rowpass16_item<P, Q, KC_ROW_W, KC_BAR_ROW>(r, rsm, item, t, red, &s_ready);
    if (r.done && !s_ready) return;
    if (r.tiles_done) rowpass16_finish_slice<KC_ROW_W, KC_BAR_ROW>(r, item, t, s_stat, &s_last);
  }
}

// ---- two column teams + one row team -------------------------------------------------------------------------------
// The column pass needs TWO gather streams per SM to keep HBM busy (stand-alone it runs two CTAs per SM; one team per SM,
// as in the kernels above, takes 0.44-0.50 ms on its own), and it needs >= 60 KB of L1 for the in-flight lines of its
// 8-byte LDGSTS gather, i.e. the CTA must stay inside the 196 KB shared-memory configuration.  Both fit
// when the column items shrink to FOUR columns: two teams x two 23 KB buffers = the 92 KB one eight-column team used,
// and a transform thread keeps two butterflies in flight instead of four (~85 registers), so 24 warps fit the register file.
//   threads   0-191  column team A (items 2 b, 2 b + 2 grid, ...)    192-383  column team B (items 2 b + 1, ...)
//   threads 384-767  row team (12 warps on a named barrier), fed by the per-slice counters the column teams bump
constexpr int K2_T = 2 * CP_WS_T + KC_ROW_T;     // 768
constexpr int K2_G = 4;
constexpr int K2_COL_SMEM = 2 * 2 * K2_G * CP_PITCH * 8;     // two teams, two buffers each
constexpr int K2_BAR_ROW = 11;                   // team A uses named barriers 1-5, team B 6-10

template <int P, int Q>
__global__ void __launch_bounds__(K2_T, 1) knee_coresident2_kernel(CoresParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem);
  __shared__ FullBarrier full_bar[4];
  __shared__ float red[KC_ROW_W];
  __shared__ float s_stat[2];
  __shared__ int s_ready, s_last;
  const int tid = threadIdx.x;
  if (tid < 4) full_init(&full_bar[tid], 32);
  __syncthreads();

  if (tid < 2 * CP_WS_T) {
    const int team = tid / CP_WS_T, t = tid - team * CP_WS_T;
    const int n_items = p.cp.n_frames * p.cp.n_groups;
    int uses[2] = {0, 0};
    const int first = 2 * blockIdx.x + team, stride = 2 * gridDim.x;
    cf* tsm = reinterpret_cast<cf*>(smem) + (size_t)team * 2 * K2_G * CP_PITCH;
    if (first < n_items) {
      const int count = (n_items - first + stride - 1) / stride;
      if (team == 0) colpass_ws_run<K2_G, 0>(p.cp, tsm, full_bar, t, first, stride, count, uses);
      else colpass_ws_run<K2_G, 5>(p.cp, tsm, full_bar + 2, t, first, stride, count, uses);
    }
    return;
  }
  const int t = tid - 2 * CP_WS_T;
  unsigned char* rsm = smem + K2_COL_SMEM;
  const RowPass16Params& r = p.rp;
  {
    Rp16Smem<P, Q> S(rsm, r);
    rp16_load_tables<KC_ROW_T>(r, S.sptw, S.sch, S.tbuf, t);
  }
  rp16_sync<K2_BAR_ROW, KC_ROW_T>();
  const int n_items = r.n_slices * r.n_tiles;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    rowpass16_item<P, Q, KC_ROW_W, K2_BAR_ROW>(r, rsm, item, t, red, &s_ready);
    if (r.done && !s_ready) return;
    if (r.tiles_done) rowpass16_finish_slice<KC_ROW_W, K2_BAR_ROW>(r, item, t, s_stat, &s_last);
  }
}

#ifndef MRIACL_EMU
// ---- TMA-fed column teams + one row team ---------------------------------------------------------------------------
// The gather of the kernels above needs L1 (8-byte LDGSTS.ca), and a CTA that also holds a row team leaves it 28-92 KB:
// every co-resident variant so far had a starved column side.  Here the TMA unit streams whole bands of k-space into
// raw slots (colpass640_tma.cuh): no L1, no load instructions, and the two transform teams only read shared memory.
// T still goes through global memory (a ring of `ring` slices keeps it in L2).
//   threads   0-319  two column transform teams      320-703  row team (12 warps)
constexpr int KT_COL_T = 2 * CP_T;               // 320
constexpr int KT_T = KT_COL_T + KC_ROW_T;        // 704
constexpr int KT_BAR_ROW = 6;                    // the column teams use named barriers 1 and 2

struct CoresTmaParams {
  ColTmaParams ct;           // ct.cp.done = per-slice counters (zeroed before the launch)
  RowPass16Params rp;        // rp.done = the same counters
  int* rows_done;            // [n_slices] row tiles consumed per slice (ring), or nullptr
};

template <int P, int Q>
__global__ void __launch_bounds__(KT_T, 1) knee_coresident_tma_kernel(const __grid_constant__ CUtensorMap map, CoresTmaParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem0);
  __shared__ unsigned long long full[CT_MAX_SLOTS];
  __shared__ float red[KC_ROW_W];
  __shared__ int s_ready;
  unsigned char* raw = smem0 + ((1024u - (ct_s32(smem0) & 1023u)) & 1023u);
  cf* work = reinterpret_cast<cf*>(raw + (size_t)p.ct.n_slots * CT_SLOT_BYTES);
  int* tabs = reinterpret_cast<int*>(work + 2 * p.ct.work_bufs * CT_WORK_CF);
  unsigned char* rsm = reinterpret_cast<unsigned char*>(tabs) + coltma_table_bytes(p.ct.cp.n_groups, p.ct.cp.n_act);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < p.ct.n_slots; ++s) ct_mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  ColTmaTables tb{};
  if (tid < KT_COL_T) tb = coltma_load_tables(p.ct, tabs, tid, KT_COL_T);
  __syncthreads();

  if (tid < KT_COL_T) {
    const int n_items = p.ct.cp.n_frames * p.ct.cp.n_groups;
    const int first = blockIdx.x;
    if (first >= n_items) return;
    const int count = (n_items - first + gridDim.x - 1) / gridDim.x;
    const int team = tid / CP_T;
    coltma_team(1 + team, &map, p.ct, tb, raw, work + team * p.ct.work_bufs * CT_WORK_CF, full, tid - team * CP_T, team, 2, first, gridDim.x, count);
    return;
  }

  const int t = tid - KT_COL_T;
  const RowPass16Params& r = p.rp;
  {
    Rp16Smem<P, Q> S(rsm, r);
    rp16_load_tables<KC_ROW_T>(r, S.sptw, S.sch, S.tbuf, t);
  }
  rp16_sync<KT_BAR_ROW, KC_ROW_T>();
  const int n_items = r.n_slices * r.n_tiles;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    rowpass16_item<P, Q, KC_ROW_W, KT_BAR_ROW>(r, rsm, item, t, red, &s_ready);
    if (r.done && !s_ready) return;                  // the wait for the slice timed out (error flag is set)
    if (p.rows_done && t == 0) atomicAdd(p.rows_done + item / r.n_tiles, 1);   // (the item's last T read is behind a team barrier)
  }
}

// ---- SM-role split: column SMs and row SMs -------------------------------------------------------------------------
// Every variant above shares each SM between the two passes, and each time the teams slowed each other down as much as
// the overlap gained (issue slots, the shared-memory pipe, barriers of a latency-bound row team).  Here the SMs are split
// instead: one CTA per SM (grid <= number of SMs, the CTA's shared memory allows no second one), CTAs 0 .. n_col-1 are
// COLUMN CTAs (four transform teams, each feeding its own raw slot by TMA: the TMA gather reaches 6.3-6.9 TB/s from
// 56-74 SMs, tools/microbench/gather_modes.cu), the others are ROW CTAs (two 12-warp row teams, exactly the two CTAs per
// SM of the stand-alone row pass).  Slices flow from the column SMs to the row SMs through a ring of T slots that stays
// in L2; per-slice counters in both directions.  Nobody shares an SM, so each side runs at its stand-alone speed and
// the step is max(column side, row side) + one row item of tail instead of their sum.
constexpr int KX_T = 768;
constexpr int KX_COL_TEAMS = 4;

struct SplitParams {
  ColTmaParams ct;           // ct.cp.done = per-slice counters (zeroed before the launch)
  RowPass16Params rp;        // rp.done = the same counters
  int* rows_done;            // [n_slices] row tiles consumed per slice (ring), or nullptr
  int n_col;                 // CTAs 0 .. n_col-1 are column CTAs
  int row_smem;              // bytes of one row team's shared memory (16-byte multiple)
};

template <int P, int Q>
__global__ void __launch_bounds__(KX_T, 1) knee_split_kernel(const __grid_constant__ CUtensorMap map, SplitParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem0);
  __shared__ unsigned long long full[CT_MAX_SLOTS];
  __shared__ float red[2][KC_ROW_W];
  __shared__ int s_ready[2];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < p.n_col) {
    // ------------------------------ column CTA ------------------------------
    unsigned char* raw = smem0 + ((1024u - (ct_s32(smem0) & 1023u)) & 1023u);
    cf* work = reinterpret_cast<cf*>(raw + (size_t)p.ct.n_slots * CT_SLOT_BYTES);
    int* tabs = reinterpret_cast<int*>(work + (size_t)KX_COL_TEAMS * p.ct.work_bufs * CT_WORK_CF);
    if (tid == 0) {
      for (int s = 0; s < p.ct.n_slots; ++s) ct_mbar_init(&full[s], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const ColTmaTables tb = coltma_load_tables(p.ct, tabs, tid, KX_T);
    __syncthreads();
    const int n_items = p.ct.cp.n_frames * p.ct.cp.n_groups;
    const int first = blockIdx.x;
    if (first >= n_items || tid >= KX_COL_TEAMS * CP_T) return;
    const int count = (n_items - first + p.n_col - 1) / p.n_col;
    const int team = tid / CP_T, tt = tid - team * CP_T;
    cf* wk = work + (size_t)team * p.ct.work_bufs * CT_WORK_CF;
    coltma_team(1 + team, &map, p.ct, tb, raw, wk, full, tt, team, KX_COL_TEAMS, first, p.n_col, count);
    return;
  }
  // ------------------------------ row CTA: two independent row teams ------------------------------
  const int team = tid / KC_ROW_T, t = tid - team * KC_ROW_T;
  unsigned char* rsm = smem0 + (size_t)team * p.row_smem;
  const RowPass16Params& r = p.rp;
  const int n_items = r.n_slices * r.n_tiles;
  const int first = 2 * ((int)blockIdx.x - p.n_col) + team, stride = 2 * ((int)gridDim.x - p.n_col);
  {
    Rp16Smem<P, Q> S(rsm, r);
    rp16_load_tables<KC_ROW_T>(r, S.sptw, S.sch, S.tbuf, t);
  }
  // one copy of the row-pass code for both teams (BAR = -1: named barrier 1 + team)
  rp16_sync<-1, KC_ROW_T>();
  for (int item = first; item < n_items; item += stride) {
    rowpass16_item<P, Q, KC_ROW_W, -1>(r, rsm, item, t, red[team], &s_ready[team]);
    if (r.done && !s_ready[team]) return;
    if (p.rows_done && t == 0) atomicAdd(p.rows_done + item / r.n_tiles, 1);
  }
}
#endif

}  // namespace mriacl
