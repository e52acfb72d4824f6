// fused640x368.cuh -- ONE persistent kernel for the whole knee plan: every CTA alternates between the two
// kinds of work, choosing dynamically:
//   * a ROW item (slice, 16-row tile of the row pass) whenever the next unclaimed one is ready, i.e. every
//     column group of its slice has been written (per-slice counters published by the column items);
//   * otherwise a run of consecutive COLUMN items (frame, 8 sampled columns) through the producer-warp pipeline.
// Row items are claimed with a compare-and-swap on the queue head only after their slice was seen complete and
// column items never wait, so no CTA ever blocks on another one (no co-residency assumption, no deadlock).
// The point of fusing: at any time roughly half of the resident CTAs stream k-space from HBM (memory-bound)
// while the others run the row transform (issue-bound) on data that is still in L2, so the two bounds overlap
// across the chip and the intermediate T never has to live long.
#pragma once
#include "colpass640.cuh"
#include "rowpass16.cuh"

namespace mriacl {

constexpr int FZ_T = 256;                 // 8 warps: column role uses 6 (5 compute + producer), row role all 8
constexpr int FZ_ROW_WARPS = 8;
enum { FZ_IDLE = 0, FZ_COL = 1, FZ_ROW = 2, FZ_EXIT = 3 };

struct FusedParams {
  ColPassParams cp;
  RowPass16Params rp;
  int* state;              // [0] next column item, [1] next row item (zeroed before the launch)
  int n_col_items, n_row_items;
  int col_batch;           // column items claimed at a time
  int done_target;         // column items per slice
  int row_ctas_first;      // which half of the grid prefers row items (tuning)
};

__device__ __forceinline__ int fz_ld(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

template <int P, int Q>
__global__ void __launch_bounds__(FZ_T, 2) fused640_kernel(FusedParams p) {
  MRIACL_DYN_SMEM(cf, sm);
  __shared__ FullBarrier full_bar[2];
  __shared__ int s_kind, s_first, s_count, s_ready;
  __shared__ float red[FZ_ROW_WARPS];
  const int tid = threadIdx.x;
  if (tid == 0) { full_init(&full_bar[0], 32); full_init(&full_bar[1], 32); }
  __syncthreads();
  int uses[2] = {0, 0};
  const bool prefer_row = p.row_ctas_first ? blockIdx.x < gridDim.x / 2 : blockIdx.x >= gridDim.x / 2;

  while (true) {
    if (tid == 0) {
      int kind = FZ_IDLE, first = 0, count = 0;
      auto try_row = [&]() {
        const int r = fz_ld(p.state + 1);
        if (r >= p.n_row_items) return;
        if (fz_ld(p.cp.done + r / p.rp.n_tiles) < p.done_target) return;
        if (atomicCAS(p.state + 1, r, r + 1) == r) { kind = FZ_ROW; first = r; __threadfence(); }
      };
      auto try_col = [&]() {
        if (fz_ld(p.state) >= p.n_col_items) return;
        const int c = atomicAdd(p.state, p.col_batch);
        if (c < p.n_col_items) { kind = FZ_COL; first = c; count = min(p.col_batch, p.n_col_items - c); }
      };
      // Static role preference keeps a steady mix on the chip (and on every SM): the first wave of CTAs streams
      // columns, the second wave transforms rows; each falls back to the other kind of work when its own queue
      // is empty or not ready yet, so nobody idles while there is work and nobody ever waits for anybody.
      if (prefer_row) { try_row(); if (kind == FZ_IDLE) try_col(); }
      else            { try_col(); if (kind == FZ_IDLE) try_row(); }
      if (kind == FZ_IDLE && fz_ld(p.state + 1) >= p.n_row_items) kind = FZ_EXIT;
      s_kind = kind; s_first = first; s_count = count;
    }
    __syncthreads();
    const int kind = s_kind, first = s_first, count = s_count;
    if (kind == FZ_EXIT) break;
    if (kind == FZ_COL) {
      if (tid < CP_WS_T) colpass_ws_run(p.cp, sm, full_bar, tid, first, 1, count, uses);
      uses[0] += (count + 1) >> 1;
      uses[1] += count >> 1;
    } else if (kind == FZ_ROW) {
      Rp16Smem<P, Q> S(sm, p.rp);
      rp16_load_tables<FZ_T>(p.rp, S.sptw, S.sch, S.tbuf, tid);
      __syncthreads();
      rowpass16_item<P, Q, FZ_ROW_WARPS>(p.rp, sm, first, tid, red, &s_ready);
    } else {
#if !defined(MRIACL_EMU)
      __nanosleep(400);
#endif
    }
    __syncthreads();
  }
}

}  // namespace mriacl
