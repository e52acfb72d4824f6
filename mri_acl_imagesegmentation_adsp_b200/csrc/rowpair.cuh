// rowpair.cuh -- row pass organised by OUTPUT PAIR: no intermediate in shared memory, no block-wide barrier.
//
// Row transform of length N = P * Q (368 = 23 * 16) with n = Q n1 + n2, k = k1 + P k2:
//     X[k1 + P k2] = sum_{n2} w_Q^{n2 k2} * ( w_N^{n2 k1} * sum_{n1} x[Q n1 + n2] w_P^{n1 k1} )  =: FFT_Q over n2 of Y[k1][n2].
// A thread owns one image row and one PAIR of first-stage outputs (k1, P - k1).  For that pair it builds all
// Q values Y[k1][.] and Y[P-k1][.] in registers -- dense residues (n2 = 0 mod STEP, the equispaced columns)
// by the symmetric direct DFT, which yields k1 and P - k1 together as A +- iB; sparse residues (the few
// extra low-frequency columns) by direct accumulation with tabulated twiddles -- runs the two Q-point FFTs
// in registers and adds |X|^2 to its 2 Q private accumulators.  Compared with the cooperative kernels in
// rowpass.cuh / rowpass16.cuh this trades a little redundant arithmetic (every pair re-reads the staged
// tile) for: no [P][Q][rows] intermediate in shared memory, no transposition, no index lookups (the staged
// tile is laid out residue-major so every operand sits at an immediate offset), no block barriers -- the
// only synchronisation is the full / empty hand-shake of the tile ring with the stager warp.
//
// Team = 6 compute warps + 1 stager warp.  A compute warp is two half-warps: lanes 0-15 are the 16 rows of
// the tile for pair 2w, lanes 16-31 the same 16 rows for pair 2w+1; both halves read the same tile words
// (shared-memory broadcast), so a 64-bit load costs one wavefront.  The stager warp does nothing but copy the
// next coils' [n_act][16 rows] blocks of T (cp.async, 16 bytes) into residue-major slots, up to n_buf tiles
// ahead; a tile's arrival is signalled by an mbarrier that the copies themselves complete, so the stager never
// waits for data.  When a tile has landed the compute threads first turn every symmetric couple
// (x[n1], x[P-n1]) of the dense residues into (a, b) = (sum, difference) in place -- once per tile, shared by
// all pairs -- and then run their pair's transform.
//
// The plan (plan.h, build_rowpair_plan) may rotate the logical index by a constant so that the dense
// residues are exactly n2 = 0 mod STEP: a rotation of the input index multiplies X[k] by a unit phase,
// which the magnitude discards.
#pragma once
#include "butterflies.cuh"
#include "rowpass.cuh"

namespace mriacl {

constexpr int RPP_ROWS = 16;          // rows per tile
constexpr int RPP_CW = 6;             // compute warps per team
constexpr int RPP_CT = RPP_CW * 32;   // compute threads
constexpr int RPP_TT = RPP_CT + 32;   // team threads (+ stager warp)
constexpr int RPP_NPAIR = 12;         // (P + 1) / 2 for P = 23: pair 0 is k1 = 0 alone

struct RowPairParams {
  const cf* T;               // [n_slices * C][n_act][ohp]
  int n_act, oh, ohp;
  const int* slot_of_j;      // [n_act] staged slot of active column j
  const int* zero_slots;     // [n_zero] slots of the dense region that no column fills (re-zeroed every stage)
  int n_zero;
  const float* tables;       // coef[12][11] float2 | dtw[ND][12] float4 | sptw[(Q-ND)*NE][12] float4
  int n_slots;
  float* out;                // [n_slices][oh][ow]
  float* partials;           // [n_slices][n_tiles][3] (count, mean, M2) or nullptr
  int ow, col0;
  int C;                     // coil frames per slice (the pair kernel serves A == 1)
  float scale;
  int n_slices, n_tiles;     // n_tiles = ceil(oh / 16)
  int n_buf;                 // tile ring depth (2 or 3)
  const int* done;           // optional [n_slices]: column-pass items finished per slice (co-resident schedule)
  int done_target;
  int* error_flag;
  int ring;                  // co-resident schedule: slice s lives in T slot s % ring (0 = no ring)
  int* rows_done;            // [n_slices] += 1 per consumed tile (the column teams wait on it before reusing a slot)
};

template <int P, int Q, int STEP, int NE> struct RowPairLayout {
  static constexpr int ND = Q / STEP;                 // dense residues
  static constexpr int NSP = Q - ND;                  // sparse residues
  static constexpr int DENSE_SLOTS = ND * P;
  static constexpr int N_SLOTS = DENSE_SLOTS + NSP * NE;
  static constexpr int HP = (P - 1) / 2;
  static constexpr int COEF_FLOATS = RPP_NPAIR * HP * 2;
  static constexpr int DTW_FLOATS = ND * RPP_NPAIR * 4;
  static constexpr int SPTW_FLOATS = NSP * NE * RPP_NPAIR * 4;
  static constexpr int TABLE_FLOATS = COEF_FLOATS + DTW_FLOATS + SPTW_FLOATS;
  static constexpr int STAGE_ELEMS = N_SLOTS * RPP_ROWS;                 // complex elements of one staged tile
  static constexpr int ACC_BYTES = RPP_CT * 2 * Q * 4;                   // 2 Q floats per compute thread
  __host__ __device__ static constexpr int smem_bytes(int n_buf, int n_act, int n_zero) {
    return n_buf * STAGE_ELEMS * 8 + ACC_BYTES + TABLE_FLOATS * 4 + ((n_act + n_zero + 3) / 4) * 16;
  }
  // n2 of the si-th sparse residue (the si-th n2 that is not a multiple of STEP)
  __host__ __device__ static constexpr int sparse_n2(int si) { return si + si / (STEP - 1) + 1; }
};

// ---- stage 1 for one coil: Ya[n2] = Y[k1][n2], Yb[n2] = Y[P - k1][n2] from the staged tile ---------------------
template <int P, int Q, int STEP, int NE>
__device__ __forceinline__ void rowpair_stage1(const cf* tb /* tile + row */, const float2* coef /* + pair * HP */,
                                               const float4* dtw /* + pair */, const float4* sptw /* + pair */,
                                               cf* Ya, cf* Yb) {
  using L = RowPairLayout<P, Q, STEP, NE>;
  constexpr int ND = L::ND, HP = L::HP;
  // dense residues: all of them advance together so that 2 ND independent FMA chains are in flight
  cf A[ND], B[ND];
  static_for<ND>([&](auto dd) { constexpr int d = dd.value; A[d] = tb[(d * P) * RPP_ROWS]; });
  static_for<HP>([&](auto nn) {
    constexpr int n = nn.value + 1;
    const float2 cs = coef[n - 1];
    static_for<ND>([&](auto dd) {
      constexpr int d = dd.value;
      const cf a = tb[(d * P + n) * RPP_ROWS], b = tb[(d * P + P - n) * RPP_ROWS];
      A[d] = pk_fma(a, bc(cs.x), A[d]);
      if constexpr (n == 1) B[d] = pk_mul(b, bc(cs.y)); else B[d] = pk_fma(b, bc(cs.y), B[d]);
    });
  });
  static_for<ND>([&](auto dd) {
    constexpr int d = dd.value;
    const cf iB = mul_i<true>(B[d]);
    const cf plus = cadd(A[d], iB), minus = csub(A[d], iB);
    if constexpr (d == 0) { Ya[0] = plus; Yb[0] = minus; }
    else {
      const float4 w = dtw[d * RPP_NPAIR];
      Ya[d * STEP] = cmul(plus, cf_make(w.x, w.y));
      Yb[d * STEP] = cmul(minus, cf_make(w.z, w.w));
    }
  });
  // sparse residues: NE tabulated entries each (absent entries point at an all-zero slot)
  static_for<L::NSP>([&](auto ss) {
    constexpr int si = ss.value;
    constexpr int n2 = L::sparse_n2(si);
    cf ya, yb;
    static_for<NE>([&](auto ee) {
      constexpr int e = ee.value;
      const cf x = tb[(L::DENSE_SLOTS + si * NE + e) * RPP_ROWS];
      const float4 w = sptw[(si * NE + e) * RPP_NPAIR];
      const cf ix = mul_i<true>(x);
      if constexpr (e == 0) {
        ya = pk_fma(ix, bc(w.y), pk_mul(x, bc(w.x)));
        yb = pk_fma(ix, bc(w.w), pk_mul(x, bc(w.z)));
      } else {
        ya = pk_fma(ix, bc(w.y), pk_fma(x, bc(w.x), ya));
        yb = pk_fma(ix, bc(w.w), pk_fma(x, bc(w.z), yb));
      }
    });
    Ya[n2] = ya; Yb[n2] = yb;
  });
}

// team-wide sum over the RPP_CT compute threads (named barrier `bar`, scratch red[RPP_CW])
__device__ __forceinline__ float rpp_team_sum(float v, float* red, int t, int bar) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((t & 31) == 0) red[t >> 5] = v;
  named_bar_sync(bar, RPP_CT);
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < RPP_CW; ++w) s += red[w];
  named_bar_sync(bar, RPP_CT);
  return s;
}

// shared-memory carve-up of one team
template <int P, int Q, int STEP, int NE> struct RowPairSmem {
  using L = RowPairLayout<P, Q, STEP, NE>;
  cf* stage; float4* acc; float* tab; int* slot_of_j; int* zero_slots;
  __device__ __forceinline__ RowPairSmem(void* base, const RowPairParams& p) {
    stage = reinterpret_cast<cf*>(base);
    acc = reinterpret_cast<float4*>(stage + (size_t)p.n_buf * L::STAGE_ELEMS);
    tab = reinterpret_cast<float*>(reinterpret_cast<char*>(acc) + L::ACC_BYTES);
    slot_of_j = reinterpret_cast<int*>(tab + L::TABLE_FLOATS);
    zero_slots = slot_of_j + p.n_act;
  }
};

// Tables into shared memory and the staged tiles zeroed (slots no column fills must read as zero); called by
// all RPP_TT threads of the team, followed by a team barrier at the caller.
template <int P, int Q, int STEP, int NE>
__device__ __forceinline__ void rowpair_setup(const RowPairParams& p, void* smem_base, int t) {
  using L = RowPairLayout<P, Q, STEP, NE>;
  RowPairSmem<P, Q, STEP, NE> S(smem_base, p);
  for (int i = t; i < L::TABLE_FLOATS; i += RPP_TT) S.tab[i] = p.tables[i];
  for (int i = t; i < p.n_act; i += RPP_TT) S.slot_of_j[i] = p.slot_of_j[i];
  for (int i = t; i < p.n_zero; i += RPP_TT) S.zero_slots[i] = p.zero_slots[i];
  for (int i = t; i < p.n_buf * L::STAGE_ELEMS; i += RPP_TT) S.stage[i] = cf_make(0.f, 0.f);
}

// One work item = (slice, 16-row tile).  `k_stage` is the team's running stage counter (tile ring position);
// it is advanced identically by the compute threads and by the stager.  Barrier ids: bar0 + b (FULL, b < n_buf),
// bar0 + 3 + b (EMPTY), bar0 + 6 (compute threads only).
template <int P, int Q, int STEP, int NE>
__device__ __forceinline__ void rowpair_compute_item(const RowPairParams& p, void* smem_base, FullBarrier* full_bar,
                                                     int item, int t, int bar0, float* red, int& k_stage) {
  using L = RowPairLayout<P, Q, STEP, NE>;
  constexpr int N = P * Q;
  RowPairSmem<P, Q, STEP, NE> S(smem_base, p);
  const int lane = t & 31, warp = t >> 5;
  const int r = lane & 15, pair = 2 * warp + (lane >> 4);
  const float2* coef = reinterpret_cast<const float2*>(S.tab) + pair * L::HP;
  const float4* dtw = reinterpret_cast<const float4*>(S.tab + L::COEF_FLOATS) + pair;
  const float4* sptw = reinterpret_cast<const float4*>(S.tab + L::COEF_FLOATS + L::DTW_FLOATS) + pair;
  float4* acc = S.acc + (size_t)warp * (2 * Q / 4) * 32 + lane;     // my quad q at acc[q * 32]
  const int s = item / p.n_tiles, tile = item - s * p.n_tiles;

  float4 fin[2 * Q / 4];
  for (int f = 0; f < p.C; ++f, ++k_stage) {
    const int buf = k_stage % p.n_buf;
    cf Ya[Q], Yb[Q];
    cf* tile_sm = S.stage + (size_t)buf * L::STAGE_ELEMS;
    full_wait(&full_bar[buf], (k_stage / p.n_buf) & 1, bar0 + buf, RPP_TT);     // tile f has landed
    if (p.n_zero) {     // slots of the dense region that no column fills hold the previous tile's a / b: clear them
      for (int i = t; i < p.n_zero * RPP_ROWS; i += RPP_CT) tile_sm[S.zero_slots[i >> 4] * RPP_ROWS + (i & 15)] = cf_make(0.f, 0.f);
      named_bar_sync(bar0 + 6, RPP_CT);
    }
    // (x[n1], x[P - n1]) -> (a, b) for every dense residue, in place, shared by all pairs
    for (int i = t; i < L::ND * L::HP * RPP_ROWS; i += RPP_CT) {
      const int rr = i & 15, c = i >> 4;
      const int d = c / L::HP, n = c - d * L::HP + 1;
      cf* pa = tile_sm + (d * P + n) * RPP_ROWS + rr;
      cf* pb = tile_sm + (d * P + P - n) * RPP_ROWS + rr;
      const cf x1 = *pa, x2 = *pb;
      *pa = cadd(x1, x2);
      *pb = csub(x1, x2);
    }
    named_bar_sync(bar0 + 6, RPP_CT);
    rowpair_stage1<P, Q, STEP, NE>(tile_sm + r, coef, dtw, sptw, Ya, Yb);
    named_bar_arrive(bar0 + 3 + buf, RPP_TT);                         // every tile word is in registers: refill may start
    fft16<true>(Ya);
    fft16<true>(Yb);
    const bool first = f == 0, last = f + 1 == p.C;
#pragma unroll
    for (int q = 0; q < 2 * Q / 4; ++q) {
      const cf* v = (q < Q / 4 ? Ya : Yb) + 4 * (q % (Q / 4));
      float4 a = first ? make_float4(0.f, 0.f, 0.f, 0.f) : acc[q * 32];
      a.x = cnorm2_acc(v[0], a.x); a.y = cnorm2_acc(v[1], a.y); a.z = cnorm2_acc(v[2], a.z); a.w = cnorm2_acc(v[3], a.w);
      if (last) fin[q] = a; else acc[q * 32] = a;
    }
  }

  // ---- epilogue: sqrt, scale, fftshift + crop into a [16][ow + 1] tile that aliases the accumulators ----
  const int bar_c = bar0 + 6;
  const int opitch = p.ow + 1;
  float* osm = reinterpret_cast<float*>(S.acc);
  named_bar_sync(bar_c, RPP_CT);                                      // all accumulator reads done
  const int k1a = pair, k1b = P - pair;
#pragma unroll
  for (int q = 0; q < 2 * Q / 4; ++q) {
    const float vv[4] = {fin[q].x, fin[q].y, fin[q].z, fin[q].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k2 = 4 * (q % (Q / 4)) + i;
      const bool second = q >= Q / 4;
      const int cc = phys_of_logical((second ? k1b : k1a) + P * k2, N) - p.col0;
      if ((!second || pair != 0) && cc >= 0 && cc < p.ow) osm[r * opitch + cc] = sqrtf(vv[i]) * p.scale;
    }
  }
  named_bar_sync(bar_c, RPP_CT);
  const int rows_here = min(RPP_ROWS, p.oh - tile * RPP_ROWS);
  const int n_here = rows_here * p.ow;
  float* dst = p.out + ((long long)s * p.oh + tile * RPP_ROWS) * p.ow;
  float lsum = 0.f;
  for (int e = t; e < n_here; e += RPP_CT) {
    const int rr = e / p.ow, cc = e - rr * p.ow;
    const float v = osm[rr * opitch + cc];
    dst[e] = v;
    lsum += v;
  }
  if (p.partials) {
    const float mean = rpp_team_sum(lsum, red, t, bar_c) / (float)n_here;
    float lq = 0.f;
    for (int e = t; e < n_here; e += RPP_CT) {
      const int rr = e / p.ow, cc = e - rr * p.ow;
      const float d = osm[rr * opitch + cc] - mean;
      lq = fmaf(d, d, lq);
    }
    const float m2 = rpp_team_sum(lq, red, t, bar_c);
    if (t == 0) {
      float* q = p.partials + ((long long)s * p.n_tiles + tile) * 3;
      q[0] = (float)n_here; q[1] = mean; q[2] = m2;
    }
  }
  named_bar_sync(bar_c, RPP_CT);                                      // tile consumed before the next item's first store
  if (p.rows_done && t == 0) atomicAdd(p.rows_done + s, 1);           // every T word of this tile has been copied out
}

// stager warp: the coil tiles of one item.  Copies only; the mbarrier of the buffer flips when they have landed.
template <int P, int Q, int STEP, int NE>
__device__ __forceinline__ void rowpair_stage_item(const RowPairParams& p, void* smem_base, FullBarrier* full_bar,
                                                   int item, int lane, int bar0, int& k_stage) {
  using L = RowPairLayout<P, Q, STEP, NE>;
  RowPairSmem<P, Q, STEP, NE> S(smem_base, p);
  const int s = item / p.n_tiles, tile = item - s * p.n_tiles;
  const long long frame_elems = (long long)p.n_act * p.ohp;
  const cf* Tit = p.T + (long long)(p.ring ? s % p.ring : s) * p.C * frame_elems + tile * RPP_ROWS;
  const int n_copies = p.n_act * (RPP_ROWS / 2);          // 16-byte pieces: 8 per column
  if (p.done) {     // co-resident schedule: the slice's columns are written by column teams running beside us
    if (lane == 0 && !rp_wait_count(p.done + s, p.done_target, p.error_flag) && p.error_flag) atomicAdd(p.error_flag, 1);
    __syncwarp();
  }
  for (int f = 0; f < p.C; ++f, ++k_stage) {
    const int buf = k_stage % p.n_buf;
    if (k_stage >= p.n_buf) named_bar_sync(bar0 + 3 + buf, RPP_TT);   // the compute warps are done with this buffer
    cf* tile_sm = S.stage + (size_t)buf * L::STAGE_ELEMS;
    const cf* src = Tit + (long long)f * frame_elems;
    for (int i = lane; i < n_copies; i += 32) {
      const int j = i >> 3, part = i & 7;
      cp_async16(tile_sm + S.slot_of_j[j] * RPP_ROWS + 2 * part, src + (long long)j * p.ohp + 2 * part);
    }
    full_signal_async(&full_bar[buf], bar0 + buf, RPP_TT);
  }
}

// Stand-alone kernel: one team per CTA, persistent over the (slice, tile) items.
template <int P, int Q, int STEP, int NE, int MINB>
__global__ void __launch_bounds__(RPP_TT, MINB) rowpair_kernel(RowPairParams p) {
  MRIACL_DYN_SMEM(unsigned char, smem);
  __shared__ float red[RPP_CW];
  __shared__ FullBarrier full_bar[3];
  const int t = threadIdx.x;
  if (t < 3) full_init(&full_bar[t], 32);
  rowpair_setup<P, Q, STEP, NE>(p, smem, t);
  __syncthreads();
  const int n_items = p.n_slices * p.n_tiles;
  int k_stage = 0;
  if (t < RPP_CT) {
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
      rowpair_compute_item<P, Q, STEP, NE>(p, smem, full_bar, item, t, 1, red, k_stage);
  } else {
    for (int item = blockIdx.x; item < n_items; item += gridDim.x)
      rowpair_stage_item<P, Q, STEP, NE>(p, smem, full_bar, item, t - RPP_CT, 1, k_stage);
    cp_async_wait<0>();
  }
}

}  // namespace mriacl
