// rowpass.cuh -- fused row pass: pruned P x Q inverse DFT along the (padded) phase-encode axis,
// |.|^2 accumulation over coils, sqrt, mean over averages, fftshift + centre crop as the store
// index map, and per-tile statistics for the instance normalisation.
//
// One work item = (slice, tile of 32 output rows); one persistent CTA per SM; lane = row, so every
// lane of a warp runs the same plan-driven control flow and every shared-memory access is stride-1
// across lanes (no bank conflicts anywhere in this kernel).  For each coil frame:
//   prefetch the frame's [n_act][32 rows] block of the intermediate T into shared memory with
//            cp.async, one frame ahead (double buffered when it fits), so HBM/L2 latency never
//            sits in front of the arithmetic;
//   stage 1  (decimation in time, n = Q n1 + n2): for each residue n2 the P-point DFT over n1 of the
//            SAMPLED columns only.  The host plan classifies each residue as dense (symmetric direct
//            DFT with immediate constants, optionally split in two halves of its outputs) or sparse
//            (<= 6 sampled columns: direct accumulation with w_N^{n k1}) and balances the pieces over
//            the warps;
//   stage 2  for each k1 the Q-point FFT over n2 in registers -> X[k1 + P k2]; |X|^2 is added to
//            per-thread accumulators that stay in registers across all coils.
// Width 368 = 23 x 16 is the knee case; the template is general in (P odd, Q = 16).
#pragma once
#include "butterflies.cuh"

namespace mriacl {

constexpr int RP_ROWS = 32;        // rows per tile (= lanes)
constexpr int RP_MAX_SPARSE = 6;   // a residue with more sampled columns than this is "dense"
constexpr int RP_SCHED_MAX = 1024; // ints of schedule kept in shared memory
constexpr int RP_SPTW_MAX = 16 * RP_MAX_SPARSE * 24;   // complex twiddles of the sparse residues (worst case)

struct RowPassParams {
  const cf* T;           // [n_slices*A*C][n_act][ohp]
  int n_act, oh, ohp;    // ohp = row pitch of T (multiple of 32)
  const int* sched;      // warp schedule, see plan.h
  int sched_len;
  const cf* sptw;        // [sptw_len] twiddle rows of the sparse residues, pitch 24
  int sptw_len;
  const cf* tw;          // w_N^k = exp(+2 pi i k / N)
  float* out;            // [n_slices][oh][ow]  (already offset to the first slice of the launch)
  float* partials;       // [n_slices][n_tiles][3] (count, mean, M2) or nullptr
  int ow, col0;
  int A, C;
  float scale;           // 1 / sqrt(H * N)
  int n_slices, n_tiles;
  int n_buf;             // 1 or 2 prefetch buffers
  const int* done;       // optional [n_slices]: items of the slice finished by a concurrently running column pass
  int done_target;       // ... the slice is ready when done[s] == done_target
  int* error_flag;       // set to 1 if the wait for a slice timed out
};

// shared memory: Y [P][Q][32] | twiddles [N] | sparse twiddle rows | schedule | T buffers | (A > 1) average tile
__host__ __device__ inline int rp_round16(int v) { return (v + 15) / 16 * 16; }
inline int rowpass_smem_bytes(int P, int Q, int sptw_len, int sched_len, int n_act, int n_buf, int ow, int A) {
  return P * Q * RP_ROWS * 8 + P * Q * 8 + rp_round16(sptw_len * 8) + rp_round16(sched_len * 4) +
         n_buf * n_act * RP_ROWS * 8 + (A > 1 ? RP_ROWS * (ow + 1) * 4 : 0);
}

// spin (one thread) until a concurrently running producer has published `target`.  Bounded (~50 ms): a scheduling
// surprise ends in a trapped kernel (sticky CUDA error -> MRIACL_ERR_CUDA from this and every later call), never in a
// hung device and never in a silently unwritten tile.
__device__ __forceinline__ bool rp_wait_count(const int* counter, int target, int* error_flag) {
#if defined(MRIACL_EMU)
  // emulator: producers may be sibling threads of the same CTA (co-resident kernel), so poll for a while
  const auto t0 = std::chrono::steady_clock::now();      // emulator: wall-clock bound (CUDA threads are OS threads here)
  while (std::chrono::steady_clock::now() - t0 < std::chrono::seconds(300)) {
    if (reinterpret_cast<const std::atomic<int>*>(counter)->load() >= target) return true;
    std::this_thread::yield();
  }
  return false;
#else
  const volatile int* c = counter;
  const volatile int* err = error_flag;
  for (int spin = 0; spin < (1 << 18); ++spin) {
    if (*c >= target) { __threadfence(); return true; }
    if (err && (spin & 63) == 63 && *err) break;
    __nanosleep(200);
  }
  // the producer never published (it was not co-resident, or died): a silent return would leave this tile unwritten
  // and the call would still report success, so the kernel aborts -- every later call on the context fails loudly
  __trap();
  return false;
#endif
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
#if defined(MRIACL_EMU)
  reinterpret_cast<float4*>(smem_dst)[0] = reinterpret_cast<const float4*>(gsrc)[0];
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cp_async16_hint(void* smem_dst, const void* gsrc, unsigned long long pol) {
#if defined(MRIACL_EMU)
  reinterpret_cast<float4*>(smem_dst)[0] = reinterpret_cast<const float4*>(gsrc)[0];
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "l"(pol) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#if !defined(MRIACL_EMU)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N_PENDING> __device__ __forceinline__ void cp_async_wait() {
#if !defined(MRIACL_EMU)
  asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory");
#endif
}

template <int NW> __device__ __forceinline__ float rp_block_sum(float v, float* red /* NW floats */) {
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

#ifdef MRIACL_EXPERIMENTAL   // the 32-row kernel: superseded by the 16-row kernel (rowpass16.cuh), kept as a measured alternative
// sparse residue with exactly NNZ sampled columns: Y'[k1] = sum_e x_e w_N^{n_e k1}; the twiddle
// rows were tabulated by the host plan, so every operand is a shared-memory load at an immediate offset
template <int P, int Q, int NNZ>
__device__ __forceinline__ void rp_sparse_unit(const int* sch, const cf* tb, const cf* sptw, cf* ycol) {
  constexpr int PITCH = 24;
  static_assert(P <= PITCH && PITCH % 2 == 0, "sptw row pitch");
  cf xe[NNZ > 0 ? NNZ : 1];
  const float4* tw4 = reinterpret_cast<const float4*>(sptw + sch[0]);
#pragma unroll
  for (int e = 0; e < NNZ; ++e) xe[e] = tb[sch[1 + e] * RP_ROWS];
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
    ycol[(2 * kp) * Q * RP_ROWS] = y0;
    if (2 * kp + 1 < P) ycol[(2 * kp + 1) * Q * RP_ROWS] = y1;
  }
}

template <int P, int Q, int NW>
__global__ void __launch_bounds__(NW * 32, NW <= 8 ? 2 : 1) rowpass_kernel(RowPassParams p) {
  static_assert(Q == 16, "stage 2 is the register-level 16-point FFT");
  constexpr int N = P * Q;
  constexpr int NT = NW * 32;
  constexpr int KPW = (P + NW - 1) / NW;
  constexpr int HP = (P - 1) / 2;          // pairs of the symmetric DFT
  constexpr int HSPLIT = (HP + 2) / 2;     // half A: X0 and pairs 1..HSPLIT-1, half B: pairs HSPLIT..HP
  MRIACL_DYN_SMEM(cf, Y);                       // [P][Q][32]
  cf* twsm = Y + N * RP_ROWS;                   // [N]
  cf* sptwsm = twsm + N;                        // [sptw_len]
  int* schsm = reinterpret_cast<int*>(reinterpret_cast<char*>(sptwsm) + rp_round16(p.sptw_len * 8));
  cf* tbuf = reinterpret_cast<cf*>(reinterpret_cast<char*>(schsm) + rp_round16(p.sched_len * 4));   // [n_buf][n_act][32]
  float* avsm = reinterpret_cast<float*>(tbuf + (size_t)p.n_buf * p.n_act * RP_ROWS);   // (A > 1)
  float* osm = reinterpret_cast<float*>(Y);     // output tile [32][ow+1], aliases Y
  __shared__ float red[NW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int opitch = p.ow + 1;
  for (int i = tid; i < N; i += NT) twsm[i] = p.tw[i];
  for (int i = tid; i < p.sched_len; i += NT) schsm[i] = p.sched[i];
  for (int i = tid; i < p.sptw_len; i += NT) sptwsm[i] = p.sptw[i];
  __syncthreads();

  const int my_off = schsm[warp];
  const int n_items = p.n_slices * p.n_tiles;
  const int n_frames = p.A * p.C;
  const long long frame_elems = (long long)p.n_act * p.ohp;
  const int tile_elems = p.n_act * RP_ROWS;           // complex elements of one prefetched block
  const int n_copies = p.n_act * (RP_ROWS / 2);       // 16-byte copies per block

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int s = item / p.n_tiles, tile = item - s * p.n_tiles;
    const cf* Tit = p.T + (long long)s * n_frames * frame_elems + tile * RP_ROWS;

    // thread i copies the 16-byte piece (j, part) = (i / 16, i % 16) of the block, then j += NT / 16
    auto prefetch = [&](int f, int buf) {
      const cf* src = Tit + (long long)f * frame_elems + (long long)(tid >> 4) * p.ohp + 2 * (tid & 15);
      cf* dst = tbuf + (size_t)buf * tile_elems + (tid >> 4) * RP_ROWS + 2 * (tid & 15);
      for (int i = tid; i < n_copies; i += NT) {
        cp_async16(dst, src);
        src += (long long)(NT / 16) * p.ohp;
        dst += (NT / 16) * RP_ROWS;
      }
      cp_async_commit();
    };

    if (p.done) {     // overlapped with the column pass: wait until every column group of slice s has landed
      __shared__ int ready;
      if (tid == 0) {
        ready = rp_wait_count(p.done + s, p.done_target, p.error_flag) ? 1 : 0;
        if (!ready && p.error_flag) atomicAdd(p.error_flag, 1);
      }
      __syncthreads();
      if (!ready) return;
    }
    if (p.A > 1) for (int i = tid; i < RP_ROWS * opitch; i += NT) avsm[i] = 0.f;
    prefetch(0, 0);

    float acc[KPW][Q];
    for (int f = 0; f < n_frames; ++f) {
      const int buf = p.n_buf == 2 ? (f & 1) : 0;
      if (f % p.C == 0) {
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk)
#pragma unroll
          for (int k2 = 0; k2 < Q; ++k2) acc[kk][k2] = 0.f;
      }
      if (p.n_buf == 2) {
        if (f + 1 < n_frames) { prefetch(f + 1, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();   // block f visible; stage 2 of frame f-1 (and the previous item's tile reads) done

      // ---------------- stage 1: pruned P-point DFTs of my pieces ----------------
      {
        const cf* tb = tbuf + (size_t)buf * tile_elems + lane;
        const int n_units = schsm[my_off];
        int off = my_off + 1;
        for (int u = 0; u < n_units; ++u) {
          const int n2 = schsm[off], type = schsm[off + 1], nnz = schsm[off + 2];
          off += 3;
          cf* ycol = Y + n2 * RP_ROWS + lane;          // + k1 * Q * 32
          if (type != 0) {
            cf x[P];
#pragma unroll
            for (int n1 = 0; n1 < P; ++n1) {
              const int j = schsm[off + n1];
              x[n1] = j >= 0 ? tb[j * RP_ROWS] : cf_make(0.f, 0.f);
            }
            off += P;
            auto emit = [&](auto kc, cf val) {
              constexpr int k1 = decltype(kc)::value;
              if (k1 != 0) val = cmul(val, twsm[(n2 * k1) % N]);
              ycol[k1 * Q * RP_ROWS] = val;
            };
            if (type == 1) dft_odd_sym_part<P, true, 1, HP + 1, true>(x, emit);
            else if (type == 2) dft_odd_sym_part<P, true, 1, HSPLIT, true>(x, emit);
            else dft_odd_sym_part<P, true, HSPLIT, HP + 1, false>(x, emit);
          } else {
            const int* sch = schsm + off;
            switch (nnz) {
              case 0: rp_sparse_unit<P, Q, 0>(sch, tb, sptwsm, ycol); break;
              case 1: rp_sparse_unit<P, Q, 1>(sch, tb, sptwsm, ycol); break;
              case 2: rp_sparse_unit<P, Q, 2>(sch, tb, sptwsm, ycol); break;
              case 3: rp_sparse_unit<P, Q, 3>(sch, tb, sptwsm, ycol); break;
              case 4: rp_sparse_unit<P, Q, 4>(sch, tb, sptwsm, ycol); break;
              case 5: rp_sparse_unit<P, Q, 5>(sch, tb, sptwsm, ycol); break;
              default: rp_sparse_unit<P, Q, 6>(sch, tb, sptwsm, ycol); break;
            }
            off += 1 + nnz;
          }
        }
      }
      __syncthreads();
      if (p.n_buf == 1 && f + 1 < n_frames) prefetch(f + 1, 0);   // single buffer: overlap with stage 2 only

      // ---------------- stage 2: Q-point FFT over n2, accumulate |X|^2 ----------------
#pragma unroll
      for (int kk = 0; kk < KPW; ++kk) {
        const int k1 = warp + NW * kk;
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

      // ---------------- end of an average: sqrt, shift + crop into the running tile ----------------
      if (p.A > 1 && (f + 1) % p.C == 0) {
#pragma unroll
        for (int kk = 0; kk < KPW; ++kk) {
          const int k1 = warp + NW * kk;
          if (k1 < P) {
#pragma unroll
            for (int k2 = 0; k2 < Q; ++k2) {
              const int cc = phys_of_logical(k1 + P * k2, N) - p.col0;
              if (cc >= 0 && cc < p.ow) avsm[lane * opitch + cc] += sqrtf(acc[kk][k2]) * p.scale;   // private owner
            }
          }
        }
      }
    }
    __syncthreads();     // all stage-2 reads of Y done: the output tile may alias it
    if (p.A == 1) {
#pragma unroll
      for (int kk = 0; kk < KPW; ++kk) {
        const int k1 = warp + NW * kk;
        if (k1 < P) {
#pragma unroll
          for (int k2 = 0; k2 < Q; ++k2) {
            const int cc = phys_of_logical(k1 + P * k2, N) - p.col0;
            if (cc >= 0 && cc < p.ow) osm[lane * opitch + cc] = sqrtf(acc[kk][k2]) * p.scale;
          }
        }
      }
      __syncthreads();
    }

    // ---------------- write the tile (coalesced) and its statistics ----------------
    const float* tile_sm = p.A > 1 ? avsm : osm;
    const float inv_a = 1.0f / (float)p.A;
    const int rows_here = min(RP_ROWS, p.oh - tile * RP_ROWS);
    const int n_here = rows_here * p.ow;
    float* dst = p.out + ((long long)s * p.oh + tile * RP_ROWS) * p.ow;
    float lsum = 0.f;
    for (int e = tid; e < n_here; e += NT) {
      const int r = e / p.ow, cc = e - r * p.ow;
      float v = tile_sm[r * opitch + cc];
      if (p.A > 1) v *= inv_a;
      dst[e] = v;
      lsum += v;
    }
    if (p.partials) {
      const float mean = rp_block_sum<NW>(lsum, red) / (float)n_here;
      float lq = 0.f;
      for (int e = tid; e < n_here; e += NT) {
        const int r = e / p.ow, cc = e - r * p.ow;
        float v = tile_sm[r * opitch + cc];
        if (p.A > 1) v *= inv_a;
        const float d = v - mean;
        lq = fmaf(d, d, lq);
      }
      const float m2 = rp_block_sum<NW>(lq, red);
      if (tid == 0) {
        float* q = p.partials + ((long long)s * p.n_tiles + tile) * 3;
        q[0] = (float)n_here; q[1] = mean; q[2] = m2;
      }
    }
    __syncthreads();   // tile fully consumed before the next item's prefetch / stage 1 reuse the buffers
  }
}

#endif  // MRIACL_EXPERIMENTAL

}  // namespace mriacl
