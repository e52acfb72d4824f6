// cuda_emu.h -- minimal CPU emulation of the CUDA execution model, TEST INFRASTRUCTURE ONLY.
//
// Lets the kernel sources under mri_acl_imagesegmentation_adsp_b200/csrc/*.cuh be compiled
// with g++ and executed one OS thread per CUDA thread (blocks run one after another), so the
// index maps, plans and barriers of the kernels can be checked against numpy in the build
// container, which has no GPU.  Nothing here is linked into libmriacl_recon.so and nothing in
// the product path can reach it.
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <algorithm>

#define MRIACL_EMU 1

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static
#define __align__(n) alignas(n)

namespace emu {

class Barrier {
 public:
  explicit Barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m_);
    int gen = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen != gen_; });
  }
 private:
  std::mutex m_; std::condition_variable cv_; int n_, count_ = 0, gen_ = 0;
};

// CUDA named barrier (bar.sync / bar.arrive with an explicit thread count)
class NamedBarrier {
 public:
  void arrive(int n, bool wait) {
    std::unique_lock<std::mutex> lk(m_);
    int gen = gen_;
    if (++count_ >= n) { count_ = 0; ++gen_; cv_.notify_all(); }
    else if (wait) cv_.wait(lk, [&] { return gen != gen_; });
  }
 private:
  std::mutex m_; std::condition_variable cv_; int count_ = 0, gen_ = 0;
};

struct BlockCtx {
  NamedBarrier named[16];
  Barrier* block_bar;
  std::vector<Barrier*> warp_bar;
  std::vector<uint64_t> shfl;   // one 8-byte slot per thread
  unsigned char* dyn_smem;
};

inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockCtx* t_ctx = nullptr;
inline thread_local int t_linear_tid = 0;

// run `body` once per CUDA thread of a 1-D grid of 1-D blocks
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  const int nt = (int)block.x;
  const int nwarps = (nt + 31) / 32;
  std::vector<unsigned char> smem(smem_bytes + 64);
  for (unsigned by = 0; by < grid.y; ++by)
  for (unsigned bx = 0; bx < grid.x; ++bx) {
    BlockCtx ctx;
    Barrier bb(nt);
    ctx.block_bar = &bb;
    std::vector<std::unique_ptr<Barrier>> wb;
    for (int w = 0; w < nwarps; ++w) {
      int lanes = std::min(32, nt - 32 * w);
      wb.emplace_back(new Barrier(lanes));
      ctx.warp_bar.push_back(wb.back().get());
    }
    ctx.shfl.assign(nt, 0);
    std::fill(smem.begin(), smem.end(), (unsigned char)0xCD);   // poison: uninitialised reads show up
    ctx.dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    std::vector<std::thread> ths;
    ths.reserve(nt);
    for (int t = 0; t < nt; ++t) {
      ths.emplace_back([&, t] {
        t_threadIdx = dim3(t); t_blockIdx = dim3(bx, by); t_blockDim = block; t_gridDim = grid;
        t_ctx = &ctx; t_linear_tid = t;
        body();
      });
    }
    for (auto& th : ths) th.join();
  }
}

template <class T> inline T shfl_generic(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl payload");
  BlockCtx* c = t_ctx;
  const int warp = t_linear_tid / 32;
  uint64_t raw = 0; std::memcpy(&raw, &v, sizeof(T));
  c->shfl[t_linear_tid] = raw;
  c->warp_bar[warp]->wait();
  uint64_t got = c->shfl[warp * 32 + (src_lane & 31)];
  c->warp_bar[warp]->wait();
  T out; std::memcpy(&out, &got, sizeof(T));
  return out;
}

}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

static inline void __syncthreads() { emu::t_ctx->block_bar->wait(); }
static inline void mriacl_emu_bar_sync(int id, int nthreads) { emu::t_ctx->named[id].arrive(nthreads, true); }
static inline void mriacl_emu_bar_arrive(int id, int nthreads) { emu::t_ctx->named[id].arrive(nthreads, false); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_ctx->warp_bar[emu::t_linear_tid / 32]->wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  return emu::shfl_generic(v, (emu::t_linear_tid & 31) ^ lane_mask);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_generic(v, src); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  int lane = emu::t_linear_tid & 31;
  return emu::shfl_generic(v, lane + d < 32 ? lane + d : lane);
}
template <class T> static inline T __ldg(const T* p) { return *p; }

static inline float atomicAdd(float* p, float v) {
  auto* a = reinterpret_cast<std::atomic<float>*>(p);
  float old = a->load();
  while (!a->compare_exchange_weak(old, old + v)) {}
  return old;
}
static inline double atomicAdd(double* p, double v) {
  auto* a = reinterpret_cast<std::atomic<double>*>(p);
  double old = a->load();
  while (!a->compare_exchange_weak(old, old + v)) {}
  return old;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return reinterpret_cast<std::atomic<unsigned>*>(p)->fetch_add(v); }
static inline int atomicAdd(int* p, int v) { return reinterpret_cast<std::atomic<int>*>(p)->fetch_add(v); }
static inline int atomicCAS(int* p, int cmp, int val) { reinterpret_cast<std::atomic<int>*>(p)->compare_exchange_strong(cmp, val); return cmp; }

static inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
using std::min; using std::max;

// dynamic shared memory
#define MRIACL_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::t_ctx->dyn_smem)
