// common.cuh -- shared device helpers for libmriacl_recon (sm_100a).
//
// The kernel headers compile in two modes: under nvcc for the product library, and under
// g++ with tests/emu/cuda_emu.h pre-included (MRIACL_EMU) for the CPU emulation tests that
// check index maps and plans in the GPU-less build container.  The emulation is test
// infrastructure; the product library has no CPU path.
#pragma once

#ifndef MRIACL_EMU
#include <cuda_runtime.h>
#define MRIACL_DYN_SMEM(type, name)                                        \
  extern __shared__ __align__(16) unsigned char mriacl_dyn_smem_raw[];     \
  type* name = reinterpret_cast<type*>(mriacl_dyn_smem_raw)
#endif

// profiling switches that skip parts of a kernel (results are garbage) exist in the experimental build only
#ifdef MRIACL_EXPERIMENTAL
#define MRIACL_DBG_SKIP(p, bits) ((p).debug_skip & (bits))
#else
#define MRIACL_DBG_SKIP(p, bits) 0
#endif

namespace mriacl {

typedef float2 cf;  // complex64: x = re, y = im

__device__ __forceinline__ cf cf_make(float re, float im) { return make_float2(re, im); }

// ---- packed fp32x2 arithmetic -------------------------------------------------------------------
// sm_100a has two-wide fp32 instructions (FFMA2 / FADD2 / FMUL2 in SASS, fma/add/mul.rn.f32x2 in PTX)
// that work on an aligned register PAIR -- exactly one complex64.  They run at half the issue rate of the
// scalar forms (same flops per clock, measured: tools/microbench/fp32x2_tput.cu) but need half the issue
// slots, and their operand modifiers make the usual complex idioms free: a scalar register broadcast to
// both halves (x, x), an immediate broadcast, and the half-swap with one negation, i.e. multiplication
// by +-i.  Every butterfly below is written on these three primitives so that the FFT passes, which are
// bound by instruction issue, execute about half as many instructions as with scalar FFMA/FADD.
#if defined(MRIACL_EMU)
__device__ __forceinline__ cf pk_add(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cf pk_mul(cf a, cf b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ cf pk_fma(cf a, cf b, cf c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ cf pk_add(cf a, cf b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ cf pk_mul(cf a, cf b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ cf pk_fma(cf a, cf b, cf c) { return __ffma2_rn(a, b, c); }
#endif
__device__ __forceinline__ cf bc(float s) { return make_float2(s, s); }            // broadcast operand
__device__ __forceinline__ cf cneg(cf a) { return make_float2(-a.x, -a.y); }       // folds into a modifier
__device__ __forceinline__ cf cadd(cf a, cf b) { return pk_add(a, b); }
__device__ __forceinline__ cf csub(cf a, cf b) { return pk_add(a, cneg(b)); }
// multiply by +i (INV) or -i (!INV): a half-swap with one sign flip, folded into the consumer's operand
template <bool INV> __device__ __forceinline__ cf mul_i(cf a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a * b for two variable operands: (a.x, a.x) * b + (a.y, a.y) * (i b)
// (ptxas folds the swap-and-negate into the FIRST multiplicand only, so the rotated operand goes first)
__device__ __forceinline__ cf cmul(cf a, cf b) { return pk_fma(mul_i<true>(b), bc(a.y), pk_mul(b, bc(a.x))); }
// a * conj(b) = a.x (b.x, -b.y) + a.y (b.y, b.x)
__device__ __forceinline__ cf cmulc(cf a, cf b) {
  return pk_fma(make_float2(b.y, b.x), bc(a.y), pk_mul(make_float2(b.x, -b.y), bc(a.x)));
}
// a * (c + i s) for constants c, s (immediates after inlining): c a + s (i a)
__device__ __forceinline__ cf cmul_k(cf a, float c, float s) { return pk_fma(mul_i<true>(a), bc(s), pk_mul(a, bc(c))); }
// acc + a * (c + i s)
__device__ __forceinline__ cf cmac_k(cf a, float c, float s, cf acc) { return pk_fma(mul_i<true>(a), bc(s), pk_fma(a, bc(c), acc)); }
__device__ __forceinline__ cf cscale(cf a, float s) { return pk_mul(a, bc(s)); }
__device__ __forceinline__ float cnorm2(cf a) { return fmaf(a.x, a.x, a.y * a.y); }
// acc + |a|^2
__device__ __forceinline__ float cnorm2_acc(cf a, float acc) { return fmaf(a.x, a.x, fmaf(a.y, a.y, acc)); }
// packed form: acc.x += re^2, acc.y += im^2 (the two halves are summed once, after the coil loop)
__device__ __forceinline__ cf cnorm2_acc2(cf a, cf acc) { return pk_fma(a, a, acc); }

// Streaming 8-byte global load: k-space is read exactly once, keep it out of L1.
__device__ __forceinline__ cf ld_stream(const cf* p) {
#if defined(MRIACL_EMU)
  return *p;
#else
  cf v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#endif
}

// ---- L2 eviction-priority hints --------------------------------------------------------------------------
// k-space is read exactly once: its lines should be the first to leave L2 (evict_first).  The intermediate T is
// written by the column pass and read back by the row pass a few hundred microseconds later, and its buffer is
// reused chunk after chunk: its lines should stay (evict_last), so that T never makes the round trip through HBM.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
#if defined(MRIACL_EMU)
  return 0ull;
#else
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#endif
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
#if defined(MRIACL_EMU)
  return 0ull;
#else
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
#endif
}
__device__ __forceinline__ void st_global_hint(cf* p, cf v, unsigned long long pol) {
#if defined(MRIACL_EMU)
  *p = v;
#else
  asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#endif
}

// named barriers: `count` threads (a multiple of 32) take part; sync waits, arrive does not
__device__ __forceinline__ void named_bar_sync(int id, int count) {
#if defined(MRIACL_EMU)
  mriacl_emu_bar_sync(id, count);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
#if defined(MRIACL_EMU)
  mriacl_emu_bar_arrive(id, count);
#else
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

// ---- "buffer full" signal driven by the completion of cp.async copies ---------------------------------
// Device: an mbarrier in shared memory; every producer lane appends a cp.async.mbarrier.arrive.noinc after
// its copies, so the barrier phase completes when all of them have LANDED -- the producer itself never
// waits and can go on issuing the next gather.  Emulator (copies are synchronous): a named barrier.
struct FullBarrier { unsigned long long word; };

__device__ __forceinline__ void full_init(FullBarrier* b, int producer_lanes) {
#if !defined(MRIACL_EMU)
  const unsigned a = (unsigned)__cvta_generic_to_shared(b);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(producer_lanes) : "memory");
#endif
}
// producer lane: after issuing this item's cp.async copies
__device__ __forceinline__ void full_signal_async(FullBarrier* b, int emu_id, int emu_count) {
#if defined(MRIACL_EMU)
  named_bar_arrive(emu_id, emu_count);
#else
  const unsigned a = (unsigned)__cvta_generic_to_shared(b);
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(a) : "memory");
#endif
}
// consumer thread: wait for phase `parity` (0, 1, 0, ... per use of this barrier)
__device__ __forceinline__ void full_wait(FullBarrier* b, int parity, int emu_id, int emu_count) {
#if defined(MRIACL_EMU)
  named_bar_sync(emu_id, emu_count);
#else
  const unsigned a = (unsigned)__cvta_generic_to_shared(b);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "MRIACL_FULL_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MRIACL_FULL_DONE_%=;\n\t"
      "bra MRIACL_FULL_WAIT_%=;\n\t"
      "MRIACL_FULL_DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
#endif
}

// one thread spins until *counter >= target (another team of this persistent kernel publishes it); bounded so that
// a scheduling surprise cannot hang the device: on a timeout the kernel traps (sticky CUDA error, reported by the call)
__device__ __forceinline__ bool wait_count_ge(const int* counter, int target) {
#if defined(MRIACL_EMU)
  const auto t0 = std::chrono::steady_clock::now();      // emulator: wall-clock bound (CUDA threads are OS threads here)
  while (std::chrono::steady_clock::now() - t0 < std::chrono::seconds(300)) {
    if (reinterpret_cast<const std::atomic<int>*>(counter)->load() >= target) return true;
    std::this_thread::yield();
  }
  return false;
#else
  const volatile int* c = counter;
  for (int spin = 0; spin < (1 << 18); ++spin) {
    if (*c >= target) { __threadfence(); return true; }
    __nanosleep(100);
  }
  __trap();          // never proceed past an unmet dependency (the caller would overwrite data still in use)
  return false;
#endif
}

// physical index of logical (un-shifted) FFT index i, and back, for a centred transform
// of length n:  ifftshift(x)[i] = x[(i + n/2) % n]  and  fftshift(y)[(m + n/2) % n] = y[m]
// (np.fft.ifftshift / fftshift, REF/src/utils/kspace.py:6-8,13-15).
__host__ __device__ __forceinline__ int phys_of_logical(int i, int n) { int p = i + n / 2; return p >= n ? p - n : p; }
__host__ __device__ __forceinline__ int logical_of_phys(int p, int n) { int i = p - n / 2; return i < 0 ? i + n : i; }

}  // namespace mriacl
