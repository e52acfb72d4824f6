// rt.h -- the few runtime calls the host orchestration needs, behind one seam.
//
// Product build (nvcc): CUDA runtime.  Emulation build (g++ with tests/emu/cuda_emu.h, used only
// by tests/emu to check the orchestration and kernels on a GPU-less machine): host memory and
// the thread-per-CUDA-thread emulator.  The emulation library is never looked for, loaded or
// linked by the product package.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace mriacl {

inline std::atomic<uint64_t>& launch_counter() { static std::atomic<uint64_t> c{0}; return c; }

#ifdef MRIACL_EMU
typedef void* rt_stream_t;
inline int rt_malloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 1; }
inline int rt_free(void* p) { std::free(p); return 0; }
inline int rt_upload(void* dst, const void* src, size_t n) { std::memcpy(dst, src, n); return 0; }
inline int rt_device() { return 0; }
inline int rt_sm_count(int) { return 2; }
inline const char* rt_last_error_string() { return "emulation"; }
inline int rt_check() { return 0; }
inline int rt_allow_smem(const void*, int, int = -1) { return 0; }
typedef int rt_event_t;
inline int rt_stream_create_high_priority(rt_stream_t* s) { *s = nullptr; return 0; }
inline int rt_event_create(rt_event_t* e) { *e = 0; return 0; }
inline int rt_event_record(rt_event_t, rt_stream_t) { return 0; }
inline int rt_stream_wait_event(rt_stream_t, rt_event_t) { return 0; }
inline int rt_memset_async(void* p, int v, size_t n, rt_stream_t) { std::memset(p, v, n); return 0; }
#define MRIACL_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
  do { emu::launch(dim3((unsigned)(grid)), dim3((unsigned)(block)), (size_t)(smem),           \
                   [&] { kern(__VA_ARGS__); });                                               \
       ::mriacl::launch_counter()++; } while (0)
#else
typedef cudaStream_t rt_stream_t;
inline int rt_malloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess ? 0 : 1; }
inline int rt_free(void* p) { return cudaFree(p) == cudaSuccess ? 0 : 1; }
inline int rt_upload(void* dst, const void* src, size_t n) {
  return cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 1;   // plan build only, once per plan
}
inline int rt_device() { int d = -1; return cudaGetDevice(&d) == cudaSuccess ? d : -1; }
inline int rt_sm_count(int dev) {
  int n = 0;
  return cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess ? n : 0;
}
inline const char* rt_last_error_string() { return cudaGetErrorString(cudaGetLastError()); }
inline int rt_check() { return cudaPeekAtLastError() == cudaSuccess ? 0 : 1; }
// opt in to large dynamic shared memory; carveout_pct >= 0 additionally sets the preferred shared-memory
// carveout (percent of the SM's 228 KB; a carveout change needs an idle SM, so kernels meant to be
// co-resident must agree on it)
inline int rt_allow_smem(const void* fn, int bytes, int carveout_pct = -1) {
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return 1;
  if (carveout_pct < 0) return 0;
  return cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct) == cudaSuccess ? 0 : 1;
}
typedef cudaEvent_t rt_event_t;
inline int rt_stream_create_high_priority(rt_stream_t* s) {
  int lo = 0, hi = 0;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return 1;
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, hi) == cudaSuccess ? 0 : 1;
}
inline int rt_event_create(rt_event_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess ? 0 : 1; }
inline int rt_event_record(rt_event_t e, rt_stream_t s) { return cudaEventRecord(e, s) == cudaSuccess ? 0 : 1; }
inline int rt_stream_wait_event(rt_stream_t s, rt_event_t e) { return cudaStreamWaitEvent(s, e, 0) == cudaSuccess ? 0 : 1; }
inline int rt_memset_async(void* p, int v, size_t n, rt_stream_t s) { return cudaMemsetAsync(p, v, n, s) == cudaSuccess ? 0 : 1; }
#define MRIACL_LAUNCH(kern, grid, block, smem, stream, ...)                                   \
  do { kern<<<(unsigned)(grid), (unsigned)(block), (size_t)(smem), (stream)>>>(__VA_ARGS__);  \
       ::mriacl::launch_counter()++; } while (0)
#endif

}  // namespace mriacl
