// hostpack.h -- host side of the end-to-end path: gather the SAMPLED phase-encode columns of host k-space into a
// dense staging buffer before the host -> device copy.
//
// The undersampling mask multiplies unsampled columns by exactly zero, so they never have to cross PCIe: at 4x
// (114 of 368 columns) the copy shrinks to 31 % of the k-space bytes.  The gather itself is memory-bound on the host
// (every 64-byte cache line of a row holds sampled elements), so it runs on a small persistent pool of threads, each
// taking contiguous blocks of rows; stores bypass the cache (the staging buffer is only read by the DMA engine).
// Plain C++ threads: no CUDA here, and nothing of this is used by the device-resident path.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#if defined(__linux__)
#include <sched.h>
#endif

namespace mriacl {

class HostPool {
 public:
  static HostPool& get() { static HostPool p; return p; }

  // run fn(block) for block in [0, n_blocks) on up to n_threads threads (the caller is one of them)
  void parallel_for(int n_blocks, int n_threads, const std::function<void(int)>& fn) {
    if (n_blocks <= 0) return;
    n_threads = std::max(1, std::min(n_threads, n_blocks));
    std::lock_guard<std::mutex> call_lock(call_mu_);           // one parallel region at a time
    grow(n_threads - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn; n_blocks_ = n_blocks; next_.store(0); active_ = n_threads - 1; wanted_ = n_threads - 1; ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return active_ == 0; });
    fn_ = nullptr;
  }

  static int default_threads() {
    int n = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
#endif
    return std::max(1, std::min(n, 64));
  }

 private:
  HostPool() = default;
  ~HostPool() {
    { std::lock_guard<std::mutex> lk(mu_); quit_ = true; ++epoch_; }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void grow(int n) {
    while ((int)threads_.size() < n) {
      const int id = (int)threads_.size();
      threads_.emplace_back([this, id] { loop(id); });
    }
  }
  void work() {
    for (;;) {
      const int b = next_.fetch_add(1);
      if (b >= n_blocks_) break;
      (*fn_)(b);
    }
  }
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return quit_ || (epoch_ != seen && id < wanted_); });
        if (quit_) return;
        seen = epoch_;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--active_ == 0) done_cv_.notify_one();
      }
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_blocks_ = 0, active_ = 0, wanted_ = 0;
  uint64_t epoch_ = 0;
  bool quit_ = false;
};

// dst[r][j] = src[r][idx[j]] for r < n_rows, j < n_idx; 8-byte elements (complex64).  stream: stores bypass the cache
inline void pack_columns_rows(const uint64_t* src, uint64_t* dst, long long r0, long long r1, int W,
                              const int* idx, int n_idx, bool stream) {
#if defined(__x86_64__)
  if (stream) {
    for (long long r = r0; r < r1; ++r) {
      const uint64_t* s = src + r * (long long)W;
      uint64_t* d = dst + r * (long long)n_idx;
      for (int j = 0; j < n_idx; ++j) _mm_stream_si64(reinterpret_cast<long long*>(d + j), (long long)s[idx[j]]);
    }
    _mm_sfence();
    return;
  }
#endif
  for (long long r = r0; r < r1; ++r) {
    const uint64_t* s = src + r * (long long)W;
    uint64_t* d = dst + r * (long long)n_idx;
    for (int j = 0; j < n_idx; ++j) d[j] = s[idx[j]];
  }
}

inline void pack_columns(const void* src, void* dst, long long n_rows, int W, const std::vector<int>& idx, int n_threads) {
  const int n_idx = (int)idx.size();
  if (n_rows <= 0 || n_idx == 0) return;
  if (n_threads <= 0) n_threads = HostPool::default_threads();
  const long long rows_per_block = std::max<long long>(64, (512 * 1024) / ((long long)W * 8));   // ~0.5 MB of source per block
  const int n_blocks = (int)((n_rows + rows_per_block - 1) / rows_per_block);
  const uint64_t* s = static_cast<const uint64_t*>(src);
  uint64_t* d = static_cast<uint64_t*>(dst);
  const int* ix = idx.data();
  static const bool stream = [] { const char* e = std::getenv("MRIACL_PACK_STREAM"); return e ? std::atoi(e) != 0 : true; }();   // non-temporal stores: no read-for-ownership of the staging lines
  HostPool::get().parallel_for(n_blocks, n_threads, [=](int b) {
    const long long r0 = b * rows_per_block, r1 = std::min(n_rows, r0 + rows_per_block);
    pack_columns_rows(s, d, r0, r1, W, ix, n_idx, stream);
  });
}

}  // namespace mriacl
