// copy_pool.h -- a few host threads that split one host-to-host copy into slices.
//
// Used by the host-pointer pipeline (bmu_host.cu) to stage PAGEABLE caller memory into a pinned ring at
// PCIe rate: one thread moves 10-14 GB/s, a B200's PCIe Gen5 x16 link takes 55 GB/s (measured with
// tools/ubench/host_copy.cu, profiles/r02_host_copy_ubench.txt).  Between begin() and end() the workers
// spin on an atomic generation counter, so that handing them a 4-8 MB piece costs about a microsecond
// instead of a condition-variable round trip; outside they sleep.
#pragma once
#include <immintrin.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace bmu {

enum CopyMode { COPY_CACHED = 0, COPY_STREAM = 1 };

// non-temporal (write-combining) copy: the destination is read next by a DMA engine, not by this core
__attribute__((target("avx512f"))) static inline void copy_stream512(char *dst, const char *src, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = src[i]; i++; }
  for (; i + 256 <= n; i += 256) {
    const __m512i a = _mm512_loadu_si512((const void *)(src + i)), b = _mm512_loadu_si512((const void *)(src + i + 64));
    const __m512i c = _mm512_loadu_si512((const void *)(src + i + 128)), d = _mm512_loadu_si512((const void *)(src + i + 192));
    _mm512_stream_si512((__m512i *)(dst + i), a);
    _mm512_stream_si512((__m512i *)(dst + i + 64), b);
    _mm512_stream_si512((__m512i *)(dst + i + 128), c);
    _mm512_stream_si512((__m512i *)(dst + i + 192), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}
__attribute__((target("avx2"))) static inline void copy_stream256(char *dst, const char *src, size_t n) {
  size_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; i++; }
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a);
    _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c);
    _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}
static inline void copy_bytes(char *dst, const char *src, size_t n, int mode) {
  static const int isa = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
  if (mode == COPY_STREAM && isa == 2) copy_stream512(dst, src, n);
  else if (mode == COPY_STREAM && isa == 1) copy_stream256(dst, src, n);
  else memcpy(dst, src, n);
}

class CopyPool {
 public:
  explicit CopyPool(int nthreads) : n_(nthreads < 1 ? 1 : nthreads) {
    for (int i = 1; i < n_; i++) workers_.emplace_back([this, i] { loop(i); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      active_.store(true);
    }
    cv_.notify_all();
    gen_.fetch_add(1);
    for (auto &t : workers_) t.join();
  }
  int threads() const { return n_; }
  // workers spin between begin() and end()
  void begin() {
    {
      std::lock_guard<std::mutex> lk(m_);
      active_.store(true);
    }
    cv_.notify_all();
  }
  void end() { active_.store(false); }
  // blocking: returns when every byte has been copied
  void copy(void *dst, const void *src, size_t bytes, int mode) {
    if (bytes == 0) return;
    const size_t kMinSlice = 256u << 10;
    int parts = (int)((bytes + kMinSlice - 1) / kMinSlice);
    if (parts > n_) parts = n_;
    if (parts <= 1 || !active_.load()) { copy_bytes((char *)dst, (const char *)src, bytes, mode); return; }
    dst_ = (char *)dst; src_ = (const char *)src; bytes_ = bytes; parts_ = parts; mode_ = mode;
    // EVERY worker acknowledges every generation (with or without a slice of it), so the job fields are
    // never rewritten while a late worker is still reading them
    pending_.store(n_ - 1, std::memory_order_relaxed);
    gen_.fetch_add(1, std::memory_order_release);
    slice(0);
    while (pending_.load(std::memory_order_acquire) != 0) _mm_pause();
  }

 private:
  void slice(int i) {
    // slices on 4 KiB boundaries so that two threads never share a page of the destination
    const size_t per = ((bytes_ + parts_ - 1) / parts_ + 4095) & ~(size_t)4095;
    const size_t lo = per * i, hi = lo + per < bytes_ ? lo + per : bytes_;
    if (lo < hi) copy_bytes(dst_ + lo, src_ + lo, hi - lo, mode_);
  }
  void loop(int i) {
    unsigned long seen = 0;            // generations start at 0: a worker that starts late still acknowledges the first job
    for (;;) {
      if (!active_.load(std::memory_order_relaxed)) {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return active_.load(); });
        if (stop_) return;
      }
      const unsigned long g = gen_.load(std::memory_order_acquire);
      if (g == seen) { _mm_pause(); continue; }
      seen = g;
      if (stop_) return;
      if (i < parts_) slice(i);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_;
  std::atomic<unsigned long> gen_{0};
  std::atomic<int> pending_{0};
  std::atomic<bool> active_{false};
  bool stop_ = false;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t bytes_ = 0;
  int parts_ = 0, mode_ = 0;
};

}  // namespace bmu
