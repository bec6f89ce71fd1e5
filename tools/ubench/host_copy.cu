// host_copy.cu -- what feeds a B200 from PAGEABLE host memory fastest?  (r02 experiment behind the
// design of csrc/bmu_host.cu; numbers in profiles/r02_host_copy_ubench.txt)
//   (a) T threads of memcpy / non-temporal AVX-512 copies: pageable -> pinned ring
//   (b) cudaHostRegister of 64 MB chunks in place (pin + unpin rate), 1..4 threads
//   (c) H2D DMA rate from pinned memory, alone and while (a) runs beside it
// build: nvcc -O3 -arch=sm_100a -Xcompiler -mavx512f,-mavx2 -o host_copy host_copy.cu -lpthread
#include <cuda_runtime.h>
#include <immintrin.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void copy_nt512(char *dst, const char *src, size_t n) {
  size_t i = 0;
  for (; i + 256 <= n; i += 256) {
    __m512i a = _mm512_loadu_si512((const void *)(src + i)), b = _mm512_loadu_si512((const void *)(src + i + 64));
    __m512i c = _mm512_loadu_si512((const void *)(src + i + 128)), d = _mm512_loadu_si512((const void *)(src + i + 192));
    _mm512_stream_si512((__m512i *)(dst + i), a);
    _mm512_stream_si512((__m512i *)(dst + i + 64), b);
    _mm512_stream_si512((__m512i *)(dst + i + 128), c);
    _mm512_stream_si512((__m512i *)(dst + i + 192), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}
static void copy_nt256(char *dst, const char *src, size_t n) {
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(src + i)), b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
    __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64)), d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
    _mm256_stream_si256((__m256i *)(dst + i), a);
    _mm256_stream_si256((__m256i *)(dst + i + 32), b);
    _mm256_stream_si256((__m256i *)(dst + i + 64), c);
    _mm256_stream_si256((__m256i *)(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n) memcpy(dst + i, src + i, n - i);
}

typedef void (*copy_fn)(char *, const char *, size_t);
static void copy_libc(char *d, const char *s, size_t n) { memcpy(d, s, n); }

// T threads copy `total` bytes from src (pageable) into a ring of pinned chunks
static double par_copy(copy_fn f, int T, char *dst, size_t ring, const char *src, size_t total, size_t chunk) {
  double t0 = now();
  for (size_t off = 0; off < total; off += chunk) {
    size_t n = total - off < chunk ? total - off : chunk;
    char *d = dst + (off % ring);
    std::vector<std::thread> th;
    size_t per = ((n + T - 1) / T + 4095) & ~(size_t)4095;
    for (int t = 0; t < T; t++) {
      size_t lo = per * t, hi = lo + per < n ? lo + per : n;
      if (lo < hi) th.emplace_back([=] { f(d + lo, src + off + lo, hi - lo); });
    }
    for (auto &x : th) x.join();
  }
  return now() - t0;
}

int main(int argc, char **argv) {
  size_t total = (size_t)2560 << 20, chunk = (size_t)64 << 20, ring = 3 * chunk;
  char *src = (char *)malloc(total);
  for (size_t i = 0; i < total; i += 4096) src[i] = (char)i;       // touch: real pages
  memset(src, 1, total);
  char *pin = nullptr, *dev = nullptr;
  cudaHostAlloc((void **)&pin, ring, cudaHostAllocDefault);
  cudaMalloc((void **)&dev, ring);
  cudaStream_t st;
  cudaStreamCreate(&st);
  printf("cores online: %u\n", std::thread::hardware_concurrency());
  struct { const char *name; copy_fn f; } fns[] = {{"memcpy", copy_libc}, {"nt256", copy_nt256}, {"nt512", copy_nt512}};
  for (auto &fn : fns)
    for (int T : {1, 2, 4, 6, 8, 12, 16}) {
      par_copy(fn.f, T, pin, ring, src, total / 4, chunk);
      double s = par_copy(fn.f, T, pin, ring, src, total, chunk);
      printf("(a) %-6s T=%2d  %.1f GB/s\n", fn.name, T, total / s / 1e9);
    }
  // (c) DMA alone
  {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (size_t off = 0; off < total; off += chunk) cudaMemcpyAsync(dev + off % ring, pin + off % ring, chunk, cudaMemcpyHostToDevice, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("(c) H2D from pinned alone: %.1f GB/s\n", total / (ms * 1e-3) / 1e9);
  }
  // (c2) DMA while T threads copy (pipelined like bmu_host.cu: copy chunk c+1 while chunk c is in flight)
  for (auto &fn : fns)
    for (int T : {4, 8, 12, 16}) {
      cudaEvent_t ev[3];
      for (int i = 0; i < 3; i++) cudaEventCreate(&ev[i]);
      double t0 = now();
      size_t ci = 0;
      for (size_t off = 0; off < total; off += chunk, ci++) {
        int b = ci % 3;
        if (ci >= 3) cudaEventSynchronize(ev[b]);
        par_copy(fn.f, T, pin + b * chunk, chunk, src + off, chunk, chunk);
        cudaMemcpyAsync(dev + b * chunk, pin + b * chunk, chunk, cudaMemcpyHostToDevice, st);
        cudaEventRecord(ev[b], st);
      }
      cudaStreamSynchronize(st);
      double s = now() - t0;
      printf("(c2) staged pipeline %-6s T=%2d: %.1f GB/s end to end\n", fn.name, T, total / s / 1e9);
    }
  // (b) register in place
  for (int T : {1, 2, 4}) {
    double t0 = now();
    std::vector<std::thread> th;
    std::atomic<size_t> next{0};
    for (int t = 0; t < T; t++)
      th.emplace_back([&] {
        for (;;) {
          size_t off = next.fetch_add(chunk);
          if (off >= total) break;
          cudaHostRegister(src + off, chunk, cudaHostRegisterDefault);
        }
      });
    for (auto &x : th) x.join();
    double s = now() - t0;
    double t1 = now();
    for (size_t off = 0; off < total; off += chunk) cudaHostUnregister(src + off);
    double u = now() - t1;
    printf("(b) cudaHostRegister 64MB chunks T=%d: %.1f GB/s pin, %.1f GB/s unpin\n", T, total / s / 1e9, total / u / 1e9);
  }
  // (b2) pipeline: register chunk c+1 (helper thread) while chunk c is DMA'd from the caller's memory
  {
    double t0 = now();
    std::thread helper;
    cudaHostRegister(src, chunk, cudaHostRegisterDefault);
    size_t ci = 0;
    for (size_t off = 0; off < total; off += chunk, ci++) {
      if (off + chunk < total) helper = std::thread([=] { cudaHostRegister(src + off + chunk, chunk, cudaHostRegisterDefault); });
      cudaMemcpyAsync(dev + (ci % 3) * chunk, src + off, chunk, cudaMemcpyHostToDevice, st);
      if (helper.joinable()) helper.join();
    }
    cudaStreamSynchronize(st);
    double s = now() - t0;
    for (size_t off = 0; off < total; off += chunk) cudaHostUnregister(src + off);
    printf("(b2) register-ahead pipeline: %.1f GB/s end to end (+ unregister %.3f s)\n", total / s / 1e9, now() - t0 - s);
  }
  // (d) plain cudaMemcpy from pageable (the driver's own staging)
  {
    double t0 = now();
    for (size_t off = 0; off < total; off += chunk) cudaMemcpy(dev + off % ring, src + off, chunk, cudaMemcpyHostToDevice);
    printf("(d) cudaMemcpy from pageable: %.1f GB/s\n", total / (now() - t0) / 1e9);
  }
  return 0;
}
