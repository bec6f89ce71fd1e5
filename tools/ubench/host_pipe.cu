// host_pipe.cu -- second r02 experiment: does a SMALL pinned ring (pieces of 2-16 MB that stay in the
// CPU's last-level cache, written with ordinary stores) beat a large ring written with non-temporal
// stores?  The DMA engine then reads the pieces from cache and DRAM only sees the source stream.
// Uses the library's own copy pool (csrc/copy_pool.h).
// build: nvcc -O3 -arch=sm_100a -o host_pipe host_pipe.cu -lpthread
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include "../../som_lvq_pak_b200/csrc/copy_pool.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main() {
  const size_t total = (size_t)2560 << 20, maxring = (size_t)256 << 20;
  char *src = (char *)malloc(total);
  memset(src, 1, total);
  char *pin = nullptr, *dev = nullptr;
  cudaHostAlloc((void **)&pin, maxring, cudaHostAllocDefault);
  cudaMalloc((void **)&dev, total);
  cudaStream_t st;
  cudaStreamCreate(&st);
  for (int T : {4, 8, 12}) {
    bmu::CopyPool pool(T);
    pool.begin();
    for (int mode : {0, 1})
      for (size_t piece : {(size_t)1 << 20, (size_t)2 << 20, (size_t)4 << 20, (size_t)8 << 20, (size_t)16 << 20, (size_t)64 << 20})
        for (int nslot : {3, 6}) {
          if (piece * nslot > maxring) continue;
          cudaEvent_t ev[8];
          for (int i = 0; i < nslot; i++) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
          double best = 0;
          for (int rep = 0; rep < 2; rep++) {
            double t0 = now();
            size_t ci = 0;
            for (size_t off = 0; off < total; off += piece, ci++) {
              const int b = ci % nslot;
              if (ci >= (size_t)nslot) cudaEventSynchronize(ev[b]);
              pool.copy(pin + b * piece, src + off, piece, mode);
              cudaMemcpyAsync(dev + off, pin + b * piece, piece, cudaMemcpyHostToDevice, st);
              cudaEventRecord(ev[b], st);
            }
            cudaStreamSynchronize(st);
            double gbs = total / (now() - t0) / 1e9;
            if (gbs > best) best = gbs;
          }
          printf("T=%2d %-6s piece %3zu MB x %d slots: %.1f GB/s\n", T, mode ? "stream" : "cached", piece >> 20, nslot, best);
          for (int i = 0; i < nslot; i++) cudaEventDestroy(ev[i]);
        }
    pool.end();
  }
  return 0;
}
