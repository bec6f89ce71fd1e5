// Grid exchange microbenchmark (K3's only grid-wide synchronisation): every CTA publishes one tagged 64-bit key and
// learns the minimum of all keys.  Measures cycles from "own key ready" to "global minimum known", averaged over
// iterations and CTAs, for several layouts of the exchange.  One warp per CTA does the exchange, as in k3_train.cu;
// a busy loop of WORK cycles between exchanges stands for the training pass.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_exchange grid_exchange.cu && ./grid_exchange
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ void st_relaxed(u64 *p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed(const u64 *p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_volatile(const u64 *p) {
  u64 v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 warp_min(u64 v) {
  for (int o = 16; o; o >>= 1) {
    u64 w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

struct Params {
  u64 *slots;        // [2][G * stride] (+ second level)
  long long *out;    // per CTA: cycles, polls
  int iters, work, delay_ns, stride, mode, pollers, sk;
};

// mode 0: every CTA polls every slot (k3_train.cu).
// mode 1: only CTAs < pollers poll all slots; the others poll ONE slot (CTA 0's) — shows how the poll round trip
//         depends on the number of pollers (the result is not distributed: timing only).
// mode 2: two levels: CTAs < pollers poll all slots and write the minimum into a result line of their group
//         (G / pollers CTAs per group); the others poll their group's result line.
// mode 3: one CTA (0) polls, writes the result into ONE line that all the others poll.
// sk (how the key is published): 0 st.relaxed.gpu by lane 0 of the polling warp (k3_train.cu), 1 st.relaxed.gpu by a
// SECOND warp, 2 weak st.global by lane 0 of the polling warp, 3 atom.exch (result unused) by lane 0 of the polling warp,
// 4 st.relaxed.gpu + __threadfence, 5 red.max (monotonic value, nothing returned)
__global__ void __launch_bounds__(64) exch(Params p) {
  const int G = gridDim.x, lane = threadIdx.x & 31, me = blockIdx.x, warp = threadIdx.x >> 5;
  if (warp == 1) {
    // publisher warp: same busy loop, then the store; one CTA barrier per exchange keeps it in step with the polling warp
    for (int it = 0; it < p.iters; it++) {
      if (p.sk == 1) {
        long long t = clock64();
        while (clock64() - t < p.work) { }
        const u64 tag = (u64)((it + 1) & 0xFF);
        const u64 key = ((u64)((me * 2654435761u + it * 40503u) & 0xFFFFFF) << 8);
        u64 *slot = p.slots + (size_t)(it & 1) * (G * p.stride + 4096);
        if (lane == 0) st_relaxed(slot + (size_t)me * p.stride, key | tag);
      }
      __syncthreads();
    }
    return;
  }
  long long cyc = 0, polls = 0, pollcyc = 0;
  u64 sink = 0;
  for (int it = 0; it < p.iters; it++) {
    // "pass": a busy loop
    long long t = clock64();
    while (clock64() - t < p.work) { }
    const u64 tag = (u64)((it + 1) & 0xFF);
    const u64 key = ((u64)((me * 2654435761u + it * 40503u) & 0xFFFFFF) << 8);
    u64 *slot = p.slots + (size_t)(it & 1) * (G * p.stride + 4096);
    u64 *res = slot + (size_t)G * p.stride;               // result lines (mode 2/3): 16 u64 apart
    const long long t0 = clock64();
    if (lane == 0) {
      if (p.sk == 0) st_relaxed(slot + (size_t)me * p.stride, key | tag);
      else if (p.sk == 2) slot[(size_t)me * p.stride] = key | tag;
      else if (p.sk == 3) atomicExch(slot + (size_t)me * p.stride, key | tag);
      else if (p.sk == 4) { st_relaxed(slot + (size_t)me * p.stride, key | tag); __threadfence(); }
      else if (p.sk == 5) atomicMax(slot + (size_t)me * p.stride, ((u64)(it + 1) << 8) | tag);     // RED: nothing returned
    }
    if (p.delay_ns > 0) __nanosleep(p.delay_ns);
    else if (p.delay_ns < 0) { while (clock64() - t0 < -p.delay_ns) { } }     // busy wait: -delay cycles after t0
    const bool full = p.mode == 0 || me < p.pollers;
    u64 m = ~0ull;
    if (full) {
      u64 v[5];
      unsigned pending = 0;
      for (int q = 0; q < 5; q++) { v[q] = ~0ull; if (lane + 32 * q < G) pending |= 1u << q; }
      while (pending) {
        polls++;                                   // lane 0's count; wpolls = the warp's
        const long long tp = clock64();
#pragma unroll
        for (int q = 0; q < 5; q++)
          if (pending & (1u << q)) v[q] = ld_relaxed(slot + (size_t)(lane + 32 * q) * p.stride);
        sink += v[0] + v[1] + v[2] + v[3] + v[4];          // all five have returned
        pollcyc += clock64() - tp;
#pragma unroll
        for (int q = 0; q < 5; q++)
          if ((pending & (1u << q)) && (((v[q] & 0xFF) - tag) & 0xFF) < 128 && v[q] != 0) pending &= ~(1u << q);   // tag reached (mode 1: a non-poller may be up to two ahead)
      }
      for (int q = 0; q < 5; q++)
        if (lane + 32 * q < G) { u64 a = v[q] & ~0xFFull; m = a < m ? a : m; }
      m = warp_min(m);
      if (p.mode == 2 && lane == 0) st_relaxed(res + 16 * me, m | tag);
      if (p.mode == 3 && lane == 0) st_relaxed(res, m | tag);
    } else if (p.mode == 1) {
      u64 v;
      do { polls++; v = ld_relaxed(slot); } while ((v & 0xFF) != tag);      // CTA 0's slot: at most one exchange ahead of it
      m = v;
    } else {
      const u64 *r = p.mode == 2 ? res + 16 * (me % p.pollers) : res;
      u64 v;
      do { polls++; v = ld_relaxed(r); } while ((v & 0xFF) != tag);
      m = v & ~0xFFull;
    }
    sink += m;
    cyc += clock64() - t0;
    __syncthreads();
  }
  if (lane == 0) { p.out[me * 4] = cyc; p.out[me * 4 + 1] = polls; p.out[me * 4 + 2] = (long long)sink; p.out[me * 4 + 3] = pollcyc; }
}

// Staggered polling: K warps of the CTA poll all slots, warp w starting gap cycles after warp w-1, so that a poll
// wave leaves the SM every gap cycles instead of once per round trip; the first warp that has seen every key
// publishes the minimum to the CTA through shared memory.
__global__ void __launch_bounds__(512) exch_stagger(Params p, int K, int gap) {
  const int G = gridDim.x, lane = threadIdx.x & 31, me = blockIdx.x, warp = threadIdx.x >> 5;
  __shared__ volatile u64 result[2];
  __shared__ volatile int done[2];
  if (threadIdx.x < 2) done[threadIdx.x] = -1;
  __syncthreads();
  long long cyc = 0, polls = 0;
  u64 sink = 0;
  for (int it = 0; it < p.iters; it++) {
    long long t = clock64();
    while (clock64() - t < p.work) { }
    __syncthreads();                                      // the CTA minimum is known here (k3: after the block reduce)
    const u64 tag = (u64)((it + 1) & 0xFF);
    const u64 key = ((u64)((me * 2654435761u + it * 40503u) & 0xFFFFFF) << 8);
    u64 *slot = p.slots + (size_t)(it & 1) * (G * p.stride + 4096);
    const long long t0 = clock64();
    if (threadIdx.x == 0) st_relaxed(slot + (size_t)me * p.stride, key | tag);
    if (warp < K) {
      const long long start = -p.delay_ns + (long long)warp * gap;
      while (clock64() - t0 < start) { }
      u64 v[5];
      unsigned pending = 0;
      for (int q = 0; q < 5; q++) { v[q] = ~0ull; if (lane + 32 * q < G) pending |= 1u << q; }
      bool mine = false;
      while (true) {
        if (done[it & 1] == it) break;
        if (warp == 0) polls++;
#pragma unroll
        for (int q = 0; q < 5; q++)
          if (pending & (1u << q)) v[q] = ld_relaxed(slot + (size_t)(lane + 32 * q) * p.stride);
#pragma unroll
        for (int q = 0; q < 5; q++)
          if ((pending & (1u << q)) && (v[q] & 0xFF) == tag) pending &= ~(1u << q);
        if (__all_sync(0xffffffffu, pending == 0)) { mine = true; break; }
      }
      if (mine) {
        u64 m = ~0ull;
        for (int q = 0; q < 5; q++)
          if (lane + 32 * q < G) { u64 a = v[q] & ~0xFFull; m = a < m ? a : m; }
        m = warp_min(m);
        if (lane == 0) { result[it & 1] = m; __threadfence_block(); done[it & 1] = it; }
      }
    }
    __syncthreads();
    sink += result[it & 1];
    cyc += clock64() - t0;
  }
  if (threadIdx.x == 0) { p.out[me * 4] = cyc; p.out[me * 4 + 1] = polls; p.out[me * 4 + 2] = (long long)sink; p.out[me * 4 + 3] = 0; }
}

// Split polling: warp w of five polls slots 32 w .. 32 w + 31 (ONE strong load per lane and wave), the five partial
// minima meet in shared memory behind the CTA barrier that follows the exchange anyway.
__global__ void __launch_bounds__(512) exch_split(Params p) {
  const int G = gridDim.x, lane = threadIdx.x & 31, me = blockIdx.x, warp = threadIdx.x >> 5;
  __shared__ u64 part[2][8];
  long long cyc = 0, polls = 0;
  u64 sink = 0;
  for (int it = 0; it < p.iters; it++) {
    long long t = clock64();
    while (clock64() - t < p.work) { }
    __syncthreads();
    const u64 tag = (u64)((it + 1) & 0xFF);
    const u64 key = ((u64)((me * 2654435761u + it * 40503u) & 0xFFFFFF) << 8);
    u64 *slot = p.slots + (size_t)(it & 1) * (size_t)(32 * (G * 16 + 64));
    const long long t0 = clock64();
    // replicas (p.pollers = R > 1): lane r of warp 0 publishes into copy r of the slot array, CTA c polls copy c % R, so a
    // line is asked for by 1/R of the CTAs
    const int R = p.pollers > 1 ? p.pollers : 1;
    const size_t copy = (size_t)G * p.stride + 64;
    if (threadIdx.x < R) st_relaxed(slot + threadIdx.x * copy + (size_t)me * p.stride, key | tag);
    slot += (me % R) * copy;
    if (warp < 5) {
      while (clock64() - t0 < -p.delay_ns) { }
      const int c = warp * 32 + lane;
      u64 v = ~0ull;
      bool pend = c < G;
      while (__any_sync(0xffffffffu, pend)) {
        if (warp == 0) polls++;
        if (pend) { v = ld_relaxed(slot + (size_t)c * p.stride); if ((v & 0xFF) == tag) pend = false; }
      }
      u64 m = c < G ? (v & ~0xFFull) : ~0ull;
      m = warp_min(m);
      if (lane == 0) part[it & 1][warp] = m;
    }
    __syncthreads();
    u64 m = part[it & 1][0];
    for (int w = 1; w < 5; w++) m = part[it & 1][w] < m ? part[it & 1][w] : m;
    sink += m;
    cyc += clock64() - t0;
  }
  if (threadIdx.x == 0) { p.out[me * 4] = cyc; p.out[me * 4 + 1] = polls; p.out[me * 4 + 2] = (long long)sink; p.out[me * 4 + 3] = 0; }
}

int main(int argc, char **argv) {
  int G = 148, iters = 20000;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  if (prop.multiProcessorCount < G) G = prop.multiProcessorCount;
  u64 *slots;
  long long *out, *h = (long long *)malloc(sizeof(long long) * 4 * G);
  const size_t nslot = 2 * (size_t)(32 * (G * 16 + 64)) + 8192;
  cudaMalloc(&slots, nslot * 8);
  cudaMalloc(&out, sizeof(long long) * 4 * G);
  struct { int mode, stride, pollers, delay, work, sk; } cfg[] = {
      {0, 4, 0, 0, 6000, 0}, {0, 4, 0, -200, 6000, 0}, {0, 4, 0, -300, 6000, 0}, {0, 4, 0, -400, 6000, 0}, {0, 4, 0, -500, 6000, 0}, {0, 4, 0, -600, 6000, 0}, {0, 4, 0, -700, 6000, 0}, {0, 4, 0, -800, 6000, 0}, {0, 4, 0, -1000, 6000, 0}, {0, 4, 0, -1200, 6000, 0}, {0, 16, 0, 0, 6000, 0}, {0, 16, 0, -200, 6000, 0}, {0, 16, 0, -300, 6000, 0}, {0, 16, 0, -400, 6000, 0}, {0, 16, 0, -500, 6000, 0}, {0, 16, 0, -600, 6000, 0}, {0, 16, 0, -700, 6000, 0}, {0, 16, 0, -800, 6000, 0}, {0, 16, 0, -1000, 6000, 0}, {0, 16, 0, -1200, 6000, 0}, {0, 2, 0, -500, 6000, 0}, {0, 8, 0, -500, 6000, 0},
  };
  printf("%-6s %-6s %-8s %-6s %-6s | %-12s %-12s %-8s\n", "mode", "stride", "pollers", "delay", "work", "cyc(pollers)", "cyc(others)",
         "polls");
  for (auto &c : cfg) {
    cudaMemset(slots, 0, nslot * 8);
    Params p{slots, out, iters, c.work, c.delay, c.stride, c.mode, c.pollers, c.sk};
    void *args[] = {&p};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)exch, dim3(G), dim3(64), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, sizeof(long long) * 4 * G, cudaMemcpyDeviceToHost);
    double a = 0, b = 0, pl = 0, pc = 0;
    int na = 0, nb = 0;
    for (int g = 0; g < G; g++) {
      const bool full = c.mode == 0 || g < c.pollers;
      if (full) { a += (double)h[g * 4] / iters; na++; pl += (double)h[g * 4 + 1] / iters; pc += (double)h[g * 4 + 3] / (double)h[g * 4 + 1]; }
      else { b += (double)h[g * 4] / iters; nb++; }
    }
    printf("sk%d %-6d %-6d %-8d %-6d %-6d | %-12.0f %-12.0f %-8.2f cyc/poll %.0f\n", c.sk, c.mode, c.stride, c.pollers, c.delay, c.work, na ? a / na : 0.0,
           nb ? b / nb : 0.0, na ? pl / na : 0.0, na ? pc / na : 0.0);
  }
  printf("staggered polling: K warps, gap cycles, first wave at d0 cycles, slot stride\n");
  struct { int K, gap, d0, stride; } sc[] = {
      {1, 0, 0, 4},   {1, 0, 500, 4}, {2, 320, 300, 4}, {4, 160, 300, 4}, {4, 160, 400, 4}, {4, 160, 500, 4}, {8, 80, 300, 4},
      {8, 80, 400, 4}, {8, 80, 500, 4}, {16, 40, 400, 4}, {4, 200, 300, 16}, {8, 100, 300, 16}, {8, 100, 500, 16}, {16, 50, 400, 16},
  };
  for (auto &c : sc) {
    cudaMemset(slots, 0, nslot * 8);
    Params p{slots, out, iters, 6000, -c.d0, c.stride, 0, 0, 0};
    int K = c.K, gap = c.gap;
    void *args[] = {&p, &K, &gap};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)exch_stagger, dim3(G), dim3(512), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, sizeof(long long) * 4 * G, cudaMemcpyDeviceToHost);
    double a = 0, pl = 0;
    for (int g = 0; g < G; g++) { a += (double)h[g * 4] / iters; pl += (double)h[g * 4 + 1] / iters; }
    printf("K %-3d gap %-4d d0 %-4d stride %-3d | cyc %-8.0f warp-0 polls %.2f\n", c.K, c.gap, c.d0, c.stride, a / G, pl / G);
  }
  printf("split polling: five warps, one load per lane and wave; first wave at d0 cycles, slot stride\n");
  struct { int d0, stride; } pc[] = {{0, 16}, {300, 16}, {500, 16}, {600, 16}, {700, 16}, {0, 8}, {300, 8}, {500, 8}, {600, 8}, {700, 8}, {500, 4}, {600, 4}};
  for (auto &c : pc) {
    cudaMemset(slots, 0, nslot * 8);
    Params p{slots, out, iters, 6000, -c.d0, c.stride, 0, 0, 0};
    void *args[] = {&p};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)exch_split, dim3(G), dim3(512), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, sizeof(long long) * 4 * G, cudaMemcpyDeviceToHost);
    double a = 0, pl = 0;
    for (int g = 0; g < G; g++) { a += (double)h[g * 4] / iters; pl += (double)h[g * 4 + 1] / iters; }
    printf("d0 %-4d stride %-3d | cyc %-8.0f warp-0 waves %.2f\n", c.d0, c.stride, a / G, pl / G);
  }
  printf("split polling with R copies of the slot array (CTA c polls copy c %% R)\n");
  for (int stride : {8, 16})
    for (int R : {1, 2, 4, 8, 16, 32})
      for (int d0 : {200, 300, 400, 500}) {
        cudaMemset(slots, 0, nslot * 8);
        Params p{slots, out, iters, 6000, -d0, stride, 0, R, 0};
        void *args[] = {&p};
        cudaError_t e = cudaLaunchCooperativeKernel((void *)exch_split, dim3(G), dim3(512), args, 0, 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, sizeof(long long) * 4 * G, cudaMemcpyDeviceToHost);
        double a = 0, pl = 0;
        for (int g = 0; g < G; g++) { a += (double)h[g * 4] / iters; pl += (double)h[g * 4 + 1] / iters; }
        printf("stride %-3d R %-3d d0 %-4d | cyc %-8.0f warp-0 waves %.2f\n", stride, R, d0, a / G, pl / G);
      }
  // the same CTA shape (512 threads, barrier before and after) with ONE polling warp, for comparison
  for (int d0 : {500, 600}) {
    cudaMemset(slots, 0, nslot * 8);
    Params p{slots, out, iters, 6000, -d0, 8, 0, 0, 0};
    int K = 1, gap = 0;
    void *args[] = {&p, &K, &gap};
    cudaLaunchCooperativeKernel((void *)exch_stagger, dim3(G), dim3(512), args, 0, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, sizeof(long long) * 4 * G, cudaMemcpyDeviceToHost);
    double a = 0;
    for (int g = 0; g < G; g++) a += (double)h[g * 4] / iters;
    printf("one warp, d0 %d stride 8 | cyc %.0f\n", d0, a / G);
  }
  return 0;
}
