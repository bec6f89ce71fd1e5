// fp32_issue.cu -- measures the non-FMA FP32 issue peak of one B200: the roofline
// denominator of the exact winner kernel K1 (3 FP32 ops per
// distance element, FMA contraction forbidden by the parity rule).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32_issue fp32_issue.cu
//   run  : ./fp32_issue            (prints lane-ops/clk/SM and T lane-ops/s per variant)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 16
#define ITERS 4096

__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b){uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b){uint64_t d; asm volatile("mul.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c){uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
__device__ __forceinline__ uint64_t add2ftz(uint64_t a, uint64_t b){uint64_t d; asm volatile("add.rn.ftz.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b){uint64_t d; asm volatile("sub.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ uint64_t pack(float lo, float hi){return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo);}
__device__ __forceinline__ float lo32(uint64_t v){return __uint_as_float((uint32_t)v);}
__device__ __forceinline__ float hi32(uint64_t v){return __uint_as_float((uint32_t)(v >> 32));}
__device__ __forceinline__ float fadd(float a, float b){float d; asm volatile("add.rn.f32 %0, %1, %2;":"=f"(d):"f"(a),"f"(b)); return d;}
__device__ __forceinline__ float fmul(float a, float b){float d; asm volatile("mul.rn.f32 %0, %1, %2;":"=f"(d):"f"(a),"f"(b)); return d;}
__device__ __forceinline__ float ffma(float a, float b, float c){float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(d):"f"(a),"f"(b),"f"(c)); return d;}

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float seed, long long* clk) {
  float a[CHAINS]; uint64_t p[CHAINS];
  for (int i = 0; i < CHAINS; i++) { a[i] = seed + i; p[i] = ((uint64_t)__float_as_uint(seed + i) << 32) | __float_as_uint(seed * i); }
  float c = seed * 0.5f; uint64_t pc = ((uint64_t)__float_as_uint(c) << 32) | __float_as_uint(c);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
      if (MODE == 0) a[i] = fadd(a[i], c);
      if (MODE == 1) a[i] = fmul(a[i], c);
      if (MODE == 2) a[i] = ffma(a[i], c, c);
      if (MODE == 3) p[i] = add2(p[i], pc);
      if (MODE == 4) p[i] = mul2(p[i], pc);
      if (MODE == 5) p[i] = fma2(p[i], pc, pc);
      if (MODE == 6) { a[i] = fadd(a[i], c); a[i] = fmul(a[i], c); a[i] = fadd(a[i], c); }       // K1 scalar mix
      if (MODE == 7) { uint64_t d = sub2(p[i], pc); uint64_t s = pack(fmul(lo32(d), lo32(d)), fmul(hi32(d), hi32(d))); p[i] = add2(p[i], s); }   // sub2, 2x scalar mul, add2
      if (MODE == 8) { uint64_t d = sub2(p[i], pc); uint64_t s = mul2(d, d); p[i] = add2ftz(p[i], s); }                                      // sub2, mul2, add2.ftz
      if (MODE == 10) { uint64_t d = sub2(p[i], pc); uint64_t s = mul2(d, d); p[i] = pack(fadd(lo32(p[i]), lo32(s)), fadd(hi32(p[i]), hi32(s))); } // sub2, mul2, 2x scalar add
      if (MODE == 11) { float d = fadd(a[i], c); float s = fmul(d, d); a[i] = fadd(a[i], s); }                                                // scalar K1 chain
      if (MODE == 9) { a[i] = fminf(a[i], c); }                                               // FMNMX (alu pipe)
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < CHAINS; i++) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, double lane_ops_per_iter_chain, int nsm, float* out, long long* clk, int warps_per_sm) {
  int blocks = nsm * (warps_per_sm / 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<blocks, 256>>>(out, 1.0001f, clk); cudaDeviceSynchronize();
  cudaEventRecord(e0); bench<MODE><<<blocks, 256>>>(out, 1.0001f, clk); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[4096]; cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
  double ops_sm = (double)ITERS * CHAINS * lane_ops_per_iter_chain * 32.0 * warps_per_sm;   // lane-ops per SM
  double total = ops_sm * nsm;
  printf("%-34s warps/SM=%2d  %.2f lane-ops/clk/SM   %.2f T lane-ops/s   (%.3f ms, %.0f clk => %.0f MHz)\n", name, warps_per_sm, ops_sm / avg,
         total / (ms * 1e-3) / 1e12, ms, avg, avg / (ms * 1e-3) / 1e6);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int nsm = p.multiProcessorCount;
  printf("device %s  SMs %d  smemOptin %zu  clock %d kHz  L2 %d MB\n", p.name, nsm, p.sharedMemPerBlockOptin, p.clockRate, p.l2CacheSize >> 20);
  float* out; long long* clk; cudaMalloc(&out, sizeof(float) * 256 * 4096); cudaMalloc(&clk, sizeof(long long) * 4096);
  for (int w : {8, 16, 32}) {
    run<0>("FADD scalar", 1, nsm, out, clk, w);
    run<1>("FMUL scalar", 1, nsm, out, clk, w);
    run<2>("FFMA scalar (1 lane-op)", 1, nsm, out, clk, w);
    run<3>("FADD2 packed (2 lane-ops)", 2, nsm, out, clk, w);
    run<4>("FMUL2 packed (2 lane-ops)", 2, nsm, out, clk, w);
    run<5>("FFMA2 packed (2 lane-ops)", 2, nsm, out, clk, w);
    run<6>("K1 mix scalar: add,mul,add", 3, nsm, out, clk, w);
    run<11>("K1 chain scalar: sub,mul,add (3)", 3, nsm, out, clk, w);
    run<7>("K1 chain: sub2,mul,mul,add2 (6)", 6, nsm, out, clk, w);
    run<10>("K1 chain: sub2,mul2,add,add (6)", 6, nsm, out, clk, w);
    run<8>("K1 chain: sub2,mul2,add2.ftz (6)", 6, nsm, out, clk, w);
    run<9>("FMNMX scalar", 1, nsm, out, clk, w);
  }
  return 0;
}
