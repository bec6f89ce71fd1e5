// common.cuh -- shared device helpers for the B200 (sm_100a) BMU engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

// ------------------------------------------------------------------------------------
// Exact FP32 arithmetic.  The reference accumulates fl(fl(m-x)^2) in component order with
// a rounding after every operation (reference lvq_pak.c:70-71, scalar subss/mulss/addss).
// The translation units are compiled with -fmad=false, but ptxas was observed to contract
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 regardless (CUDA 12.9), so the scalar helpers use
// the never-contracted intrinsics and the packed helpers use an .ftz add, which ptxas does
// not fuse with a non-.ftz mul.  The .ftz add is bit-identical to the plain add as long as
// no operand or result is subnormal; rows that could produce subnormals (|v| < 2^-40) never
// reach the packed kernels (see classify in k1_search.cu).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float sq_acc(float acc, float m, float x) {
  float d = __fsub_rn(m, x);
  return __fadd_rn(acc, __fmul_rn(d, d));
}

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2_ftz(u64 a, u64 b) {
  u64 d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes multiple of 16,
// both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------------------------
// relaxed / volatile global accesses for the software grid barrier of K3
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// row classification bits (k1_search.cu)
#define ROW_NONFINITE 1u   // NaN or +-Inf in an unmasked component
#define ROW_TINY 2u        // non-zero magnitude below 2^-40 in an unmasked component
#define ROW_MASKED 4u      // at least one masked component
#define ROW_ALLMASKED 8u   // every component masked
