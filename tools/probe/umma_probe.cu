// umma_probe.cu -- bring-up test for the tcgen05 building blocks used by K2:
//   cp.async.bulk of pre-arranged operand images -> tcgen05.mma (kind::f16, bf16 in, f32
//   accum in TMEM, no-swizzle K-major canonical layout) -> tcgen05.ld -> compare with CPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int M = 128, NT = 256;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

__global__ void __launch_bounds__(160, 1) probe(const __nv_bfloat16 *Aimg, const __nv_bfloat16 *Bimg, int Kp, float *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16 *sA = (__nv_bfloat16 *)smem;                 // Kp/8 chunks x 128 rows x 16 B
  __nv_bfloat16 *sB = sA + (size_t)M * Kp;                   // Kp/8 chunks x 256 rows x 16 B
  uint64_t *bar_full = (uint64_t *)(sB + (size_t)NT * Kp);
  uint64_t *bar_mma = bar_full + 1;
  uint32_t *tmem_ptr = (uint32_t *)(bar_mma + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_ptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4 && lane == 0) {
    uint32_t bytesA = (uint32_t)M * Kp * 2, bytesB = (uint32_t)NT * Kp * 2;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar_full)), "r"(bytesA + bytesB) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sA)), "l"(Aimg), "r"(bytesA), "r"(smem_u32(bar_full)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sB)), "l"(Bimg), "r"(bytesB), "r"(smem_u32(bar_full)) : "memory");
    // wait for the bytes
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}" ::"r"(smem_u32(bar_full)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // instruction descriptor: c=F32(1)<<4, a=BF16(1)<<7, b=BF16(1)<<10, K-major both, N>>3 <<17, M>>4 <<24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < Kp / 16; ks++) {
      uint64_t da = make_desc(smem_u32(sA) + ks * 2 * (M * 16), M * 16, 128);
      uint64_t db = make_desc(smem_u32(sB) + ks * 2 * (NT * 16), NT * 16, 128);
      uint32_t acc = ks > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)) : "memory");
  }
  if (warp < 4) {
    asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2;\nbra W2;\nD2:\n}" ::"r"(smem_u32(bar_mma)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < NT; c0 += 32) {
      uint32_t v[32];
      uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 32; c++) out[(size_t)row * NT + c0 + c] = __uint_as_float(v[c]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
}

int main() {
  const int Kp = 208;
  std::vector<float> A((size_t)M * Kp), B((size_t)NT * Kp);
  srand(1);
  for (auto &v : A) v = (rand() % 2001 - 1000) / 1000.0f;
  for (auto &v : B) v = (rand() % 2001 - 1000) / 1000.0f;
  std::vector<__nv_bfloat16> Ai((size_t)M * Kp), Bi((size_t)NT * Kp);
  std::vector<float> Ar(A.size()), Br(B.size());
  for (int r = 0; r < M; r++) for (int k = 0; k < Kp; k++) {
    __nv_bfloat16 h = __float2bfloat16(A[(size_t)r * Kp + k]);
    Ar[(size_t)r * Kp + k] = __bfloat162float(h);
    Ai[((size_t)(k / 8) * M + r) * 8 + (k % 8)] = h;      // [kc][row][8]  (row = rg*8 + r8: 128 B per row group)
  }
  for (int r = 0; r < NT; r++) for (int k = 0; k < Kp; k++) {
    __nv_bfloat16 h = __float2bfloat16(B[(size_t)r * Kp + k]);
    Br[(size_t)r * Kp + k] = __bfloat162float(h);
    Bi[((size_t)(k / 8) * NT + r) * 8 + (k % 8)] = h;
  }
  __nv_bfloat16 *dA, *dB; float *dO;
  cudaMalloc(&dA, Ai.size() * 2); cudaMalloc(&dB, Bi.size() * 2); cudaMalloc(&dO, (size_t)M * NT * 4);
  cudaMemcpy(dA, Ai.data(), Ai.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bi.data(), Bi.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, (size_t)M * NT * 4);
  size_t smem = (size_t)(M + NT) * Kp * 2 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 160, smem>>>(dA, dB, Kp, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> O((size_t)M * NT);
  cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int r = 0; r < M; r++) for (int c = 0; c < NT; c++) {
    double ref = 0; for (int k = 0; k < Kp; k++) ref += (double)Ar[(size_t)r * Kp + k] * Br[(size_t)c * Kp + k];
    double err = fabs(ref - O[(size_t)r * NT + c]);
    if (err > maxerr) maxerr = err;
    if (err > 1e-3) { if (bad < 5) printf("bad r=%d c=%d got %f ref %f\n", r, c, O[(size_t)r * NT + c], ref); bad++; }
  }
  printf("max abs err %.3e  bad %d of %d\n", maxerr, bad, M * NT);
  return bad != 0;
}
