// umma_pair_probe.cu -- evaluation of tcgen05.mma cta_group::2 (256-row UMMA over a CTA pair) for K2's streaming GEMM
// at the C4 shape (rows x 4096 codes x K = 512, fp16 in, fp32 accumulators in TMEM, no-swizzle K-major operands):
//   CG = 1: the product's scheme -- one CTA per SM, its 128-row A tile resident, whole 256-code x 64-K slabs of B staged
//           through a ring (32 KB per stage), 128x256x16 MMAs.
//   CG = 2: a cluster of two CTAs, each with its own 128-row A tile resident and HALF of every B slab (128 codes x 64 K,
//           16 KB per stage, twice the stages), the leader issues 256x256x16 MMAs that read both halves: half the L2->SM
//           bytes per flop.
// The peer's half of a slab is announced to the leader by a relay thread (wait on the local barrier, remote arrive):
// a non-tensor cp.async.bulk cannot complete on a barrier of another CTA than its destination (tried: it hangs).
// Both variants run the same skeleton (producer warp, MMA warp, four epilogue warps that only hand the accumulator
// back), so the difference is the operand feed.  Mode "check" computes one tile pair and compares with the CPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_pair_probe umma_pair_probe.cu
//   ./umma_pair_probe check ; ./umma_pair_probe time [row_tiles]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int TM = 128, TN = 256, KS = 64;          // row tile, code tile, K slab
constexpr int THREADS = 192;                        // warp 0 producer, warp 1 MMA (peer CTA: relay), warps 2-5 epilogue

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// completion of all MMAs issued so far -> one arrival on `bar` (CG = 2: on that barrier in BOTH CTAs of the pair)
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Operand images (built on the host): A [row tile][kc = K/8][128 rows][8 halfs];
// B, CG = 1: [code tile][kc][256 codes][8];  CG = 2: [code tile][half][kc][128 codes][8].
template <int CG>
__global__ void __launch_bounds__(THREADS, 1) gemm(const __half *__restrict__ Aimg, const __half *__restrict__ Bimg, int Kp,
                                                    long ntiles, int nct, int nst, float *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int BN = TN / CG;                             // codes of a B slab held by this CTA
  const uint32_t a_bytes = (uint32_t)TM * Kp * 2, st_bytes = (uint32_t)BN * KS * 2;
  unsigned char *sA = smem;
  unsigned char *sB = smem + a_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)nst * st_bytes);
  uint64_t *full = bars, *empty = bars + 8, *pfull = bars + 16, *tfull = bars + 24, *tempty = bars + 26, *afull = bars + 28;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 30);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const long unit = CG == 2 ? blockIdx.x / 2 : blockIdx.x, nunits = CG == 2 ? gridDim.x / 2 : gridDim.x;
  const long npass = (ntiles + CG - 1) / CG;              // a pass = CG row tiles against all code tiles
  const int nslab = Kp / KS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&pfull[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4 * CG); }
    mbar_init(afull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();                            // the peer's barriers are initialised before anyone arrives on them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  // ---- producer: own A tile per pass, then this CTA's part of every B slab
  if (warp == 0) {
   if (lane == 0) {
    uint32_t it = 0;
    long pass_i = 0;
    for (long ps = unit; ps < npass; ps += nunits, pass_i++) {
      long tile = ps * CG + rank;
      if (tile >= ntiles) tile = ntiles - 1;
      // A of the previous pass is free once its MMAs have completed: commits arrive in order, so the latest phase of
      // every stage's empty barrier covers them all (the ring drains once per pass)
      if (pass_i > 0) {
        for (int s = 0; s < nst; s++) {
          const uint32_t cnt = it > (uint32_t)s ? (it - s + nst - 1) / nst : 0;     // fills of stage s so far
          if (cnt > 0) mbar_wait(&empty[s], (cnt - 1) & 1);
        }
      }
      mbar_expect_tx(afull, a_bytes);
      bulk_g2s(sA, Aimg + (size_t)tile * TM * Kp, a_bytes, afull);
      for (int ct = 0; ct < nct; ct++) {
        const __half *bsrc = Bimg + (size_t)ct * TN * Kp + (CG == 2 ? (size_t)rank * BN * Kp : 0);
        for (int sl = 0; sl < nslab; sl++, it++) {
          const int s = it % nst;
          const uint32_t use = it / nst;
          if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
          mbar_expect_tx(&full[s], st_bytes);
          bulk_g2s(sB + (size_t)s * st_bytes, bsrc + (size_t)sl * (KS / 8) * BN * 8, st_bytes, &full[s]);
        }
      }
    }
   }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---- MMA issuer (leader CTA)
      const bool leader = elect_one();
      const uint32_t idesc = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((TM * CG) >> 4) << 24);
      const uint64_t dA0 = umma_desc(smem_u32(sA), TM * 16, 128);
      uint32_t it = 0, acc_it = 0;
      long pass_i = 0;
      for (long ps = unit; ps < npass; ps += nunits, pass_i++) {
        mbar_wait(afull, pass_i & 1);
        for (int ct = 0; ct < nct; ct++, acc_it++) {
          const int buf = acc_it & 1;
          const uint32_t tuse = acc_it >> 1;
          if (tuse > 0) mbar_wait(&tempty[buf], (tuse - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem_base + buf * TN;
          for (int sl = 0; sl < nslab; sl++, it++) {
            const int s = it % nst;
            const uint32_t use = it / nst;
            mbar_wait(&full[s], use & 1);
            if (CG == 2) mbar_wait(&pfull[s], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader) {
              const uint64_t dB0 = umma_desc(smem_u32(sB + (size_t)s * st_bytes), BN * 16, 128);
#pragma unroll
              for (int k = 0; k < KS / 16; k++) {
                const uint64_t da = dA0 + (uint64_t)(((sl * (KS / 16) + k) * 2 * TM * 16) >> 4);
                const uint64_t db = dB0 + (uint64_t)((k * 2 * BN * 16) >> 4);
                umma<CG>(d_tmem, da, db, idesc, (sl | k) != 0);
              }
              umma_commit<CG>(&empty[s]);
              if (sl == nslab - 1) umma_commit<CG>(&tfull[buf]);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ---- relay (peer CTA): its half of a slab has landed -> tell the leader
      if (lane == 0) {
        uint32_t it = 0;
        long pass_i = 0;
        for (long ps = unit; ps < npass; ps += nunits, pass_i++) {
          mbar_wait(afull, pass_i & 1);                    // this CTA's A tile too: the first slab of a pass vouches for it
          for (int ct = 0; ct < nct; ct++)
            for (int sl = 0; sl < nslab; sl++, it++) {
              const int s = it % nst;
              mbar_wait(&full[s], (it / nst) & 1);
              mbar_arrive_cluster(&pfull[s], 0);
            }
        }
      }
    }
  } else {
    // ---- epilogue warps: hand the accumulator back (check mode: write it out first)
    const int q = warp - 2;                               // TMEM lane quarter of this warp: warp id % 4
    const int quarter = warp & 3;
    (void)q;
    uint32_t acc_it = 0;
    long pass_i = 0;
    for (long ps = unit; ps < npass; ps += nunits, pass_i++) {
      long tile = ps * CG + rank;
      for (int ct = 0; ct < nct; ct++, acc_it++) {
        const int buf = acc_it & 1;
        mbar_wait(&tfull[buf], (acc_it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (out && tile < ntiles) {
          const int row = quarter * 32 + lane;
          for (int c0 = 0; c0 < TN; c0 += 8) {
            uint32_t v[8];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * TN + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 8; c++)
              out[((size_t)tile * TM + row) * ((size_t)nct * TN) + (size_t)ct * TN + c0 + c] = __uint_as_float(v[c]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (CG == 1 || rank == 0) mbar_arrive(&tempty[buf]);
          else mbar_arrive_cluster(&tempty[buf], 0);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync();
  if (warp == 2) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

static void build_images(int CG, long ntiles, int nct, int Kp, const std::vector<float> &A, const std::vector<float> &B,
                         std::vector<__half> &Ai, std::vector<__half> &Bi) {
  Ai.assign((size_t)ntiles * TM * Kp, __float2half(0.0f));
  Bi.assign((size_t)nct * TN * Kp, __float2half(0.0f));
  for (long t = 0; t < ntiles; t++)
    for (int r = 0; r < TM; r++)
      for (int k = 0; k < Kp; k++)
        Ai[(size_t)t * TM * Kp + ((size_t)(k / 8) * TM + r) * 8 + (k % 8)] = __float2half(A[((size_t)t * TM + r) * Kp + k]);
  const int BN = TN / CG;
  for (int ct = 0; ct < nct; ct++)
    for (int c = 0; c < TN; c++)
      for (int k = 0; k < Kp; k++) {
        const int half = c / BN, cl = c % BN;
        Bi[(size_t)ct * TN * Kp + (size_t)half * BN * Kp + ((size_t)(k / 8) * BN + cl) * 8 + (k % 8)] =
            __float2half(B[((size_t)ct * TN + c) * Kp + k]);
      }
}

template <int CG>
static int launch(const __half *dA, const __half *dB, int Kp, long ntiles, int nct, int grid, float *dOut, float *ms, int nst = 0) {
  if (!nst) nst = CG == 2 ? 4 : 2;
  const size_t smem = (size_t)TM * Kp * 2 + (size_t)nst * (TN / CG) * KS * 2 + 32 * 8;
  cudaFuncSetAttribute(gemm<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm<CG>, dA, dB, Kp, ntiles, nct, nst, dOut);
  cudaEventRecord(e1);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CG %d: %s\n", CG, cudaGetErrorString(e)); return 1; }
  cudaEventElapsedTime(ms, e0, e1);
  return 0;
}

int main(int argc, char **argv) {
  const bool check = argc < 2 || strcmp(argv[1], "check") == 0;
  const int Kp = 512;
  if (check) {
    const long ntiles = 3;                                 // ragged: the second pair has one tile
    const int nct = 2;
    std::vector<float> A((size_t)ntiles * TM * Kp), B((size_t)nct * TN * Kp);
    srand(1);
    for (auto &v : A) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto &v : B) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto &v : A) v = __half2float(__float2half(v));   // operands exactly representable in fp16
    for (auto &v : B) v = __half2float(__float2half(v));
    int rc = 0;
    for (int CG = 1; CG <= 2; CG++) {
      std::vector<__half> Ai, Bi;
      build_images(CG, ntiles, nct, Kp, A, B, Ai, Bi);
      __half *dA, *dB; float *dO;
      cudaMalloc(&dA, Ai.size() * 2); cudaMalloc(&dB, Bi.size() * 2);
      cudaMalloc(&dO, (size_t)ntiles * TM * nct * TN * 4);
      cudaMemcpy(dA, Ai.data(), Ai.size() * 2, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, Bi.data(), Bi.size() * 2, cudaMemcpyHostToDevice);
      cudaMemset(dO, 0, (size_t)ntiles * TM * nct * TN * 4);
      float ms;
      if (CG == 1 ? launch<1>(dA, dB, Kp, ntiles, nct, 2, dO, &ms) : launch<2>(dA, dB, Kp, ntiles, nct, 2, dO, &ms, 4)) return 1;
      printf("CG %d kernel done, %.3f ms\n", CG, ms); fflush(stdout);
      std::vector<float> O((size_t)ntiles * TM * nct * TN);
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; long bad = 0;
      for (long r = 0; r < ntiles * TM; r++)
        for (int c = 0; c < nct * TN; c++) {
          double ref = 0;
          for (int k = 0; k < Kp; k++)
            ref += (double)A[(size_t)r * Kp + k] * B[(size_t)c * Kp + k];
          const double err = fabs(ref - O[(size_t)r * nct * TN + c]);
          if (err > maxerr) maxerr = err;
          if (err > 2e-3) { if (bad < 5) printf("CG %d bad r=%ld c=%d got %f ref %f\n", CG, r, c, O[(size_t)r * nct * TN + c], ref); bad++; }
        }
      fflush(stdout);
      printf("CG %d check: max abs err %.3e, bad %ld of %ld\n", CG, maxerr, bad, ntiles * TM * (long)nct * TN);
      rc |= bad != 0;
      cudaFree(dA); cudaFree(dB); cudaFree(dO);
    }
    return rc;
  }
  // ---- timing at the C4 shape: rows x 4096 codes x K 512
  const long ntiles = argc > 2 ? atol(argv[2]) : 7813;
  const int nct = 16;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount & ~1;
  __half *dA, *dB;
  cudaMalloc(&dA, (size_t)ntiles * TM * Kp * 2); cudaMalloc(&dB, (size_t)nct * TN * Kp * 2);
  {
    // random finite halfs in (-1, 1): constant operands would flatter the power draw of the tensor pipe
    std::vector<__half> h((size_t)16 << 20);
    srand(3);
    for (auto &v : h) v = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (size_t off = 0; off < (size_t)ntiles * TM * Kp; off += h.size())
      cudaMemcpy(dA + off, h.data(), std::min(h.size(), (size_t)ntiles * TM * Kp - off) * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, h.data(), (size_t)nct * TN * Kp * 2, cudaMemcpyHostToDevice);
  }
  const double flop = 2.0 * (double)ntiles * TM * nct * TN * Kp;
  struct { int CG, nst; } cfgs[] = {{1, 2}, {2, 4}, {2, 6}};
  for (int rep = 0; rep < 3; rep++)
    for (auto &c : cfgs) {
      float ms;
      if (c.CG == 1 ? launch<1>(dA, dB, Kp, ntiles, nct, sms, nullptr, &ms, c.nst)
                    : launch<2>(dA, dB, Kp, ntiles, nct, sms, nullptr, &ms, c.nst)) return 1;
      printf("rep %d CG %d stages %d: %.3f ms  %.1f TFLOP/s  (B from L2: %.1f GB -> %.2f TB/s)\n", rep, c.CG, c.nst,
             ms, flop / ms * 1e-9, (double)ntiles / c.CG * nct * TN * Kp * 2 * 1e-9,
             (double)ntiles / c.CG * nct * TN * Kp * 2 / ms * 1e-9);
      fflush(stdout);
    }
  return 0;
}
