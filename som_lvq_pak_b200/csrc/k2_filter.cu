// k2_filter.cu -- K2: tcgen05 tensor-core GEMM filter + exact FP32 re-rank (sm_100a).
//
// For large M*D the search  argmin_j ||x - m_j||^2  is dominated by the contraction x.m_j.
// K2 computes an APPROXIMATE score  s~_j = ||m'_j||^2 - 2 x'.m'_j  (primes: vectors centred
// on the codebook mean) on the 5th-generation tensor cores, keeps a few candidates per row
// in the GEMM epilogue, and then decides the winner(s) with the reference's EXACT FP32 sum
// (lvq_pak.c:63-73) over those candidates only.  A per-row certificate proves that no code
// outside the candidate set can win or tie; rows whose certificate fails are answered by the
// exact kernel K1 (k1_warp_kernel).  Results are therefore bit-identical to K1 / the reference.
//
//   operands   bf16 3-term split:  x'.m' ~ xh.mh + xh.ml + xl.mh   (K = 3*Dp + 3 -> Kp)
//              the -2 factor and ||m'||^2 (3 bf16 terms against a column of ones) are folded
//              into the B operand, so the accumulator IS the score: no FP32 op per element
//   GEMM       tcgen05.mma.cta_group::1.kind::f16, M=128 x N=256 x K=16 per instruction,
//              FP32 accumulators double-buffered in TMEM (2 x 256 columns), operands staged
//              by cp.async.bulk (UBLKCP) from images pre-arranged in the canonical no-swizzle
//              K-major core-matrix layout, mbarrier ring, one MMA-issuing thread
//   epilogue   4 warps, tcgen05.ld 32x32b.x32 (one row per thread); column index packed into
//              the 8 low mantissa bits, branch-free top-2/top-4 per code tile with FMNMX,
//              merged into a per-row sorted candidate list
//   re-rank    exact distances of the candidates, reference tie rules, certificate
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "k2_filter.h"

namespace bmu {

constexpr int K2_TM = 128;   // rows per sample tile (UMMA M)
constexpr int K2_TN = 256;   // codes per code tile (UMMA N)
constexpr int K2_KS = 64;    // K elements per pipeline stage (4 MMAs)
constexpr int K2_NSTAGE = 4;
constexpr int K2_THREADS = 192;   // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
constexpr int K2_ARES_MAX_KP = 320;   // A image stays resident in smem up to this Kp
constexpr int K2_GROUP_MAX_KP = 0;    // group-minimum epilogue (k == 1) up to this Kp; 0 = off: its
                                      // re-rank gathers 64 code rows per sample and costs more than it saves

struct CbStats {     // maxima over the codebook (centred), device side
  float nm;          // max ||m'||
  float nmlo;        // max ||m'_lo||
  float nrm;         // max ||m' - m'_hi - m'_lo||
  float nm2;         // max ||m'||^2
};

__host__ __device__ inline int k2_dp(int D) { return (D + 7) & ~7; }
__host__ __device__ inline int k2_kp(int D) { return (3 * k2_dp(D) + 3 + 15) & ~15; }

__device__ __forceinline__ void atomic_max_pos(float *addr, float v) {
  atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));   // non-negative floats order as ints
}

// ---------------------------------------------------------------- codebook side
__global__ void k2_mean_kernel(const float *__restrict__ codes, long M, int D, float *__restrict__ mean) {
  // one thread per component; deterministic (sequential double sum)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D) return;
  double s = 0.0;
  for (long j = 0; j < M; j++) s += (double)codes[j * D + i];
  mean[i] = (float)(s / (double)M);
}

// one warp per code: centre, split, write the B image, accumulate the maxima
__global__ void __launch_bounds__(256)
k2_cb_prep_kernel(const float *__restrict__ codes, long M, int D, const float *__restrict__ mean,
                  __nv_bfloat16 *__restrict__ Bimg, CbStats *__restrict__ st) {
  const int lane = threadIdx.x & 31;
  const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nct = (M + K2_TN - 1) / K2_TN;
  if (w >= nct * K2_TN) return;
  const int Dp = k2_dp(D), Kp = k2_kp(D);
  const long ct = w / K2_TN;
  const int r = (int)(w % K2_TN);
  __nv_bfloat16 *img = Bimg + ct * (long)K2_TN * Kp;           // [kc][TN][8]
  auto put = [&](int k, float v) { img[((long)(k >> 3) * K2_TN + r) * 8 + (k & 7)] = __float2bfloat16(v); };
  // zero everything this row owns first (pads included)
  for (int k = lane; k < Kp; k += 32) put(k, 0.0f);
  __syncwarp();
  if (w >= M) {            // padding code: a huge score so it never becomes a candidate
    if (lane == 0) put(3 * Dp, 1e30f);
    return;
  }
  double n2 = 0.0, nlo2 = 0.0, nr2 = 0.0;
  for (int i = lane; i < D; i += 32) {
    float c = __fsub_rn(codes[w * D + i], mean[i]);
    __nv_bfloat16 h = __float2bfloat16(c);
    float lo_f = __fsub_rn(c, __bfloat162float(h));
    __nv_bfloat16 l = __float2bfloat16(lo_f);
    float res = __fsub_rn(lo_f, __bfloat162float(l));
    // the factor -2 is exact in bf16
    put(i, -2.0f * __bfloat162float(h));            // pairs with x_hi
    put(Dp + i, -2.0f * __bfloat162float(l));       // pairs with x_hi
    put(2 * Dp + i, -2.0f * __bfloat162float(h));   // pairs with x_lo
    n2 += (double)c * c;
    nlo2 += (double)__bfloat162float(l) * __bfloat162float(l);
    nr2 += (double)res * res;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    n2 += __shfl_xor_sync(0xffffffffu, n2, off);
    nlo2 += __shfl_xor_sync(0xffffffffu, nlo2, off);
    nr2 += __shfl_xor_sync(0xffffffffu, nr2, off);
  }
  if (lane == 0) {
    float nf = (float)n2;
    __nv_bfloat16 a = __float2bfloat16(nf);
    float r1 = nf - __bfloat162float(a);
    __nv_bfloat16 b = __float2bfloat16(r1);
    float r2 = r1 - __bfloat162float(b);
    __nv_bfloat16 c3 = __float2bfloat16(r2);
    put(3 * Dp, __bfloat162float(a));
    put(3 * Dp + 1, __bfloat162float(b));
    put(3 * Dp + 2, __bfloat162float(c3));
    const float up = 1.0001f;
    atomic_max_pos(&st->nm, (float)sqrt(n2) * up);
    atomic_max_pos(&st->nmlo, (float)sqrt(nlo2) * up);
    atomic_max_pos(&st->nrm, (float)sqrt(nr2) * up);
    atomic_max_pos(&st->nm2, (float)n2 * up);
  }
}

// ---------------------------------------------------------------- row side
// One CTA per tile of 128 rows: classification + work lists (same rules as K1's
// data_prep_kernel), centred bf16 split written as the A image [kc][128][8], row bounds.
struct RowStats {
  double nx2;     // ||x'||^2
  float nx;       // ||x'||      (rounded up)
  float nxlo;     // ||x'_lo||
  float nrx;      // ||x' - x'_hi - x'_lo||
  float pad;
};

__device__ __forceinline__ unsigned k2_classify(float v) {
  unsigned b = __float_as_uint(v) & 0x7fffffffu;
  unsigned f = 0;
  if (b >= 0x7f800000u) f |= ROW_NONFINITE;
  if (b != 0u && b < 0x2b800000u) f |= ROW_TINY;
  return f;
}

__device__ __forceinline__ uint32_t bf16x2_bits(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);      // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t *>(&v);
}

constexpr int K2_PS = 32;     // components per row-prep slab

__global__ void __launch_bounds__(256)
k2_row_prep_kernel(const float *__restrict__ data, const unsigned char *__restrict__ mask, long N,
                   int D, int k, const float *__restrict__ mean, __nv_bfloat16 *__restrict__ Aimg,
                   RowStats *__restrict__ rs, unsigned char *__restrict__ flags,
                   int *__restrict__ listW, int *__restrict__ listS, int *__restrict__ counters,
                   int32_t *__restrict__ idx, float *__restrict__ diff, int32_t *__restrict__ nfound) {
  // staged through shared memory so that both the row reads and the image writes are coalesced
  __shared__ float xs[K2_TM][K2_PS + 1];                      // centred inputs of the slab
  __shared__ unsigned char ms[K2_TM][K2_PS];                  // mask bytes of the slab
  __shared__ __align__(16) uint4 img_s[3][K2_PS / 8][K2_TM];  // three image regions of the slab
  // row-combine scratch aliases the image staging buffer (used only after the slab loop)
  double (*red)[3] = reinterpret_cast<double (*)[3]>(&img_s[0][0][0]);
  int (*redi)[2] = reinterpret_cast<int (*)[2]>(&img_s[1][0][0]);
  const int Dp = k2_dp(D), Kp = k2_kp(D);
  const long tile = blockIdx.x;
  const long n0 = tile * K2_TM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (K2_TM - 1), half = tid >> 7;           // thread = (row, 16-component half)
  uint4 *img = reinterpret_cast<uint4 *>(Aimg + tile * (long)K2_TM * Kp);   // [kc][128] uint4
  const long nrow = n0 + row;
  double n2 = 0.0, nlo2 = 0.0, nr2 = 0.0;
  unsigned f = 0;
  int nmasked = 0;

  for (int d0 = 0; d0 < Dp; d0 += K2_PS) {
    __syncthreads();
    // phase 1: coalesced load of 128 rows x 32 components, centred
    for (int r = warp; r < K2_TM; r += 8) {
      const long n = n0 + r;
      const int i = d0 + lane;
      float v = 0.0f;
      unsigned char mk = 1;                                     // beyond D / N counts as "masked" (zero)
      if (n < N && i < D) {
        v = data[n * (long)D + i];
        mk = mask ? mask[n * (long)D + i] : 0;
      }
      xs[r][lane] = v;
      ms[r][lane] = mk;
    }
    __syncthreads();
    // phase 2: split; each thread packs 2 x 8 components of one row
#pragma unroll
    for (int grp = 0; grp < 2; grp++) {
      uint32_t hi_w[4], lo_w[4];
#pragma unroll
      for (int p = 0; p < 4; p++) {
        float hv[2], lv[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int il = half * 16 + grp * 8 + p * 2 + q;
          const int i = d0 + il;
          hv[q] = 0.0f; lv[q] = 0.0f;
          if (i < D && nrow < N) {
            if (ms[row][il]) { nmasked++; }
            else {
              const float v = xs[row][il];
              f |= k2_classify(v);
              const float c = __fsub_rn(v, mean[i]);
              const __nv_bfloat16 h = __float2bfloat16(c);
              const float lo_f = __fsub_rn(c, __bfloat162float(h));
              const __nv_bfloat16 l = __float2bfloat16(lo_f);
              const float res = __fsub_rn(lo_f, __bfloat162float(l));
              hv[q] = __bfloat162float(h);
              lv[q] = __bfloat162float(l);
              n2 += (double)c * c;
              nlo2 += (double)lv[q] * lv[q];
              nr2 += (double)res * res;
            }
          }
        }
        hi_w[p] = bf16x2_bits(hv[0], hv[1]);
        lo_w[p] = bf16x2_bits(lv[0], lv[1]);
      }
      const int ch = half * 2 + grp;                            // chunk of 8 inside the slab
      const uint4 hq = make_uint4(hi_w[0], hi_w[1], hi_w[2], hi_w[3]);
      const uint4 lq = make_uint4(lo_w[0], lo_w[1], lo_w[2], lo_w[3]);
      img_s[0][ch][row] = hq;       // x_hi against -2 m_hi
      img_s[1][ch][row] = hq;       // x_hi against -2 m_lo
      img_s[2][ch][row] = lq;       // x_lo against -2 m_hi
    }
    __syncthreads();
    // phase 3: the three regions are contiguous runs of (chunks x 128) uint4 in the image
    const int nch = min(K2_PS, Dp - d0) / 8;
    for (int t = tid; t < 3 * nch * K2_TM; t += 256) {
      const int reg = t / (nch * K2_TM), rem = t % (nch * K2_TM);
      const int ch = rem / K2_TM, r = rem % K2_TM;
      img[((long)(reg * Dp + d0) / 8 + ch) * K2_TM + r] = img_s[reg][ch][r];
    }
  }
  // tail: the column(s) of ones that pick up ||m'||^2, then zero padding up to Kp
  for (int t = tid; t < (Kp - 3 * Dp) / 8 * K2_TM; t += 256) {
    const int ch = t / K2_TM, r = t % K2_TM;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ch == 0 && n0 + r < N) { v.x = bf16x2_bits(1.0f, 1.0f); v.y = bf16x2_bits(1.0f, 0.0f); }
    img[((long)(3 * Dp) / 8 + ch) * K2_TM + r] = v;
  }
  // combine the two halves of every row
  __syncthreads();
  if (half == 1) { red[row][0] = n2; red[row][1] = nlo2; red[row][2] = nr2; redi[row][0] = (int)f; redi[row][1] = nmasked; }
  __syncthreads();
  if (half == 0 && nrow < N) {
    n2 += red[row][0]; nlo2 += red[row][1]; nr2 += red[row][2];
    f |= (unsigned)redi[row][0]; nmasked += redi[row][1];
    if (nmasked > 0) f |= ROW_MASKED;
    if (nmasked == D) f |= ROW_ALLMASKED;
    const long n = nrow;
    flags[n] = (unsigned char)f;
    RowStats s;
    const float up = 1.0001f;
    s.nx2 = n2;
    s.nx = (float)sqrt(n2) * up;
    s.nxlo = (float)sqrt(nlo2) * up;
    s.nrx = (float)sqrt(nr2) * up;
    s.pad = 0.0f;
    rs[n] = s;
    if (f & ROW_ALLMASKED) {
      nfound[n] = 0;
      for (int t = 0; t < k; t++) { idx[n * k + t] = -1; diff[n * k + t] = (k == 1) ? -1.0f : FLT_MAX; }
    } else if (f & ROW_NONFINITE) {
      listS[atomicAdd(&counters[1], 1)] = (int)n;
    } else if (f & (ROW_TINY | ROW_MASKED)) {
      listW[atomicAdd(&counters[0], 1)] = (int)n;
    }
  }
}

// ---------------------------------------------------------------- GEMM + fused top-k
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // no-swizzle K-major canonical layout: core matrix = 8 rows x 16 B, contiguous 128 B;
  // LBO = byte distance between the two K chunks of one MMA, SBO = between 8-row groups
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version 1 (sm_100)
  return d;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]) :: "memory");
}

struct K2Smem {
  static size_t bytes(int Kp, bool a_res) {
    size_t stage = (size_t)K2_TN * K2_KS * 2 + (a_res ? 0 : (size_t)K2_TM * K2_KS * 2);
    return (a_res ? (size_t)K2_TM * Kp * 2 : 0) + K2_NSTAGE * stage + 256;
  }
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// tcgen05.wait::ld with the destination registers as in/out operands: the compiler must not
// schedule any use of v[] between the asynchronous tcgen05.ld and this wait
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]) :: "memory");
}

__device__ __forceinline__ float min32(const uint32_t (&v)[32]) {
  // 3-input minima (FMNMX3): 32 -> 11 -> 4 -> 2 -> 1
  float a[11];
#pragma unroll
  for (int t = 0; t < 10; t++)
    a[t] = fminf(fminf(__uint_as_float(v[3 * t]), __uint_as_float(v[3 * t + 1])), __uint_as_float(v[3 * t + 2]));
  a[10] = fminf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  float b0 = fminf(fminf(a[0], a[1]), a[2]), b1 = fminf(fminf(a[3], a[4]), a[5]);
  float b2 = fminf(fminf(a[6], a[7]), a[8]), b3 = fminf(a[9], a[10]);
  return fminf(fminf(b0, b1), fminf(b2, b3));
}

// TG: candidates kept per row, TT: smallest keys tracked per code tile.
// GROUP mode (k == 1): the epilogue only tracks the minimum of every group of 32 columns
// (~0.8 ALU op per score instead of 4) and keeps the TG best GROUPS per row; all 32 codes of
// those groups are re-ranked exactly, and every code outside them has a score >= the
// (TG+1)-th smallest group minimum, which is what the certificate needs.
template <int TG, int TT, bool GROUP>
__global__ void __launch_bounds__(K2_THREADS, 1)
k2_gemm_kernel(const __nv_bfloat16 *__restrict__ Aimg, const __nv_bfloat16 *__restrict__ Bimg, long N,
               long M, int Kp, int a_res, int32_t *__restrict__ cand, float *__restrict__ thr) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const size_t a_res_bytes = a_res ? (size_t)K2_TM * Kp * 2 : 0;
  const size_t b_stage_bytes = (size_t)K2_TN * K2_KS * 2;
  const size_t a_stage_bytes = a_res ? 0 : (size_t)K2_TM * K2_KS * 2;
  const size_t stage_bytes = b_stage_bytes + a_stage_bytes;
  unsigned char *sAres = smem;
  unsigned char *sStage = smem + a_res_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStage + K2_NSTAGE * stage_bytes);
  uint64_t *full = bars, *empty = bars + K2_NSTAGE;
  uint64_t *tfull = bars + 2 * K2_NSTAGE, *tempty = tfull + 2;
  uint64_t *afull = tempty + 2, *aempty = afull + 1;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(aempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long ntiles = (N + K2_TM - 1) / K2_TM;
  const int nct = (int)((M + K2_TN - 1) / K2_TN);
  const int nslab = (Kp + K2_KS - 1) / K2_KS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < K2_NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    mbar_init(afull, 1);
    mbar_init(aempty, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      unsigned seq = 0, tcount = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
        const unsigned char *gA = reinterpret_cast<const unsigned char *>(Aimg) + (size_t)tile * K2_TM * Kp * 2;
        if (a_res) {
          mbar_wait(aempty, (tcount & 1) ^ 1);           // previous tile's MMAs have consumed A
          // a bulk copy moves < 1 MiB; the resident image is at most 80 KB: one transaction
          mbar_arrive_expect_tx(afull, (uint32_t)a_res_bytes);
          bulk_g2s(sAres, gA, (uint32_t)a_res_bytes, afull);
        }
        for (int ct = 0; ct < nct; ct++) {
          const unsigned char *gB = reinterpret_cast<const unsigned char *>(Bimg) + (size_t)ct * K2_TN * Kp * 2;
          for (int sl = 0; sl < nslab; sl++, seq++) {
            const int st = seq % K2_NSTAGE;
            const int kc = min(K2_KS, Kp - sl * K2_KS);               // K elements in this slab
            mbar_wait(&empty[st], ((seq / K2_NSTAGE) & 1) ^ 1);
            const uint32_t bB = (uint32_t)K2_TN * kc * 2, bA = a_res ? 0u : (uint32_t)K2_TM * kc * 2;
            mbar_arrive_expect_tx(&full[st], bB + bA);
            unsigned char *dst = sStage + (size_t)st * stage_bytes;
            bulk_g2s(dst, gB + (size_t)sl * K2_KS * K2_TN * 2, bB, &full[st]);
            if (!a_res) bulk_g2s(dst + b_stage_bytes, gA + (size_t)sl * K2_KS * K2_TM * 2, bA, &full[st]);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=BF16, both K-major, N=256, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(K2_TN >> 3) << 17) |
                             ((uint32_t)(K2_TM >> 4) << 24);
      unsigned seq = 0, acc_seq = 0, tcount = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
        if (a_res) { mbar_wait(afull, tcount & 1); tc_fence_after(); }
        for (int ct = 0; ct < nct; ct++, acc_seq++) {
          const int buf = acc_seq & 1;
          mbar_wait(&tempty[buf], ((acc_seq >> 1) & 1) ^ 1);        // epilogue drained this buffer
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * K2_TN;
          for (int sl = 0; sl < nslab; sl++, seq++) {
            const int st = seq % K2_NSTAGE;
            const int kc = min(K2_KS, Kp - sl * K2_KS);
            mbar_wait(&full[st], (seq / K2_NSTAGE) & 1);
            tc_fence_after();
            const uint32_t bBase = smem_u32(sStage + (size_t)st * stage_bytes);
            const uint32_t aBase = a_res ? smem_u32(sAres) + (uint32_t)sl * K2_KS * K2_TM * 2
                                         : bBase + (uint32_t)b_stage_bytes;
            for (int kk = 0; kk < kc / 16; kk++) {
              // one MMA consumes K=16 = two 8-element chunks, K2_T? * 16 bytes apart
              const uint64_t da = umma_desc(aBase + kk * 2 * (K2_TM * 16), K2_TM * 16, 128);
              const uint64_t db = umma_desc(bBase + kk * 2 * (K2_TN * 16), K2_TN * 16, 128);
              umma_bf16(d_tmem, da, db, idesc, (sl | kk) ? 1u : 0u);
            }
            umma_commit(&empty[st]);                // smem slot reusable once these MMAs retire
          }
          umma_commit(&tfull[buf]);                 // accumulator ready for the epilogue
        }
        if (a_res) umma_commit(aempty);
      }
    }
  } else {
    if constexpr (GROUP) {
    // ===================== epilogue, group mode: one row per thread =====================
    unsigned acc_seq = 0;
    const int row = warp * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      float gk[TG + 1];
      int gi[TG + 1];
#pragma unroll
      for (int t = 0; t <= TG; t++) { gk[t] = INFINITY; gi[t] = -1; }
      for (int ct = 0; ct < nct; ct++, acc_seq++) {
        const int buf = acc_seq & 1;
        mbar_wait(&tfull[buf], (acc_seq >> 1) & 1);
        tc_fence_after();
        // tile-level top-(TG+1) group minima; the 3 low mantissa bits carry the group number
        float tk[TG + 1];
#pragma unroll
        for (int t = 0; t <= TG; t++) tk[t] = INFINITY;
        uint32_t va[32], vb[32];
        tmem_ld32_nowait(lane_base + buf * K2_TN, va);
#pragma unroll
        for (int g = 0; g < K2_TN / 32; g += 2) {
          tmem_ld_wait32(va);
          tmem_ld32_nowait(lane_base + buf * K2_TN + (g + 1) * 32, vb);     // overlap with the min tree
          {
            float key = __uint_as_float((__float_as_uint(min32(va)) & 0xFFFFFFF8u) | (uint32_t)g);
#pragma unroll
            for (int t = 0; t <= TG; t++) { float lo = fminf(tk[t], key); key = fmaxf(tk[t], key); tk[t] = lo; }
          }
          tmem_ld_wait32(vb);
          if (g + 2 < K2_TN / 32) tmem_ld32_nowait(lane_base + buf * K2_TN + (g + 2) * 32, va);
          {
            float key = __uint_as_float((__float_as_uint(min32(vb)) & 0xFFFFFFF8u) | (uint32_t)(g + 1));
#pragma unroll
            for (int t = 0; t <= TG; t++) { float lo = fminf(tk[t], key); key = fmaxf(tk[t], key); tk[t] = lo; }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
#pragma unroll
        for (int t = 0; t <= TG; t++) {
          float key = tk[t];
          int j = ct * (K2_TN / 32) + (int)(__float_as_uint(key) & 7u);      // global group index
          if (key < gk[TG]) {
            bool ins = false;
#pragma unroll
            for (int p = 0; p <= TG; p++) {
              if (ins || key < gk[p]) {
                float t2 = gk[p]; int i2 = gi[p];
                gk[p] = key; gi[p] = j;
                key = t2; j = i2;
                ins = true;
              }
            }
          }
        }
      }
      const long n = tile * K2_TM + row;
      if (n < N) {
#pragma unroll
        for (int t = 0; t < TG; t++) cand[n * TG + t] = gi[t];
        thr[n] = gk[TG];           // every group that was not kept has a minimum >= this key
      }
    }
    } else {
    // ===================== epilogue: one row per thread =====================
    unsigned acc_seq = 0;
    const int row = warp * 32 + lane;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      float gk[TG];
      int gi[TG];
#pragma unroll
      for (int t = 0; t < TG; t++) { gk[t] = INFINITY; gi[t] = -1; }
      float tmin = INFINITY;
      for (int ct = 0; ct < nct; ct++, acc_seq++) {
        const int buf = acc_seq & 1;
        mbar_wait(&tfull[buf], (acc_seq >> 1) & 1);
        tc_fence_after();
        float b[TT];
#pragma unroll
        for (int t = 0; t < TT; t++) b[t] = INFINITY;
        // two register buffers: the tcgen05.ld of the next 32 columns is in flight while the
        // current 32 are folded into the tile's TT smallest keys
        uint32_t va[32], vb[32];
        const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * K2_TN;
        auto fold = [&](const uint32_t (&v)[32], int c0) {
          if constexpr (TT == 2) {
            // pairs: 2 LOP3 + min,max,min,max,min3 = 3.5 ALU ops per score instead of 4
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const float k0 = __uint_as_float((v[c] & 0xFFFFFF00u) | (uint32_t)(c0 + c));
              const float k1 = __uint_as_float((v[c + 1] & 0xFFFFFF00u) | (uint32_t)(c0 + c + 1));
              const float lo = fminf(k0, k1), hi = fmaxf(k0, k1);
              const float t = fmaxf(b[0], lo);
              b[0] = fminf(b[0], lo);
              b[1] = fminf(fminf(b[1], hi), t);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; c++) {
              // 8 low mantissa bits <- column index inside the tile (perturbs the score by < 2^-15 |s|)
              float key = __uint_as_float((v[c] & 0xFFFFFF00u) | (uint32_t)(c0 + c));
#pragma unroll
              for (int t = 0; t < TT; t++) {
                float lo = fminf(b[t], key);
                key = fmaxf(b[t], key);
                b[t] = lo;
              }
            }
          }
        };
        tmem_ld32_nowait(tbase, va);
#pragma unroll 1
        for (int c0 = 0; c0 < K2_TN; c0 += 64) {
          tmem_ld_wait32(va);
          tmem_ld32_nowait(tbase + c0 + 32, vb);
          fold(va, c0);
          tmem_ld_wait32(vb);
          if (c0 + 64 < K2_TN) tmem_ld32_nowait(tbase + c0 + 64, va);
          fold(vb, c0 + 32);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        // everything of this tile that is not kept is >= the tile's TT-th smallest key
        tmin = fminf(tmin, b[TT - 1]);
#pragma unroll
        for (int t = 0; t < TT; t++) {
          float key = b[t];
          int j = ct * K2_TN + (int)(__float_as_uint(key) & 0xFFu);
          if (key < gk[TG - 1]) {
            // sorted insertion with static indexing: once placed, everything below shifts down
            // and the old last entry (the largest) is the one that drops out
            bool ins = false;
#pragma unroll
            for (int p = 0; p < TG; p++) {
              if (ins || key < gk[p]) {
                float tk = gk[p]; int ti = gi[p];
                gk[p] = key; gi[p] = j;
                key = tk; j = ti;
                ins = true;
              }
            }
          }
        }
      }
      const long n = tile * K2_TM + row;
      if (n < N) {
#pragma unroll
        for (int t = 0; t < TG; t++) cand[n * TG + t] = gi[t];
        // candidates dropped from the list are >= its last key; kept ones are re-ranked exactly
        thr[n] = fminf(tmin, gk[TG - 1]);
      }
    }
  }
    }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ---------------------------------------------------------------- exact re-rank + certificate
// LPR lanes per row (power of two >= TG): lane g of a row's group computes the exact distance
// of candidate g, so the loads of x are shared by the group and every lane streams one code
// row; the group leader then orders the candidates and evaluates the certificate.
template <int TG, int LPR>
__global__ void __launch_bounds__(256)
k2_rerank_kernel(const float *__restrict__ data, const float *__restrict__ codes, long N, long M, int D,
                 int k, int Kp, const unsigned char *__restrict__ flags, const RowStats *__restrict__ rs,
                 const CbStats *__restrict__ cst, const int32_t *__restrict__ cand,
                 const float *__restrict__ thr, int *__restrict__ listW, int *__restrict__ counters,
                 int32_t *__restrict__ idx, float *__restrict__ diff, int32_t *__restrict__ nfound) {
  constexpr int RPW = 32 / LPR;                         // rows per warp
  const int lane = threadIdx.x & 31;
  const int g = lane % LPR;
  const long warp_id = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long n = warp_id * RPW + lane / LPR;
  const bool row_ok = n < N && flags[n] == 0;           // other rows are answered by K1
  float myd = INFINITY;
  int myj = -1;
  if (row_ok && g < TG) {
    const int j = cand[n * TG + g];
    if (j >= 0 && j < M) {
      const float *x = data + n * (long)D;
      const float *c = codes + (long)j * D;
      float acc = 0.0f;
      if ((D & 3) == 0) {
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        const float4 *c4 = reinterpret_cast<const float4 *>(c);
#pragma unroll 4
        for (int i = 0; i < D / 4; i++) {
          const float4 xv = x4[i], cv = __ldg(c4 + i);
          acc = sq_acc(acc, cv.x, xv.x);               // the reference's sum, component order
          acc = sq_acc(acc, cv.y, xv.y);
          acc = sq_acc(acc, cv.z, xv.z);
          acc = sq_acc(acc, cv.w, xv.w);
        }
      } else {
        for (int i = 0; i < D; i++) acc = sq_acc(acc, __ldg(c + i), x[i]);
      }
      myd = acc;
      myj = j;
    }
  }
  // gather the group's results on every lane (only the leader uses them)
  float cd[TG];
  int ci[TG];
  int nc = 0;
#pragma unroll
  for (int t = 0; t < TG; t++) {
    cd[t] = __shfl_sync(0xffffffffu, myd, (lane / LPR) * LPR + t);
    ci[t] = __shfl_sync(0xffffffffu, myj, (lane / LPR) * LPR + t);
    if (ci[t] >= 0) nc++; else cd[t] = INFINITY;
  }
  if (!row_ok || g != 0) return;
  // order by the reference's rule: k == 1 -> (diff asc, idx asc), k >= 2 -> (diff asc, idx desc)
  const bool knn_rule = k > 1;
#pragma unroll
  for (int a = 0; a < TG; a++)
#pragma unroll
    for (int b = a + 1; b < TG; b++) {
      bool sw;
      if (ci[b] < 0) sw = false;
      else if (ci[a] < 0) sw = true;
      else sw = cd[b] < cd[a] || (cd[b] == cd[a] && (knn_rule ? ci[b] > ci[a] : ci[b] < ci[a]));
      if (sw) { float td = cd[a]; cd[a] = cd[b]; cd[b] = td; int ti = ci[a]; ci[a] = ci[b]; ci[b] = ti; }
    }
  // ---- certificate (double arithmetic; any NaN makes it fail)
  const RowStats s = rs[n];
  const CbStats cs = *cst;
  const double nx = s.nx, NM = cs.nm;
  const double amag = 2.0 * nx * NM + (double)cs.nm2;                  // bound on |partial sums|, |score|
  const double e_dot = 2.0 * ((double)s.nxlo * cs.nmlo + (double)s.nrx * NM + nx * (double)cs.nrm);
  const double e_norm = ldexp((double)cs.nm2, -25);
  const double e_acc = 2.0 * (double)(Kp / 16) * 17.0 * ldexp(amag, -23);
  const double e_pack = ldexp(amag, -14);
  const double E = (e_dot + e_norm + e_acc + e_pack) * 1.0001;
  const double Lc = s.nx2 + (double)thr[n] - E;                         // lower bound, centred exact distance
  const double eta = ldexp(nx + NM, -23);                               // centring rounding, both vectors
  bool ok = false;
  double L = 0.0;
  if (Lc > 0.0) {
    const double r = sqrt(Lc) - eta;
    if (r > 0.0) {
      const double gamma = (double)(D + 2) * ldexp(1.0, -24) * 1.01;     // reference's own rounding
      L = r * r * (1.0 - gamma) * (1.0 - 1e-6);
      // the k-th winner must be strictly below every non-candidate; all real candidates needed
      const int need = k < (int)M ? k : (int)M;
      ok = nc >= need && need >= 1 && (double)cd[need - 1] < L && cd[need - 1] < FLT_MAX;
    }
  }
  if (M <= TG && nc == (int)M) ok = true;          // every code is a candidate: nothing to certify
  if (ok && k == 1 && !(cd[0] < FLT_MAX)) ok = false;
  if (!ok) {
    listW[atomicAdd(&counters[0], 1)] = (int)n;
    atomicAdd(&counters[3], 1);
    return;
  }
  atomicAdd(&counters[2], 1);
  for (int t = 0; t < k; t++) {
    const bool have = t < nc && t < TG;
    if (k == 1) { idx[n] = have ? ci[0] : -1; diff[n] = have ? cd[0] : -1.0f; }
    else { idx[n * k + t] = have ? ci[t] : -1; diff[n * k + t] = have ? cd[t] : FLT_MAX; }
  }
  nfound[n] = k;
}

// ---------------------------------------------------------------- group re-rank (k == 1)
// One warp per row: lane l computes the exact distance to code 32*g + l of each candidate
// group g, the warp takes the (diff, index) minimum -- first minimum wins, lvq_pak.c:79 -- and
// lane 0 evaluates the certificate against the (NG+1)-th smallest group minimum.
template <int NG>
__global__ void __launch_bounds__(256)
k2_rerank_group_kernel(const float *__restrict__ data, const float *__restrict__ codes, long N, long M,
                       int D, int Kp, const unsigned char *__restrict__ flags,
                       const RowStats *__restrict__ rs, const CbStats *__restrict__ cst,
                       const int32_t *__restrict__ cand, const float *__restrict__ thr,
                       int *__restrict__ listW, int *__restrict__ counters, int32_t *__restrict__ idx,
                       float *__restrict__ diff, int32_t *__restrict__ nfound) {
  const int lane = threadIdx.x & 31;
  const long n = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  if (n >= N || flags[n] != 0) return;                 // warp-uniform
  const float *x = data + n * (long)D;
  u64 best = ~0ull;
  int ncodes = 0;
#pragma unroll
  for (int gsel = 0; gsel < NG; gsel++) {
    const int gid = cand[n * NG + gsel];
    if (gid < 0) continue;
    const long j = (long)gid * 32 + lane;
    ncodes += (int)min(32L, max(0L, M - (long)gid * 32));
    if (j >= M) continue;
    const float *c = codes + j * D;
    float acc = 0.0f;
    if ((D & 3) == 0) {
      const float4 *x4 = reinterpret_cast<const float4 *>(x);
      const float4 *c4 = reinterpret_cast<const float4 *>(c);
#pragma unroll 4
      for (int i = 0; i < D / 4; i++) {
        const float4 xv = __ldg(x4 + i), cv = __ldg(c4 + i);
        acc = sq_acc(acc, cv.x, xv.x);
        acc = sq_acc(acc, cv.y, xv.y);
        acc = sq_acc(acc, cv.z, xv.z);
        acc = sq_acc(acc, cv.w, xv.w);
      }
    } else {
      for (int i = 0; i < D; i++) acc = sq_acc(acc, __ldg(c + i), __ldg(x + i));
    }
    // only d < FLT_MAX can win; non-negative floats order like their bit patterns
    if (acc < FLT_MAX) {
      const u64 key = ((u64)__float_as_uint(acc) << 32) | (unsigned)j;
      best = key < best ? key : best;
    }
  }
  {
    const unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
    const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    best = ((u64)mh << 32) | ml;
  }
  if (lane != 0) return;
  const float dbest = __uint_as_float((unsigned)(best >> 32));
  const int jbest = (int)(unsigned)best;
  bool ok = false;
  if (best != ~0ull) {
    const RowStats s = rs[n];
    const CbStats cs = *cst;
    const double nx = s.nx, NM = cs.nm;
    const double amag = 2.0 * nx * NM + (double)cs.nm2;
    const double e_dot = 2.0 * ((double)s.nxlo * cs.nmlo + (double)s.nrx * NM + nx * (double)cs.nrm);
    const double e_norm = ldexp((double)cs.nm2, -25);
    const double e_acc = 2.0 * (double)(Kp / 16) * 17.0 * ldexp(amag, -23);
    const double e_pack = ldexp(amag, -19);                 // 3 low mantissa bits carry the group number
    const double E = (e_dot + e_norm + e_acc + e_pack) * 1.0001;
    const double Lc = s.nx2 + (double)thr[n] - E;
    const double eta = ldexp(nx + NM, -23);
    if (Lc > 0.0) {
      const double r = sqrt(Lc) - eta;
      if (r > 0.0) {
        const double gamma = (double)(D + 2) * ldexp(1.0, -24) * 1.01;
        const double L = r * r * (1.0 - gamma) * (1.0 - 1e-6);
        ok = (double)dbest < L;
      }
    }
    if (ncodes >= M) ok = true;                             // every code was re-ranked
  }
  if (!ok) {
    listW[atomicAdd(&counters[0], 1)] = (int)n;
    atomicAdd(&counters[3], 1);
    return;
  }
  atomicAdd(&counters[2], 1);
  idx[n] = jbest;
  diff[n] = dbest;
  nfound[n] = 1;
}

// ---------------------------------------------------------------- host side
struct K2Scratch {      // carved out of one grow-only device buffer
  __nv_bfloat16 *Aimg;
  RowStats *rs;
  int32_t *cand;
  float *thr;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool k2_eligible(int path, long M, int D, long N, int k, unsigned cb_flags) {
  if (path == 1) return false;                    // BMU_PATH_EXACT
  if (cb_flags != 0) return false;                // non-finite / tiny codebook: exact kernels only
  if (k < 1 || k > 16 || D < 1 || D > 8192) return false;
  if (path == 2) return true;                     // BMU_PATH_FILTER: forced
  // AUTO: the filter pays off once the contraction dominates
  return M >= 512 && N >= 4096 && (double)M * D >= 32768.0;
}

void k2_codebook_invalidate(K2Codebook *c) { c->valid = 0; }

void k2_codebook_free(K2Codebook *c) {
  if (c->d_ops) cudaFree(c->d_ops);
  if (c->d_norm) cudaFree(c->d_norm);
  c->d_ops = nullptr;
  c->d_norm = nullptr;
  c->valid = 0;
}

static cudaError_t k2_build_codebook(K2Codebook *c, const K1Args &a, cudaStream_t st) {
  const int Kp = k2_kp(a.D);
  const long nct = (a.M + K2_TN - 1) / K2_TN;
  const size_t need = (size_t)nct * K2_TN * Kp * 2;
  cudaError_t e;
  if (need > c->ops_bytes) {
    if (c->d_ops) cudaFree(c->d_ops);
    c->d_ops = nullptr;
    if ((e = cudaMalloc(&c->d_ops, need)) != cudaSuccess) return e;
    c->ops_bytes = need;
  }
  if (!c->d_norm) {
    // [CbStats (16 B)] [mean: D floats]
    if ((e = cudaMalloc((void **)&c->d_norm, 64 + sizeof(float) * 8192)) != cudaSuccess) return e;
  }
  if ((e = cudaMemsetAsync(c->d_norm, 0, 64, st)) != cudaSuccess) return e;
  float *mean = c->d_norm + 16;
  k2_mean_kernel<<<(a.D + 127) / 128, 128, 0, st>>>(a.codes, a.M, a.D, mean);
  const long warps = nct * K2_TN;
  k2_cb_prep_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
      a.codes, a.M, a.D, mean, (__nv_bfloat16 *)c->d_ops, (CbStats *)c->d_norm);
  k1_count_launch(2);
  c->Kp = Kp;
  c->valid = 1;
  return cudaGetLastError();
}

static cudaEvent_t g_k2ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static bool g_k2ev_valid = false;

template <int TG, int TT, bool GROUP>
static cudaError_t k2_run(K2Codebook *c, const K1Args &a, const K2Scratch &s, cudaStream_t st) {
  const int Kp = c->Kp;
  const bool a_res = Kp <= K2_ARES_MAX_KP;
  const size_t smem = K2Smem::bytes(Kp, a_res);
  cudaError_t e = cudaFuncSetAttribute(k2_gemm_kernel<TG, TT, GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long ntiles = (a.N + K2_TM - 1) / K2_TM;
  const int grid = (int)(ntiles < a.num_sms ? ntiles : a.num_sms);
  k2_gemm_kernel<TG, TT, GROUP><<<grid, K2_THREADS, smem, st>>>(s.Aimg, (const __nv_bfloat16 *)c->d_ops, a.N, a.M, Kp,
                                                        a_res ? 1 : 0, s.cand, s.thr);
  k1_count_launch(1);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[2], st);
  if constexpr (GROUP) {
    k2_rerank_group_kernel<TG><<<(unsigned)((a.N + 7) / 8), 256, 0, st>>>(
        a.data, a.codes, a.N, a.M, a.D, Kp, a.flags, s.rs, (const CbStats *)c->d_norm, s.cand, s.thr,
        a.listW, a.counters, a.idx, a.diff, a.nfound);
  } else {
    constexpr int LPR = TG <= 4 ? 4 : (TG <= 16 ? 16 : 32);
    const long rr_warps = (a.N + (32 / LPR) - 1) / (32 / LPR);
    k2_rerank_kernel<TG, LPR><<<(unsigned)((rr_warps + 7) / 8), 256, 0, st>>>(
        a.data, a.codes, a.N, a.M, a.D, a.k, Kp, a.flags, s.rs, (const CbStats *)c->d_norm, s.cand, s.thr,
        a.listW, a.counters, a.idx, a.diff, a.nfound);
  }
  k1_count_launch(1);
  return cudaGetLastError();
}

cudaError_t k2_last_kernel_ms(float out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  if (!g_k2ev_valid) return cudaSuccess;
  cudaError_t e = cudaEventSynchronize(g_k2ev[4]);
  if (e != cudaSuccess) return e;
  for (int i = 0; i < 4; i++)
    if ((e = cudaEventElapsedTime(&out[i], g_k2ev[i], g_k2ev[i + 1])) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t k2_search(K2Codebook *c, const K1Args &a, void **scratch, size_t *scratch_bytes, cudaStream_t st) {
  cudaError_t e;
  if (!g_k2ev[0])
    for (int i = 0; i < 5; i++)
      if ((e = cudaEventCreate(&g_k2ev[i])) != cudaSuccess) return e;
  if (!c->valid && (e = k2_build_codebook(c, a, st)) != cudaSuccess) return e;
  const int Kp = c->Kp;
  const int TG = a.k == 1 ? 4 : (a.k <= 5 ? 10 : 20);
  const long ntiles = (a.N + K2_TM - 1) / K2_TM;
  // scratch layout
  size_t off = 0;
  const size_t oA = off; off = align_up(off + (size_t)ntiles * K2_TM * Kp * 2, 256);
  const size_t oR = off; off = align_up(off + (size_t)a.N * sizeof(RowStats), 256);
  const size_t oC = off; off = align_up(off + (size_t)a.N * TG * 4, 256);
  const size_t oT = off; off = align_up(off + (size_t)a.N * 4, 256);
  if (off > *scratch_bytes) {
    if (*scratch) cudaFree(*scratch);
    *scratch = nullptr;
    *scratch_bytes = 0;
    if ((e = cudaMalloc(scratch, off)) != cudaSuccess) return e;
    *scratch_bytes = off;
  }
  K2Scratch s;
  unsigned char *base = (unsigned char *)*scratch;
  s.Aimg = (__nv_bfloat16 *)(base + oA);
  s.rs = (RowStats *)(base + oR);
  s.cand = (int32_t *)(base + oC);
  s.thr = (float *)(base + oT);

  if ((e = cudaMemsetAsync(a.counters, 0, 4 * sizeof(int), st)) != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[0], st);
  k2_row_prep_kernel<<<(unsigned)ntiles, 256, 0, st>>>(a.data, a.mask, a.N, a.D, a.k, c->d_norm + 16, s.Aimg,
                                                      s.rs, a.flags, a.listW, a.listS, a.counters, a.idx,
                                                      a.diff, a.nfound);
  k1_count_launch(1);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[1], st);
  // k == 1: group mode for short contractions (the 4-op/score epilogue would outlast the MMAs);
  // long contractions keep the element mode, whose re-rank touches 4 instead of 64 code rows
  if (a.k == 1 && Kp <= K2_GROUP_MAX_KP) e = k2_run<2, 2, true>(c, a, s, st);
  else if (a.k == 1) e = k2_run<4, 2, false>(c, a, s, st);
  else if (a.k <= 5) e = k2_run<10, 4, false>(c, a, s, st);
  else e = k2_run<20, 4, false>(c, a, s, st);
  if (e != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[3], st);
  // rows that failed the certificate + masked / tiny rows, then the non-finite rows
  if ((e = k1_run_lists(a, st)) != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[4], st);
  g_k2ev_valid = true;
  return cudaSuccess;
}

}  // namespace bmu
