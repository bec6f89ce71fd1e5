// k2_filter.cu -- K2: tcgen05 tensor-core GEMM filter + exact FP32 re-rank (sm_100a).
//
// For large M*D the search  argmin_j ||x - m_j||^2  is dominated by the contraction x.m_j.
// K2 computes an APPROXIMATE score  s~_j = ||m'_j||^2 - 2 x'.m'_j  (primes: vectors centred
// on the codebook mean and scaled by a power of two) on the 5th-generation tensor cores,
// keeps a few candidates per row in the GEMM epilogue, and then decides the winner(s) with
// the reference's EXACT FP32 sum (lvq_pak.c:63-73) over those candidates only.  A per-row
// certificate proves that no code outside the candidate set can win or tie; rows whose
// certificate fails are answered by the exact kernel K1 (k1_warp_kernel).  Results are
// therefore bit-identical to K1 / the reference whatever the data looks like; only the speed
// depends on how well fp16 resolves the data.
//
//   operands   fp16, ONE term:  x'.m' ~ xh.mh  (K = Dp + 3 -> Kp); the per-row rounding
//              residual ||x' - xh|| is measured in the prep kernel and enters the error bound E,
//              so E is rigorous for every row.  The -2 factor and ||m'||^2 (3 fp16 terms against
//              columns of ones) are folded into the B operand: the accumulator IS the score.
//   scaling    s = 2^e with s^2 max||m'||^2 in [2^11, 2^13): products stay far from the fp16
//              range limits; rows whose scaled values overflow fp16 go to K1.
//   GEMM       tcgen05.mma.cta_group::1.kind::f16 (K = 16 per instruction), FP32 accumulators in TMEM,
//              operands staged by cp.async.bulk (UBLKCP) from images pre-arranged in the canonical
//              no-swizzle K-major core-matrix layout, mbarrier rings, one elected MMA-issuing lane.
//   k2_rec_kernel  (k == 1, short contractions; C3): 704 threads.  4 row tiles share every staged code
//              tile (L2 -> smem traffic / 4); 128x128 MMAs into four 128-column accumulators, one per
//              row tile; 16 epilogue warps with the "record" epilogue: minima of 4-code groups per 32
//              columns (0.62 ALU op per score) and, only when a row's chunk minimum is below its running
//              threshold best+delta, a short predicated update that records the group of the minimum;
//              4 re-rank warps compute the exact distances of the recorded groups of the previous pass
//              and evaluate the certificate inside the same kernel.
//   k2_gemm_kernel (k >= 2 or long contractions; C4): 128x256 MMAs, K streamed in slabs, two 256-column
//              accumulators; epilogue keeps the TT smallest keys per code tile with the column index packed
//              into the low mantissa bits; k2_rerank_kernel orders the candidates by the reference's k-NN
//              rule and evaluates the certificate, overlapped with the next sub-batch's GEMM.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "k2_filter.h"
#include "api_internal.h"

namespace bmu {

constexpr int K2_TM = 128;   // rows per sample tile (UMMA M)
constexpr int K2_TN = 256;   // codes per code tile (UMMA N)
constexpr int K2_KS = 64;    // K elements per pipeline stage (4 MMAs), streaming kernel
constexpr int K2_NSTAGE = 4;      // slab stages of the default configuration
constexpr int K2_NSTAGE_MAX = 8;  // barrier slots
constexpr int K2_THREADS = 192;   // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer.  (Eight epilogue warps, two per
                                  // row splitting the 256 columns, were measured on C4: GEMM 4.0 -> 4.5 ms (k = 1), 4.95 -> 5.85
                                  // (k = 5), and the re-rank blocks that run beside the GEMM CTA got fewer registers.)
constexpr int K2_ARES_MAX_KP = 320;   // A image stays resident in smem up to this Kp (streaming kernel)

constexpr int K2R_R = 4;          // row tiles per CTA pass of the record kernel = epilogue groups = accumulators
constexpr int K2R_BST = 2;        // staged code tiles (3 measured no faster; 2 leave 60 KB of the SM's shared memory
                                  // to the prep / re-rank blocks that run beside the GEMM CTA)
constexpr int K2R_TNH = 128;      // accumulator width: half a code tile (4 x 128 columns fill TMEM).  Measured and rejected:
                                  // two 64-column accumulators per row tile (the MMAs fill one while the epilogue drains
                                  // the other): 13.5 instead of 11.7 ms, a 128 x 64 x 16 MMA takes about as long as a
                                  // 128 x 128 x 16 one; and an issuer in which one elected lane also does the barrier
                                  // waits while the other lanes park: 12.3 instead of 12.0 ms on the same box.
constexpr int K2R_THREADS = 704;  // warps 0-15 epilogue (group g = warp / 4 owns row tile g), warp 16 producer, warp 17 MMA,
                                  // warps 18-21 exact re-rank of the pass the epilogue finished before
constexpr int K2R_GW = 4;         // candidate granularity: groups of 4 consecutive codes
constexpr int K2R_NG = 2;         // candidate groups kept per row
constexpr int K2R_MAX_KP = 96;    // code tile (256 x Kp fp16) <= 48 KB

struct CbStats {     // maxima over the codebook (centred, scaled), device side
  float nm2_raw;     // max ||m'||^2 before scaling (rounded up)
  float scale;       // s (a power of two)
  float inv_s2;      // 1 / s^2
  float nm;          // max ||s m'||
  float nrm;         // max ||s m' - fp16(s m')||
  float nm2;         // max ||s m'||^2
};

__host__ __device__ inline int k2_dp(int D) { return (D + 7) & ~7; }
__host__ __device__ inline int k2_kp(int D) { return (k2_dp(D) + 3 + 15) & ~15; }

__device__ __forceinline__ void atomic_max_pos(float *addr, float v) {
  atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));   // non-negative floats order as ints
}

// s = 2^e such that s^2 * nm2_raw lies in [2^11, 2^13); 1 for degenerate codebooks
__device__ __forceinline__ float k2_scale_from(float nm2_raw) {
  if (!(nm2_raw > 0.0f) || !(nm2_raw < INFINITY)) return 1.0f;
  int ex;
  frexpf(nm2_raw, &ex);                         // nm2_raw = f * 2^ex, f in [0.5, 1)
  int se = (13 - ex) >> 1;                      // floor((13 - ex) / 2)
  se = max(-60, min(60, se));
  return ldexpf(1.0f, se);
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

// one warp per code: max ||m'||^2 (fixes the scale)
__global__ void __launch_bounds__(256)
k2_cb_norm_kernel(const float *__restrict__ codes, long M, int D, const float *__restrict__ mean,
                  CbStats *__restrict__ st) {
  const int lane = threadIdx.x & 31;
  const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  if (w >= M) return;
  double n2 = 0.0;
  for (int i = lane; i < D; i += 32) {
    const float c = __fsub_rn(codes[w * D + i], mean[i]);
    n2 += (double)c * c;
  }
  for (int off = 16; off >= 1; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
  if (lane == 0) {
    float v = (float)n2 * 1.0001f;
    if (!(v < INFINITY)) v = INFINITY;          // NaN / overflow: degenerate, scale 1
    atomic_max_pos(&st->nm2_raw, v);
  }
}

// one warp per code: centre, scale, round to fp16, write the B image, accumulate the maxima
__global__ void __launch_bounds__(256)
k2_cb_prep_kernel(const float *__restrict__ codes, long M, int D, const float *__restrict__ mean,
                  __half *__restrict__ Bimg, CbStats *__restrict__ st) {
  const int lane = threadIdx.x & 31;
  const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nct = (M + K2_TN - 1) / K2_TN;
  if (w >= nct * K2_TN) return;
  const int Dp = k2_dp(D), Kp = k2_kp(D);
  const float sc = k2_scale_from(st->nm2_raw);
  if (w == 0 && lane == 0) { st->scale = sc; st->inv_s2 = 1.0f / (sc * sc); }   // powers of two: exact
  const long ct = w / K2_TN;
  const int r = (int)(w % K2_TN);
  __half *img = Bimg + ct * (long)K2_TN * Kp;           // [kc][TN][8]
  auto put = [&](int k, float v) { img[((long)(k >> 3) * K2_TN + r) * 8 + (k & 7)] = __float2half_rn(v); };
  // zero everything this row owns first (pads included)
  for (int k = lane; k < Kp; k += 32) put(k, 0.0f);
  __syncwarp();
  if (w >= M) {            // padding code: a large score (never a candidate: column >= M is masked / range-checked)
    if (lane == 0) { put(Dp, 65504.0f); put(Dp + 1, 65504.0f); put(Dp + 2, 65504.0f); }
    return;
  }
  double n2 = 0.0, nr2 = 0.0;
  for (int i = lane; i < D; i += 32) {
    const float c = __fmul_rn(__fsub_rn(codes[w * D + i], mean[i]), sc);
    const float h = __half2float(__float2half_rn(c));
    const float res = __fsub_rn(c, h);           // exact (Sterbenz / fp16 grid is a subset of fp32)
    put(i, -2.0f * h);                           // |h| <= 2^6.5: the factor -2 is exact in fp16
    n2 += (double)c * c;
    nr2 += (double)res * res;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    n2 += __shfl_xor_sync(0xffffffffu, n2, off);
    nr2 += __shfl_xor_sync(0xffffffffu, nr2, off);
  }
  if (lane == 0) {
    const float nf = (float)n2;                  // < 2^13 * 1.0001 by the choice of the scale
    const float a = __half2float(__float2half_rn(nf));
    const float r1 = nf - a;
    const float b = __half2float(__float2half_rn(r1));
    const float r2 = r1 - b;
    put(Dp, a);
    put(Dp + 1, b);
    put(Dp + 2, r2);
    const float up = 1.0001f;
    atomic_max_pos(&st->nm, (float)sqrt(n2) * up);
    atomic_max_pos(&st->nrm, (float)sqrt(nr2) * up);
    atomic_max_pos(&st->nm2, (float)n2 * up);
  }
}

// FP32 codebook regrouped for the re-rank warps of k2_rec_kernel: [group of K2R_GW codes][4-component chunk]
// [code in group][4 comps], so that the lanes of a group read contiguous bytes per float4
__global__ void k2_cb_regroup_kernel(const float *__restrict__ codes, long M, int D, float *__restrict__ grp) {
  const int Dq = (D + 3) / 4;
  const long total = ((M + K2R_GW - 1) / K2R_GW) * (long)Dq * K2R_GW * 4;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const int e = (int)(t & 3), l = (int)((t >> 2) % K2R_GW);
    const long rest = t / (4 * K2R_GW);
    const int c4 = (int)(rest % Dq);
    const long gidx = rest / Dq;
    const long j = gidx * K2R_GW + l;
    const int i = c4 * 4 + e;
    grp[t] = (j < M && i < D) ? codes[j * D + i] : 0.0f;
  }
}

// ---------------------------------------------------------------- row side
// One CTA per tile of 128 rows: classification + work lists (same rules as K1's
// data_prep_kernel), centred+scaled fp16 image [kc][128][8], and the per-row error bound.
struct RowStats {
  double nx2;     // ||s x'||^2 (fp32 centred values, before the fp16 rounding)
  float nx;       // ||s x'||   (rounded up)
  float E;        // bound on |score - (||s x' - s m'||^2 - nx2)| over all codes (scaled units)
  float delta;    // candidate window of the record epilogue: > 2E + rounding slack
  float pad;
};

#define ROW_RANGE 16u       // a scaled component overflows fp16: row answered by K1

__device__ __forceinline__ unsigned k2_classify(float v) {
  unsigned b = __float_as_uint(v) & 0x7fffffffu;
  unsigned f = 0;
  if (b >= 0x7f800000u) f |= ROW_NONFINITE;
  if (b != 0u && b < 0x2b800000u) f |= ROW_TINY;
  return f;
}

__device__ __forceinline__ uint32_t half2_bits(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);      // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t *>(&v);
}

constexpr int K2_PS = 32;     // components per row-prep slab

// pack_bits: low mantissa bits the GEMM epilogue overwrites with the column index (0 or 8)
__global__ void __launch_bounds__(256)
k2_row_prep_kernel(const float *__restrict__ data, const unsigned char *__restrict__ mask, long N, long row0,
                   int D, int k, int pack_bits, const float *__restrict__ mean,
                   const CbStats *__restrict__ cst, __half *__restrict__ Aimg,
                   RowStats *__restrict__ rs, unsigned char *__restrict__ flags,
                   int *__restrict__ listW, int *__restrict__ listS, int *__restrict__ counters,
                   int32_t *__restrict__ idx, float *__restrict__ diff, int32_t *__restrict__ nfound) {
  // staged through shared memory so that both the row reads and the image writes are coalesced
  __shared__ float xs[K2_TM][K2_PS + 1];                      // inputs of the slab
  __shared__ unsigned char ms[K2_TM][K2_PS];                  // mask bytes of the slab
  __shared__ __align__(16) uint4 img_s[K2_PS / 8][K2_TM];     // image chunks of the slab
  __shared__ float red[K2_TM][2];
  __shared__ int redi[K2_TM][2];
  const int Dp = k2_dp(D), Kp = k2_kp(D);
  const long tile = blockIdx.x;
  const long n0 = tile * K2_TM;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (K2_TM - 1), half = tid >> 7;           // thread = (row, 16-component half)
  uint4 *img = reinterpret_cast<uint4 *>(Aimg + tile * (long)K2_TM * Kp);   // [kc][128] uint4
  const long nrow = n0 + row;
  const float sc = cst->scale;
  // ||x'||^2 and the squared fp16 residual, accumulated in FP32: the relative error (D+4) 2^-24 of
  // these sums is charged to the bounds below, which is far cheaper than accumulating in double
  float n2 = 0.0f, nr2 = 0.0f;
  unsigned f = 0;
  int nmasked = 0;
  __shared__ float mean_s[K2_PS];
  const bool fast = mask == nullptr && n0 + K2_TM <= N && (D % K2_PS) == 0 &&   // full tile: no per-element checks,
                    (reinterpret_cast<uintptr_t>(data) & 15) == 0;              // rows readable as float4

  if (fast) {
    for (int d0 = 0; d0 < D; d0 += K2_PS) {
      __syncthreads();
      if (tid < K2_PS) mean_s[tid] = mean[d0 + tid];
      // phase 1: 128 rows x 32 components as float4 loads, all four of a thread in flight before the first
      // use (the kernel is HBM-bound: with one scalar load per lane and row the SM had ~28 KB in flight,
      // below the ~35 KB per SM that 6.5 TB/s at ~0.8 us latency needs; measured 1.15 -> see DESIGN.md)
      {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int f4 = tid + q * 256;                      // float4 index inside the slab: 8 per row
          v[q] = __ldcs(reinterpret_cast<const float4 *>(data + (n0 + (f4 >> 3)) * (long)D + d0) + (f4 & 7));
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int f4 = tid + q * 256;
          float *dst = &xs[f4 >> 3][(f4 & 7) * 4];            // row stride 33 floats: the 32 lanes hit 32 banks
          dst[0] = v[q].x; dst[1] = v[q].y; dst[2] = v[q].z; dst[3] = v[q].w;
        }
      }
      __syncthreads();
      // phase 2: centre, scale, round; each thread packs 2 x 8 components of one row
#pragma unroll
      for (int grp = 0; grp < 2; grp++) {
        uint32_t hw[4];
#pragma unroll
        for (int pq = 0; pq < 8; pq += 2) {
          float hv[2];
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const int il = half * 16 + grp * 8 + pq + q;
            const float v = xs[row][il];
            f |= k2_classify(v);
            const float c = __fmul_rn(__fsub_rn(v, mean_s[il]), sc);
            const float h = __half2float(__float2half_rn(c));
            const float res = __fsub_rn(c, h);
            hv[q] = h;
            n2 = __fadd_rn(n2, __fmul_rn(c, c));
            nr2 = __fadd_rn(nr2, __fmul_rn(res, res));
          }
          if (!(fabsf(hv[0]) < INFINITY) || !(fabsf(hv[1]) < INFINITY)) f |= ROW_RANGE;
          hw[pq >> 1] = half2_bits(hv[0], hv[1]);
        }
        img_s[half * 2 + grp][row] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      }
      __syncthreads();
      // phase 3: (4 chunks x 128) uint4 form one contiguous run of the image
      for (int t = tid; t < (K2_PS / 8) * K2_TM; t += 256)
        img[((long)d0 / 8 + t / K2_TM) * K2_TM + (t % K2_TM)] = img_s[t / K2_TM][t % K2_TM];
    }
  } else
  for (int d0 = 0; d0 < Dp; d0 += K2_PS) {
    __syncthreads();
    // phase 1: coalesced load of 128 rows x 32 components
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
    // phase 2: centre, scale, round; each thread packs 2 x 8 components of one row
#pragma unroll
    for (int grp = 0; grp < 2; grp++) {
      uint32_t hw[4];
#pragma unroll
      for (int p = 0; p < 4; p++) {
        float hv[2];
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int il = half * 16 + grp * 8 + p * 2 + q;
          const int i = d0 + il;
          hv[q] = 0.0f;
          if (i < D && nrow < N) {
            if (ms[row][il]) { nmasked++; }
            else {
              const float v = xs[row][il];
              f |= k2_classify(v);
              const float c = __fmul_rn(__fsub_rn(v, mean[i]), sc);
              const float h = __half2float(__float2half_rn(c));
              if (!(fabsf(h) < INFINITY)) f |= ROW_RANGE;
              const float res = __fsub_rn(c, h);
              hv[q] = h;
              n2 = __fadd_rn(n2, __fmul_rn(c, c));
              nr2 = __fadd_rn(nr2, __fmul_rn(res, res));
            }
          }
        }
        hw[p] = half2_bits(hv[0], hv[1]);
      }
      img_s[half * 2 + grp][row] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    }
    __syncthreads();
    // phase 3: (chunks x 128) uint4 form one contiguous run of the image
    const int nch = min(K2_PS, Dp - d0) / 8;
    for (int t = tid; t < nch * K2_TM; t += 256) {
      const int ch = t / K2_TM, r = t % K2_TM;
      img[((long)d0 / 8 + ch) * K2_TM + r] = img_s[ch][r];
    }
  }
  // tail: the columns of ones that pick up ||m'||^2, then zero padding up to Kp
  for (int t = tid; t < (Kp - Dp) / 8 * K2_TM; t += 256) {
    const int ch = t / K2_TM, r = t % K2_TM;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ch == 0 && n0 + r < N) { v.x = half2_bits(1.0f, 1.0f); v.y = half2_bits(1.0f, 0.0f); }
    img[((long)Dp / 8 + ch) * K2_TM + r] = v;
  }
  // combine the two halves of every row
  __syncthreads();
  if (half == 1) { red[row][0] = n2; red[row][1] = nr2; redi[row][0] = (int)f; redi[row][1] = nmasked; }
  __syncthreads();
  if (half == 0 && nrow < N) {
    n2 += red[row][0]; nr2 += red[row][1];
    f |= (unsigned)redi[row][0]; nmasked += redi[row][1];
    if (nmasked > 0) f |= ROW_MASKED;
    if (nmasked == D) f |= ROW_ALLMASKED;
    const long n = nrow;
    flags[n] = (unsigned char)f;
    // ---- error bound of the tensor-core score for this row (double; scaled units)
    const CbStats cs = *cst;
    const double up = 1.0001;
    const double facc = (double)(D + 4) * ldexp(1.0, -23);               // FP32 accumulation of n2 / nr2 (x2 margin)
    const double nx = sqrt((double)n2 * (1.0 + facc)) * up, nrx = sqrt((double)nr2 * (1.0 + facc)) * up;
    const double NM = cs.nm, nrm = cs.nrm, nm2 = cs.nm2;
    const double nxh = nx + nrx, NMh = NM + nrm;
    const double amag = 2.0 * nxh * NMh + nm2;                         // bound on |partial sums|, |score|
    const double e_dot = 2.0 * (nrx * NM + nxh * nrm);                 // fp16 rounding of both operands
    const double e_norm = ldexp(nm2, -23) + ldexp(1.0, -22);           // ||m'||^2 as three fp16 terms
    const double e_acc = 2.0 * (double)(Kp / 16) * 17.0 * ldexp(amag, -23);   // FP32 accumulation in the tensor core
    const double e_pack = pack_bits ? ldexp(amag, pack_bits - 22) : 0.0;      // index bits in the mantissa
    const double E = (e_dot + e_norm + e_acc + e_pack) * up;
    // slack of the certificate's other roundings: centring (eta), the reference's own sum (gamma)
    const double dmax = (nx + NM) * (nx + NM);
    const double eta = ldexp(nx + NM, -23);
    const double gamma = (double)(D + 2) * ldexp(1.0, -24) * 1.01 + 1e-6;
    // nx2 enters the certificate as a LOWER bound of ||x'||^2; what that gives away goes into the window
    const double nx2_lo = (double)n2 * (1.0 - facc);
    const double slack = 4.0 * (2.0 * eta * (nx + NM) + gamma * dmax) + 2.0 * facc * (double)n2;
    RowStats s;
    s.nx2 = nx2_lo;
    s.nx = (float)nx * 1.0001f;
    s.E = (float)E * 1.0001f;
    s.delta = (float)(2.02 * E + slack) * 1.0001f;
    s.pad = 0.0f;
    rs[n] = s;
    if (f & ROW_ALLMASKED) {
      nfound[n] = 0;
      for (int t = 0; t < k; t++) { idx[n * k + t] = -1; diff[n * k + t] = (k == 1) ? -1.0f : FLT_MAX; }
    } else if (f & ROW_NONFINITE) {
      listS[atomicAdd(&counters[1], 1)] = (int)(row0 + n);      // list entries are rows of the whole call
    } else if (f & (ROW_TINY | ROW_MASKED | ROW_RANGE)) {
      listW[atomicAdd(&counters[0], 1)] = (int)(row0 + n);
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
struct K2Smem {
  static size_t bytes(int Kp, bool a_res, int nst, int ks = K2_KS) {
    size_t stage = (size_t)K2_TN * ks * 2 + (a_res ? 0 : (size_t)K2_TM * ks * 2);
    return (a_res ? (size_t)K2_TM * Kp * 2 : 0) + nst * stage + 256;
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

// instruction descriptor: D=F32, A=B=F16 (format 0), both K-major, N=256, M=128
__device__ __forceinline__ uint32_t k2_idesc() {
  return (1u << 4) | ((uint32_t)(K2_TN >> 3) << 17) | ((uint32_t)(K2_TM >> 4) << 24);
}

// ---------------------------------------------------------------- streaming kernel (k >= 2 / long K)
// TG: candidates kept per row, TT: smallest keys tracked per code tile.
template <int TG, int TT>
__global__ void __launch_bounds__(K2_THREADS, 1)
k2_gemm_kernel(const __half *__restrict__ Aimg, const __half *__restrict__ Bimg,
               const RowStats *__restrict__ rs, long N, long M, int Kp, int a_res, int nst, int ks, int k,
               int32_t *__restrict__ cand, float *__restrict__ thr) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const size_t a_res_bytes = a_res ? (size_t)K2_TM * Kp * 2 : 0;
  const size_t b_stage_bytes = (size_t)K2_TN * ks * 2;
  const size_t a_stage_bytes = a_res ? 0 : (size_t)K2_TM * ks * 2;
  const size_t stage_bytes = b_stage_bytes + a_stage_bytes;
  unsigned char *sAres = smem;
  unsigned char *sStage = smem + a_res_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStage + nst * stage_bytes);   // nst <= K2_NSTAGE slab stages
  uint64_t *full = bars, *empty = bars + K2_NSTAGE_MAX;
  uint64_t *tfull = bars + 2 * K2_NSTAGE_MAX, *tempty = tfull + 2;
  uint64_t *afull = tempty + 2, *aempty = afull + 1;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(aempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long ntiles = (N + K2_TM - 1) / K2_TM;
  const int nct = (int)((M + K2_TN - 1) / K2_TN);
  const int nslab = (Kp + ks - 1) / ks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < K2_NSTAGE_MAX; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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
          mbar_arrive_expect_tx(afull, (uint32_t)a_res_bytes);
          bulk_g2s(sAres, gA, (uint32_t)a_res_bytes, afull);
        }
        for (int ct = 0; ct < nct; ct++) {
          const unsigned char *gB = reinterpret_cast<const unsigned char *>(Bimg) + (size_t)ct * K2_TN * Kp * 2;
          for (int sl = 0; sl < nslab; sl++, seq++) {
            const int st = seq % nst;
            const int kc = min(ks, Kp - sl * ks);               // K elements in this slab
            mbar_wait(&empty[st], ((seq / nst) & 1) ^ 1);
            const uint32_t bB = (uint32_t)K2_TN * kc * 2, bA = a_res ? 0u : (uint32_t)K2_TM * kc * 2;
            mbar_arrive_expect_tx(&full[st], bB + bA);
            unsigned char *dst = sStage + (size_t)st * stage_bytes;
            bulk_g2s(dst, gB + (size_t)sl * ks * K2_TN * 2, bB, &full[st]);
            if (!a_res) bulk_g2s(dst + b_stage_bytes, gA + (size_t)sl * ks * K2_TM * 2, bA, &full[st]);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    // As in k2_rec_kernel: the whole warp runs the loop (uniform control flow, barrier waits by every lane), one
    // elected lane issues, and the descriptors are built ONCE and advanced by constant offsets.  Building two
    // descriptors per MMA inside a `lane == 0` branch cost ~140 cycles per MMA there, more than the 128 cycles a
    // 128x256x16 MMA takes -- the issuing thread, not the tensor pipe, was the bound of this kernel too.
    const uint32_t idesc = k2_idesc();
    const bool leader = elect_one();
    const uint64_t dA0 = umma_desc(smem_u32(sAres), K2_TM * 16, 128);            // resident A (a_res)
    const uint64_t dS0 = umma_desc(smem_u32(sStage), K2_TN * 16, 128);           // B of stage 0
    const uint64_t dSA0 = umma_desc(smem_u32(sStage) + (uint32_t)b_stage_bytes, K2_TM * 16, 128);   // streamed A of stage 0
    const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
    const uint64_t a_slab_step = (uint64_t)((ks * K2_TM * 2) >> 4);
    constexpr uint64_t A_KK = (uint64_t)((2 * K2_TM * 16) >> 4), B_KK = (uint64_t)((2 * K2_TN * 16) >> 4);
    unsigned seq = 0, acc_seq = 0, tcount = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tcount++) {
      if (a_res) { mbar_wait(afull, tcount & 1); tc_fence_after(); }
      for (int ct = 0; ct < nct; ct++, acc_seq++) {
        const int buf = acc_seq & 1;
        mbar_wait(&tempty[buf], ((acc_seq >> 1) & 1) ^ 1);        // epilogue drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * K2_TN;
        for (int sl = 0; sl < nslab; sl++, seq++) {
          const int st = seq % nst;
          const int nk = min(ks, Kp - sl * ks) >> 4;               // MMAs in this slab
          mbar_wait(&full[st], (seq / nst) & 1);
          tc_fence_after();
          if (leader) {
            const uint64_t db = dS0 + (uint64_t)st * stage_step;
            const uint64_t da = a_res ? dA0 + (uint64_t)sl * a_slab_step : dSA0 + (uint64_t)st * stage_step;
#pragma unroll 4
            for (int kk = 0; kk < nk; kk++)
              umma_f16(d_tmem, da + (uint64_t)kk * A_KK, db + (uint64_t)kk * B_KK, idesc, (sl | kk) ? 1u : 0u);
            umma_commit(&empty[st]);                // smem slot reusable once these MMAs retire
          }
          __syncwarp();
        }
        if (leader) umma_commit(&tfull[buf]);       // accumulator ready for the epilogue
        __syncwarp();
      }
      if (a_res && leader) umma_commit(aempty);
      __syncwarp();
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
          } else if constexpr (TT == 4) {
            // pairs merged into the sorted four by rank selection (the k-th smallest of two sorted lists is
            // min over i of max(b[i-1], p[k-i])): 2 LOP3 + 2 (sort the pair) + 1 + 2 + 3 + 3 = 6.5 ALU ops per
            // score instead of the 9 of two insertions, which made the k = 5 epilogue slower than its MMAs
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const float k0 = __uint_as_float((v[c] & 0xFFFFFF00u) | (uint32_t)(c0 + c));
              const float k1 = __uint_as_float((v[c + 1] & 0xFFFFFF00u) | (uint32_t)(c0 + c + 1));
              const float lo = fminf(k0, k1), hi = fmaxf(k0, k1);
              const float r0 = fminf(b[0], lo);
              const float r1 = fminf(fminf(hi, fmaxf(b[0], lo)), b[1]);
              const float r2 = fminf(fminf(fmaxf(b[0], hi), fmaxf(b[1], lo)), b[2]);
              const float r3 = fminf(fminf(fmaxf(b[1], hi), fmaxf(b[2], lo)), b[3]);
              b[0] = r0; b[1] = r1; b[2] = r2; b[3] = r3;
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
        // TT == 3, 4: a second, independent chain (the upper 32 columns of every 64): one warp per
        // SM partition runs this epilogue, and a single chain of dependent min/max leaves its issue slots idle
        float b2[TT];
#pragma unroll
        for (int t = 0; t < TT; t++) b2[t] = INFINITY;
        if constexpr (TT == 4 || TT == 3) {
          tmem_ld32_nowait(tbase, va);
          tmem_ld32_nowait(tbase + 32, vb);
#pragma unroll 1
          for (int c0 = 0; c0 < K2_TN; c0 += 64) {
            tmem_ld_wait32(va);
            tmem_ld_wait32(vb);
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const float k0 = __uint_as_float((va[c] & 0xFFFFFF00u) | (uint32_t)(c0 + c));
              const float k1 = __uint_as_float((va[c + 1] & 0xFFFFFF00u) | (uint32_t)(c0 + c + 1));
              const float q0 = __uint_as_float((vb[c] & 0xFFFFFF00u) | (uint32_t)(c0 + 32 + c));
              const float q1 = __uint_as_float((vb[c + 1] & 0xFFFFFF00u) | (uint32_t)(c0 + 32 + c + 1));
              const float lo = fminf(k0, k1), hi = fmaxf(k0, k1), lo2 = fminf(q0, q1), hi2 = fmaxf(q0, q1);
              const float r0 = fminf(b[0], lo), s0 = fminf(b2[0], lo2);
              const float r1 = fminf(fminf(hi, fmaxf(b[0], lo)), b[1]), s1 = fminf(fminf(hi2, fmaxf(b2[0], lo2)), b2[1]);
              const float r2 = fminf(fminf(fmaxf(b[0], hi), fmaxf(b[1], lo)), b[2]);
              const float s2 = fminf(fminf(fmaxf(b2[0], hi2), fmaxf(b2[1], lo2)), b2[2]);
              if constexpr (TT == 4) {
                const float r3 = fminf(fminf(fmaxf(b[1], hi), fmaxf(b[2], lo)), b[3]);
                const float s3 = fminf(fminf(fmaxf(b2[1], hi2), fmaxf(b2[2], lo2)), b2[3]);
                b[3] = r3; b2[3] = s3;
              }
              b[0] = r0; b[1] = r1; b[2] = r2;
              b2[0] = s0; b2[1] = s1; b2[2] = s2;
            }
            if (c0 + 64 < K2_TN) {
              tmem_ld32_nowait(tbase + c0 + 64, va);
              tmem_ld32_nowait(tbase + c0 + 96, vb);
            }
          }
        } else {
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
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        // everything of this tile that is not kept is >= the TT-th smallest key of its chain
        tmin = fminf(tmin, fminf(b[TT - 1], b2[TT - 1]));
#pragma unroll
        for (int t = 0; t < ((TT == 4 || TT == 3) ? 2 * TT : TT); t++) {
          float key = t < TT ? b[t < TT ? t : 0] : b2[t >= TT ? t - TT : 0];
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
        // only keys within delta (> 2E) of the k-th smallest can still win or tie: the others
        // become non-candidates bounded by `bound`, which spares the re-rank their code rows
        const int kk = k < TG ? k : TG;
        float kth = gk[0];
#pragma unroll
        for (int t = 1; t < TG; t++) kth = (t < kk) ? gk[t] : kth;
        const float bound = __fadd_ru(kth, rs[n].delta);
#pragma unroll
        for (int t = 0; t < TG; t++) cand[n * TG + t] = (gk[t] < bound) ? gi[t] : -1;
        // candidates dropped from the list are >= its last key; kept ones are re-ranked exactly
        thr[n] = fminf(fminf(tmin, gk[TG - 1]), bound);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ---------------------------------------------------------------- record kernel (k == 1, short K)
// per-row state of the record epilogue: running best score, threshold best + delta, the two
// smallest group minima seen so far with the first code of their group, and `lost` = lower
// bound of every group minimum that was looked at but is not (or no longer) one of the two
struct K2RRow {
  float thr, lost, k0, k1;      // thr = k0 + delta (rounded up): k0 is also the running minimum
  int i0, i1;
};


// N = 128 variant of the instruction descriptor
__device__ __forceinline__ uint32_t k2r_idesc() {
  return (1u << 4) | ((uint32_t)(K2R_TNH >> 3) << 17) | ((uint32_t)(K2_TM >> 4) << 24);
}

// Record kernel.  CTA = 4 row tiles (A resident) x all code tiles (staged whole, K2R_BST-deep ring).
// MMA order per code tile: (half h, row tile r) -> accumulator r (128 TMEM columns each), so an
// accumulator is rewritten every 4th MMA group and its epilogue group has 3 MMA groups of time.
// Epilogue per 32 columns of a row: minima of the eight 4-column groups and of the chunk
// (20 FMNMX3/FMNMX, 0.63 ALU-pipe op per score).  Only if some lane's chunk minimum is below its
// running threshold thr = k0 + delta (vote) does the warp run the short predicated update (41
// instructions) that records the GROUP holding the minimum; the exact distances of the <= 2 x 4 codes
// of the recorded groups are computed by the re-rank warps (k2r_rerank_row).  Invariant for the
// certificate: a group that is not recorded has a minimum >= min(final k0 + delta, lost).
// Exact re-rank of ONE row inside the record kernel (one thread): the reference's sum (lvq_pak.c:63-73)
// over the <= 2 x K2R_GW codes of the recorded groups, first-minimum rule (lvq_pak.c:79), then the
// certificate.  Returns false when the row has to be answered by the exact kernel K1.
struct K2RRerankArgs {
  const float *data, *grp;
  const unsigned char *flags;
  const RowStats *rs;
  const CbStats *cst;
  int *listW, *counters;
  int32_t *idx, *nfound;
  float *diff;
  long N, M;
  int D;
};
__device__ __forceinline__ void k2r_rerank_row(const K2RRerankArgs &A, long n, int g0, int g1, float thr) {
  if (n >= A.N || A.flags[n] != 0) return;               // flagged rows are answered by K1 (lists)
  const int D = A.D, Dq = (D + 3) / 4;
  const float *x = A.data + n * (long)D;
  float dbest = INFINITY;
  int jbest = -1;
#pragma unroll 1
  for (int gsel = 0; gsel < K2R_NG; gsel++) {
    const int gfirst = gsel ? g1 : g0;
    if (gfirst < 0) continue;
    const float4 *c4 = reinterpret_cast<const float4 *>(A.grp) + ((long)(gfirst / K2R_GW) * Dq) * K2R_GW;
    float acc[K2R_GW];
#pragma unroll
    for (int l = 0; l < K2R_GW; l++) acc[l] = 0.0f;
    if ((D & 3) == 0) {
      const float4 *x4 = reinterpret_cast<const float4 *>(x);
#pragma unroll 2
      for (int i = 0; i < Dq; i++) {
        const float4 xv = __ldg(x4 + i);
#pragma unroll
        for (int l = 0; l < K2R_GW; l++) {
          const float4 cv = __ldg(c4 + i * K2R_GW + l);
          acc[l] = sq_acc(acc[l], cv.x, xv.x);           // component order, one rounding per operation
          acc[l] = sq_acc(acc[l], cv.y, xv.y);
          acc[l] = sq_acc(acc[l], cv.z, xv.z);
          acc[l] = sq_acc(acc[l], cv.w, xv.w);
        }
      }
    } else {
      for (int i = 0; i < Dq; i++) {
#pragma unroll
        for (int l = 0; l < K2R_GW; l++) {
          const float4 cv = __ldg(c4 + i * K2R_GW + l);
          const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
          for (int e = 0; e < 4; e++)
            if (4 * i + e < D) acc[l] = sq_acc(acc[l], cc[e], __ldg(x + 4 * i + e));
        }
      }
    }
#pragma unroll
    for (int l = 0; l < K2R_GW; l++) {
      const int j = gfirst + l;
      // only d < FLT_MAX can win; equal distances: the lower index
      if (j < A.M && acc[l] < FLT_MAX && (acc[l] < dbest || (acc[l] == dbest && j < jbest))) { dbest = acc[l]; jbest = j; }
    }
  }
  bool ok = false;
  if (jbest >= 0) {
    // in scaled units every code outside the candidate groups has ||s x' - s m'||^2 >= nx2 + thr - E
    const RowStats s = A.rs[n];
    const CbStats cs = *A.cst;
    const double Lc = s.nx2 + (double)thr - (double)s.E;
    const double eta = ldexp((double)s.nx + (double)cs.nm, -23);
    if (Lc > 0.0) {
      const double r = sqrt(Lc) - eta;
      if (r > 0.0) {
        const double gamma = (double)(D + 2) * ldexp(1.0, -24) * 1.01;
        const double L = r * r * (1.0 - gamma) * (1.0 - 1e-6) * (double)cs.inv_s2;
        ok = (double)dbest < L;
      }
    }
  }
  if (!ok) {
    A.listW[atomicAdd(&A.counters[0], 1)] = (int)n;
    atomicAdd(&A.counters[3], 1);
    return;
  }
  atomicAdd(&A.counters[2], 1);
  A.idx[n] = jbest;
  A.diff[n] = dbest;
  A.nfound[n] = 1;
}

template <int R, int NK>
__global__ void __launch_bounds__(K2R_THREADS, 1)
k2_rec_kernel(const __half *__restrict__ Aimg, const __half *__restrict__ Bimg,
              const RowStats *__restrict__ rs, long N, long M, int Kp, int nst, const K2RRerankArgs RA) {
  static_assert(R == 4, "one epilogue group and one 128-column accumulator per row tile");
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t a_tile_bytes = (uint32_t)K2_TM * Kp * 2, b_tile_bytes = (uint32_t)K2_TN * Kp * 2;
  unsigned char *sA = smem;                                   // R row tiles
  unsigned char *sB = smem + (size_t)R * a_tile_bytes;        // nst (<= K2R_BST) code tiles
  uint64_t *bars = reinterpret_cast<uint64_t *>(sB + (size_t)nst * b_tile_bytes);
  uint64_t *full = bars, *empty = bars + K2R_BST;
  uint64_t *tfull = bars + 2 * K2R_BST, *tempty = tfull + R;
  uint64_t *afull = tempty + R, *aempty = afull + 1;
  uint64_t *rfull = aempty + 1, *rempty = rfull + 2;               // hand-off to the re-rank warps, double buffered
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(rempty + 2);
  int4 *hand = reinterpret_cast<int4 *>(reinterpret_cast<unsigned char *>(bars) + 256);   // [2][R * 128] {g0, g1, thr, -}

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long ntiles = (N + K2_TM - 1) / K2_TM;
  const long nsuper = (ntiles + R - 1) / R;
  const int nct = (int)((M + K2_TN - 1) / K2_TN);
  // 128-column accumulations per row tile: the upper half of the last code tile is skipped when it
  // holds padding only (M = 10000: 79 instead of 80)
  const int nacc = (int)((M + K2R_TNH - 1) / K2R_TNH);

  if (threadIdx.x == 0) {
    for (int s = 0; s < K2R_BST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < R; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    mbar_init(afull, 1);
    mbar_init(aempty, 1);
    for (int b = 0; b < 2; b++) { mbar_init(&rfull[b], 16); mbar_init(&rempty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 17) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_ptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 16) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      unsigned bseq = 0, tcount = 0;
      for (long st = blockIdx.x; st < nsuper; st += gridDim.x, tcount++) {
        // the image is allocated for nsuper*R tiles; tiles past the last row hold don't-care bits
        const unsigned char *gA = reinterpret_cast<const unsigned char *>(Aimg) + (size_t)st * R * a_tile_bytes;
        mbar_wait(aempty, (tcount & 1) ^ 1);             // previous pass's MMAs have consumed A
        mbar_arrive_expect_tx(afull, (uint32_t)R * a_tile_bytes);
#pragma unroll
        for (int r = 0; r < R; r++) bulk_g2s(sA + (size_t)r * a_tile_bytes, gA + (size_t)r * a_tile_bytes, a_tile_bytes, afull);
        for (int ct = 0; ct < nct; ct++, bseq++) {
          const int s = bseq % nst;
          mbar_wait(&empty[s], ((bseq / nst) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], b_tile_bytes);
          bulk_g2s(sB + (size_t)s * b_tile_bytes,
                   reinterpret_cast<const unsigned char *>(Bimg) + (size_t)ct * b_tile_bytes, b_tile_bytes, &full[s]);
        }
      }
    }
  } else if (warp == 17) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (uniform control flow, barrier waits by every lane); one
    // elected lane issues.  Descriptors are built once and advanced by constant offsets: the
    // issue loop has to stay well under the 64 cycles that one 128x128x16 MMA takes.
    const uint32_t idesc = k2r_idesc();
    const bool leader = elect_one();
    const uint64_t da0 = umma_desc(smem_u32(sA), K2_TM * 16, 128);
    const uint64_t db0 = umma_desc(smem_u32(sB), K2_TN * 16, 128);
    const uint64_t a_tile_step = (uint64_t)(a_tile_bytes >> 4), b_tile_step = (uint64_t)(b_tile_bytes >> 4);
    unsigned bseq = 0, tcount = 0, use = 0;              // `use`: accumulations issued per accumulator so far
    for (long st = blockIdx.x; st < nsuper; st += gridDim.x, tcount++) {
      mbar_wait(afull, tcount & 1);
      tc_fence_after();
      for (int ct = 0; ct < nct; ct++, bseq++) {
        const int s = bseq % nst;
        mbar_wait(&full[s], (bseq / nst) & 1);
        tc_fence_after();
        const uint64_t dbs = db0 + (uint64_t)s * b_tile_step;
        const int nh = (2 * ct + 1 < nacc) ? 2 : 1;
#pragma unroll 1
        for (int h = 0; h < nh; h++, use++) {
#pragma unroll
          for (int r = 0; r < R; r++) {
            mbar_wait(&tempty[r], (use & 1) ^ 1);          // epilogue group r drained its accumulator
            tc_fence_after();
            if (leader) {
              const uint64_t da = da0 + (uint64_t)r * a_tile_step;
              const uint64_t db = dbs + (uint64_t)(h * ((K2R_TNH * 16) >> 4));     // codes 128h.. of every K chunk
              const uint32_t d_tmem = tmem_base + r * K2R_TNH;
#pragma unroll
              for (int kk = 0; kk < NK; kk++)
                umma_f16(d_tmem, da + (uint64_t)(kk * ((2 * K2_TM * 16) >> 4)),
                         db + (uint64_t)(kk * ((2 * K2_TN * 16) >> 4)), idesc, kk ? 1u : 0u);
              umma_commit(&tfull[r]);
            }
            __syncwarp();
          }
        }
        if (leader) umma_commit(&empty[s]);               // code tile slot reusable once all products retire
        __syncwarp();
      }
      if (leader) umma_commit(aempty);
      __syncwarp();
    }
  } else if (warp >= 18) {
    // ===================== exact re-rank of the pass the epilogue handed over =====================
    // 128 threads, 4 rows each (one per row tile of the pass).  FMA-pipe work (sub, mul, add) on rows
    // fetched through L2, beside an epilogue that is bound by the ALU pipe: it costs the GEMM almost
    // nothing and replaces a separate kernel that re-read candidates, thresholds and row statistics.
    const int t = threadIdx.x - 18 * 32;                  // 0 .. 127
    unsigned cnt = 0;
    for (long st = blockIdx.x; st < nsuper; st += gridDim.x, cnt++) {
      const int par = cnt & 1;
      mbar_wait(&rfull[par], (cnt >> 1) & 1);
#pragma unroll 1
      for (int r = 0; r < R; r++) {
        const int4 h = hand[par * (R * K2_TM) + r * K2_TM + t];
        k2r_rerank_row(RA, (st * R + r) * K2_TM + t, h.x, h.y, __int_as_float(h.z));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rempty[par]);
    }
  } else {
    // ===================== epilogue: group g = warp / 4 owns row tile g and accumulator g =====
    const int g = warp >> 2, quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + g * K2R_TNH;
    unsigned use = 0, cnt = 0;
    for (long st = blockIdx.x; st < nsuper; st += gridDim.x, cnt++) {
      const long n = (st * R + g) * K2_TM + row;
      const float delta = n < N ? rs[n].delta : 0.0f;
      K2RRow r = {INFINITY, INFINITY, INFINITY, INFINITY, -1, -1};
      for (int q = 0; q < nacc; q++, use++) {               // q = 2 * code tile + half
        mbar_wait(&tfull[g], use & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32_nowait(tbase, v);
#pragma unroll 1
        for (int c0 = 0; c0 < K2R_TNH; c0 += 32) {
          tmem_ld_wait32(v);
          // minima of the eight 4-column groups (FMNMX3 + FMNMX), then of the chunk: 20 ALU ops
          float gm[8];
#pragma unroll
          for (int t = 0; t < 8; t++)
            gm[t] = fminf(fminf(fminf(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1])), __uint_as_float(v[4 * t + 2])),
                          __uint_as_float(v[4 * t + 3]));
          if (c0 + 32 < K2R_TNH) {
            // the scores are dead now: the next chunk streams into the same registers
            tmem_ld32_nowait(tbase + c0 + 32, v);
          } else {
            // the last 32 columns have been consumed: the accumulator can be refilled while the rest of this
            // chunk is processed
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[g]);
          }
          const float a3 = fminf(fminf(gm[0], gm[1]), gm[2]), b3 = fminf(fminf(gm[3], gm[4]), gm[5]), c2 = fminf(gm[6], gm[7]);
          const float m = fminf(fminf(a3, b3), c2);
          const bool ins = m < r.thr;
          if (__any_sync(0xffffffffu, ins)) {
            // ---- slow path: warp-uniform entry, straight-line predicated code
            // First group holding the minimum, and the smallest minimum of the other seven, found along
            // the reduction tree (which triple, then which member: 23 ALU-pipe instructions instead of the
            // 33 of a linear search over the eight groups).  Ties resolve to the lower group; the tied
            // value then shows up in s2, which only makes the certificate more conservative.
            const bool inA = a3 == m, inB = !inA && b3 == m;
            const float t0 = inA ? gm[0] : (inB ? gm[3] : gm[6]);
            const float t1 = inA ? gm[1] : (inB ? gm[4] : gm[7]);
            const float t2 = inA ? gm[2] : (inB ? gm[5] : INFINITY);
            const bool e0 = t0 == m, e1 = !e0 && t1 == m;
            const int qs = (inA ? 0 : (inB ? 3 : 6)) + (e0 ? 0 : (e1 ? 1 : 2));
            const float s2 = fminf(fminf(fminf(inA ? b3 : a3, (inA || inB) ? c2 : b3), e0 ? t1 : t0), (e0 || e1) ? t2 : t1);
            const int gi = q * K2R_TNH + c0 + K2R_GW * qs;     // first code of that group
            const bool first = m < r.k0, second = ins && !first && m < r.k1;     // m < k0 <= thr implies ins
            // the minimum that leaves the pair (or m itself when it does not enter) bounds what is dropped
            r.lost = fminf(r.lost, (first || second) ? r.k1 : (ins ? m : INFINITY));
            r.k1 = first ? r.k0 : (second ? m : r.k1);
            r.i1 = first ? r.i0 : (second ? gi : r.i1);
            r.k0 = first ? m : r.k0;
            r.i0 = first ? gi : r.i0;
            r.thr = first ? __fadd_ru(m, delta) : r.thr;           // a new running minimum tightens the threshold
            // only one group of this chunk is recorded; the others are >= s2
            r.lost = (ins && s2 < r.thr) ? fminf(r.lost, s2) : r.lost;
          }
        }
      }
      {
        // hand the row over to the re-rank warps (the buffer of two passes ago must have been drained)
        const int par = cnt & 1;
        mbar_wait(&rempty[par], ((cnt >> 1) & 1) ^ 1);
        const float bound = r.thr;                      // final best + delta (rounded up)
        int4 h;
        h.x = (r.k0 < bound && r.i0 < M) ? r.i0 : -1;
        h.y = (r.k1 < bound && r.i1 < M) ? r.i1 : -1;
        // groups never recorded: minimum >= best + delta; recorded and dropped: >= lost
        h.z = __float_as_int(fminf(bound, r.lost));
        h.w = 0;
        hand[par * (R * K2_TM) + g * K2_TM + row] = h;
        __syncwarp();
        if (lane == 0) mbar_arrive(&rfull[par]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ---------------------------------------------------------------- exact re-rank + certificate
// LPR lanes per row (power of two >= TG): lane g of a row's group computes the exact distance
// of candidate g, so the loads of x are shared by the group and every lane streams one code
// row; the group leader then orders the candidates and evaluates the certificate.
template <int TG, int LPR>
__global__ void __launch_bounds__(256)
k2_rerank_kernel(const float *__restrict__ data, const float *__restrict__ codes, long N, long M, long row0, int D,
                 int k, const unsigned char *__restrict__ flags, const RowStats *__restrict__ rs,
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
  // ---- certificate (double arithmetic; any NaN makes it fail).  In scaled units every code
  // that is not a candidate has  ||s x' - s m'||^2 >= nx2 + thr - E =: Lc.
  const RowStats s = rs[n];
  const CbStats cs = *cst;
  const double Lc = s.nx2 + (double)thr[n] - (double)s.E;
  const double eta = ldexp((double)s.nx + (double)cs.nm, -23);         // centring rounding, both vectors
  bool ok = false;
  if (Lc > 0.0) {
    const double r = sqrt(Lc) - eta;
    if (r > 0.0) {
      const double gamma = (double)(D + 2) * ldexp(1.0, -24) * 1.01;     // reference's own rounding
      const double L = r * r * (1.0 - gamma) * (1.0 - 1e-6) * (double)cs.inv_s2;
      // the k-th winner must be strictly below every non-candidate; all real candidates needed
      const int need = k < (int)M ? k : (int)M;
      ok = nc >= need && need >= 1 && (double)cd[need - 1] < L && cd[need - 1] < FLT_MAX;
    }
  }
  if (M <= TG && nc == (int)M) ok = true;          // every code is a candidate: nothing to certify
  if (ok && k == 1 && !(cd[0] < FLT_MAX)) ok = false;
  if (!ok) {
    listW[atomicAdd(&counters[0], 1)] = (int)(row0 + n);     // the other arrays are passed pre-offset by row0
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

// ---------------------------------------------------------------- host side
struct K2Scratch {      // carved out of one grow-only device buffer
  __half *Aimg;
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
  if (c->d_grp) cudaFree(c->d_grp);
  c->d_ops = nullptr;
  c->d_norm = nullptr;
  c->d_grp = nullptr;
  c->ops_bytes = c->grp_bytes = 0;
  c->valid = 0;
}

struct K2CbArgs { const float *codes; long M; int D; };
static cudaError_t k2_build_codebook(K2Codebook *c, const K2CbArgs &a, cudaStream_t st) {
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
    // [CbStats (64 B reserved)] [mean: D floats]
    if ((e = cudaMalloc((void **)&c->d_norm, 64 + sizeof(float) * 8192)) != cudaSuccess) return e;
  }
  if ((e = cudaMemsetAsync(c->d_norm, 0, 64, st)) != cudaSuccess) return e;
  float *mean = c->d_norm + 16;
  CbStats *cst = (CbStats *)c->d_norm;
  k2_mean_kernel<<<(a.D + 127) / 128, 128, 0, st>>>(a.codes, a.M, a.D, mean);
  k2_cb_norm_kernel<<<(unsigned)((a.M * 32 + 255) / 256), 256, 0, st>>>(a.codes, a.M, a.D, mean, cst);
  const long warps = nct * K2_TN;
  k2_cb_prep_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
      a.codes, a.M, a.D, mean, (__half *)c->d_ops, cst);
  const size_t gneed = (size_t)((a.M + K2R_GW - 1) / K2R_GW) * ((a.D + 3) / 4) * K2R_GW * 4 * sizeof(float);
  if (gneed > c->grp_bytes) {
    if (c->d_grp) cudaFree(c->d_grp);
    c->d_grp = nullptr;
    if ((e = cudaMalloc((void **)&c->d_grp, gneed)) != cudaSuccess) return e;
    c->grp_bytes = gneed;
  }
  k2_cb_regroup_kernel<<<1024, 256, 0, st>>>(a.codes, a.M, a.D, c->d_grp);
  k1_count_launch(4);
  c->Kp = Kp;
  c->valid = 1;
  return cudaGetLastError();
}

// CUDA events around the phases of the last K_EV_RING k2_search calls (a ring, so that a benchmark can
// read every step of its timed region after the closing synchronisation instead of the last one only)
// (the ring and the auxiliary stream live in the device context of the calling thread, api_internal.h)
#define g_k2ev (ctx()->k2ring[ctx()->k2calls % K_EV_RING])
#define g_k2aux (ctx()->k2aux)
#define g_k2sub (ctx()->k2sub)
#define g_k2join (ctx()->k2join)

cudaError_t k2_prepare_codebook(K2Codebook *c, const float *d_codes, long M, int D, cudaStream_t st) {
  K2CbArgs a = {d_codes, M, D};
  return k2_build_codebook(c, a, st);
}

// Sub-batch pipeline shared by both GEMM kernels: the re-rank of sub-batch i runs on a second stream
// beside the GEMM kernel of sub-batch i+1 (see k2_run_record).

static cudaError_t k2_pipeline_init() {
  cudaError_t e;
  if (g_k2aux) return cudaSuccess;
  if ((e = cudaStreamCreateWithFlags(&g_k2aux, cudaStreamNonBlocking)) != cudaSuccess) return e;
  for (int i = 0; i < 8; i++)
    if ((e = cudaEventCreateWithFlags(&g_k2sub[i], cudaEventDisableTiming)) != cudaSuccess) return e;
  return cudaEventCreateWithFlags(&g_k2join, cudaEventDisableTiming);
}
// Sub-batches of WHOLE waves of the persistent kernel (passes = a multiple of the SM count, so no
// sub-batch ends with a partly filled wave), at least ~16 waves each and at most 8 of them.
// Returns the number of sub-batches and the passes per sub-batch.
static int k2_subbatches(long passes, int num_sms, long *passes_per) {
  const long waves = (passes + num_sms - 1) / num_sms;
  long n = waves / 16;
  n = n < 1 ? 1 : (n > 8 ? 8 : n);
  *passes_per = (waves + n - 1) / n * num_sms;
  return (int)n;
}

template <int TG>
static cudaError_t k2_run_rerank(K2Codebook *c, const K1Args &a, const K2Scratch &s, long row0, long n, cudaStream_t st) {
  constexpr int LPR = TG <= 4 ? 4 : (TG <= 16 ? 16 : 32);
  const long rr_warps = (n + (32 / LPR) - 1) / (32 / LPR);
  k2_rerank_kernel<TG, LPR><<<(unsigned)((rr_warps + 7) / 8), 256, 0, st>>>(
      a.data + row0 * a.D, a.codes, n, a.M, row0, a.D, a.k, a.flags + row0, s.rs + row0, (const CbStats *)c->d_norm,
      s.cand + row0 * TG, s.thr + row0, a.listW, a.counters, a.idx + row0 * a.k, a.diff + row0 * a.k, a.nfound + row0);
  k1_count_launch(1);
  return cudaGetLastError();
}

template <int TG, int TT>
static cudaError_t k2_run_stream(K2Codebook *c, const K1Args &a, const K2Scratch &s, cudaStream_t st) {
  const int Kp = c->Kp;
  // (A resident beyond K2_ARES_MAX_KP with the two slab stages that then fit was measured at K = 528:
  // slower, 4.08 -> 4.23 ms for k = 1; the slab ring needs its depth.)
  // Round 2 repeated the experiment with HALF-size slabs (32 K columns, 16 KB of B per stage, five stages next to the
  // 135 KB of A, the code tiles the only stream from L2): 5.50 instead of 4.00 ms (k = 1), 6.84 instead of 5.19 ms
  // (k = 5).  Two MMAs per barrier round trip do not keep the tensor pipe fed; the slab size stays 64.
  const bool a_res = Kp <= K2_ARES_MAX_KP;
  const int nst = K2_NSTAGE, ks = K2_KS;
  const size_t smem = K2Smem::bytes(Kp, a_res, nst, ks);
  cudaError_t e = cudaFuncSetAttribute(k2_gemm_kernel<TG, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long ntiles_all = (a.N + K2_TM - 1) / K2_TM;
  long passes_per;
  const int nsub = k2_subbatches(ntiles_all, a.num_sms, &passes_per);
  if (nsub > 1 && (e = k2_pipeline_init()) != cudaSuccess) return e;
  const long per = passes_per * K2_TM;
  if (nsub > 1) {
    cudaEventRecord(g_k2join, st);
    cudaStreamWaitEvent(g_k2aux, g_k2join, 0);
  }
  for (int i = 0; i < nsub; i++) {
    const long row0 = i * per, n = a.N - row0 < per ? a.N - row0 : per;
    if (n <= 0) break;
    const long ntiles = (n + K2_TM - 1) / K2_TM;
    const int grid = (int)(ntiles < a.num_sms ? ntiles : a.num_sms);
    k2_gemm_kernel<TG, TT><<<grid, K2_THREADS, smem, st>>>(s.Aimg + (size_t)row0 * Kp, (const __half *)c->d_ops,
                                                          s.rs + row0, n, a.M, Kp, a_res ? 1 : 0, nst, ks, a.k,
                                                          s.cand + row0 * TG, s.thr + row0);
    k1_count_launch(1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    cudaStream_t rr = st;
    if (nsub > 1) {
      cudaEventRecord(g_k2sub[i], st);
      cudaStreamWaitEvent(g_k2aux, g_k2sub[i], 0);
      rr = g_k2aux;
    } else {
      cudaEventRecord(g_k2ev[2], st);
    }
    if ((e = k2_run_rerank<TG>(c, a, s, row0, n, rr)) != cudaSuccess) return e;
  }
  if (nsub > 1) {
    cudaEventRecord(g_k2ev[2], st);
    cudaEventRecord(g_k2join, g_k2aux);
    cudaStreamWaitEvent(st, g_k2join, 0);
  }
  return cudaSuccess;
}

// staged code tiles: as many as fit next to the four resident row tiles (3 up to K = 80, 2 at K = 96)
static int k2r_stages(int Kp) {
  const size_t a = (size_t)K2R_R * K2_TM * Kp * 2, b = (size_t)K2_TN * Kp * 2;
  int n = K2R_BST;
  while (n > 2 && a + n * b + 256 + 2 * (size_t)K2R_R * K2_TM * 16 > 227 * 1024) n--;
  return n;
}
static size_t k2r_smem_bytes(int Kp) {
  return (size_t)K2R_R * K2_TM * Kp * 2 + (size_t)k2r_stages(Kp) * K2_TN * Kp * 2 + 256 +
         2 * (size_t)K2R_R * K2_TM * sizeof(int4);        // + hand-off to the re-rank warps
}

template <int NK>
static cudaError_t k2_launch_record(K2Codebook *c, const K1Args &a, const K2Scratch &s, cudaStream_t st) {
  const long row0 = 0, n = a.N;
  const int Kp = c->Kp;
  const size_t smem = k2r_smem_bytes(Kp);
  cudaError_t e = cudaFuncSetAttribute(k2_rec_kernel<K2R_R, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long ntiles = (n + K2_TM - 1) / K2_TM;
  const long nsuper = (ntiles + K2R_R - 1) / K2R_R;
  const int grid = (int)(nsuper < a.num_sms ? nsuper : a.num_sms);
  K2RRerankArgs ra;
  ra.data = a.data + row0 * a.D; ra.grp = c->d_grp; ra.flags = a.flags + row0; ra.rs = s.rs + row0;
  ra.cst = (const CbStats *)c->d_norm; ra.listW = a.listW; ra.counters = a.counters;
  ra.idx = a.idx + row0; ra.nfound = a.nfound + row0; ra.diff = a.diff + row0;
  ra.N = n; ra.M = a.M; ra.D = a.D;
  k2_rec_kernel<K2R_R, NK><<<grid, K2R_THREADS, smem, st>>>(s.Aimg + (size_t)row0 * Kp, (const __half *)c->d_ops,
                                                           s.rs + row0, n, a.M, Kp, k2r_stages(Kp), ra);
  return cudaGetLastError();
}

static cudaError_t k2_launch_prep(K2Codebook *c, const K1Args &a, const K2Scratch &s, long row0, long n, int pack_bits,
                                  cudaStream_t st) {
  const long ntiles = (n + K2_TM - 1) / K2_TM;
  k2_row_prep_kernel<<<(unsigned)ntiles, 256, 0, st>>>(
      a.data + row0 * a.D, a.mask ? a.mask + row0 * a.D : nullptr, n, row0, a.D, a.k, pack_bits, c->d_norm + 16,
      (const CbStats *)c->d_norm, s.Aimg + (size_t)row0 * c->Kp, s.rs + row0, a.flags + row0, a.listW, a.listS,
      a.counters, a.idx + row0 * a.k, a.diff + row0 * a.k, a.nfound + row0);
  k1_count_launch(1);
  return cudaGetLastError();
}

// Record path (k == 1, short K): row prep, then ONE kernel that does the GEMM filter and the exact
// re-rank.  History (measured on C3, same box): a separate re-rank kernel after the GEMM 16.5 ms per
// step; that kernel on a second stream beside the GEMM of the next sub-batch 15.7 ms, but the GEMM
// launches stretched from 12.3 to 13.4 ms (its L2 gathers competed with the code-tile stream and the
// pipeline needed sub-batch tails); re-rank warps inside the GEMM kernel: see DESIGN.md.
// Also tried and rejected: the row prep of sub-batches 1.. on a second stream beside the record kernel of
// the sub-batch before (prep CTAs held to 32 registers so that they fit on an SM next to the record
// CTA): the record launches stretched by exactly the prep time that was hidden (12.2 -> 13.1 ms at 4
// sub-batches, step 13.9 ms either way), so prep stays one launch in front.
static cudaError_t k2_run_record(K2Codebook *c, const K1Args &a, const K2Scratch &s, cudaStream_t st) {
  cudaError_t e = k2_launch_prep(c, a, s, 0, a.N, 0, st);
  if (e != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[1], st);
  switch (c->Kp / 16) {                                 // K2R_MAX_KP / 16 = 6 unrolled issue loops
    case 1: e = k2_launch_record<1>(c, a, s, st); break;
    case 2: e = k2_launch_record<2>(c, a, s, st); break;
    case 3: e = k2_launch_record<3>(c, a, s, st); break;
    case 4: e = k2_launch_record<4>(c, a, s, st); break;
    case 5: e = k2_launch_record<5>(c, a, s, st); break;
    default: e = k2_launch_record<6>(c, a, s, st); break;
  }
  k1_count_launch(1);
  cudaEventRecord(g_k2ev[2], st);                       // the re-rank is inside the kernel: its phase is empty
  return e;
}

cudaError_t k2_kernel_ms_history(int back, float out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  DevCtx *cx = ctx();
  if (back < 0 || back >= K_EV_RING || back >= cx->k2calls) return cudaSuccess;
  cudaEvent_t *ev = cx->k2ring[(cx->k2calls - 1 - back) % K_EV_RING];
  cudaError_t e = cudaEventSynchronize(ev[4]);
  if (e != cudaSuccess) return e;
  for (int i = 0; i < 4; i++)
    if ((e = cudaEventElapsedTime(&out[i], ev[i], ev[i + 1])) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t k2_last_kernel_ms(float out[4]) { return k2_kernel_ms_history(0, out); }

cudaError_t k2_search(K2Codebook *c, const K1Args &a, void **scratch, size_t *scratch_bytes, cudaStream_t st) {
  cudaError_t e;
  if (!g_k2ev[0])
    for (int i = 0; i < 5; i++)
      if ((e = cudaEventCreate(&g_k2ev[i])) != cudaSuccess) return e;
  if (!c->valid && (e = k2_prepare_codebook(c, a.codes, a.M, a.D, st)) != cudaSuccess) return e;
  const int Kp = c->Kp;
  // k == 1 and a code tile that fits a shared-memory stage: record kernel; else K streamed in slabs
  const bool record = a.k == 1 && Kp <= K2R_MAX_KP;
  const int TG = a.k == 1 ? 4 : (a.k <= 5 ? 10 : 20);
  const long ntiles = (a.N + K2_TM - 1) / K2_TM;
  const long ntiles_alloc = (ntiles + K2R_R - 1) / K2R_R * K2R_R;   // the record kernel loads whole groups of tiles
  // scratch layout
  size_t off = 0;
  const size_t oA = off; off = align_up(off + (size_t)ntiles_alloc * K2_TM * Kp * 2, 256);
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
  s.Aimg = (__half *)(base + oA);
  s.rs = (RowStats *)(base + oR);
  s.cand = (int32_t *)(base + oC);
  s.thr = (float *)(base + oT);

  if ((e = cudaMemsetAsync(a.counters, 0, 4 * sizeof(int), st)) != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[0], st);
  if (record) {
    e = k2_run_record(c, a, s, st);                     // row prep, record GEMM, group re-rank
  } else {
    if ((e = k2_launch_prep(c, a, s, 0, a.N, 8, st)) != cudaSuccess) return e;
    cudaEventRecord(g_k2ev[1], st);
    if (a.k == 1) e = k2_run_stream<4, 3>(c, a, s, st);
    else if (a.k <= 5) e = k2_run_stream<10, 4>(c, a, s, st);
    else e = k2_run_stream<20, 4>(c, a, s, st);
  }
  if (e != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[3], st);
  // rows that failed the certificate + masked / tiny rows, then the non-finite rows
  if ((e = k1_run_lists(a, st)) != cudaSuccess) return e;
  cudaEventRecord(g_k2ev[4], st);
  ctx()->k2calls++;
  return cudaSuccess;
}

}  // namespace bmu
