// k1_search.cu -- K1: exact FP32 batch winner search (1-NN and k-NN) for sm_100a.
//
// Replaces the per-sample loops over find_winner_euc (reference lvq_pak.c:41-94) and
// find_winner_knn (lvq_pak.c:152-221).  Every distance is the reference's own sum: FP32
// sub, mul, add with a rounding after each operation, accumulated in component order
// (lvq_pak.c:63-73); no FMA, no reassociation, no split over the dimension.  Parallelism
// comes from samples and code vectors only.
//
// Kernels
//   cb_prep_kernel    codebook  -> tile layout cT[tile][i][128]           (once per codebook)
//   data_prep_kernel  data rows -> classification flags, work lists and, for k == 1, the
//                     duplicated/transposed tile layout xT[tile][i][128][2]
//   k1_fast_kernel    k == 1, regular rows: 128 samples x 128 codes per CTA pass, both
//                     operands staged in shared memory by TMA bulk copies (UBLKCP) through
//                     an mbarrier ring, 8x8 register tile per thread, packed f32x2 math
//   k1_warp_kernel    any k <= 16, masks: one warp per sample, lanes over code vectors
//   k1_seq_kernel     rows with NaN/Inf: one thread per sample, the reference's loop
//                     including its early exit (whose side effects are visible with NaN)
#include <stdlib.h>
#include "common.cuh"
#include "k1_search.h"
#include "api_internal.h"

namespace bmu {

// position of code c (0..127) inside a tile row, chosen so that the 16 code groups of a
// warp read 256 contiguous bytes per LDS.128:  c = cg*8 + q*4 + e  ->  q*64 + cg*4 + e
__host__ __device__ __forceinline__ int tile_pos(int c) {
  return ((c >> 2) & 1) * 64 + (c >> 3) * 4 + (c & 3);
}
__host__ __device__ __forceinline__ int tile_code(int pos) {
  return ((pos & 63) >> 2) * 8 + (pos >> 6) * 4 + (pos & 3);
}

__device__ __forceinline__ unsigned classify_value(float v) {
  unsigned b = __float_as_uint(v) & 0x7fffffffu;
  unsigned f = 0;
  if (b >= 0x7f800000u) f |= ROW_NONFINITE;
  // 2^-40 = 0x2b800000 ; zero is fine
  if (b != 0u && b < 0x2b800000u) f |= ROW_TINY;
  return f;
}

// ---------------------------------------------------------------- codebook prep
__global__ void cb_prep_kernel(const float *__restrict__ codes, long M, int D,
                               float *__restrict__ cT, unsigned *__restrict__ cb_flags) {
  long nct = (M + K1_TC - 1) / K1_TC;
  long total = nct * (long)D * K1_TC;
  unsigned f = 0;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    // t enumerates (tile, c, i) with i fastest so that the reads are coalesced
    int i = (int)(t % D);
    long r = t / D;
    int c = (int)(r % K1_TC);
    long ct = r / K1_TC;
    long j = ct * K1_TC + c;
    float v = 0.0f;
    if (j < M) {
      v = codes[j * (long)D + i];
      f |= classify_value(v);
    }
    cT[(ct * D + i) * K1_TC + tile_pos(c)] = v;
  }
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(cb_flags, f);
}

// ---------------------------------------------------------------- data prep
// One CTA per tile of 128 rows.  Phase A: one warp per row classifies it and routes it.
// Phase B (k == 1 only): transposed + duplicated tile for the packed kernel.
__global__ void __launch_bounds__(256)
data_prep_kernel(const float *__restrict__ data, const unsigned char *__restrict__ mask, long N,
                 int D, int k, const unsigned *__restrict__ cb_flags, float *__restrict__ xT,
                 unsigned char *__restrict__ flags, int *__restrict__ listW,
                 int *__restrict__ listS, int *__restrict__ counters, int32_t *__restrict__ idx,
                 float *__restrict__ diff, int32_t *__restrict__ nfound, int want_tiles) {
  __shared__ float tile[K1_TS][33];
  const long tile_id = blockIdx.x;
  const long n0 = tile_id * K1_TS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned cbf = *cb_flags;

  for (int r = warp; r < K1_TS; r += 8) {
    long n = n0 + r;
    if (n >= N) break;
    const float *row = data + n * (long)D;
    const unsigned char *mrow = mask ? mask + n * (long)D : nullptr;
    unsigned f = 0;
    int nmasked = 0;
    for (int i = lane; i < D; i += 32) {
      if (mrow && mrow[i]) { nmasked++; continue; }
      f |= classify_value(row[i]);
    }
    f = __reduce_or_sync(0xffffffffu, f);
    nmasked = __reduce_add_sync(0xffffffffu, nmasked);
    if (nmasked > 0) f |= ROW_MASKED;
    if (nmasked == D) f |= ROW_ALLMASKED;
    if (lane == 0) {
      flags[n] = (unsigned char)f;
      if (f & ROW_ALLMASKED) {
        // reference: return 0 with the initial winner_info contents (lvq_pak.c:50-52,165-170,75-76)
        nfound[n] = 0;
        for (int t = 0; t < k; t++) {
          idx[n * k + t] = -1;
          diff[n * k + t] = (k == 1) ? -1.0f : FLT_MAX;
        }
      } else if ((f | cbf) & ROW_NONFINITE) {
        listS[atomicAdd(&counters[1], 1)] = (int)n;
      } else if (k > 1 || ((f | cbf) & (ROW_TINY | ROW_MASKED))) {
        listW[atomicAdd(&counters[0], 1)] = (int)n;
      }
    }
  }
  if (!want_tiles) return;

  float *xt = xT + tile_id * (long)D * K1_TS * 2;
  for (int d0 = 0; d0 < D; d0 += 32) {
    __syncthreads();
    for (int r = warp; r < K1_TS; r += 8) {
      long n = n0 + r;
      int i = d0 + lane;
      tile[r][lane] = (n < N && i < D) ? data[n * (long)D + i] : 0.0f;
    }
    __syncthreads();
    // 32 components x 128 samples, written as float2 {x,x}: consecutive threads -> consecutive s
    for (int t = threadIdx.x; t < 32 * K1_TS; t += 256) {
      int il = t / K1_TS, s = t % K1_TS;
      int i = d0 + il;
      if (i < D) {
        float v = tile[s][il];
        reinterpret_cast<float2 *>(xt)[(long)i * K1_TS + s] = make_float2(v, v);
      }
    }
  }
}

// ---------------------------------------------------------------- K1-fast
// smem carve-up (bytes): c ring NC x 32 KB | x ring NX x 64 KB | barriers
struct FastSmem {
  static constexpr int C_STAGE_FLOATS = K1_DC * K1_TC;        // 8192 floats = 32 KB
  static constexpr int X_STAGE_FLOATS = K1_DC * K1_TS * 2;    // 16384 floats = 64 KB
  static constexpr int NC = 3, NX = 2;
  static constexpr size_t BYTES =
      (size_t)(NC * C_STAGE_FLOATS + NX * X_STAGE_FLOATS) * 4 + 2 * (NC + NX) * 8;
};

__global__ void __launch_bounds__(288, 1)
k1_fast_kernel(const float *__restrict__ xT, const float *__restrict__ cT, long N, long M, int D,
               const unsigned char *__restrict__ flags, int32_t *__restrict__ idx,
               float *__restrict__ diff, int32_t *__restrict__ nfound) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *cs = reinterpret_cast<float *>(smem_raw);
  float *xs = cs + FastSmem::NC * FastSmem::C_STAGE_FLOATS;
  uint64_t *bars = reinterpret_cast<uint64_t *>(xs + FastSmem::NX * FastSmem::X_STAGE_FLOATS);
  uint64_t *c_full = bars, *c_empty = bars + FastSmem::NC;
  uint64_t *x_full = bars + 2 * FastSmem::NC, *x_empty = x_full + FastSmem::NX;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long ntiles = (N + K1_TS - 1) / K1_TS;
  const int nct = (int)((M + K1_TC - 1) / K1_TC);
  const int ndch = (D + K1_DC - 1) / K1_DC;

  if (tid == 0) {
    for (int s = 0; s < FastSmem::NC; s++) { mbar_init(&c_full[s], 1); mbar_init(&c_empty[s], 8); }
    for (int s = 0; s < FastSmem::NX; s++) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 8); }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 8) {
    // ===== TMA producer: one lane streams x and c chunks through the two rings =====
    if (lane == 0) {
      unsigned cseq = 0, xseq = 0;
      for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int ct = 0; ct < nct; ct++) {
          for (int dch = 0; dch < ndch; dch++) {
            const int dc = min(K1_DC, D - dch * K1_DC);
            if (ndch > 1 || ct == 0) {
              int slot = xseq % FastSmem::NX;
              mbar_wait(&x_empty[slot], ((xseq / FastSmem::NX) & 1) ^ 1);
              uint32_t bytes = (uint32_t)dc * K1_TS * 2 * 4;
              mbar_arrive_expect_tx(&x_full[slot], bytes);
              bulk_g2s(xs + slot * FastSmem::X_STAGE_FLOATS,
                       xT + ((long)tile * D + (long)dch * K1_DC) * K1_TS * 2, bytes, &x_full[slot]);
              xseq++;
            }
            int slot = cseq % FastSmem::NC;
            mbar_wait(&c_empty[slot], ((cseq / FastSmem::NC) & 1) ^ 1);
            uint32_t bytes = (uint32_t)dc * K1_TC * 4;
            mbar_arrive_expect_tx(&c_full[slot], bytes);
            bulk_g2s(cs + slot * FastSmem::C_STAGE_FLOATS,
                     cT + ((long)ct * D + (long)dch * K1_DC) * K1_TC, bytes, &c_full[slot]);
            cseq++;
          }
        }
      }
    }
    return;
  }

  // ===== compute warps: thread = (sample group sg, code group cg), 8 samples x 8 codes =====
  const int cg = tid & 15, sg = tid >> 4;
  unsigned cseq = 0, xseq = 0;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    float best[8];
    int bidx[8];
#pragma unroll
    for (int s = 0; s < 8; s++) { best[s] = FLT_MAX; bidx[s] = 0x7fffffff; }

    for (int ct = 0; ct < nct; ct++) {
      u64 acc[8][4];
#pragma unroll
      for (int s = 0; s < 8; s++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[s][c] = 0ull;

      for (int dch = 0; dch < ndch; dch++) {
        const int dc = min(K1_DC, D - dch * K1_DC);
        const int xslot = xseq % FastSmem::NX;
        if (ndch > 1 || ct == 0) mbar_wait(&x_full[xslot], (xseq / FastSmem::NX) & 1);
        const int cslot = cseq % FastSmem::NC;
        mbar_wait(&c_full[cslot], (cseq / FastSmem::NC) & 1);

        const ulonglong2 *xp =
            reinterpret_cast<const ulonglong2 *>(xs + xslot * FastSmem::X_STAGE_FLOATS) + sg * 4;
        const ulonglong2 *cp =
            reinterpret_cast<const ulonglong2 *>(cs + cslot * FastSmem::C_STAGE_FLOATS) + cg;
#pragma unroll 2
        for (int i = 0; i < dc; i++) {
          // x: 8 samples duplicated {x,x}: 4 x LDS.128 ; c: 8 codes: 2 x LDS.128
          ulonglong2 xv0 = xp[i * 64 + 0], xv1 = xp[i * 64 + 1], xv2 = xp[i * 64 + 2],
                     xv3 = xp[i * 64 + 3];
          ulonglong2 cv0 = cp[i * 32], cv1 = cp[i * 32 + 16];
          u64 xq[8] = {xv0.x, xv0.y, xv1.x, xv1.y, xv2.x, xv2.y, xv3.x, xv3.y};
          u64 cq[4] = {cv0.x, cv0.y, cv1.x, cv1.y};
#pragma unroll
          for (int s = 0; s < 8; s++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
              u64 d = sub2(cq[c], xq[s]);          // code - sample (lvq_pak.c:70)
              acc[s][c] = add2_ftz(acc[s][c], mul2(d, d));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&c_empty[cslot]);
        cseq++;
        if (ndch > 1 || ct == nct - 1) {
          if (lane == 0) mbar_arrive(&x_empty[xslot]);
          xseq++;
        }
      }
      // fold this code tile into the running winners: ascending code index, strict '<'
      // keeps the first minimum (lvq_pak.c:79)
      const int jbase = ct * K1_TC + cg * 8;
#pragma unroll
      for (int s = 0; s < 8; s++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float lo, hi;
          unpack2(acc[s][c], lo, hi);
          int j = jbase + 2 * c;
          if (j < M && lo < best[s]) { best[s] = lo; bidx[s] = j; }
          if (j + 1 < M && hi < best[s]) { best[s] = hi; bidx[s] = j + 1; }
        }
    }
    // merge the 16 code groups of each sample group (lanes differing in the low 4 bits)
#pragma unroll
    for (int s = 0; s < 8; s++) {
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best[s], off);
        int oi = __shfl_xor_sync(0xffffffffu, bidx[s], off);
        if (ob < best[s] || (ob == best[s] && oi < bidx[s])) { best[s] = ob; bidx[s] = oi; }
      }
    }
    if (cg == 0) {
#pragma unroll
      for (int s = 0; s < 8; s++) {
        long n = tile * K1_TS + sg * 8 + s;
        if (n < N && flags[n] == 0) {     // flagged rows are answered by the other kernels
          bool found = bidx[s] != 0x7fffffff;
          idx[n] = found ? bidx[s] : -1;
          diff[n] = found ? best[s] : -1.0f;
          nfound[n] = 1;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- K1-warp
// order of the k-NN list: (diff asc, index desc) for k >= 2 (lvq_pak.c:197), first minimum
// for k == 1 (lvq_pak.c:79,160-161).  Lanes do not visit codes in ascending index order, so
// ties are resolved on the index explicitly.
__global__ void __launch_bounds__(256)
k1_warp_kernel(const float *__restrict__ data, const unsigned char *__restrict__ mask,
               const float *__restrict__ cT, long M, int D, int k, const int *__restrict__ list,
               const int *__restrict__ count, int max_cnt, int32_t *__restrict__ idx,
               float *__restrict__ diff, int32_t *__restrict__ nfound) {
  __shared__ float sl_d[8][BMU_KMAX_];
  __shared__ int sl_i[8][BMU_KMAX_];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = *count;
  if (cnt >= max_cnt) return;                            // long list: k1_warp8_kernel answers it
  const int nct = (int)((M + K1_TC - 1) / K1_TC);
  const bool knn_rule = k > 1;
  // short lists (K2 certificate failures, a few masked rows): S warps of a CTA share one row
  // and split its code tiles, so that the tail does not run at one warp per SM
  int S = 1;
  const long total_warps = (long)gridDim.x * 8;
  while (S < 8 && S * 2 <= nct && (long)cnt * S * 2 <= total_warps) S *= 2;
  const int rows_per_cta = 8 / S;
  const int rloc = warp / S, sub = warp % S;

  for (long base = (long)blockIdx.x * rows_per_cta; base < cnt; base += (long)gridDim.x * rows_per_cta) {
    const long w = base + rloc;
    const bool valid = w < cnt;
    const long n = valid ? list[w] : 0;
    const float *x = data + n * (long)D;
    const unsigned char *mk = mask ? mask + n * (long)D : nullptr;
    float ld[BMU_KMAX_];
    int li[BMU_KMAX_];
#pragma unroll
    for (int t = 0; t < BMU_KMAX_; t++) { ld[t] = FLT_MAX; li[t] = -1; }

    if (valid)
    for (int ct = sub; ct < nct; ct += S) {
      const float *cbase = cT + (long)ct * D * K1_TC + lane;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (mk) {
        for (int i = 0; i < D; i++) {
          if (mk[i]) continue;           // warp-uniform
          const float xi = __ldg(x + i);
          const float *cr = cbase + (long)i * K1_TC;
          a0 = sq_acc(a0, cr[0], xi);
          a1 = sq_acc(a1, cr[32], xi);
          a2 = sq_acc(a2, cr[64], xi);
          a3 = sq_acc(a3, cr[96], xi);
        }
      } else {
#pragma unroll 16
        for (int i = 0; i < D; i++) {
          const float xi = __ldg(x + i);
          const float *cr = cbase + (long)i * K1_TC;
          a0 = sq_acc(a0, cr[0], xi);
          a1 = sq_acc(a1, cr[32], xi);
          a2 = sq_acc(a2, cr[64], xi);
          a3 = sq_acc(a3, cr[96], xi);
        }
      }
      float av[4] = {a0, a1, a2, a3};
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int j = ct * K1_TC + tile_code(lane + 32 * r);
        const float d = av[r];
        if (j >= M) continue;
        if (k == 1) {
          if (d < ld[0] || (d == ld[0] && li[0] >= 0 && j < li[0])) { ld[0] = d; li[0] = j; }
        } else {
          // insertion; +Inf / NaN / > FLT_MAX never enter (lvq_pak.c:197 with FLT_MAX init)
          if (!(d <= FLT_MAX)) continue;
          int p = 0;
          while (p < k && !(d < ld[p] || (d == ld[p] && (li[p] < 0 || j > li[p])))) p++;
          if (p < k) {
            for (int t = k - 1; t > p; t--) { ld[t] = ld[t - 1]; li[t] = li[t - 1]; }
            ld[p] = d; li[p] = j;
          }
        }
      }
    }
    // warp merge: k rounds, each takes the best head among the 32 per-lane lists
    int head = 0;
    for (int t = 0; t < k; t++) {
      float d = head < k ? ld[head] : FLT_MAX;
      int j = head < k ? li[head] : -1;
      float bd = d; int bj = j;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, bd, off);
        int oj = __shfl_xor_sync(0xffffffffu, bj, off);
        bool take;
        if (oj < 0) take = false;
        else if (bj < 0) take = true;
        else take = od < bd || (od == bd && (knn_rule ? oj > bj : oj < bj));
        if (take) { bd = od; bj = oj; }
      }
      if (bj >= 0 && bj == j) head++;     // exactly one lane owns code bj
      if (lane == 0) { sl_d[warp][t] = bd; sl_i[warp][t] = bj; }
    }
    __syncthreads();
    if (valid && sub == 0 && lane == 0) {
      // merge the S sorted lists of this row's warps
      int hp[8];
      for (int q = 0; q < S; q++) hp[q] = 0;
      for (int t = 0; t < k; t++) {
        float bd = FLT_MAX; int bj = -1, bq = -1;
        for (int q = 0; q < S; q++) {
          if (hp[q] >= k) continue;
          const float od = sl_d[warp + q][hp[q]];
          const int oj = sl_i[warp + q][hp[q]];
          if (oj < 0) continue;
          if (bj < 0 || od < bd || (od == bd && (knn_rule ? oj > bj : oj < bj))) { bd = od; bj = oj; bq = q; }
        }
        if (bq >= 0) hp[bq]++;
        if (k == 1) {
          idx[n] = bj;
          diff[n] = bj >= 0 ? bd : -1.0f;
        } else {
          idx[n * k + t] = bj;
          diff[n * k + t] = bj >= 0 ? bd : FLT_MAX;
        }
      }
      nfound[n] = k;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- K1-list8 (k == 1, no masks)
// The rows of a work list -- the K2 certificate failures, tiny-magnitude rows -- answered exactly.  One
// row per warp streams the whole codebook through L2 by itself (C4: 294 rows x 8 MB = 2.4 GB, 0.5 ms), so a
// warp takes EIGHT listed rows against each 128-code tile (8 rows x 4 codes per lane in registers), which
// divides that traffic by eight; and because a short list then has far fewer groups than the GPU has
// warps, the code tiles of a group are cut into S slices that go to different warps ANYWHERE in the grid.
// The slices of a row meet in a 64-bit atomicMin on (distance bits << 32 | index) -- distances are
// non-negative, so the unsigned order is (distance, then lower index): the first-minimum rule of
// lvq_pak.c:79 -- and the warp that finishes a group last writes the results and resets the scratch.
__global__ void __launch_bounds__(256, 2)
k1_list8_kernel(const float *__restrict__ data, const float *__restrict__ cT, long M, int D,
                const int *__restrict__ list, const int *__restrict__ count, const unsigned char *__restrict__ flags,
                const unsigned *__restrict__ cb_flags, u64 *__restrict__ keys,
                int *__restrict__ done, int32_t *__restrict__ idx, float *__restrict__ diff,
                int32_t *__restrict__ nfound) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = *count;
  if (cnt <= 0) return;
  const bool cb_plain = *cb_flags == 0;                  // no tiny / non-finite code vector
  const long W = (long)gridDim.x * 8, wid = (long)blockIdx.x * 8 + warp;
  const long groups = (cnt + 7) / 8;
  const int nct = (int)((M + K1_TC - 1) / K1_TC);
  long S = W / groups;                                   // slices per group: uniform over the grid
  S = S < 1 ? 1 : (S > nct ? nct : S);
  const long items = groups * S;
  for (long item = wid; item < items; item += W) {
    const long gidx = item / S;
    const int sl = (int)(item % S);
    const int ct0 = (int)((long)sl * nct / S), ct1 = (int)((long)(sl + 1) * nct / S);
    const float *xp[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      long w = gidx * 8 + r;
      if (w >= cnt) w = gidx * 8;                          // harmless duplicate, never written
      xp[r] = data + (long)list[w] * D;
    }
    // Rows without tiny magnitudes (the K2 certificate failures; flags bit ROW_TINY clear) against a plain codebook
    // take the PACKED arithmetic of k1_fast_kernel -- sub / mul / ftz-add on two rows at a time, the same results
    // bit for bit (common.cuh) in 56 instead of 96 issue slots per component and code quad; any tiny row in the
    // group sends the whole group through the scalar non-ftz sequence.
    bool packed = cb_plain;
    {
      long w = gidx * 8 + (lane & 7);
      if (w >= cnt) w = gidx * 8;
      const bool tiny = (flags[list[w]] & ROW_TINY) != 0;
      packed = packed && !__any_sync(0xffffffffu, tiny);
    }
    u64 best[8];
#pragma unroll
    for (int r = 0; r < 8; r++) best[r] = ~0ull;
    for (int ct = ct0; ct < ct1; ct++) {
      const float *cbase = cT + (long)ct * D * K1_TC + lane;
      float acc[8][4];
      if (packed) {
        u64 a2[4][4];
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
          for (int q = 0; q < 4; q++) a2[p][q] = pack2(0.0f, 0.0f);
#pragma unroll 4
        for (int i = 0; i < D; i++) {
          const float *cr = cbase + (long)i * K1_TC;
          const float c0 = cr[0], c1 = cr[32], c2 = cr[64], c3 = cr[96];
          const u64 cc[4] = {pack2(c0, c0), pack2(c1, c1), pack2(c2, c2), pack2(c3, c3)};
#pragma unroll
          for (int p = 0; p < 4; p++) {
            const u64 xx = pack2(__ldg(xp[2 * p] + i), __ldg(xp[2 * p + 1] + i));
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const u64 d = sub2(cc[q], xx);                 // code - sample (lvq_pak.c:70)
              a2[p][q] = add2_ftz(a2[p][q], mul2(d, d));
            }
          }
        }
#pragma unroll
        for (int p = 0; p < 4; p++)
#pragma unroll
          for (int q = 0; q < 4; q++) unpack2(a2[p][q], acc[2 * p][q], acc[2 * p + 1][q]);
      } else {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int q = 0; q < 4; q++) acc[r][q] = 0.0f;
#pragma unroll 4
        for (int i = 0; i < D; i++) {
          const float *cr = cbase + (long)i * K1_TC;
          const float c0 = cr[0], c1 = cr[32], c2 = cr[64], c3 = cr[96];
#pragma unroll
          for (int r = 0; r < 8; r++) {
            const float xi = __ldg(xp[r] + i);
            acc[r][0] = sq_acc(acc[r][0], c0, xi);
            acc[r][1] = sq_acc(acc[r][1], c1, xi);
            acc[r][2] = sq_acc(acc[r][2], c2, xi);
            acc[r][3] = sq_acc(acc[r][3], c3, xi);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int j = ct * K1_TC + tile_code(lane + 32 * q);
        if (j >= M) continue;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          // strict < against the FLT_MAX start value (lvq_pak.c:60,79); NaN never wins; the
          // 64-bit key orders by distance, then by the lower index
          if (acc[r][q] < FLT_MAX) {
            const u64 key = ((u64)__float_as_uint(acc[r][q]) << 32) | (unsigned)j;
            best[r] = key < best[r] ? key : best[r];
          }
        }
      }
    }
    u64 mine = ~0ull;                                       // lane r < 8 ends up with row r's minimum
#pragma unroll
    for (int r = 0; r < 8; r++) {
      u64 b = best[r];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const u64 o = __shfl_xor_sync(0xffffffffu, b, off);
        b = o < b ? o : b;
      }
      if (lane == r) mine = b;
    }
    const long w = gidx * 8 + lane;
    if (lane < 8 && w < cnt && mine != ~0ull) atomicMin(&keys[w], mine);
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence();
      last = atomicAdd(&done[gidx], 1) == (int)S - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      if (lane < 8 && w < cnt) {
        const u64 b = atomicExch(&keys[w], ~0ull);          // read the minimum and leave the slot reset
        const long n = list[w];
        idx[n] = b == ~0ull ? -1 : (int)(unsigned)b;
        diff[n] = b == ~0ull ? -1.0f : __uint_as_float((unsigned)(b >> 32));
        nfound[n] = 1;
      }
      if (lane == 0) done[gidx] = 0;
    }
  }
}

// ---------------------------------------------------------------- K1-listk (2 <= k <= 5, no masks)
// The k-NN version of k1_list8_kernel: FOUR listed rows per warp against each 128-code tile (codebook traffic
// through L2 / 4 compared with k1_warp_kernel, which this replaces for unmasked rows), per-lane sorted lists of
// 64-bit keys (distance bits << 32 | ~index: the unsigned order is the reference's k-NN order, distance ascending
// then index DESCENDING, lvq_pak.c:197), k rounds of warp minimum per row.  A short list is cut into S slices of
// code tiles handled by warps anywhere in the grid; every slice leaves its sorted k keys per row in `parts`, the
// warp that finishes a group last merges the S lists.  Long lists (more groups than warps) run with S = 1 and
// write their results directly.
#define K1_LISTK_KT 5
#define K1_LISTK_SMAX 32
#define K1_LISTK_CAP 16384        // listed rows that can be sliced (beyond that the list is long enough for S = 1)
__global__ void __launch_bounds__(256, 2)
k1_listk_kernel(const float *__restrict__ data, const float *__restrict__ cT, long M, int D, int k,
                const int *__restrict__ list, const int *__restrict__ count, u64 *__restrict__ parts,
                int *__restrict__ done, int32_t *__restrict__ idx, float *__restrict__ diff,
                int32_t *__restrict__ nfound) {
  constexpr int KT = K1_LISTK_KT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = *count;
  if (cnt <= 0) return;
  const long W = (long)gridDim.x * 8, wid = (long)blockIdx.x * 8 + warp;
  const long groups = (cnt + 3) / 4;
  const int nct = (int)((M + K1_TC - 1) / K1_TC);
  long S = cnt <= K1_LISTK_CAP ? W / groups : 1;
  S = S < 1 ? 1 : (S > nct ? nct : (S > K1_LISTK_SMAX ? K1_LISTK_SMAX : S));
  const long items = groups * S;
  for (long item = wid; item < items; item += W) {
    const long gidx = item / S;
    const int sl = (int)(item % S);
    const int ct0 = (int)((long)sl * nct / S), ct1 = (int)((long)(sl + 1) * nct / S);
    const float *xp[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      long w = gidx * 4 + r;
      if (w >= cnt) w = gidx * 4;
      xp[r] = data + (long)list[w] * D;
    }
    u64 ld[4][KT];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int t = 0; t < KT; t++) ld[r][t] = ~0ull;
    for (int ct = ct0; ct < ct1; ct++) {
      const float *cbase = cT + (long)ct * D * K1_TC + lane;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[r][q] = 0.0f;
#pragma unroll 4
      for (int i = 0; i < D; i++) {
        const float *cr = cbase + (long)i * K1_TC;
        const float c0 = cr[0], c1 = cr[32], c2 = cr[64], c3 = cr[96];
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const float xi = __ldg(xp[r] + i);
          acc[r][0] = sq_acc(acc[r][0], c0, xi);
          acc[r][1] = sq_acc(acc[r][1], c1, xi);
          acc[r][2] = sq_acc(acc[r][2], c2, xi);
          acc[r][3] = sq_acc(acc[r][3], c3, xi);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int j = ct * K1_TC + tile_code(lane + 32 * q);
        if (j >= M) continue;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          const float d = acc[r][q];
          if (!(d <= FLT_MAX)) continue;              // +Inf / NaN never enter (lvq_pak.c:197 against the FLT_MAX start)
          u64 key = ((u64)__float_as_uint(d) << 32) | (u64)(0xFFFFFFFFu - (unsigned)j);
          if (key < ld[r][KT - 1]) {
            // sorted insertion with static indexing: once placed, the rest shifts down and the largest drops out
#pragma unroll
            for (int t = 0; t < KT; t++) {
              const u64 cur = ld[r][t];
              const bool sw = key < cur;
              ld[r][t] = sw ? key : cur;
              key = sw ? cur : key;
            }
          }
        }
      }
    }
    // k rounds of warp minimum per row: lane t < k keeps the t-th key of row r in out[r]
    u64 out[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      out[r] = ~0ull;
      for (int t = 0; t < k; t++) {
        u64 b = ld[r][0];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const u64 o = __shfl_xor_sync(0xffffffffu, b, off);
          b = o < b ? o : b;
        }
        if (b != ~0ull && ld[r][0] == b) {            // exactly one lane owns this (distance, code) pair: pop it
#pragma unroll
          for (int u = 0; u + 1 < KT; u++) ld[r][u] = ld[r][u + 1];
          ld[r][KT - 1] = ~0ull;
        }
        if (lane == t) out[r] = b;
      }
    }
    if (S == 1) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const long w = gidx * 4 + r;
        if (w < cnt && lane < k) {
          const long n = list[w];
          const u64 b = out[r];
          idx[n * k + lane] = b == ~0ull ? -1 : (int)(0xFFFFFFFFu - (unsigned)b);
          diff[n * k + lane] = b == ~0ull ? FLT_MAX : __uint_as_float((unsigned)(b >> 32));
          if (lane == 0) nfound[n] = k;
        }
      }
      continue;
    }
    // sliced: leave the slice's sorted keys, the last warp of the group merges the S lists of every row
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const long w = gidx * 4 + r;
      if (w < cnt && lane < KT) parts[((size_t)w * K1_LISTK_SMAX + sl) * KT + lane] = lane < k ? out[r] : ~0ull;
    }
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence();
      last = atomicAdd(&done[gidx], 1) == (int)S - 1;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      const long w = gidx * 4 + lane;
      if (lane < 4 && w < cnt) {
        const long n = list[w];
        const volatile u64 *pp = parts + (size_t)w * K1_LISTK_SMAX * KT;
        unsigned char hp[K1_LISTK_SMAX];
        for (int q = 0; q < (int)S; q++) hp[q] = 0;
        for (int t = 0; t < k; t++) {
          u64 b = ~0ull;
          int bq = -1;
          for (int q = 0; q < (int)S; q++) {
            if (hp[q] >= KT) continue;
            const u64 o = pp[q * KT + hp[q]];
            if (o < b) { b = o; bq = q; }
          }
          if (bq >= 0) hp[bq]++;
          idx[n * k + t] = b == ~0ull ? -1 : (int)(0xFFFFFFFFu - (unsigned)b);
          diff[n * k + t] = b == ~0ull ? FLT_MAX : __uint_as_float((unsigned)(b >> 32));
        }
        nfound[n] = k;
      }
      if (lane == 0) done[gidx] = 0;
    }
  }
}

// ---------------------------------------------------------------- K1-seq
// One thread per listed row: the reference's loops verbatim in behaviour, including the
// early exit `if (difference > bound) break` (lvq_pak.c:72,195).  With NaN in the inputs the
// early exit changes which code vectors reach the (NaN-accepting) k-NN insertion, so these
// rows cannot use full sums.
__global__ void __launch_bounds__(128)
k1_seq_kernel(const float *__restrict__ data, const unsigned char *__restrict__ mask,
              const float *__restrict__ codes, long M, int D, int k, const int *__restrict__ list,
              const int *__restrict__ count, int32_t *__restrict__ idx, float *__restrict__ diff,
              int32_t *__restrict__ nfound) {
  const int cnt = *count;
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < cnt; w += gridDim.x * blockDim.x) {
    const long n = list[w];
    const float *x = data + n * (long)D;
    const unsigned char *mk = mask ? mask + n * (long)D : nullptr;
    float ld[BMU_KMAX_];
    int li[BMU_KMAX_];
    if (k == 1) { ld[0] = -1.0f; li[0] = -1; }
    else for (int t = 0; t < k; t++) { ld[t] = FLT_MAX; li[t] = -1; }
    float bound = FLT_MAX;
    for (long j = 0; j < M; j++) {
      const float *c = codes + j * (long)D;
      float acc = 0.0f;
      for (int i = 0; i < D; i++) {
        if (mk && mk[i]) continue;
        acc = sq_acc(acc, c[i], x[i]);
        if (acc > bound) break;
      }
      if (k == 1) {
        if (acc < bound) { bound = acc; ld[0] = acc; li[0] = (int)j; }
      } else {
        int p = 0;
        while (p < k && acc > ld[p]) p++;
        if (p < k) {
          for (int t = k - 1; t > p; t--) { ld[t] = ld[t - 1]; li[t] = li[t - 1]; }
          ld[p] = acc; li[p] = (int)j;
        }
        bound = ld[k - 1];
      }
    }
    for (int t = 0; t < k; t++) { idx[n * k + t] = li[t]; diff[n * k + t] = ld[t]; }
    nfound[n] = k;
  }
}

// ---------------------------------------------------------------- launchers
// (launch counter: bmu_api.cu, one atomic for all device contexts)

size_t k1_cT_floats(long M, int D) { return (size_t)((M + K1_TC - 1) / K1_TC) * D * K1_TC; }
size_t k1_listk_parts_bytes() { return (size_t)K1_LISTK_CAP * K1_LISTK_SMAX * K1_LISTK_KT * sizeof(u64); }
size_t k1_xT_floats(long N, int D) { return (size_t)((N + K1_TS - 1) / K1_TS) * D * K1_TS * 2; }

cudaError_t k1_prepare_codebook(const float *d_codes, long M, int D, float *d_cT,
                                unsigned *d_cb_flags, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d_cb_flags, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  size_t total = k1_cT_floats(M, D);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cb_prep_kernel<<<blocks, 256, 0, st>>>(d_codes, M, D, d_cT, d_cb_flags);
  k1_count_launch(1);
  return cudaGetLastError();
}

// CUDA events around each kernel of the last k1_search call, on the launching stream
// (the ring lives in the device context of the calling thread, api_internal.h)

cudaError_t k1_kernel_ms_history(int back, float out[4]) {
  for (int i = 0; i < 4; i++) out[i] = 0.0f;
  DevCtx *c = ctx();
  if (back < 0 || back >= K_EV_RING || back >= c->k1calls) return cudaSuccess;
  cudaEvent_t *ev = c->k1ring[(c->k1calls - 1 - back) % K_EV_RING];
  cudaError_t e = cudaEventSynchronize(ev[4]);
  if (e != cudaSuccess) return e;
  for (int i = 0; i < 4; i++) {
    e = cudaEventElapsedTime(&out[i], ev[i], ev[i + 1]);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t k1_last_kernel_ms(float out[4]) { return k1_kernel_ms_history(0, out); }

cudaError_t k1_search(const K1Args &a, cudaStream_t st) {
  cudaError_t e;
  DevCtx *c = ctx();
  cudaEvent_t *g_ev = c->k1ring[c->k1calls % K_EV_RING];
  if (!g_ev[0])
    for (int i = 0; i < 5; i++)
      if ((e = cudaEventCreate(&g_ev[i])) != cudaSuccess) return e;
  // (function attributes are per device: set on every call, it is a cheap host-side table update)
  e = cudaFuncSetAttribute(k1_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FastSmem::BYTES);
  if (e != cudaSuccess) return e;
  if (a.N <= 0) return cudaSuccess;
  const long ntiles = (a.N + K1_TS - 1) / K1_TS;
  e = cudaMemsetAsync(a.counters, 0, 4 * sizeof(int), st);
  if (e != cudaSuccess) return e;
  const int want_tiles = (a.k == 1 && !a.skip_fast) ? 1 : 0;
  cudaEventRecord(g_ev[0], st);
  data_prep_kernel<<<(unsigned)ntiles, 256, 0, st>>>(a.data, a.mask, a.N, a.D, a.k, a.cb_flags,
                                                     a.xT, a.flags, a.listW, a.listS, a.counters,
                                                     a.idx, a.diff, a.nfound, want_tiles);
  k1_count_launch(1);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  cudaEventRecord(g_ev[1], st);
  if (want_tiles) {
    int grid = (int)(ntiles < a.num_sms ? ntiles : a.num_sms);
    k1_fast_kernel<<<grid, 288, FastSmem::BYTES, st>>>(a.xT, a.cT, a.N, a.M, a.D, a.flags, a.idx,
                                                       a.diff, a.nfound);
    k1_count_launch(1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  cudaEventRecord(g_ev[2], st);
  if ((e = k1_run_warp_list(a, st)) != cudaSuccess) return e;
  cudaEventRecord(g_ev[3], st);
  if ((e = k1_run_seq_list(a, st)) != cudaSuccess) return e;
  cudaEventRecord(g_ev[4], st);
  c->k1calls++;
  return cudaSuccess;
}

// rows in listW (count in counters[0]): masked / tiny rows, k >= 2 rows, K2 certificate failures
cudaError_t k1_run_warp_list(const K1Args &a, cudaStream_t st) {
  // The list length is only known on the device.  k == 1 without masks: k1_list8_kernel (eight rows per
  // warp, code tiles sliced over the whole grid, 2 CTAs per SM); 2 <= k <= 5 without masks: k1_listk_kernel
  // (four rows per warp); larger k or masks: k1_warp_kernel, whose warps share a row.  Both are persistent over the list and return at once when it is empty.
  if (a.k == 1 && a.mask == nullptr && a.lkeys) {
    k1_list8_kernel<<<a.num_sms * 2, 256, 0, st>>>(a.data, a.cT, a.M, a.D, a.listW, a.counters + 0, a.flags, a.cb_flags,
                                                  a.lkeys, a.ldone, a.idx, a.diff, a.nfound);
    k1_count_launch(1);
    return cudaGetLastError();
  }
  if (a.k >= 2 && a.k <= K1_LISTK_KT && a.mask == nullptr && a.lparts) {
    k1_listk_kernel<<<a.num_sms * 2, 256, 0, st>>>(a.data, a.cT, a.M, a.D, a.k, a.listW, a.counters + 0, a.lparts, a.ldone,
                                                  a.idx, a.diff, a.nfound);
    k1_count_launch(1);
    return cudaGetLastError();
  }
  long warps = a.N < 8L * a.num_sms * 8 ? a.N : 8L * a.num_sms * 8;   // persistent over the list
  if (warps < 8) warps = 8;
  int grid = (int)((warps + 7) / 8);
  k1_warp_kernel<<<grid, 256, 0, st>>>(a.data, a.mask, a.cT, a.M, a.D, a.k, a.listW,
                                       a.counters + 0, 0x7fffffff, a.idx, a.diff, a.nfound);
  k1_count_launch(1);
  return cudaGetLastError();
}

// rows in listS (count in counters[1]): NaN / Inf inputs
cudaError_t k1_run_seq_list(const K1Args &a, cudaStream_t st) {
  k1_seq_kernel<<<a.num_sms * 2, 128, 0, st>>>(a.data, a.mask, a.codes, a.M, a.D, a.k, a.listS,
                                               a.counters + 1, a.idx, a.diff, a.nfound);
  k1_count_launch(1);
  return cudaGetLastError();
}

cudaError_t k1_run_lists(const K1Args &a, cudaStream_t st) {
  cudaError_t e = k1_run_warp_list(a, st);
  if (e != cudaSuccess) return e;
  return k1_run_seq_list(a, st);
}

}  // namespace bmu
