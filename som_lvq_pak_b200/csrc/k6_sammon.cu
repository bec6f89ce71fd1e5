// k6_sammon.cu -- K6: Sammon's mapping of a codebook (sammon_iterate, sammon.c:129-262) and the
// zero-distance pair search of remove_identicals (sammon.c:83-127).
//
// The reference keeps the M(M-1)/2 mutual distances vector_dist_euc(a, b) in a triangular table and
// runs `length` Jacobi sweeps: for every point j a float/double mixed sum over all other points k IN
// INDEX ORDER (sammon.c:200-218), then xu[j] = x[j] + 0.2*e1x/|e2x| (220-221), a sequential centre of
// mass (225-233) and x = xu - centre (234-237).  Every operation below repeats the reference's
// operand types and order (float ops where C computes in float, double where an operand is double;
// -fmad=false, IEEE division and square root), so the positions are bit-identical:
//   k6_pair_dist_kernel   the full symmetric M x M distance matrix (row j contiguous, so a warp reads
//                         its row coalesced) from the shared 64 x 64 pair tile (pairtile.cuh), and the
//                         list of pairs at distance exactly 0 for remove_identicals;
//   k6_sweep_kernel       a CTA per 32 points j: seven warps compute the four terms of every (j, k) pair
//                         into shared memory, chunk by chunk, while the eighth advances the 32 running
//                         sums, one per lane, in k order (a float add is not associative, so each chain
//                         is sequential; giving a chain a lane keeps the conversions it needs fully used);
//   k6_center_kernel      one CTA: sequential float sums of xu / yu, then the parallel subtraction.
// FP64-pipe and latency bound; the distance matrix (4 M^2 bytes) stays resident across sweeps.
#include "common.cuh"
#include "pairtile.cuh"
#include "k6_sammon.h"

namespace bmu {

template <bool MASKED>
__global__ void __launch_bounds__(256)
k6_pair_dist_kernel(const float *__restrict__ codes, const unsigned char *__restrict__ mask, long M, int D,
                    int ntiles, float *__restrict__ dd, int32_t *__restrict__ zero_pairs, long cap,
                    unsigned long long *__restrict__ nzero) {
  __shared__ PairTileSmem<MASKED> ts;
  int ti, tj;
  pair_tile_index(blockIdx.x, ntiles, ti, tj);
  const long i0 = (long)ti * PT_T, j0 = (long)tj * PT_T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
  int nmask[4][4];
  pair_tile_sums<MASKED>(codes, mask, M, D, i0, j0, ts, acc, nmask);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const long gi = i0 + ty * 4 + r;
    if (gi >= M) continue;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const long gj = j0 + tx * 4 + q;
      if (gj >= M) continue;
      float d = (float)sqrt((double)acc[r][q]);                  // lvq_pak.c:315
      if (MASKED && nmask[r][q] == D) d = -1.0f;                 // lvq_pak.c:312-313
      if (dd) {
        dd[gi * M + gj] = d;
        dd[gj * M + gi] = d;
      }
      if (zero_pairs && gj > gi && d == 0.0f) {
        const unsigned long long slot = atomicAdd(nzero, 1ULL);
        if ((long)slot < cap) { zero_pairs[2 * slot] = (int32_t)gi; zero_pairs[2 * slot + 1] = (int32_t)gj; }
      }
    }
  }
}

#define K6_JT 32          // points per CTA: one lane of the summing warp each
#define K6_KC 32          // partners per chunk
#define K6_THREADS 256    // warp 0 sums, warps 1..7 compute terms

struct K6Terms {
  float t1x[K6_KC][K6_JT], t1y[K6_KC][K6_JT];
  double t2x[K6_KC][K6_JT], t2y[K6_KC][K6_JT];
};

// One CTA owns 32 points j.  The running sums of a point are a strictly sequential float/double chain
// over all partners k, so the chain is given a LANE (32 chains advance per instruction of warp 0)
// while the other seven warps compute the terms of the next chunk of 32 partners into the second
// shared-memory buffer ([k][j]: conflict free for both sides).  dd is symmetric, so the producers read
// dd[k][j0 + lane], contiguous across the lanes.
__global__ void __launch_bounds__(K6_THREADS)
k6_sweep_kernel(const float *__restrict__ dd, long M, const float *__restrict__ x, const float *__restrict__ y,
                float *__restrict__ xu, float *__restrict__ yu) {
  __shared__ K6Terms buf[2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long j = (long)blockIdx.x * K6_JT + lane;
  const long jc = j < M ? j : M - 1;                       // idle lanes of the last CTA repeat the last point
  const float xj = x[jc], yj = y[jc];
  float e1x = 0.0f, e1y = 0.0f;                            // warp 0 only
  double e2x = 0.0, e2y = 0.0;                             // float-valued: rounded to float after every add
  // (rounding the significand in integer arithmetic instead of the two conversions was measured: slower)
  const long nchunks = (M + K6_KC - 1) / K6_KC;
  constexpr int NP = (K6_KC + 6) / 7;                      // partners per producer warp and chunk
  // a producer's inputs for the NEXT chunk are loaded while it computes the current one (the dd row
  // segments come from HBM: 4 M^2 bytes per sweep)
  float nxt_x[NP], nxt_y[NP], nxt_d[NP];
#pragma unroll
  for (int it = 0; it < NP; it++) {
    const int kl = w - 1 + 7 * it;
    const bool ok = w > 0 && kl < K6_KC && kl < M;
    nxt_x[it] = ok ? x[kl] : 0.0f;
    nxt_y[it] = ok ? y[kl] : 0.0f;
    nxt_d[it] = ok ? dd[(long)kl * M + jc] : 0.0f;
  }
  for (long c = 0; c <= nchunks; c++) {
    if (w > 0) {
      if (c < nchunks) {
        K6Terms &t = buf[c & 1];
        const long k0 = c * K6_KC;
        float cur_x[NP], cur_y[NP], cur_d[NP];
#pragma unroll
        for (int it = 0; it < NP; it++) {
          cur_x[it] = nxt_x[it]; cur_y[it] = nxt_y[it]; cur_d[it] = nxt_d[it];
          const int kl = w - 1 + 7 * it;
          const long kn = k0 + K6_KC + kl;
          const bool ok = kl < K6_KC && kn < M;
          nxt_x[it] = ok ? x[kn] : 0.0f;
          nxt_y[it] = ok ? y[kn] : 0.0f;
          nxt_d[it] = ok ? dd[kn * M + jc] : 0.0f;
        }
        // five partners per producer warp, unrolled: their square-root / division chains are independent
#pragma unroll
        for (int it = 0; it < NP; it++) {
          const int kl = w - 1 + 7 * it;
          const long k = k0 + kl;
          if (kl >= K6_KC || k >= M) continue;
          const float xd = __fsub_rn(xj, cur_x[it]), yd = __fsub_rn(yj, cur_y[it]);
          // (float) sqrt((double) xd * xd + yd * yd): double product + float product, added in double
          const float dpj = (float)sqrt(__dadd_rn(__dmul_rn((double)xd, (double)xd), (double)__fmul_rn(yd, yd)));
          const float dt = cur_d[it];
          const float dq = __fsub_rn(dt, dpj), dr = __fmul_rn(dt, dpj);
          t.t1x[kl][lane] = __fdiv_rn(__fmul_rn(xd, dq), dr);                     // xd * dq / dr
          t.t1y[kl][lane] = __fdiv_rn(__fmul_rn(yd, dq), dr);
          const double u = __dadd_rn(1.0, (double)__fdiv_rn(dq, dpj));           // 1.0 + dq / dpj
          const double ddq = (double)dq, ddpj = (double)dpj, ddr = (double)dr;
          // (dq - xd * xd * (1.0 + dq / dpj) / dpj) / dr
          t.t2x[kl][lane] = __ddiv_rn(__dsub_rn(ddq, __ddiv_rn(__dmul_rn((double)__fmul_rn(xd, xd), u), ddpj)), ddr);
          t.t2y[kl][lane] = __ddiv_rn(__dsub_rn(ddq, __ddiv_rn(__dmul_rn((double)__fmul_rn(yd, yd), u), ddpj)), ddr);
        }
      }
    } else if (c > 0) {
      const K6Terms &t = buf[(c - 1) & 1];
      const long k0 = (c - 1) * K6_KC;
      const int n = (M - k0 < K6_KC) ? (int)(M - k0) : K6_KC;
#pragma unroll 4
      for (int kl = 0; kl < n; kl++) {
        if (k0 + kl == j) continue;                                               // sammon.c:202-203
        e1x = __fadd_rn(e1x, t.t1x[kl][lane]);
        e1y = __fadd_rn(e1y, t.t1y[kl][lane]);
        e2x = (double)(float)__dadd_rn(e2x, t.t2x[kl][lane]);                 // e2x += <double>: float result
        e2y = (double)(float)__dadd_rn(e2y, t.t2y[kl][lane]);
      }
    }
    __syncthreads();
  }
  if (w == 0 && j < M) {
    // x[j] + MAGIC * e1x / fabs(e2x): all in double, rounded on the store (sammon.c:220-221)
    xu[j] = (float)__dadd_rn((double)xj, __ddiv_rn(__dmul_rn(0.2, (double)e1x), fabs(e2x)));
    yu[j] = (float)__dadd_rn((double)yj, __ddiv_rn(__dmul_rn(0.2, (double)e1y), fabs(e2y)));
  }
}

__global__ void __launch_bounds__(1024)
k6_center_kernel(long M, const float *__restrict__ xu, const float *__restrict__ yu, float *__restrict__ x,
                 float *__restrict__ y) {
  __shared__ float bx[1024], by[1024];
  __shared__ float cx, cy;
  float xx = 0.0f, yy = 0.0f;                            // meaningful in thread 0 only
  for (long j0 = 0; j0 < M; j0 += 1024) {
    const long j = j0 + threadIdx.x;
    if (j < M) { bx[threadIdx.x] = xu[j]; by[threadIdx.x] = yu[j]; }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int n = (M - j0 < 1024) ? (int)(M - j0) : 1024;
      for (int l = 0; l < n; l++) { xx = __fadd_rn(xx, bx[l]); yy = __fadd_rn(yy, by[l]); }   // sammon.c:226-231
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { cx = __fdiv_rn(xx, (float)M); cy = __fdiv_rn(yy, (float)M); }       // xx /= noc
  __syncthreads();
  for (long j = threadIdx.x; j < M; j += 1024) {
    x[j] = __fsub_rn(xu[j], cx);
    y[j] = __fsub_rn(yu[j], cy);
  }
}

cudaError_t k6_pair_dist(const float *d_codes, const unsigned char *d_mask, long M, int D, float *d_dd,
                         int32_t *d_zero_pairs, long cap, unsigned long long *d_nzero, cudaStream_t st) {
  const long nt = (M + PT_T - 1) / PT_T;
  const long nblocks = nt * (nt + 1) / 2;
  if (nblocks > 0x7fffffffL) return cudaErrorInvalidValue;
  if (d_mask)
    k6_pair_dist_kernel<true><<<(unsigned)nblocks, 256, 0, st>>>(d_codes, d_mask, M, D, (int)nt, d_dd, d_zero_pairs,
                                                                 cap, d_nzero);
  else
    k6_pair_dist_kernel<false><<<(unsigned)nblocks, 256, 0, st>>>(d_codes, d_mask, M, D, (int)nt, d_dd, d_zero_pairs,
                                                                  cap, d_nzero);
  return cudaGetLastError();
}

cudaError_t k6_sweep(const float *d_dd, long M, float *d_x, float *d_y, float *d_xu, float *d_yu, int num_sms,
                     cudaStream_t st) {
  (void)num_sms;
  const long blocks = (M + K6_JT - 1) / K6_JT;
  k6_sweep_kernel<<<(unsigned)blocks, K6_THREADS, 0, st>>>(d_dd, M, d_x, d_y, d_xu, d_yu);
  k6_center_kernel<<<1, 1024, 0, st>>>(M, d_xu, d_yu, d_x, d_y);
  return cudaGetLastError();
}

}  // namespace bmu
