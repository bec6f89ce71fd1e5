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
//   k6_sweep_kernel       one warp per point j: lanes compute the four terms of 32 consecutive k in
//                         parallel, park them in shared memory, and the running sums are then advanced
//                         in k order (a float add is not associative, so this chain cannot be a tree);
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

#define K6_WARPS 8

__global__ void __launch_bounds__(K6_WARPS * 32)
k6_sweep_kernel(const float *__restrict__ dd, long M, const float *__restrict__ x, const float *__restrict__ y,
                float *__restrict__ xu, float *__restrict__ yu) {
  __shared__ float s1x[K6_WARPS][32], s1y[K6_WARPS][32];
  __shared__ double s2x[K6_WARPS][32], s2y[K6_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long warp0 = (long)blockIdx.x * K6_WARPS + w, nwarps = (long)gridDim.x * K6_WARPS;
  for (long j = warp0; j < M; j += nwarps) {
    const float xj = x[j], yj = y[j];
    const float *row = dd + j * M;
    float e1x = 0.0f, e1y = 0.0f, e2x = 0.0f, e2y = 0.0f;
    for (long k0 = 0; k0 < M; k0 += 32) {
      const long k = k0 + lane;
      if (k < M && k != j) {
        const float xd = __fsub_rn(xj, x[k]), yd = __fsub_rn(yj, y[k]);
        // (float) sqrt((double) xd * xd + yd * yd): double product + float product, added in double
        const float dpj = (float)sqrt(__dadd_rn(__dmul_rn((double)xd, (double)xd), (double)__fmul_rn(yd, yd)));
        const float dt = row[k];
        const float dq = __fsub_rn(dt, dpj), dr = __fmul_rn(dt, dpj);
        s1x[w][lane] = __fdiv_rn(__fmul_rn(xd, dq), dr);                       // xd * dq / dr
        s1y[w][lane] = __fdiv_rn(__fmul_rn(yd, dq), dr);
        const double u = __dadd_rn(1.0, (double)__fdiv_rn(dq, dpj));           // 1.0 + dq / dpj
        // (dq - xd * xd * (1.0 + dq / dpj) / dpj) / dr
        s2x[w][lane] = __ddiv_rn(__dsub_rn((double)dq, __ddiv_rn(__dmul_rn((double)__fmul_rn(xd, xd), u), (double)dpj)),
                                 (double)dr);
        s2y[w][lane] = __ddiv_rn(__dsub_rn((double)dq, __ddiv_rn(__dmul_rn((double)__fmul_rn(yd, yd), u), (double)dpj)),
                                 (double)dr);
      }
      __syncwarp();
      const int n = (M - k0 < 32) ? (int)(M - k0) : 32;
      const int skip = (j >= k0 && j < k0 + 32) ? (int)(j - k0) : -1;
      // every lane advances the same chain (uniform, broadcast reads); lane 0's copy is stored
#pragma unroll 8
      for (int l = 0; l < n; l++) {
        if (l == skip) continue;
        e1x = __fadd_rn(e1x, s1x[w][l]);
        e1y = __fadd_rn(e1y, s1y[w][l]);
        e2x = (float)__dadd_rn((double)e2x, s2x[w][l]);
        e2y = (float)__dadd_rn((double)e2y, s2y[w][l]);
      }
      __syncwarp();
    }
    if (lane == 0) {
      // x[j] + MAGIC * e1x / fabs(e2x): all in double, rounded on the store (sammon.c:220-221)
      xu[j] = (float)__dadd_rn((double)xj, __ddiv_rn(__dmul_rn(0.2, (double)e1x), fabs((double)e2x)));
      yu[j] = (float)__dadd_rn((double)yj, __ddiv_rn(__dmul_rn(0.2, (double)e1y), fabs((double)e2y)));
    }
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
  long blocks = (M + K6_WARPS - 1) / K6_WARPS;
  const long cap = (long)num_sms * 8;
  if (blocks > cap) blocks = cap;
  k6_sweep_kernel<<<(unsigned)blocks, K6_WARPS * 32, 0, st>>>(d_dd, M, d_x, d_y, d_xu, d_yu);
  k6_center_kernel<<<1, 1024, 0, st>>>(M, d_xu, d_yu, d_x, d_y);
  return cudaGetLastError();
}

}  // namespace bmu
