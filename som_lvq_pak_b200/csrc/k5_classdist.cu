// k5_classdist.cu -- K5: for every code vector the distance to the nearest LATER code vector of the
// same class, the O(M^2 D) inner loops of min_distances / med_distances (lvq_rout.c:280-492) that
// `mindist` and `balance` run on a codebook.
//
// The reference calls vector_dist_euc(later, earlier) (lvq_pak.c:291-316) for every same-class pair
// and keeps the smallest value per earlier entry.  sqrt is monotone, so the kernel keeps the smallest
// SQUARED sum (exact: fl(fl(a-b)^2) added in component order, components masked in either vector
// skipped) and the host takes the one square root per entry.  Work is tiled like a matrix product:
// a CTA owns a 64 x 64 block of (earlier, later) pairs of the upper triangle, both 64-vector tiles
// are staged through shared memory 32 components at a time, and each thread carries a 4 x 4 block
// of running sums in registers.  The class and order predicate is applied to the finished sums; the
// per-entry minimum goes through shared-memory and then global atomicMin on the bit pattern (the
// sums are non-negative, so unsigned order is float order; NaN and +Inf sort above FLT_MAX and are
// never taken, like `dist < dissf` in the reference).  FP32 issue bound, 3 lane-ops per element.
#include "common.cuh"
#include "pairtile.cuh"
#include "k5_classdist.h"

namespace bmu {

#define K5_T PT_T

template <bool MASKED>
__global__ void __launch_bounds__(256)
k5_class_nearest_kernel(const float *__restrict__ codes, const unsigned char *__restrict__ mask,
                        const int32_t *__restrict__ label, long M, int D, int ntiles,
                        uint32_t *__restrict__ d2bits, uint32_t *__restrict__ flags) {
  __shared__ PairTileSmem<MASKED> ts;
  __shared__ uint32_t smin[K5_T], sflag[K5_T];
  int ti, tj;
  pair_tile_index(blockIdx.x, ntiles, ti, tj);
  const long i0 = (long)ti * K5_T, j0 = (long)tj * K5_T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  if (tid < K5_T) { smin[tid] = 0x7f800000u; sflag[tid] = 0u; }

  float acc[4][4];
  int nmask[4][4];
  pair_tile_sums<MASKED>(codes, mask, M, D, i0, j0, ts, acc, nmask);   // later - earlier, lvq_pak.c:306

#pragma unroll
  for (int r = 0; r < 4; r++) {
    const long gi = i0 + ty * 4 + r;
    if (gi >= M) continue;
    const int li = label[gi];
    uint32_t best = 0x7f800000u, fl = 0u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const long gj = j0 + tx * 4 + q;
      if (gj >= M || gj <= gi || label[gj] != li) continue;
      fl |= 1u;
      if (MASKED && nmask[r][q] == D) { fl |= 2u; continue; }
      const uint32_t bits = __float_as_uint(acc[r][q]);
      if (bits < best) best = bits;
    }
    if (fl) {
      atomicMin(&smin[ty * 4 + r], best);
      atomicOr(&sflag[ty * 4 + r], fl);
    }
  }
  __syncthreads();
  if (tid < K5_T && i0 + tid < M && sflag[tid]) {
    atomicMin(&d2bits[i0 + tid], smin[tid]);
    atomicOr(&flags[i0 + tid], sflag[tid]);
  }
}

cudaError_t k5_class_nearest(const float *d_codes, const unsigned char *d_mask, const int32_t *d_label,
                             long M, int D, uint32_t *d_d2bits, uint32_t *d_flags, cudaStream_t st) {
  const long nt = (M + K5_T - 1) / K5_T;
  const long nblocks = nt * (nt + 1) / 2;
  if (nblocks > 0x7fffffffL) return cudaErrorInvalidValue;
  if (d_mask)
    k5_class_nearest_kernel<true><<<(unsigned)nblocks, 256, 0, st>>>(d_codes, d_mask, d_label, M, D, (int)nt,
                                                                     d_d2bits, d_flags);
  else
    k5_class_nearest_kernel<false><<<(unsigned)nblocks, 256, 0, st>>>(d_codes, d_mask, d_label, M, D, (int)nt,
                                                                      d_d2bits, d_flags);
  return cudaGetLastError();
}

}  // namespace bmu
