// pairtile.cuh -- a 64 x 64 tile of exact vector_dist_euc sums (lvq_pak.c:291-316) between code
// vectors i0.. (rows) and j0.. (columns): both 64-vector tiles are staged through shared memory 32
// components at a time as [component][vector] (padded), each of the 256 threads carries a 4 x 4 block
// of running sums (rows ty*4.., columns tx*4..) -- fl(fl(a-b)^2) added in component order, components
// masked in either vector skipped and counted.  Shared by K5 (class distances) and K6 (Sammon).
#pragma once
#include "common.cuh"

namespace bmu {

#define PT_T 64          // pairs tile edge
#define PT_DC 32         // components per stage
#define PT_LD (PT_T + 1) // padded row of the [component][vector] tiles

template <bool MASKED>
struct PairTileSmem {
  float sa[PT_DC * PT_LD], sb[PT_DC * PT_LD];
  unsigned char ma[MASKED ? PT_DC * PT_LD : 1], mb[MASKED ? PT_DC * PT_LD : 1];
};

template <bool MASKED>
__device__ __forceinline__ void pair_tile_sums(const float *__restrict__ codes,
                                               const unsigned char *__restrict__ mask, long M, int D,
                                               long i0, long j0, PairTileSmem<MASKED> &s,
                                               float (&acc)[4][4], int (&nmask)[4][4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) { acc[r][c] = 0.0f; nmask[r][c] = 0; }
  for (int d0 = 0; d0 < D; d0 += PT_DC) {
    __syncthreads();
    // 64 vectors x 32 components per tile; consecutive threads read consecutive components
    for (int e = tid; e < PT_T * PT_DC; e += 256) {
      const int v = e >> 5, c = e & 31;
      const bool cin = d0 + c < D;
      const long gi = i0 + v, gj = j0 + v;
      s.sa[c * PT_LD + v] = (cin && gi < M) ? codes[gi * D + d0 + c] : 0.0f;
      s.sb[c * PT_LD + v] = (cin && gj < M) ? codes[gj * D + d0 + c] : 0.0f;
      if (MASKED) {
        s.ma[c * PT_LD + v] = (cin && gi < M) ? mask[gi * D + d0 + c] : 1;
        s.mb[c * PT_LD + v] = (cin && gj < M) ? mask[gj * D + d0 + c] : 1;
      }
    }
    __syncthreads();
    const int dc = (D - d0 < PT_DC) ? D - d0 : PT_DC;
#pragma unroll 4
    for (int c = 0; c < dc; c++) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; r++) { a[r] = s.sa[c * PT_LD + ty * 4 + r]; b[r] = s.sb[c * PT_LD + tx * 4 + r]; }
      if (MASKED) {
        unsigned char xa[4], xb[4];
#pragma unroll
        for (int r = 0; r < 4; r++) { xa[r] = s.ma[c * PT_LD + ty * 4 + r]; xb[r] = s.mb[c * PT_LD + tx * 4 + r]; }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int q = 0; q < 4; q++) {
            if (xa[r] | xb[q]) nmask[r][q]++;
            else acc[r][q] = sq_acc(acc[r][q], b[q], a[r]);
          }
      } else {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
          for (int q = 0; q < 4; q++) acc[r][q] = sq_acc(acc[r][q], b[q], a[r]);
      }
    }
  }
}

// linear block index -> tile pair (ti, tj) of the upper triangle, tj >= ti
__device__ __forceinline__ void pair_tile_index(int block, int ntiles, int &ti, int &tj) {
  int rem = block;
  ti = 0;
  while (rem >= ntiles - ti) { rem -= ntiles - ti; ti++; }
  tj = ti + rem;
}

}  // namespace bmu
