// k4_qerror2.cu -- K4: the second pass of `qerror -qetype 1` (find_qerror2, som_rout.c:823-891).
//
// For every sample with winner (bx, by) the reference walks ALL map units in index order and
// accumulates, in one float,   bubble:   d*d           for units with mapdist <= radius
//                              gaussian: (alp*d)*d     alp = (float)exp(-dd*dd / (2 r r))
// where d = vector_dist_euc(unit, sample) = (float)sqrt((double)sum) (lvq_pak.c:291-316; components
// masked in the sample are skipped).  d*d is NOT the squared sum again (sqrt rounds), and the
// float accumulation is order dependent, so the kernel reproduces both: one warp per sample,
// lanes over 32 consecutive units (coalesced reads of a component-major codebook, the sample
// broadcast from L1), exact per-unit sums in component order, then lane 0 adds the 32 terms in
// unit order.  Units outside a bubble contribute +0.0f, which leaves a non-negative float sum
// unchanged bit for bit.  FP32 issue bound: 3*M*D lane-ops per sample, like K1.
#include "common.cuh"
#include "lattice.cuh"
#include "k4_qerror2.h"

namespace bmu {

long k4_mp(long M) { return (M + 31) / 32 * 32; }

__global__ void k4_transpose_kernel(const float *__restrict__ codes, long M, int D, long Mp,
                                    float *__restrict__ cq) {
  __shared__ float tile[32][33];
  const long u0 = (long)blockIdx.x * 32;
  const int i0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;           // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const long u = u0 + r;
    const int i = i0 + tx;
    tile[r][tx] = (u < M && i < D) ? codes[u * D + i] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r;
    const long u = u0 + tx;
    if (i < D && u < Mp) cq[(long)i * Mp + u] = tile[tx][r];
  }
}

cudaError_t k4_transpose_codebook(const float *d_codes, long M, int D, float *d_cq, cudaStream_t st) {
  const long Mp = k4_mp(M);
  dim3 grid((unsigned)(Mp / 32), (unsigned)((D + 31) / 32)), block(32, 8);
  k4_transpose_kernel<<<grid, block, 0, st>>>(d_codes, M, D, Mp, d_cq);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
k4_qerror2_kernel(const float *__restrict__ cq, long M, long Mp, int D, int xdim, int topol, int neigh,
                  float radius, const float *__restrict__ data, const unsigned char *__restrict__ mask,
                  long N, const int32_t *__restrict__ idx, const int32_t *__restrict__ nfound,
                  float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const bool gaussian = neigh == 2;                        // NEIGH_GAUSSIAN
  for (long n = warp0; n < N; n += nwarps) {
    const int w = idx[n];
    if (nfound[n] == 0 || w < 0) {                         // find_winner returned 0: sample ignored (som_rout.c:867)
      if (lane == 0) out[n] = 0.0f;
      continue;
    }
    const int bx = w % xdim, by = w / xdim;
    const float *x = data + n * (long)D;
    const unsigned char *mk = mask ? mask + n * (long)D : nullptr;
    float q = 0.0f;
    for (long u0 = 0; u0 < M; u0 += 32) {
      const long u = u0 + lane;
      float term = 0.0f;
      bool use = false;
      float dd = 0.0f;
      if (u < M) {
        const int tx = (int)(u % xdim), ty = (int)(u / xdim);
        dd = topol == 4 ? rect_dist_dev(bx, by, tx, ty) : hexa_dist_dev(bx, by, tx, ty);
        use = gaussian || dd <= radius;
      }
      if (!__any_sync(0xffffffffu, use)) continue;         // bubble: whole block outside the radius
      if (use) {
        const float *c = cq + u;
        float acc = 0.0f;
        int masked = 0;
        if (mk) {
          for (int i = 0; i < D; i++) {
            if (mk[i]) masked++;
            else acc = sq_acc(acc, c[(long)i * Mp], x[i]);
          }
        } else {
#pragma unroll 4
          for (int i = 0; i < D; i++) acc = sq_acc(acc, c[(long)i * Mp], __ldg(x + i));
        }
        const float d = masked == D ? -1.0f : (float)__dsqrt_rn((double)acc);
        if (gaussian) {
          // som_rout.c:802-806: alp = exp((double)(-dd*dd / (2.0*radius*radius))), qerror += alp*d*d
          const float num = __fmul_rn(-dd, dd);
          const double den = __dmul_rn(__dmul_rn(2.0, (double)radius), (double)radius);
          const float alp = (float)exp(__ddiv_rn((double)num, den));
          term = __fmul_rn(__fmul_rn(alp, d), d);
        } else {
          term = __fmul_rn(d, d);
        }
      }
      // the reference's accumulation order: unit index ascending, one float
#pragma unroll
      for (int l = 0; l < 32; l++) {
        const float t = __shfl_sync(0xffffffffu, term, l);
        q = __fadd_rn(q, t);
      }
    }
    if (lane == 0) out[n] = q;
  }
}

cudaError_t k4_qerror2(const float *d_cq, long M, int D, int xdim, int topol, int neigh, float radius,
                       const float *d_data, const unsigned char *d_mask, long N, const int32_t *d_idx,
                       const int32_t *d_nfound, float *d_out, int num_sms, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  long blocks = (N + 7) / 8;
  const long cap = (long)num_sms * 8;
  if (blocks > cap) blocks = cap;
  k4_qerror2_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_cq, M, k4_mp(M), D, xdim, topol, neigh, radius, d_data,
                                                     d_mask, N, d_idx, d_nfound, d_out);
  return cudaGetLastError();
}

}  // namespace bmu
