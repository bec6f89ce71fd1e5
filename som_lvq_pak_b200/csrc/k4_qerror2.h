// k4_qerror2.h -- internal interface of K4: neighbourhood-weighted quantization error (qerror -qetype 1)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bmu {

// component-major copy of the codebook: cq[i * Mp + u], Mp = M rounded up to 32 (zero padded)
long k4_mp(long M);
cudaError_t k4_transpose_codebook(const float *d_codes, long M, int D, float *d_cq, cudaStream_t st);
// per-sample value of bubble_qerror / gaussian_qerror around the winner idx[n] (som_rout.c:734-819);
// out[n] = 0 for rows without a winner
cudaError_t k4_qerror2(const float *d_cq, long M, int D, int xdim, int topol, int neigh, float radius,
                       const float *d_data, const unsigned char *d_mask, long N, const int32_t *d_idx,
                       const int32_t *d_nfound, float *d_out, int num_sms, cudaStream_t st);

}  // namespace bmu
