// k2_filter.h -- internal interface of K2: tcgen05 GEMM filter + exact FP32 re-rank
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "k1_search.h"

namespace bmu {

struct K2Codebook {
  void *d_ops = nullptr;        // fp16 operand image of the codebook, UMMA tile layout
  float *d_norm = nullptr;      // codebook statistics (scale, maxima) and the mean vector
  float *d_grp = nullptr;       // FP32 codebook regrouped [M/GW][D/4][GW codes][4 comps] for the group re-rank
  size_t grp_bytes = 0;
  size_t ops_bytes = 0;
  int valid = 0;
  int Kp = 0;
};

bool k2_eligible(int path, long M, int D, long N, int k, unsigned cb_flags);
// centred / scaled fp16 operand image, statistics and the regrouped FP32 copy, queued on `st`
cudaError_t k2_prepare_codebook(K2Codebook *c, const float *d_codes, long M, int D, cudaStream_t st);
void k2_codebook_invalidate(K2Codebook *c);
void k2_codebook_free(K2Codebook *c);
// scratch: grow-only device buffer owned by the caller
// device time of [row_prep, gemm, rerank, fallback lists] of the last k2_search call (ms)
cudaError_t k2_last_kernel_ms(float out[4]);
// the same for the call `back` calls ago (0 = last; zeros beyond the ring of K_EV_RING calls)
cudaError_t k2_kernel_ms_history(int back, float out[4]);
cudaError_t k2_search(K2Codebook *c, const K1Args &a, void **scratch, size_t *scratch_bytes,
                      cudaStream_t st);

}  // namespace bmu
