// k6_sammon.h -- internal interface of K6: Sammon's mapping (sammon.c:83-262)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bmu {

// d_dd (nullable): full symmetric M x M matrix of vector_dist_euc values (-1 = all components masked);
// d_zero_pairs (nullable): (i, j), i < j, of the pairs at distance exactly 0, in no particular order;
// *d_nzero counts them all (also those beyond cap) and must be zeroed by the caller
cudaError_t k6_pair_dist(const float *d_codes, const unsigned char *d_mask, long M, int D, float *d_dd,
                         int32_t *d_zero_pairs, long cap, unsigned long long *d_nzero, cudaStream_t st);
// one sweep of sammon_iterate: (x, y) -> (x, y), xu / yu scratch of M floats each; two launches
cudaError_t k6_sweep(const float *d_dd, long M, float *d_x, float *d_y, float *d_xu, float *d_yu, int num_sms,
                     cudaStream_t st);

}  // namespace bmu
