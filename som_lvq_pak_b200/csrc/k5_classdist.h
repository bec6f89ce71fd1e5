// k5_classdist.h -- internal interface of K5: nearest later same-class code (min_distances / med_distances)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bmu {

// d2bits[i] = bit pattern of the smallest finite squared distance from code i to a code j > i with
//             label[j] == label[i] (0x7f800000 when there is none); the caller fills it with 0x7f800000
// flags[i]  = bit 0: a later code of the same class exists; bit 1: one such pair has every component
//             masked (vector_dist_euc returns -1 for it); the caller zeroes it
cudaError_t k5_class_nearest(const float *d_codes, const unsigned char *d_mask, const int32_t *d_label,
                             long M, int D, uint32_t *d_d2bits, uint32_t *d_flags, cudaStream_t st);

}  // namespace bmu
