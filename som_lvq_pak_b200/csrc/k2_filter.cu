// k2_filter.cu -- K2: tcgen05 GEMM filter + exact FP32 re-rank.  (placeholder: the filter
// path is not enabled yet; k2_eligible() returns false so every search runs through K1.)
#include "common.cuh"
#include "k2_filter.h"

namespace bmu {

bool k2_eligible(int, long, int, long, int, unsigned) { return false; }
void k2_codebook_invalidate(K2Codebook *c) { c->valid = 0; }
void k2_codebook_free(K2Codebook *c) {
  if (c->d_ops) cudaFree(c->d_ops);
  if (c->d_norm) cudaFree(c->d_norm);
  c->d_ops = nullptr; c->d_norm = nullptr; c->valid = 0;
}
cudaError_t k2_search(K2Codebook *, const K1Args &, void **, size_t *, cudaStream_t) {
  return cudaErrorNotSupported;
}

}  // namespace bmu
