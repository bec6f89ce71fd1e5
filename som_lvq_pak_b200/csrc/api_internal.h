// api_internal.h -- state shared by the translation units that implement include/bmu.h
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/bmu.h"
#include "k2_filter.h"

namespace bmu {

extern char g_err[512];
extern int g_dev, g_sms;
extern size_t g_smem_optin;
extern cudaStream_t g_compute, g_copy, g_out;

int fail(int code, const char *fmt, ...);
int ensure_init();

#define CK(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return bmu::fail(BMU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                       __FILE__, __LINE__);                                                 \
  } while (0)

// grow-only device scratch
struct Scratch {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need);
  void release();
};

}  // namespace bmu

struct bmu_codebook {
  long M;
  int D;
  float *d_codes;      // M x D row-major
  float *d_cT;         // K1 tile layout
  unsigned *d_flags;   // ROW_* bits of the codebook
  unsigned h_flags;
  bmu::K2Codebook k2;  // operands of the tcgen05 filter (built lazily)
  float *d_cq;         // component-major copy for K4 (qerror2), built lazily; nullptr = stale
};
