// k1_search.h -- internal interface of the exact winner-search kernels (k1_search.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define K1_TS 128       // samples per CTA tile
#define K1_TC 128       // code vectors per CTA tile
#define K1_DC 64        // components per shared-memory stage
#define BMU_KMAX_ 16

namespace bmu {

struct K1Args {
  const float *data;            // N x D row-major (device)
  const unsigned char *mask;    // N x D or nullptr (device)
  const float *codes;           // M x D row-major (device)
  const float *cT;              // tile layout of the codebook (k1_prepare_codebook)
  const unsigned *cb_flags;     // device: ROW_* bits of the whole codebook
  long N, M;
  int D, k;
  int skip_fast;                // host copy of (*cb_flags != 0)
  int num_sms;
  int short_list;               // listW is expected to be short (K2 certificate failures): 8 rows per warp
  // scratch
  float *xT;                    // k1_xT_floats(N, D) floats (only touched when k == 1)
  unsigned char *flags;         // N
  int *listW, *listS;           // N each
  int *counters;                // 4 ints
  unsigned long long *lkeys;    // k1_list8_kernel: N keys, all ones between launches (nullptr: not available)
  int *ldone;                   // k1_list8 / k1_listk: N / 4 + 1 arrival counters, zero between launches
  unsigned long long *lparts;   // k1_listk_kernel: k1_listk_parts_bytes() of per-slice keys (nullptr: not available)
  // outputs
  int32_t *idx;
  float *diff;
  int32_t *nfound;
};

size_t k1_cT_floats(long M, int D);
size_t k1_xT_floats(long N, int D);
size_t k1_listk_parts_bytes();
cudaError_t k1_prepare_codebook(const float *d_codes, long M, int D, float *d_cT,
                                unsigned *d_cb_flags, cudaStream_t st);
cudaError_t k1_search(const K1Args &a, cudaStream_t st);
cudaError_t k1_run_warp_list(const K1Args &a, cudaStream_t st);
cudaError_t k1_run_seq_list(const K1Args &a, cudaStream_t st);
cudaError_t k1_run_lists(const K1Args &a, cudaStream_t st);
// device time of [data_prep, k1_fast, k1_warp, k1_seq] of the last k1_search call (ms)
#define K_EV_RING 32
cudaError_t k1_last_kernel_ms(float out[4]);
// the same for the call `back` calls ago (0 = last; zeros beyond the ring of K_EV_RING calls)
cudaError_t k1_kernel_ms_history(int back, float out[4]);
long k1_launch_count();
void k1_count_launch(int n);

}  // namespace bmu
