// k3_train.h -- internal interface of K3, the persistent online-training kernel
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bmu {

enum K3Mode { K3_SOM_BUBBLE = 0, K3_SOM_GAUSSIAN = 1, K3_LVQ1 = 2, K3_LVQ2 = 3, K3_LVQ3 = 4, K3_OLVQ1 = 5 };

#define K3_THREADS 256
// Pause between publishing a CTA's key and the first poll of the grid exchange (see k3_train.cu): > 0 nanoseconds of
// __nanosleep, < 0 cycles of busy waiting, K3_POLL_DELAY_AUTO = the measured default of the kernel that is launched
#define K3_POLL_DELAY_AUTO (-2147483647)
#define K3_POLL_DELAY_GENERIC 100      // generic kernel, grids of more than 64 CTAs (smaller grids poll at once)
#define K3_POLL_DELAY_FUSED (-475)     // fused large-map kernel: 475 cycles (split polling; -450 is the measured optimum, -400 starts to miss)
#define K3_MAX_GRID 160              // CTA slots the grid exchange polls (5 x 32 lanes)
#define K3_MASK_SENTINEL 0x7fc00b00u   // quiet NaN payload that marks a masked component

struct K3Params {
  float *codes;                  // M x D row-major, updated in place
  const float *data;             // N x D row-major; masked components hold K3_MASK_SENTINEL
  const unsigned char *valid;    // N: 0 = every component masked (step is skipped)
  long N, M;
  int D;
  int mode;
  // SOM
  int xdim, ydim, topol;
  const short *fixed_xy;         // N x 2 or nullptr
  // LVQ
  const int *code_label;         // M
  const int *data_label;         // N
  float win_thr, epsilon, alpha_cap;
  float *unit_alpha;             // M (OLVQ1) or nullptr
  // per-step schedule (device arrays)
  const int *sample;
  const float *talp;
  const float *trad;
  long nsteps;
  // grid-wide exchange
  unsigned long long *slots;     // [2][grid][2], zeroed before every launch
  float *gslice;                 // [grid][D][Us] when the slices do not fit shared memory
  int U, Us;                     // units per CTA and padded row stride
  int slice_in_smem;
  int poll_delay_ns;             // pause between publishing the CTA's key and the first poll of the others' slots
  long long *prof;               // nullptr, or [grid][8] cycle counters of the fused kernel's phases ($BMU_K3_PROF)
};

struct K3Plan {
  int grid, U, Us, slice_in_smem;
  size_t smem_bytes, gslice_floats;
};

K3Plan k3_plan(long M, int D, int num_sms, size_t smem_optin);
cudaError_t k3_launch(const K3Params &p, const K3Plan &plan, bool has_mask, cudaStream_t st);
// data (+mask) -> device copy with sentinels, and the per-row valid flags
cudaError_t k3_encode_mask(float *d_data, const unsigned char *d_mask, unsigned char *d_valid,
                           long N, int D, cudaStream_t st);

}  // namespace bmu
