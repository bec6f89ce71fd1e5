// api_internal.h -- state shared by the translation units that implement include/bmu.h
//
// All device state lives in ONE DevCtx per GPU.  The single-GPU entry points work on the primary
// context (the device given to bmu_init); the multi-GPU entry points (bmu_multi_*) run one host
// thread per GPU, each bound to its own context through a thread-local pointer, so the kernels'
// launchers (k1_search.cu, k2_filter.cu) never see which GPU they are on.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "../../include/bmu.h"
#include "k1_search.h"
#include "k2_filter.h"

namespace bmu {

// grow-only device scratch
struct Scratch {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need);
  void release();
};

// scratch of one search (row classes, work lists, counters, K2 operand image)
struct SearchScratch {
  Scratch xT, flags, listW, listS, counters, k2, lkeys, ldone, lparts;
};

// pinned host ring that feeds the device from pageable caller memory (bmu_search, host pointers)
struct HostRing;

#define BMU_MAX_GPUS 16
#define BMU_NSLOT 3          // chunks in flight of the host-pointer pipeline

struct DevCtx {
  int dev = -1, sms = 0;
  size_t smem_optin = 0;
  cudaStream_t compute = nullptr, copy = nullptr, out = nullptr;
  // ONE search scratch per context.  Searches may be launched on any stream (bmu_search_dev takes the
  // caller's); every search waits for `ss_done` of the one before and records it again when its last
  // kernel is queued, so that searches of one context never overlap on the shared scratch, the work
  // lists, the counters or the K2 operand image -- whatever streams they were given.
  SearchScratch ss;
  cudaEvent_t ss_done = nullptr;
  int ss_used = 0;
  // host-pointer chunk pipeline: device staging per slot and the events that order the three streams
  Scratch stage_in[BMU_NSLOT], stage_mask[BMU_NSLOT], stage_idx[BMU_NSLOT], stage_diff[BMU_NSLOT], stage_nf[BMU_NSLOT],
      stage_lab[BMU_NSLOT];
  cudaEvent_t ev_in[BMU_NSLOT] = {}, ev_work[BMU_NSLOT] = {}, ev_out[BMU_NSLOT] = {};
  Scratch q2_out;
  HostRing *ring = nullptr;
  // statistics of a sharded search: {sum sqrt(diff)} and {n_found, hist[M], confusion[L*L]}
  Scratch stat_f64, stat_i64, stat_part;
  // per-kernel CUDA events of the last K_EV_RING searches (k1_search.cu / k2_filter.cu)
  cudaEvent_t k1ring[K_EV_RING][5] = {};
  long k1calls = 0;
  cudaEvent_t k2ring[K_EV_RING][5] = {};
  long k2calls = 0;
  cudaStream_t k2aux = nullptr;
  cudaEvent_t k2sub[8] = {}, k2join = nullptr;
  // breakdown of the last search
  const int *last_counters = nullptr;
  long last_rows = 0;
  int last_used_k2 = 0;
  // NCCL communicator of this device (void*: nccl.h stays out of the other translation units)
  void *comm = nullptr;
  int comm_rank = 0, comm_nranks = 1;
};

extern thread_local char g_err[512];
DevCtx *ctx();                 // context of the calling thread (worker binding, else the primary)
void bind_ctx(DevCtx *c);      // worker threads: cudaSetDevice + thread-local binding
DevCtx *ctx_of_device(int dev);
int ctx_open(DevCtx *c, int device);     // streams / events of one device
void ctx_close(DevCtx *c);
void host_ring_free(DevCtx *c);          // bmu_host.cu
void multi_shutdown();                   // bmu_multi.cu

int fail(int code, const char *fmt, ...);
int ensure_init();

// names the single-GPU code was written with
#define g_dev (bmu::ctx()->dev)
#define g_sms (bmu::ctx()->sms)
#define g_smem_optin (bmu::ctx()->smem_optin)
#define g_compute (bmu::ctx()->compute)
#define g_copy (bmu::ctx()->copy)
#define g_out (bmu::ctx()->out)

#define CK(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return bmu::fail(BMU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                       __FILE__, __LINE__);                                                 \
  } while (0)

// bmu_api.cu: one search of device-resident rows on the calling thread's context
int search_dev_impl(bmu_codebook *cb, const float *d_data, const unsigned char *d_mask, long N, int k,
                    int32_t *d_idx, float *d_diff, int32_t *d_nfound, cudaStream_t st);
// bmu_api.cu: accumulate the statistics of a finished search into the context's buffers
int stats_accumulate(DevCtx *c, const int32_t *d_idx, const float *d_diff, const int32_t *d_nfound, long N, int k,
                     long M, double *d_sum, long long *d_nfound_total, long long *d_hist, const int32_t *d_slabel,
                     const int32_t *d_clabel, int L, long long *d_conf, cudaStream_t st);
// bmu_host.cu: the host-pointer chunk pipeline on the calling thread's context
struct HostStats {             // optional reduction fused into the pipeline (device buffers of the context)
  const int32_t *sample_label; // host, N (or nullptr)
  int n_labels;
  int want_hist;
};
int search_host_pipeline(bmu_codebook *cb, const float *data, const unsigned char *mask, long N, int k,
                         int32_t *idx, float *diff, int32_t *nfound, const HostStats *hs);

}  // namespace bmu

struct bmu_codebook {
  long M;
  int D;
  bmu::DevCtx *owner;  // the context (device) this copy lives on
  float *d_codes;      // M x D row-major
  float *d_cT;         // K1 tile layout
  unsigned *d_flags;   // ROW_* bits of the codebook
  unsigned h_flags;
  bmu::K2Codebook k2;  // operands of the tcgen05 filter (built with the codebook)
  float *d_cq;         // component-major copy for K4 (qerror2), built lazily; nullptr = stale
  int32_t *d_label;    // class label per code vector (confusion counts), nullptr = none
};
