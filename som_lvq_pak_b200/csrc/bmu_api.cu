// bmu_api.cu -- the C ABI declared in include/bmu.h: contexts, codebooks, the batch winner
// search entry points (host- and device-pointer), per-shard statistics and the pure-C host
// helpers (sample order, schedules).  No CPU fallback: without a usable sm_100 device every
// entry point fails with BMU_ERR_NODEV / BMU_ERR_CUDA.
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bmu.h"
#include "common.cuh"
#include "k1_search.h"
#include "k2_filter.h"
#include "k3_train.h"
#include "k4_qerror2.h"
#include "k5_classdist.h"
#include "k6_sammon.h"
#include "api_internal.h"

using namespace bmu;

// ------------------------------------------------------------------ context
namespace bmu {
char g_err[512] = "";
int g_dev = -1;
int g_sms = 0;
size_t g_smem_optin = 0;
cudaStream_t g_compute = nullptr, g_copy = nullptr, g_out = nullptr;

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int ensure_init() {
  if (g_dev >= 0) return BMU_OK;
  return bmu_init(0);
}
int Scratch::ensure(size_t need) {
  if (need <= bytes) return BMU_OK;
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
  size_t want = need + need / 8;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    if (cudaMalloc(&p, need) != cudaSuccess) {
      cudaGetLastError();
      return fail(BMU_ERR_NOMEM, "cudaMalloc of %zu bytes failed", need);
    }
    want = need;
  }
  bytes = want;
  return BMU_OK;
}
void Scratch::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}
}  // namespace bmu

namespace {

int g_path = BMU_PATH_AUTO;
const int *g_last_counters = nullptr;     // device counters of the last search call
long g_last_rows = 0;
int g_last_used_k2 = 0;

struct SearchScratch {
  Scratch xT, flags, listW, listS, counters, k2;
};
SearchScratch g_ss[2];          // two sets so that chunked host searches can overlap
Scratch g_stage_in[2], g_stage_mask[2], g_stage_idx[2], g_stage_diff[2], g_stage_nf[2];
Scratch g_q2_out;                // per-sample values of bmu_qerror2

}  // namespace

extern "C" {

int bmu_init(int device) {
  if (g_dev == device && g_compute) return BMU_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(BMU_ERR_NODEV, "no CUDA device: %s (this library has no CPU fallback)",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return fail(BMU_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  CK(cudaSetDevice(device));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return fail(BMU_ERR_NODEV, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, p.major, p.minor);
  if (g_compute) bmu_shutdown();
  g_dev = device;
  g_sms = p.multiProcessorCount;
  g_smem_optin = p.sharedMemPerBlockOptin;
  CK(cudaStreamCreateWithFlags(&g_compute, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g_out, cudaStreamNonBlocking));
  return BMU_OK;
}

void bmu_shutdown(void) {
  if (g_dev < 0) return;
  cudaDeviceSynchronize();
  for (int i = 0; i < 2; i++) {
    g_ss[i].xT.release(); g_ss[i].flags.release(); g_ss[i].listW.release();
    g_ss[i].listS.release(); g_ss[i].counters.release(); g_ss[i].k2.release();
    g_stage_in[i].release(); g_stage_mask[i].release(); g_stage_idx[i].release();
    g_stage_diff[i].release(); g_stage_nf[i].release();
  }
  g_q2_out.release();
  if (g_compute) cudaStreamDestroy(g_compute);
  if (g_copy) cudaStreamDestroy(g_copy);
  if (g_out) cudaStreamDestroy(g_out);
  g_compute = g_copy = g_out = nullptr;
  g_dev = -1;
}

const char *bmu_last_error(void) { return g_err; }

int bmu_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin) {
  int rc = ensure_init();
  if (rc) return rc;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, g_dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
  return BMU_OK;
}

int bmu_set_search_path(int path) {
  if (path < BMU_PATH_AUTO || path > BMU_PATH_FILTER) return fail(BMU_ERR_ARG, "bad path %d", path);
  g_path = path;
  return BMU_OK;
}

long bmu_launch_count(void) { return k1_launch_count(); }

int bmu_last_search_kernel_ms(float out[8]) {
  CK(k1_last_kernel_ms(out));
  CK(k2_last_kernel_ms(out + 4));
  return BMU_OK;
}

int bmu_search_kernel_ms_history(int back, float out[8]) {
  CK(k1_kernel_ms_history(back, out));
  CK(k2_kernel_ms_history(back, out + 4));
  return BMU_OK;
}

int bmu_last_search_breakdown(long out[5]) {
  for (int i = 0; i < 5; i++) out[i] = 0;
  if (!g_last_counters) return BMU_OK;
  int h[4];
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, g_last_counters, sizeof(h), cudaMemcpyDeviceToHost));
  out[0] = g_last_rows;
  out[1] = h[0];
  out[2] = h[1];
  out[3] = g_last_used_k2 ? h[2] : 0;
  out[4] = g_last_used_k2 ? h[3] : 0;
  return BMU_OK;
}

// ------------------------------------------------------------------ codebook
static bmu_codebook *codebook_from_dev(const float *d_src, const float *h_src, long M, int D) {
  if (ensure_init()) return nullptr;
  if (M <= 0 || D <= 0 || M > 0x7fffff00L) {
    fail(BMU_ERR_ARG, "bad codebook shape M=%ld D=%d", M, D);
    return nullptr;
  }
  bmu_codebook *cb = (bmu_codebook *)calloc(1, sizeof(bmu_codebook));
  if (!cb) { fail(BMU_ERR_NOMEM, "calloc"); return nullptr; }
  cb->M = M;
  cb->D = D;
  size_t bytes = (size_t)M * D * sizeof(float);
  bool ok = cudaMalloc(&cb->d_codes, bytes) == cudaSuccess &&
            cudaMalloc(&cb->d_cT, k1_cT_floats(M, D) * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&cb->d_flags, sizeof(unsigned)) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    fail(BMU_ERR_NOMEM, "cudaMalloc for a %ld x %d codebook failed", M, D);
    bmu_codebook_destroy(cb);
    return nullptr;
  }
  cudaError_t e = h_src ? cudaMemcpyAsync(cb->d_codes, h_src, bytes, cudaMemcpyHostToDevice, g_compute)
                        : cudaMemcpyAsync(cb->d_codes, d_src, bytes, cudaMemcpyDeviceToDevice, g_compute);
  if (e == cudaSuccess) e = k1_prepare_codebook(cb->d_codes, M, D, cb->d_cT, cb->d_flags, g_compute);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(&cb->h_flags, cb->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, g_compute);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_compute);
  if (e != cudaSuccess) {
    fail(BMU_ERR_CUDA, "codebook upload failed: %s", cudaGetErrorString(e));
    bmu_codebook_destroy(cb);
    return nullptr;
  }
  return cb;
}

bmu_codebook *bmu_codebook_create(const float *codes, long M, int D) {
  if (!codes) { fail(BMU_ERR_ARG, "codes is NULL"); return nullptr; }
  return codebook_from_dev(nullptr, codes, M, D);
}

bmu_codebook *bmu_codebook_create_dev(const float *d_codes, long M, int D) {
  if (!d_codes) { fail(BMU_ERR_ARG, "d_codes is NULL"); return nullptr; }
  return codebook_from_dev(d_codes, nullptr, M, D);
}

int bmu_codebook_update(bmu_codebook *cb, const float *codes) {
  if (!cb || !codes) return fail(BMU_ERR_ARG, "NULL argument");
  CK(cudaMemcpyAsync(cb->d_codes, codes, (size_t)cb->M * cb->D * sizeof(float),
                     cudaMemcpyHostToDevice, g_compute));
  CK(k1_prepare_codebook(cb->d_codes, cb->M, cb->D, cb->d_cT, cb->d_flags, g_compute));
  CK(cudaMemcpyAsync(&cb->h_flags, cb->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, g_compute));
  CK(cudaStreamSynchronize(g_compute));
  k2_codebook_invalidate(&cb->k2);
  if (cb->d_cq) { cudaFree(cb->d_cq); cb->d_cq = nullptr; }
  return BMU_OK;
}

void bmu_codebook_destroy(bmu_codebook *cb) {
  if (!cb) return;
  if (cb->d_codes) cudaFree(cb->d_codes);
  if (cb->d_cT) cudaFree(cb->d_cT);
  if (cb->d_flags) cudaFree(cb->d_flags);
  if (cb->d_cq) cudaFree(cb->d_cq);
  k2_codebook_free(&cb->k2);
  free(cb);
}

// ------------------------------------------------------------------ search
static int search_dev_impl(bmu_codebook *cb, const float *d_data, const unsigned char *d_mask,
                           long N, int k, int32_t *d_idx, float *d_diff, int32_t *d_nfound,
                           cudaStream_t st, SearchScratch &ss) {
  if (!cb || !d_data || !d_idx || !d_diff || !d_nfound) return fail(BMU_ERR_ARG, "NULL argument");
  if (k < 1 || k > BMU_KMAX) return fail(BMU_ERR_ARG, "k=%d outside 1..%d", k, BMU_KMAX);
  if (N < 0 || N > 0x7fffff00L) return fail(BMU_ERR_ARG, "bad N=%ld", N);
  if (N == 0) return BMU_OK;
  const int D = cb->D;
  int rc;
  const bool use_k2 = k2_eligible(g_path, cb->M, D, N, k, cb->h_flags);
  const bool need_tiles = (k == 1) && !cb->h_flags && !use_k2;
  if (need_tiles && (rc = ss.xT.ensure(k1_xT_floats(N, D) * sizeof(float)))) return rc;
  if ((rc = ss.flags.ensure((size_t)N))) return rc;
  if ((rc = ss.listW.ensure((size_t)N * sizeof(int)))) return rc;
  if ((rc = ss.listS.ensure((size_t)N * sizeof(int)))) return rc;
  if ((rc = ss.counters.ensure(16 * sizeof(int)))) return rc;

  K1Args a;
  a.data = d_data; a.mask = d_mask; a.codes = cb->d_codes; a.cT = cb->d_cT;
  a.cb_flags = cb->d_flags; a.N = N; a.M = cb->M; a.D = D; a.k = k;
  a.skip_fast = need_tiles ? 0 : 1;
  a.num_sms = g_sms;
  a.short_list = use_k2 ? 1 : 0;
  a.xT = (float *)ss.xT.p; a.flags = (unsigned char *)ss.flags.p;
  a.listW = (int *)ss.listW.p; a.listS = (int *)ss.listS.p; a.counters = (int *)ss.counters.p;
  a.idx = d_idx; a.diff = d_diff; a.nfound = d_nfound;
  g_last_counters = a.counters;
  g_last_rows = N;
  g_last_used_k2 = use_k2 ? 1 : 0;
  if (use_k2) {
    cudaError_t e = k2_search(&cb->k2, a, &ss.k2.p, &ss.k2.bytes, st);
    if (e != cudaSuccess) return fail(BMU_ERR_CUDA, "k2_search: %s", cudaGetErrorString(e));
    return BMU_OK;
  }
  cudaError_t e = k1_search(a, st);
  if (e != cudaSuccess) return fail(BMU_ERR_CUDA, "k1_search: %s", cudaGetErrorString(e));
  return BMU_OK;
}

int bmu_search_dev(bmu_codebook *cb, const float *d_data, const unsigned char *d_mask, long N,
                   int k, int32_t *d_idx, float *d_diff, int32_t *d_nfound, void *stream) {
  int rc = ensure_init();
  if (rc) return rc;
  return search_dev_impl(cb, d_data, d_mask, N, k, d_idx, d_diff, d_nfound, (cudaStream_t)stream,
                         g_ss[0]);
}

// Host-pointer search: rows are processed in chunks on three streams; chunk c+1 is copied in
// (g_copy) while chunk c is searched (g_compute) and chunk c-1's results drain (g_out).
int bmu_search(bmu_codebook *cb, const float *data, const unsigned char *mask, long N, int k,
               int32_t *idx, float *diff, int32_t *nfound) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!cb || !data || !idx || !diff || !nfound) return fail(BMU_ERR_ARG, "NULL argument");
  if (k < 1 || k > BMU_KMAX) return fail(BMU_ERR_ARG, "k=%d outside 1..%d", k, BMU_KMAX);
  if (N <= 0) return N == 0 ? BMU_OK : fail(BMU_ERR_ARG, "bad N");
  const int D = cb->D;
  // chunk: about 256 MB of input, a multiple of the 128-row tile
  long chunk = (256L << 20) / ((long)D * 4);
  chunk = (chunk / K1_TS) * K1_TS;
  if (chunk < K1_TS) chunk = K1_TS;
  if (chunk > N) chunk = N;
  for (int b = 0; b < 2; b++) {
    if ((rc = g_stage_in[b].ensure((size_t)chunk * D * 4))) return rc;
    if (mask && (rc = g_stage_mask[b].ensure((size_t)chunk * D))) return rc;
    if ((rc = g_stage_idx[b].ensure((size_t)chunk * k * 4))) return rc;
    if ((rc = g_stage_diff[b].ensure((size_t)chunk * k * 4))) return rc;
    if ((rc = g_stage_nf[b].ensure((size_t)chunk * 4))) return rc;
  }
  cudaEvent_t in_done[2], work_done[2], out_done[2];
  for (int b = 0; b < 2; b++) {
    CK(cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&work_done[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming));
  }
  const long nchunks = (N + chunk - 1) / chunk;
  int status = BMU_OK;
  for (long c = 0; c < nchunks && status == BMU_OK; c++) {
    const int b = (int)(c & 1);
    const long n0 = c * chunk, n = (N - n0 < chunk) ? N - n0 : chunk;
    // the staging buffers of slot b are free once chunk c-2 has been searched and drained
    if (c >= 2) CK(cudaStreamWaitEvent(g_copy, work_done[b], 0));
    CK(cudaMemcpyAsync(g_stage_in[b].p, data + n0 * (long)D, (size_t)n * D * 4,
                       cudaMemcpyHostToDevice, g_copy));
    if (mask)
      CK(cudaMemcpyAsync(g_stage_mask[b].p, mask + n0 * (long)D, (size_t)n * D,
                         cudaMemcpyHostToDevice, g_copy));
    CK(cudaEventRecord(in_done[b], g_copy));
    CK(cudaStreamWaitEvent(g_compute, in_done[b], 0));
    if (c >= 2) CK(cudaStreamWaitEvent(g_compute, out_done[b], 0));
    status = search_dev_impl(cb, (const float *)g_stage_in[b].p,
                             mask ? (const unsigned char *)g_stage_mask[b].p : nullptr, n, k,
                             (int32_t *)g_stage_idx[b].p, (float *)g_stage_diff[b].p,
                             (int32_t *)g_stage_nf[b].p, g_compute, g_ss[b]);
    if (status) break;
    CK(cudaEventRecord(work_done[b], g_compute));
    CK(cudaStreamWaitEvent(g_out, work_done[b], 0));
    CK(cudaMemcpyAsync(idx + n0 * (long)k, g_stage_idx[b].p, (size_t)n * k * 4,
                       cudaMemcpyDeviceToHost, g_out));
    CK(cudaMemcpyAsync(diff + n0 * (long)k, g_stage_diff[b].p, (size_t)n * k * 4,
                       cudaMemcpyDeviceToHost, g_out));
    CK(cudaMemcpyAsync(nfound + n0, g_stage_nf[b].p, (size_t)n * 4, cudaMemcpyDeviceToHost, g_out));
    CK(cudaEventRecord(out_done[b], g_out));
  }
  cudaError_t e1 = cudaStreamSynchronize(g_compute), e2 = cudaStreamSynchronize(g_out);
  cudaStreamSynchronize(g_copy);
  for (int b = 0; b < 2; b++) {
    cudaEventDestroy(in_done[b]); cudaEventDestroy(work_done[b]); cudaEventDestroy(out_done[b]);
  }
  if (status) return status;
  if (e1 != cudaSuccess) return fail(BMU_ERR_CUDA, "search failed: %s", cudaGetErrorString(e1));
  if (e2 != cudaSuccess) return fail(BMU_ERR_CUDA, "search copy failed: %s", cudaGetErrorString(e2));
  return BMU_OK;
}

// ------------------------------------------------------------------ qerror -qetype 1
// find_qerror2 (som_rout.c:823-891): winner search (k = 1) followed by the neighbourhood-weighted
// pass K4 over the same resident chunk; the host adds out[] in data order (one float, som_rout.c:872).
int bmu_qerror2(bmu_codebook *cb, int xdim, int ydim, int topol, int neigh, float radius,
                const float *data, const unsigned char *mask, long N, float *out) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!cb || !data || !out) return fail(BMU_ERR_ARG, "NULL argument");
  if (xdim < 1 || ydim < 1 || (long)xdim * ydim != cb->M)
    return fail(BMU_ERR_ARG, "map %d x %d does not match the codebook (%ld units)", xdim, ydim, cb->M);
  if (topol != BMU_TOPOL_HEXA && topol != BMU_TOPOL_RECT) return fail(BMU_ERR_ARG, "bad topology %d", topol);
  if (neigh != BMU_NEIGH_BUBBLE && neigh != BMU_NEIGH_GAUSSIAN) return fail(BMU_ERR_ARG, "bad neighbourhood %d", neigh);
  if (N <= 0) return N == 0 ? BMU_OK : fail(BMU_ERR_ARG, "bad N");
  const int D = cb->D;
  if (!cb->d_cq) {
    if (cudaMalloc((void **)&cb->d_cq, (size_t)k4_mp(cb->M) * D * sizeof(float)) != cudaSuccess) {
      cb->d_cq = nullptr;
      return fail(BMU_ERR_NOMEM, "cudaMalloc of the component-major codebook failed");
    }
    CK(k4_transpose_codebook(cb->d_codes, cb->M, D, cb->d_cq, g_compute));
    k1_count_launch(1);
  }
  long chunk = (256L << 20) / ((long)D * 4);
  if (chunk < 1) chunk = 1;
  if (chunk > N) chunk = N;
  if ((rc = g_stage_in[0].ensure((size_t)chunk * D * 4))) return rc;
  if (mask && (rc = g_stage_mask[0].ensure((size_t)chunk * D))) return rc;
  if ((rc = g_stage_idx[0].ensure((size_t)chunk * 4))) return rc;
  if ((rc = g_stage_diff[0].ensure((size_t)chunk * 4))) return rc;
  if ((rc = g_stage_nf[0].ensure((size_t)chunk * 4))) return rc;
  if ((rc = g_q2_out.ensure((size_t)chunk * 4))) return rc;
  for (long n0 = 0; n0 < N; n0 += chunk) {
    const long n = (N - n0 < chunk) ? N - n0 : chunk;
    CK(cudaMemcpyAsync(g_stage_in[0].p, data + n0 * (long)D, (size_t)n * D * 4, cudaMemcpyHostToDevice, g_compute));
    if (mask)
      CK(cudaMemcpyAsync(g_stage_mask[0].p, mask + n0 * (long)D, (size_t)n * D, cudaMemcpyHostToDevice, g_compute));
    const unsigned char *d_mask = mask ? (const unsigned char *)g_stage_mask[0].p : nullptr;
    if ((rc = search_dev_impl(cb, (const float *)g_stage_in[0].p, d_mask, n, 1, (int32_t *)g_stage_idx[0].p,
                              (float *)g_stage_diff[0].p, (int32_t *)g_stage_nf[0].p, g_compute, g_ss[0])))
      return rc;
    CK(k4_qerror2(cb->d_cq, cb->M, D, xdim, topol, neigh, radius, (const float *)g_stage_in[0].p, d_mask, n,
                  (const int32_t *)g_stage_idx[0].p, (const int32_t *)g_stage_nf[0].p, (float *)g_q2_out.p,
                  g_sms, g_compute));
    k1_count_launch(1);
    CK(cudaMemcpyAsync(out + n0, g_q2_out.p, (size_t)n * 4, cudaMemcpyDeviceToHost, g_compute));
    CK(cudaStreamSynchronize(g_compute));
  }
  return BMU_OK;
}

// ------------------------------------------------------------------ class distances
// The pair loops of min_distances / med_distances (lvq_rout.c:280-492): dist[i] = dissf of entry i,
// the smallest vector_dist_euc(later, i) over later entries of the same class (first label).
int bmu_class_nearest(const float *codes, const unsigned char *mask, const int32_t *label, long M, int D,
                      float *dist, int32_t *found) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!codes || !label || !dist || !found) return fail(BMU_ERR_ARG, "NULL argument");
  if (M < 0 || D < 1) return fail(BMU_ERR_ARG, "bad M or D");
  if (M == 0) return BMU_OK;
  float *d_codes = nullptr;
  unsigned char *d_mask = nullptr;
  int32_t *d_label = nullptr;
  uint32_t *d_out = nullptr;           // [0, M) squared-distance bits, [M, 2M) flags
  uint32_t *h_out = (uint32_t *)malloc((size_t)M * 8);
  cudaError_t e = cudaSuccess;
  if (!h_out) return fail(BMU_ERR_NOMEM, "out of host memory");
  for (long i = 0; i < M; i++) { h_out[i] = 0x7f800000u; h_out[M + i] = 0u; }
  if ((e = cudaMalloc((void **)&d_codes, (size_t)M * D * 4)) == cudaSuccess &&
      (e = cudaMalloc((void **)&d_label, (size_t)M * 4)) == cudaSuccess &&
      (e = cudaMalloc((void **)&d_out, (size_t)M * 8)) == cudaSuccess &&
      (!mask || (e = cudaMalloc((void **)&d_mask, (size_t)M * D)) == cudaSuccess) &&
      (e = cudaMemcpyAsync(d_codes, codes, (size_t)M * D * 4, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess &&
      (e = cudaMemcpyAsync(d_label, label, (size_t)M * 4, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess &&
      (e = cudaMemcpyAsync(d_out, h_out, (size_t)M * 8, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess &&
      (!mask || (e = cudaMemcpyAsync(d_mask, mask, (size_t)M * D, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess) &&
      (e = k5_class_nearest(d_codes, d_mask, d_label, M, D, d_out, d_out + M, g_compute)) == cudaSuccess &&
      (e = cudaMemcpyAsync(h_out, d_out, (size_t)M * 8, cudaMemcpyDeviceToHost, g_compute)) == cudaSuccess)
    e = cudaStreamSynchronize(g_compute);
  cudaFree(d_codes); cudaFree(d_mask); cudaFree(d_label); cudaFree(d_out);
  if (e != cudaSuccess) {
    free(h_out);
    cudaGetLastError();
    return fail(BMU_ERR_CUDA, "bmu_class_nearest: %s", cudaGetErrorString(e));
  }
  k1_count_launch(1);
  for (long i = 0; i < M; i++) {
    const uint32_t fl = h_out[M + i], bits = h_out[i];
    float d2;
    memcpy(&d2, &bits, 4);
    found[i] = (fl & 1u) ? 1 : 0;
    if (fl & 2u) dist[i] = -1.0f;                                  // an all-masked pair: -1 < anything
    else if (!(fl & 1u) || bits == 0x7f800000u) dist[i] = FLT_MAX; // dissf never lowered
    else dist[i] = (float)sqrt((double)d2);                        // lvq_pak.c:315
  }
  free(h_out);
  return BMU_OK;
}

// ------------------------------------------------------------------ Sammon's mapping
// remove_identicals (sammon.c:83-127) needs the pairs at distance exactly 0; they are returned sorted
// by (i, j) so that the caller can replay the reference's removal walk.
static int cmp_pair(const void *a, const void *b) {
  const int32_t *p = (const int32_t *)a, *q = (const int32_t *)b;
  if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1;
  return p[1] < q[1] ? -1 : (p[1] > q[1] ? 1 : 0);
}

int bmu_identical_pairs(const float *codes, const unsigned char *mask, long M, int D, int32_t *pairs, long cap,
                        long *npairs) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!codes || !npairs || (cap > 0 && !pairs)) return fail(BMU_ERR_ARG, "NULL argument");
  if (M < 0 || D < 1 || cap < 0) return fail(BMU_ERR_ARG, "bad M, D or cap");
  *npairs = 0;
  if (M < 2) return BMU_OK;
  float *d_codes = nullptr;
  unsigned char *d_mask = nullptr;
  int32_t *d_pairs = nullptr;
  unsigned long long *d_n = nullptr, n = 0;
  cudaError_t e;
  if ((e = cudaMalloc((void **)&d_codes, (size_t)M * D * 4)) == cudaSuccess &&
      (e = cudaMalloc((void **)&d_pairs, (size_t)(cap > 0 ? cap : 1) * 8)) == cudaSuccess &&
      (e = cudaMalloc((void **)&d_n, 8)) == cudaSuccess &&
      (!mask || (e = cudaMalloc((void **)&d_mask, (size_t)M * D)) == cudaSuccess) &&
      (e = cudaMemcpyAsync(d_codes, codes, (size_t)M * D * 4, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess &&
      (!mask || (e = cudaMemcpyAsync(d_mask, mask, (size_t)M * D, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess) &&
      (e = cudaMemsetAsync(d_n, 0, 8, g_compute)) == cudaSuccess &&
      (e = k6_pair_dist(d_codes, d_mask, M, D, nullptr, d_pairs, cap, d_n, g_compute)) == cudaSuccess &&
      (e = cudaMemcpyAsync(&n, d_n, 8, cudaMemcpyDeviceToHost, g_compute)) == cudaSuccess &&
      (e = cudaStreamSynchronize(g_compute)) == cudaSuccess) {
    const long got = (long)n < cap ? (long)n : cap;
    if (got > 0) e = cudaMemcpy(pairs, d_pairs, (size_t)got * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && got > 0) qsort(pairs, (size_t)got, 8, cmp_pair);
  }
  cudaFree(d_codes); cudaFree(d_mask); cudaFree(d_pairs); cudaFree(d_n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BMU_ERR_CUDA, "bmu_identical_pairs: %s", cudaGetErrorString(e));
  }
  k1_count_launch(1);
  *npairs = (long)n;
  if ((long)n > cap) return fail(BMU_ERR_ARG, "%ld identical pairs, room for %ld", (long)n, cap);
  return BMU_OK;
}

// sammon_iterate (sammon.c:129-262): `length` sweeps from the caller's initial (x, y) (sammon.c:159-162).
// err (nullable, `length` values): the mapping error the reference prints per sweep at -v 2
// (sammon.c:240-254), sequential float sums over all pairs, computed on the host from the downloaded
// positions (order dependent; costs a synchronisation per sweep, so leave it NULL unless needed).
int bmu_sammon(const float *codes, const unsigned char *mask, long M, int D, long length, float *x, float *y,
               float *err) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!codes || !x || !y) return fail(BMU_ERR_ARG, "NULL argument");
  if (M < 1 || D < 1 || length < 0) return fail(BMU_ERR_ARG, "bad M, D or length");
  float *d_codes = nullptr, *d_dd = nullptr, *d_xy = nullptr, *h_dd = nullptr;
  unsigned char *d_mask = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc((void **)&d_codes, (size_t)M * D * 4)) == cudaSuccess &&
      (e = cudaMalloc((void **)&d_dd, (size_t)M * M * 4)) == cudaSuccess &&
      (e = cudaMalloc((void **)&d_xy, (size_t)M * 16)) == cudaSuccess &&
      (!mask || (e = cudaMalloc((void **)&d_mask, (size_t)M * D)) == cudaSuccess) &&
      (e = cudaMemcpyAsync(d_codes, codes, (size_t)M * D * 4, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess &&
      (!mask || (e = cudaMemcpyAsync(d_mask, mask, (size_t)M * D, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess) &&
      (e = cudaMemcpyAsync(d_xy, x, (size_t)M * 4, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess &&
      (e = cudaMemcpyAsync(d_xy + M, y, (size_t)M * 4, cudaMemcpyHostToDevice, g_compute)) == cudaSuccess &&
      (e = k6_pair_dist(d_codes, d_mask, M, D, d_dd, nullptr, 0, nullptr, g_compute)) == cudaSuccess) {
    k1_count_launch(1);
    if (err && length > 0) {
      h_dd = (float *)malloc((size_t)M * M * 4);
      if (!h_dd) e = cudaErrorMemoryAllocation;
      else e = cudaMemcpyAsync(h_dd, d_dd, (size_t)M * M * 4, cudaMemcpyDeviceToHost, g_compute);
    }
    for (long it = 0; it < length && e == cudaSuccess; it++) {
      e = k6_sweep(d_dd, M, d_xy, d_xy + M, d_xy + 2 * M, d_xy + 3 * M, g_sms, g_compute);
      k1_count_launch(2);
      if (err && e == cudaSuccess) {
        if ((e = cudaMemcpyAsync(x, d_xy, (size_t)M * 4, cudaMemcpyDeviceToHost, g_compute)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(y, d_xy + M, (size_t)M * 4, cudaMemcpyDeviceToHost, g_compute)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(g_compute)) != cudaSuccess) break;
        float ee = 0.0f, tot = 0.0f;
        for (long j = 1; j < M; j++)
          for (long k = 0; k < j; k++) {                             // sammon.c:243-252
            const float d = h_dd[j * M + k];
            tot += d;
            const float xd = x[j] - x[k], yd = y[j] - y[k];
            const float df = d - (float)sqrt((double)xd * xd + yd * yd);
            ee += (df * df / d);
          }
        err[it] = ee / tot;
      }
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(x, d_xy, (size_t)M * 4, cudaMemcpyDeviceToHost, g_compute);
    if (e == cudaSuccess) e = cudaMemcpyAsync(y, d_xy + M, (size_t)M * 4, cudaMemcpyDeviceToHost, g_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_compute);
  }
  free(h_dd);
  cudaFree(d_codes); cudaFree(d_mask); cudaFree(d_dd); cudaFree(d_xy);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? BMU_ERR_NOMEM : BMU_ERR_CUDA, "bmu_sammon: %s", cudaGetErrorString(e));
  }
  return BMU_OK;
}

// ------------------------------------------------------------------ statistics
__global__ void stats_kernel(const int32_t *__restrict__ idx, const float *__restrict__ diff,
                             const int32_t *__restrict__ nfound, long N, int k, long M,
                             double *__restrict__ stats, unsigned long long *__restrict__ hist,
                             const int32_t *__restrict__ slabel, const int32_t *__restrict__ clabel,
                             int L, unsigned long long *__restrict__ conf) {
  double s = 0.0;
  unsigned long long cnt = 0;
  for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < N; n += (long)gridDim.x * blockDim.x) {
    const int j = idx[n * k];
    if (nfound[n] == 0 || j < 0) continue;
    s += sqrt((double)diff[n * k]);
    cnt++;
    if (hist && j < M) atomicAdd(&hist[j], 1ull);
    if (conf) {
      const int a = slabel[n], b = clabel[j];
      if (a >= 0 && a < L && b >= 0 && b < L) atomicAdd(&conf[(long)a * L + b], 1ull);
    }
  }
  for (int off = 16; off >= 1; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&stats[0], s);
    atomicAdd(&stats[1], (double)cnt);
  }
}

int bmu_search_stats_dev(const int32_t *d_idx, const float *d_diff, const int32_t *d_nfound,
                         long N, int k, long M, double *d_stats, long long *d_hist,
                         const int32_t *d_sample_label, const int32_t *d_code_label,
                         int n_labels, long long *d_confusion, void *stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!d_idx || !d_diff || !d_nfound || !d_stats) return fail(BMU_ERR_ARG, "NULL argument");
  if (d_confusion && (!d_sample_label || !d_code_label || n_labels <= 0))
    return fail(BMU_ERR_ARG, "confusion counts need labels");
  if (N <= 0) return BMU_OK;
  int grid = g_sms * 4;
  stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      d_idx, d_diff, d_nfound, N, k, M, d_stats, (unsigned long long *)d_hist, d_sample_label,
      d_code_label, n_labels, (unsigned long long *)d_confusion);
  k1_count_launch(1);
  CK(cudaGetLastError());
  return BMU_OK;
}

// ------------------------------------------------------------------ host helpers
// Sample order of `-rand seed`: the reference's LCG (lvq_pak.c:459-473) drives one pass of
// swaps (datafile.c:1169-1175).
void bmu_rand_order(long n, int seed, int32_t *order) {
  unsigned long state = (unsigned long)seed;
  for (long i = 0; i < n; i++) order[i] = (int32_t)i;
  for (long i = 0; i < n; i++) {
    state = (state * 23UL) % 100000001UL;
    long j = (long)(int)(state % 32767UL) % n;
    int32_t t = order[i];
    order[i] = order[j];
    order[j] = t;
  }
}

// som_rout.c:97-152.  The maximum starts at FLT_MIN (the smallest POSITIVE float), a quirk kept on
// purpose; the value expression mixes float and double exactly as the reference's does:
//   mival + (maval - mival) * ((float) orand() / 32768.0)
void bmu_randinit_codes(const float *data, const unsigned char *mask, long N, int D, long M, int seed,
                        float *codes) {
  float *mx = (float *)malloc(sizeof(float) * 2 * (size_t)D), *mn = mx + D;
  long *cnt = (long *)calloc((size_t)D, sizeof(long));
  unsigned long state = (unsigned long)seed;
  if (!mx || !cnt) { free(mx); free(cnt); return; }
  for (int i = 0; i < D; i++) { mx[i] = FLT_MIN; mn[i] = FLT_MAX; }
  for (long n = 0; n < N; n++)
    for (int i = 0; i < D; i++) {
      if (mask && mask[n * (long)D + i]) continue;
      const float v = data[n * (long)D + i];
      cnt[i]++;
      if (mx[i] < v) mx[i] = v;
      if (mn[i] > v) mn[i] = v;
    }
  for (long u = 0; u < M; u++)
    for (int i = 0; i < D; i++) {
      if (cnt[i] > 0) {
        state = (state * 23UL) % 100000001UL;
        const long r = (long)(int)(state % 32767UL);
        codes[u * (long)D + i] = (float)((double)mn[i] + (double)(mx[i] - mn[i]) * ((double)(float)r / 32768.0));
      } else {
        codes[u * (long)D + i] = 0.0f;
      }
    }
  free(mx);
  free(cnt);
}

static float alpha_at(long le, long length, float alpha, int alpha_type) {
  if (alpha_type == BMU_ALPHA_INVERSE_T) {           // lvq_pak.c:914-921
    float c = (float)length / 100.0f;
    return alpha * c / (c + (float)le);
  }
  return alpha * (float)(length - le) / (float)length;   // lvq_pak.c:903-906
}

void bmu_som_schedule(long le0, long le1, long length, float alpha, float radius,
                      int alpha_type, long N, const int32_t *order, const int16_t *weight,
                      int32_t *sample, float *talp, float *trad) {
  for (long le = le0; le < le1; le++) {
    const long pos = le % N;                                // cyclic list order, som_rout.c:602-610
    const long s = order ? order[pos] : pos;
    float a = alpha_at(le, length, alpha, alpha_type);
    if (weight && weight[s] > 0)                            // som_rout.c:622-624
      a = (float)(1.0 - (double)(float)pow(1.0 - (double)a, (double)(float)weight[s]));
    sample[le - le0] = (int32_t)s;
    talp[le - le0] = a;
    // som_rout.c:615, a double expression rounded once on assignment
    trad[le - le0] = (float)(1.0 + ((double)radius - 1.0) * (double)(float)(length - le) /
                                       (double)(float)length);
  }
}

void bmu_lvq_schedule(long le0, long le1, long length, float alpha, int alpha_type, long N,
                      const int32_t *order, int32_t *sample, float *talp) {
  for (long le = le0; le < le1; le++) {
    const long pos = le % N;
    sample[le - le0] = (int32_t)(order ? order[pos] : pos);
    if (talp) talp[le - le0] = alpha_at(le, length, alpha, alpha_type);
  }
}

float bmu_replay_qerror(const float *diff, const int32_t *nfound, long N, int k) {
  float q = 0.0f;
  for (long n = 0; n < N; n++) {
    if (nfound[n] == 0) continue;
    q = (float)((double)q + sqrt((double)diff[n * (long)k]));
  }
  return q;
}

}  // extern "C"
