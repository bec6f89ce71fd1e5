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

#include <atomic>

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
thread_local char g_err[512] = "";
static DevCtx g_primary_ctx;                 // the device of bmu_init (single-GPU entry points)
static thread_local DevCtx *t_ctx = nullptr; // worker threads of the multi-GPU path bind their own
static std::atomic<long> g_launches{0};

DevCtx *ctx() { return t_ctx ? t_ctx : &g_primary_ctx; }
void bind_ctx(DevCtx *c) {
  t_ctx = c;
  if (c && c->dev >= 0) cudaSetDevice(c->dev);
}
long k1_launch_count() { return g_launches.load(); }
void k1_count_launch(int n) { g_launches += n; }

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int ensure_init() {
  DevCtx *c = ctx();
  if (c->dev < 0) return bmu_init(0);
  // the calling thread may have been left on another device (the multi-GPU entry points, the host's own
  // CUDA code): the streams and buffers of this context belong to c->dev
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != c->dev)
    if (cudaSetDevice(c->dev) != cudaSuccess) return fail(BMU_ERR_CUDA, "cudaSetDevice(%d) failed", c->dev);
  return BMU_OK;
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

int ctx_open(DevCtx *c, int device) {
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
  if (c->compute) ctx_close(c);
  c->dev = device;
  c->sms = p.multiProcessorCount;
  c->smem_optin = p.sharedMemPerBlockOptin;
  {
    // the trainers allocate from the stream-ordered pool (bmu_train_api.cu): keep up to 1 GiB of freed blocks in it
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = 1ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  CK(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->out, cudaStreamNonBlocking));
  // every event the entry points use is created once here (not per call)
  CK(cudaEventCreateWithFlags(&c->ss_done, cudaEventDisableTiming));
  for (int b = 0; b < BMU_NSLOT; b++) {
    CK(cudaEventCreateWithFlags(&c->ev_in[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_work[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_out[b], cudaEventDisableTiming));
  }
  c->ss_used = 0;
  return BMU_OK;
}

void ctx_close(DevCtx *c) {
  if (c->dev < 0) return;
  cudaSetDevice(c->dev);
  cudaDeviceSynchronize();
  host_ring_free(c);
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c->dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    cudaGetLastError();
  }
  c->ss.xT.release(); c->ss.flags.release(); c->ss.listW.release();
  c->ss.listS.release(); c->ss.counters.release(); c->ss.k2.release(); c->ss.lkeys.release(); c->ss.ldone.release(); c->ss.lparts.release();
  for (int i = 0; i < BMU_NSLOT; i++) {
    c->stage_in[i].release(); c->stage_mask[i].release(); c->stage_idx[i].release();
    c->stage_diff[i].release(); c->stage_nf[i].release(); c->stage_lab[i].release();
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_work[i]) cudaEventDestroy(c->ev_work[i]);
    if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    c->ev_in[i] = c->ev_work[i] = c->ev_out[i] = nullptr;
  }
  c->q2_out.release(); c->stat_f64.release(); c->stat_i64.release(); c->stat_part.release();
  if (c->ss_done) cudaEventDestroy(c->ss_done);
  c->ss_done = nullptr;
  for (int r = 0; r < K_EV_RING; r++)
    for (int i = 0; i < 5; i++) {
      if (c->k1ring[r][i]) cudaEventDestroy(c->k1ring[r][i]);
      if (c->k2ring[r][i]) cudaEventDestroy(c->k2ring[r][i]);
      c->k1ring[r][i] = c->k2ring[r][i] = nullptr;
    }
  c->k1calls = c->k2calls = 0;
  for (int i = 0; i < 8; i++) { if (c->k2sub[i]) cudaEventDestroy(c->k2sub[i]); c->k2sub[i] = nullptr; }
  if (c->k2join) cudaEventDestroy(c->k2join);
  c->k2join = nullptr;
  if (c->k2aux) cudaStreamDestroy(c->k2aux);
  if (c->compute) cudaStreamDestroy(c->compute);
  if (c->copy) cudaStreamDestroy(c->copy);
  if (c->out) cudaStreamDestroy(c->out);
  c->k2aux = c->compute = c->copy = c->out = nullptr;
  c->last_counters = nullptr;
  c->dev = -1;
}
}  // namespace bmu

namespace {
int g_path = BMU_PATH_AUTO;
}  // namespace

extern "C" {

int bmu_init(int device) {
  DevCtx *c = &g_primary_ctx;
  if (c->dev == device && c->compute) return cudaSetDevice(device) == cudaSuccess ? BMU_OK : fail(BMU_ERR_CUDA, "cudaSetDevice");
  return ctx_open(c, device);
}

void bmu_shutdown(void) {
  multi_shutdown();
  ctx_close(&g_primary_ctx);
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
  DevCtx *c = ctx();
  for (int i = 0; i < 5; i++) out[i] = 0;
  if (!c->last_counters) return BMU_OK;
  int h[4];
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, c->last_counters, sizeof(h), cudaMemcpyDeviceToHost));
  out[0] = c->last_rows;
  out[1] = h[0];
  out[2] = h[1];
  out[3] = c->last_used_k2 ? h[2] : 0;
  out[4] = c->last_used_k2 ? h[3] : 0;
  return BMU_OK;
}

// ------------------------------------------------------------------ codebook
static bool k2_worth_building(long M, int D, unsigned cb_flags) {
  // the operands of the tensor-core filter are built WITH the codebook (synchronously, on the
  // context's stream) whenever a later search could pick the filter on its own; a search that is
  // forced onto the filter afterwards (bmu_set_search_path) builds them inside the search, ordered
  // like every search by the context's ss_done event
  return k2_eligible(g_path, M, D, 1L << 30, 1, cb_flags);
}

static int codebook_finish(bmu_codebook *cb) {       // after d_codes is written on g_compute
  cudaError_t e = k1_prepare_codebook(cb->d_codes, cb->M, cb->D, cb->d_cT, cb->d_flags, g_compute);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(&cb->h_flags, cb->d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, g_compute);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g_compute);
  if (e != cudaSuccess) return fail(BMU_ERR_CUDA, "codebook upload failed: %s", cudaGetErrorString(e));
  k2_codebook_invalidate(&cb->k2);
  if (k2_worth_building(cb->M, cb->D, cb->h_flags)) {
    e = k2_prepare_codebook(&cb->k2, cb->d_codes, cb->M, cb->D, g_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_compute);
    if (e != cudaSuccess) return fail(BMU_ERR_CUDA, "k2_prepare_codebook: %s", cudaGetErrorString(e));
  }
  return BMU_OK;
}

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
  cb->owner = ctx();
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
  cudaError_t e = cudaSuccess;
  if (h_src) e = cudaMemcpyAsync(cb->d_codes, h_src, bytes, cudaMemcpyHostToDevice, g_compute);
  else if (d_src) e = cudaMemcpyAsync(cb->d_codes, d_src, bytes, cudaMemcpyDeviceToDevice, g_compute);
  if (e != cudaSuccess) {
    fail(BMU_ERR_CUDA, "codebook upload failed: %s", cudaGetErrorString(e));
    bmu_codebook_destroy(cb);
    return nullptr;
  }
  if ((h_src || d_src) && codebook_finish(cb)) {
    bmu_codebook_destroy(cb);
    return nullptr;
  }
  return cb;
}

}  // extern "C"
namespace bmu {
// multi-GPU path: an empty replica whose d_codes the caller fills (ncclBroadcast) before codebook_ready
bmu_codebook *codebook_alloc(long M, int D) { return codebook_from_dev(nullptr, nullptr, M, D); }
int codebook_ready(bmu_codebook *cb) { return codebook_finish(cb); }
int codebook_set_labels(bmu_codebook *cb, const int32_t *label) {
  if (cb->d_label) { cudaFree(cb->d_label); cb->d_label = nullptr; }
  if (!label) return BMU_OK;
  if (cudaMalloc((void **)&cb->d_label, (size_t)cb->M * 4) != cudaSuccess) {
    cudaGetLastError();
    return fail(BMU_ERR_NOMEM, "cudaMalloc of the code labels failed");
  }
  CK(cudaMemcpyAsync(cb->d_label, label, (size_t)cb->M * 4, cudaMemcpyHostToDevice, g_compute));
  CK(cudaStreamSynchronize(g_compute));
  return BMU_OK;
}
}  // namespace bmu
extern "C" {

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
  DevCtx *c = ctx();
  // a search still in flight on another stream reads the images that are rewritten here
  if (c->ss_used) CK(cudaStreamWaitEvent(g_compute, c->ss_done, 0));
  CK(cudaMemcpyAsync(cb->d_codes, codes, (size_t)cb->M * cb->D * sizeof(float),
                     cudaMemcpyHostToDevice, g_compute));
  if (cb->d_cq) { cudaFree(cb->d_cq); cb->d_cq = nullptr; }
  return codebook_finish(cb);
}

void bmu_codebook_destroy(bmu_codebook *cb) {
  if (!cb) return;
  if (cb->d_codes) cudaFree(cb->d_codes);
  if (cb->d_cT) cudaFree(cb->d_cT);
  if (cb->d_flags) cudaFree(cb->d_flags);
  if (cb->d_cq) cudaFree(cb->d_cq);
  if (cb->d_label) cudaFree(cb->d_label);
  k2_codebook_free(&cb->k2);
  free(cb);
}

// ------------------------------------------------------------------ search
}  // extern "C"
namespace bmu {
int search_dev_impl(bmu_codebook *cb, const float *d_data, const unsigned char *d_mask, long N, int k,
                    int32_t *d_idx, float *d_diff, int32_t *d_nfound, cudaStream_t st) {
  if (!cb || !d_data || !d_idx || !d_diff || !d_nfound) return fail(BMU_ERR_ARG, "NULL argument");
  if (k < 1 || k > BMU_KMAX) return fail(BMU_ERR_ARG, "k=%d outside 1..%d", k, BMU_KMAX);
  if (N < 0 || N > 0x7fffff00L) return fail(BMU_ERR_ARG, "bad N=%ld", N);
  if (N == 0) return BMU_OK;
  DevCtx *c = ctx();
  if (cb->owner != c) return fail(BMU_ERR_ARG, "codebook belongs to another device context");
  SearchScratch &ss = c->ss;
  const int D = cb->D;
  int rc;
  const bool use_k2 = k2_eligible(g_path, cb->M, D, N, k, cb->h_flags);
  const bool need_tiles = (k == 1) && !cb->h_flags && !use_k2;
  // the scratch (and the K2 operands of the codebook) are shared by every search of this context:
  // wait for the search before, whatever stream it ran on.  Growing a buffer frees the old one, which
  // the CUDA runtime orders after all work already queued on the device.
  if (c->ss_used) CK(cudaStreamWaitEvent(st, c->ss_done, 0));
  if (need_tiles && (rc = ss.xT.ensure(k1_xT_floats(N, D) * sizeof(float)))) return rc;
  if ((rc = ss.flags.ensure((size_t)N))) return rc;
  if ((rc = ss.listW.ensure((size_t)N * sizeof(int)))) return rc;
  if ((rc = ss.listS.ensure((size_t)N * sizeof(int)))) return rc;
  if ((rc = ss.counters.ensure(16 * sizeof(int)))) return rc;
  const bool list8 = k == 1 && d_mask == nullptr;      // k1_list8_kernel: keys (all ones) + arrival counters (zero)
  const bool listk = k >= 2 && k <= 5 && d_mask == nullptr;   // k1_listk_kernel: per-slice keys + arrival counters
  if (list8 || listk) {
    // the kernels leave the arrays as they found them; a new (larger) allocation is initialised once
    if (list8 && ss.lkeys.bytes < (size_t)N * 8) {
      if ((rc = ss.lkeys.ensure((size_t)N * 8))) return rc;
      CK(cudaMemsetAsync(ss.lkeys.p, 0xff, ss.lkeys.bytes, st));
    }
    if (ss.ldone.bytes < ((size_t)N / 4 + 1) * sizeof(int)) {
      if ((rc = ss.ldone.ensure(((size_t)N / 4 + 1) * sizeof(int)))) return rc;
      CK(cudaMemsetAsync(ss.ldone.p, 0, ss.ldone.bytes, st));
    }
    if (listk && (rc = ss.lparts.ensure(k1_listk_parts_bytes()))) return rc;
  }

  K1Args a;
  a.data = d_data; a.mask = d_mask; a.codes = cb->d_codes; a.cT = cb->d_cT;
  a.cb_flags = cb->d_flags; a.N = N; a.M = cb->M; a.D = D; a.k = k;
  a.skip_fast = need_tiles ? 0 : 1;
  a.num_sms = c->sms;
  a.short_list = use_k2 ? 1 : 0;
  a.xT = (float *)ss.xT.p; a.flags = (unsigned char *)ss.flags.p;
  a.listW = (int *)ss.listW.p; a.listS = (int *)ss.listS.p; a.counters = (int *)ss.counters.p;
  a.lkeys = nullptr; a.ldone = nullptr;
  a.lparts = nullptr;
  if (list8) { a.lkeys = (unsigned long long *)ss.lkeys.p; a.ldone = (int *)ss.ldone.p; }
  if (listk) { a.lparts = (unsigned long long *)ss.lparts.p; a.ldone = (int *)ss.ldone.p; }
  a.idx = d_idx; a.diff = d_diff; a.nfound = d_nfound;
  c->last_counters = a.counters;
  c->last_rows = N;
  c->last_used_k2 = use_k2 ? 1 : 0;
  cudaError_t e = use_k2 ? k2_search(&cb->k2, a, &ss.k2.p, &ss.k2.bytes, st) : k1_search(a, st);
  if (e != cudaSuccess) return fail(BMU_ERR_CUDA, "%s: %s", use_k2 ? "k2_search" : "k1_search", cudaGetErrorString(e));
  CK(cudaEventRecord(c->ss_done, st));
  c->ss_used = 1;
  return BMU_OK;
}
}  // namespace bmu
extern "C" {

int bmu_search_dev(bmu_codebook *cb, const float *d_data, const unsigned char *d_mask, long N,
                   int k, int32_t *d_idx, float *d_diff, int32_t *d_nfound, void *stream) {
  int rc = ensure_init();
  if (rc) return rc;
  return search_dev_impl(cb, d_data, d_mask, N, k, d_idx, d_diff, d_nfound, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ qerror -qetype 1
// find_qerror2 (som_rout.c:823-891): winner search (k = 1) followed by the neighbourhood-weighted
// pass K4 over the same resident chunk; the host adds out[] in data order (one float, som_rout.c:872).
// Chunks are double buffered: the rows of chunk c+1 are copied in (copy stream) and the values of
// chunk c-1 drain (out stream) while chunk c is searched and weighted (compute stream).
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
  DevCtx *c = ctx();
  if (cb->owner != c) return fail(BMU_ERR_ARG, "codebook belongs to another device context");
  const int D = cb->D;
  if (!cb->d_cq) {
    if (cudaMalloc((void **)&cb->d_cq, (size_t)k4_mp(cb->M) * D * sizeof(float)) != cudaSuccess) {
      cb->d_cq = nullptr;
      return fail(BMU_ERR_NOMEM, "cudaMalloc of the component-major codebook failed");
    }
    CK(k4_transpose_codebook(cb->d_codes, cb->M, D, cb->d_cq, g_compute));
    k1_count_launch(1);
  }
  long chunk = (64L << 20) / ((long)D * 4);
  {
    const char *env = getenv("SOMLVQ_CHUNK_ROWS");       // tests: many chunks from a small call
    if (env && atol(env) > 0) chunk = atol(env);
  }
  if (chunk < 1) chunk = 1;
  if (chunk > N) chunk = N;
  Scratch *q2[2] = {&c->q2_out, &c->stage_lab[0]};          // per-sample values of the two slots
  for (int b = 0; b < 2; b++) {
    if ((rc = c->stage_in[b].ensure((size_t)chunk * D * 4))) return rc;
    if (mask && (rc = c->stage_mask[b].ensure((size_t)chunk * D))) return rc;
    if ((rc = c->stage_idx[b].ensure((size_t)chunk * 4))) return rc;
    if ((rc = c->stage_diff[b].ensure((size_t)chunk * 4))) return rc;
    if ((rc = c->stage_nf[b].ensure((size_t)chunk * 4))) return rc;
    if ((rc = q2[b]->ensure((size_t)chunk * 4))) return rc;
  }
  int status = BMU_OK;
  long cidx = 0;
  for (long n0 = 0; n0 < N && status == BMU_OK; n0 += chunk, cidx++) {
    const int b = (int)(cidx & 1);
    const long n = (N - n0 < chunk) ? N - n0 : chunk;
    cudaError_t e = cudaSuccess;
    // slot b is free once chunk c-2 has been computed (inputs) and drained (values)
    if (cidx >= 2) e = cudaStreamWaitEvent(g_copy, c->ev_work[b], 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(c->stage_in[b].p, data + n0 * (long)D, (size_t)n * D * 4, cudaMemcpyHostToDevice, g_copy);
    if (e == cudaSuccess && mask)
      e = cudaMemcpyAsync(c->stage_mask[b].p, mask + n0 * (long)D, (size_t)n * D, cudaMemcpyHostToDevice, g_copy);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_in[b], g_copy);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_compute, c->ev_in[b], 0);
    if (e == cudaSuccess && cidx >= 2) e = cudaStreamWaitEvent(g_compute, c->ev_out[b], 0);
    if (e != cudaSuccess) { status = fail(BMU_ERR_CUDA, "bmu_qerror2: %s", cudaGetErrorString(e)); break; }
    const unsigned char *d_mask = mask ? (const unsigned char *)c->stage_mask[b].p : nullptr;
    if ((status = search_dev_impl(cb, (const float *)c->stage_in[b].p, d_mask, n, 1, (int32_t *)c->stage_idx[b].p,
                                  (float *)c->stage_diff[b].p, (int32_t *)c->stage_nf[b].p, g_compute)))
      break;
    e = k4_qerror2(cb->d_cq, cb->M, D, xdim, topol, neigh, radius, (const float *)c->stage_in[b].p, d_mask, n,
                   (const int32_t *)c->stage_idx[b].p, (const int32_t *)c->stage_nf[b].p, (float *)q2[b]->p, c->sms,
                   g_compute);
    k1_count_launch(1);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_work[b], g_compute);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_out, c->ev_work[b], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out + n0, q2[b]->p, (size_t)n * 4, cudaMemcpyDeviceToHost, g_out);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_out[b], g_out);
    if (e != cudaSuccess) status = fail(BMU_ERR_CUDA, "bmu_qerror2: %s", cudaGetErrorString(e));
  }
  // single exit: nothing that references the caller's buffers is left queued, error or not
  cudaError_t e0 = cudaStreamSynchronize(g_copy), e1 = cudaStreamSynchronize(g_compute), e2 = cudaStreamSynchronize(g_out);
  if (status) return status;
  if (e0 != cudaSuccess || e1 != cudaSuccess || e2 != cudaSuccess)
    return fail(BMU_ERR_CUDA, "bmu_qerror2: %s", cudaGetErrorString(e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2)));
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
// Per-shard sums of a finished search (find_qerror's accumulator som_rout.c:710-721, the hit counts of
// vcal.c:109-131 / accuracy.c:82-118 / cmatr.c:84-109), in the form ONE all-reduce can combine:
// counts are int64 (exact whatever the order), the sum of sqrt(diff) is a double that is reduced in a
// FIXED order -- per-thread grid-stride partial, xor-shuffle tree, warp partials in warp order, block
// partials in block order by the last block to finish -- so it is bit-identical from run to run for a
// given (N, grid).  It is NOT the reference's sequential float sum; bmu_replay_qerror is.
}  // extern "C"
// SH: BMU hit counts go to a shared-memory histogram first (M <= STATS_SMEM_BINS: 48 KB of 32-bit counters, one CTA
// of 1024 threads per SM) and reach the global int64 counters once per CTA and bin -- 10 M global atomics become
// 148 x M; larger maps count with global atomics directly.
#define STATS_SMEM_BINS 12288
template <int THREADS, bool SH>
__global__ void __launch_bounds__(THREADS)
stats_kernel(const int32_t *__restrict__ idx, const float *__restrict__ diff,
             const int32_t *__restrict__ nfound, long N, int k, long M, double *__restrict__ sum_out,
             unsigned long long *__restrict__ nfound_out, unsigned long long *__restrict__ hist,
             const int32_t *__restrict__ slabel, const int32_t *__restrict__ clabel, int L,
             unsigned long long *__restrict__ conf, double *__restrict__ part, unsigned *__restrict__ ticket) {
  extern __shared__ unsigned sh_hist[];
  __shared__ double ws[THREADS / 32];
  __shared__ unsigned long long wc[THREADS / 32];
  __shared__ bool last;
  const bool shist = SH && hist != nullptr;
  if (shist) {
    for (long b = threadIdx.x; b < M; b += THREADS) sh_hist[b] = 0u;
    __syncthreads();
  }
  double s = 0.0;
  unsigned long long cnt = 0;
  for (long n = blockIdx.x * (long)blockDim.x + threadIdx.x; n < N; n += (long)gridDim.x * blockDim.x) {
    const int j = idx[n * k];
    if (nfound[n] == 0 || j < 0) continue;
    s += sqrt((double)diff[n * k]);
    cnt++;
    if (hist && j < M) {
      if (shist) atomicAdd(&sh_hist[j], 1u);
      else atomicAdd(&hist[j], 1ull);
    }
    if (conf) {
      const int a = slabel[n], b = clabel[j];
      if (a >= 0 && a < L && b >= 0 && b < L) atomicAdd(&conf[(long)a * L + b], 1ull);
    }
  }
  for (int off = 16; off >= 1; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = s; wc[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (shist)
    for (long b = threadIdx.x; b < M; b += THREADS) {
      const unsigned v = sh_hist[b];
      if (v) atomicAdd(&hist[b], (unsigned long long)v);
    }
  if (threadIdx.x == 0) {
    double bs = 0.0;
    unsigned long long bc = 0;
    for (int w = 0; w < THREADS / 32; w++) { bs += ws[w]; bc += wc[w]; }
    part[blockIdx.x] = bs;
    if (bc) atomicAdd(nfound_out, bc);
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned b2 = 0; b2 < gridDim.x; b2++) t += ((volatile double *)part)[b2];
    *sum_out += t;                       // accumulates over the chunks of a call, in chunk order
    *ticket = 0;
  }
}

namespace bmu {
int stats_accumulate(DevCtx *c, const int32_t *d_idx, const float *d_diff, const int32_t *d_nfound, long N, int k,
                     long M, double *d_sum, long long *d_nfound_total, long long *d_hist, const int32_t *d_slabel,
                     const int32_t *d_clabel, int L, long long *d_conf, cudaStream_t st) {
  if (N <= 0) return BMU_OK;
  const bool sh = d_hist != nullptr && M <= STATS_SMEM_BINS;
  const int grid = sh ? c->sms : c->sms * 4;
  if (c->stat_part.bytes < (size_t)(c->sms * 4 + 2) * 8) {
    int rc = c->stat_part.ensure((size_t)(c->sms * 4 + 2) * 8);
    if (rc) return rc;
    CK(cudaMemsetAsync(c->stat_part.p, 0, c->stat_part.bytes, st));
  }
  double *part = (double *)c->stat_part.p;
  if (sh) {
    const size_t smem = (size_t)M * sizeof(unsigned);
    CK(cudaFuncSetAttribute(stats_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stats_kernel<1024, true><<<grid, 1024, smem, st>>>(d_idx, d_diff, d_nfound, N, k, M, d_sum,
                                                      (unsigned long long *)d_nfound_total, (unsigned long long *)d_hist,
                                                      d_slabel, d_clabel, L, (unsigned long long *)d_conf, part + 1,
                                                      (unsigned *)part);
  } else {
    stats_kernel<256, false><<<grid, 256, 0, st>>>(d_idx, d_diff, d_nfound, N, k, M, d_sum,
                                                  (unsigned long long *)d_nfound_total, (unsigned long long *)d_hist,
                                                  d_slabel, d_clabel, L, (unsigned long long *)d_conf, part + 1,
                                                  (unsigned *)part);
  }
  k1_count_launch(1);
  CK(cudaGetLastError());
  return BMU_OK;
}
}  // namespace bmu
extern "C" {

int bmu_search_stats_dev(const int32_t *d_idx, const float *d_diff, const int32_t *d_nfound,
                         long N, int k, long M, double *d_sum, long long *d_nfound_total, long long *d_hist,
                         const int32_t *d_sample_label, const int32_t *d_code_label,
                         int n_labels, long long *d_confusion, void *stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if (!d_idx || !d_diff || !d_nfound || !d_sum || !d_nfound_total) return fail(BMU_ERR_ARG, "NULL argument");
  if (d_confusion && (!d_sample_label || !d_code_label || n_labels <= 0))
    return fail(BMU_ERR_ARG, "confusion counts need labels");
  return stats_accumulate(ctx(), d_idx, d_diff, d_nfound, N, k, M, d_sum, d_nfound_total, d_hist, d_sample_label,
                          d_code_label, n_labels, d_confusion, (cudaStream_t)stream);
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

// The row used at every training step.  Without -buffer (or with a buffer larger than the file) the list
// is shuffled once when the file is read (datafile.c:340-341) and walked cyclically (som_rout.c:602-610).
// With `-buffer B` the reference holds B entries at a time: every chunk it reads is shuffled on its own
// (read_entries, datafile.c:237-344), at the end of the file it rewinds and reads -- and shuffles -- the
// chunks again, and the generator's state (lvq_pak.c:459-473) simply runs on from shuffle to shuffle.  A
// host that keeps the whole file in memory reproduces that order from the row counts alone.
void bmu_sample_sequence(long N, long buffer, int seed, long nsteps, int32_t *sample) {
  unsigned long state = (unsigned long)seed;
  if (N <= 0 || nsteps <= 0) return;
  const long chunk = (buffer <= 0 || buffer > N) ? N : buffer;
  const bool reread = chunk == buffer;                      // buffered: every pass re-reads and re-shuffles
  int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)chunk);
  int32_t *once = reread ? nullptr : (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
  if (!tmp || (!reread && !once)) { free(tmp); free(once); return; }
  auto shuffle = [&](int32_t *t, long n, long first) {
    for (long i = 0; i < n; i++) t[i] = (int32_t)(first + i);
    if (!seed) return;
    for (long i = 0; i < n; i++) {                          // datafile.c:1169-1175
      state = (state * 23UL) % 100000001UL;
      const long j = (long)(int)(state % 32767UL) % n;
      const int32_t v = t[i];
      t[i] = t[j];
      t[j] = v;
    }
  };
  if (!reread) {
    shuffle(once, N, 0);
    for (long le = 0; le < nsteps; le++) sample[le] = once[le % N];
  } else {
    long le = 0;
    while (le < nsteps)
      for (long c0 = 0; c0 < N && le < nsteps; c0 += chunk) {
        const long n = N - c0 < chunk ? N - c0 : chunk;
        shuffle(tmp, n, c0);
        for (long i = 0; i < n && le < nsteps; i++) sample[le++] = tmp[i];
      }
  }
  free(tmp);
  free(once);
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
