// bmu_host.cu -- the host-pointer side of the batch search: bmu_search() and the chunk pipeline the
// multi-GPU entry points share.
//
// The reference's hosts hold their vectors in plain calloc'd memory (datafile.c:472), so the caller's
// buffers are normally PAGEABLE.  A cudaMemcpyAsync from pageable memory is staged by the driver through
// one small pinned buffer by one thread (11 GB/s measured) and never overlaps anything.  Here the
// staging is ours:
//
//   caller rows --(T copy threads)--> small pinned ring (pieces of 8 MB) --(DMA, copy stream)--> device
//        chunk slot --> search kernels (+ statistics) on the compute stream --> results --(DMA, out
//        stream)--> pinned --(copy threads)--> caller's idx / diff / nfound
//
// with BMU_NSLOT device chunks in flight, so that the staging of chunk c+1, the H2D copy of chunk c, the
// kernels of chunk c-1 and the D2H copy of chunk c-2 all run at the same time.  The pinned ring is kept
// SMALL on purpose (6 x 8 MB): written with ordinary stores it stays in the CPU's last-level cache, the
// DMA engine reads the pieces from there, and DRAM only sees the one read of the caller's rows.  Measured
// on the B200 box (tools/ubench/host_pipe.cu, profiles/r02_host_copy_ubench.txt): 54 GB/s end to end with
// 8 threads against 55.3 GB/s for the DMA alone from pinned memory; a large ring written with
// non-temporal stores reached 52 GB/s, glibc memcpy into a large ring 37-47 GB/s, cudaHostRegister of the
// caller's pages on the fly 7-10 GB/s.  Buffers that are already page-locked (bmu_host_alloc,
// bmu_host_register, cudaHostAlloc, torch pin_memory) are recognised with cudaPointerGetAttributes and
// read by the DMA engine directly.
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "api_internal.h"
#include "common.cuh"
#include "copy_pool.h"

namespace bmu {

// ------------------------------------------------------------------ pinned ring
struct Pinned {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return BMU_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    if (cudaHostAlloc(&p, need, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return fail(BMU_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed", need);
    }
    bytes = need;
    return BMU_OK;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
  }
};

#define RING_PIECES 6
#define RING_PIECE_MAX ((size_t)8 << 20)

struct HostRing {
  Pinned pieces;                                   // RING_PIECES x piece_bytes, H2D staging
  size_t piece_bytes = 0;
  cudaEvent_t piece_done[RING_PIECES] = {};        // the DMA that read piece i has finished
  bool piece_used[RING_PIECES] = {};
  unsigned long next_piece = 0;
  Pinned idx[BMU_NSLOT], diff[BMU_NSLOT], nf[BMU_NSLOT];   // D2H landing buffers per device chunk slot
  CopyPool *pool = nullptr;
};

static std::atomic<int> g_copy_threads{0};     // 0 = decide from the machine

void host_set_copy_threads(int n) { g_copy_threads = n; }

static int copy_threads_default() {
  int n = g_copy_threads.load();
  if (n <= 0) {
    const char *s = getenv("SOMLVQ_COPY_THREADS");
    if (s && atoi(s) > 0) n = atoi(s);
  }
  if (n <= 0) {
    // share the cores with the other ranks of this box (torchrun's LOCAL_WORLD_SIZE) -- the copy threads
    // of eight processes must not oversubscribe the cores that also run eight launch threads
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
    int sharers = 1;
    const char *lw = getenv("LOCAL_WORLD_SIZE");
    if (lw && atoi(lw) > 1) sharers = atoi(lw);
    n = (int)(cores / sharers);
    if (n > 8) n = 8;
  }
  if (n < 1) n = 1;
  if (n > 64) n = 64;
  return n;
}

static std::atomic<int> g_host_sharers{0};      // shards of this process that stage at the same time (bmu_multi_search)
void host_set_sharers(int n) { g_host_sharers = n; }

// The ring only pays while it stays in the last-level cache, and every rank / shard that feeds a GPU from
// this host has one: 8 MB pieces for a single feeder (54 GB/s, profiles/r02_host_copy_ubench.txt), smaller
// ones when several share the cache ($SOMLVQ_RING_PIECE_MB overrides).
static size_t ring_piece_bytes() {
  const char *env = getenv("SOMLVQ_RING_PIECE_MB");
  if (env && atoi(env) > 0) return (size_t)atoi(env) << 20;
  int sharers = g_host_sharers.load();
  const char *lw = getenv("LOCAL_WORLD_SIZE");
  if (lw && atoi(lw) > sharers) sharers = atoi(lw);
  size_t piece = RING_PIECE_MAX;
  while (sharers > 1 && piece > ((size_t)1 << 20)) { piece >>= 1; sharers >>= 1; }
  return piece;
}

static HostRing *ring_of(DevCtx *c, int threads_hint) {
  if (!c->ring) c->ring = new HostRing();
  if (!c->ring->pool) c->ring->pool = new CopyPool(threads_hint > 0 ? threads_hint : copy_threads_default());
  return c->ring;
}

void host_ring_free(DevCtx *c) {
  if (!c->ring) return;
  c->ring->pieces.release();
  for (int i = 0; i < RING_PIECES; i++)
    if (c->ring->piece_done[i]) cudaEventDestroy(c->ring->piece_done[i]);
  for (int b = 0; b < BMU_NSLOT; b++) { c->ring->idx[b].release(); c->ring->diff[b].release(); c->ring->nf[b].release(); }
  delete c->ring->pool;
  delete c->ring;
  c->ring = nullptr;
}

static bool is_pinned(const void *p) {
  if (!p) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// rows per chunk: about 64 MB of input, whole waves of the persistent kernels when the call is that large
static long chunk_rows(long N, int D, int sms) {
  // $SOMLVQ_CHUNK_ROWS: rows per chunk, for tests that want many chunks from a small call
  const char *env = getenv("SOMLVQ_CHUNK_ROWS");
  if (env && atol(env) > 0) return atol(env) < N ? atol(env) : N;
  long rows = (64L << 20) / ((long)D * 4);
  const long wave = (long)sms * 512;                   // 4 row tiles of 128 rows per CTA pass (k2_rec_kernel)
  if (rows >= wave) rows = rows / wave * wave;
  else rows = rows / K1_TS * K1_TS;
  if (rows < K1_TS) rows = K1_TS;
  if (rows > N) rows = N;
  return rows;
}

// The chunk pipeline on the calling thread's context.  `hs` (nullable) adds the per-shard statistics:
// they accumulate over the chunks in the context's stat_f64 / stat_i64 buffers (zeroed here).
int search_host_pipeline(bmu_codebook *cb, const float *data, const unsigned char *mask, long N, int k,
                         int32_t *idx, float *diff, int32_t *nfound, const HostStats *hs) {
  DevCtx *c = ctx();
  if (!cb || !data || !idx || !diff || !nfound) return fail(BMU_ERR_ARG, "NULL argument");
  if (cb->owner != c) return fail(BMU_ERR_ARG, "codebook belongs to another device context");
  if (k < 1 || k > BMU_KMAX) return fail(BMU_ERR_ARG, "k=%d outside 1..%d", k, BMU_KMAX);
  if (N < 0) return fail(BMU_ERR_ARG, "bad N");
  const int D = cb->D;
  const long M = cb->M;
  int rc;
  // ---- statistics buffers: [sum] and [n_found, hist[M], confusion[L*L]]
  const int L = hs ? hs->n_labels : 0;
  const bool want_conf = hs && hs->sample_label && cb->d_label && L > 0;
  const size_t ncounts = hs ? 1 + (hs->want_hist ? (size_t)M : 0) + (want_conf ? (size_t)L * L : 0) : 0;
  if (hs) {
    if ((rc = c->stat_f64.ensure(sizeof(double)))) return rc;
    if ((rc = c->stat_i64.ensure(ncounts * sizeof(long long)))) return rc;
    CK(cudaMemsetAsync(c->stat_f64.p, 0, sizeof(double), c->compute));
    CK(cudaMemsetAsync(c->stat_i64.p, 0, ncounts * sizeof(long long), c->compute));
  }
  if (N == 0) {
    if (hs) CK(cudaStreamSynchronize(c->compute));
    return BMU_OK;
  }
  const long chunk = chunk_rows(N, D, c->sms);
  const long nchunks = (N + chunk - 1) / chunk;
  // pageable buffers are staged through the pinned ring; tiny calls (the demo recipes) are left to the driver
  const bool tiny = (size_t)N * D * 4 < (1u << 20) && !getenv("SOMLVQ_CHUNK_ROWS");
  const bool st_in = !tiny && !is_pinned(data), st_mask = !tiny && mask && !is_pinned(mask);
  const bool st_out = !tiny && !(is_pinned(idx) && is_pinned(diff) && is_pinned(nfound));
  const bool st_lab = !tiny && want_conf && !is_pinned(hs->sample_label);
  HostRing *ring = (st_in || st_mask || st_out || st_lab) ? ring_of(c, 0) : nullptr;
  if (ring && (st_in || st_mask || st_lab)) {
    ring->piece_bytes = ring_piece_bytes();
    if ((rc = ring->pieces.ensure(RING_PIECES * ring->piece_bytes))) return rc;
    for (int i = 0; i < RING_PIECES; i++)
      if (!ring->piece_done[i]) CK(cudaEventCreateWithFlags(&ring->piece_done[i], cudaEventDisableTiming));
  }
  struct PoolSession {                       // the copy threads spin for the duration of this call only
    CopyPool *p;
    explicit PoolSession(CopyPool *q) : p(q) { if (p) p->begin(); }
    ~PoolSession() { if (p) p->end(); }
  } session(ring ? ring->pool : nullptr);
  // host -> device on the copy stream: page-locked sources are read by the DMA engine as they are,
  // pageable ones go through the ring piece by piece (copy threads fill piece i+1 while the DMA reads piece i)
  auto h2d = [&](void *dst, const void *src, size_t bytes, bool staged) -> cudaError_t {
    if (!staged) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->copy);
    const size_t pb = ring->piece_bytes;
    for (size_t off = 0; off < bytes; off += pb) {
      const size_t len = bytes - off < pb ? bytes - off : pb;
      const int r = (int)(ring->next_piece++ % RING_PIECES);
      char *piece = (char *)ring->pieces.p + (size_t)r * pb;
      cudaError_t e;
      if (ring->piece_used[r] && (e = cudaEventSynchronize(ring->piece_done[r])) != cudaSuccess) return e;
      ring->pool->copy(piece, (const char *)src + off, len, COPY_CACHED);
      if ((e = cudaMemcpyAsync((char *)dst + off, piece, len, cudaMemcpyHostToDevice, c->copy)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(ring->piece_done[r], c->copy)) != cudaSuccess) return e;
      ring->piece_used[r] = true;
    }
    return cudaSuccess;
  };
  const int nslot = (int)(nchunks < BMU_NSLOT ? nchunks : BMU_NSLOT);
  for (int b = 0; b < nslot; b++) {
    if ((rc = c->stage_in[b].ensure((size_t)chunk * D * 4))) return rc;
    if (mask && (rc = c->stage_mask[b].ensure((size_t)chunk * D))) return rc;
    if ((rc = c->stage_idx[b].ensure((size_t)chunk * k * 4))) return rc;
    if ((rc = c->stage_diff[b].ensure((size_t)chunk * k * 4))) return rc;
    if ((rc = c->stage_nf[b].ensure((size_t)chunk * 4))) return rc;
    if (want_conf && (rc = c->stage_lab[b].ensure((size_t)chunk * 4))) return rc;
    if (st_out) {
      if ((rc = ring->idx[b].ensure((size_t)chunk * k * 4))) return rc;
      if ((rc = ring->diff[b].ensure((size_t)chunk * k * 4))) return rc;
      if ((rc = ring->nf[b].ensure((size_t)chunk * 4))) return rc;
    }
  }
  double *d_sum = (double *)c->stat_f64.p;
  long long *d_cnt = (long long *)c->stat_i64.p;
  long long *d_hist = (hs && hs->want_hist) ? d_cnt + 1 : nullptr;
  long long *d_conf = want_conf ? d_cnt + 1 + (hs->want_hist ? M : 0) : nullptr;

  // results of chunk `cc` (slot b): wait for its D2H copies, then hand them to the caller's arrays
  auto drain = [&](long cc) -> cudaError_t {
    const int b = (int)(cc % BMU_NSLOT);
    cudaError_t e = cudaEventSynchronize(c->ev_out[b]);
    if (e != cudaSuccess || !st_out) return e;
    const long n0 = cc * chunk, n = (N - n0 < chunk) ? N - n0 : chunk;
    ring->pool->copy(idx + n0 * (long)k, ring->idx[b].p, (size_t)n * k * 4, COPY_CACHED);
    ring->pool->copy(diff + n0 * (long)k, ring->diff[b].p, (size_t)n * k * 4, COPY_CACHED);
    ring->pool->copy(nfound + n0, ring->nf[b].p, (size_t)n * 4, COPY_CACHED);
    return cudaSuccess;
  };

  int status = BMU_OK;
  cudaError_t e = cudaSuccess;
  long drained = 0;                          // chunks whose results have reached the caller
  for (long cc = 0; cc < nchunks && status == BMU_OK && e == cudaSuccess; cc++) {
    const int b = (int)(cc % BMU_NSLOT);
    const long n0 = cc * chunk, n = (N - n0 < chunk) ? N - n0 : chunk;
    if (cc >= BMU_NSLOT) {
      // slot b comes free: its results (chunk cc - NSLOT) go to the caller
      if ((e = drain(cc - BMU_NSLOT)) != cudaSuccess) break;
      drained = cc - BMU_NSLOT + 1;
      if ((e = cudaStreamWaitEvent(c->copy, c->ev_work[b], 0)) != cudaSuccess) break;   // device slot searched
    }
    const void *src = data + n0 * (long)D, *msrc = mask ? mask + n0 * (long)D : nullptr;
    const void *lsrc = want_conf ? hs->sample_label + n0 : nullptr;
    e = h2d(c->stage_in[b].p, src, (size_t)n * D * 4, st_in);
    if (e == cudaSuccess && mask) e = h2d(c->stage_mask[b].p, msrc, (size_t)n * D, st_mask);
    if (e == cudaSuccess && want_conf) e = h2d(c->stage_lab[b].p, lsrc, (size_t)n * 4, st_lab);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_in[b], c->copy);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->compute, c->ev_in[b], 0);
    if (e == cudaSuccess && cc >= BMU_NSLOT) e = cudaStreamWaitEvent(c->compute, c->ev_out[b], 0);
    if (e != cudaSuccess) break;
    status = search_dev_impl(cb, (const float *)c->stage_in[b].p,
                             mask ? (const unsigned char *)c->stage_mask[b].p : nullptr, n, k,
                             (int32_t *)c->stage_idx[b].p, (float *)c->stage_diff[b].p,
                             (int32_t *)c->stage_nf[b].p, c->compute);
    if (status) break;
    if (hs && (status = stats_accumulate(c, (const int32_t *)c->stage_idx[b].p, (const float *)c->stage_diff[b].p,
                                         (const int32_t *)c->stage_nf[b].p, n, k, M, d_sum, d_cnt, d_hist,
                                         (const int32_t *)c->stage_lab[b].p, cb->d_label, L, d_conf, c->compute)))
      break;
    if ((e = cudaEventRecord(c->ev_work[b], c->compute)) != cudaSuccess) break;
    if ((e = cudaStreamWaitEvent(c->out, c->ev_work[b], 0)) != cudaSuccess) break;
    void *oi = st_out ? ring->idx[b].p : (void *)(idx + n0 * (long)k);
    void *od = st_out ? ring->diff[b].p : (void *)(diff + n0 * (long)k);
    void *on = st_out ? ring->nf[b].p : (void *)(nfound + n0);
    e = cudaMemcpyAsync(oi, c->stage_idx[b].p, (size_t)n * k * 4, cudaMemcpyDeviceToHost, c->out);
    if (e == cudaSuccess) e = cudaMemcpyAsync(od, c->stage_diff[b].p, (size_t)n * k * 4, cudaMemcpyDeviceToHost, c->out);
    if (e == cudaSuccess) e = cudaMemcpyAsync(on, c->stage_nf[b].p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->out);
    if (e == cudaSuccess) e = cudaEventRecord(c->ev_out[b], c->out);
  }
  // single exit path: drain what was queued (results to the caller when everything went well), and
  // leave nothing behind that still references the caller's buffers
  if (status == BMU_OK && e == cudaSuccess)
    for (long cc = drained; cc < nchunks && e == cudaSuccess; cc++) e = drain(cc);
  cudaError_t e0 = cudaStreamSynchronize(c->copy), e1 = cudaStreamSynchronize(c->compute),
              e2 = cudaStreamSynchronize(c->out);
  if (status) return status;
  if (e == cudaSuccess) e = e0 != cudaSuccess ? e0 : (e1 != cudaSuccess ? e1 : e2);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(BMU_ERR_CUDA, "search failed: %s", cudaGetErrorString(e));
  }
  return BMU_OK;
}

}  // namespace bmu

using namespace bmu;

extern "C" {

// Host-pointer search on the device of bmu_init.  This is bench.py's e2e call.
int bmu_search(bmu_codebook *cb, const float *data, const unsigned char *mask, long N, int k,
               int32_t *idx, float *diff, int32_t *nfound) {
  int rc = ensure_init();
  if (rc) return rc;
  return search_host_pipeline(cb, data, mask, N, k, idx, diff, nfound, nullptr);
}

// page-locked host memory for callers that want the DMA to read their arrays directly (the C host's
// loader parses into it; reference: the calloc of datafile.c:472)
void *bmu_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (ensure_init()) return nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    fail(BMU_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed", bytes);
    return nullptr;
  }
  return p;
}
void bmu_host_free(void *p) {
  if (p) cudaFreeHost(p);
}
int bmu_host_register(void *p, size_t bytes) {
  int rc = ensure_init();
  if (rc) return rc;
  CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return BMU_OK;
}
int bmu_host_unregister(void *p) {
  CK(cudaHostUnregister(p));
  return BMU_OK;
}
int bmu_set_copy_threads(int n) {
  if (n < 0 || n > 64) return fail(BMU_ERR_ARG, "copy threads %d outside 0..64", n);
  host_set_copy_threads(n);
  return BMU_OK;
}

}  // extern "C"
