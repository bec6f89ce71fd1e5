// bmu_multi.cu -- the data-parallel split of the batch search behind the C ABI.
//
// Samples are independent and the codebook is read-only, so a search shards over rows with NO data-path
// collective: the codebook is replicated once (ncclBroadcast), every shard searches a contiguous slice
// of the caller's rows and writes its per-row results straight into the caller's arrays (data order is
// kept, which the hosts' in-order replays need), and only the small statistics vector
//     { double sum sqrt(diff) } + { int64 n_found, hist[M], confusion[L*L] }
// is combined, by ONE grouped NCCL all-reduce over NVLink (reference: the accumulators of find_qerror
// som_rout.c:710-721, compute_accuracy accuracy.c:82-118, compute_cmatr cmatr.c:84-109).
//
// Two ways to run it:
//   * one process, all GPUs (the C hosts: bmu_pak qerror / accuracy / cmatr / knntest ...):
//     bmu_multi_init + bmu_mcodebook_* + bmu_multi_search.  One host thread per shard drives the chunk
//     pipeline of bmu_host.cu on its own device context; communicators from ncclCommInitAll.
//   * one process per GPU (torchrun / MPI style launchers, bench.py --gpus N): the launcher moves a
//     ncclUniqueId between the ranks, bmu_comm_init_rank binds the rank's device, and
//     bmu_comm_broadcast_dev / bmu_comm_allreduce_stats_dev run on device buffers.
// NCCL is resolved at run time (dlopen of libnccl.so.2): a single-GPU host does not need it installed;
// with more than one device a missing NCCL is an error, never a silent host-side reduction.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <thread>
#include <vector>

#include "api_internal.h"
#include "common.cuh"

namespace bmu {
bmu_codebook *codebook_alloc(long M, int D);
int codebook_ready(bmu_codebook *cb);
int codebook_set_labels(bmu_codebook *cb, const int32_t *label);
void host_set_copy_threads(int n);
void host_set_sharers(int n);
}  // namespace bmu

using namespace bmu;

// ------------------------------------------------------------------ NCCL, resolved at run time
namespace {
struct Nccl {
  void *h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
} g_nccl;

int nccl_load() {
  if (g_nccl.h) return BMU_OK;
  // NCCL writes its version banner and debug lines to STDOUT unless told otherwise; the hosts' stdout is compared byte
  // for byte with the reference's, so the log goes to stderr (a user's own NCCL_DEBUG_FILE is respected)
  setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
  // A copy that the process has loaded already (PyTorch ships its own libnccl.so.2) is reused: two NCCL
  // builds in one process, the second one opened RTLD_GLOBAL, made a later `import torch` bind to the wrong
  // one.  Otherwise the system library is opened with local scope.
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names)
    if ((h = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;
  if (!h)
    for (const char *n : names)
      if ((h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
  if (!h) return fail(BMU_ERR_NODEV, "NCCL is needed for more than one GPU and libnccl.so.2 could not be loaded: %s", dlerror());
#define SYM(field, name)                                                     \
  g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);                     \
  if (!g_nccl.field) return fail(BMU_ERR_NODEV, "libnccl has no symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommInitAll, "ncclCommInitAll");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.h = h;
  return BMU_OK;
}

#define NC(call)                                                                               \
  do {                                                                                         \
    ncclResult_t r_ = (call);                                                                  \
    if (r_ != ncclSuccess)                                                                     \
      return fail(BMU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_),     \
                  __FILE__, __LINE__);                                                         \
  } while (0)

// ------------------------------------------------------------------ shards of the one-process mode
struct Multi {
  int nshards = 0, ndev = 0;
  DevCtx ctxs[BMU_MAX_GPUS];
  int leader[BMU_MAX_GPUS];        // first shard on the same device (== own index for leaders)
  std::vector<int> leaders;        // one shard per distinct device, in device order
} g_multi;

// dst += src for the per-shard statistics of shards that share a device (logical shards > devices)
__global__ void stats_add_kernel(double *dsum, const double *ssum, long long *dcnt, const long long *scnt, long n) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i == 0) *dsum += *ssum;
  if (i < n) dcnt[i] += scnt[i];
}
}  // namespace

struct bmu_mcodebook {
  long M;
  int D;
  bmu_codebook *rep[BMU_MAX_GPUS];     // one replica per shard context
  int has_labels;
};

namespace bmu {
void multi_shutdown() {
  Multi &m = g_multi;
  for (int s = 0; s < m.nshards; s++) {
    if (m.ctxs[s].comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)m.ctxs[s].comm);
    m.ctxs[s].comm = nullptr;
    ctx_close(&m.ctxs[s]);
  }
  m.nshards = m.ndev = 0;
  m.leaders.clear();
  DevCtx *p = ctx();
  if (p->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((ncclComm_t)p->comm); p->comm = nullptr; }
}
}  // namespace bmu

extern "C" {

int bmu_multi_init(int nshards) {
  Multi &m = g_multi;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(BMU_ERR_NODEV, "no CUDA device (this library has no CPU fallback)");
  }
  if (nshards <= 0) {
    const char *s = getenv("SOMLVQ_GPUS");
    nshards = (s && atoi(s) > 0) ? atoi(s) : ndev;
  }
  if (nshards > BMU_MAX_GPUS) return fail(BMU_ERR_ARG, "%d shards, at most %d", nshards, BMU_MAX_GPUS);
  if (m.nshards == nshards) return BMU_OK;
  if (m.nshards) multi_shutdown();
  int cur = 0;
  cudaGetDevice(&cur);
  m.ndev = nshards < ndev ? nshards : ndev;
  m.leaders.clear();
  for (int s = 0; s < nshards; s++) {
    const int dev = s % m.ndev;            // more shards than devices: logical shards share a device
    int rc = ctx_open(&m.ctxs[s], dev);
    if (rc) { m.nshards = s; multi_shutdown(); cudaSetDevice(cur); return rc; }
    m.leader[s] = dev;                     // shard `dev` is the first one on device `dev`
    if (s < m.ndev) m.leaders.push_back(s);
  }
  m.nshards = nshards;
  if (m.ndev > 1) {
    int rc = nccl_load();
    if (rc) { multi_shutdown(); cudaSetDevice(cur); return rc; }
    ncclComm_t comms[BMU_MAX_GPUS];
    int devs[BMU_MAX_GPUS];
    for (int d = 0; d < m.ndev; d++) devs[d] = d;
    ncclResult_t r = g_nccl.CommInitAll(comms, m.ndev, devs);
    if (r != ncclSuccess) {
      multi_shutdown();
      cudaSetDevice(cur);
      return fail(BMU_ERR_CUDA, "ncclCommInitAll over %d devices failed: %s", m.ndev, g_nccl.GetErrorString(r));
    }
    for (int d = 0; d < m.ndev; d++) {
      m.ctxs[d].comm = comms[d];
      m.ctxs[d].comm_rank = d;
      m.ctxs[d].comm_nranks = m.ndev;
    }
  }
  cudaSetDevice(cur);
  return BMU_OK;
}

int bmu_multi_shards(void) { return g_multi.nshards; }
int bmu_multi_devices(void) { return g_multi.ndev; }

void bmu_mcodebook_destroy(bmu_mcodebook *mc) {
  if (!mc) return;
  int cur = 0;
  cudaGetDevice(&cur);
  for (int s = 0; s < BMU_MAX_GPUS; s++)
    if (mc->rep[s]) {
      bind_ctx(mc->rep[s]->owner);
      bmu_codebook_destroy(mc->rep[s]);
    }
  bind_ctx(nullptr);
  cudaSetDevice(cur);
  free(mc);
}

// replicate `codes` (host) into every shard's replica: one H2D copy to device 0, ONE ncclBroadcast to the
// other devices, device-to-device copies for shards that share a device; then every replica builds its
// kernel-side images
static int mcodebook_fill(bmu_mcodebook *mc, const float *codes) {
  Multi &m = g_multi;
  const size_t bytes = (size_t)mc->M * mc->D * sizeof(float);
  bind_ctx(&m.ctxs[0]);
  CK(cudaMemcpyAsync(mc->rep[0]->d_codes, codes, bytes, cudaMemcpyHostToDevice, m.ctxs[0].compute));
  if (m.ndev > 1) {
    NC(g_nccl.GroupStart());
    for (int d = 0; d < m.ndev; d++) {
      ncclResult_t r = g_nccl.Broadcast(mc->rep[d]->d_codes, mc->rep[d]->d_codes, (size_t)mc->M * mc->D, ncclFloat, 0,
                                        (ncclComm_t)m.ctxs[d].comm, m.ctxs[d].compute);
      if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(BMU_ERR_CUDA, "ncclBroadcast: %s", g_nccl.GetErrorString(r)); }
    }
    NC(g_nccl.GroupEnd());
  }
  for (int d = 0; d < m.ndev; d++) {
    bind_ctx(&m.ctxs[d]);
    CK(cudaStreamSynchronize(m.ctxs[d].compute));
  }
  for (int s = 0; s < m.nshards; s++) {
    bind_ctx(&m.ctxs[s]);
    if (s >= m.ndev)
      CK(cudaMemcpyAsync(mc->rep[s]->d_codes, mc->rep[m.leader[s]]->d_codes, bytes, cudaMemcpyDeviceToDevice,
                         m.ctxs[s].compute));
    int rc = codebook_ready(mc->rep[s]);
    if (rc) return rc;
  }
  return BMU_OK;
}

bmu_mcodebook *bmu_mcodebook_create(const float *codes, long M, int D) {
  Multi &m = g_multi;
  if (!codes) { fail(BMU_ERR_ARG, "codes is NULL"); return nullptr; }
  if (!m.nshards && bmu_multi_init(0)) return nullptr;
  bmu_mcodebook *mc = (bmu_mcodebook *)calloc(1, sizeof(bmu_mcodebook));
  if (!mc) { fail(BMU_ERR_NOMEM, "calloc"); return nullptr; }
  mc->M = M;
  mc->D = D;
  int cur = 0;
  cudaGetDevice(&cur);
  int rc = BMU_OK;
  for (int s = 0; s < m.nshards && !rc; s++) {
    bind_ctx(&m.ctxs[s]);
    mc->rep[s] = codebook_alloc(M, D);
    if (!mc->rep[s]) rc = BMU_ERR_NOMEM;
  }
  if (!rc) rc = mcodebook_fill(mc, codes);
  if (rc) {
    char keep[512];
    memcpy(keep, g_err, sizeof(keep));
    bmu_mcodebook_destroy(mc);
    memcpy(g_err, keep, sizeof(keep));
    mc = nullptr;
  }
  bind_ctx(nullptr);
  cudaSetDevice(cur);
  return mc;
}

int bmu_mcodebook_update(bmu_mcodebook *mc, const float *codes) {
  if (!mc || !codes) return fail(BMU_ERR_ARG, "NULL argument");
  int cur = 0;
  cudaGetDevice(&cur);
  for (int s = 0; s < g_multi.nshards; s++)
    if (mc->rep[s]->d_cq) { cudaFree(mc->rep[s]->d_cq); mc->rep[s]->d_cq = nullptr; }
  int rc = mcodebook_fill(mc, codes);
  bind_ctx(nullptr);
  cudaSetDevice(cur);
  return rc;
}

int bmu_mcodebook_set_labels(bmu_mcodebook *mc, const int32_t *code_label) {
  if (!mc) return fail(BMU_ERR_ARG, "NULL argument");
  int cur = 0, rc = BMU_OK;
  cudaGetDevice(&cur);
  for (int s = 0; s < g_multi.nshards && !rc; s++) {
    bind_ctx(&g_multi.ctxs[s]);
    rc = codebook_set_labels(mc->rep[s], code_label);
  }
  mc->has_labels = code_label && !rc;
  bind_ctx(nullptr);
  cudaSetDevice(cur);
  return rc;
}

void bmu_multi_shard_bounds(long N, int nshards, int shard, long *lo, long *hi) {
  // contiguous and balanced; large calls are cut at multiples of 512 rows (whole CTA passes of the
  // search kernels), small ones row by row
  const long unit = (N / (nshards > 0 ? nshards : 1) >= 4096) ? 512 : 1;
  const long units = (N + unit - 1) / unit;
  long a = units * shard / nshards * unit, b = units * (shard + 1) / nshards * unit;
  if (a > N) a = N;
  if (b > N || shard == nshards - 1) b = N;
  *lo = a;
  *hi = b;
}

int bmu_multi_search(bmu_mcodebook *mc, const float *data, const unsigned char *mask, long N, int k, int32_t *idx,
                     float *diff, int32_t *nfound, bmu_stats *stats) {
  Multi &m = g_multi;
  if (!mc || !data || !idx || !diff || !nfound) return fail(BMU_ERR_ARG, "NULL argument");
  if (!m.nshards) return fail(BMU_ERR_ARG, "bmu_multi_init has not been called");
  if (N < 0) return fail(BMU_ERR_ARG, "bad N");
  if (stats && stats->confusion && (!stats->sample_label || !mc->has_labels || stats->n_labels <= 0))
    return fail(BMU_ERR_ARG, "confusion counts need sample labels, bmu_mcodebook_set_labels and n_labels");
  const int S = m.nshards, D = mc->D;
  int cur = 0;
  cudaGetDevice(&cur);
  HostStats hs;
  hs.sample_label = stats ? stats->sample_label : nullptr;
  hs.n_labels = (stats && stats->confusion) ? stats->n_labels : 0;
  hs.want_hist = stats && stats->hist;
  if (stats && !stats->confusion) hs.sample_label = nullptr;
  // the copy threads of all shards share the machine's cores
  {
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    int per = (int)(cores / S);
    host_set_copy_threads(per < 1 ? 1 : (per > 8 ? 8 : per));
    host_set_sharers(S);
  }
  int rcs[BMU_MAX_GPUS] = {0};
  char errs[BMU_MAX_GPUS][512];
  auto work = [&](int s) {
    bind_ctx(&m.ctxs[s]);
    long lo, hi;
    bmu_multi_shard_bounds(N, S, s, &lo, &hi);
    HostStats h = hs;
    if (h.sample_label) h.sample_label += lo;
    rcs[s] = search_host_pipeline(mc->rep[s], data + lo * (long)D, mask ? mask + lo * (long)D : nullptr, hi - lo, k,
                                  idx + lo * (long)k, diff + lo * (long)k, nfound + lo, stats ? &h : nullptr);
    if (rcs[s]) memcpy(errs[s], g_err, sizeof(errs[s]));
    bind_ctx(nullptr);
  };
  if (S == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int s = 0; s < S; s++) th.emplace_back(work, s);
    for (auto &t : th) t.join();
  }
  host_set_copy_threads(0);
  host_set_sharers(0);
  for (int s = 0; s < S; s++)
    if (rcs[s]) {
      memcpy(g_err, errs[s], sizeof(errs[s]));
      cudaSetDevice(cur);
      return rcs[s];
    }
  int rc = BMU_OK;
  if (stats) {
    const size_t ncounts = 1 + (hs.want_hist ? (size_t)mc->M : 0) + (hs.n_labels ? (size_t)hs.n_labels * hs.n_labels : 0);
    auto reduce = [&]() -> int {
      // shards that share a device: add into the device's first shard (all pipelines have synchronised)
      for (int s = m.ndev; s < S; s++) {
        DevCtx &dst = m.ctxs[m.leader[s]], &src = m.ctxs[s];
        bind_ctx(&dst);
        stats_add_kernel<<<(unsigned)((ncounts + 255) / 256), 256, 0, dst.compute>>>(
            (double *)dst.stat_f64.p, (const double *)src.stat_f64.p, (long long *)dst.stat_i64.p,
            (const long long *)src.stat_i64.p, (long)ncounts);
        k1_count_launch(1);
        CK(cudaGetLastError());
      }
      // ONE grouped all-reduce over the devices: the double sum and the int64 counts
      if (m.ndev > 1) {
        NC(g_nccl.GroupStart());
        for (int d = 0; d < m.ndev; d++) {
          DevCtx &c = m.ctxs[d];
          ncclResult_t r = g_nccl.AllReduce(c.stat_f64.p, c.stat_f64.p, 1, ncclDouble, ncclSum, (ncclComm_t)c.comm, c.compute);
          if (r == ncclSuccess)
            r = g_nccl.AllReduce(c.stat_i64.p, c.stat_i64.p, ncounts, ncclInt64, ncclSum, (ncclComm_t)c.comm, c.compute);
          if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(BMU_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(r)); }
        }
        NC(g_nccl.GroupEnd());
      }
      bind_ctx(&m.ctxs[0]);
      std::vector<long long> h(ncounts);
      CK(cudaMemcpyAsync(&stats->sum_sqrt, m.ctxs[0].stat_f64.p, sizeof(double), cudaMemcpyDeviceToHost, m.ctxs[0].compute));
      CK(cudaMemcpyAsync(h.data(), m.ctxs[0].stat_i64.p, ncounts * sizeof(long long), cudaMemcpyDeviceToHost, m.ctxs[0].compute));
      for (int d = 0; d < m.ndev; d++) {
        bind_ctx(&m.ctxs[d]);
        CK(cudaStreamSynchronize(m.ctxs[d].compute));
      }
      stats->n_found = h[0];
      size_t off = 1;
      if (hs.want_hist) { memcpy(stats->hist, h.data() + off, (size_t)mc->M * sizeof(long long)); off += (size_t)mc->M; }
      if (hs.n_labels) memcpy(stats->confusion, h.data() + off, (size_t)hs.n_labels * hs.n_labels * sizeof(long long));
      return BMU_OK;
    };
    rc = reduce();
  }
  bind_ctx(nullptr);
  cudaSetDevice(cur);
  return rc;
}

// ------------------------------------------------------------------ one process per GPU
int bmu_comm_unique_id(unsigned char id[128]) {
  int rc = nccl_load();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  NC(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, 128);
  return BMU_OK;
}

int bmu_comm_init_rank(int nranks, int rank, const unsigned char id[128]) {
  int rc = ensure_init();
  if (rc) return rc;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(BMU_ERR_ARG, "bad rank %d of %d", rank, nranks);
  if ((rc = nccl_load())) return rc;
  DevCtx *c = ctx();
  if (c->comm) { g_nccl.CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t comm;
  CK(cudaSetDevice(c->dev));
  NC(g_nccl.CommInitRank(&comm, nranks, u, rank));
  c->comm = comm;
  c->comm_rank = rank;
  c->comm_nranks = nranks;
  return BMU_OK;
}

int bmu_comm_destroy(void) {
  DevCtx *c = ctx();
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->comm);
  c->comm = nullptr;
  c->comm_nranks = 1;
  c->comm_rank = 0;
  return BMU_OK;
}

int bmu_comm_broadcast_dev(void *d_buf, size_t bytes, int root, void *stream) {
  DevCtx *c = ctx();
  if (!c->comm) return c->comm_nranks == 1 ? BMU_OK : fail(BMU_ERR_ARG, "no communicator");
  NC(g_nccl.Broadcast(d_buf, d_buf, bytes, ncclChar, root, (ncclComm_t)c->comm, (cudaStream_t)stream));
  return BMU_OK;
}

int bmu_comm_allreduce_stats_dev(double *d_sum, long nsum, long long *d_counts, long ncounts, void *stream) {
  DevCtx *c = ctx();
  if (!c->comm) return BMU_OK;                          // a single rank: nothing to combine
  NC(g_nccl.GroupStart());
  ncclResult_t r = ncclSuccess;
  if (d_sum && nsum > 0) r = g_nccl.AllReduce(d_sum, d_sum, (size_t)nsum, ncclDouble, ncclSum, (ncclComm_t)c->comm, (cudaStream_t)stream);
  if (r == ncclSuccess && d_counts && ncounts > 0)
    r = g_nccl.AllReduce(d_counts, d_counts, (size_t)ncounts, ncclInt64, ncclSum, (ncclComm_t)c->comm, (cudaStream_t)stream);
  if (r != ncclSuccess) { g_nccl.GroupEnd(); return fail(BMU_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(r)); }
  NC(g_nccl.GroupEnd());
  return BMU_OK;
}

}  // extern "C"
