// bmu_train_api.cu -- C ABI for online training (bmu_trainer_*, bmu_som_train,
// bmu_lvq_train) on top of the persistent kernel K3.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "api_internal.h"
#include "common.cuh"
#include "k1_search.h"
#include "k3_train.h"

using namespace bmu;

// Device memory of a trainer comes from the device's stream-ordered pool (cudaMallocAsync on the context's compute
// stream; ctx_open keeps up to 1 GiB of freed blocks cached in it).  Creating and destroying trainers back to back --
// vfind's trials, the Python wrappers -- otherwise pays cudaMalloc/cudaFree every time, and on a busy host a single
// cudaFree was measured at 0.4-1.2 s (tools/probe/olvq_probe.py) next to a 50 ms training run.
struct TBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

struct bmu_trainer {
  long M, N;
  int D;
  DevCtx *owner = nullptr;       // context (device, compute stream) the trainer lives on
  cudaStream_t st = nullptr;
  float *d_codes = nullptr, *d_data = nullptr, *d_unit_alpha = nullptr, *d_gslice = nullptr;
  unsigned char *d_valid = nullptr;
  short *d_fixed = nullptr;
  int *d_code_label = nullptr, *d_data_label = nullptr;
  unsigned long long *d_slots = nullptr;
  bool has_mask = false;
  int mode = -1;
  int xdim = 0, ydim = 0, topol = 0;
  float win_thr = 0, epsilon = 0, alpha_cap = 0;
  K3Plan plan{};
  TBuf sched_sample, sched_talp, sched_trad;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.0f;
};

namespace {
int t_malloc(bmu_trainer *t, void **p, size_t bytes) {
  if (cudaMallocAsync(p, bytes, t->st) != cudaSuccess) {
    cudaGetLastError();
    *p = nullptr;
    return fail(BMU_ERR_NOMEM, "cudaMallocAsync of %zu bytes failed", bytes);
  }
  return BMU_OK;
}
void t_free(bmu_trainer *t, void *p) {
  if (!p) return;
  // the context may have been closed (or reopened) since: then the stream is gone and plain cudaFree does it
  if (t->owner && t->owner->compute == t->st && t->st) {
    if (cudaFreeAsync(p, t->st) == cudaSuccess) return;
    cudaGetLastError();
  }
  cudaFree(p);
}
int t_ensure(bmu_trainer *t, TBuf *b, size_t need) {
  if (need <= b->bytes) return BMU_OK;
  t_free(t, b->p);
  b->p = nullptr; b->bytes = 0;
  int rc = t_malloc(t, &b->p, need + need / 8);
  if (!rc) b->bytes = need + need / 8;
  return rc;
}
template <typename T>
int dev_upload(bmu_trainer *t, T **dst, const T *src, size_t count) {
  if (*dst) { t_free(t, *dst); *dst = nullptr; }
  if (int rc = t_malloc(t, (void **)dst, count * sizeof(T))) return rc;
  CK(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, t->st));
  return BMU_OK;
}
}  // namespace

extern "C" {

bmu_trainer *bmu_trainer_create(const float *codes, long M, int D, const float *data,
                                const unsigned char *mask, long N) {
  if (ensure_init()) return nullptr;
  if (!codes || !data || M <= 0 || N <= 0 || D <= 0 || M > 0xFFFFFEL) {
    fail(BMU_ERR_ARG, "bad trainer arguments (M=%ld N=%ld D=%d)", M, N, D);
    return nullptr;
  }
  bmu_trainer *t = new bmu_trainer();
  t->M = M; t->N = N; t->D = D;
  t->owner = ctx(); t->st = g_compute;
  int rc = dev_upload(t, &t->d_codes, codes, (size_t)M * D);
  if (!rc) rc = dev_upload(t, &t->d_data, data, (size_t)N * D);
  if (!rc && mask) {
    // masks travel inside the data as a NaN sentinel (k3_train.h); only done if any bit is set
    bool any = false;
    for (size_t i = 0; i < (size_t)N * D && !any; i++) any = mask[i] != 0;
    if (any) {
      unsigned char *d_mask = nullptr;
      rc = dev_upload(t, &d_mask, mask, (size_t)N * D);
      if (!rc) rc = t_malloc(t, (void **)&t->d_valid, (size_t)N);
      if (!rc) {
        cudaError_t e = k3_encode_mask(t->d_data, d_mask, t->d_valid, N, D, g_compute);
        k1_count_launch(1);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_compute);
        if (e != cudaSuccess) rc = fail(BMU_ERR_CUDA, "mask encoding: %s", cudaGetErrorString(e));
      }
      t_free(t, d_mask);
      t->has_mask = true;
    }
  }
  if (!rc) {
    t->plan = k3_plan(M, D, g_sms, g_smem_optin);
    rc = t_malloc(t, (void **)&t->d_slots, sizeof(unsigned long long) * 2 * 16 * t->plan.grid);
    if (!rc && t->plan.gslice_floats) rc = t_malloc(t, (void **)&t->d_gslice, t->plan.gslice_floats * sizeof(float));
  }
  if (!rc && (cudaEventCreate(&t->ev0) != cudaSuccess || cudaEventCreate(&t->ev1) != cudaSuccess))
    rc = fail(BMU_ERR_CUDA, "cudaEventCreate failed");
  if (!rc && cudaStreamSynchronize(g_compute) != cudaSuccess) rc = fail(BMU_ERR_CUDA, "upload failed");
  if (rc) { bmu_trainer_destroy(t); return nullptr; }
  return t;
}

int bmu_trainer_set_som(bmu_trainer *t, int xdim, int ydim, int topol, int neigh,
                        const int16_t *fixed_xy) {
  if (int rc0 = ensure_init()) return rc0;
  if (!t) return fail(BMU_ERR_ARG, "NULL trainer");
  if ((long)xdim * ydim != t->M) return fail(BMU_ERR_ARG, "xdim*ydim=%ld != M=%ld", (long)xdim * ydim, t->M);
  if (topol != BMU_TOPOL_HEXA && topol != BMU_TOPOL_RECT) return fail(BMU_ERR_ARG, "bad topology %d", topol);
  if (neigh != BMU_NEIGH_BUBBLE && neigh != BMU_NEIGH_GAUSSIAN) return fail(BMU_ERR_ARG, "bad neighbourhood %d", neigh);
  t->mode = neigh == BMU_NEIGH_GAUSSIAN ? K3_SOM_GAUSSIAN : K3_SOM_BUBBLE;
  t->xdim = xdim; t->ydim = ydim; t->topol = topol;
  if (t->d_fixed) { t_free(t, t->d_fixed); t->d_fixed = nullptr; }
  if (fixed_xy) {
    int rc = dev_upload(t, &t->d_fixed, (const short *)fixed_xy, (size_t)t->N * 2);
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_compute));
  }
  return BMU_OK;
}

int bmu_trainer_set_lvq(bmu_trainer *t, int algo, const int32_t *code_label,
                        const int32_t *data_label, float win_thr, float epsilon,
                        float alpha_cap, const float *unit_alpha) {
  if (int rc0 = ensure_init()) return rc0;
  if (!t || !code_label || !data_label) return fail(BMU_ERR_ARG, "NULL argument");
  switch (algo) {
    case BMU_LVQ1: t->mode = K3_LVQ1; break;
    case BMU_LVQ2: t->mode = K3_LVQ2; break;
    case BMU_LVQ3: t->mode = K3_LVQ3; break;
    case BMU_OLVQ1: t->mode = K3_OLVQ1; break;
    default: return fail(BMU_ERR_ARG, "bad LVQ algorithm %d", algo);
  }
  if (algo == BMU_OLVQ1 && !unit_alpha) return fail(BMU_ERR_ARG, "OLVQ1 needs unit_alpha");
  if ((algo == BMU_LVQ2 || algo == BMU_LVQ3) && t->M < 2) return fail(BMU_ERR_ARG, "LVQ2/3 need M >= 2");
  t->win_thr = win_thr; t->epsilon = epsilon; t->alpha_cap = alpha_cap;
  int rc = dev_upload(t, &t->d_code_label, (const int *)code_label, (size_t)t->M);
  if (!rc) rc = dev_upload(t, &t->d_data_label, (const int *)data_label, (size_t)t->N);
  if (!rc && unit_alpha) rc = dev_upload(t, &t->d_unit_alpha, unit_alpha, (size_t)t->M);
  if (rc) return rc;
  CK(cudaStreamSynchronize(g_compute));
  return BMU_OK;
}

int bmu_trainer_steps(bmu_trainer *t, const int32_t *sample, const float *talp,
                      const float *trad, long nsteps) {
  if (int rc0 = ensure_init()) return rc0;
  if (!t || !sample) return fail(BMU_ERR_ARG, "NULL argument");
  if (t->mode < 0) return fail(BMU_ERR_ARG, "trainer mode not set (bmu_trainer_set_som/_lvq)");
  if (nsteps <= 0) return BMU_OK;
  const bool som = t->mode <= K3_SOM_GAUSSIAN;
  if (som && (!talp || !trad)) return fail(BMU_ERR_ARG, "SOM training needs talp and trad");
  if (!som && t->mode != K3_OLVQ1 && !talp) return fail(BMU_ERR_ARG, "LVQ training needs talp");
  for (long i = 0; i < nsteps; i++)
    if (sample[i] < 0 || sample[i] >= t->N) return fail(BMU_ERR_ARG, "sample[%ld]=%d out of range", i, sample[i]);
  int rc;
  if ((rc = t_ensure(t, &t->sched_sample, (size_t)nsteps * 4))) return rc;
  CK(cudaMemcpyAsync(t->sched_sample.p, sample, (size_t)nsteps * 4, cudaMemcpyHostToDevice, g_compute));
  if (talp) {
    if ((rc = t_ensure(t, &t->sched_talp, (size_t)nsteps * 4))) return rc;
    CK(cudaMemcpyAsync(t->sched_talp.p, talp, (size_t)nsteps * 4, cudaMemcpyHostToDevice, g_compute));
  }
  if (trad) {
    if ((rc = t_ensure(t, &t->sched_trad, (size_t)nsteps * 4))) return rc;
    CK(cudaMemcpyAsync(t->sched_trad.p, trad, (size_t)nsteps * 4, cudaMemcpyHostToDevice, g_compute));
  }
  K3Params p{};
  p.codes = t->d_codes; p.data = t->d_data; p.valid = t->d_valid;
  p.N = t->N; p.M = t->M; p.D = t->D; p.mode = t->mode;
  p.xdim = t->xdim; p.ydim = t->ydim; p.topol = t->topol; p.fixed_xy = t->d_fixed;
  p.code_label = t->d_code_label; p.data_label = t->d_data_label;
  p.win_thr = t->win_thr; p.epsilon = t->epsilon; p.alpha_cap = t->alpha_cap;
  p.unit_alpha = t->d_unit_alpha;
  p.sample = (const int *)t->sched_sample.p;
  p.talp = talp ? (const float *)t->sched_talp.p : nullptr;
  p.trad = trad ? (const float *)t->sched_trad.p : nullptr;
  p.nsteps = nsteps;
  p.slots = t->d_slots; p.gslice = t->d_gslice;
  p.U = t->plan.U; p.Us = t->plan.Us; p.slice_in_smem = t->plan.slice_in_smem;
  long long *d_prof = nullptr;
  if (getenv("BMU_K3_PROF")) {                       // diagnostic: phase cycles of the fused large-map kernel
    if (cudaMalloc((void **)&d_prof, sizeof(long long) * 8 * (size_t)t->plan.grid) != cudaSuccess) d_prof = nullptr;
    else cudaMemsetAsync(d_prof, 0, sizeof(long long) * 8 * (size_t)t->plan.grid, g_compute);
  }
  p.prof = d_prof;
  {
    const char *env = getenv("BMU_K3_POLL_DELAY_NS");
    p.poll_delay_ns = env ? atoi(env) : K3_POLL_DELAY_AUTO;
  }
  CK(cudaEventRecord(t->ev0, g_compute));
  cudaError_t e = k3_launch(p, t->plan, t->has_mask, g_compute);
  k1_count_launch(1);
  if (e != cudaSuccess) return fail(BMU_ERR_CUDA, "k3_launch: %s", cudaGetErrorString(e));
  CK(cudaEventRecord(t->ev1, g_compute));
  CK(cudaStreamSynchronize(g_compute));
  CK(cudaEventElapsedTime(&t->last_ms, t->ev0, t->ev1));
  if (d_prof) {
    const int G = t->plan.grid;
    long long *h = (long long *)malloc(sizeof(long long) * 8 * (size_t)G);
    if (h && cudaMemcpy(h, d_prof, sizeof(long long) * 8 * (size_t)G, cudaMemcpyDeviceToHost) == cudaSuccess) {
      static const char *names[6] = {"cta_min", "exchange", "barrier", "weights", "pass", "polls x1000"};
      fprintf(stderr, "K3 phase cycles per step (mean over %d CTAs / min / max), %ld steps, %.3f us per step:\n", G, nsteps,
              1e3 * t->last_ms / (double)nsteps);
      for (int i = 0; i < 6; i++) {
        double sum = 0, mn = 1e300, mx = 0;
        for (int g = 0; g < G; g++) {
          const double v = (double)h[(size_t)g * 8 + i] * (i == 5 ? 1000.0 : 1.0) / (double)nsteps;
          sum += v; mn = v < mn ? v : mn; mx = v > mx ? v : mx;
        }
        fprintf(stderr, "  %-9s %8.0f %8.0f %8.0f\n", names[i], sum / G, mn, mx);
      }
    }
    free(h);
    cudaFree(d_prof);
  }
  return BMU_OK;
}

int bmu_trainer_get_codes(bmu_trainer *t, float *codes) {
  if (int rc0 = ensure_init()) return rc0;
  if (!t || !codes) return fail(BMU_ERR_ARG, "NULL argument");
  CK(cudaMemcpyAsync(codes, t->d_codes, (size_t)t->M * t->D * 4, cudaMemcpyDeviceToHost, g_compute));
  CK(cudaStreamSynchronize(g_compute));
  return BMU_OK;
}

int bmu_trainer_get_unit_alpha(bmu_trainer *t, float *unit_alpha) {
  if (int rc0 = ensure_init()) return rc0;
  if (!t || !unit_alpha || !t->d_unit_alpha) return fail(BMU_ERR_ARG, "no unit_alpha state");
  CK(cudaMemcpyAsync(unit_alpha, t->d_unit_alpha, (size_t)t->M * 4, cudaMemcpyDeviceToHost, g_compute));
  CK(cudaStreamSynchronize(g_compute));
  return BMU_OK;
}

float bmu_trainer_last_ms(bmu_trainer *t) { return t ? t->last_ms : 0.0f; }

void bmu_trainer_destroy(bmu_trainer *t) {
  if (!t) return;
  void *ptrs[] = {t->d_codes, t->d_data, t->d_unit_alpha, t->d_gslice, t->d_valid, t->d_fixed,
                  t->d_code_label, t->d_data_label, t->d_slots};
  int cur = -1;
  cudaGetDevice(&cur);
  if (t->owner && t->owner->dev >= 0 && t->owner->dev != cur) cudaSetDevice(t->owner->dev);
  for (void *q : ptrs) t_free(t, q);
  t_free(t, t->sched_sample.p); t_free(t, t->sched_talp.p); t_free(t, t->sched_trad.p);
  if (t->owner && t->owner->dev >= 0 && t->owner->dev != cur && cur >= 0) cudaSetDevice(cur);
  if (t->ev0) cudaEventDestroy(t->ev0);
  if (t->ev1) cudaEventDestroy(t->ev1);
  delete t;
}

int bmu_som_train(float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                  const float *data, const unsigned char *mask, long N,
                  const int16_t *fixed_xy, const int32_t *sample, const float *talp,
                  const float *trad, long nsteps) {
  bmu_trainer *t = bmu_trainer_create(codes, M, D, data, mask, N);
  if (!t) return BMU_ERR_CUDA;
  int rc = bmu_trainer_set_som(t, xdim, ydim, topol, neigh, fixed_xy);
  if (!rc) rc = bmu_trainer_steps(t, sample, talp, trad, nsteps);
  if (!rc) rc = bmu_trainer_get_codes(t, codes);
  bmu_trainer_destroy(t);
  return rc;
}

int bmu_lvq_train(int algo, float *codes, const int32_t *code_label, long M, int D,
                  const float *data, const unsigned char *mask, const int32_t *data_label,
                  long N, const int32_t *sample, const float *talp, long nsteps,
                  float win_thr, float epsilon, float alpha_cap, float *unit_alpha) {
  bmu_trainer *t = bmu_trainer_create(codes, M, D, data, mask, N);
  if (!t) return BMU_ERR_CUDA;
  int rc = bmu_trainer_set_lvq(t, algo, code_label, data_label, win_thr, epsilon, alpha_cap, unit_alpha);
  if (!rc) rc = bmu_trainer_steps(t, sample, talp, nullptr, nsteps);
  if (!rc) rc = bmu_trainer_get_codes(t, codes);
  if (!rc && algo == BMU_OLVQ1) rc = bmu_trainer_get_unit_alpha(t, unit_alpha);
  bmu_trainer_destroy(t);
  return rc;
}

}  // extern "C"
