/* bmu_glue.c -- the reference-side binding of the B200 BMU engine, written against the REFERENCE'S OWN
 * headers and structs (struct entries / data_entry / winner_info / teach_params, lvq_pak.h:73-124,186-204).
 *
 * This file is what a maintainer of SOM_PAK / LVQ_PAK adds to the package: it is compiled with
 * -I/root/reference next to the reference's unmodified sources (glue/Makefile) and linked with
 * libbmu_b200.so.  Nothing in the reference is edited; three groups of its symbols are displaced at
 * compile time (-Dname=ref_name on the one file that defines them) and re-defined here:
 *
 *   find_winner_euc / find_winner_knn (lvq_pak.c:41-94,152-221)    the WINNER_FUNCTION slots.  The
 *       reference calls them once per sample; here the first call for a sample flattens that sample and
 *       EVERYTHING BEHIND IT IN ITS LIST (->next up to the end of the list, i.e. the rest of the file, or
 *       of the -buffer chunk) into one array, answers all of them with ONE bmu_multi_search, and serves
 *       the following per-sample calls from that batch.  Every consumer of the package walks its data in
 *       list order (find_qerror som_rout.c:710-721, compute_accuracy accuracy.c:82, compute_knnaccuracy
 *       knntest.c:98, compute_visual_data visual.c:113, find_labels vcal.c:109, compute_classifications
 *       classify.c:66, compute_cmatr cmatr.c:84), so their loops stay as they are.
 *   som_training (som_rout.c:556-671)                              entries_flatten -> bmu_som_schedule ->
 *       bmu_trainer_* (cut at the snapshot steps) -> entries_scatter.
 *   lvq1/olvq1/lvq2/lvq3_training (lvq_rout.c:498-916)             the same with bmu_lvq_schedule.
 *
 * The batch answers are valid while the codebook does not change, which holds for the search-only
 * programs; the training loops that do change it are the ones replaced wholesale.  A fingerprint of the
 * code vectors is checked on every batch so that a stale codebook image is never searched.
 */
#include <float.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lvq_pak.h"
#include "datafile.h"
#include "labels.h"
#include "som_rout.h"
#include "lvq_rout.h"

#include "bmu.h"

/* ------------------------------------------------------------------ flat views of the reference's lists */
struct flat {
  long n;
  int dim;
  float *points;               /* n x dim */
  unsigned char *mask;         /* n x dim, or NULL when no entry carries a mask */
  struct data_entry **node;    /* index -> list node (win->winner aliases a node, lvq_pak.c:86) */
};

static void flat_free(struct flat *f) {
  free(f->points); free(f->mask); free(f->node);
  memset(f, 0, sizeof(*f));
}

/* entries_flatten: the nodes from `first` to the end of its list, in list order */
static int entries_flatten(struct data_entry *first, int dim, struct flat *f) {
  long n = 0, i;
  int any_mask = 0;
  struct data_entry *e;
  memset(f, 0, sizeof(*f));
  for (e = first; e != NULL; e = e->next) { n++; if (e->mask) any_mask = 1; }
  f->n = n;
  f->dim = dim;
  f->points = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * dim);
  f->node = (struct data_entry **)malloc(sizeof(struct data_entry *) * (size_t)(n > 0 ? n : 1));
  if (any_mask) f->mask = (unsigned char *)calloc((size_t)(n > 0 ? n : 1) * dim, 1);
  if (!f->points || !f->node || (any_mask && !f->mask)) { flat_free(f); return 1; }
  for (e = first, i = 0; e != NULL; e = e->next, i++) {
    memcpy(f->points + i * dim, e->points, sizeof(float) * dim);
    if (e->mask) memcpy(f->mask + i * dim, e->mask, (size_t)dim);
    f->node[i] = e;
  }
  return 0;
}

/* entries_scatter: the (trained) flat array back into the list nodes */
static void entries_scatter(const struct flat *f) {
  long i;
  for (i = 0; i < f->n; i++) memcpy(f->node[i]->points, f->points + i * f->dim, sizeof(float) * f->dim);
}

static int engine_error(const char *what) {
  fprintf(stderr, "%s: %s\n", what, bmu_last_error());
  return 1;
}

/* ------------------------------------------------------------------ batched WINNER_FUNCTION */
static struct {
  struct entries *codes;       /* codebook the batch was searched against */
  uint64_t codes_print;        /* fingerprint of its vectors at that time */
  struct flat cflat;           /* its nodes (index -> node) */
  bmu_mcodebook *cb;
  int knn;
  struct flat batch;           /* the samples of the batch */
  int32_t *idx, *nfound;
  float *diff;
  long cursor;                 /* consumers ask in list order: next expected sample */
} W;

static uint64_t codes_fingerprint(struct entries *codes) {
  uint64_t h = 1469598103934665603ULL;
  struct data_entry *e;
  int i, dim = codes->dimension;
  for (e = codes->entries; e != NULL; e = e->next)
    for (i = 0; i < dim; i++) {
      uint32_t b;
      memcpy(&b, &e->points[i], 4);
      h = (h ^ b) * 1099511628211ULL;
    }
  return h ^ (uint64_t)(uintptr_t)codes->entries;
}

static int winner_batch(struct entries *codes, struct data_entry *sample, int knn) {
  uint64_t print;
  if (codes->entries == NULL) {               /* not read yet: the reference loads on the first rewind (lvq_pak.c:55) */
    eptr p;
    rewind_entries(codes, &p);
  }
  print = codes_fingerprint(codes);
  if (W.cb == NULL || W.codes != codes || W.codes_print != print) {
    if (W.cb) bmu_mcodebook_destroy(W.cb);
    flat_free(&W.cflat);
    W.cb = NULL;
    if (entries_flatten(codes->entries, codes->dimension, &W.cflat)) return 1;
    {
      /* small searches (the demo recipes) stay on one GPU, large ones use every visible one */
      const char *env = getenv("SOMLVQ_GPUS");
      if (bmu_multi_init((env && atoi(env) > 0) ? atoi(env) : 1)) return engine_error("bmu_multi_init");
    }
    W.cb = bmu_mcodebook_create(W.cflat.points, W.cflat.n, codes->dimension);
    if (!W.cb) return engine_error("bmu_mcodebook_create");
    W.codes = codes;
    W.codes_print = print;
  }
  flat_free(&W.batch);
  free(W.idx); free(W.nfound); free(W.diff);
  W.idx = W.nfound = NULL; W.diff = NULL;
  if (entries_flatten(sample, codes->dimension, &W.batch)) return 1;
  W.idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)W.batch.n * knn);
  W.diff = (float *)malloc(sizeof(float) * (size_t)W.batch.n * knn);
  W.nfound = (int32_t *)malloc(sizeof(int32_t) * (size_t)W.batch.n);
  if (!W.idx || !W.diff || !W.nfound) return 1;
  if (bmu_multi_search(W.cb, W.batch.points, W.batch.mask, W.batch.n, knn, W.idx, W.diff, W.nfound, NULL))
    return engine_error("bmu_multi_search");
  W.knn = knn;
  W.cursor = 0;
  return 0;
}

static int winner_from_batch(struct entries *codes, struct data_entry *sample, struct winner_info *win, int knn) {
  long i = -1, t;
  if (W.cb && W.codes == codes && W.knn == knn && W.batch.n > 0) {
    if (W.cursor < W.batch.n && W.batch.node[W.cursor] == sample) i = W.cursor;
    else
      for (t = 0; t < W.batch.n; t++)
        if (W.batch.node[t] == sample) { i = t; break; }
    /* the node is known, but is it still the vector that was searched?  (-buffer re-uses nothing, a
     * consumer that edits its samples would be caught here) */
    if (i >= 0 && memcmp(W.batch.points + i * W.batch.dim, sample->points, sizeof(float) * W.batch.dim) != 0) i = -1;
    /* cheap per-call check of the codebook (its list head); the full fingerprint is taken per batch */
    if (i >= 0 && (W.cflat.n == 0 || codes->entries != W.cflat.node[0])) i = -1;
  }
  if (i < 0) {
    if (winner_batch(codes, sample, knn)) {
      fprintf(stderr, "bmu_glue: batch winner search failed\n");
      exit(1);                                        /* no CPU fallback */
    }
    i = 0;
  }
  W.cursor = i + 1;
  for (t = 0; t < knn; t++) {
    const int32_t j = W.idx[i * knn + t];
    win[t].index = j;
    win[t].winner = j >= 0 ? W.cflat.node[j] : NULL;
    win[t].diff = W.diff[i * knn + t];
  }
  return W.nfound[i];
}

/* the two WINNER_FUNCTION slots of struct teach_params (set_teach_params datafile.c:1248-1282 and the
 * per-program overrides knntest.c:206, lvqtrain.c:224,228 take their addresses) */
int find_winner_euc(struct entries *codes, struct data_entry *sample, struct winner_info *win, int knn) {
  (void)knn;                                           /* lvq_pak.c:41-94 always looks for one winner */
  return winner_from_batch(codes, sample, win, 1);
}
int find_winner_knn(struct entries *codes, struct data_entry *sample, struct winner_info *win, int knn) {
  if (knn < 1) knn = 1;
  return winner_from_batch(codes, sample, win, knn);   /* knn == 1: the 1-NN rule, lvq_pak.c:160-161 */
}

/* ------------------------------------------------------------------ training loops */
/* the whole data set in list order (loads the file if it has not been read yet, datafile.c:789-832) */
static int flatten_all(struct entries *set, struct flat *f, const char *who) {
  eptr p;
  struct data_entry *first = rewind_entries(set, &p);
  if (first == NULL) { fprintf(stderr, "%s: can't get data\n", who); return 1; }
  if (set->flags.loadmode == LOADMODE_BUFFER) {
    fprintf(stderr, "%s: -buffer is not supported for training on the B200 engine (load the file whole)\n", who);
    return 1;
  }
  return entries_flatten(first, set->dimension, f);
}

/* steps [le0, le1) of a run; snapshots are written after the steps `le % interval == 0 && le > 0`
 * (som_rout.c:650, lvq_rout.c:562), so the run is cut right behind every such step */
static long next_cut(const struct snapshot_info *snap, long le0, long length) {
  long le;
  if (!snap || snap->interval <= 0) return length;
  le = (le0 + snap->interval - 1) / snap->interval * snap->interval;     /* first snapshot step >= le0 ... */
  if (le == 0) le = snap->interval;                                        /* ... that is > 0 */
  return le + 1 < length ? le + 1 : length;
}

struct entries *som_training(struct teach_params *teach) {
  struct entries *data = teach->data, *codes = teach->codes;
  struct snapshot_info *snap = teach->snapshot;
  const long length = teach->length;
  struct flat fd, fc;
  int16_t *weight = NULL, *fixed = NULL;
  int32_t *sample;
  float *talp, *trad;
  bmu_trainer *t;
  long i, le0;

  if (set_som_params(teach)) { fprintf(stderr, "som_training: can't set SOM parameters\n"); return NULL; }
  if (flatten_all(data, &fd, "som_training")) return NULL;
  if (data->dimension != codes->dimension) {
    fprintf(stderr, "code dimension (%d) != data dimension (%d)\n", codes->dimension, data->dimension);
    return NULL;
  }
  {
    eptr p;
    rewind_entries(codes, &p);                             /* make sure the map is loaded (lvq_pak.c:55) */
  }
  if (entries_flatten(codes->entries, codes->dimension, &fc)) return NULL;
  if (use_weights(-1)) {                                                   /* som_rout.c:622-624 */
    weight = (int16_t *)malloc(sizeof(int16_t) * (size_t)fd.n);
    for (i = 0; i < fd.n; i++) weight[i] = fd.node[i]->weight;
  }
  if (use_fixed(-1)) {                                                     /* som_rout.c:628-632 */
    fixed = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)fd.n);
    for (i = 0; i < fd.n; i++) {
      fixed[2 * i] = fd.node[i]->fixed ? fd.node[i]->fixed->xfix : -1;
      fixed[2 * i + 1] = fd.node[i]->fixed ? fd.node[i]->fixed->yfix : -1;
    }
  }
  sample = (int32_t *)malloc(sizeof(int32_t) * (size_t)(length > 0 ? length : 1));
  talp = (float *)malloc(sizeof(float) * (size_t)(length > 0 ? length : 1));
  trad = (float *)malloc(sizeof(float) * (size_t)(length > 0 ? length : 1));
  /* the list IS the sample order (a -rand shuffle happened when the file was read, datafile.c:340-341) */
  bmu_som_schedule(0, length, length, teach->alpha, teach->radius, teach->alpha_type, fd.n, NULL, weight, sample,
                   talp, trad);
  time(&teach->start_time);
  t = bmu_trainer_create(fc.points, fc.n, fc.dim, fd.points, fd.mask, fd.n);
  if (!t || bmu_trainer_set_som(t, codes->xdim, codes->ydim, teach->topol, teach->neigh, fixed)) {
    engine_error("som_training");
    return NULL;
  }
  for (le0 = 0; le0 < length;) {
    const long le1 = next_cut(snap, le0, length);
    if (bmu_trainer_steps(t, sample + le0, talp + le0, trad + le0, le1 - le0)) { engine_error("som_training"); return NULL; }
    if (snap && le1 - 1 > 0 && ((le1 - 1) % snap->interval) == 0) {
      if (bmu_trainer_get_codes(t, fc.points)) { engine_error("som_training"); return NULL; }
      entries_scatter(&fc);
      if (save_snapshot(teach, le1 - 1)) fprintf(stderr, "snapshot failed, continuing teaching\n");
    }
    le0 = le1;
  }
  if (bmu_trainer_get_codes(t, fc.points)) { engine_error("som_training"); return NULL; }
  bmu_trainer_destroy(t);
  entries_scatter(&fc);
  time(&teach->end_time);
  free(sample); free(talp); free(trad); free(weight); free(fixed);
  flat_free(&fd); flat_free(&fc);
  return codes;
}

static struct entries *lvq_run(struct teach_params *teach, int algo, float winlen, float epsilon, char *infile,
                               char *outfile, const char *who) {
  struct entries *data = teach->data, *codes = teach->codes;
  struct snapshot_info *snap = teach->snapshot;
  const long length = teach->length;
  struct flat fd, fc;
  int32_t *sample, *code_label, *data_label;
  float *talp, *unit_alpha = NULL;
  float alpha = teach->alpha, win_thr = 0.0f;
  bmu_trainer *t;
  eptr p;
  long i, le0;

  rewind_entries(codes, &p);                              /* make sure codes are loaded (lvq_rout.c:607) */
  if (flatten_all(data, &fd, who)) return NULL;
  if (entries_flatten(codes->entries, codes->dimension, &fc)) return NULL;
  code_label = (int32_t *)malloc(sizeof(int32_t) * (size_t)fc.n);
  data_label = (int32_t *)malloc(sizeof(int32_t) * (size_t)fd.n);
  for (i = 0; i < fc.n; i++) code_label[i] = get_entry_label(fc.node[i]);     /* the FIRST label, labels.h:45 */
  for (i = 0; i < fd.n; i++) data_label[i] = get_entry_label(fd.node[i]);
  if (algo == BMU_OLVQ1) {                                /* lvq_rout.c:609-627 */
    unit_alpha = (float *)malloc(sizeof(float) * (size_t)fc.n);
    if (alpha == 0.0) {
      if (!alpha_read(unit_alpha, fc.n, infile)) {
        alpha = 0.3;
        for (i = 0; i < fc.n; i++) unit_alpha[i] = alpha;
      }
    } else {
      for (i = 0; i < fc.n; i++) unit_alpha[i] = alpha;
    }
  }
  if (algo == BMU_LVQ2 || algo == BMU_LVQ3) win_thr = (1 - winlen) / (1 + winlen);    /* lvq_rout.c:770,876 */
  sample = (int32_t *)malloc(sizeof(int32_t) * (size_t)(length > 0 ? length : 1));
  talp = (float *)malloc(sizeof(float) * (size_t)(length > 0 ? length : 1));
  bmu_lvq_schedule(0, length, length, alpha, teach->alpha_type, fd.n, NULL, sample, talp);
  t = bmu_trainer_create(fc.points, fc.n, fc.dim, fd.points, fd.mask, fd.n);
  if (!t || bmu_trainer_set_lvq(t, algo, code_label, data_label, win_thr, epsilon, alpha, unit_alpha)) {
    engine_error(who);
    return NULL;
  }
  for (le0 = 0; le0 < length;) {
    const long le1 = next_cut(snap, le0, length);
    if (bmu_trainer_steps(t, sample + le0, algo == BMU_OLVQ1 ? NULL : talp + le0, NULL, le1 - le0)) { engine_error(who); return NULL; }
    if (snap && le1 - 1 > 0 && ((le1 - 1) % snap->interval) == 0) {
      if (bmu_trainer_get_codes(t, fc.points)) { engine_error(who); return NULL; }
      entries_scatter(&fc);
      if (save_snapshot(teach, le1 - 1)) fprintf(stderr, "snapshot failed\n");
    }
    le0 = le1;
  }
  if (bmu_trainer_get_codes(t, fc.points)) { engine_error(who); return NULL; }
  if (algo == BMU_OLVQ1) {
    if (bmu_trainer_get_unit_alpha(t, unit_alpha)) { engine_error(who); return NULL; }
    alpha_write(unit_alpha, fc.n, outfile);               /* lvq_rout.c:694 */
  }
  bmu_trainer_destroy(t);
  entries_scatter(&fc);
  free(sample); free(talp); free(code_label); free(data_label); free(unit_alpha);
  flat_free(&fd); flat_free(&fc);
  return codes;
}

struct entries *lvq1_training(struct teach_params *teach) {
  return lvq_run(teach, BMU_LVQ1, 0.0f, 0.0f, NULL, NULL, "lvq1_training");
}
struct entries *olvq1_training(struct teach_params *teach, char *infile, char *outfile) {
  return lvq_run(teach, BMU_OLVQ1, 0.0f, 0.0f, infile, outfile, "olvq1_training");
}
struct entries *lvq2_training(struct teach_params *teach, float winlen) {
  return lvq_run(teach, BMU_LVQ2, winlen, 0.0f, NULL, NULL, "lvq2_training");
}
struct entries *lvq3_training(struct teach_params *teach, float epsilon, float winlen) {
  return lvq_run(teach, BMU_LVQ3, winlen, epsilon, NULL, NULL, "lvq3_training");
}
