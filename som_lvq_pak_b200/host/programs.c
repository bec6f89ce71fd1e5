/* programs.c -- the SOM_PAK / LVQ_PAK programs whose loops sit on the BMU hot path, re-hosted on
 * libbmu_b200 (include/bmu.h): same options, same stdout, same output files.
 *
 *   reference per-sample loop                         here
 *   find_qerror / find_qerror2  som_rout.c:678-891    bmu_search + bmu_replay_qerror / bmu_qerror2
 *   compute_visual_data         visual.c:48-155       bmu_search, rows written in data order
 *   find_labels                 vcal.c:45-167         bmu_search, per-unit hitlists replayed in data order
 *   compute_accuracy            accuracy.c:39-137     bmu_search, hitlists replayed in data order
 *   compute_classifications     classify.c:41-95      bmu_search
 *   compute_knnaccuracy         knntest.c:41-157      bmu_search (k), majority vote = head of the hitlist
 *   compute_cmatr               cmatr.c:41-171        bmu_search, confusion counts replayed in data order
 *   find_labels (setlabel)      setlabel.c:41-96      bmu_search (k) with the data set as the searched set
 *   eliminate_codes             elimin.c:42-126       bmu_search (k) of the data set against itself
 *   som_training                som_rout.c:556-671    bmu_som_schedule + bmu_som_train
 *   lvq1/olvq1/lvq2/lvq3        lvq_rout.c:498-916    bmu_lvq_schedule + bmu_lvq_train
 *   vfind trials                vfind.c:247-306       bmu_randinit_codes + 2 x bmu_som_train + qerror, per trial
 *   pick_inside_codes           lvq_rout.c:151-211    bmu_search (k) of the data set against itself (eveninit / propinit)
 *
 * -buffer N: qerror and accuracy stream the data file in chunks of N entries (the next chunk is parsed
 * while the current one is searched); vsom / lvq1..3 / olvq1 reproduce the reference's chunk-wise -rand
 * order (bmu_sample_sequence); the other programs read the file whole -- for them -buffer changes no output.
 * Not carried over (outside SURVEY.md section 8): -selfuncs, background (forked) snapshots.
 * Compressed (.gz/.Z) and piped (|cmd) file names ARE handled by the file layer (entries.c).
 */
#include "somhost.h"

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/bmu.h"

/* ------------------------------------------------------------------ options */
static const char *opt(int argc, char **argv, const char *name) {       /* lvq_pak.c:583-612, OPTION */
  int i;
  for (i = 0; i < argc - 1; i++)
    if (strcmp(argv[i], name) == 0) return argv[i + 1];
  return NULL;
}
static const char *need(int argc, char **argv, const char *name) {      /* ALWAYS */
  const char *v = opt(argc, argv, name);
  if (!v) { fprintf(stderr, "Can't find asked option %s\n", name); exit(-1); }
  return v;
}
static int flag(int argc, char **argv, const char *name) {              /* OPTION2 */
  int i;
  for (i = 0; i < argc; i++)
    if (strcmp(argv[i], name) == 0) return 1;
  return 0;
}
static int verbose_level = 1;
static void global_options(int argc, char **argv) {                     /* lvq_pak.c:618-661 */
  const char *s = getenv("LVQSOM_MASK_STR");
  pak_mask_string = "x";                                             /* `bmu_pak batch` runs several programs */
  if (s) pak_mask_string = s;
  s = opt(argc, argv, "-mask_str");
  if (s) pak_mask_string = s;
  s = opt(argc, argv, "-v");
  verbose_level = s ? atoi(s) : 1;
  if (opt(argc, argv, "-selfuncs"))
    fprintf(stderr, "note: -selfuncs is not supported by the B200 host; ignored\n");
}

static void lra_name(const char *codefile, char *out, size_t outsz);

static int engine_failed(const char *what) {
  fprintf(stderr, "%s: %s\n", what, bmu_last_error());
  return 1;
}

/* winners of every data vector: idx, squared diff, return value of the reference's winner function */
struct winners {
  int32_t *idx, *nfound;
  float *diff;
};
static void winners_free(struct winners *w) { free(w->idx); free(w->nfound); free(w->diff); }
/* One funnel for every program's winner loop (find_qerror som_rout.c:710-721, accuracy.c:82, classify.c:66,
 * knntest.c:98, vcal.c:109, visual.c:113, cmatr.c:84, elimin.c:81, setlabel.c:73): the rows are sharded over
 * every GPU the process can see ($SOMLVQ_GPUS limits them), results come back in data order. */
static int find_winners(const struct pak_entries *codes, const struct pak_entries *data, int knn, struct winners *w) {
  bmu_mcodebook *cb;
  int rc;
  size_t n = (size_t)(data->n > 0 ? data->n : 1);
  w->idx = (int32_t *)malloc(sizeof(int32_t) * n * knn);
  w->diff = (float *)malloc(sizeof(float) * n * knn);
  w->nfound = (int32_t *)malloc(sizeof(int32_t) * n);
  if (!w->idx || !w->diff || !w->nfound) { fprintf(stderr, "out of memory\n"); return 1; }
  if (data->n == 0) return 0;
  {
    /* Opening every device and an NCCL communicator costs one to two seconds; one B200 streams 54 GB/s of rows over
     * PCIe and searches 4e14 distance elements a second.  The rows are sharded over all visible GPUs only when one
     * GPU would need longer than that; $SOMLVQ_GPUS overrides. */
    const char *env = getenv("SOMLVQ_GPUS");
    const double work = (double)data->n * (double)codes->n * (double)codes->dim;
    const double bytes = (double)data->n * (double)codes->dim * 4.0;
    const double one_gpu_s = work / 4e14 > bytes / 5e10 ? work / 4e14 : bytes / 5e10;
    const int shards = (env && atoi(env) > 0) ? atoi(env) : (one_gpu_s < 2.0 ? 1 : 0);
    if (bmu_multi_init(shards)) return engine_failed("bmu_multi_init");
  }
  cb = bmu_mcodebook_create(codes->points, codes->n, codes->dim);
  if (!cb) return engine_failed("bmu_mcodebook_create");
  rc = bmu_multi_search(cb, data->points, data->mask, data->n, knn, w->idx, w->diff, w->nfound, NULL);
  bmu_mcodebook_destroy(cb);
  return rc ? engine_failed("bmu_multi_search") : 0;
}

static int open_pair(int argc, char **argv, int data_labels_needed, int code_labels_needed, int skip_empty,
                     int want_map, struct pak_entries **data, struct pak_entries **codes) {
  const char *din = need(argc, argv, "-din"), *cin = need(argc, argv, "-cin");
  *data = pak_load(din, data_labels_needed, skip_empty);
  if (!*data) { fprintf(stderr, "Can't open data file '%s'\n", din); return 1; }
  *codes = pak_load(cin, code_labels_needed, 1);
  if (!*codes) { fprintf(stderr, "Can't open code file '%s'\n", cin); return 1; }
  if (want_map && (*codes)->topol < TOPOL_HEXA) { fprintf(stderr, "File %s is not a map file\n", cin); return 1; }
  if ((*data)->dim != (*codes)->dim) {
    fprintf(stderr, "Data and codebook vectors have different dimensions (%d != %d)", (*data)->dim, (*codes)->dim);
    return 1;
  }
  if (bmu_init(0)) return engine_failed("bmu_init");
  return 0;
}

/* ------------------------------------------------------------------ qerror */
/* data file opened as a stream of chunks (`-buffer N`, datafile.c:237-344; without it one chunk = the whole
 * file) plus the code file, checked like open_pair */
static int open_stream_pair(int argc, char **argv, int data_labels_needed, int code_labels_needed, int skip_empty,
                            int want_map, struct pak_stream **ds, struct pak_entries **codes) {
  const char *din = need(argc, argv, "-din"), *cin = need(argc, argv, "-cin"), *b = opt(argc, argv, "-buffer");
  *ds = pak_stream_open(din, data_labels_needed, skip_empty, b ? atol(b) : 0);
  if (!*ds) { fprintf(stderr, "Can't open data file '%s'\n", din); return 1; }
  *codes = pak_load(cin, code_labels_needed, 1);
  if (!*codes) { fprintf(stderr, "Can't open code file '%s'\n", cin); return 1; }
  if (want_map && (*codes)->topol < TOPOL_HEXA) { fprintf(stderr, "File %s is not a map file\n", cin); return 1; }
  if (pak_stream_header(*ds)->dim != (*codes)->dim) {
    fprintf(stderr, "Data and codebook vectors have different dimensions (%d != %d)", pak_stream_header(*ds)->dim,
            (*codes)->dim);
    return 1;
  }
  if (bmu_init(0)) return engine_failed("bmu_init");
  return 0;
}

int qerror_main(int argc, char **argv) {
  struct pak_stream *ds = NULL;
  struct pak_entries *data = NULL, *codes = NULL;
  const char *s;
  float radius, qerror = 0.0f;
  long total = 0;
  int qmode;
  bmu_codebook *cb = NULL;
  global_options(argc, argv);
  s = opt(argc, argv, "-radius");
  radius = s ? (float)atof(s) : 1.0f;
  s = opt(argc, argv, "-qetype");
  qmode = s ? atoi(s) : 0;
  if (open_stream_pair(argc, argv, 0, 0, 1, 1, &ds, &codes)) return 1;
  /* chunk by chunk (one chunk without -buffer): the next chunk is being read and parsed while this one is
   * searched; the accumulator is ONE float carried across the chunks in data order (som_rout.c:697,715) */
  while ((data = pak_stream_next(ds)) != NULL) {
    long i;
    if (qmode > 0) {
      /* find_qerror2: per-sample neighbourhood-weighted error on the GPU, added in data order */
      float *per = (float *)malloc(sizeof(float) * (size_t)data->n);
      if (!cb) cb = bmu_codebook_create(codes->points, codes->n, codes->dim);
      if (!per || !cb) return engine_failed("bmu_codebook_create");
      if (bmu_qerror2(cb, codes->xdim, codes->ydim, codes->topol, codes->neigh, radius, data->points, data->mask,
                      data->n, per))
        return engine_failed("bmu_qerror2");
      for (i = 0; i < data->n; i++) qerror += per[i];                 /* som_rout.c:872 */
      free(per);
    } else {
      struct winners w;
      if (find_winners(codes, data, 1, &w)) return 1;
      for (i = 0; i < data->n; i++)                                   /* som_rout.c:712-715 */
        if (w.nfound[i] != 0) qerror += sqrt((double)w.diff[i]);
      winners_free(&w);
    }
    total += data->n;
    pak_free(data);
  }
  if (pak_stream_failed(ds)) { fprintf(stderr, "Can't read data file '%s'\n", need(argc, argv, "-din")); return 1; }
  if (cb) bmu_codebook_destroy(cb);
  if (verbose_level >= 1)                                          /* qerror.c:114-118 */
    fprintf(stdout, "Quantization error of %s with map %s is %f per sample (%ld samples)\n",
            need(argc, argv, "-din"), need(argc, argv, "-cin"), qerror / (float)total, total);
  else
    fprintf(stdout, "%f\n", qerror / (float)total);
  pak_stream_close(ds);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ visual */
int visual_main(int argc, char **argv) {
  struct pak_entries *data = NULL, *codes = NULL, head;
  struct winners w;
  const char *dout;
  FILE *fp;
  long i, l;
  int emptylab;
  global_options(argc, argv);
  dout = need(argc, argv, "-dout");
  if (open_pair(argc, argv, 0, 0, !flag(argc, argv, "-noskip"), 1, &data, &codes)) return 1;
  emptylab = label_index("EMPTY_LINE");                            /* visual.c:67 */
  if (find_winners(codes, data, 1, &w)) return 1;
  fp = fopen(dout, "w");
  if (!fp) { fprintf(stderr, "can't open file for output: '%s'\n", dout); return 1; }
  head = *codes;
  head.dim = 3;
  pak_write_header(fp, &head);                                     /* visual.c:74-100 */
  for (i = 0; i < data->n; i++) {
    if (w.nfound[i] == 0) {                                        /* visual.c:113-122 */
      fprintf(fp, "%g %g %g %s \n", -1.0f, -1.0f, -1.0f, label_string(emptylab));
      continue;
    }
    {
      const long b = w.idx[i];
      const float x = (float)(b % codes->xdim), y = (float)(b / codes->xdim);
      const float qe = (float)sqrt((double)w.diff[i]);             /* visual.c:132 */
      fprintf(fp, "%g %g %g ", x, y, qe);
      for (l = codes->lab_off[b]; l < codes->lab_off[b + 1]; l++) fprintf(fp, "%s ", label_string(codes->lab_pool[l]));
      fprintf(fp, "\n");
    }
  }
  fclose(fp);
  winners_free(&w);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ vcal */
int vcal_main(int argc, char **argv) {
  struct pak_entries *data = NULL, *codes = NULL;
  struct winners w;
  struct pak_hitlist *hits;
  const char *s, *cout_name;
  int numlabs, *nlab, *labs;
  long i, tot = 0, pos = 0;
  global_options(argc, argv);
  cout_name = need(argc, argv, "-cout");
  s = opt(argc, argv, "-numlabs");
  numlabs = s ? atoi(s) : 1;
  if (numlabs < 0) numlabs = 0;
  /* vcal.c:196-205: the data file is opened while labels are still required */
  if (open_pair(argc, argv, 1, 0, 1, 1, &data, &codes)) return 1;
  if (find_winners(codes, data, 1, &w)) return 1;
  hits = (struct pak_hitlist *)malloc(sizeof(*hits) * (size_t)codes->n);
  nlab = (int *)malloc(sizeof(int) * (size_t)codes->n);
  if (!hits || !nlab) return 1;
  for (i = 0; i < codes->n; i++) hit_init(&hits[i]);
  for (i = 0; i < data->n; i++) {                                  /* vcal.c:104-118: data order */
    const int lab = pak_label(data, i);
    if (w.nfound[i] == 0 || lab == LABEL_EMPTY) continue;
    hit_add(&hits[w.idx[i]], lab);
  }
  for (i = 0; i < codes->n; i++) {                                 /* vcal.c:140-160 */
    nlab[i] = (int)(numlabs == 0 ? hits[i].n : (hits[i].n < numlabs ? hits[i].n : numlabs));
    tot += nlab[i];
  }
  labs = (int *)malloc(sizeof(int) * (size_t)(tot > 0 ? tot : 1));
  if (!labs) return 1;
  for (i = 0; i < codes->n; i++) {
    int t;
    for (t = 0; t < nlab[i]; t++) labs[pos++] = (int)hits[i].label[t];
    hit_free(&hits[i]);
  }
  if (pak_set_labels(codes, nlab, labs)) return 1;
  pak_save(codes, cout_name);
  free(hits); free(nlab); free(labs);
  winners_free(&w);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ accuracy / knntest */
static void print_accuracy(const struct pak_hitlist *totals, const struct pak_hitlist *correct, long total,
                           long stotal, int knn_style) {
  long i;
  fprintf(stdout, "\nRecognition accuracy:\n\n");
  for (i = 0; i < totals->n; i++) {
    const long tot = totals->freq[i], res = hit_freq(correct, totals->label[i]);
    if (knn_style) {                                               /* knntest.c:133-142 */
      fprintf(stdout, "%14s: ", label_string((int)totals->label[i]));
      fprintf(stdout, "%6.2f %%\n", 100.0 * (float)res / tot);
    } else {                                                       /* accuracy.c:118-127 */
      fprintf(stdout, "%9s: %4ld entries ", label_string((int)totals->label[i]), tot);
      fprintf(stdout, "%6.2f %%\n", 100.0 * (float)res / tot);
    }
  }
  if (knn_style) fprintf(stdout, "\nTotal accuracy: %6.2f %%\n\n", 100.0 * (float)stotal / total);
  else fprintf(stdout, "\nTotal accuracy: %5ld entries %6.2f %%\n\n", total, 100.0 * (float)stotal / total);
}

int accuracy_main(int argc, char **argv) {
  struct pak_stream *ds = NULL;
  struct pak_entries *data = NULL, *codes = NULL;
  struct pak_hitlist correct, totals;
  const char *cfout;
  FILE *ocf = NULL;
  long i, total = 0, stotal = 0;
  global_options(argc, argv);
  cfout = opt(argc, argv, "-cfout");
  if (open_stream_pair(argc, argv, 1, 1, 1, 0, &ds, &codes)) return 1;
  if (cfout && !(ocf = fopen(cfout, "w"))) { fprintf(stderr, "Cannot open '%s' for output\n", cfout); return 1; }
  hit_init(&correct);
  hit_init(&totals);
  while ((data = pak_stream_next(ds)) != NULL) {                     /* one chunk unless -buffer N */
    struct winners w;
    if (find_winners(codes, data, 1, &w)) return 1;
    for (i = 0; i < data->n; i++) {                                  /* accuracy.c:80-105 */
      const int datalabel = pak_label(data, i);
      const int winlabel = w.idx[i] >= 0 ? pak_label(codes, w.idx[i]) : -1;
      if (winlabel == datalabel) {
        stotal++;
        hit_add(&correct, datalabel);
        if (ocf) fprintf(ocf, "1\n");
      } else if (ocf) {
        fprintf(ocf, "0\n");
      }
      hit_add(&totals, datalabel);
      total++;
    }
    winners_free(&w);
    pak_free(data);
  }
  if (pak_stream_failed(ds)) { fprintf(stderr, "Can't read data file '%s'\n", need(argc, argv, "-din")); return 1; }
  print_accuracy(&totals, &correct, total, stotal, 0);
  if (ocf) fclose(ocf);
  hit_free(&correct); hit_free(&totals);
  pak_stream_close(ds);
  pak_free(codes);
  return 0;
}

int knntest_main(int argc, char **argv) {
  struct pak_entries *data = NULL, *codes = NULL;
  struct winners w;
  struct pak_hitlist hits, correct, totals;
  const char *s;
  long i, total = 0, stotal = 0;
  int knn, t;
  global_options(argc, argv);
  s = opt(argc, argv, "-knn");
  knn = s ? atoi(s) : 5;                                           /* knntest.c:177 */
  if (knn < 1) knn = 1;
  if (knn > BMU_KMAX) { fprintf(stderr, "-knn %d is larger than the engine's limit %d\n", knn, BMU_KMAX); return 1; }
  if (open_pair(argc, argv, 1, 1, 1, 0, &data, &codes)) return 1;
  if (find_winners(codes, data, knn, &w)) return 1;
  hit_init(&hits); hit_init(&correct); hit_init(&totals);
  for (i = 0; i < data->n; i++) {                                  /* knntest.c:96-122 */
    const int datalabel = pak_label(data, i);
    hit_clear(&hits);
    for (t = 0; t < knn; t++) {
      const int j = w.idx[i * knn + t];
      if (j >= 0) hit_add(&hits, pak_label(codes, j));
    }
    if (hits.n > 0 && hits.label[0] == datalabel) {
      stotal++;
      hit_add(&correct, datalabel);
    }
    hit_add(&totals, datalabel);
    total++;
  }
  print_accuracy(&totals, &correct, total, stotal, 1);
  hit_free(&hits); hit_free(&correct); hit_free(&totals);
  winners_free(&w);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ cmatr */
int cmatr_main(int argc, char **argv) {                              /* cmatr.c:41-171 */
  struct pak_entries *data = NULL, *codes = NULL;
  struct winners w;
  struct pak_hitlist correct, totals, confusion;
  const char *cfout;
  FILE *ocf = NULL;
  long i, j, total = 0, stotal = 0;
  global_options(argc, argv);
  cfout = opt(argc, argv, "-cfout");
  if (open_pair(argc, argv, 1, 1, 1, 0, &data, &codes)) return 1;
  if (cfout && !(ocf = fopen(cfout, "w"))) { fprintf(stderr, "Cannot open '%s' for output\n", cfout); return 1; }
  if (find_winners(codes, data, 1, &w)) return 1;
  hit_init(&correct); hit_init(&totals); hit_init(&confusion);
  for (i = 0; i < data->n; i++) {
    const int datalabel = pak_label(data, i);
    int label;
    if (w.nfound[i] == 0) continue;                                  /* invalid data vector */
    label = pak_label(codes, w.idx[i]);
    if (label == datalabel) {
      stotal++;
      hit_add(&correct, datalabel);
      if (ocf) fprintf(ocf, "1\n");
    } else if (ocf) {
      fprintf(ocf, "0\n");
    }
    hit_add(&confusion, (long)datalabel * 65536 + label);
    hit_add(&totals, datalabel);
    total++;
  }
  print_accuracy(&totals, &correct, total, stotal, 0);
  fprintf(stdout, "Confusion matrix:\n\n");
  fprintf(stdout, "          ");
  for (i = 0; i < totals.n; i++) fprintf(stdout, " %4s", label_string((int)totals.label[i]));
  fprintf(stdout, "\n\n");
  for (i = 0; i < totals.n; i++) {
    fprintf(stdout, "%9s: ", label_string((int)totals.label[i]));
    for (j = 0; j < totals.n; j++)
      fprintf(stdout, "%4ld ", hit_freq(&confusion, totals.label[i] * 65536 + totals.label[j]));
    fprintf(stdout, "\n");
  }
  fprintf(stdout, "\n");
  if (ocf) fclose(ocf);
  hit_free(&correct); hit_free(&totals); hit_free(&confusion);
  winners_free(&w);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ setlabel */
/* labels every codebook vector by the majority among its knn nearest DATA vectors
 * (setlabel.c:41-96): the roles are swapped, the data set is the searched "codebook" */
int setlabel_main(int argc, char **argv) {
  struct pak_entries *data = NULL, *codes = NULL;
  struct winners w;
  struct pak_hitlist hits;
  const char *s, *cout_name;
  int knn, t, *nlab, *labs;
  long i, pos = 0;
  global_options(argc, argv);
  cout_name = need(argc, argv, "-cout");
  s = opt(argc, argv, "-knn");
  knn = s ? atoi(s) : 5;
  if (knn < 1) knn = 1;
  if (knn > BMU_KMAX) { fprintf(stderr, "-knn %d is larger than the engine's limit %d\n", knn, BMU_KMAX); return 1; }
  if (open_pair(argc, argv, 1, 0, 1, 0, &data, &codes)) return 1;
  if (find_winners(data, codes, knn, &w)) return 1;                  /* queries = code vectors */
  nlab = (int *)malloc(sizeof(int) * (size_t)(codes->n > 0 ? codes->n : 1));
  labs = (int *)malloc(sizeof(int) * (size_t)(codes->n > 0 ? codes->n : 1));
  if (!nlab || !labs) return 1;
  hit_init(&hits);
  for (i = 0; i < codes->n; i++) {
    hit_clear(&hits);
    for (t = 0; t < knn; t++) {
      const int j = w.idx[i * knn + t];
      if (j >= 0) hit_add(&hits, pak_label(data, j));
    }
    nlab[i] = (hits.n > 0 && hits.label[0] != LABEL_EMPTY) ? 1 : 0;   /* set_entry_label, labels.c:146-155 */
    if (nlab[i]) labs[pos++] = (int)hits.label[0];
  }
  if (pak_set_labels(codes, nlab, labs)) return 1;
  pak_save(codes, cout_name);
  hit_free(&hits);
  free(nlab); free(labs);
  winners_free(&w);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ elimin */
/* keeps the entries whose knn nearest neighbours in the SAME set (itself included) are mostly of
 * its own class (elimin.c:42-126) */
int elimin_main(int argc, char **argv) {
  struct pak_entries *data, *out;
  struct winners w;
  const char *s, *din, *cout_name;
  int knn, t;
  long i, kept = 0, nl = 0;
  char lra[2048];
  global_options(argc, argv);
  din = need(argc, argv, "-din");
  cout_name = need(argc, argv, "-cout");
  s = opt(argc, argv, "-knn");
  knn = s ? atoi(s) : 5;
  if (knn > 10) { fprintf(stderr, "Can use only %d neighbors", 10); knn = 10; }   /* elimin.c:51-54 */
  if (knn < 1) knn = 1;
  data = pak_load(din, 1, 1);
  if (!data) { fprintf(stderr, "Can't open data file '%s'\n", din); return 1; }
  if (bmu_init(0)) return engine_failed("bmu_init");
  if (find_winners(data, data, knn, &w)) return 1;
  out = pak_alloc(data->dim, data->n);
  if (!out) return 1;
  out->topol = data->topol; out->neigh = data->neigh; out->xdim = data->xdim; out->ydim = data->ydim;
  if (data->mask) out->mask = (unsigned char *)calloc((size_t)data->n * data->dim, 1);
  out->lab_pool = (int *)malloc(sizeof(int) * (size_t)(data->lab_off[data->n] > 0 ? data->lab_off[data->n] : 1));
  if (!out->lab_pool || (data->mask && !out->mask)) return 1;
  for (i = 0; i < data->n; i++) {
    long correct = 0, incorrect = 0, l;
    const int datalabel = pak_label(data, i);
    if (w.nfound[i] != knn) continue;                                /* did not find winners */
    for (t = 0; t < knn; t++) {
      const int j = w.idx[i * knn + t];
      if (j >= 0 && pak_label(data, j) == datalabel) correct++;
      else incorrect++;
    }
    if (correct <= incorrect) continue;
    memcpy(out->points + (size_t)kept * data->dim, data->points + (size_t)i * data->dim, sizeof(float) * data->dim);
    if (data->mask) memcpy(out->mask + (size_t)kept * data->dim, data->mask + (size_t)i * data->dim, (size_t)data->dim);
    for (l = data->lab_off[i]; l < data->lab_off[i + 1]; l++) out->lab_pool[nl++] = data->lab_pool[l];
    kept++;
    out->lab_off[kept] = nl;
  }
  out->n = kept;
  pak_save(out, cout_name);
  lra_name(cout_name, lra, sizeof lra);                              /* elimin.c:209 invalidate_alphafile */
  {
    FILE *fp = fopen(lra, "r");
    if (fp) {
      if (verbose_level >= 1) fprintf(stdout, "Removing the learning rate file %s\n", lra);
      fclose(fp);
      if (remove(lra)) fprintf(stderr, "Can not remove %s", lra);
    }
  }
  winners_free(&w);
  pak_free(data);
  pak_free(out);
  return 0;
}

/* ------------------------------------------------------------------ classify */
int classify_main(int argc, char **argv) {
  struct pak_entries *data = NULL, *codes = NULL;
  struct winners w;
  const char *cfout, *dout;
  FILE *ocf = NULL;
  int *nlab, *labs;
  long i;
  global_options(argc, argv);
  cfout = opt(argc, argv, "-cfout");
  dout = need(argc, argv, "-dout");
  if (open_pair(argc, argv, 0, 1, 1, 0, &data, &codes)) return 1;  /* classify.c:117-129 */
  if (cfout && !(ocf = fopen(cfout, "w"))) { fprintf(stderr, "Cannot write to %s\n", cfout); return 1; }
  if (find_winners(codes, data, 1, &w)) return 1;
  nlab = (int *)malloc(sizeof(int) * (size_t)(data->n > 0 ? data->n : 1));
  labs = (int *)malloc(sizeof(int) * (size_t)(data->n > 0 ? data->n : 1));
  if (!nlab || !labs) return 1;
  {
    long pos = 0, keep = 0;
    /* entries without a winner keep their labels (classify.c:63-67); count them first */
    for (i = 0; i < data->n; i++) if (w.nfound[i] == 0) keep += data->lab_off[i + 1] - data->lab_off[i];
    if (keep > 0) {
      int *t = (int *)realloc(labs, sizeof(int) * (size_t)(data->n + keep));
      if (!t) return 1;
      labs = t;
    }
    for (i = 0; i < data->n; i++) {                                /* classify.c:62-80 */
      int label;
      if (w.nfound[i] == 0) {
        long l;
        label = label_index("# empty datavector");
        nlab[i] = (int)(data->lab_off[i + 1] - data->lab_off[i]);
        for (l = data->lab_off[i]; l < data->lab_off[i + 1]; l++) labs[pos++] = data->lab_pool[l];
      } else {
        label = pak_label(codes, w.idx[i]);                        /* only the first label */
        nlab[i] = label != LABEL_EMPTY ? 1 : 0;
        if (label != LABEL_EMPTY) labs[pos++] = label;
      }
      if (ocf) fprintf(ocf, "%s\n", label_string(label) ? label_string(label) : "(null)");
    }
  }
  if (pak_set_labels(data, nlab, labs)) return 1;
  if (ocf) fclose(ocf);
  pak_save(data, dout);
  free(nlab); free(labs);
  winners_free(&w);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ training */
static int alpha_type_of(int argc, char **argv, int *type) {
  const char *s = opt(argc, argv, "-alpha_type");
  *type = BMU_ALPHA_LINEAR;
  if (!s) return 0;
  if (strcasecmp(s, "linear") == 0) return 0;
  if (strcasecmp(s, "inverse_t") == 0) { *type = BMU_ALPHA_INVERSE_T; return 0; }
  fprintf(stderr, "Unknown alpha type %s\n", s);
  return 1;
}

/* the data row of every training step: list order walked cyclically; `-rand seed` shuffles the list when
 * it is read (datafile.c:1152-1188; seed 0 means "seed from the clock", lvq_pak.c:476-484); `-buffer B`
 * makes the reference hold B entries at a time and shuffle every chunk each time it is (re-)read
 * (datafile.c:237-344) -- reproduced from the row counts by bmu_sample_sequence, the file itself stays
 * resident on the device.  Returns `length` row indices (the caller passes them to the schedule helpers
 * as `order` with N = length), or NULL for plain list order. */
static int32_t *sample_sequence(int argc, char **argv, long n, long length) {
  const char *s = opt(argc, argv, "-rand"), *b = opt(argc, argv, "-buffer");
  const long buffer = b ? atol(b) : 0;
  int32_t *seq;
  int seed = 0;
  if (!s && buffer <= 0) return NULL;
  if (s) {
    seed = atoi(s);
    if (seed == 0) seed = (int)time(NULL);
  }
  seq = (int32_t *)malloc(sizeof(int32_t) * (size_t)(length > 0 ? length : 1));
  if (seq) bmu_sample_sequence(n, buffer, seed, length, seq);
  return seq;
}

/* ---- snapshots (lvq_pak.c:663-764, som_rout.c:650-658): every `interval` steps the codebook is
 * written out.  The resident trainer (bmu_trainer_*) keeps data and codebook on the device between
 * the chunks of steps, so a snapshot costs one download of the codebook. */
struct snap {
  long interval;           /* 0 = none */
  const char *pattern;     /* file name, may hold one %d / %ld for the iteration */
  int keepopen;            /* -snaptype keepopen: all snapshots in one file between #start n / #end */
  FILE *fp;
  int counter;
};
static int snap_setup(struct snap *sn, int argc, char **argv, const char *default_name) {
  const char *s = opt(argc, argv, "-snapinterval");
  memset(sn, 0, sizeof *sn);
  sn->interval = s ? atol(s) : 0;
  if (!sn->interval) return 0;
  sn->pattern = opt(argc, argv, "-snapfile");
  if (!sn->pattern) {
    sn->pattern = default_name;
    fprintf(stderr, "snapshot file not specified, using '%s'", default_name);
  }
  s = opt(argc, argv, "-snaptype");
  if (s && strcasecmp(s, "keepopen") == 0) sn->keepopen = 1;
  else if (s && strcasecmp(s, "file") != 0) fprintf(stderr, "note: snapshot type %s is written synchronously\n", s);
  return 0;
}
/* the -snapfile name may hold ONE integer conversion for the iteration (lvq_pak.c:700-703 passes it to
 * sprintf as it is); anything else a user-supplied string could smuggle into a format is refused */
static int snap_pattern_ok(const char *p) {
  int conversions = 0;
  for (; *p; p++) {
    if (*p != '%') continue;
    p++;
    if (*p == '%') continue;
    while (*p == '0' || (*p >= '1' && *p <= '9')) p++;      /* width */
    if (*p == 'l') p++;
    if (*p != 'd' && *p != 'i') return 0;
    if (++conversions > 1) return 0;
  }
  return 1;
}
static int snap_save(struct snap *sn, const struct pak_entries *codes, long iter, long length) {
  FILE *fp = sn->fp;
  long i, l;
  int c;
  sn->counter++;
  if (!fp) {
    char name[1024];
    if (snap_pattern_ok(sn->pattern)) snprintf(name, sizeof name, sn->pattern, iter);
    else snprintf(name, sizeof name, "%s", sn->pattern);
    fp = fopen(name, "w");
    if (!fp) return 1;
    if (sn->keepopen) sn->fp = fp;
  }
  if (sn->keepopen) fprintf(fp, "#start %d\n", sn->counter);
  pak_write_header(fp, codes);
  fprintf(fp, "#SNAPSHOT FILE\n#iterations: %ld/%ld\n", iter, length);
  for (i = 0; i < codes->n; i++) {
    const float *pt = codes->points + (size_t)i * codes->dim;
    const unsigned char *mk = codes->mask ? codes->mask + (size_t)i * codes->dim : NULL;
    for (c = 0; c < codes->dim; c++) {
      if (mk && mk[c]) fprintf(fp, "%s ", pak_mask_string);
      else fprintf(fp, "%g ", pt[c]);
    }
    for (l = codes->lab_off[i]; l < codes->lab_off[i + 1]; l++) fprintf(fp, "%s ", label_string(codes->lab_pool[l]));
    fprintf(fp, "\n");
  }
  if (sn->keepopen) { fprintf(fp, "#end\n"); fflush(fp); }
  else fclose(fp);
  return 0;
}
/* first step index >= from after which a snapshot is due (le % interval == 0 && le > 0), or -1 */
static long snap_next(const struct snap *sn, long from, long length) {
  long le;
  if (!sn->interval) return -1;
  le = from < 1 ? 1 : from;
  le = (le + sn->interval - 1) / sn->interval * sn->interval;
  return le < length ? le : -1;
}

int vsom_main(int argc, char **argv) {
  struct pak_entries *data = NULL, *codes = NULL;
  const char *cout_name;
  long length;
  float alpha, radius;
  int alpha_type, use_fixed, use_weights, rc;
  int32_t *order, *sample;
  float *talp, *trad;
  struct snap sn;
  global_options(argc, argv);
  cout_name = need(argc, argv, "-cout");
  length = atol(need(argc, argv, "-rlen"));
  alpha = (float)atof(need(argc, argv, "-alpha"));
  radius = (float)atof(need(argc, argv, "-radius"));
  use_fixed = flag(argc, argv, "-fixed");
  use_weights = flag(argc, argv, "-weights");
  if (alpha_type_of(argc, argv, &alpha_type)) return 1;
  if (open_pair(argc, argv, 0, 0, 1, 1, &data, &codes)) return 1;
  snap_setup(&sn, argc, argv, cout_name);
  if (length > 0 && data->n > 0) {
    order = sample_sequence(argc, argv, data->n, length);
    sample = (int32_t *)malloc(sizeof(int32_t) * (size_t)length);
    talp = (float *)malloc(sizeof(float) * (size_t)length);
    trad = (float *)malloc(sizeof(float) * (size_t)length);
    if (!sample || !talp || !trad) { fprintf(stderr, "out of memory\n"); return 1; }
    /* with a sequence the modulus is the run length: step le uses order[le] */
    bmu_som_schedule(0, length, length, alpha, radius, alpha_type, order ? length : data->n, order,
                     use_weights ? data->weight : NULL, sample, talp, trad);
    if (!sn.interval) {
      rc = bmu_som_train(codes->points, codes->n, codes->dim, codes->xdim, codes->ydim, codes->topol, codes->neigh,
                         data->points, data->mask, data->n, use_fixed ? data->fixed_xy : NULL, sample, talp, trad,
                         length);
      if (rc) return engine_failed("bmu_som_train");
    } else {
      bmu_trainer *t = bmu_trainer_create(codes->points, codes->n, codes->dim, data->points, data->mask, data->n);
      long le0 = 0;
      if (!t || bmu_trainer_set_som(t, codes->xdim, codes->ydim, codes->topol, codes->neigh,
                                    use_fixed ? data->fixed_xy : NULL))
        return engine_failed("bmu_trainer");
      while (le0 < length) {
        const long due = snap_next(&sn, le0, length), le1 = due >= 0 ? due + 1 : length;
        if (bmu_trainer_steps(t, sample + le0, talp + le0, trad + le0, le1 - le0)) return engine_failed("bmu_trainer_steps");
        if (due >= 0) {
          if (bmu_trainer_get_codes(t, codes->points)) return engine_failed("bmu_trainer_get_codes");
          if (snap_save(&sn, codes, due, length)) fprintf(stderr, "snapshot failed, continuing teaching\n");
        }
        le0 = le1;
      }
      if (bmu_trainer_get_codes(t, codes->points)) return engine_failed("bmu_trainer_get_codes");
      bmu_trainer_destroy(t);
    }
    free(order); free(sample); free(talp); free(trad);
  }
  pak_save(codes, cout_name);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* datafile.c:1030-1046: strtok(basename, ".") -- leading dots are skipped, the name ends at the next one */
static void lra_name(const char *codefile, char *out, size_t outsz) {
  size_t i = 0;
  while (codefile[i] == '.' && i + 5 < outsz) { out[i] = codefile[i]; i++; }
  while (codefile[i] && codefile[i] != '.' && i + 5 < outsz) { out[i] = codefile[i]; i++; }
  strcpy(out + i, ".lra");
}

int lvqtrain_main(int argc, char **argv, const char *progname) {
  struct pak_entries *data = NULL, *codes = NULL;
  const char *cin_name, *cout_name, *s;
  long length, i;
  float alpha = 0.0f, winlen = 0.0f, epsilon = 0.0f, win_thr = 0.0f, *talp, *unit_alpha = NULL;
  int algo, alpha_type, rc;
  int32_t *order, *sample, *code_label, *data_label;
  struct snap sn;
  char lra[2048];
  global_options(argc, argv);
  s = opt(argc, argv, "-type");
  if (s) progname = s;
  if (strcasecmp(progname, "lvq1") == 0) algo = BMU_LVQ1;
  else if (strcasecmp(progname, "lvq2") == 0) algo = BMU_LVQ2;
  else if (strcasecmp(progname, "lvq3") == 0) algo = BMU_LVQ3;
  else if (strcasecmp(progname, "olvq1") == 0) algo = BMU_OLVQ1;
  else { fprintf(stderr, "Unknown LVQ type %s\n", progname); return 1; }
  cin_name = need(argc, argv, "-cin");
  cout_name = need(argc, argv, "-cout");
  length = atol(need(argc, argv, "-rlen"));
  if (algo == BMU_OLVQ1) { s = opt(argc, argv, "-alpha"); alpha = s ? (float)atof(s) : 0.0f; }   /* lvqtrain.c:151-166 */
  else alpha = (float)atof(need(argc, argv, "-alpha"));
  if (algo == BMU_LVQ2 || algo == BMU_LVQ3) winlen = (float)atof(need(argc, argv, "-win"));
  if (algo == BMU_LVQ3) epsilon = (float)atof(need(argc, argv, "-epsilon"));
  if (alpha_type_of(argc, argv, &alpha_type)) return 1;
  if (open_pair(argc, argv, 1, 1, 1, 0, &data, &codes)) return 1;

  code_label = (int32_t *)malloc(sizeof(int32_t) * (size_t)(codes->n > 0 ? codes->n : 1));
  data_label = (int32_t *)malloc(sizeof(int32_t) * (size_t)(data->n > 0 ? data->n : 1));
  if (!code_label || !data_label) return 1;
  for (i = 0; i < codes->n; i++) code_label[i] = pak_label(codes, i);
  for (i = 0; i < data->n; i++) data_label[i] = pak_label(data, i);

  if (algo == BMU_OLVQ1) {                                          /* lvq_rout.c:609-627 */
    unit_alpha = (float *)malloc(sizeof(float) * (size_t)codes->n);
    if (!unit_alpha) return 1;
    rc = 0;
    if (alpha == 0.0f) {
      FILE *fp;
      lra_name(cin_name, lra, sizeof lra);
      fp = fopen(lra, "r");
      if (fp) {
        rc = 1;
        for (i = 0; i < codes->n; i++)
          if (fscanf(fp, "%g\n", &unit_alpha[i]) < 0) { rc = 0; break; }
        fclose(fp);
      } else if (verbose_level >= 1) {
        fprintf(stderr, "Can't open alpha file %s", lra);
      }
      if (!rc) alpha = 0.3f;
    }
    if (!rc)
      for (i = 0; i < codes->n; i++) unit_alpha[i] = alpha;
  }
  {
    const float w = winlen;
    win_thr = (1 - w) / (1 + w);                                     /* lvq_rout.c:770, float arithmetic */
  }
  if (length > 0 && data->n > 0) {
    order = sample_sequence(argc, argv, data->n, length);
    sample = (int32_t *)malloc(sizeof(int32_t) * (size_t)length);
    talp = (float *)malloc(sizeof(float) * (size_t)length);
    if (!sample || !talp) { fprintf(stderr, "out of memory\n"); return 1; }
    bmu_lvq_schedule(0, length, length, alpha, alpha_type, order ? length : data->n, order, sample, talp);
    snap_setup(&sn, argc, argv, cout_name);
    if (!sn.interval) {
      rc = bmu_lvq_train(algo, codes->points, code_label, codes->n, codes->dim, data->points, data->mask, data_label,
                         data->n, sample, talp, length, win_thr, epsilon, alpha, unit_alpha);
      if (rc) return engine_failed("bmu_lvq_train");
    } else {                                                         /* lvq_rout.c:560-568 and alike */
      bmu_trainer *t = bmu_trainer_create(codes->points, codes->n, codes->dim, data->points, data->mask, data->n);
      long le0 = 0;
      if (!t || bmu_trainer_set_lvq(t, algo, code_label, data_label, win_thr, epsilon, alpha, unit_alpha))
        return engine_failed("bmu_trainer");
      while (le0 < length) {
        const long due = snap_next(&sn, le0, length), le1 = due >= 0 ? due + 1 : length;
        if (bmu_trainer_steps(t, sample + le0, talp + le0, NULL, le1 - le0)) return engine_failed("bmu_trainer_steps");
        if (due >= 0) {
          if (bmu_trainer_get_codes(t, codes->points)) return engine_failed("bmu_trainer_get_codes");
          if (snap_save(&sn, codes, due, length)) fprintf(stderr, "snapshot failed\n");
        }
        le0 = le1;
      }
      if (bmu_trainer_get_codes(t, codes->points)) return engine_failed("bmu_trainer_get_codes");
      if (algo == BMU_OLVQ1 && bmu_trainer_get_unit_alpha(t, unit_alpha)) return engine_failed("bmu_trainer_get_unit_alpha");
      bmu_trainer_destroy(t);
    }
    free(order); free(sample); free(talp);
  }
  if (algo == BMU_OLVQ1) {                                          /* lvq_rout.c:694, datafile.c:1061-1086 */
    FILE *fp;
    lra_name(cout_name, lra, sizeof lra);
    fp = fopen(lra, "w+");
    if (fp) {
      for (i = 0; i < codes->n; i++) fprintf(fp, "%g\n", unit_alpha[i]);
      fclose(fp);
    } else {
      fprintf(stderr, "Can't open alpha file %s for writing", lra);
    }
  }
  pak_save(codes, cout_name);
  /* lvqtrain.c:248-249 invalidate_alphafile(): every lvq program, olvq1 included, removes the
   * .lra that belongs to the output name again */
  lra_name(cout_name, lra, sizeof lra);
  {
    FILE *fp = fopen(lra, "r");
    if (fp) {                                                       /* datafile.c:1088-1108 */
      if (verbose_level >= 1) fprintf(stdout, "Removing the learning rate file %s\n", lra);
      fclose(fp);
      if (remove(lra)) fprintf(stderr, "Can not remove %s", lra);
    }
  }
  free(code_label); free(data_label); free(unit_alpha);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ vfind */
/* vfind.c:138-330: answers are read from stdin after the same prompts; every trial is
 * randinit (seed = trial number) -> two training phases -> quantization error on the test file;
 * the best map is saved.  Trials are independent: with BMU_TRIAL_STRIDE / BMU_TRIAL_OFFSET (set
 * by a launcher, one process per GPU) a process runs every stride-th trial and reports its best
 * on stdout, so that the launcher can pick the overall best (som_lvq_pak_b200/distributed.py). */
static long ask_int(const char *q, long def) {
  char str[100];
  printf("%s: ", q);
  if (!fgets(str, sizeof str, stdin)) return def;
  return atol(str);
}
static float ask_float(const char *q, float def) {
  char str[100];
  printf("%s: ", q);
  if (!fgets(str, sizeof str, stdin)) return def;
  return (float)atof(str);
}
static char *ask_str(const char *q) {
  char str[100], *t;
  printf("%s: ", q);
  if (!fgets(str, sizeof str, stdin)) { printf("Can't read required data\n"); exit(1); }
  t = strdup(str);
  if (strchr(t, ' ')) *strchr(t, ' ') = '\0';
  if (strchr(t, '\n')) *strchr(t, '\n') = '\0';
  return t;
}
static int name_to_id(const char *s, const char *a, int ia, const char *b, int ib) {
  if (strcasecmp(s, a) == 0) return ia;
  if (strcasecmp(s, b) == 0) return ib;
  return 0;
}

int vfind_main(int argc, char **argv) {
  struct pak_entries *data, *test, *best = NULL;
  long trials, bnot = 0, length1, length2, lmax, not;
  int xdim, ydim, topol, neigh, alpha_type, use_fixed, use_weights, qmode, stride = 1, offset = 0;
  float alpha1, radius1, alpha2, radius2, qerrorb = FLT_MAX;
  char *din, *tin, *cout_name;
  const char *s;
  int32_t *sample;
  float *talp, *trad, *per = NULL;
  global_options(argc, argv);
  printf("Repeated initialization, training and testing of a Self-Organizing Map (vfind); the best map\n"
         "(smallest quantization error on the test file) is saved.\n\n");
  trials = ask_int("Give the number of trials", 0);
  din = ask_str("Give the input data file name");
  tin = ask_str("Give the input test file name");
  cout_name = ask_str("Give the output map file name");
  topol = name_to_id(ask_str("Give the topology type"), "hexa", TOPOL_HEXA, "rect", TOPOL_RECT);
  if (!topol) topol = TOPOL_HEXA;
  neigh = name_to_id(ask_str("Give the neighborhood type"), "bubble", NEIGH_BUBBLE, "gaussian", NEIGH_GAUSSIAN);
  if (!neigh) neigh = NEIGH_BUBBLE;
  xdim = (int)ask_int("Give the x-dimension", 0);
  ydim = (int)ask_int("Give the y-dimension", 0);
  length1 = ask_int("Give the training length of first part", 0);
  alpha1 = ask_float("Give the training rate of first part", 0.0f);
  radius1 = ask_float("Give the radius in first part", 0.0f);
  length2 = ask_int("Give the training length of second part", 0);
  alpha2 = ask_float("Give the training rate of second part", 0.0f);
  radius2 = ask_float("Give the radius in second part", 0.0f);
  printf("\n");
  s = opt(argc, argv, "-fixed");   use_fixed = s ? atoi(s) : 0;        /* vfind.c:186-187: valued options */
  s = opt(argc, argv, "-weights"); use_weights = s ? atoi(s) : 0;
  s = opt(argc, argv, "-qetype");  qmode = s ? atoi(s) : 0;
  if (alpha_type_of(argc, argv, &alpha_type)) return 1;
  if ((s = getenv("BMU_TRIAL_STRIDE")) != NULL && atoi(s) > 0) stride = atoi(s);
  if ((s = getenv("BMU_TRIAL_OFFSET")) != NULL) offset = atoi(s);
  data = pak_load(din, 0, 1);
  if (!data) { fprintf(stderr, "Can't open data file '%s'\n", din); return 1; }
  test = pak_load(tin, 0, 1);
  if (!test) { fprintf(stderr, "Can't open test data file '%s'\n", tin); return 1; }
  if ((long)xdim * ydim <= 0 || xdim < 0) { fprintf(stderr, "Dimensions of map (%d %d) are incorrect\n", xdim, ydim); return 1; }
  if (bmu_init(0)) return engine_failed("bmu_init");
  lmax = length1 > length2 ? length1 : length2;
  sample = (int32_t *)malloc(sizeof(int32_t) * (size_t)(lmax > 0 ? lmax : 1));
  talp = (float *)malloc(sizeof(float) * (size_t)(lmax > 0 ? lmax : 1));
  trad = (float *)malloc(sizeof(float) * (size_t)(lmax > 0 ? lmax : 1));
  if (qmode > 0) per = (float *)malloc(sizeof(float) * (size_t)(test->n > 0 ? test->n : 1));
  if (!sample || !talp || !trad || (qmode > 0 && !per)) return 1;
  for (not = trials; not > 0; not--) {                                  /* vfind.c:247-306 */
    struct pak_entries *codes;
    float qerror = 0.0f;
    int phase;
    if ((int)((trials - not) % stride) != offset) continue;
    codes = pak_alloc(data->dim, (long)xdim * ydim);
    if (!codes) return 1;
    codes->topol = topol; codes->neigh = neigh; codes->xdim = xdim; codes->ydim = ydim;
    bmu_randinit_codes(data->points, data->mask, data->n, data->dim, codes->n, (int)not, codes->points);
    for (phase = 0; phase < 2; phase++) {
      const long len = phase ? length2 : length1;
      if (len <= 0 || data->n == 0) continue;
      bmu_som_schedule(0, len, len, phase ? alpha2 : alpha1, phase ? radius2 : radius1, alpha_type, data->n, NULL,
                       use_weights ? data->weight : NULL, sample, talp, trad);
      if (bmu_som_train(codes->points, codes->n, codes->dim, xdim, ydim, topol, neigh, data->points, data->mask,
                        data->n, use_fixed ? data->fixed_xy : NULL, sample, talp, trad, len))
        return engine_failed("bmu_som_train");
    }
    {
      bmu_codebook *cb = bmu_codebook_create(codes->points, codes->n, codes->dim);
      if (!cb) return engine_failed("bmu_codebook_create");
      if (qmode > 0) {
        long i;
        if (bmu_qerror2(cb, xdim, ydim, topol, neigh, radius2, test->points, test->mask, test->n, per))
          return engine_failed("bmu_qerror2");
        for (i = 0; i < test->n; i++) qerror += per[i];
      } else {
        struct winners w;
        bmu_codebook_destroy(cb);
        cb = NULL;
        if (find_winners(codes, test, 1, &w)) return 1;
        qerror = bmu_replay_qerror(w.diff, w.nfound, test->n, 1);
        winners_free(&w);
      }
      if (cb) bmu_codebook_destroy(cb);
    }
    if (qerror < qerrorb) {
      qerrorb = qerror;
      bnot = not;
      pak_free(best);
      best = codes;
    } else {
      pak_free(codes);
    }
    if (verbose_level >= 1) fprintf(stderr, "%3ld: %f\n", not, qerror / (float)test->n);
  }
  if (best) {
    pak_save(best, cout_name);
    if (verbose_level >= 1)
      fprintf(stdout, "Smallest error with random seed %3ld: %f\n", bnot, qerrorb / (float)test->n);
  }
  pak_free(best); pak_free(data); pak_free(test);
  free(sample); free(talp); free(trad); free(per);
  return 0;
}

/* ------------------------------------------------------------------ randinit */
/* mapinit.c:52-181 with the random initialisation (randinit / mapinit -init rand); the linear
 * initialisation (eigenvectors of the data) is not on the BMU path and is not carried over */
/* ------------------------------------------------------------------ lininit */
/* lininit_codes / find_eigenvectors (som_rout.c:166-429): the map is laid out on the plane of the two
 * largest eigenvectors of the data's autocorrelation matrix, found by ten power iterations from a start
 * drawn with the reference's generator.  Host arithmetic (every sum is a sequential float sum over the
 * data, so its order is part of the result); each operation keeps the reference's operand types. */
static void lin_normalize(float *v, int n) {                         /* som_rout.c:166-174 */
  float sum = 0.0;
  int j;
  for (j = 0; j < n; j++) sum += v[j] * v[j];
  sum = sqrt(sum);
  for (j = 0; j < n; j++) v[j] /= sum;
}
static float lin_dotprod(const float *v, const float *w, int n) {    /* som_rout.c:177-184 */
  float sum = 0.0;
  int j;
  for (j = 0; j < n; j++) sum += v[j] * w[j];
  return sum;
}
static int lin_gram_schmidt(float *v, int n, int e) {                /* som_rout.c:187-208 */
  int i, j, p, t;
  float sum, *w = (float *)malloc(sizeof(float) * (size_t)n * e);
  if (!w) return 1;
  for (i = 0; i < e; i++) {
    for (t = 0; t < n; t++) {
      sum = v[i * n + t];
      for (j = 0; j < i; j++)
        for (p = 0; p < n; p++) sum -= w[j * n + t] * w[j * n + p] * v[i * n + p];
      w[i * n + t] = sum;
    }
    lin_normalize(w + i * n, n);
  }
  memcpy(v, w, sizeof(float) * (size_t)n * e);
  free(w);
  return 0;
}
/* mean[n], eigen1[n], eigen2[n]; returns non-zero when the reference would fail (fewer than three
 * entries, a zero eigenvalue estimate) */
static int lin_eigenvectors(const struct pak_entries *data, unsigned long *next, float *mean, float *eigen1, float *eigen2) {
  const int n = data->dim;
  float *r = (float *)calloc((size_t)n * n, sizeof(float)), *m = mean;
  float *u = (float *)malloc(sizeof(float) * 2 * (size_t)n), *v = (float *)malloc(sizeof(float) * 2 * (size_t)n);
  long *k2 = (long *)calloc((size_t)n, sizeof(long)), i, j, k, e;
  float mu[2], sum;
  if (!r || !u || !v || !k2) return 1;
  for (i = 0; i < n; i++) m[i] = 0.0;
  for (e = 0; e < data->n; e++) {                                     /* som_rout.c:244-254 */
    const float *pt = data->points + (size_t)e * n;
    const unsigned char *mk = data->mask ? data->mask + (size_t)e * n : NULL;
    for (i = 0; i < n; i++)
      if (!mk || mk[i] == 0) { m[i] += pt[i]; k2[i]++; }
  }
  k = data->n;
  if (k < 3) return 1;
  for (i = 0; i < n; i++) m[i] /= k2[i];
  for (e = 0; e < data->n; e++) {                                     /* som_rout.c:269-283 */
    const float *pt = data->points + (size_t)e * n;
    const unsigned char *mk = data->mask ? data->mask + (size_t)e * n : NULL;
    for (i = 0; i < n; i++) {
      if (mk && mk[i] != 0) continue;
      for (j = i; j < n; j++) {
        if (mk && mk[j] != 0) continue;
        r[i * n + j] += (pt[i] - m[i]) * (pt[j] - m[j]);
      }
    }
  }
  for (i = 0; i < n; i++)
    for (j = i; j < n; j++) r[j * n + i] = r[i * n + j] /= k;
  for (i = 0; i < 2; i++) {                                           /* som_rout.c:289-293 */
    for (j = 0; j < n; j++) {
      *next = (*next * 23) % 100000001;                               /* orand, lvq_pak.c:470-473 */
      u[i * n + j] = (long)(int)(*next % 32767L) / 16384.0 - 1.0;
    }
    lin_normalize(u + i * n, n);
    mu[i] = 1.0;
  }
  for (k = 0; k < 10; k++) {                                          /* som_rout.c:295-311 */
    for (i = 0; i < 2; i++)
      for (j = 0; j < n; j++) v[i * n + j] = mu[i] * lin_dotprod(r + j * n, u + i * n, n) + u[i * n + j];
    if (lin_gram_schmidt(v, n, 2)) return 1;
    sum = 0.0;                                                        /* NOT reset between the two vectors */
    for (i = 0; i < 2; i++) {
      for (j = 0; j < n; j++) sum += fabs(v[i * n + j] / lin_dotprod(r + j * n, v + i * n, n));
      mu[i] = sum / n;
    }
    memcpy(u, v, sizeof(float) * 2 * (size_t)n);
  }
  if (mu[0] == 0.0 || mu[1] == 0.0) return 1;
  for (j = 0; j < n; j++) { eigen1[j] = u[j]; eigen1[j] /= sqrt(mu[0]); }
  for (j = 0; j < n; j++) { eigen2[j] = u[n + j]; eigen2[j] /= sqrt(mu[1]); }
  free(r); free(u); free(v); free(k2);
  return 0;
}
static int lininit_codes(const struct pak_entries *data, int xdim, int ydim, long seed, float *codes) {
  const int dim = data->dim;
  unsigned long next = (unsigned long)(int)seed;
  float *mean = (float *)malloc(sizeof(float) * 3 * (size_t)dim), *eigen1, *eigen2, xf, yf;
  long index, i;
  if (!mean) return 1;
  eigen1 = mean + dim;
  eigen2 = eigen1 + dim;
  if (lin_eigenvectors(data, &next, mean, eigen1, eigen2)) {
    fprintf(stderr, "lininit_codes: Can't find eigenvectors\n");
    free(mean);
    return 1;
  }
  for (index = 0; index < (long)xdim * ydim; index++) {               /* som_rout.c:412-419 */
    xf = 4.0 * (float)(index % xdim) / (xdim - 1.0) - 2.0;
    yf = 4.0 * (float)(index / xdim) / (ydim - 1.0) - 2.0;
    for (i = 0; i < dim; i++) codes[(size_t)index * dim + i] = mean[i] + xf * eigen1[i] + yf * eigen2[i];
  }
  free(mean);
  return 0;
}

int randinit_main(int argc, char **argv, const char *progname) {
  struct pak_entries *data, *codes;
  const char *din, *cout_name, *s;
  int topol, neigh, xdim, ydim;
  long seed;
  FILE *fp;
  long i;
  int c, lin = -1;
  global_options(argc, argv);
  din = need(argc, argv, "-din");
  cout_name = need(argc, argv, "-cout");
  s = opt(argc, argv, "-rand");
  seed = s ? atol(s) : 0;
  s = need(argc, argv, "-topol");
  topol = name_to_id(s, "hexa", TOPOL_HEXA, "rect", TOPOL_RECT);
  if (!topol) { fprintf(stderr, "Unknown topology type %s\n", s); return 1; }
  s = need(argc, argv, "-neigh");
  neigh = name_to_id(s, "bubble", NEIGH_BUBBLE, "gaussian", NEIGH_GAUSSIAN);
  if (!neigh) { fprintf(stderr, "Unknown neighborhood type %s\n", s); return 1; }
  xdim = atoi(need(argc, argv, "-xdim"));
  ydim = atoi(need(argc, argv, "-ydim"));
  s = opt(argc, argv, "-init");
  if (strcasecmp(progname, "lininit") == 0) lin = 1;                  /* mapinit.c:72-75, 104-119 */
  else if (strcasecmp(progname, "randinit") == 0) lin = 0;
  if (s) lin = strcmp(s, "lin") == 0 ? 1 : (strcmp(s, "rand") == 0 ? 0 : -1);
  if (lin < 0) { fprintf(stderr, "Unknown initialization type %s\n", s ? s : progname); return 1; }
  if ((long)xdim * ydim <= 0 || xdim < 0) { fprintf(stderr, "Dimensions of map (%d %d) are incorrect\n", xdim, ydim); return 1; }
  data = pak_load(din, 0, 1);
  if (!data) { fprintf(stderr, "Can't open data file '%s'\n", din); return 1; }
  codes = pak_alloc(data->dim, (long)xdim * ydim);
  if (!codes) return 1;
  codes->topol = topol; codes->neigh = neigh; codes->xdim = xdim; codes->ydim = ydim;
  if (lin) {
    if (lininit_codes(data, xdim, ydim, seed ? seed : (long)time(NULL), codes->points)) {
      fprintf(stderr, "initialization failure\n");
      return 1;
    }
  } else {
    bmu_randinit_codes(data->points, data->mask, data->n, data->dim, codes->n,
                       (int)(seed ? seed : (long)time(NULL)), codes->points);     /* init_random, lvq_pak.c:478-484 */
  }
  fp = fopen(cout_name, "w");
  if (!fp) { fprintf(stderr, "save_entries: Can't open file '%s'\n", cout_name); return 1; }
  pak_write_header(fp, codes);
  fprintf(fp, "# random seed: %ld\n", seed);                                      /* mapinit.c:176-177 */
  for (i = 0; i < codes->n; i++) {
    for (c = 0; c < codes->dim; c++) fprintf(fp, "%g ", codes->points[(size_t)i * codes->dim + c]);
    fprintf(fp, "\n");
  }
  fclose(fp);
  pak_free(data);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ eveninit / propinit */
/* eveninit.c:46-160 + pick_inside_codes / correct_by_knn (lvq_rout.c:38-211): codebook vectors are
 * picked from the data, class by class, among the entries that their own knn nearest neighbours in
 * the data set classify correctly.  The reference runs one k-NN search per visited entry; here the
 * whole self-search is ONE bmu_search call and the picking loop replays it in data order. */
static void pick_inside(const struct pak_entries *data, const unsigned char *inside, struct pak_hitlist *classes,
                        long *picked, long *npicked) {
  long total = 0, i, c;
  for (c = 0; c < classes->n; c++) total += classes->freq[c];
  for (i = 0; i < data->n && total > 0; i++) {
    const long lab = pak_label(data, i);
    for (c = 0; c < classes->n; c++)
      if (classes->label[c] == lab) break;
    if (c == classes->n || classes->freq[c] <= 0 || !inside[i]) continue;
    total--;
    picked[(*npicked)++] = i;
    classes->freq[c]--;
  }
}

int eveninit_main(int argc, char **argv, const char *progname) {
  struct pak_entries *data, *out;
  struct winners w;
  struct pak_hitlist classes, hits;
  const char *din, *cout_name, *s;
  unsigned char *inside;
  long noc, nol, tot, nic, emp = 0, i, c, npicked = 0, *picked, nl = 0;
  int knn, prop = -1, t;
  char lra[2048];
  global_options(argc, argv);
  if (strcasecmp(progname, "propinit") == 0) prop = 1;
  else if (strcasecmp(progname, "eveninit") == 0) prop = 0;
  s = opt(argc, argv, "-type");
  if (s && strcasecmp(s, "propinit") == 0) prop = 1;
  else if (s && strcasecmp(s, "eveninit") == 0) prop = 0;
  if (prop < 0) { fprintf(stderr, "unknown init type\n"); return 1; }
  din = need(argc, argv, "-din");
  cout_name = need(argc, argv, "-cout");
  noc = atol(need(argc, argv, "-noc"));
  s = opt(argc, argv, "-knn");
  knn = s ? atoi(s) : 5;
  if (knn < 1) knn = 1;
  if (knn > BMU_KMAX) { fprintf(stderr, "-knn %d is larger than the engine's limit %d\n", knn, BMU_KMAX); return 1; }
  data = pak_load(din, 1, 1);
  if (!data) { fprintf(stderr, "Can't open data file '%s'\n", din); return 1; }
  if (bmu_init(0)) return engine_failed("bmu_init");
  if (find_winners(data, data, knn, &w)) return 1;
  /* correct_by_knn for every entry: majority label of its knn neighbours == its own label */
  inside = (unsigned char *)calloc((size_t)(data->n > 0 ? data->n : 1), 1);
  picked = (long *)malloc(sizeof(long) * (size_t)(data->n > 0 ? data->n : 1));
  if (!inside || !picked) return 1;
  hit_init(&hits);
  for (i = 0; i < data->n; i++) {
    if (w.nfound[i] != knn) { inside[i] = 1; continue; }          /* -1 from correct_by_knn counts as true (lvq_rout.c:173) */
    hit_clear(&hits);
    for (t = 0; t < knn; t++) hit_add(&hits, pak_label(data, w.idx[i * knn + t]));
    inside[i] = hits.n > 0 && hits.label[0] == pak_label(data, i);
  }
  hit_init(&classes);
  for (i = 0; i < data->n; i++) hit_add(&classes, pak_label(data, i));
  nol = classes.n;
  tot = data->n;
  if (nol > noc) fprintf(stderr, "There are more different classes than requested codes");
  nic = nol ? noc / nol : 0;
  for (c = 0; c < nol; c++) {                                      /* eveninit.c:86-94 */
    if (prop) {
      classes.freq[c] = (long)(classes.freq[c] * (float)noc / tot);
      if (classes.freq[c] < 1) classes.freq[c] = 1;
    } else {
      classes.freq[c] = nic;
    }
  }
  pick_inside(data, inside, &classes, picked, &npicked);
  for (c = 0; c < nol; c++) if (classes.freq[c] == 0) emp++;
  if (npicked < noc) {                                             /* eveninit.c:116-143: second pass */
    float frac = 0.0f, err = 0.0f;
    if (emp != 0) frac = (noc - npicked) / (float)emp;
    for (c = 0; c < nol; c++) {
      if (classes.freq[c] == 0) {
        classes.freq[c] = (int)(frac + err);
        err = frac + err - classes.freq[c];
      } else {
        classes.freq[c] = 0;
      }
    }
    pick_inside(data, inside, &classes, picked, &npicked);
  }
  out = pak_alloc(data->dim, npicked);
  if (!out) return 1;
  out->topol = TOPOL_LVQ;
  out->neigh = data->neigh; out->xdim = data->xdim; out->ydim = data->ydim;
  if (data->mask) out->mask = (unsigned char *)calloc((size_t)(npicked > 0 ? npicked : 1) * data->dim, 1);
  out->lab_pool = (int *)malloc(sizeof(int) * (size_t)(data->lab_off[data->n] > 0 ? data->lab_off[data->n] : 1));
  if (!out->lab_pool || (data->mask && !out->mask)) return 1;
  for (i = 0; i < npicked; i++) {
    const long r = picked[i];
    long l;
    memcpy(out->points + (size_t)i * data->dim, data->points + (size_t)r * data->dim, sizeof(float) * data->dim);
    if (data->mask) memcpy(out->mask + (size_t)i * data->dim, data->mask + (size_t)r * data->dim, (size_t)data->dim);
    for (l = data->lab_off[r]; l < data->lab_off[r + 1]; l++) out->lab_pool[nl++] = data->lab_pool[l];
    out->lab_off[i + 1] = nl;
  }
  pak_save(out, cout_name);
  lra_name(cout_name, lra, sizeof lra);                              /* eveninit.c:229 invalidate_alphafile */
  {
    FILE *fp = fopen(lra, "r");
    if (fp) {
      if (verbose_level >= 1) fprintf(stdout, "Removing the learning rate file %s\n", lra);
      fclose(fp);
      if (remove(lra)) fprintf(stderr, "Can not remove %s", lra);
    }
  }
  hit_free(&hits); hit_free(&classes);
  free(inside); free(picked);
  winners_free(&w);
  pak_free(data);
  pak_free(out);
  return 0;
}

/* ------------------------------------------------------------------ mindist */
/* med_distances (lvq_rout.c:383-492): per class, the median over its entries of the distance to
 * the nearest LATER entry of the same class (vector_dist_euc, lvq_pak.c:291-316), as `mindist`
 * prints it (mindist.c:93-105).  The O(M^2 D) pair loops run on the device (bmu_class_nearest);
 * class order, medians and printing stay here. */
static int cmp_float(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

int mindist_main(int argc, char **argv) {
  struct pak_entries *codes;
  struct pak_hitlist classes;
  const char *cin_name;
  float *meds, *near;
  int32_t *label, *found;
  long c, i;
  global_options(argc, argv);
  cin_name = need(argc, argv, "-cin");
  if (opt(argc, argv, "-din")) fprintf(stderr, "note: standard deviations (-din) are not provided by the B200 host\n");
  codes = pak_load(cin_name, 1, 1);
  if (!codes) { fprintf(stderr, "Can't read code file '%s'\n", cin_name); return 1; }
  if (bmu_init(0)) return engine_failed("bmu_init");
  hit_init(&classes);
  for (i = 0; i < codes->n; i++) hit_add(&classes, pak_label(codes, i));
  {
    const size_t n = (size_t)(codes->n > 0 ? codes->n : 1);
    meds = (float *)malloc(sizeof(float) * n);
    near = (float *)malloc(sizeof(float) * n);
    label = (int32_t *)malloc(sizeof(int32_t) * n);
    found = (int32_t *)malloc(sizeof(int32_t) * n);
  }
  if (!meds || !near || !label || !found) return 1;
  for (i = 0; i < codes->n; i++) label[i] = pak_label(codes, i);
  if (bmu_class_nearest(codes->points, codes->mask, label, codes->n, codes->dim, near, found))
    return engine_failed("bmu_class_nearest");
  for (c = 0; c < classes.n; c++) {
    const long lab = classes.label[c];
    long not = 0;
    float dist = 0.0f;
    for (i = 0; i < codes->n; i++)
      if (label[i] == lab && found[i]) meds[not++] = near[i];
    if (not > 0) {
      qsort(meds, (size_t)not, sizeof(float), cmp_float);
      dist = meds[not / 2];
    }
    fprintf(stdout, "In class %9s %3d units, min dist.: %6.3f\n", label_string((int)lab), (int)classes.freq[c], dist);
  }
  free(meds); free(near); free(label); free(found);
  hit_free(&classes);
  pak_free(codes);
  return 0;
}

/* ------------------------------------------------------------------ balance */
/* balance.c:45-226: med_distances per class (K5), classes whose median is far below / above the mean
 * lose / gain one codebook vector, new vectors are picked among the data entries their own knn
 * neighbours classify correctly (pick_inside_codes, one batched self-search), then one pass of OLVQ1
 * (length = number of data entries, alpha 0.3) and the medians again.
 *
 * Reference quirk, reproduced on purpose: balance.c:160-190 appends the new vectors without counting
 * them (its comment there says so, in Finnish), so olvq1_training (lvq_rout.c:607-627) sizes its per-unit rate
 * array by the count BEFORE the additions and the appended units read their rate from beyond the array.
 * With glibc's allocator those bytes are zero / a denormal chunk header, i.e. the appended units are
 * never adapted (the reference's ex1b.cod holds them as untouched data vectors) and the .lra file has
 * one line per counted unit only.  Here: rate 0 for the appended units, .lra of the counted length. */
static int class_medians(const struct pak_entries *codes, const int32_t *label, struct pak_hitlist *classes,
                         float *dists) {
  const size_t n = (size_t)(codes->n > 0 ? codes->n : 1);
  float *near = (float *)malloc(sizeof(float) * n), *meds = (float *)malloc(sizeof(float) * n);
  int32_t *found = (int32_t *)malloc(sizeof(int32_t) * n);
  long c, i;
  if (!near || !meds || !found) return 1;
  hit_clear(classes);
  for (i = 0; i < codes->n; i++) hit_add(classes, label[i]);
  if (bmu_class_nearest(codes->points, codes->mask, label, codes->n, codes->dim, near, found))
    return engine_failed("bmu_class_nearest");
  for (c = 0; c < classes->n; c++) {                                   /* lvq_rout.c:430-470 */
    long not = 0;
    dists[c] = 0.0f;
    for (i = 0; i < codes->n; i++)
      if (label[i] == classes->label[c] && found[i]) meds[not++] = near[i];
    if (not > 0) {
      qsort(meds, (size_t)not, sizeof(float), cmp_float);
      dists[c] = meds[not / 2];
    }
  }
  free(near); free(meds); free(found);
  return 0;
}

int balance_main(int argc, char **argv) {
  const double BAL = 1.3;                                              /* balance.c:30 */
  struct pak_entries *data, *codes, *out;
  struct pak_hitlist classes, more, hits;
  struct winners w;
  const char *din, *cin_name, *cout_name, *s;
  int32_t *label, *data_label, *out_label, *sample;
  unsigned char *inside;
  float *dists, *unit_alpha, aver;
  long *picked, npicked = 0, i, c, nol, kept = 0, counted, nl = 0, total;
  int *diff, note, knn, t;
  char lra[2048];
  global_options(argc, argv);
  din = need(argc, argv, "-din");
  cin_name = need(argc, argv, "-cin");
  cout_name = need(argc, argv, "-cout");
  s = opt(argc, argv, "-knn");
  knn = s ? atoi(s) : 5;
  if (knn < 1) knn = 1;
  if (knn > BMU_KMAX) { fprintf(stderr, "-knn %d is larger than the engine's limit %d\n", knn, BMU_KMAX); return 1; }
  if (verbose_level >= 2) fprintf(stderr, "Input entries are read from file %s\n", din);
  data = pak_load(din, 1, 1);
  if (!data) { fprintf(stderr, "Can't open data file '%s'\n", din); return 1; }
  if (verbose_level >= 2) fprintf(stderr, "Codebook entries are read from file %s\n", cin_name);
  codes = pak_load(cin_name, 1, 1);
  if (!codes) { fprintf(stderr, "Can't open code file '%s'\n", cin_name); return 1; }
  if (data->dim != codes->dim) { fprintf(stderr, "Data and codes have different dimensions\n"); return 1; }
  if (bmu_init(0)) return engine_failed("bmu_init");

  label = (int32_t *)malloc(sizeof(int32_t) * (size_t)(codes->n > 0 ? codes->n : 1));
  dists = (float *)malloc(sizeof(float) * (size_t)(codes->n + data->n + 1));
  if (!label || !dists) return 1;
  for (i = 0; i < codes->n; i++) label[i] = pak_label(codes, i);
  hit_init(&classes);
  if (verbose_level >= 2) fprintf(stderr, "Medians of the shortest distances are computed\n");
  if (class_medians(codes, label, &classes, dists)) return 1;
  nol = classes.n;
  diff = (int *)calloc((size_t)(nol > 0 ? nol : 1), sizeof(int));
  if (!diff) return 1;
  aver = 0.0f;
  note = 0;
  for (c = 0; c < nol; c++)
    if (classes.freq[c] > 1) { aver += dists[c]; note++; }
  aver /= note;                                                        /* balance.c:82-90 */
  note = 0;
  if (verbose_level >= 2) fprintf(stderr, "Medians of different classes are compared\n");
  for (c = 0; c < nol; c++) {                                          /* balance.c:95-104 */
    if ((aver > BAL * dists[c]) && (classes.freq[c] > 1)) { diff[c]--; note++; }
    if (BAL * aver < dists[c]) { diff[c]++; note--; }
  }
  for (c = 0; c < nol; c++) {                                          /* balance.c:121-134 */
    if ((aver > BAL * dists[c]) && ((classes.freq[c] + diff[c]) > 1))
      if (note < 0) { diff[c]--; note++; }
    if (BAL * aver < dists[c])
      if (note > 0) { diff[c]++; note--; }
  }

  /* new vectors: diff[c] more for the classes with diff > 0, picked in data order */
  hit_init(&more);
  total = 0;
  for (c = 0; c < nol; c++)
    for (t = 0; t < diff[c]; t++) { hit_add(&more, classes.label[c]); total++; }
  inside = (unsigned char *)calloc((size_t)(data->n > 0 ? data->n : 1), 1);
  picked = (long *)malloc(sizeof(long) * (size_t)(data->n > 0 ? data->n : 1));
  data_label = (int32_t *)malloc(sizeof(int32_t) * (size_t)(data->n > 0 ? data->n : 1));
  if (!inside || !picked || !data_label) return 1;
  for (i = 0; i < data->n; i++) data_label[i] = pak_label(data, i);
  if (verbose_level >= 1) fprintf(stderr, "Some codebook vectors are removed\n");
  if (verbose_level >= 1) fprintf(stderr, "Some new codebook vectors are picked\n");
  if (total > 0) {
    if (find_winners(data, data, knn, &w)) return 1;
    hit_init(&hits);
    for (i = 0; i < data->n; i++) {
      if (w.nfound[i] != knn) { inside[i] = 1; continue; }            /* lvq_rout.c:173 */
      hit_clear(&hits);
      for (t = 0; t < knn; t++) hit_add(&hits, pak_label(data, w.idx[i * knn + t]));
      inside[i] = hits.n > 0 && hits.label[0] == data_label[i];
    }
    pick_inside(data, inside, &more, picked, &npicked);
    hit_free(&hits);
    winners_free(&w);
  }

  /* the balanced codebook: survivors in list order (the first -diff entries of a class go,
   * balance.c:141-163), then the picked data entries */
  out = pak_alloc(codes->dim, codes->n + npicked);
  out_label = (int32_t *)malloc(sizeof(int32_t) * (size_t)(codes->n + npicked + 1));
  unit_alpha = (float *)malloc(sizeof(float) * (size_t)(codes->n + npicked + 1));
  if (!out || !out_label || !unit_alpha) return 1;
  out->topol = codes->topol; out->neigh = codes->neigh; out->xdim = codes->xdim; out->ydim = codes->ydim;
  if (codes->mask || data->mask) out->mask = (unsigned char *)calloc((size_t)(codes->n + npicked + 1) * codes->dim, 1);
  out->lab_pool = (int *)malloc(sizeof(int) * (size_t)(codes->lab_off[codes->n] + data->lab_off[data->n] + 1));
  if (!out->lab_pool || ((codes->mask || data->mask) && !out->mask)) return 1;
  for (i = 0; i < codes->n; i++) {
    long l;
    for (c = 0; c < nol; c++)
      if (classes.label[c] == label[i]) break;
    if (c < nol && diff[c] < 0) { diff[c]++; continue; }
    memcpy(out->points + (size_t)kept * codes->dim, codes->points + (size_t)i * codes->dim, sizeof(float) * codes->dim);
    if (codes->mask) memcpy(out->mask + (size_t)kept * codes->dim, codes->mask + (size_t)i * codes->dim, (size_t)codes->dim);
    for (l = codes->lab_off[i]; l < codes->lab_off[i + 1]; l++) out->lab_pool[nl++] = codes->lab_pool[l];
    out_label[kept] = label[i];
    unit_alpha[kept] = 0.3f;
    kept++;
    out->lab_off[kept] = nl;
  }
  counted = kept;                                                      /* what the reference's num_entries says */
  for (i = 0; i < npicked; i++) {
    const long j = picked[i];
    long l;
    memcpy(out->points + (size_t)kept * codes->dim, data->points + (size_t)j * codes->dim, sizeof(float) * codes->dim);
    if (data->mask) memcpy(out->mask + (size_t)kept * codes->dim, data->mask + (size_t)j * codes->dim, (size_t)codes->dim);
    for (l = data->lab_off[j]; l < data->lab_off[j + 1]; l++) out->lab_pool[nl++] = data->lab_pool[l];
    out_label[kept] = data_label[j];
    unit_alpha[kept] = 0.0f;                                           /* read from beyond the rate array: see above */
    kept++;
    out->lab_off[kept] = nl;
  }
  out->n = kept;

  if (verbose_level >= 1) fprintf(stderr, "Codebook vectors are redistributed\n");
  if (data->n > 0 && out->n > 0) {                                     /* balance.c:196-202: one pass of OLVQ1 */
    sample = (int32_t *)malloc(sizeof(int32_t) * (size_t)data->n);
    if (!sample) return 1;
    bmu_lvq_schedule(0, data->n, data->n, 0.3f, BMU_ALPHA_LINEAR, data->n, NULL, sample, dists + codes->n);
    if (bmu_lvq_train(BMU_OLVQ1, out->points, out_label, out->n, out->dim, data->points, data->mask, data_label,
                      data->n, sample, NULL, data->n, 0.0f, 0.0f, 0.3f, unit_alpha))
      return engine_failed("bmu_lvq_train");
    free(sample);
  }
  lra_name(cout_name, lra, sizeof lra);                                /* alpha_write, datafile.c:1061-1086 */
  {
    FILE *fp = fopen(lra, "w+");
    if (!fp) fprintf(stderr, "Can't open alpha file %s for writing", lra);
    else {
      for (i = 0; i < counted; i++) fprintf(fp, "%g\n", unit_alpha[i]);
      fclose(fp);
    }
  }
  if (verbose_level >= 2) fprintf(stderr, "Medians of the shortest distances are computed\n");
  if (class_medians(out, out_label, &classes, dists)) return 1;
  if (verbose_level > 0)
    for (c = 0; c < classes.n; c++)
      fprintf(stdout, "In class %9s %3d units, min dist.: %.3f\n", label_string((int)classes.label[c]),
              (int)classes.freq[c], dists[c]);
  if (verbose_level >= 2) fprintf(stderr, "Codebook entries are saved to file %s\n", cout_name);
  if (pak_save(out, cout_name)) return 1;
  free(label); free(dists); free(diff); free(inside); free(picked); free(data_label); free(out_label); free(unit_alpha);
  hit_free(&classes); hit_free(&more);
  pak_free(data); pak_free(codes); pak_free(out);
  return 0;
}

/* ------------------------------------------------------------------ sammon */
/* sammon.c:420-490: remove_identicals (83-127) + sammon_iterate (129-262) + save_entries.  The
 * O(M^2 D) distance loops and the O(M^2) sweeps run on the device (bmu_identical_pairs, bmu_sammon);
 * the removal walk with its message numbering, the generator for the initial positions and the file
 * stay here.  The PostScript picture (-eps / -ps) is not produced. */
int sammon_main(int argc, char **argv) {
  struct pak_entries *codes, *out;
  const char *cin_name, *cout_name, *s;
  long length, i, j, noc, nl = 0, npairs = 0, cap;
  unsigned long next;
  int32_t *pairs;
  unsigned char *alive;
  float *x, *y, *err = NULL;
  global_options(argc, argv);
  cin_name = need(argc, argv, "-cin");
  cout_name = need(argc, argv, "-cout");
  length = atol(need(argc, argv, "-rlen"));
  s = opt(argc, argv, "-rand");
  next = (unsigned long)(int)(s ? atol(s) : 0);
  if (!next) next = (unsigned long)(int)time(NULL);                   /* init_random, lvq_pak.c:478-484 */
  if (flag(argc, argv, "-eps") || flag(argc, argv, "-ps"))
    fprintf(stderr, "note: the PostScript picture (-eps / -ps) is not provided by the B200 host\n");
  if (verbose_level >= 2) fprintf(stderr, "Code entries from file %s\n", cin_name);
  codes = pak_load(cin_name, 0, 1);
  if (!codes) { fprintf(stderr, "can't open code file %s\n", cin_name); return 1; }
  if (bmu_init(0)) return engine_failed("bmu_init");

  /* remove_identicals: for the entry at place ii of the CURRENT list, every later entry still in the
   * list at distance 0 goes; the counter ij steps by two after a removal (sammon.c:104-118) */
  cap = 4 * codes->n + 1024;
  pairs = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cap);
  alive = (unsigned char *)malloc((size_t)(codes->n > 0 ? codes->n : 1));
  if (!pairs || !alive) return 1;
  if (bmu_identical_pairs(codes->points, codes->mask, codes->n, codes->dim, pairs, cap, &npairs)) {
    if (npairs <= cap) return engine_failed("bmu_identical_pairs");
    free(pairs);                                                      /* a codebook full of duplicates */
    cap = npairs;
    pairs = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cap);
    if (!pairs || bmu_identical_pairs(codes->points, codes->mask, codes->n, codes->dim, pairs, cap, &npairs))
      return engine_failed("bmu_identical_pairs");
  }
  memset(alive, 1, (size_t)(codes->n > 0 ? codes->n : 1));
  {
    long p = 0, before = 0, at = 0;                                   /* alive entries in front of `at` */
    while (p < npairs) {
      const long cur = pairs[2 * p];
      long q = p, ij;
      while (q < npairs && pairs[2 * q] == cur) q++;
      for (; at < cur; at++) before += alive[at];
      if (alive[cur]) {
        const long ii = before + 1;
        long t = p;
        ij = ii + 1;
        for (j = cur + 1; j < codes->n && t < q; j++) {
          if (!alive[j]) continue;
          while (t < q && pairs[2 * t + 1] < j) t++;
          if (t < q && pairs[2 * t + 1] == j) {
            fprintf(stderr, "Identical entries in codebook ");
            fprintf(stderr, "(entries %d, %d), removing one.\n", (int)ii, (int)ij);
            alive[j] = 0;
            ij += 2;
          } else {
            ij++;
          }
        }
      }
      p = q;
    }
  }
  noc = 0;
  for (i = 0; i < codes->n; i++) noc += alive[i];

  out = pak_alloc(2, noc);
  x = (float *)malloc(sizeof(float) * (size_t)(noc > 0 ? noc : 1));
  y = (float *)malloc(sizeof(float) * (size_t)(noc > 0 ? noc : 1));
  if (verbose_level >= 2 && length > 0) err = (float *)malloc(sizeof(float) * (size_t)length);
  if (!out || !x || !y) return 1;
  out->topol = codes->topol; out->neigh = codes->neigh; out->xdim = codes->xdim; out->ydim = codes->ydim;
  out->lab_pool = (int *)malloc(sizeof(int) * (size_t)(codes->lab_off[codes->n] > 0 ? codes->lab_off[codes->n] : 1));
  if (!out->lab_pool) return 1;
  for (i = 0, j = 0; i < codes->n; i++) {                              /* compact the survivors in place */
    long l;
    if (!alive[i]) continue;
    if (j != i) {
      memmove(codes->points + (size_t)j * codes->dim, codes->points + (size_t)i * codes->dim, sizeof(float) * codes->dim);
      if (codes->mask) memmove(codes->mask + (size_t)j * codes->dim, codes->mask + (size_t)i * codes->dim, (size_t)codes->dim);
    }
    for (l = codes->lab_off[i]; l < codes->lab_off[i + 1]; l++) out->lab_pool[nl++] = codes->lab_pool[l];
    j++;
    out->lab_off[j] = nl;
  }
  for (i = 0; i < noc; i++) {                                          /* sammon.c:159-162 */
    next = (next * 23) % 100000001;                                   /* orand, lvq_pak.c:470-473 */
    x[i] = (float)((long)(int)(next % 32767L) % noc) / noc;
    y[i] = (float)(i) / noc;
  }
  if (noc > 0 && bmu_sammon(codes->points, codes->mask, noc, codes->dim, length, x, y, err))
    return engine_failed("bmu_sammon");
  if (err)
    for (i = 0; i < length; i++) fprintf(stdout, "Mapping error: %7.3f\n", err[i]);
  for (i = 0; i < noc; i++) { out->points[2 * i] = x[i]; out->points[2 * i + 1] = y[i]; }
  if (verbose_level >= 2) fprintf(stderr, "Save code entries to file %s\n", cout_name);
  if (pak_save(out, cout_name)) return 1;
  free(pairs); free(alive); free(x); free(y); free(err);
  pak_free(codes);
  pak_free(out);
  return 0;
}

/* ------------------------------------------------------------------ pakstat */
/* load only: entries, dimension, a checksum of the values and the time the loader took */
int pakstat_main(int argc, char **argv) {
  struct pak_entries *e;
  struct timespec t0, t1;
  double sum = 0.0;
  long i, nmask = 0;
  global_options(argc, argv);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  e = pak_load(need(argc, argv, "-din"), 0, !flag(argc, argv, "-noskip"));
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (!e) return 1;
  for (i = 0; i < e->n * e->dim; i++) { sum += e->points[i]; if (e->mask && e->mask[i]) nmask++; }
  if (opt(argc, argv, "-rawout")) {                               /* the parsed values as raw float32 */
    FILE *fp = fopen(opt(argc, argv, "-rawout"), "wb");
    if (!fp || fwrite(e->points, sizeof(float), (size_t)e->n * e->dim, fp) != (size_t)e->n * e->dim) return 1;
    fclose(fp);
  }
  fprintf(stdout, "entries %ld dim %d masked %ld labels %ld sum %.9g load_seconds %.4f\n", e->n, e->dim, nmask,
          e->lab_off[e->n], sum, (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
  pak_free(e);
  return 0;
}

/* ------------------------------------------------------------------ paksynth */
/* a synthetic .dat / map file for wall-time measurements (tools/bench_cli_qerror.py): the counter-based
 * generator of bench.py (splitmix64(seed, index) -> 24-bit uniform in [0,1)), written with the package's
 * own "%g " entry grammar */
int paksynth_main(int argc, char **argv) {
  const long rows = atol(need(argc, argv, "-rows"));
  const int dim = atoi(need(argc, argv, "-dim"));
  const char *s = opt(argc, argv, "-seed"), *xd = opt(argc, argv, "-xdim");
  const unsigned long long seed = s ? (unsigned long long)atoll(s) : 1ULL;
  FILE *fp = fopen(need(argc, argv, "-dout"), "w");
  char *buf;
  long r;
  int c;
  if (!fp || rows < 0 || dim < 1) return 1;
  buf = (char *)malloc((size_t)dim * 16 + 16);
  if (!buf) return 1;
  if (xd) fprintf(fp, "%d hexa %d %ld bubble\n", dim, atoi(xd), rows / atoi(xd));
  else fprintf(fp, "%d\n", dim);
  for (r = 0; r < rows; r++) {
    size_t n = 0;
    for (c = 0; c < dim; c++) {
      unsigned long long z = (unsigned long long)(r * (long)dim + c) + seed * 0x632BE59BD9B4E019ULL;
      z *= 0x9E3779B97F4A7C15ULL;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
      z ^= z >> 31;
      n += (size_t)sprintf(buf + n, "%g ", (double)((float)((z >> 40) & 0xFFFFFF) * (1.0f / 16777216.0f)));
    }
    buf[n++] = '\n';
    fwrite(buf, 1, n, fp);
  }
  free(buf);
  return fclose(fp) != 0;
}

/* ------------------------------------------------------------------ pakcat */
int pakcat_main(int argc, char **argv) {
  struct pak_entries *e;
  const char *b;
  global_options(argc, argv);
  b = opt(argc, argv, "-buffer");
  if (b) {
    /* through the streamed reader: chunks of N entries, written out one after the other */
    struct pak_stream *st = pak_stream_open(need(argc, argv, "-din"), 0, !flag(argc, argv, "-noskip"), atol(b));
    const char *dout = need(argc, argv, "-dout");
    FILE *fp;
    long chunks = 0;
    if (!st) return 1;
    fp = fopen(dout, "w");
    if (!fp) return 1;
    pak_write_header(fp, pak_stream_header(st));
    while ((e = pak_stream_next(st)) != NULL) {
      if (atol(b) > 0 && e->n > atol(b)) { fprintf(stderr, "chunk of %ld entries\n", e->n); return 1; }
      pak_write_entries(fp, e);
      pak_free(e);
      chunks++;
    }
    fclose(fp);
    if (pak_stream_failed(st)) return 1;
    pak_stream_close(st);
    fprintf(stderr, "%ld chunks\n", chunks);
    return 0;
  }
  e = pak_load(need(argc, argv, "-din"), 0, !flag(argc, argv, "-noskip"));
  if (!e) return 1;
  if (pak_save(e, need(argc, argv, "-dout"))) return 1;
  pak_free(e);
  return 0;
}
