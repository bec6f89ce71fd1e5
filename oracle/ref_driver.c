/* ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin driver of OUR OWN that links against the UNMODIFIED reference objects compiled by
 * oracle/Makefile (`make ref`) from /root/reference.  It converts flat row-major arrays
 * into the reference's linked lists (struct entries / struct data_entry, lvq_pak.h:73-113)
 * and then calls the reference's own functions -- find_winner_euc / find_winner_knn
 * (lvq_pak.c:41-94,152-221), som_training (som_rout.c:556-671), find_qerror / find_qerror2
 * (som_rout.c:678-731,823-891), lvq1/olvq1/lvq2/lvq3_training (lvq_rout.c:498-916) and
 * randomize_entry_order (datafile.c:1152-1188) -- so that their outputs can be used
 *   (1) to pin oracle/oracle.c (our restatement) and generate tests/golden/ fixtures, and
 *   (2) as the `"kind": "reference"` CPU baseline of bench.py.
 * Nothing under som_lvq_pak_b200/ may link or load this file.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lvq_pak.h"
#include "datafile.h"
#include "labels.h"
#include "som_rout.h"
#include "lvq_rout.h"

/* Build an in-memory entries list from a flat array.  totlen_known=1 keeps
 * rewind_entries (datafile.c:790-840) from trying to read a file. */
static struct entries *build_entries(const float *x, const unsigned char *mask,
                                     const int *label, const short *weight,
                                     const short *fixed_xy, long n, int dim,
                                     int topol, int neigh, int xdim, int ydim,
                                     struct data_entry ***table_out)
{
  struct entries *e = alloc_entries();
  struct data_entry *prev = NULL, **tbl;
  long i;
  int d;

  if (!e) return NULL;
  e->dimension = dim;
  e->topol = topol;
  e->neigh = neigh;
  e->xdim = xdim;
  e->ydim = ydim;
  e->flags.loadmode = LOADMODE_ALL;
  e->flags.totlen_known = 1;
  e->num_entries = n;
  e->num_loaded = n;
  tbl = malloc(sizeof(*tbl) * (n > 0 ? n : 1));
  for (i = 0; i < n; i++) {
    struct data_entry *de = alloc_entry(e);
    memcpy(de->points, x + i * (long)dim, sizeof(float) * dim);
    if (mask) {
      int any = 0;
      for (d = 0; d < dim; d++) any |= mask[i * (long)dim + d];
      if (any) {
        de->mask = malloc(dim);
        memcpy(de->mask, mask + i * (long)dim, dim);
      }
    }
    if (label) set_entry_label(de, label[i]);
    if (weight) de->weight = weight[i];
    if (fixed_xy && fixed_xy[2 * i] >= 0) {
      de->fixed = malloc(sizeof(struct fixpoint));
      de->fixed->xfix = fixed_xy[2 * i];
      de->fixed->yfix = fixed_xy[2 * i + 1];
    }
    if (prev) prev->next = de; else e->entries = de;
    prev = de;
    tbl[i] = de;
  }
  if (table_out) *table_out = tbl; else free(tbl);
  return e;
}

static void scatter_back(struct entries *e, float *x, long n, int dim)
{
  struct data_entry *de = e->entries;
  long i;
  for (i = 0; i < n && de; i++, de = de->next)
    memcpy(x + i * (long)dim, de->points, sizeof(float) * dim);
}

/* -------- batch winner search through the reference's per-sample functions -------- */
/* ret[n] = the function's return value (0 = all components masked). idx/diff: N x k. */
int ref_search(const float *codes, long M, int D, const float *data,
               const unsigned char *mask, long N, int k,
               int *idx, float *diff, int *ret)
{
  struct entries *ce, *de;
  struct data_entry *s;
  struct winner_info *win;
  long n;
  int j;

  verbose(0);
  ce = build_entries(codes, NULL, NULL, NULL, NULL, M, D, TOPOL_LVQ, 0, 0, 0, NULL);
  de = build_entries(data, mask, NULL, NULL, NULL, N, D, TOPOL_DATA, 0, 0, 0, NULL);
  win = malloc(sizeof(*win) * (k > 0 ? k : 1));
  for (n = 0, s = de->entries; s; s = s->next, n++) {
    for (j = 0; j < k; j++) { win[j].index = -7; win[j].diff = -7.0f; win[j].winner = NULL; }
    if (k == 1) ret[n] = find_winner_euc(ce, s, win, 1);
    else        ret[n] = find_winner_knn(ce, s, win, k);
    for (j = 0; j < k; j++) {
      idx[n * k + j] = (int)win[j].index;
      diff[n * k + j] = win[j].diff;
    }
  }
  free(win);
  close_entries(ce);
  close_entries(de);
  return 0;
}

/* timing variant: no outputs kept except a checksum, so the loop is the reference's own */
double ref_search_time_only(const float *codes, long M, int D, const float *data, long N, int k,
                            long *checksum)
{
  struct entries *ce, *de;
  struct data_entry *s;
  struct winner_info win[64];
  struct timespec t0, t1;
  long cs = 0;

  verbose(0);
  ce = build_entries(codes, NULL, NULL, NULL, NULL, M, D, TOPOL_LVQ, 0, 0, 0, NULL);
  de = build_entries(data, NULL, NULL, NULL, NULL, N, D, TOPOL_DATA, 0, 0, 0, NULL);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (s = de->entries; s; s = s->next) {
    if (k == 1) find_winner_euc(ce, s, win, 1);
    else        find_winner_knn(ce, s, win, k);
    cs += win[0].index;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (checksum) *checksum = cs;
  close_entries(ce);
  close_entries(de);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* -------- sample order produced by `-rand seed` (datafile.c:340-341,1152-1188) -------- */
int ref_shuffle_order(long N, int seed, int *order)
{
  struct entries *de;
  struct data_entry **tbl, *s;
  float *tag = malloc(sizeof(float) * N);
  long i;
  for (i = 0; i < N; i++) tag[i] = 0.0f;
  de = build_entries(tag, NULL, NULL, NULL, NULL, N, 1, TOPOL_DATA, 0, 0, 0, &tbl);
  /* tag each node with its original position through the weight-free label slot */
  for (i = 0; i < N; i++) { tbl[i]->lab.label = (int)i; tbl[i]->num_labs = 1; }
  init_random(seed);
  de->entries = randomize_entry_order(de->entries);
  for (i = 0, s = de->entries; s; s = s->next, i++) order[i] = s->lab.label;
  free(tbl); free(tag);
  close_entries(de);
  return 0;
}

long ref_orand_after_seed(int seed, int ncalls)
{
  long v = 0; int i;
  init_random(seed);
  for (i = 0; i < ncalls; i++) v = orand();
  return v;
}

/* -------- SOM training: the reference's som_training on in-memory lists -------- */
/* rand_seed < 0: list order (no -rand); otherwise init_random(seed) + one shuffle, exactly
 * what vsom.c:168-172 + read_entries do.  weight/fixed_xy may be NULL. */
int ref_som_train(float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                  const float *data, const unsigned char *mask, const short *weight,
                  const short *fixed_xy, long N,
                  long length, float alpha, float radius, int alpha_type, int rand_seed)
{
  struct entries *ce, *de;
  struct teach_params params;

  verbose(0);
  use_weights(weight ? 1 : 0);
  use_fixed(fixed_xy ? 1 : 0);
  ce = build_entries(codes, NULL, NULL, NULL, NULL, M, D, topol, neigh, xdim, ydim, NULL);
  de = build_entries(data, mask, NULL, weight, fixed_xy, N, D, TOPOL_DATA, 0, 0, 0, NULL);
  memset(&params, 0, sizeof(params));
  set_teach_params(&params, ce, de, 0, NULL);
  set_som_params(&params);
  if (rand_seed >= 0) {
    init_random(rand_seed);
    de->entries = randomize_entry_order(de->entries);
  }
  params.alpha_type = alpha_type;
  params.alpha_func = (alpha_type == ALPHA_INVERSE_T) ? inverse_t_alpha : linear_alpha;
  params.length = length;
  params.alpha = alpha;
  params.radius = radius;
  if (som_training(&params) == NULL) return 1;
  scatter_back(ce, codes, M, D);
  close_entries(ce);
  close_entries(de);
  use_weights(0);
  use_fixed(0);
  return 0;
}

/* qetype 0: find_qerror, 1: find_qerror2 (neighbourhood weighted; radius used) */
float ref_qerror(const float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                 const float *data, const unsigned char *mask, long N, int qetype, float radius)
{
  struct entries *ce, *de;
  struct teach_params params;
  float q;

  verbose(0);
  ce = build_entries(codes, NULL, NULL, NULL, NULL, M, D, topol, neigh, xdim, ydim, NULL);
  de = build_entries(data, mask, NULL, NULL, NULL, N, D, TOPOL_DATA, 0, 0, 0, NULL);
  memset(&params, 0, sizeof(params));
  set_teach_params(&params, ce, de, 0, NULL);
  set_som_params(&params);
  params.radius = radius;
  q = qetype ? find_qerror2(&params) : find_qerror(&params);
  close_entries(ce);
  close_entries(de);
  return q;
}

/* -------- LVQ training -------- */
/* algo: 1 lvq1, 2 lvq2, 3 lvq3, 4 olvq1.  lra_base: file base name for the olvq1 alpha
 * file (written as <base>.lra by alpha_write, datafile.c:1061-1086); may be NULL for
 * algo != 4.  olvq1 with alpha == 0 reads <lra_in>.lra (lvq_rout.c:614-627). */
int ref_lvq_train(int algo, float *codes, const int *code_label, long M, int D,
                  const float *data, const unsigned char *mask, const int *data_label, long N,
                  long length, float alpha, int alpha_type, float winlen, float epsilon,
                  int rand_seed, const char *lra_in, const char *lra_out)
{
  struct entries *ce, *de, *r = NULL;
  struct teach_params params;

  verbose(0);
  ce = build_entries(codes, NULL, code_label, NULL, NULL, M, D, TOPOL_LVQ, 0, 0, 0, NULL);
  de = build_entries(data, mask, data_label, NULL, NULL, N, D, TOPOL_DATA, 0, 0, 0, NULL);
  memset(&params, 0, sizeof(params));
  set_teach_params(&params, ce, de, 0, NULL);
  if (rand_seed >= 0) {
    init_random(rand_seed);
    de->entries = randomize_entry_order(de->entries);
  }
  params.alpha_type = alpha_type;
  params.alpha_func = (alpha_type == ALPHA_INVERSE_T) ? inverse_t_alpha : linear_alpha;
  params.length = length;
  params.alpha = alpha;
  switch (algo) {
  case 1: r = lvq1_training(&params); break;
  case 2: params.winner = find_winner_knn; r = lvq2_training(&params, winlen); break;
  case 3: params.winner = find_winner_knn; r = lvq3_training(&params, epsilon, winlen); break;
  case 4: r = olvq1_training(&params, (char *)lra_in, (char *)lra_out); break;
  default: break;
  }
  if (r == NULL) return 1;
  scatter_back(ce, codes, M, D);
  close_entries(ce);
  close_entries(de);
  return 0;
}

/* scalar helpers so the oracle's restatements can be pinned one by one */
float ref_hexa_dist(int bx, int by, int tx, int ty) { return hexa_dist(bx, by, tx, ty); }
float ref_rect_dist(int bx, int by, int tx, int ty) { return rect_dist(bx, by, tx, ty); }
float ref_linear_alpha(long it, long len, float a) { return linear_alpha(it, len, a); }
float ref_inverse_t_alpha(long it, long len, float a) { return inverse_t_alpha(it, len, a); }

float ref_vector_dist(const float *a, const unsigned char *ma, const float *b,
                      const unsigned char *mb, int dim)
{
  struct data_entry ea, eb;
  memset(&ea, 0, sizeof(ea)); memset(&eb, 0, sizeof(eb));
  ea.points = (float *)a; ea.mask = (char *)ma;
  eb.points = (float *)b; eb.mask = (char *)mb;
  return vector_dist_euc(&ea, &eb, dim);
}

/* majority vote via the reference's hitlist (labels.c:370-410): returns head label */
long ref_hitlist_vote(const long *labels, int n)
{
  struct hitlist *h = new_hitlist();
  long r = -1; int i;
  for (i = 0; i < n; i++) add_hit(h, labels[i]);
  if (h->head) r = h->head->label;
  free_hitlist(h);
  return r;
}

/* min_distances / med_distances (lvq_rout.c:280-492) on a flat codebook; near/found are not
 * available from the reference (it only returns the per-class values).  Returns #classes. */
long ref_class_dists(const float *codes, const unsigned char *mask, const int *label, long M, int D,
                     int median, int *out_class, int *out_noe, float *out_dists,
                     float *near, int *found)
{
  struct entries *e = build_entries(codes, mask, label, NULL, NULL, M, D, TOPOL_LVQ, NEIGH_UNKNOWN, 0, 0, NULL);
  struct mindists *md;
  long i, n;
  (void)near; (void)found;
  if (!e) return -1;
  md = median ? med_distances(e, vector_dist_euc) : min_distances(e, vector_dist_euc);
  if (!md) return -1;
  n = md->num_classes;
  for (i = 0; i < n; i++) {
    out_class[i] = md->class[i];
    out_noe[i] = md->noe[i];
    out_dists[i] = md->dists[i];
  }
  free_mindists(md);
  close_entries(e);
  return n;
}

/* the reference's sammon.c compiled with -Dmain=ref_sammon_main (oracle/Makefile): init_random(seed),
 * remove_identicals, sammon_iterate -- exactly what its main() does (sammon.c:462-467).  x/y receive
 * the positions of the surviving entries; returns how many survived, -1 on failure. */
struct entries *remove_identicals(struct entries *codes, int *rem);
struct entries *sammon_iterate(struct entries *codes, int length);

long ref_sammon(const float *codes, const unsigned char *mask, long M, int D, long length, int seed,
                float *x, float *y)
{
  struct entries *e = build_entries(codes, mask, NULL, NULL, NULL, M, D, TOPOL_LVQ, NEIGH_UNKNOWN, 0, 0, NULL);
  struct entries *sp;
  struct data_entry *de;
  eptr p;
  int removed, old = verbose(-1);
  long n = 0;
  if (!e) return -1;
  verbose(0);
  init_random(seed);
  e = remove_identicals(e, &removed);
  sp = sammon_iterate(e, (int)length);
  verbose(old);
  if (!sp) return -1;
  for (de = rewind_entries(sp, &p); de != NULL; de = next_entry(&p), n++) {
    x[n] = de->points[0];
    y[n] = de->points[1];
  }
  close_entries(e);
  return n;
}
