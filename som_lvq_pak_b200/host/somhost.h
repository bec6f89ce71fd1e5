/* somhost.h -- C host layer of the B200 BMU engine: the data model, file formats and program
 * loops of SOM_PAK / LVQ_PAK that sit on either side of the hot path, calling libbmu_b200
 * (include/bmu.h) instead of the per-sample function pointers of struct teach_params.
 *
 * The reference keeps `struct entries` as a linked list of `struct data_entry`
 * (reference lvq_pak.h:73-113) and reaches its winner function once per sample.  Here a file
 * is loaded ONCE into flat arrays (what bmu_search / bmu_som_train / bmu_lvq_train take), the
 * per-sample loop becomes one batch call, and the consumers replay the results in data order,
 * which is what makes stdout and the output files byte-identical to the reference programs.
 * File grammar: SURVEY.md appendix B (datafile.c:112-148, 552-748, 396-447).
 */
#ifndef SOMHOST_H
#define SOMHOST_H

#include <stdio.h>

#define TOPOL_UNKNOWN 0
#define TOPOL_DATA 1
#define TOPOL_LVQ 2
#define TOPOL_HEXA 3
#define TOPOL_RECT 4
#define NEIGH_UNKNOWN 0
#define NEIGH_BUBBLE 1
#define NEIGH_GAUSSIAN 2
#define LABEL_EMPTY 0

/* ---- label table (reference labels.c:36-128): index 0 is the empty label ---------- */
int label_index(const char *str);        /* adds the label when it is new */
const char *label_string(int ind);       /* NULL for LABEL_EMPTY / unknown */
void label_reset(void);                   /* forget every label (between programs of one process) */

/* ---- entries: one .dat / .cod file in flat arrays ------------------------------------- */
struct pak_entries {
  int dim, topol, neigh, xdim, ydim;
  long n;                  /* number of entries kept (all-masked ones are dropped unless asked) */
  float *points;           /* n x dim, masked components stored as 0 (datafile.c:623,659) */
  unsigned char *mask;     /* n x dim, non-zero = component ignored; NULL when the file has none */
  long *lab_off;           /* n + 1 offsets into lab_pool: entry i has labels lab_pool[lab_off[i] .. lab_off[i+1]) */
  int *lab_pool;
  short *weight;           /* n, `weight=N` terms (0 when absent) */
  short *fixed_xy;         /* n x 2, `fixed=x,y` terms (-1,-1 when absent) */
};

/* labels_needed: a line without a label is an error (datafile.c:737-745);
 * skip_empty: drop entries whose components are all masked (datafile.c:677-690) */
struct pak_entries *pak_load(const char *name, int labels_needed, int skip_empty);
/* streamed reading, the reference's `-buffer N` (datafile.c:237-344): the file is handed out in chunks of
 * N entries (N <= 0: one chunk, the whole file); the next block of text is read and parsed on a helper
 * thread while the caller works on the current chunk */
struct pak_stream;
struct pak_stream *pak_stream_open(const char *name, int labels_needed, int skip_empty, long buffer);
const struct pak_entries *pak_stream_header(const struct pak_stream *st);   /* dim, topology (no rows) */
struct pak_entries *pak_stream_next(struct pak_stream *st);                 /* NULL: end of file or error */
int pak_stream_failed(const struct pak_stream *st);
void pak_stream_close(struct pak_stream *st);
struct pak_entries *pak_alloc(int dim, long n);
void pak_free(struct pak_entries *e);
int pak_save(const struct pak_entries *e, const char *name);
void pak_write_header(FILE *fp, const struct pak_entries *e);
void pak_write_entries(FILE *fp, const struct pak_entries *e);   /* the entry lines, datafile.c:420-447 */
/* first label of entry i (get_entry_label, labels.h:45) */
int pak_label(const struct pak_entries *e, long i);
/* replace the labels of every entry: nlab[i] labels taken from labs (concatenated) */
int pak_set_labels(struct pak_entries *e, const int *nlab, const int *labs);
extern const char *pak_mask_string;      /* "x" unless -mask_str / LVQSOM_MASK_STR */

/* ---- hitlist (labels.c:286-444): (label, count) pairs, highest count first; a label that
 * ties with the one in front of it stays behind it */
struct pak_hitlist {
  long n, cap;
  long *label, *freq;
};
void hit_init(struct pak_hitlist *h);
void hit_clear(struct pak_hitlist *h);
void hit_free(struct pak_hitlist *h);
long hit_add(struct pak_hitlist *h, long label);
long hit_freq(const struct pak_hitlist *h, long label);

/* ---- programs: same options, stdout and output files as the reference's ----------------- */
int vsom_main(int argc, char **argv);
int qerror_main(int argc, char **argv);
int visual_main(int argc, char **argv);
int vcal_main(int argc, char **argv);
int accuracy_main(int argc, char **argv);
int classify_main(int argc, char **argv);
int knntest_main(int argc, char **argv);
int cmatr_main(int argc, char **argv);
int setlabel_main(int argc, char **argv);
int elimin_main(int argc, char **argv);
int lvqtrain_main(int argc, char **argv, const char *progname);
int vfind_main(int argc, char **argv);
int randinit_main(int argc, char **argv, const char *progname);
int eveninit_main(int argc, char **argv, const char *progname);
int mindist_main(int argc, char **argv);
int sammon_main(int argc, char **argv);
int balance_main(int argc, char **argv);
int pakcat_main(int argc, char **argv);   /* load + save: exercises the file layer alone */
int pakstat_main(int argc, char **argv);
int paksynth_main(int argc, char **argv); /* synthetic .dat / map files for wall-time measurements */  /* load only, prints counts and the load time */

#endif
