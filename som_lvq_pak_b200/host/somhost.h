/* somhost.h -- C host layer of the B200 BMU engine: the data model, file formats and loops of
 * SOM_PAK / LVQ_PAK that sit on either side of the hot path, calling libbmu_b200 (include/bmu.h).
 *
 * The struct and function names mirror the reference's interface (reference lvq_pak.h:73-124,
 * 186-204; datafile.h; labels.h; som_rout.h; lvq_rout.h) so that the reference's programs read
 * the same against this layer; the implementation is new: entries are loaded into flat arrays
 * once, the per-sample winner loop is one bmu_search() call, and the consumers replay the
 * results in data order (which is what makes stdout and the output files byte-identical).
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
#define ALPHA_LINEAR 1
#define ALPHA_INVERSE_T 2
#define LABEL_EMPTY 0

struct fixpoint { short xfix, yfix; };

struct data_entry {                 /* reference lvq_pak.h:73-87 */
  float *points;
  int *labels;                      /* label ids, num_labs of them */
  short num_labs;
  short weight;
  struct data_entry *next;
  char *mask;                       /* non-zero = component ignored */
  struct fixpoint *fixed;
};

struct entries {                    /* reference lvq_pak.h:89-113 */
  short dimension, topol, neigh, xdim, ydim;
  struct data_entry *entries;
  long num_entries;
  int skip_empty, labels_needed, random_order;
};

struct winner_info {                /* reference lvq_pak.h:120-124 */
  long index;
  struct data_entry *winner;
  float diff;
};

struct teach_params {               /* reference lvq_pak.h:186-204 (the fields the loops use) */
  short topol, neigh, alpha_type;
  float radius, alpha;
  long length;
  int knn;
  struct entries *codes, *data;
  long snap_interval;               /* 0 = no snapshots */
  const char *snap_file;
};

struct hit_entry { struct hit_entry *next, *prev; long label, freq; };
struct hitlist { struct hit_entry *head, *tail; long entries; };

/* ---- global switches (reference datafile.c:1310-1319, lvq_pak.c:486-495) */
int label_not_needed(int level);
int use_weights(int level);
int use_fixed(int level);
extern int verbose_level;
extern const char *masked_string;

/* ---- labels (labels.c) */
int find_conv_to_ind(const char *str);
const char *find_conv_to_lab(int ind);
int get_entry_label(const struct data_entry *e);
void set_entry_label(struct data_entry *e, int label);
void add_entry_label(struct data_entry *e, int label);
void clear_entry_labels(struct data_entry *e);
struct hitlist *new_hitlist(void);
void clear_hitlist(struct hitlist *hl);
void free_hitlist(struct hitlist *hl);
long add_hit(struct hitlist *hl, long label);
long hitlist_label_freq(struct hitlist *hl, long label);

/* ---- entries (datafile.c) */
struct entries *open_entries(const char *name);       /* loads the whole file */
struct entries *alloc_entries(void);
void close_entries(struct entries *e);
int save_entries(struct entries *e, const char *name);
int write_header(FILE *fp, const struct entries *e);
int write_entry(FILE *fp, const struct entries *e, const struct data_entry *d);
void init_random(int seed);
void randomize_entry_order(struct entries *e);          /* datafile.c:1152-1188 */

/* ---- the loops on the hot path, now batch calls into libbmu_b200 */
float find_qerror(struct teach_params *teach);                       /* som_rout.c:678-731 */
float find_qerror2(struct teach_params *teach);                      /* som_rout.c:823-891 */
struct entries *som_training(struct teach_params *teach);            /* som_rout.c:556-671 */
struct entries *lvq1_training(struct teach_params *teach);           /* lvq_rout.c:498-577 */
struct entries *olvq1_training(struct teach_params *teach, const char *in, const char *out);
struct entries *lvq2_training(struct teach_params *teach, float winlen);
struct entries *lvq3_training(struct teach_params *teach, float epsilon, float winlen);
/* batch WINNER_FUNCTION: win is N x knn, ret[n] = what find_winner_euc/knn would return */
int find_winners_batch(struct entries *codes, struct entries *data, int knn,
                       struct winner_info *win, int *ret);

/* ---- programs (one main each in the reference) */
int vsom_main(int argc, char **argv);
int qerror_main(int argc, char **argv);
int visual_main(int argc, char **argv);
int vcal_main(int argc, char **argv);
int accuracy_main(int argc, char **argv);
int classify_main(int argc, char **argv);
int knntest_main(int argc, char **argv);
int lvqtrain_main(int argc, char **argv, const char *progname);

#endif
