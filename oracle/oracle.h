/* oracle.h -- CPU restatement of the SOM_PAK/LVQ_PAK best-matching-unit path.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product (som_lvq_pak_b200/)
 * never links, loads or calls anything declared here.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py checks every function below
 * against (a) the unmodified reference compiled by `make -C oracle ref`
 * (oracle/_ref/libref_driver.so, when present) and (b) golden vectors under
 * tests/golden/ that were generated from that same build by tests/golden/make_golden.py.
 */
#ifndef SOMLVQ_ORACLE_H
#define SOMLVQ_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* topology / neighbourhood / alpha codes: values of lvq_pak.h:206-224 */
#define ORC_TOPOL_HEXA 3
#define ORC_TOPOL_RECT 4
#define ORC_NEIGH_BUBBLE 1
#define ORC_NEIGH_GAUSSIAN 2
#define ORC_ALPHA_LINEAR 1
#define ORC_ALPHA_INVERSE_T 2

/* lvq_pak.c:459-473 */
void orc_osrand(int seed);
long orc_orand(void);
/* datafile.c:1152-1188: order[i] = original position of the i-th entry after `-rand seed` */
void orc_shuffle_order(long n, int seed, int *order);

/* lvq_pak.c:41-94 (k==1) and 152-221 (k>=2). idx/diff hold k slots. returns 0 if all
 * components are masked, else k. */
int orc_find_winner(const float *codes, long M, int D, const float *x,
                    const unsigned char *mask, int k, int *idx, float *diff);
void orc_search(const float *codes, long M, int D, const float *data,
                const unsigned char *mask, long N, int k, int *idx, float *diff, int *ret);

/* lvq_pak.c:291-316 */
float orc_vector_dist(const float *a, const unsigned char *ma, const float *b,
                      const unsigned char *mb, int D);
/* lvq_pak.c:339-351 */
void orc_adapt_vector(float *c, const float *x, const unsigned char *mask, int D, float alpha);
/* som_rout.c:434-468 */
float orc_hexa_dist(int bx, int by, int tx, int ty);
float orc_rect_dist(int bx, int by, int tx, int ty);
/* lvq_pak.c:903-921 */
float orc_linear_alpha(long iter, long length, float alpha);
float orc_inverse_t_alpha(long iter, long length, float alpha);

/* som_rout.c:556-671 with bubble_adapt 472-506 / gaussian_adapt 511-549.
 * order[N]: list order of the samples (identity if no -rand). weight/fixed_xy nullable. */
int orc_som_train(float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                  const float *data, const unsigned char *mask, const short *weight,
                  const short *fixed_xy, long N, const int *order,
                  long length, float alpha, float radius, int alpha_type);

int orc_som_train_prefix(float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                         const float *data, const unsigned char *mask, const short *weight,
                         const short *fixed_xy, long N, const int *order,
                         long length, long nsteps, float alpha, float radius, int alpha_type);

/* som_rout.c:678-731 (qetype 0) and 734-891 (qetype 1) */
float orc_qerror(const float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                 const float *data, const unsigned char *mask, long N, int qetype, float radius);

/* lvq_rout.c:498-916.  algo 1 lvq1, 2 lvq2, 3 lvq3, 4 olvq1.  unit_alpha[M] is the olvq1
 * per-unit rate state (in/out); alpha is the cap (lvq_rout.c:670-672). */
int orc_lvq_train(int algo, float *codes, const int *code_label, long M, int D,
                  const float *data, const unsigned char *mask, const int *data_label, long N,
                  const int *order, long length, float alpha, int alpha_type,
                  float winlen, float epsilon, float *unit_alpha);

/* labels.c:370-410 majority vote: the label that first reaches the final maximum count,
 * scanning in rank order. */
long orc_hitlist_vote(const long *labels, int n);

/* lvq_rout.c:280-361 (median=0: min_distances, mean) / 375-473 (median=1: med_distances); classes in
 * add_hit order (labels.c:370-410).  near/found nullable: dissf / fou per entry.  Returns #classes. */
long orc_class_dists(const float *codes, const unsigned char *mask, const int *label, long M, int D,
                     int median, int *out_class, int *out_noe, float *out_dists,
                     float *near, int *found);

/* sammon.c:83-127: keep[i] = 0 for the entries remove_identicals drops; returns the number kept */
long orc_remove_identicals(const float *codes, const unsigned char *mask, long M, int D, int *keep);
/* sammon.c:129-262 from the given initial positions; err nullable (`length` mapping errors, 240-254) */
int orc_sammon(const float *codes, const unsigned char *mask, long noc, int D, long length,
               float *x, float *y, float *err);

#ifdef __cplusplus
}
#endif
#endif
