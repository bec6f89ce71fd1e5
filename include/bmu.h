/* bmu.h -- C ABI of the B200 best-matching-unit engine for SOM_PAK / LVQ_PAK.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The
 * reference reaches its hot path one sample at a time through the function pointers of
 * struct teach_params (reference lvq_pak.h:131-148,186-204):
 *     WINNER_FUNCTION  find_winner_euc / find_winner_knn   (lvq_pak.c:41-94, 152-221)
 *     VECTOR_ADAPT     adapt_vector                        (lvq_pak.c:339-351)
 *     NEIGH_ADAPT      bubble_adapt / gaussian_adapt       (som_rout.c:472-549)
 * A GPU cannot be fed one sample at a time, so the entry points below replace the loops
 * that call those pointers (one level up):
 *     bmu_search*        <- the per-sample loops of find_qerror (som_rout.c:710-721),
 *                           compute_accuracy (accuracy.c:82), compute_classifications
 *                           (classify.c:66), compute_knnaccuracy (knntest.c:98), find_labels
 *                           (vcal.c:109), compute_visual_data (visual.c:113), compute_cmatr
 *                           (cmatr.c:84), scan_data_traj (planes.c:242), correct_by_knn
 *                           (lvq_rout.c:61), elimin.c:81, setlabel.c:73
 *     bmu_som_train*     <- som_training             (som_rout.c:556-671)
 *     bmu_lvq_train*     <- lvq1/olvq1/lvq2/lvq3_training (lvq_rout.c:498-916)
 *     bmu_qerror2        <- find_qerror2 + bubble/gaussian_qerror (som_rout.c:734-891)
 *     bmu_search_stats*  <- the sums/counts those callers accumulate, for the multi-GPU
 *                           all-reduce
 * Results are bit-identical to the reference for winner indices and squared distances
 * (exact FP32: sub, mul, add with a rounding after every operation, summed in component
 * order; first minimum wins for k = 1, (distance asc, index desc) for k >= 2).
 *
 * There is NO CPU fallback: every entry point returns BMU_ERR_NODEV / BMU_ERR_CUDA when no
 * sm_100 device is usable.  All functions return 0 on success, a BMU_ERR_* code otherwise;
 * bmu_last_error() gives the message.  Entry points are called from one host thread.
 */
#ifndef BMU_H
#define BMU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMU_OK 0
#define BMU_ERR_CUDA 1   /* a CUDA call failed (message in bmu_last_error)            */
#define BMU_ERR_ARG 2    /* bad argument                                              */
#define BMU_ERR_NOMEM 3  /* host or device allocation failed                          */
#define BMU_ERR_NODEV 4  /* no CUDA device / not an sm_100 device                     */

#define BMU_KMAX 16      /* largest k of bmu_search (reference uses k <= 10, elimin.c:30) */

/* values of the reference's enums (lvq_pak.h:206-224) so that hosts can pass theirs through */
#define BMU_TOPOL_HEXA 3
#define BMU_TOPOL_RECT 4
#define BMU_NEIGH_BUBBLE 1
#define BMU_NEIGH_GAUSSIAN 2
#define BMU_ALPHA_LINEAR 1
#define BMU_ALPHA_INVERSE_T 2
#define BMU_LVQ1 1
#define BMU_LVQ2 2
#define BMU_LVQ3 3
#define BMU_OLVQ1 4

/* which kernel family a search may use (bmu_set_search_path); AUTO picks per shape */
#define BMU_PATH_AUTO 0
#define BMU_PATH_EXACT 1   /* K1: direct FP32 difference-squared kernels only             */
#define BMU_PATH_FILTER 2  /* K2: tcgen05 GEMM filter + exact FP32 re-rank (+K1 fallback) */

/* ---- library / device ------------------------------------------------------------ */
int bmu_init(int device);                 /* select device, create streams; idempotent  */
void bmu_shutdown(void);
const char *bmu_last_error(void);
int bmu_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin);
int bmu_set_search_path(int path);
/* number of kernels this library has launched since bmu_init (bench.py: gpu_launches) */
long bmu_launch_count(void);
/* counters of the last bmu_search_dev call (of the last chunk for bmu_search): [0] rows in the
 * call, [1] rows answered by the warp-per-sample exact kernel (masked / tiny-magnitude rows,
 * k >= 2 on the exact path, K2 certificate failures), [2] rows answered by the sequential
 * emulation kernel (NaN/Inf), [3] rows certified by the K2 filter, [4] K2 rows whose
 * certificate failed (re-done exactly, included in [1]).  Synchronises the device. */
int bmu_last_search_breakdown(long out[5]);
/* device time (ms, CUDA events on the launching stream) of the kernels of the last search of
 * each family: K1 [0] data_prep, [1] k1_fast, [2] k1_warp, [3] k1_seq; K2 [4] row_prep,
 * [5] gemm + fused top-k, [6] exact re-rank, [7] K1 fallback lists.  Synchronises. */
int bmu_last_search_kernel_ms(float out[8]);
/* the same for the search `back` calls ago within each family (0 = the last one; zeros beyond the 32
 * calls that are kept), so that a timed loop can be read after its closing synchronisation */
int bmu_search_kernel_ms_history(int back, float out[8]);

/* ---- codebook (replicated on every GPU; reference: struct entries *codes) ---------
 * All kernel-side images (K1 tiles, the fp16 operands of the K2 filter) are built synchronously
 * inside create / update. */
typedef struct bmu_codebook bmu_codebook;
bmu_codebook *bmu_codebook_create(const float *codes, long M, int D);        /* host ptr  */
bmu_codebook *bmu_codebook_create_dev(const float *d_codes, long M, int D);  /* device ptr */
int bmu_codebook_update(bmu_codebook *cb, const float *codes);               /* host ptr  */
void bmu_codebook_destroy(bmu_codebook *cb);

/* ---- batch winner search ---------------------------------------------------------- */
/* data: N x D row-major; mask: NULL or N x D bytes (non-zero = component ignored, only the
 * SAMPLE's mask counts, lvq_pak.c:65-69); idx/diff: N x k; nfound: N (the reference's return
 * value: 0 = every component masked, else k).  Unfilled slots: idx -1, diff -1.0 (k == 1)
 * or FLT_MAX (k >= 2), exactly what the reference leaves in struct winner_info.
 * Host-pointer version: copies in chunks overlapped with compute (this is bench.py's e2e). */
int bmu_search(bmu_codebook *cb, const float *data, const unsigned char *mask, long N, int k,
               int32_t *idx, float *diff, int32_t *nfound);
/* device-pointer version, asynchronous on `stream` (a cudaStream_t, NULL = default).  All searches
 * of one device share one scratch set (work lists, counters, operand images): the library orders
 * every search after the one before it with an event, whatever streams they were launched on, so
 * calls from several streams are safe but do not overlap each other. */
int bmu_search_dev(bmu_codebook *cb, const float *d_data, const unsigned char *d_mask, long N,
                   int k, int32_t *d_idx, float *d_diff, int32_t *d_nfound, void *stream);

/* per-shard partial statistics of a finished search, to be summed over GPUs by ONE small
 * all-reduce (reference accumulators: find_qerror som_rout.c:710-721, find_labels
 * vcal.c:109-131, compute_accuracy accuracy.c:82-118, compute_cmatr cmatr.c:84-109).
 * *d_sum += sum of sqrt(diff[:,0]) over found rows -- a double, reduced on the device in a fixed
 * order (bit-identical from run to run for a given N; NOT the reference's sequential float sum:
 * the byte-exact qerror is bmu_replay_qerror over the gathered diffs in data order);
 * *d_nfound_total += number of found rows; d_hist (nullable): M int64 BMU hit counts;
 * d_confusion (nullable): n_labels x n_labels int64, [sample label][winner's label], labels from
 * d_sample_label[N] / d_code_label[M] in 0..n_labels-1.  All counters are int64 and accumulate (+=):
 * zero them before the first shard/chunk.  Asynchronous on `stream`. */
int bmu_search_stats_dev(const int32_t *d_idx, const float *d_diff, const int32_t *d_nfound,
                         long N, int k, long M, double *d_sum, long long *d_nfound_total,
                         long long *d_hist, const int32_t *d_sample_label,
                         const int32_t *d_code_label, int n_labels, long long *d_confusion,
                         void *stream);

/* ---- the data-parallel split: rows sharded over GPUs, codebook replicated ---- */
/* Host-side view of the combined statistics (bmu_multi_search).  hist / confusion: NULL = not wanted,
 * else caller arrays of M / n_labels*n_labels int64 that receive the totals over ALL shards. */
typedef struct bmu_stats {
  double sum_sqrt;              /* out: sum of sqrt(diff[:,0]) over found rows                 */
  long long n_found;            /* out: rows with a winner                                     */
  long long *hist;              /* out (nullable): BMU hits per code vector                    */
  long long *confusion;         /* out (nullable): [sample label][winner's label]              */
  const int32_t *sample_label;  /* in: N labels in 0..n_labels-1 (only with confusion)         */
  int n_labels;
} bmu_stats;

/* (A) one process, all GPUs -- what the C hosts use.  nshards = 0: $SOMLVQ_GPUS if set, else every
 * visible device.  One shard per device; asking for more shards than devices places the extra
 * (logical) shards round-robin on the devices, each with its own streams and scratch (used by the
 * tests to run the sharded path on one GPU; statistics of shards that share a device are added on
 * the device, devices are combined by NCCL).  Needs libnccl.so.2 at run time when > 1 device. */
int bmu_multi_init(int nshards);          /* a different shard count than before closes the old contexts: destroy
                                             the bmu_mcodebook handles made with them first */
int bmu_multi_shards(void);
int bmu_multi_devices(void);
/* rows [lo, hi) of shard `shard`: contiguous, balanced, cut at multiples of 512 rows */
void bmu_multi_shard_bounds(long N, int nshards, int shard, long *lo, long *hi);
typedef struct bmu_mcodebook bmu_mcodebook;
/* codes (host) -> device 0 -> ONE ncclBroadcast to the other devices */
bmu_mcodebook *bmu_mcodebook_create(const float *codes, long M, int D);
int bmu_mcodebook_update(bmu_mcodebook *cb, const float *codes);
int bmu_mcodebook_set_labels(bmu_mcodebook *cb, const int32_t *code_label);   /* M, for confusion */
void bmu_mcodebook_destroy(bmu_mcodebook *cb);
/* bmu_search over all shards: every shard streams its contiguous slice of the caller's rows through
 * its own GPU and writes idx/diff/nfound straight into the caller's arrays (data order kept, no
 * gather); stats (nullable) = totals over all shards after ONE grouped NCCL all-reduce of
 * {double sum} + {int64 n_found, hist[M], confusion[L*L]}.  Results are identical to bmu_search. */
int bmu_multi_search(bmu_mcodebook *cb, const float *data, const unsigned char *mask, long N, int k,
                     int32_t *idx, float *diff, int32_t *nfound, bmu_stats *stats);

/* (B) one process per GPU (torchrun / MPI style launchers): rank 0 makes the id, the launcher hands
 * the 128 bytes to every rank, every rank binds its bmu_init device.  The collectives run on device
 * buffers, asynchronously on `stream`; with no communicator (a single rank) they are no-ops. */
int bmu_comm_unique_id(unsigned char id[128]);
int bmu_comm_init_rank(int nranks, int rank, const unsigned char id[128]);
int bmu_comm_destroy(void);
int bmu_comm_broadcast_dev(void *d_buf, size_t bytes, int root, void *stream);   /* codebook replicate */
/* ONE grouped all-reduce (sum) of a double vector and an int64 vector, in place */
int bmu_comm_allreduce_stats_dev(double *d_sum, long nsum, long long *d_counts, long ncounts,
                                 void *stream);

/* ---- page-locked host memory -------------------------------------------------------- */
/* bmu_search / bmu_multi_search accept ANY host memory.  Pageable buffers (malloc/calloc, what the
 * reference's hosts have, datafile.c:472) are staged through a pinned ring by a few copy threads,
 * chunk by chunk, overlapped with the DMA and the kernels; buffers that are already page-locked are
 * read by the DMA engine directly.  A host that owns its allocation can skip the staging hop: */
void *bmu_host_alloc(size_t bytes);            /* page-locked; NULL on failure */
void bmu_host_free(void *p);
int bmu_host_register(void *p, size_t bytes);  /* pin an existing allocation in place */
int bmu_host_unregister(void *p);
/* threads that stage pageable memory (0 = default: min(8, cores / ranks on this box), or
 * $SOMLVQ_COPY_THREADS) */
int bmu_set_copy_threads(int n);

/* ---- online training (sequential; one GPU) ---------------------------------------- */
/* Schedules are produced on the host with the reference's own formulas (bmu_som_schedule /
 * bmu_lvq_schedule below) and passed as per-step arrays, so that the device never has to
 * reproduce libm's pow() and the sample order is the reference's -rand order:
 *   sample[t]  index of the data row used at step t (list order, wrap, shuffle resolved)
 *   talp[t]    learning rate of step t   (lvq_pak.c:903-921, som_rout.c:617-624)
 *   trad[t]    neighbourhood radius      (som_rout.c:615)
 * fixed_xy: NULL or N x 2 int16 (x,y) with x < 0 = no fixed point (som_rout.c:628-632).
 * codes (M x D, M = xdim*ydim) is updated in place.
 * Two documented deviations from the reference, both outside what its own loaders can produce:
 * (1) masks travel to the device as a NaN sentinel inside the data, so once ANY component of the
 * data set is masked, a genuine NaN in an unmasked component is treated as masked too (the
 * reference would propagate it into the codebook); (2) a step whose sample has no winner (every
 * component masked) is skipped, as som_rout.c:635-640 does -- data whose distance is NaN for every
 * unit (index -1 in the reference, which then adapts around unit -1) is skipped as well. */
int bmu_som_train(float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                  const float *data, const unsigned char *mask, long N,
                  const int16_t *fixed_xy, const int32_t *sample, const float *talp,
                  const float *trad, long nsteps);

/* algo BMU_LVQ1/2/3/OLVQ1 (lvq_rout.c:498-916).  win_thr = (1-win)/(1+win) in float
 * (lvq_rout.c:770); unit_alpha: OLVQ1 per-unit rates, M floats in/out (NULL otherwise);
 * alpha_cap: OLVQ1 cap (lvq_rout.c:670-672).  talp unused for OLVQ1 (may be NULL). */
int bmu_lvq_train(int algo, float *codes, const int32_t *code_label, long M, int D,
                  const float *data, const unsigned char *mask, const int32_t *data_label,
                  long N, const int32_t *sample, const float *talp, long nsteps,
                  float win_thr, float epsilon, float alpha_cap, float *unit_alpha);

/* resident trainer: keeps data + codebook on the device between chunks of steps, so a host
 * can stop at snapshot boundaries (som_rout.c:650-658) without re-uploading anything */
typedef struct bmu_trainer bmu_trainer;
bmu_trainer *bmu_trainer_create(const float *codes, long M, int D, const float *data,
                                const unsigned char *mask, long N);
int bmu_trainer_set_som(bmu_trainer *t, int xdim, int ydim, int topol, int neigh,
                        const int16_t *fixed_xy);
int bmu_trainer_set_lvq(bmu_trainer *t, int algo, const int32_t *code_label,
                        const int32_t *data_label, float win_thr, float epsilon,
                        float alpha_cap, const float *unit_alpha);
/* run steps [0, nsteps) of the given per-step arrays (host pointers) */
int bmu_trainer_steps(bmu_trainer *t, const int32_t *sample, const float *talp,
                      const float *trad, long nsteps);
int bmu_trainer_get_codes(bmu_trainer *t, float *codes);
int bmu_trainer_get_unit_alpha(bmu_trainer *t, float *unit_alpha);
/* device time of the last bmu_trainer_steps call in milliseconds (CUDA events) */
float bmu_trainer_last_ms(bmu_trainer *t);
void bmu_trainer_destroy(bmu_trainer *t);

/* ---- neighbourhood-weighted quantization error (qerror -qetype 1) ------------------ */
/* per-sample value of bubble_qerror / gaussian_qerror (som_rout.c:734-819) around the BMU;
 * out[n] for found rows, 0 for all-masked rows.  The host sums out[] in data order. */
int bmu_qerror2(bmu_codebook *cb, int xdim, int ydim, int topol, int neigh, float radius,
                const float *data, const unsigned char *mask, long N, float *out);

/* The pair loops of min_distances / med_distances (lvq_rout.c:325-350, 430-470; called by
 * mindist.c:93-105 and balance.c:81-83): for every code vector i, dist[i] = the smallest
 * vector_dist_euc(j, i) (lvq_pak.c:291-316, components masked in either vector skipped, -1 when
 * all are) over the LATER code vectors j > i whose class label[j] equals label[i]; found[i] = 0
 * and dist[i] = FLT_MAX when there is none (`fou` in the reference).  The per-class mean / median
 * of dist[] over found entries stays with the caller (it follows the class hitlist order).
 * mask: NULL or M x D. */
int bmu_class_nearest(const float *codes, const unsigned char *mask, const int32_t *label, long M, int D,
                      float *dist, int32_t *found);

/* remove_identicals (sammon.c:83-127): the pairs (i, j), i < j, of code vectors with
 * vector_dist_euc == 0.0, sorted by (i, j), in pairs[2*p], pairs[2*p+1]; *npairs = how many there
 * are (BMU_ERR_ARG when more than cap).  The caller replays the reference's removal walk. */
int bmu_identical_pairs(const float *codes, const unsigned char *mask, long M, int D, int32_t *pairs, long cap,
                        long *npairs);
/* sammon_iterate (sammon.c:129-262): `length` sweeps of Sammon's mapping of the M code vectors,
 * starting from the caller's initial positions (sammon.c:159-162: x[i] = (float)(orand() % M) / M,
 * y[i] = (float)i / M) and returning the final ones in place.  err: NULL, or `length` floats that
 * receive the mapping error the reference prints per sweep at -v 2 (sammon.c:240-254). */
int bmu_sammon(const float *codes, const unsigned char *mask, long M, int D, long length, float *x, float *y,
               float *err);

/* ---- host-side helpers (pure C arithmetic, no device) ------------------------------ */
/* lvq_pak.c:459-473 + datafile.c:1152-1188: order[i] = row used at list position i after
 * `-rand seed` (seed != 0; the reference maps seed 0 to time()). */
void bmu_rand_order(long n, int seed, int32_t *order);
/* the data row used at each of `nsteps` training steps (som_rout.c:602-610, lvq_rout.c:531-540): list
 * order walked cyclically; with seed != 0 (`-rand`) the list is shuffled when it is read; with
 * 0 < buffer <= N (`-buffer`) the file is held `buffer` entries at a time, every chunk is shuffled on its
 * own each time it is (re-)read and the generator state runs on (datafile.c:237-344).  Feed the result to
 * bmu_som_schedule / bmu_lvq_schedule as `order` with N = nsteps. */
void bmu_sample_sequence(long N, long buffer, int seed, long nsteps, int32_t *sample);
/* randinit_codes (som_rout.c:34-157): M = xdim*ydim code vectors drawn uniformly between the
 * per-component minimum and maximum of the (unmasked) data with the reference's generator seeded
 * by `seed` (init_random, lvq_pak.c:478-484); components without data become 0.  Used by the
 * multi-trial map search (vfind.c:247-306), one trial per seed. */
void bmu_randinit_codes(const float *data, const unsigned char *mask, long N, int D, long M, int seed,
                        float *codes);
/* fill sample/talp/trad for steps [le0, le1) of a run of `length` steps.  order: NULL =
 * identity; weight: NULL or N shorts (vsom -weights, som_rout.c:622-624). */
void bmu_som_schedule(long le0, long le1, long length, float alpha, float radius,
                      int alpha_type, long N, const int32_t *order, const int16_t *weight,
                      int32_t *sample, float *talp, float *trad);
void bmu_lvq_schedule(long le0, long le1, long length, float alpha, int alpha_type, long N,
                      const int32_t *order, int32_t *sample, float *talp);
/* find_qerror's accumulator (som_rout.c:697,715): float q += sqrt((double)diff) over the found
 * rows IN DATA ORDER -- the order-dependent float sum that makes qerror's stdout byte-exact.
 * diff is N x k (first column used). */
float bmu_replay_qerror(const float *diff, const int32_t *nfound, long N, int k);

#ifdef __cplusplus
}
#endif
#endif /* BMU_H */
