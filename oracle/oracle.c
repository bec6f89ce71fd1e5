/* oracle.c -- CPU restatement of the SOM_PAK / LVQ_PAK best-matching-unit path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Written from the behaviour of the reference
 * on flat row-major arrays instead of linked lists; every function cites the reference
 * lines it restates.  Must be compiled WITHOUT fp contraction (-ffp-contract=off, no
 * -march=native, no -ffast-math): the reference is scalar SSE sub/mul/add with a rounding
 * after every operation (FLT_EVAL_METHOD == 0).
 *
 * Parity status: PINNED against the compiled reference and tests/golden/ (oracle.h).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* ---------------------------------------------------------------- RNG + sample order */

static unsigned long orc_next = 1;

/* lvq_pak.c:464-473 : next = next*23 mod 100000001 ; value = next mod 32767 */
void orc_osrand(int seed) { orc_next = (unsigned long)seed; }

long orc_orand(void)
{
  orc_next = (orc_next * 23UL) % 100000001UL;
  return (long)(int)(orc_next % 32767UL);
}

/* datafile.c:1152-1188 : one pass, position i swapped with position orand() % n.
 * (partner < 32767 always -- a quirk that is reproduced, not fixed) */
void orc_shuffle_order(long n, int seed, int *order)
{
  long i, j;
  int t;
  for (i = 0; i < n; i++) order[i] = (int)i;
  orc_osrand(seed);
  for (i = 0; i < n; i++) {
    j = orc_orand() % n;
    t = order[i]; order[i] = order[j]; order[j] = t;
  }
}

/* ---------------------------------------------------------------- winner search */

/* squared distance accumulated in index order with the reference's early exit:
 * the loop stops as soon as the partial sum exceeds `bound` (lvq_pak.c:63-73 / 184-195).
 * *nmasked counts the masked components seen before the loop ended. */
static float partial_dist(const float *m, const float *x, const unsigned char *mask,
                          int D, float bound, int *nmasked)
{
  float acc = 0.0f, d;
  int i, nm = 0;
  for (i = 0; i < D; i++) {
    if (mask && mask[i]) { nm++; continue; }
    d = m[i] - x[i];
    acc += d * d;
    if (acc > bound) break;
  }
  *nmasked = nm;
  return acc;
}

int orc_find_winner(const float *codes, long M, int D, const float *x,
                    const unsigned char *mask, int k, int *idx, float *diff)
{
  long j;
  int i, t, nm;
  float acc;

  if (k == 1) {                          /* lvq_pak.c:41-94 (and 160-161) */
    float best = FLT_MAX;
    idx[0] = -1; diff[0] = -1.0f;
    for (j = 0; j < M; j++) {
      acc = partial_dist(codes + j * (long)D, x, mask, D, best, &nm);
      if (nm == D) return 0;
      if (acc < best) { best = acc; idx[0] = (int)j; diff[0] = acc; }  /* strict: first min wins */
    }
    return 1;
  }
  for (i = 0; i < k; i++) { idx[i] = -1; diff[i] = FLT_MAX; }          /* lvq_pak.c:165-170 */
  for (j = 0; j < M; j++) {
    acc = partial_dist(codes + j * (long)D, x, mask, D, diff[k - 1], &nm);
    if (nm == D) return 0;
    for (i = 0; i < k && acc > diff[i]; i++) ;                          /* lvq_pak.c:197 */
    if (i < k) {                                                        /* equal goes BEFORE */
      for (t = k - 1; t > i; t--) { diff[t] = diff[t - 1]; idx[t] = idx[t - 1]; }
      diff[i] = acc; idx[i] = (int)j;
    }
  }
  return k;
}

void orc_search(const float *codes, long M, int D, const float *data,
                const unsigned char *mask, long N, int k, int *idx, float *diff, int *ret)
{
  long n;
  for (n = 0; n < N; n++)
    ret[n] = orc_find_winner(codes, M, D, data + n * (long)D,
                             mask ? mask + n * (long)D : NULL, k,
                             idx + n * (long)k, diff + n * (long)k);
}

/* lvq_pak.c:291-316 : masked in either vector => skipped; sqrt in double, float result */
float orc_vector_dist(const float *a, const unsigned char *ma, const float *b,
                      const unsigned char *mb, int D)
{
  float acc = 0.0f, d;
  int i, nm = 0;
  for (i = 0; i < D; i++) {
    if ((ma && ma[i]) || (mb && mb[i])) { nm++; continue; }
    d = a[i] - b[i];
    acc += d * d;
  }
  if (nm == D) return -1.0f;
  return (float)sqrt((double)acc);
}

/* lvq_pak.c:339-351 */
void orc_adapt_vector(float *c, const float *x, const unsigned char *mask, int D, float alpha)
{
  int i;
  for (i = 0; i < D; i++) {
    if (mask && mask[i]) continue;
    c[i] += alpha * (x[i] - c[i]);
  }
}

/* ---------------------------------------------------------------- lattice + schedules */

/* som_rout.c:434-455 */
float orc_hexa_dist(int bx, int by, int tx, int ty)
{
  float dx = (float)(bx - tx), dy, r;
  if (((by - ty) % 2) != 0) {
    if ((by % 2) == 0) dx = (float)((double)dx - 0.5);
    else               dx = (float)((double)dx + 0.5);
  }
  r = dx * dx;
  dy = (float)(by - ty);
  r = (float)((double)r + 0.75 * (double)dy * (double)dy);
  return (float)sqrt((double)r);
}

/* som_rout.c:457-468 */
float orc_rect_dist(int bx, int by, int tx, int ty)
{
  float dx = (float)(bx - tx), dy = (float)(by - ty), r;
  r = dx * dx;
  r += dy * dy;
  return (float)sqrt((double)r);
}

/* lvq_pak.c:903-906 */
float orc_linear_alpha(long iter, long length, float alpha)
{
  return alpha * (float)(length - iter) / (float)length;
}

/* lvq_pak.c:908-921 : c = length/100.0f ; alpha*c/(c+iter) all in float */
float orc_inverse_t_alpha(long iter, long length, float alpha)
{
  float c = (float)length / 100.0f;
  return alpha * c / (c + (float)iter);
}

static float map_dist(int topol, int bx, int by, int tx, int ty)
{
  return topol == ORC_TOPOL_RECT ? orc_rect_dist(bx, by, tx, ty) : orc_hexa_dist(bx, by, tx, ty);
}

/* gaussian weight, som_rout.c:541-542 : -dd*dd in float, 2.0*r*r in double, exp in double */
static float gauss_weight(float dd, float radius)
{
  float num = -dd * dd;
  return (float)exp((double)num / (2.0 * (double)radius * (double)radius));
}

/* ---------------------------------------------------------------- SOM training */

int orc_som_train(float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                  const float *data, const unsigned char *mask, const short *weight,
                  const short *fixed_xy, long N, const int *order,
                  long length, float alpha, float radius, int alpha_type)
{
  return orc_som_train_prefix(codes, M, D, xdim, ydim, topol, neigh, data, mask, weight, fixed_xy,
                              N, order, length, length, alpha, radius, alpha_type);
}

/* the first `nsteps` steps of a run of `length` steps (schedules depend on length) */
int orc_som_train_prefix(float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                         const float *data, const unsigned char *mask, const short *weight,
                         const short *fixed_xy, long N, const int *order,
                         long length, long nsteps, float alpha, float radius, int alpha_type)
{
  long le, pos = 0, u;
  (void)ydim;
  if (N <= 0) return 1;
  for (le = 0; le < nsteps; le++, pos++) {
    long s;
    const float *x;
    const unsigned char *mk;
    float trad, talp, w;
    int bx, by, widx;
    float wdiff;

    if (pos == N) pos = 0;                                     /* som_rout.c:602-610 */
    s = order ? order[pos] : pos;
    x = data + s * (long)D;
    mk = mask ? mask + s * (long)D : NULL;

    /* som_rout.c:615 : double expression rounded to float on assignment */
    trad = (float)(1.0 + ((double)radius - 1.0) * (double)(float)(length - le) / (double)(float)length);
    talp = alpha_type == ORC_ALPHA_INVERSE_T ? orc_inverse_t_alpha(le, length, alpha)
                                             : orc_linear_alpha(le, length, alpha);
    w = weight ? (float)weight[s] : 0.0f;
    if (weight && w > 0.0f)                                    /* som_rout.c:622-624 */
      talp = (float)(1.0 - (double)(float)pow(1.0 - (double)talp, (double)w));

    if (fixed_xy && fixed_xy[2 * s] >= 0) {                    /* som_rout.c:628-632 */
      bx = fixed_xy[2 * s]; by = fixed_xy[2 * s + 1];
    } else {
      if (orc_find_winner(codes, M, D, x, mk, 1, &widx, &wdiff) == 0) continue;
      bx = widx % xdim; by = widx / xdim;                       /* som_rout.c:641-642 */
    }
    for (u = 0; u < M; u++) {
      int tx = (int)(u % xdim), ty = (int)(u / xdim);
      float dd = map_dist(topol, bx, by, tx, ty);
      if (neigh == ORC_NEIGH_GAUSSIAN) {                         /* som_rout.c:511-549 */
        float a = talp * gauss_weight(dd, trad);
        orc_adapt_vector(codes + u * (long)D, x, mk, D, a);
      } else if (dd <= trad) {                                   /* som_rout.c:472-506 */
        orc_adapt_vector(codes + u * (long)D, x, mk, D, talp);
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- quantization error */

float orc_qerror(const float *codes, long M, int D, int xdim, int ydim, int topol, int neigh,
                 const float *data, const unsigned char *mask, long N, int qetype, float radius)
{
  float q = 0.0f;
  long n, u;
  (void)ydim;
  for (n = 0; n < N; n++) {
    const float *x = data + n * (long)D;
    const unsigned char *mk = mask ? mask + n * (long)D : NULL;
    int widx; float wdiff;
    if (orc_find_winner(codes, M, D, x, mk, 1, &widx, &wdiff) == 0) continue;
    if (!qetype) {                                              /* som_rout.c:715 */
      q = (float)((double)q + sqrt((double)wdiff));
    } else {                                                    /* som_rout.c:734-819,870-875 */
      int bx = widx % xdim, by = widx / xdim;
      float e = 0.0f;
      for (u = 0; u < M; u++) {
        int tx = (int)(u % xdim), ty = (int)(u / xdim);
        float dd = map_dist(topol, bx, by, tx, ty), d;
        if (neigh == ORC_NEIGH_GAUSSIAN) {
          float a = gauss_weight(dd, radius);
          d = orc_vector_dist(codes + u * (long)D, NULL, x, mk, D);
          e += a * d * d;
        } else if (dd <= radius) {
          d = orc_vector_dist(codes + u * (long)D, NULL, x, mk, D);
          e += d * d;
        }
      }
      q += e;
    }
  }
  return q;
}

/* ---------------------------------------------------------------- LVQ training */

int orc_lvq_train(int algo, float *codes, const int *code_label, long M, int D,
                  const float *data, const unsigned char *mask, const int *data_label, long N,
                  const int *order, long length, float alpha, int alpha_type,
                  float winlen, float epsilon, float *unit_alpha)
{
  long le, pos = 0;
  /* lvq_rout.c:770,876 : (1-winlen)/(1+winlen) in float */
  float wthr = (1 - winlen) / (1 + winlen);
  if (N <= 0) return 1;
  for (le = 0; le < length; le++, pos++) {
    long s;
    const float *x;
    const unsigned char *mk;
    int idx[2], dl;
    float diff[2], talp;

    if (pos == N) pos = 0;
    s = order ? order[pos] : pos;
    x = data + s * (long)D;
    mk = mask ? mask + s * (long)D : NULL;
    dl = data_label[s];
    talp = alpha_type == ORC_ALPHA_INVERSE_T ? orc_inverse_t_alpha(le, length, alpha)
                                             : orc_linear_alpha(le, length, alpha);

    if (algo == 1 || algo == 4) {
      float *c;
      if (orc_find_winner(codes, M, D, x, mk, 1, idx, diff) == 0 || idx[0] < 0) continue;
      c = codes + idx[0] * (long)D;
      if (algo == 1) {                                          /* lvq_rout.c:542-555 */
        orc_adapt_vector(c, x, mk, D, code_label[idx[0]] == dl ? talp : -talp);
      } else {                                                  /* lvq_rout.c:650-673 */
        float *ta = unit_alpha + idx[0];
        if (code_label[idx[0]] == dl) {
          orc_adapt_vector(c, x, mk, D, *ta);
          *ta = *ta / (1 + *ta);
        } else {
          orc_adapt_vector(c, x, mk, D, -*ta);
          *ta = *ta / (1 - *ta);
          if (*ta > alpha) *ta = alpha;
        }
      }
    } else {                                                    /* lvq_rout.c:750-781,855-896 */
      int b, nb, l0, l1;
      if (orc_find_winner(codes, M, D, x, mk, 2, idx, diff) == 0 || idx[0] < 0 || idx[1] < 0)
        continue;
      b = idx[0]; nb = idx[1];
      l0 = code_label[b]; l1 = code_label[nb];
      if (l0 != l1) {
        if ((l0 == dl || l1 == dl) && (diff[0] / diff[1]) > wthr) {
          if (l1 == dl) { int t = b; b = nb; nb = t; }
          orc_adapt_vector(codes + b * (long)D, x, mk, D, talp);
          orc_adapt_vector(codes + nb * (long)D, x, mk, D, -talp);
        }
      } else if (algo == 3 && l0 == dl) {
        orc_adapt_vector(codes + b * (long)D, x, mk, D, talp * epsilon);
        orc_adapt_vector(codes + nb * (long)D, x, mk, D, talp * epsilon);
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- hitlist vote */

/* labels.c:370-410 : list kept ordered by frequency; a label moves ahead of its
 * predecessor only while the predecessor's count is strictly smaller. */
long orc_hitlist_vote(const long *labels, int n)
{
  long *lab, *frq, r;
  int cnt = 0, i, p;
  if (n <= 0) return -1;
  lab = malloc(sizeof(long) * n);
  frq = malloc(sizeof(long) * n);
  for (i = 0; i < n; i++) {
    for (p = 0; p < cnt && lab[p] != labels[i]; p++) ;
    if (p == cnt) { lab[cnt] = labels[i]; frq[cnt] = 1; cnt++; continue; }
    frq[p]++;
    while (p > 0 && frq[p - 1] < frq[p]) {
      long tl = lab[p - 1], tf = frq[p - 1];
      lab[p - 1] = lab[p]; frq[p - 1] = frq[p];
      lab[p] = tl; frq[p] = tf;
      p--;
    }
  }
  r = lab[0];
  free(lab); free(frq);
  return r;
}

/* ---------------------------------------------------------------- class distances */
/* lvq_rout.c:280-361 (min_distances: mean) and 375-473 (med_distances: median) with the class
 * list built by add_hit (labels.c:370-410).  near/found (nullable, M each) receive dissf / fou
 * of every entry.  Returns the number of classes. */
static int cmp_float_asc(const void *a, const void *b)
{
  float x = *(const float *)a, y = *(const float *)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

long orc_class_dists(const float *codes, const unsigned char *mask, const int *label, long M, int D,
                     int median, int *out_class, int *out_noe, float *out_dists,
                     float *near, int *found)
{
  long ncls = 0, c, i, j, k;
  long *cls = malloc(sizeof(long) * (M > 0 ? M : 1)), *freq = malloc(sizeof(long) * (M > 0 ? M : 1));
  float *meds = malloc(sizeof(float) * (M > 0 ? M : 1));
  for (i = 0; i < M; i++) {                       /* add_hit */
    for (k = 0; k < ncls && cls[k] != label[i]; k++) ;
    if (k == ncls) { cls[ncls] = label[i]; freq[ncls++] = 1; continue; }
    freq[k]++;
    while (k > 0 && freq[k - 1] < freq[k]) {
      long t = cls[k]; cls[k] = cls[k - 1]; cls[k - 1] = t;
      t = freq[k]; freq[k] = freq[k - 1]; freq[k - 1] = t;
      k--;
    }
  }
  for (c = 0; c < ncls; c++) {
    long note = 0;
    float sum = 0.0f;
    for (i = 0; i < M; i++) {
      float dissf = FLT_MAX;
      int fou = 0;
      if (label[i] != cls[c]) continue;
      for (j = i + 1; j < M; j++) {
        float dist;
        if (label[j] != cls[c]) continue;
        fou = 1;
        dist = orc_vector_dist(codes + j * (long)D, mask ? mask + j * (long)D : NULL,
                               codes + i * (long)D, mask ? mask + i * (long)D : NULL, D);
        if (dist < dissf) dissf = dist;
      }
      if (near) near[i] = dissf;
      if (found) found[i] = fou;
      if (fou) { sum += dissf; meds[note++] = dissf; }
    }
    out_class[c] = (int)cls[c];
    out_noe[c] = (int)freq[c];
    out_dists[c] = 0.0f;
    if (note > 0) {
      if (median) { qsort(meds, note, sizeof(float), cmp_float_asc); out_dists[c] = meds[note / 2]; }
      else out_dists[c] = sum / note;
    }
  }
  free(cls); free(freq); free(meds);
  return ncls;
}

/* ---------------------------------------------------------------- Sammon's mapping */
/* sammon.c:83-127: keep[i] = 0 for the entries remove_identicals drops (a later entry still in the
 * list at distance exactly 0 from the current one).  Returns the number kept. */
long orc_remove_identicals(const float *codes, const unsigned char *mask, long M, int D, int *keep)
{
  long i, j, n = 0;
  for (i = 0; i < M; i++) keep[i] = 1;
  for (i = 0; i < M; i++) {
    if (!keep[i]) continue;
    for (j = i + 1; j < M; j++) {
      if (!keep[j]) continue;
      if (orc_vector_dist(codes + i * (long)D, mask ? mask + i * (long)D : NULL,
                          codes + j * (long)D, mask ? mask + j * (long)D : NULL, D) == 0.0)
        keep[j] = 0;
    }
  }
  for (i = 0; i < M; i++) n += keep[i];
  return n;
}

/* sammon.c:129-262 from given initial positions (the reference draws them at 159-162).
 * err: nullable, `length` mapping errors (240-254). */
int orc_sammon(const float *codes, const unsigned char *mask, long noc, int D, long length,
               float *x, float *y, float *err)
{
  long i, j, k, mutual;
  float e1x, e1y, e2x, e2y, dpj, dq, dr, dt, xd, yd, xx, yy, e, tot, d, ee;
  float *xu = malloc(sizeof(float) * (noc > 0 ? noc : 1)), *yu = malloc(sizeof(float) * (noc > 0 ? noc : 1));
  float *dd = malloc(sizeof(float) * (noc > 1 ? (size_t)noc * (noc - 1) / 2 : 1));
  if (!xu || !yu || !dd) return 1;
  mutual = 0;
  for (j = 1; j < noc; j++)
    for (k = 0; k < j; k++)
      dd[mutual++] = orc_vector_dist(codes + j * (long)D, mask ? mask + j * (long)D : NULL,
                                     codes + k * (long)D, mask ? mask + k * (long)D : NULL, D);
  for (i = 0; i < length; i++) {
    for (j = 0; j < noc; j++) {
      e1x = e1y = e2x = e2y = 0.0;
      for (k = 0; k < noc; k++) {
        if (j == k) continue;
        xd = x[j] - x[k];
        yd = y[j] - y[k];
        dpj = (float)sqrt((double)xd * xd + yd * yd);
        if (k > j) dt = dd[k * (k - 1) / 2 + j];
        else dt = dd[j * (j - 1) / 2 + k];
        dq = dt - dpj;
        dr = dt * dpj;
        e1x += xd * dq / dr;
        e1y += yd * dq / dr;
        e2x += (dq - xd * xd * (1.0 + dq / dpj) / dpj) / dr;
        e2y += (dq - yd * yd * (1.0 + dq / dpj) / dpj) / dr;
      }
      xu[j] = x[j] + 0.2 * e1x / fabs(e2x);
      yu[j] = y[j] + 0.2 * e1y / fabs(e2y);
    }
    xx = yy = 0.0;
    for (j = 0; j < noc; j++) { xx += xu[j]; yy += yu[j]; }
    xx /= noc;
    yy /= noc;
    for (j = 0; j < noc; j++) { x[j] = xu[j] - xx; y[j] = yu[j] - yy; }
    if (err) {
      e = tot = 0.0;
      mutual = 0;
      for (j = 1; j < noc; j++)
        for (k = 0; k < j; k++) {
          d = dd[mutual];
          tot += d;
          xd = x[j] - x[k];
          yd = y[j] - y[k];
          ee = d - (float)sqrt((double)xd * xd + yd * yd);
          e += (ee * ee / d);
          mutual++;
        }
      e /= tot;
      err[i] = e;
    }
  }
  free(xu); free(yu); free(dd);
  return 0;
}
