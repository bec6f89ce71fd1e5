/* entries.c -- label table, .dat/.cod reader and writer, hitlists (host layer, see somhost.h).
 *
 * Behaviour follows the reference's file layer (datafile.c, labels.c, fileio.c) so that files
 * written here are byte-identical to the reference's; the storage is flat arrays, not lists.
 */
#include "somhost.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------ label table */
static char **g_labels = NULL;
static int g_nlabels = 0, g_labcap = 0;
/* the streamed reader (pak_stream_*) interns the labels of the next chunk on its own thread while the
 * program prints labels of the current one */
static pthread_mutex_t g_label_lock = PTHREAD_MUTEX_INITIALIZER;

static int label_index_locked(const char *str) {
  int i;
  for (i = 0; i < g_nlabels; i++)
    if (strcmp(g_labels[i], str) == 0) return i + 1;
  if (g_nlabels == g_labcap) {
    char **t = (char **)realloc(g_labels, sizeof(char *) * (g_labcap + 128));
    if (!t) return -1;
    g_labels = t;
    g_labcap += 128;
  }
  g_labels[g_nlabels] = strdup(str);
  if (!g_labels[g_nlabels]) return -1;
  return ++g_nlabels;
}
int label_index(const char *str) {
  int r;
  if (str == NULL || str[0] == '\0') return LABEL_EMPTY;          /* labels.c:86-91 */
  pthread_mutex_lock(&g_label_lock);
  r = label_index_locked(str);
  pthread_mutex_unlock(&g_label_lock);
  return r;
}

void label_reset(void) {                   /* a fresh table per program of `bmu_pak batch` */
  int i;
  for (i = 0; i < g_nlabels; i++) free(g_labels[i]);
  g_nlabels = 0;
}

const char *label_string(int ind) {
  const char *r = NULL;
  pthread_mutex_lock(&g_label_lock);
  if (ind > 0 && ind <= g_nlabels) r = g_labels[ind - 1];    /* the strings themselves never move */
  pthread_mutex_unlock(&g_label_lock);
  return r;
}

/* ------------------------------------------------------------------ entries */
const char *pak_mask_string = "x";

struct pak_entries *pak_alloc(int dim, long n) {
  struct pak_entries *e = (struct pak_entries *)calloc(1, sizeof(*e));
  if (!e) return NULL;
  e->dim = dim;
  e->n = n;
  e->topol = TOPOL_DATA;
  e->points = (float *)calloc((size_t)(n > 0 ? n : 1) * dim, sizeof(float));
  e->lab_off = (long *)calloc((size_t)n + 1, sizeof(long));
  e->weight = (short *)calloc((size_t)(n > 0 ? n : 1), sizeof(short));
  e->fixed_xy = (short *)malloc((size_t)(n > 0 ? n : 1) * 2 * sizeof(short));
  if (!e->points || !e->lab_off || !e->weight || !e->fixed_xy) { pak_free(e); return NULL; }
  memset(e->fixed_xy, 0xff, (size_t)(n > 0 ? n : 1) * 2 * sizeof(short));
  return e;
}

void pak_free(struct pak_entries *e) {
  if (!e) return;
  free(e->points); free(e->mask); free(e->lab_off); free(e->lab_pool); free(e->weight); free(e->fixed_xy);
  free(e);
}

int pak_label(const struct pak_entries *e, long i) {
  return e->lab_off[i + 1] > e->lab_off[i] ? e->lab_pool[e->lab_off[i]] : LABEL_EMPTY;
}

int pak_set_labels(struct pak_entries *e, const int *nlab, const int *labs) {
  long i, tot = 0;
  int *pool;
  for (i = 0; i < e->n; i++) tot += nlab[i];
  pool = (int *)malloc(sizeof(int) * (size_t)(tot > 0 ? tot : 1));
  if (!pool) return 1;
  memcpy(pool, labs, sizeof(int) * (size_t)tot);
  free(e->lab_pool);
  e->lab_pool = pool;
  e->lab_off[0] = 0;
  for (i = 0; i < e->n; i++) e->lab_off[i + 1] = e->lab_off[i] + nlab[i];
  return 0;
}

static const char *topol_names[] = {NULL, "data", "lvq", "hexa", "rect"};
static const char *neigh_names[] = {NULL, "bubble", "gaussian"};

static int name_id(const char **names, int count, const char *s) {
  int i;
  if (s)
    for (i = 1; i < count; i++)
      if (strcasecmp(names[i], s) == 0) return i;
  return 0;
}

/* the n-th token (0-based) of the header line, split at blanks only (datafile.c:947-1023) */
static char *header_token(const char *line, int n, char *out, size_t outsz) {
  char *dup = strdup(line), *tok;
  int i;
  out[0] = '\0';
  if (!dup) return NULL;
  tok = strtok(dup, " ");
  for (i = 0; i < n && tok; i++) tok = strtok(NULL, " ");
  if (tok) { strncpy(out, tok, outsz - 1); out[outsz - 1] = '\0'; }
  free(dup);
  return out[0] ? out : NULL;
}

/* Decimal -> float.  scanf's %f is strtof, and glibc's strtof costs ~180 ns per value; almost every
 * token of a .dat file is a short decimal, for which the correctly rounded result is reachable
 * with one exact double operation (Clinger's fast path).  Towards float this needs one more
 * guard: mantissa < 2^53 and |exp10| <= 22 make mant * 10^e exact (checked against a table) or
 * mant / 10^e correct to half a double ulp; the second rounding double -> float is then the
 * correct rounding of the decimal unless the double lies within an ulp of a float rounding
 * midpoint (or is subnormal for float), and those tokens -- like anything unusual (inf, nan,
 * hex, > 18 digits) -- go to strtof.  tests/test_host_files.py compares against libc's strtof. */
static const double k_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
static float pak_strtof(const char *s, char **end) {
  const char *p = s;
  unsigned long long mant = 0;
  int neg = 0, nd = 0, dp = 0, seen = 0, e10 = 0;
  double q;
  unsigned long long bits, r;
  if (*p == '-') { neg = 1; p++; } else if (*p == '+') p++;
  if (p[0] == '0' && (p[1] == 'x' || p[1] == 'X')) goto slow;       /* hexadecimal floats */
  while (*p >= '0' && *p <= '9') { if (nd < 19) { mant = mant * 10 + (unsigned)(*p - '0'); if (mant) nd++; } else goto slow; seen = 1; p++; }
  if (*p == '.') {
    p++;
    while (*p >= '0' && *p <= '9') { if (nd < 19) { mant = mant * 10 + (unsigned)(*p - '0'); if (mant) nd++; dp++; } else goto slow; seen = 1; p++; }
  }
  if (!seen) goto slow;
  if (*p == 'e' || *p == 'E') {
    const char *pe = p + 1;
    int eneg = 0, ev = 0, ed = 0;
    if (*pe == '-') { eneg = 1; pe++; } else if (*pe == '+') pe++;
    while (*pe >= '0' && *pe <= '9') { if (ev < 10000) ev = ev * 10 + (*pe - '0'); ed = 1; pe++; }
    if (ed) { e10 = eneg ? -ev : ev; p = pe; }
  }
  e10 -= dp;
  if (mant == 0) { *end = (char *)p; return neg ? -0.0f : 0.0f; }
  if (mant >= (1ull << 53) || e10 > 22 || e10 < -22) goto slow;
  if (e10 >= 0) {
    q = (double)mant * k_pow10[e10];
    if (q >= 9007199254740992.0) goto slow;              /* product not exactly representable */
  } else {
    q = (double)mant / k_pow10[-e10];
  }
  if (q < 1.1754943508222875e-38 || q > 3.4028234e38) goto slow;      /* float subnormal / overflow range */
  memcpy(&bits, &q, sizeof bits);
  r = bits & ((1ull << 29) - 1);
  if (r >= (1ull << 28) - 2 && r <= (1ull << 28) + 2) goto slow;        /* next to a float rounding midpoint */
  *end = (char *)p;
  return neg ? -(float)q : (float)q;
slow:
  return strtof(s, end);
}

/* ---- loader.  The reference parses one line at a time with sscanf("%f") (datafile.c:552-748),
 * which at the 10 M-row scale of the batch search costs minutes while the search itself takes
 * milliseconds.  Here the file is read into memory once, cut at line
 * boundaries into blocks, and the blocks are parsed in parallel (pthreads; strtof converts
 * exactly like scanf's %f); the blocks are then stitched together in file order, and labels
 * are interned in file order so that the label table does not depend on the thread count. */
struct blk {
  char *beg, *end;                  /* [beg, end): whole lines */
  int dim, labels_needed, skip_empty;
  const char *mask_str, *name;
  long n, cap, nlab, labcap, lines;
  float *points;
  unsigned char *mask;              /* NULL until the block sees a masked component */
  long *lab_off;                    /* cap + 1 */
  char **labs;                      /* label tokens (pointers into the buffer) */
  short *weight, *fixed_xy;
  int err;                          /* 0 ok, 1 unreadable component, 2 out of memory, 3 bad fixed=, 4 label missing */
  int err_comp;
  long err_line;                    /* line inside the block */
};

static int blk_grow(struct blk *b) {
  const long ncap = b->cap ? b->cap * 2 : 1024;
  float *p = (float *)realloc(b->points, sizeof(float) * (size_t)ncap * b->dim);
  long *lo = (long *)realloc(b->lab_off, sizeof(long) * (size_t)(ncap + 1));
  short *w = (short *)realloc(b->weight, sizeof(short) * (size_t)ncap);
  short *f = (short *)realloc(b->fixed_xy, sizeof(short) * 2 * (size_t)ncap);
  if (p) b->points = p;
  if (lo) b->lab_off = lo;
  if (w) b->weight = w;
  if (f) b->fixed_xy = f;
  if (!p || !lo || !w || !f) return 1;
  if (b->mask) {
    unsigned char *m = (unsigned char *)realloc(b->mask, (size_t)ncap * b->dim);
    if (!m) return 1;
    memset(m + (size_t)b->cap * b->dim, 0, (size_t)(ncap - b->cap) * b->dim);
    b->mask = m;
  }
  if (b->cap == 0) b->lab_off[0] = 0;
  b->cap = ncap;
  return 0;
}

static void *blk_parse(void *arg) {
  struct blk *b = (struct blk *)arg;
  char *line = b->beg;
  const int dim = b->dim;
  {                                           /* one allocation per array: entries <= lines */
    long nlines = 1;
    const char *q = b->beg;
    while (q < b->end && (q = (const char *)memchr(q, '\n', (size_t)(b->end - q))) != NULL) { nlines++; q++; }
    b->cap = 0;
    while (b->cap < nlines) {
      const long want = nlines;
      b->points = (float *)malloc(sizeof(float) * (size_t)want * dim);
      b->lab_off = (long *)malloc(sizeof(long) * (size_t)(want + 1));
      b->weight = (short *)malloc(sizeof(short) * (size_t)want);
      b->fixed_xy = (short *)malloc(sizeof(short) * 2 * (size_t)want);
      if (!b->points || !b->lab_off || !b->weight || !b->fixed_xy) { b->err = 2; return NULL; }
      b->lab_off[0] = 0;
      b->cap = want;
    }
  }
  while (line < b->end) {
    char *nl = (char *)memchr(line, '\n', (size_t)(b->end - line));
    char *next = nl ? nl + 1 : b->end, *save = NULL, *tok;
    int i, maskcnt = 0, label_found = 0;
    float *pt;
    unsigned char *mk;
    if (nl) *nl = '\0';                       /* the caller guarantees a terminator after the last line */
    b->lines++;
    if (line[0] == '#') { line = next; continue; }
    tok = strtok_r(line, " \r\t", &save);
    if (!tok) { line = next; continue; }      /* empty line */
    if (b->n == b->cap && blk_grow(b)) { b->err = 2; return NULL; }
    pt = b->points + (size_t)b->n * dim;
    mk = b->mask ? b->mask + (size_t)b->n * dim : NULL;
    if (mk) memset(mk, 0, (size_t)dim);
    for (i = 0; i < dim; i++) {
      if (i > 0) tok = strtok_r(NULL, " \r\t", &save);
      if (tok && strcmp(tok, b->mask_str) == 0) {
        if (!b->mask) {                       /* first masked component of the block */
          b->mask = (unsigned char *)calloc((size_t)b->cap * dim, 1);
          if (!b->mask) { b->err = 2; return NULL; }
          mk = b->mask + (size_t)b->n * dim;
        }
        mk[i] = 1;
        maskcnt++;
        pt[i] = 0.0f;
      } else {
        char *endp = NULL;
        if (tok) pt[i] = pak_strtof(tok, &endp);  /* same value as sscanf("%f"), prefix match included */
        if (!tok || endp == tok) {
          b->err = 1;
          b->err_comp = i;
          b->err_line = b->lines;
          return NULL;
        }
      }
    }
    if (maskcnt == dim && b->skip_empty) { line = next; continue; }      /* datafile.c:677-690 */
    b->weight[b->n] = 0;
    b->fixed_xy[2 * b->n] = b->fixed_xy[2 * b->n + 1] = -1;
    while ((tok = strtok_r(NULL, " \r\t", &save)) != NULL) {
      if (strncmp(tok, "weight=", 7) == 0) {
        b->weight[b->n] = (short)atoi(tok + 7);
      } else if (strncmp(tok, "fixed=", 6) == 0) {
        const char *comma = strchr(tok, ',');
        if (!comma) {
          b->err = 3;
          b->err_line = b->lines;
          return NULL;
        }
        b->fixed_xy[2 * b->n] = (short)atoi(tok + 6);
        b->fixed_xy[2 * b->n + 1] = (short)atoi(comma + 1);
      } else {
        if (b->nlab == b->labcap) {
          const long ncap = b->labcap ? b->labcap * 2 : 1024;
          char **t = (char **)realloc(b->labs, sizeof(char *) * (size_t)ncap);
          if (!t) { b->err = 2; return NULL; }
          b->labs = t;
          b->labcap = ncap;
        }
        b->labs[b->nlab++] = tok;
        label_found++;
      }
    }
    if (b->labels_needed && !label_found) {
      b->err = 4;
      b->err_line = b->lines;
      return NULL;
    }
    b->n++;
    b->lab_off[b->n] = b->nlab;
    line = next;
  }
  return NULL;
}

static void blk_release(struct blk *b) {
  free(b->points); free(b->mask); free(b->lab_off); free(b->labs); free(b->weight); free(b->fixed_xy);
}

/* whole file (or stdin) into one buffer with a terminating NUL */
static char *slurp(FILE *fp, size_t *len) {
  size_t cap = 1 << 20, n = 0, got;
  char *buf = (char *)malloc(cap + 1);
  if (!buf) return NULL;
  while ((got = fread(buf + n, 1, cap - n, fp)) > 0) {
    n += got;
    if (n == cap) {
      char *t = (char *)realloc(buf, cap * 2 + 1);
      if (!t) { free(buf); return NULL; }
      buf = t;
      cap *= 2;
    }
  }
  buf[n] = '\0';
  *len = n;
  return buf;
}

/* open_file (fileio.c:60-190): "-" is stdin / stdout, a name ending in .gz / .z / .Z goes through gzip
 * ("gzip -d -c %s" / "gzip -9 -c >%s", fileio.h:33-38), a name starting with '|' is a command. */
static FILE *pak_open(const char *name, int writing, int *piped) {
  char cmd[4200];
  const char *dot;
  *piped = 0;
  if (strcmp(name, "-") == 0) return writing ? stdout : stdin;
  if (name[0] == '|') {
    *piped = 1;
    return popen(name + 1, writing ? "w" : "r");
  }
  dot = strrchr(name, '.');
  if (dot && (strcmp(dot, ".gz") == 0 || strcmp(dot, ".z") == 0 || strcmp(dot, ".Z") == 0)) {
    /* the reference interpolates the name as it is ("gzip -d -c %s", fileio.h:33-38); here it is single-quoted
     * so that a file name cannot carry shell syntax ('|cmd' names are commands by definition) */
    char quoted[4100];
    size_t q = 0;
    const char *c;
    if (strlen(name) > 1000) return NULL;
    quoted[q++] = '\'';
    for (c = name; *c; c++) {
      if (*c == '\'') { memcpy(quoted + q, "'\\''", 4); q += 4; }
      else quoted[q++] = *c;
    }
    quoted[q++] = '\'';
    quoted[q] = '\0';
    snprintf(cmd, sizeof cmd, writing ? "gzip -9 -c >%s" : "gzip -d -c %s", quoted);
    *piped = 1;
    return popen(cmd, writing ? "w" : "r");
  }
  return fopen(name, writing ? "w" : "r");
}
static void pak_close(FILE *fp, int piped) {
  if (fp == stdin || fp == stdout) return;
  if (piped) pclose(fp); else fclose(fp);
}

/* Parses the text in buf[0, len) (NUL-terminated, whole lines; consumed: the tokenizer writes into it).
 * head == NULL: the text starts with the file header; otherwise it is a later part of the file whose
 * header was `head`.  *lines_io: lines of the file in front of this text (for error messages), advanced. */
static struct pak_entries *parse_text(char *buf, size_t len, const char *name, int labels_needed, int skip_empty,
                                      const struct pak_entries *head, long *lines_io) {
  char *p, *endbuf, tokbuf[64];
  long header_lines = lines_io ? *lines_io : 0, total = 0, totlab = 0, i;
  struct pak_entries *e = NULL;
  struct blk *blks = NULL;
  pthread_t *tids = NULL;
  int dim, nthreads, nb = 0, t, any_mask = 0, failed = 0;
  const char *env = getenv("BMU_PAK_THREADS");
  endbuf = buf + len;
  /* header: first line that is not a comment (datafile.c:112-148) */
  p = buf;
  if (head) {
    e = (struct pak_entries *)calloc(1, sizeof(*e));
    if (!e) return NULL;
    dim = head->dim;
    e->dim = head->dim; e->topol = head->topol; e->neigh = head->neigh; e->xdim = head->xdim; e->ydim = head->ydim;
  } else
  for (;;) {
    char *nl;
    if (p >= endbuf) { fprintf(stderr, "Can't read file %s", name); return NULL; }
    nl = (char *)memchr(p, '\n', (size_t)(endbuf - p));
    if (nl) *nl = '\0';
    header_lines++;
    if (p[0] != '#') {
      char *line = p;
      p = nl ? nl + 1 : endbuf;
      if (sscanf(line, "%d", &dim) <= 0 || dim <= 0) {
        fprintf(stderr, "Can't read dimension parameter in file %s", name);
        return NULL;
      }
      e = (struct pak_entries *)calloc(1, sizeof(*e));
      if (!e) return NULL;
      e->dim = dim;
      e->topol = name_id(topol_names, 5, header_token(line, 1, tokbuf, sizeof tokbuf));
      e->xdim = header_token(line, 2, tokbuf, sizeof tokbuf) ? atoi(tokbuf) : 0;
      e->ydim = header_token(line, 3, tokbuf, sizeof tokbuf) ? atoi(tokbuf) : 0;
      e->neigh = name_id(neigh_names, 3, header_token(line, 4, tokbuf, sizeof tokbuf));
      break;
    }
    p = nl ? nl + 1 : endbuf;
  }
  /* blocks of whole lines, at least 1 MiB each */
  nthreads = env ? atoi(env) : (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  if ((size_t)(endbuf - p) / nthreads < (1u << 20)) nthreads = (int)((size_t)(endbuf - p) >> 20) + 1;
  blks = (struct blk *)calloc((size_t)nthreads, sizeof(*blks));
  tids = (pthread_t *)calloc((size_t)nthreads, sizeof(*tids));
  if (!blks || !tids) goto fail;
  {
    char *start = p;
    const size_t chunk = (size_t)(endbuf - p) / nthreads + 1;
    while (start < endbuf) {
      char *stop = start + chunk < endbuf ? start + chunk : endbuf;
      if (stop < endbuf) {                                 /* extend to the end of the line */
        char *nl = (char *)memchr(stop, '\n', (size_t)(endbuf - stop));
        stop = nl ? nl + 1 : endbuf;
      }
      blks[nb].beg = start; blks[nb].end = stop;
      blks[nb].dim = dim; blks[nb].labels_needed = labels_needed; blks[nb].skip_empty = skip_empty;
      blks[nb].mask_str = pak_mask_string; blks[nb].name = name;
      nb++;
      start = stop;
    }
  }
  for (t = 1; t < nb; t++)
    if (pthread_create(&tids[t], NULL, blk_parse, &blks[t])) { blk_parse(&blks[t]); tids[t] = 0; }
  if (nb > 0) blk_parse(&blks[0]);
  for (t = 1; t < nb; t++)
    if (tids[t]) pthread_join(tids[t], NULL);
  /* first error in file order, with its line number in the file */
  {
    long lines_before = header_lines;
    for (t = 0; t < nb; t++) {
      const long ln = lines_before + blks[t].err_line;
      if (blks[t].err == 1)
        fprintf(stderr, "load_entry: can't read entry in file %s on line %ld, component %d\n", name, ln, blks[t].err_comp);
      if (blks[t].err == 2) fprintf(stderr, "out of memory while reading %s\n", name);
      if (blks[t].err == 3) fprintf(stderr, "bad fixed point, line %ld of file %s\n", ln, name);
      if (blks[t].err == 4) fprintf(stderr, "Required label missing on line %ld of file %s\n", ln, name);
      if (blks[t].err) { failed = 1; break; }
      lines_before += blks[t].lines;
    }
  }
  if (failed) goto fail;
  /* stitch the blocks together in file order */
  for (t = 0; t < nb; t++) { total += blks[t].n; totlab += blks[t].nlab; if (blks[t].mask) any_mask = 1; }
  e->n = total;
  e->points = (float *)malloc(sizeof(float) * (size_t)(total > 0 ? total : 1) * dim);
  e->lab_off = (long *)calloc((size_t)total + 1, sizeof(long));
  e->lab_pool = (int *)malloc(sizeof(int) * (size_t)(totlab > 0 ? totlab : 1));
  e->weight = (short *)calloc((size_t)(total > 0 ? total : 1), sizeof(short));
  e->fixed_xy = (short *)malloc(sizeof(short) * 2 * (size_t)(total > 0 ? total : 1));
  if (any_mask) e->mask = (unsigned char *)calloc((size_t)(total > 0 ? total : 1) * dim, 1);
  if (!e->points || !e->lab_off || !e->lab_pool || !e->weight || !e->fixed_xy || (any_mask && !e->mask)) goto fail;
  memset(e->fixed_xy, 0xff, sizeof(short) * 2 * (size_t)(total > 0 ? total : 1));
  {
    long row = 0, lab = 0;
    for (t = 0; t < nb; t++) {
      struct blk *b = &blks[t];
      if (b->n == 0) continue;
      memcpy(e->points + (size_t)row * dim, b->points, sizeof(float) * (size_t)b->n * dim);
      if (b->mask) memcpy(e->mask + (size_t)row * dim, b->mask, (size_t)b->n * dim);
      memcpy(e->weight + row, b->weight, sizeof(short) * (size_t)b->n);
      memcpy(e->fixed_xy + 2 * row, b->fixed_xy, sizeof(short) * 2 * (size_t)b->n);
      for (i = 0; i < b->n; i++) {
        long l;
        for (l = b->lab_off[i]; l < b->lab_off[i + 1]; l++) {
          const int id = label_index(b->labs[l]);          /* file order: same table for any thread count */
          if (id != LABEL_EMPTY) e->lab_pool[lab++] = id;
        }
        e->lab_off[row + i + 1] = lab;
      }
      row += b->n;
    }
  }
  if (lines_io) {
    *lines_io = header_lines;
    for (t = 0; t < nb; t++) *lines_io += blks[t].lines;
  }
  for (t = 0; t < nb; t++) blk_release(&blks[t]);
  free(blks); free(tids);
  return e;
fail:
  if (blks) for (t = 0; t < nb; t++) blk_release(&blks[t]);
  free(blks); free(tids);
  pak_free(e);
  return NULL;
}

/* ------------------------------------------------------------------ binary cache of a parsed file (opt-in)
 * $BMU_PAK_CACHE=1 keeps `<file>.bmuc` next to a regular .dat / .cod file after its first parse: the flat arrays
 * as they are in memory plus the label STRINGS in the order they first appear, so that loading the cache interns
 * them exactly as the parse would have.  The cache is used only while the source's size and modification time,
 * the mask string and the two loader flags are the ones it was made with.  ($BMU_PAK_CACHE=<directory> puts the
 * caches there instead.)  A 577 MB file of 1 M x 64 values loads in ~0.1 s instead of ~1.3 s. */
#include <sys/stat.h>
struct cache_head {
  char magic[8];                       /* "BMUC0001" */
  long long src_size, src_mtime_s, src_mtime_ns;
  int labels_needed, skip_empty, dim, topol, neigh, xdim, ydim, has_mask;
  long long n, nlab, nstrings, string_bytes;
  char mask_str[32];
};

static int cache_path(const char *name, char *out, size_t outsz) {
  const char *env = getenv("BMU_PAK_CACHE");
  const char *dot = strrchr(name, '.');
  if (!env || !env[0] || strcmp(env, "0") == 0) return 0;
  if (strcmp(name, "-") == 0 || name[0] == '|') return 0;
  if (dot && (strcmp(dot, ".gz") == 0 || strcmp(dot, ".z") == 0 || strcmp(dot, ".Z") == 0)) return 0;
  if (strcmp(env, "1") == 0) return snprintf(out, outsz, "%s.bmuc", name) < (int)outsz;
  {
    const char *base = strrchr(name, '/');
    return snprintf(out, outsz, "%s/%s.bmuc", env, base ? base + 1 : name) < (int)outsz;
  }
}

static void cache_fill_head(struct cache_head *h, const struct stat *sb, int labels_needed, int skip_empty) {
  memset(h, 0, sizeof(*h));
  memcpy(h->magic, "BMUC0001", 8);
  h->src_size = (long long)sb->st_size;
  h->src_mtime_s = (long long)sb->st_mtim.tv_sec;
  h->src_mtime_ns = (long long)sb->st_mtim.tv_nsec;
  h->labels_needed = labels_needed;
  h->skip_empty = skip_empty;
  strncpy(h->mask_str, pak_mask_string, sizeof(h->mask_str) - 1);
}

static struct pak_entries *cache_load(const char *name, const char *cpath, int labels_needed, int skip_empty) {
  struct stat sb;
  struct cache_head h, want;
  struct pak_entries *e = NULL;
  char *strings = NULL;
  int *map = NULL;
  long long i;
  FILE *fp;
  if (stat(name, &sb) != 0 || !S_ISREG(sb.st_mode)) return NULL;
  fp = fopen(cpath, "rb");
  if (!fp) return NULL;
  cache_fill_head(&want, &sb, labels_needed, skip_empty);
  if (fread(&h, sizeof(h), 1, fp) != 1 || memcmp(h.magic, want.magic, 8) != 0 || h.src_size != want.src_size ||
      h.src_mtime_s != want.src_mtime_s || h.src_mtime_ns != want.src_mtime_ns || h.labels_needed != labels_needed ||
      h.skip_empty != skip_empty || strcmp(h.mask_str, want.mask_str) != 0 || h.n < 0 || h.dim < 1)
    goto bad;
  e = (struct pak_entries *)calloc(1, sizeof(*e));
  if (!e) goto bad;
  e->dim = h.dim; e->topol = h.topol; e->neigh = h.neigh; e->xdim = h.xdim; e->ydim = h.ydim; e->n = (long)h.n;
  {
    const size_t n1 = (size_t)(h.n > 0 ? h.n : 1);
    e->points = (float *)malloc(sizeof(float) * n1 * h.dim);
    e->lab_off = (long *)malloc(sizeof(long) * ((size_t)h.n + 1));
    e->lab_pool = (int *)malloc(sizeof(int) * (size_t)(h.nlab > 0 ? h.nlab : 1));
    e->weight = (short *)malloc(sizeof(short) * n1);
    e->fixed_xy = (short *)malloc(sizeof(short) * 2 * n1);
    if (h.has_mask) e->mask = (unsigned char *)malloc(n1 * h.dim);
    strings = (char *)malloc((size_t)(h.string_bytes > 0 ? h.string_bytes : 1));
    map = (int *)malloc(sizeof(int) * (size_t)(h.nstrings > 0 ? h.nstrings : 1));
    if (!e->points || !e->lab_off || !e->lab_pool || !e->weight || !e->fixed_xy || (h.has_mask && !e->mask) || !strings || !map)
      goto bad;
    if (fread(e->points, sizeof(float) * h.dim, (size_t)h.n, fp) != (size_t)h.n) goto bad;
    if (h.has_mask && fread(e->mask, (size_t)h.dim, (size_t)h.n, fp) != (size_t)h.n) goto bad;
    if (fread(e->lab_off, sizeof(long), (size_t)h.n + 1, fp) != (size_t)h.n + 1) goto bad;
    if (h.nlab && fread(e->lab_pool, sizeof(int), (size_t)h.nlab, fp) != (size_t)h.nlab) goto bad;
    if (fread(e->weight, sizeof(short), (size_t)h.n, fp) != (size_t)h.n) goto bad;
    if (fread(e->fixed_xy, sizeof(short) * 2, (size_t)h.n, fp) != (size_t)h.n) goto bad;
    if (h.string_bytes && fread(strings, 1, (size_t)h.string_bytes, fp) != (size_t)h.string_bytes) goto bad;
  }
  /* intern the label strings in first-appearance order, then map the cache's local ids to table ids */
  {
    const char *p = strings, *end = strings + h.string_bytes;
    for (i = 0; i < h.nstrings; i++) {
      if (p >= end) goto bad;
      map[i] = label_index(p);
      p += strlen(p) + 1;
    }
    for (i = 0; i < h.nlab; i++) {
      if (e->lab_pool[i] < 0 || e->lab_pool[i] >= h.nstrings) goto bad;
      e->lab_pool[i] = map[e->lab_pool[i]];
    }
  }
  fclose(fp);
  free(strings); free(map);
  return e;
bad:
  fclose(fp);
  free(strings); free(map);
  pak_free(e);
  return NULL;
}

static void cache_save(const char *name, const char *cpath, const struct pak_entries *e, int labels_needed, int skip_empty) {
  struct stat sb;
  struct cache_head h;
  char tmp[4200];
  FILE *fp;
  int *local = NULL, *ids = NULL, nids = 0, idcap = 0, ok = 1;
  long i;
  size_t bytes = 0;
  const long nlab = e->lab_off[e->n];
  if (stat(name, &sb) != 0 || !S_ISREG(sb.st_mode)) return;
  if (snprintf(tmp, sizeof tmp, "%s.tmp%d", cpath, (int)getpid()) >= (int)sizeof tmp) return;
  /* local ids in first-appearance order */
  local = (int *)malloc(sizeof(int) * (size_t)(nlab > 0 ? nlab : 1));
  if (!local) return;
  for (i = 0; i < nlab; i++) {
    int j, id = e->lab_pool[i];
    for (j = 0; j < nids; j++)
      if (ids[j] == id) break;
    if (j == nids) {
      if (nids >= 4096) { free(local); free(ids); return; }      /* files with that many distinct labels are not cached */
      if (nids == idcap) {
        int *t = (int *)realloc(ids, sizeof(int) * (size_t)(idcap + 256));
        if (!t) { free(local); free(ids); return; }
        ids = t;
        idcap += 256;
      }
      ids[nids++] = id;
    }
    local[i] = j;
  }
  for (i = 0; i < nids; i++) bytes += strlen(label_string(ids[i])) + 1;
  cache_fill_head(&h, &sb, labels_needed, skip_empty);
  h.dim = e->dim; h.topol = e->topol; h.neigh = e->neigh; h.xdim = e->xdim; h.ydim = e->ydim; h.has_mask = e->mask != NULL;
  h.n = e->n; h.nlab = nlab; h.nstrings = nids; h.string_bytes = (long long)bytes;
  fp = fopen(tmp, "wb");
  if (!fp) { free(local); free(ids); return; }
  ok = fwrite(&h, sizeof(h), 1, fp) == 1;
  ok = ok && fwrite(e->points, sizeof(float) * e->dim, (size_t)e->n, fp) == (size_t)e->n;
  if (e->mask) ok = ok && fwrite(e->mask, (size_t)e->dim, (size_t)e->n, fp) == (size_t)e->n;
  ok = ok && fwrite(e->lab_off, sizeof(long), (size_t)e->n + 1, fp) == (size_t)e->n + 1;
  if (nlab) ok = ok && fwrite(local, sizeof(int), (size_t)nlab, fp) == (size_t)nlab;
  ok = ok && fwrite(e->weight, sizeof(short), (size_t)e->n, fp) == (size_t)e->n;
  ok = ok && fwrite(e->fixed_xy, sizeof(short) * 2, (size_t)e->n, fp) == (size_t)e->n;
  for (i = 0; i < nids && ok; i++) {
    const char *str = label_string(ids[i]);
    ok = fwrite(str, 1, strlen(str) + 1, fp) == strlen(str) + 1;
  }
  ok = (fclose(fp) == 0) && ok;
  if (ok) ok = rename(tmp, cpath) == 0;         /* atomic: a reader never sees half a cache */
  if (!ok) remove(tmp);
  free(local); free(ids);
}

struct pak_entries *pak_load(const char *name, int labels_needed, int skip_empty) {
  int piped = 0;
  FILE *fp;
  struct pak_entries *e;
  char *buf, cpath[4200];
  size_t len = 0;
  const int cached = cache_path(name, cpath, sizeof cpath);
  if (cached && (e = cache_load(name, cpath, labels_needed, skip_empty)) != NULL) return e;
  fp = pak_open(name, 0, &piped);
  if (!fp) return NULL;
  buf = slurp(fp, &len);
  pak_close(fp, piped);
  if (!buf) { fprintf(stderr, "Can't read file %s", name); return NULL; }
  e = parse_text(buf, len, name, labels_needed, skip_empty, NULL, NULL);
  free(buf);
  if (e && cached) cache_save(name, cpath, e, labels_needed, skip_empty);
  return e;
}

/* ------------------------------------------------------------------ streamed reading (-buffer N)
 * The reference's `-buffer N` (datafile.c:237-344) keeps N entries in memory at a time.  Here the file is
 * read in text blocks of PAK_TEXT_BLOCK bytes, each parsed by the block-parallel parser above, and handed
 * out in chunks of exactly N entries; the NEXT text block is read and parsed on a helper thread while
 * the caller searches the current chunk (parse || H2D || kernel). */
#define PAK_TEXT_BLOCK pak_text_block()
static size_t pak_text_block(void) {       /* $BMU_PAK_TEXT_BLOCK: bytes per text block (tests use small ones) */
  const char *env = getenv("BMU_PAK_TEXT_BLOCK");
  return (env && atol(env) > 0) ? (size_t)atol(env) : ((size_t)64 << 20);
}
struct pak_stream {
  FILE *fp;
  int piped, labels_needed, skip_empty, eof, failed;
  char *name;
  long buffer, lines;
  struct pak_entries head;          /* dim / topol / ... of the file (no rows) */
  int have_head;
  char *carry;                      /* the unfinished last line of the block before */
  size_t carry_len;
  struct pak_entries *pending;      /* parsed, not handed out yet */
  long pending_pos;
  pthread_t th;
  int th_running;
  struct pak_entries *fetched;      /* result of the helper thread */
  int fetched_eof, fetched_failed;
};

/* read and parse the next text block; returns NULL at the end of the file or on error (st->failed) */
static struct pak_entries *stream_block(struct pak_stream *st, int *eof, int *failed) {
  const size_t block = PAK_TEXT_BLOCK;
  size_t cap = block + st->carry_len + 1, n = st->carry_len, got, cut;
  char *buf = (char *)malloc(cap + 1);
  struct pak_entries *e;
  *failed = 0;
  if (!buf) { *failed = 1; return NULL; }
  if (st->carry_len) memcpy(buf, st->carry, st->carry_len);
  free(st->carry);
  st->carry = NULL;
  st->carry_len = 0;
  got = fread(buf + n, 1, block, st->fp);
  n += got;
  if (got < block) *eof = 1;
  if (n == 0) { free(buf); return NULL; }
  cut = n;
  if (!*eof) {                                         /* keep the unfinished last line for the next block */
    while (cut > 0 && buf[cut - 1] != '\n') cut--;
    if (cut == 0) {                                    /* one line longer than a block: read on */
      st->carry = buf;
      st->carry_len = n;
      return stream_block(st, eof, failed);
    }
    st->carry_len = n - cut;
    st->carry = (char *)malloc(st->carry_len + 1);
    if (!st->carry) { free(buf); *failed = 1; return NULL; }
    memcpy(st->carry, buf + cut, st->carry_len);
  }
  buf[cut] = '\0';
  e = parse_text(buf, cut, st->name, st->labels_needed, st->skip_empty, st->have_head ? &st->head : NULL, &st->lines);
  free(buf);
  if (!e) { *failed = 1; return NULL; }
  if (!st->have_head) {
    st->head = *e;
    st->head.n = 0;
    st->head.points = NULL; st->head.mask = NULL; st->head.lab_off = NULL; st->head.lab_pool = NULL;
    st->head.weight = NULL; st->head.fixed_xy = NULL;
    st->have_head = 1;
  }
  return e;
}

static void *stream_fetch(void *arg) {
  struct pak_stream *st = (struct pak_stream *)arg;
  st->fetched_eof = 0;
  st->fetched = stream_block(st, &st->fetched_eof, &st->fetched_failed);
  return NULL;
}

struct pak_stream *pak_stream_open(const char *name, int labels_needed, int skip_empty, long buffer) {
  struct pak_stream *st = (struct pak_stream *)calloc(1, sizeof(*st));
  if (!st) return NULL;
  st->fp = pak_open(name, 0, &st->piped);
  if (!st->fp) { free(st); return NULL; }
  st->name = strdup(name);
  st->labels_needed = labels_needed;
  st->skip_empty = skip_empty;
  st->buffer = buffer;
  /* the first block is parsed here: it carries the header, which the caller wants before the first chunk */
  st->pending = stream_block(st, &st->eof, &st->failed);
  if (st->failed || !st->have_head) { pak_stream_close(st); return NULL; }
  return st;
}

const struct pak_entries *pak_stream_header(const struct pak_stream *st) { return &st->head; }

/* rows [from, from + n) of `src` appended to `dst` (same file, same dim) */
static int entries_append(struct pak_entries *dst, const struct pak_entries *src, long from, long n) {
  const int dim = dst->dim;
  const long nn = dst->n + n, lab0 = dst->lab_off ? dst->lab_off[dst->n] : 0;
  const long nl = src->lab_off[from + n] - src->lab_off[from];
  long i;
  float *p = (float *)realloc(dst->points, sizeof(float) * (size_t)(nn > 0 ? nn : 1) * dim);
  long *lo = (long *)realloc(dst->lab_off, sizeof(long) * (size_t)(nn + 1));
  int *lp = (int *)realloc(dst->lab_pool, sizeof(int) * (size_t)(lab0 + nl > 0 ? lab0 + nl : 1));
  short *w = (short *)realloc(dst->weight, sizeof(short) * (size_t)(nn > 0 ? nn : 1));
  short *f = (short *)realloc(dst->fixed_xy, sizeof(short) * 2 * (size_t)(nn > 0 ? nn : 1));
  if (p) dst->points = p;
  if (lo) dst->lab_off = lo;
  if (lp) dst->lab_pool = lp;
  if (w) dst->weight = w;
  if (f) dst->fixed_xy = f;
  if (!p || !lo || !lp || !w || !f) return 1;
  if (dst->n == 0) dst->lab_off[0] = 0;
  if (src->mask || dst->mask) {
    unsigned char *m = (unsigned char *)realloc(dst->mask, (size_t)(nn > 0 ? nn : 1) * dim);
    if (!m) return 1;
    if (!dst->mask) memset(m, 0, (size_t)dst->n * dim);
    dst->mask = m;
    if (src->mask) memcpy(m + (size_t)dst->n * dim, src->mask + (size_t)from * dim, (size_t)n * dim);
    else memset(m + (size_t)dst->n * dim, 0, (size_t)n * dim);
  }
  memcpy(dst->points + (size_t)dst->n * dim, src->points + (size_t)from * dim, sizeof(float) * (size_t)n * dim);
  memcpy(dst->weight + dst->n, src->weight + from, sizeof(short) * (size_t)n);
  memcpy(dst->fixed_xy + 2 * dst->n, src->fixed_xy + 2 * from, sizeof(short) * 2 * (size_t)n);
  memcpy(dst->lab_pool + lab0, src->lab_pool + src->lab_off[from], sizeof(int) * (size_t)nl);
  for (i = 1; i <= n; i++) dst->lab_off[dst->n + i] = lab0 + (src->lab_off[from + i] - src->lab_off[from]);
  dst->n = nn;
  return 0;
}

/* the next chunk: `buffer` entries (fewer at the end of the file; the whole file when buffer <= 0);
 * NULL when the file is exhausted or on error (pak_stream_failed tells which).  The caller frees it. */
struct pak_entries *pak_stream_next(struct pak_stream *st) {
  struct pak_entries *out;
  if (st->failed) return NULL;
  out = (struct pak_entries *)calloc(1, sizeof(*out));
  if (!out) { st->failed = 1; return NULL; }
  *out = st->head;
  for (;;) {
    const long want = st->buffer > 0 ? st->buffer - out->n : -1;
    if (st->pending) {
      const long have = st->pending->n - st->pending_pos, take = (want < 0 || have < want) ? have : want;
      if (take > 0 && entries_append(out, st->pending, st->pending_pos, take)) { st->failed = 1; break; }
      st->pending_pos += take;
      if (st->pending_pos >= st->pending->n) { pak_free(st->pending); st->pending = NULL; st->pending_pos = 0; }
    }
    if (st->buffer > 0 && out->n >= st->buffer) break;
    if (st->pending) continue;
    /* need more text: take what the helper thread fetched, or read now */
    if (st->th_running) {
      pthread_join(st->th, NULL);
      st->th_running = 0;
      st->pending = st->fetched;
      st->fetched = NULL;
      st->eof = st->fetched_eof;
      if (st->fetched_failed) { st->failed = 1; break; }
    } else if (!st->eof) {
      st->pending = stream_block(st, &st->eof, &st->failed);
      if (st->failed) break;
    }
    if (!st->pending && st->eof) break;
  }
  /* read ahead while the caller works on this chunk */
  if (!st->failed && !st->eof && !st->th_running && (!st->pending || st->pending->n - st->pending_pos < st->buffer || st->buffer <= 0)) {
    if (pthread_create(&st->th, NULL, stream_fetch, st) == 0) st->th_running = 1;
  }
  if (st->failed || out->n == 0) { pak_free(out); return NULL; }
  return out;
}

int pak_stream_failed(const struct pak_stream *st) { return st->failed; }

void pak_stream_close(struct pak_stream *st) {
  if (!st) return;
  if (st->th_running) { pthread_join(st->th, NULL); pak_free(st->fetched); }
  if (st->fp) pak_close(st->fp, st->piped);
  pak_free(st->pending);
  free(st->carry);
  free(st->name);
  free(st);
}



void pak_write_header(FILE *fp, const struct pak_entries *e) {     /* datafile.c:396-415 */
  fprintf(fp, "%d", e->dim);
  if (e->topol > TOPOL_DATA) {
    fprintf(fp, " %s", topol_names[e->topol]);
    if (e->topol > TOPOL_LVQ) fprintf(fp, " %d %d %s", e->xdim, e->ydim, neigh_names[e->neigh] ? neigh_names[e->neigh] : "(null)");
  }
  fprintf(fp, "\n");
}

void pak_write_entries(FILE *fp, const struct pak_entries *e) {     /* datafile.c:420-447 */
  long i, l;
  int c;
  for (i = 0; i < e->n; i++) {
    const float *pt = e->points + (size_t)i * e->dim;
    const unsigned char *mk = e->mask ? e->mask + (size_t)i * e->dim : NULL;
    for (c = 0; c < e->dim; c++) {
      if (mk && mk[c]) fprintf(fp, "%s ", pak_mask_string);
      else fprintf(fp, "%g ", pt[c]);
    }
    for (l = e->lab_off[i]; l < e->lab_off[i + 1]; l++) fprintf(fp, "%s ", label_string(e->lab_pool[l]));
    fprintf(fp, "\n");
  }
}

int pak_save(const struct pak_entries *e, const char *name) {
  int piped = 0;
  FILE *fp = pak_open(name, 1, &piped);
  if (!fp) { fprintf(stderr, "save_entries: Can't open file '%s'\n", name); return 1; }
  pak_write_header(fp, e);
  pak_write_entries(fp, e);
  pak_close(fp, piped);
  return 0;
}

/* ------------------------------------------------------------------ hitlists */
void hit_init(struct pak_hitlist *h) { h->n = h->cap = 0; h->label = h->freq = NULL; }
void hit_clear(struct pak_hitlist *h) { h->n = 0; }
void hit_free(struct pak_hitlist *h) { free(h->label); free(h->freq); hit_init(h); }

long hit_add(struct pak_hitlist *h, long label) {                  /* labels.c:370-410 */
  long i;
  for (i = 0; i < h->n; i++)
    if (h->label[i] == label) break;
  if (i < h->n) {
    h->freq[i]++;
    /* move towards the head while the entry in front has a strictly smaller count */
    while (i > 0 && h->freq[i - 1] < h->freq[i]) {
      long tl = h->label[i - 1], tf = h->freq[i - 1];
      h->label[i - 1] = h->label[i]; h->freq[i - 1] = h->freq[i];
      h->label[i] = tl; h->freq[i] = tf;
      i--;
    }
    return h->freq[i];
  }
  if (h->n == h->cap) {
    long ncap = h->cap ? h->cap * 2 : 8;
    long *l = (long *)realloc(h->label, sizeof(long) * (size_t)ncap);
    long *f = (long *)realloc(h->freq, sizeof(long) * (size_t)ncap);
    if (l) h->label = l;
    if (f) h->freq = f;
    if (!l || !f) return 0;
    h->cap = ncap;
  }
  h->label[h->n] = label;
  h->freq[h->n] = 1;
  h->n++;
  return 1;
}

long hit_freq(const struct pak_hitlist *h, long label) {
  long i;
  for (i = 0; i < h->n; i++)
    if (h->label[i] == label) return h->freq[i];
  return 0;
}
