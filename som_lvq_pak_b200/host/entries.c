/* entries.c -- label table, .dat/.cod reader and writer, hitlists (host layer, see somhost.h).
 *
 * Behaviour follows the reference's file layer (datafile.c, labels.c, fileio.c) so that files
 * written here are byte-identical to the reference's; the storage is flat arrays, not lists.
 */
#include "somhost.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ label table */
static char **g_labels = NULL;
static int g_nlabels = 0, g_labcap = 0;

int label_index(const char *str) {
  int i;
  if (str == NULL || str[0] == '\0') return LABEL_EMPTY;          /* labels.c:86-91 */
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

const char *label_string(int ind) {
  if (ind <= 0 || ind > g_nlabels) return NULL;
  return g_labels[ind - 1];
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

/* whole line without the newline, any length (fileio.c:283-375); NULL at end of file */
static char *read_line(FILE *fp, char **buf, size_t *cap) {
  size_t len = 0;
  int c;
  if (!*buf) { *cap = 4096; *buf = (char *)malloc(*cap); if (!*buf) return NULL; }
  while ((c = fgetc(fp)) != EOF && c != '\n') {
    if (len + 2 > *cap) {
      char *t = (char *)realloc(*buf, *cap * 2);
      if (!t) return NULL;
      *buf = t;
      *cap *= 2;
    }
    (*buf)[len++] = (char)c;
  }
  (*buf)[len] = '\0';
  if (c == EOF && len == 0) return NULL;
  return *buf;
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

struct grow {       /* growing arrays while the number of entries is unknown */
  long cap, labcap, nlab;
};

static int grow_entries(struct pak_entries *e, struct grow *g, int have_mask) {
  long ncap = g->cap ? g->cap * 2 : 1024;
  float *p = (float *)realloc(e->points, sizeof(float) * (size_t)ncap * e->dim);
  long *lo;
  short *w, *f;
  if (!p) return 1;
  e->points = p;
  if (have_mask) {
    unsigned char *m = (unsigned char *)realloc(e->mask, (size_t)ncap * e->dim);
    if (!m) return 1;
    memset(m + (size_t)g->cap * e->dim, 0, (size_t)(ncap - g->cap) * e->dim);
    e->mask = m;
  }
  lo = (long *)realloc(e->lab_off, sizeof(long) * (size_t)(ncap + 1));
  w = (short *)realloc(e->weight, sizeof(short) * (size_t)ncap);
  f = (short *)realloc(e->fixed_xy, sizeof(short) * 2 * (size_t)ncap);
  if (lo) e->lab_off = lo;
  if (w) e->weight = w;
  if (f) e->fixed_xy = f;
  if (!lo || !w || !f) return 1;
  g->cap = ncap;
  return 0;
}

struct pak_entries *pak_load(const char *name, int labels_needed, int skip_empty) {
  FILE *fp = strcmp(name, "-") == 0 ? stdin : fopen(name, "r");
  char *buf = NULL, *line, tokbuf[64];
  size_t cap = 0;
  long row = 0;
  struct pak_entries *e = NULL;
  struct grow g = {0, 0, 0};
  int dim;
  if (!fp) return NULL;
  /* header: first line that is not a comment (datafile.c:112-148) */
  do {
    line = read_line(fp, &buf, &cap);
    row++;
    if (!line) { fprintf(stderr, "Can't read file %s", name); goto fail; }
  } while (line[0] == '#');
  if (sscanf(line, "%d", &dim) <= 0 || dim <= 0) {
    fprintf(stderr, "Can't read dimension parameter in file %s", name);
    goto fail;
  }
  e = (struct pak_entries *)calloc(1, sizeof(*e));
  if (!e) goto fail;
  e->dim = dim;
  e->topol = name_id(topol_names, 5, header_token(line, 1, tokbuf, sizeof tokbuf));
  e->xdim = header_token(line, 2, tokbuf, sizeof tokbuf) ? atoi(tokbuf) : 0;
  e->ydim = header_token(line, 3, tokbuf, sizeof tokbuf) ? atoi(tokbuf) : 0;
  e->neigh = name_id(neigh_names, 3, header_token(line, 4, tokbuf, sizeof tokbuf));
  e->lab_off = (long *)calloc(1, sizeof(long));
  if (!e->lab_off) goto fail;

  /* entries (datafile.c:552-748) */
  while ((line = read_line(fp, &buf, &cap)) != NULL) {
    char *tok;
    int i, maskcnt = 0, label_found = 0;
    float *pt;
    unsigned char *mk;
    long first_lab;
    row++;
    if (line[0] == '#') continue;
    tok = strtok(line, " \r\t");
    if (!tok) continue;                                           /* empty line */
    if (e->n == g.cap && grow_entries(e, &g, e->mask != NULL)) goto fail;
    pt = e->points + (size_t)e->n * dim;
    mk = e->mask ? e->mask + (size_t)e->n * dim : NULL;
    if (mk) memset(mk, 0, (size_t)dim);
    for (i = 0; i < dim; i++) {
      if (i > 0) tok = strtok(NULL, " \r\t");
      if (!tok) {
        fprintf(stderr, "load_entry: can't read entry in file %s on line %ld, component %d\n", name, row, i);
        goto fail;
      }
      if (strcmp(tok, pak_mask_string) == 0) {
        if (!e->mask) {                                           /* first masked component of the file */
          e->mask = (unsigned char *)calloc((size_t)g.cap * dim, 1);
          if (!e->mask) goto fail;
          mk = e->mask + (size_t)e->n * dim;
        }
        mk[i] = 1;
        maskcnt++;
        pt[i] = 0.0f;
      } else if (sscanf(tok, "%f", &pt[i]) <= 0) {
        fprintf(stderr, "load_entry: can't read entry in file %s on line %ld, component %d\n", name, row, i);
        goto fail;
      }
    }
    if (maskcnt == dim && skip_empty) continue;                    /* datafile.c:677-690 */
    e->weight[e->n] = 0;
    e->fixed_xy[2 * e->n] = e->fixed_xy[2 * e->n + 1] = -1;
    first_lab = g.nlab;
    while ((tok = strtok(NULL, " \r\t")) != NULL) {
      if (strncmp(tok, "weight=", 7) == 0) {
        e->weight[e->n] = (short)atoi(tok + 7);
      } else if (strncmp(tok, "fixed=", 6) == 0) {
        const char *comma = strchr(tok, ',');
        if (!comma) { fprintf(stderr, "bad fixed point, line %ld of file %s\n", row, name); goto fail; }
        e->fixed_xy[2 * e->n] = (short)atoi(tok + 6);
        e->fixed_xy[2 * e->n + 1] = (short)atoi(comma + 1);
      } else {
        int lab = label_index(tok);
        if (lab == LABEL_EMPTY) continue;
        if (g.nlab == g.labcap) {
          long ncap = g.labcap ? g.labcap * 2 : 1024;
          int *t = (int *)realloc(e->lab_pool, sizeof(int) * (size_t)ncap);
          if (!t) goto fail;
          e->lab_pool = t;
          g.labcap = ncap;
        }
        e->lab_pool[g.nlab++] = lab;
        label_found++;
      }
    }
    (void)first_lab;
    if (labels_needed && !label_found) {
      fprintf(stderr, "Required label missing on line %ld of file %s\n", row, name);
      goto fail;
    }
    e->n++;
    e->lab_off[e->n] = g.nlab;
  }
  if (e->n == 0 && !e->points) {               /* keep the invariants of pak_alloc for empty files */
    if (grow_entries(e, &g, 0)) goto fail;
  }
  free(buf);
  if (fp != stdin) fclose(fp);
  return e;
fail:
  free(buf);
  if (fp != stdin) fclose(fp);
  pak_free(e);
  return NULL;
}

void pak_write_header(FILE *fp, const struct pak_entries *e) {     /* datafile.c:396-415 */
  fprintf(fp, "%d", e->dim);
  if (e->topol > TOPOL_DATA) {
    fprintf(fp, " %s", topol_names[e->topol]);
    if (e->topol > TOPOL_LVQ) fprintf(fp, " %d %d %s", e->xdim, e->ydim, neigh_names[e->neigh] ? neigh_names[e->neigh] : "(null)");
  }
  fprintf(fp, "\n");
}

int pak_save(const struct pak_entries *e, const char *name) {
  FILE *fp = strcmp(name, "-") == 0 ? stdout : fopen(name, "w");
  long i, l;
  int c;
  if (!fp) { fprintf(stderr, "save_entries: Can't open file '%s'\n", name); return 1; }
  pak_write_header(fp, e);
  for (i = 0; i < e->n; i++) {                                     /* datafile.c:420-447 */
    const float *pt = e->points + (size_t)i * e->dim;
    const unsigned char *mk = e->mask ? e->mask + (size_t)i * e->dim : NULL;
    for (c = 0; c < e->dim; c++) {
      if (mk && mk[c]) fprintf(fp, "%s ", pak_mask_string);
      else fprintf(fp, "%g ", pt[c]);
    }
    for (l = e->lab_off[i]; l < e->lab_off[i + 1]; l++) fprintf(fp, "%s ", label_string(e->lab_pool[l]));
    fprintf(fp, "\n");
  }
  if (fp != stdout) fclose(fp);
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
