/* main.c -- multi-call front end of the B200 host programs: `bmu_pak <program> <options>` or a
 * link named after the program (vsom, qerror, visual, vcal, accuracy, classify, knntest, cmatr,
 * setlabel, elimin, lvq1, olvq1, lvq2, lvq3), as the reference installs one binary per program (reference Makefile). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "somhost.h"

static int dispatch(const char *prog, int argc, char **argv) {
  if (strcmp(prog, "vsom") == 0) return vsom_main(argc, argv);
  if (strcmp(prog, "qerror") == 0) return qerror_main(argc, argv);
  if (strcmp(prog, "visual") == 0) return visual_main(argc, argv);
  if (strcmp(prog, "vcal") == 0) return vcal_main(argc, argv);
  if (strcmp(prog, "accuracy") == 0) return accuracy_main(argc, argv);
  if (strcmp(prog, "classify") == 0) return classify_main(argc, argv);
  if (strcmp(prog, "knntest") == 0) return knntest_main(argc, argv);
  if (strcmp(prog, "vfind") == 0) return vfind_main(argc, argv);
  if (strcmp(prog, "randinit") == 0 || strcmp(prog, "lininit") == 0 || strcmp(prog, "mapinit") == 0)
    return randinit_main(argc, argv, prog);
  if (strcmp(prog, "eveninit") == 0 || strcmp(prog, "propinit") == 0) return eveninit_main(argc, argv, prog);
  if (strcmp(prog, "mindist") == 0) return mindist_main(argc, argv);
  if (strcmp(prog, "sammon") == 0) return sammon_main(argc, argv);
  if (strcmp(prog, "balance") == 0) return balance_main(argc, argv);
  if (strcmp(prog, "cmatr") == 0) return cmatr_main(argc, argv);
  if (strcmp(prog, "setlabel") == 0) return setlabel_main(argc, argv);
  if (strcmp(prog, "elimin") == 0) return elimin_main(argc, argv);
  if (strcmp(prog, "pakcat") == 0) return pakcat_main(argc, argv);
  if (strcmp(prog, "pakstat") == 0) return pakstat_main(argc, argv);
  if (strcmp(prog, "paksynth") == 0) return paksynth_main(argc, argv);
  if (strcmp(prog, "lvq1") == 0 || strcmp(prog, "lvq2") == 0 || strcmp(prog, "lvq3") == 0 ||
      strcmp(prog, "olvq1") == 0 || strcmp(prog, "lvqtrain") == 0)
    return lvqtrain_main(argc, argv, prog);
  return -2;
}

/* `bmu_pak batch [file]`: one program per line ("vsom -din ex.dat ..."; '#' starts a comment; "..." or
 * '...' keep blanks in a word), run one after the other in THIS process.  Creating the CUDA context costs
 * about a second, far more than a demo-size program needs; a recipe of several programs pays it once.
 * Every line starts from a fresh label table; the first failing program ends the batch with its code. */
static int batch_main(int argc, char **argv) {
  FILE *fp = argc > 1 && strcmp(argv[1], "-") != 0 ? fopen(argv[1], "r") : stdin;
  char line[8192];
  int rc = 0;
  if (!fp) { fprintf(stderr, "batch: can't open %s\n", argv[1]); return 1; }
  while (rc == 0 && fgets(line, sizeof line, fp)) {
    char *words[512], *p = line;
    int n = 0;
    while (*p && n < 511) {
      while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') p++;
      if (!*p || *p == '#') break;
      if (*p == '"' || *p == '\'') {
        const char q = *p++;
        words[n++] = p;
        while (*p && *p != q) p++;
      } else {
        words[n++] = p;
        while (*p && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') p++;
      }
      if (*p) *p++ = '\0';
    }
    if (n == 0) continue;
    words[n] = NULL;
    label_reset();
    {
      struct timespec t0, t1;
      clock_gettime(CLOCK_MONOTONIC, &t0);
      rc = dispatch(words[0], n, words);
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if (getenv("BMU_PAK_BATCH_TIMING"))
        fprintf(stderr, "[batch] %-10s %.3f s\n", words[0],
                (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
    }
    fflush(stdout);
    if (rc == -2) { fprintf(stderr, "batch: unknown program '%s'\n", words[0]); rc = 2; }
  }
  if (fp != stdin) fclose(fp);
  return rc;
}

int main(int argc, char **argv) {
  const char *base = strrchr(argv[0], '/');
  int rc;
  base = base ? base + 1 : argv[0];
  rc = dispatch(base, argc, argv);
  if (rc == -2 && argc > 1) {
    if (strcmp(argv[1], "batch") == 0) return batch_main(argc - 1, argv + 1);
    rc = dispatch(argv[1], argc - 1, argv + 1);
  }
  if (rc == -2) {
    fprintf(stderr, "usage: bmu_pak <randinit|lininit|eveninit|propinit|balance|mindist|sammon|vsom|vfind|qerror|visual|vcal|accuracy|classify|knntest|cmatr|setlabel|elimin|lvq1|olvq1|lvq2|lvq3|pakcat> <options>\n"
                    "       bmu_pak batch [file]     one program per line, run in one process\n");
    return 2;
  }
  return rc;
}
