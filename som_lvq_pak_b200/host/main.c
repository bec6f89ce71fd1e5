/* main.c -- multi-call front end of the B200 host programs: `bmu_pak <program> <options>` or a
 * link named after the program (vsom, qerror, visual, vcal, accuracy, classify, knntest, cmatr,
 * setlabel, elimin, lvq1, olvq1, lvq2, lvq3), as the reference installs one binary per program (reference Makefile). */
#include <stdio.h>
#include <string.h>

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
  if (strcmp(prog, "randinit") == 0 || strcmp(prog, "mapinit") == 0) return randinit_main(argc, argv, prog);
  if (strcmp(prog, "eveninit") == 0 || strcmp(prog, "propinit") == 0) return eveninit_main(argc, argv, prog);
  if (strcmp(prog, "mindist") == 0) return mindist_main(argc, argv);
  if (strcmp(prog, "sammon") == 0) return sammon_main(argc, argv);
  if (strcmp(prog, "balance") == 0) return balance_main(argc, argv);
  if (strcmp(prog, "cmatr") == 0) return cmatr_main(argc, argv);
  if (strcmp(prog, "setlabel") == 0) return setlabel_main(argc, argv);
  if (strcmp(prog, "elimin") == 0) return elimin_main(argc, argv);
  if (strcmp(prog, "pakcat") == 0) return pakcat_main(argc, argv);
  if (strcmp(prog, "pakstat") == 0) return pakstat_main(argc, argv);
  if (strcmp(prog, "lvq1") == 0 || strcmp(prog, "lvq2") == 0 || strcmp(prog, "lvq3") == 0 ||
      strcmp(prog, "olvq1") == 0 || strcmp(prog, "lvqtrain") == 0)
    return lvqtrain_main(argc, argv, prog);
  return -2;
}

int main(int argc, char **argv) {
  const char *base = strrchr(argv[0], '/');
  int rc;
  base = base ? base + 1 : argv[0];
  rc = dispatch(base, argc, argv);
  if (rc == -2 && argc > 1) rc = dispatch(argv[1], argc - 1, argv + 1);
  if (rc == -2) {
    fprintf(stderr, "usage: bmu_pak <randinit|eveninit|propinit|balance|mindist|sammon|vsom|vfind|qerror|visual|vcal|accuracy|classify|knntest|cmatr|setlabel|elimin|lvq1|olvq1|lvq2|lvq3|pakcat> <options>\n");
    return 2;
  }
  return rc;
}
