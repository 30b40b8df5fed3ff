/* TEST INFRASTRUCTURE ONLY. Minimal dispatcher so that the UNMODIFIED reference sources in
 * /root/reference/lib/aln can be run as `biscuit_ref index|align` (the reference's own
 * dispatcher, src/main.c:105-159, needs htslib/utils which are absent offline). */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

int main_biscuit_index(int argc, char *argv[]);
int main_align(int argc, char *argv[]);
extern char *bwa_pg;

int main(int argc, char *argv[]) {
  if (argc < 2) { fprintf(stderr, "usage: biscuit_ref index|align ...\n"); return 1; }
  bwa_pg = NULL; /* no @PG line: parity is compared modulo @PG */
  if (strcmp(argv[1], "index") == 0) return main_biscuit_index(argc - 1, argv + 1);
  if (strcmp(argv[1], "align") == 0) return main_align(argc - 1, argv + 1);
  fprintf(stderr, "unknown command %s\n", argv[1]);
  return 1;
}
