/* TEST INFRASTRUCTURE ONLY.  Dispatcher so that the UNMODIFIED reference sources src/pileup.c, src/vcf2bed.c and
 * src/mergecg.c run as `biscuit_ref_src pileup|vcf2bed|mergecg ...` (the reference's own dispatcher, src/main.c:105-159,
 * needs htslib's version.h and every other subcommand). */
#include <stdio.h>
#include <string.h>
int main_pileup(int argc, char *argv[]);
int main_vcf2bed(int argc, char *argv[]);
int main_mergecg(int argc, char *argv[]);
int main(int argc, char *argv[]) {
  if (argc < 2) { fprintf(stderr, "usage: biscuit_ref_src pileup|vcf2bed|mergecg ...\n"); return 1; }
  int r = 1;
  if (strcmp(argv[1], "pileup") == 0) r = main_pileup(argc - 1, argv + 1);
  else if (strcmp(argv[1], "vcf2bed") == 0) r = main_vcf2bed(argc - 1, argv + 1);
  else if (strcmp(argv[1], "mergecg") == 0) r = main_mergecg(argc - 1, argv + 1);
  else fprintf(stderr, "unknown command %s\n", argv[1]);
  fflush(stdout);
  return r;
}
