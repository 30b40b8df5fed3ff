/* bq_vcf2bed.c -- `biscuit vcf2bed` (src/vcf2bed.c:146-382) and `biscuit mergecg` (src/mergecg.c:52-231):
 * the text transforms that turn the pileup VCF into the methylation BED files.  The reference parses VCF/BED
 * through huishenlab/utils wzvcf.h / wzbed.h (not vendored); the column rules below are the VCF 4.1 ones.
 */
#include <ctype.h>
#include <errno.h>
#include <getopt.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <zlib.h>
#include "bq_plp.h"

/* ---- line reader over gz or plain text ---- */
typedef struct { gzFile fp; char *buf; size_t cap; } lines_t;

static int next_line(lines_t *L) {
  size_t n = 0;
  for (;;) {
    if (L->cap - n < 4096) { L->cap = L->cap ? L->cap * 2 : 1 << 16; L->buf = realloc(L->buf, L->cap); }
    if (!gzgets(L->fp, L->buf + n, (int)(L->cap - n))) { if (n == 0) return 0; break; }
    n += strlen(L->buf + n);
    if (n && L->buf[n - 1] == '\n') break;
  }
  while (n && (L->buf[n - 1] == '\n' || L->buf[n - 1] == '\r')) L->buf[--n] = 0;
  return 1;
}

static int split(char *s, char sep, char **f, int max) {
  int n = 0;
  f[n++] = s;
  for (; *s; ++s)
    if (*s == sep) { *s = 0; if (n < max) f[n++] = s + 1; else break; }
  return n;
}

static int is_number(const char *s) { /* optional sign, digits, optional fraction / exponent */
  char *e;
  if (!*s) return 0;
  strtod(s, &e);
  return *e == 0 && e != s;
}

/* value of INFO key, or NULL; result is malloc()ed */
static char *info_get(const char *info, const char *key) {
  const size_t lk = strlen(key);
  const char *p = info;
  while (p && *p) {
    const char *e = strchr(p, ';');
    const size_t l = e ? (size_t)(e - p) : strlen(p);
    if (l > lk && strncmp(p, key, lk) == 0 && p[lk] == '=') return strndup(p + lk + 1, l - lk - 1);
    p = e ? e + 1 : 0;
  }
  return 0;
}

/* index of key in the colon-separated FORMAT string, -1 if absent */
static int fmt_index(const char *fmt, const char *key) {
  const size_t lk = strlen(key);
  int i = 0;
  const char *p = fmt;
  while (*p) {
    const char *e = strchr(p, ':');
    const size_t l = e ? (size_t)(e - p) : strlen(p);
    if (l == lk && strncmp(p, key, lk) == 0) return i;
    if (!e) break;
    p = e + 1; ++i;
  }
  return -1;
}

/* i-th colon-separated field of a sample column, copied into out ("." if missing) */
static void sample_field(const char *col, int idx, char *out, size_t cap) {
  const char *p = col;
  for (int i = 0; i < idx && p; ++i) { p = strchr(p, ':'); if (p) ++p; }
  if (!p) { strcpy(out, "."); return; }
  const char *e = strchr(p, ':');
  size_t l = e ? (size_t)(e - p) : strlen(p);
  if (l >= cap) l = cap - 1;
  memcpy(out, p, l); out[l] = 0;
}

typedef struct { char target[5]; int mincov, showctxt, showmu; } v2b_conf_t;

static int v2b_usage(void) {
  fprintf(stderr, "\n");
  fprintf(stderr, "Usage: biscuit vcf2bed [options] <in.vcf>\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "Options:\n");
  fprintf(stderr, "    -t STR    Extract type {c, cg, ch, hcg, gch, snp} [CG]\n");
  fprintf(stderr, "    -k INT    Minimum coverage (see Note 1) [1]\n");
  fprintf(stderr, "    -s STR    Sample, (takes \"FIRST\", \"LAST\", \"ALL\", or specific\n");
  fprintf(stderr, "                  sample names separated by \",\") [FIRST]\n");
  fprintf(stderr, "    -e        Show context (reference base, context group {CG,CHG,CHH},\n");
  fprintf(stderr, "                  2-base {CA,CC,CG,CT} and 5-base context) before beta\n");
  fprintf(stderr, "                  value and coverage column\n");
  fprintf(stderr, "    -c        Output Beta-M-U instead of Beta-Cov.\n");
  fprintf(stderr, "    -h        This help\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "Note 1: Starting with version 1.6.0, the default minimum coverage was changed from\n");
  fprintf(stderr, "        three (3) to one (1) to better match other tools and serve as a better default\n");
  fprintf(stderr, "        for single-cell experiments.\n");
  fprintf(stderr, "\n");
  return 1;
}

int bq_main_vcf2bed(int argc, char **argv) {
  v2b_conf_t conf;
  conf.mincov = 1; conf.showctxt = 0; conf.showmu = 0; strcpy(conf.target, "CG");
  char *target_samples = 0;
  int c;
  if (argc < 2) return v2b_usage();
  while ((c = getopt(argc, argv, ":t:k:s:ech")) >= 0) {
    switch (c) {
      case 'k': conf.mincov = atoi(optarg); break;
      case 't': if (strlen(optarg) > 4) bq_fatal("Invalid option for -t: %s.\n", optarg); strcpy(conf.target, optarg); break;
      case 's': target_samples = strdup(optarg); break;
      case 'e': conf.showctxt = 1; break;
      case 'c': conf.showmu = 1; break;
      case 'h': return v2b_usage();
      case ':': v2b_usage(); bq_fatal("Option needs an argument: -%c\n", optopt); break;
      default: v2b_usage(); bq_fatal("Unrecognized option: -%c\n", optopt); break;
    }
  }
  if (!target_samples) target_samples = strdup("FIRST");
  if (optind >= argc) { v2b_usage(); bq_fatal("Please provide input vcf.\n"); }
  char *raw_target = strdup(conf.target);
  for (char *p = conf.target; *p; ++p) *p = (char)toupper((unsigned char)*p);
  if (strcmp(conf.target, "CG") && strcmp(conf.target, "CH") && strcmp(conf.target, "C") && strcmp(conf.target, "HCG") &&
      strcmp(conf.target, "GCH") && strcmp(conf.target, "SNP"))
    bq_fatal("Invalid option for -t: %s.\n", raw_target);
  free(raw_target);
  const int snp = strcmp(conf.target, "SNP") == 0;
  const char *cx = conf.target;

  lines_t L = {0, 0, 0};
  L.fp = strcmp(argv[optind], "-") == 0 ? gzdopen(0, "r") : gzopen(argv[optind], "r");
  if (!L.fp) bq_fatal("Cannot open %s\n", argv[optind]);
  gzbuffer(L.fp, 1 << 20);
  int n_samples = 0, n_sel = 0, *sel = 0;
  char **f = malloc(4096 * sizeof(char *));
  double *betas = 0;
  int *covs = 0;
  setvbuf(stdout, 0, _IOFBF, 1 << 22);
  while (next_line(&L)) {
    if (L.buf[0] == '#') {
      if (strncmp(L.buf, "#CHROM", 6) == 0) { /* sample columns -> selection (index_vcf_samples) */
        char *hl = strdup(L.buf);
        const int nf = split(hl, '\t', f, 4096);
        n_samples = nf > 9 ? nf - 9 : 0;
        sel = calloc((size_t)n_samples + 1, sizeof(int));
        if (strcmp(target_samples, "FIRST") == 0) { if (n_samples) sel[n_sel++] = 0; }
        else if (strcmp(target_samples, "LAST") == 0) { if (n_samples) sel[n_sel++] = n_samples - 1; }
        else if (strcmp(target_samples, "ALL") == 0) { for (int i = 0; i < n_samples; ++i) sel[n_sel++] = i; }
        else {
          char *ts = strdup(target_samples), *names[256];
          const int nn = split(ts, ',', names, 256);
          for (int k = 0; k < nn; ++k) {
            int found = -1;
            for (int i = 0; i < n_samples; ++i) if (strcmp(f[9 + i], names[k]) == 0) found = i;
            if (found < 0) bq_fatal("Sample %s not found in the VCF header.\n", names[k]);
            sel[n_sel++] = found;
          }
          free(ts);
        }
        betas = calloc((size_t)n_sel + 1, sizeof(double));
        covs = calloc((size_t)n_sel + 1, sizeof(int));
        free(hl);
      }
      continue;
    }
    if (!sel) bq_fatal("Malformed VCF file: no #CHROM line.\n");
    const int nf = split(L.buf, '\t', f, 4096);
    if (nf < 9 + n_samples) continue;
    const char *chrom = f[0], *ref = f[3], *alt = f[4], *info = f[7], *fmt = f[8];
    const long pos = atol(f[1]);
    if (!snp) {
      char *info_cx = info_get(info, "CX"), *info_n5 = info_get(info, "N5");
      if (!info_cx) { free(info_n5); continue; }
      int skip = 0;
      if (strcmp(cx, "C") == 0) { if (ref[0] != 'C' && ref[0] != 'G') skip = 1; }
      else if (strcmp(cx, "CH") == 0) { if (strcmp(info_cx, "CHH") != 0 && strcmp(info_cx, "CHG") != 0) skip = 1; }
      else if (strcmp(info_cx, cx) != 0) skip = 1;
      if (!skip) {
        const int ibt = fmt_index(fmt, "BT"), icv = fmt_index(fmt, "CV");
        int pass = 0;
        for (int k = 0; k < n_sel; ++k) {
          char v[64];
          betas[k] = -1.0; covs[k] = 0;
          if (ibt >= 0) { sample_field(f[9 + sel[k]], ibt, v, sizeof v); if (is_number(v) && strcmp(v, ".") != 0) betas[k] = atof(v); }
          if (icv >= 0) { sample_field(f[9 + sel[k]], icv, v, sizeof v); if (is_number(v) && strcmp(v, ".") != 0) covs[k] = atoi(v); }
          if (covs[k] >= conf.mincov) pass = 1;
        }
        if (pass) {
          char n5[6];
          if (!info_n5 || strlen(info_n5) != 5) strcpy(n5, "NNNNN"); else strcpy(n5, info_n5);
          fprintf(stdout, "%s\t%ld\t%ld", chrom, pos - 1, pos);
          if (conf.showctxt) fprintf(stdout, "\t%c\t%s\t%.2s\t%.5s", ref[0], info_cx, n5 + 2, n5);
          for (int k = 0; k < n_sel; ++k) {
            if (conf.showmu) {
              const int M = (int)round(covs[k] * betas[k]);
              if (betas[k] < 0) fputs("\t.", stdout); else fprintf(stdout, "\t%d", (int)round(betas[k] * 100));
              fprintf(stdout, "\t%d\t%d", M, covs[k] - M);
            } else {
              if (betas[k] < 0) fputs("\t.", stdout); else fprintf(stdout, "\t%1.3f", betas[k]);
              fprintf(stdout, "\t%d", covs[k]);
            }
          }
          if (fputc('\n', stdout) < 0 && errno == EPIPE) exit(1);
        }
      }
      free(info_cx); free(info_n5);
    } else if (strcmp(alt, ".") != 0) {
      const int igt = fmt_index(fmt, "GT"), isp = fmt_index(fmt, "SP"), iac = fmt_index(fmt, "AC"), iaf = fmt_index(fmt, "AF1");
      if (igt < 0 || isp < 0 || iac < 0 || iaf < 0) bq_fatal("Malformed VCF file (unmatched no. records) in %s\n", chrom);
      int highest_cov = 0;
      double highest_af = 0.0;
      char v[256];
      for (int k = 0; k < n_sel; ++k) {
        sample_field(f[9 + sel[k]], iac, v, sizeof v);
        const int cov = atoi(v);
        if (cov > highest_cov) highest_cov = cov;
        sample_field(f[9 + sel[k]], iaf, v, sizeof v);
        const double af = atof(v);
        if (af > highest_af) highest_af = af;
      }
      if (highest_cov < conf.mincov || highest_af <= 0.0) continue;
      fprintf(stdout, "%s\t%ld\t%ld\t%s\t%s", chrom, pos - 1, pos, ref, alt);
      for (int k = 0; k < n_sel; ++k) {
        const int idx[4] = {igt, isp, iac, iaf};
        for (int q = 0; q < 4; ++q) { sample_field(f[9 + sel[k]], idx[q], v, sizeof v); putchar('\t'); fputs(v, stdout); }
      }
      if (fputc('\n', stdout) < 0 && errno == EPIPE) exit(1);
    }
  }
  gzclose(L.fp);
  free(L.buf); free(f); free(sel); free(betas); free(covs); free(target_samples);
  return 0;
}

/* ------------------------------------------------------------------ mergecg ---- */

typedef struct {
  char *chrom;
  long beg, end;
  int valid; /* tid >= 0 */
  char ref;
  int nsamples;
  double *c_betas, *g_betas;
  int *c_depts, *g_depts;
} mcg_row_t;

typedef struct { int nome_mode, min_depth, show_mu; } mcg_conf_t;

static void mcg_output(mcg_row_t *p, char base_before, char base_after, mcg_conf_t conf) { /* format_output, src/mergecg.c:90-137 */
  int max_depth = 0;
  for (int i = 0; i < p->nsamples; ++i)
    if (p->c_depts[i] + p->g_depts[i] > max_depth) max_depth = p->c_depts[i] + p->g_depts[i];
  if (max_depth == 0 || max_depth < conf.min_depth) return;
  if (p->ref == 'C' && base_after == 'G') p->end++;
  else if (p->ref == 'G' && base_before == 'C') p->beg--;
  printf("%s\t%ld\t%ld", p->chrom, p->beg, p->end);
  for (int i = 0; i < p->nsamples; ++i) {
    const int cov = p->c_depts[i] + p->g_depts[i];
    if (cov == 0) fputs(conf.show_mu ? "\t.\t0\t0" : "\t.\t0", stdout);
    else {
      const float c_ret = rintf(p->c_betas[i] * p->c_depts[i]);
      const float g_ret = rintf(p->g_betas[i] * p->g_depts[i]);
      const int M = c_ret + g_ret;
      if (conf.show_mu) printf("\t%d\t%d\t%d", (int)round(M / (double)cov * 100), M, cov - M);
      else printf("\t%1.3f\t%d", M / (double)cov, cov);
    }
    if (p->c_depts[i] == 0) fputs("\tC:.:0", stdout); else printf("\tC:%1.3f:%d", p->c_betas[i], p->c_depts[i]);
    if (p->g_depts[i] == 0) fputs(",G:.:0", stdout); else printf(",G:%1.3f:%d", p->g_betas[i], p->g_depts[i]);
  }
  putchar('\n');
}

static int mcg_usage(int min_depth) {
  fprintf(stderr, "\n");
  fprintf(stderr, "Usage: biscuit mergecg [options] <ref.fa> <in.bed>\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "Options:\n");
  fprintf(stderr, "    -N        NOMe-seq mode, only merge C,G both in HCGD context\n");
  fprintf(stderr, "    -c        Output Beta-M-U instead of Beta-Cov (input is still Beta-Cov).\n");
  fprintf(stderr, "    -k INT    Minimum depth after merging - applies to the maximum depth\n");
  fprintf(stderr, "                  across samples [%d]\n", min_depth);
  fprintf(stderr, "    -h        This help\n");
  fprintf(stderr, "\n");
  fprintf(stderr, "Note, in.bed is a position sorted bed file with beta values and coverages found\n");
  fprintf(stderr, "    in columns 4 and 5, respectively. Additional beta value-coverage column\n");
  fprintf(stderr, "    pairs are added for each additional sample. This is the format that would be\n");
  fprintf(stderr, "    found in the output of biscuit vcf2bed without the '-e' flag included.\n");
  fprintf(stderr, "\n");
  return 1;
}

int bq_main_mergecg(int argc, char **argv) {
  int c;
  mcg_conf_t conf = {0, 0, 0};
  if (argc < 2) return mcg_usage(conf.min_depth);
  while ((c = getopt(argc, argv, ":k:hNc")) >= 0) {
    switch (c) {
      case 'N': conf.nome_mode = 1; break;
      case 'k': conf.min_depth = atoi(optarg); break;
      case 'h': return mcg_usage(conf.min_depth);
      case 'c': conf.show_mu = 1; break;
      case ':': mcg_usage(conf.min_depth); bq_fatal("Option needs an argument: -%c\n", optopt); break;
      default: mcg_usage(conf.min_depth); bq_fatal("Unrecognized option: -%c\n", optopt); break;
    }
  }
  if (optind + 2 > argc) { mcg_usage(conf.min_depth); bq_fatal("Please supply reference file and sorted bed file.\n"); }
  bq_fasta_t fa;
  if (bq_fasta_load(argv[optind], &fa) != 0) bq_fatal("Cannot open reference %s\n", argv[optind]);
  optind++;
  lines_t L = {0, 0, 0};
  L.fp = strcmp(argv[optind], "-") == 0 ? gzdopen(0, "r") : gzopen(argv[optind], "r");
  if (!L.fp) bq_fatal("Cannot open %s\n", argv[optind]);
  gzbuffer(L.fp, 1 << 20);
  setvbuf(stdout, 0, _IOFBF, 1 << 22);
  mcg_row_t rows[2];
  memset(rows, 0, sizeof rows);
  mcg_row_t *b = &rows[0], *p = &rows[1];
  char p_before = 'N', p_after = 'N', b_before = 'N', b_after = 'N';
  char *cur_chrom = 0;
  uint8_t *ref = 0;
  int64_t ref_len = 0;
  char **f = malloc(4096 * sizeof(char *));
  static const char nt4c[5] = "ACGTN";
  while (next_line(&L)) {
    if (L.buf[0] == 0 || L.buf[0] == '#') continue;
    const int nf = split(L.buf, '\t', f, 4096);
    if (nf < 5) bq_fatal("Malformed bed file.\n");
    /* parse_data_meth, src/mergecg.c:64-88 */
    const int start = (strcmp(f[3], "C") == 0 || strcmp(f[3], "G") == 0) ? 7 : 3;
    if (b->nsamples <= 0) {
      b->nsamples = (nf - start) / 2;
      if (b->nsamples <= 0) bq_fatal("No sample data identified.\n");
      b->c_betas = calloc((size_t)b->nsamples, sizeof(double)); b->g_betas = calloc((size_t)b->nsamples, sizeof(double));
      b->c_depts = calloc((size_t)b->nsamples, sizeof(int)); b->g_depts = calloc((size_t)b->nsamples, sizeof(int));
    } else if (b->nsamples * 2 + start != nf) bq_fatal("Malformed bed file.\n");
    free(b->chrom);
    b->chrom = strdup(f[0]); b->beg = atol(f[1]); b->end = atol(f[2]); b->valid = 1;
    for (int i = 0; i < b->nsamples; ++i) {
      b->c_betas[i] = atof(f[start + 2 * i]); b->c_depts[i] = atoi(f[start + 1 + 2 * i]);
      b->g_betas[i] = 0; b->g_depts[i] = 0;
    }
    if (!cur_chrom || strcmp(cur_chrom, b->chrom) != 0) {
      free(ref); ref = 0;
      ref_len = bq_fasta_fetch_nt4(&fa, b->chrom, &ref);
      if (ref_len < 0) bq_fatal("Contig %s is not in the reference.\n", b->chrom);
      free(cur_chrom); cur_chrom = strdup(b->chrom);
    }
#define BASE1(pos1) (((pos1) >= 1 && (pos1) <= ref_len) ? nt4c[ref[(pos1)-1]] : 'N')
    b->ref = BASE1(b->end);
    b_before = b->end - 1 < 0 ? 'N' : BASE1(b->end - 1);
    b_after = b->end == ref_len ? 'N' : BASE1(b->end + 1);
    if (b->ref == 'G') {
      memcpy(b->g_betas, b->c_betas, (size_t)b->nsamples * sizeof(double)); memcpy(b->g_depts, b->c_depts, (size_t)b->nsamples * sizeof(int));
      memset(b->c_betas, 0, (size_t)b->nsamples * sizeof(double)); memset(b->c_depts, 0, (size_t)b->nsamples * sizeof(int));
    }
    if (p->valid && b->valid && strcmp(b->chrom, p->chrom) == 0 && b->beg == p->beg + 1 && b->end == p->end + 1 && b->ref == 'G' &&
        p->ref == 'C' && (!conf.nome_mode || (p_before != 'G' && b_after != 'C'))) {
      if (p->nsamples != b->nsamples) bq_fatal("Missing sample at %s:%ld-%ld.\n", b->chrom, b->beg, b->end);
      memcpy(p->g_betas, b->g_betas, (size_t)b->nsamples * sizeof(double)); memcpy(p->g_depts, b->g_depts, (size_t)b->nsamples * sizeof(int));
      b->valid = 0;
    }
    if (p->valid) mcg_output(p, p_before, p_after, conf);
    mcg_row_t *tmp = p; p = b; b = tmp;
    p_before = b_before; p_after = b_after;
  }
  if (p->valid) mcg_output(p, p_before, p_after, conf);
  for (int i = 0; i < 2; ++i) { free(rows[i].chrom); free(rows[i].c_betas); free(rows[i].g_betas); free(rows[i].c_depts); free(rows[i].g_depts); }
  free(ref); free(cur_chrom); free(f); free(L.buf);
  gzclose(L.fp);
  bq_fasta_free(&fa);
  return 0;
}
