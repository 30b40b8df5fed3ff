/* TEST INFRASTRUCTURE ONLY -- stand-in for huishenlab/utils wzvcf.h as used by /root/reference/src/vcf2bed.c (see
 * README.md): VCF line tokenising and INFO / FORMAT look-ups.  Which rows are printed and how is the reference's code. */
#ifndef BSQ_SHIM_SRC_WZVCF_H
#define BSQ_SHIM_SRC_WZVCF_H
#include "wzbed.h"
typedef struct vcf_file_t {
  char *file_path;
  gzFile fh;
  char *line; size_t cap;
  int nsamples; char **samples;
  int n_tsamples; int *tsample_indices;
  wz_target_v *targets;
} vcf_file_t;
typedef struct vcf_record_t {
  int tid; char *chrm; int64_t pos; char *id; char *ref; char *alt; char *qual; char *filter; char *info;
  char **fmt; int nfmt;   /* fmt[0] = FORMAT column, fmt[1..] = sample columns, NULL-terminated */
} vcf_record_t;
static inline vcf_file_t *init_vcf_file(char *file_path) {
  vcf_file_t *vcf = (vcf_file_t *)calloc(1, sizeof(vcf_file_t));
  vcf->fh = strcmp(file_path, "-") == 0 ? gzdopen(fileno(stdin), "r") : gzopen(file_path, "r");
  if (!vcf->fh) wzfatal("Could not read file: %s\n", file_path);
  vcf->file_path = strdup(file_path);
  vcf->targets = init_wz_target_v(16);
  return vcf;
}
static inline void free_vcf_file(vcf_file_t *vcf) {
  gzclose(vcf->fh); free(vcf->line); free(vcf->file_path); wz_free_targets(vcf->targets);
  free_char_array(vcf->samples, vcf->nsamples); free(vcf->tsample_indices); free(vcf);
}
/* read the header up to and including #CHROM, then choose the target samples: FIRST, LAST, ALL or names */
static inline void index_vcf_samples(vcf_file_t *vcf, char *sample_str) {
  while (wz_gzreadline(vcf->fh, &vcf->line, &vcf->cap)) {
    if (strncmp(vcf->line, "#CHROM", 6) == 0) {
      char **f; int n;
      line_get_fields(vcf->line, "\t", &f, &n);
      vcf->nsamples = n > 9 ? n - 9 : 0;
      vcf->samples = (char **)calloc((size_t)vcf->nsamples + 1, sizeof(char *));
      for (int i = 0; i < vcf->nsamples; ++i) vcf->samples[i] = strdup(f[9 + i]);
      free_char_array(f, n);
      break;
    }
  }
  if (vcf->nsamples <= 0) wzfatal("No sample found in %s\n", vcf->file_path);
  vcf->tsample_indices = (int *)calloc((size_t)vcf->nsamples + 1, sizeof(int));
  if (strcmp(sample_str, "FIRST") == 0) { vcf->n_tsamples = 1; vcf->tsample_indices[0] = 0; }
  else if (strcmp(sample_str, "LAST") == 0) { vcf->n_tsamples = 1; vcf->tsample_indices[0] = vcf->nsamples - 1; }
  else if (strcmp(sample_str, "ALL") == 0) { vcf->n_tsamples = vcf->nsamples; for (int i = 0; i < vcf->nsamples; ++i) vcf->tsample_indices[i] = i; }
  else {
    char **f; int n;
    line_get_fields(sample_str, ",", &f, &n);
    for (int k = 0; k < n; ++k) {
      int hit = -1;
      for (int i = 0; i < vcf->nsamples; ++i) if (strcmp(vcf->samples[i], f[k]) == 0) hit = i;
      if (hit < 0) wzfatal("Sample %s not found in %s\n", f[k], vcf->file_path);
      vcf->tsample_indices[vcf->n_tsamples++] = hit;
    }
    free_char_array(f, n);
  }
}
static inline vcf_record_t *init_vcf_record(void) { return (vcf_record_t *)calloc(1, sizeof(vcf_record_t)); }
static inline void wz_clear_vcf_record(vcf_record_t *r) {
  free(r->chrm); free(r->id); free(r->ref); free(r->alt); free(r->qual); free(r->filter); free(r->info);
  free_char_array(r->fmt, r->nfmt);
  memset(r, 0, sizeof *r);
}
static inline void free_vcf_record(vcf_record_t *r) { wz_clear_vcf_record(r); free(r); }
static inline int vcf_read_record(vcf_file_t *vcf, vcf_record_t *rec) {
  for (;;) {
    if (!wz_gzreadline(vcf->fh, &vcf->line, &vcf->cap)) return 0;
    if (vcf->line[0] && vcf->line[0] != '#') break;
  }
  char **f; int n;
  line_get_fields(vcf->line, "\t", &f, &n);
  if (n < 8) wzfatal("[%s:%d] Invalid VCF line: %s\n", __func__, __LINE__, vcf->line);
  wz_clear_vcf_record(rec);
  rec->chrm = f[0]; rec->pos = atoll(f[1]); rec->id = f[2]; rec->ref = f[3]; rec->alt = f[4];
  rec->qual = f[5]; rec->filter = f[6]; rec->info = f[7];
  free(f[1]);
  rec->nfmt = n > 8 ? n - 8 : 0;
  rec->fmt = (char **)calloc((size_t)rec->nfmt + 1, sizeof(char *));
  for (int i = 0; i < rec->nfmt; ++i) rec->fmt[i] = f[8 + i];
  free(f);
  rec->tid = wz_get_tid(vcf->targets, rec->chrm, 1);
  return 1;
}
/* value of KEY in the ';'-separated INFO column (strdup'ed), NULL when absent */
static inline char *get_vcf_record_info(const char *key, char *info) {
  size_t kl = strlen(key);
  for (const char *p = info; p && *p;) {
    size_t l = strcspn(p, ";");
    if (l > kl && strncmp(p, key, kl) == 0 && p[kl] == '=') return strndup(p + kl + 1, l - kl - 1);
    p += l;
    if (*p) ++p;
  }
  return NULL;
}
/* sub-field KEY of every target sample; *out = NULL, *n = 0 when FORMAT has no such key */
static inline void get_vcf_record_fmt(const char *key, char **fmt, vcf_file_t *vcf, char ***out, int *n) {
  *out = NULL; *n = 0;
  if (!fmt || !fmt[0]) return;
  char **keys; int nkeys, idx = -1;
  line_get_fields(fmt[0], ":", &keys, &nkeys);
  for (int i = 0; i < nkeys; ++i) if (strcmp(keys[i], key) == 0) { idx = i; break; }
  free_char_array(keys, nkeys);
  if (idx < 0) return;
  *out = (char **)calloc((size_t)vcf->n_tsamples + 1, sizeof(char *));
  for (int k = 0; k < vcf->n_tsamples; ++k) {
    const char *col = fmt[1 + vcf->tsample_indices[k]];
    char **vals; int nvals;
    line_get_fields(col ? col : ".", ":", &vals, &nvals);
    (*out)[k] = strdup(idx < nvals ? vals[idx] : ".");
    free_char_array(vals, nvals);
  }
  *n = vcf->n_tsamples;
}
#endif
