/* TEST INFRASTRUCTURE ONLY -- stand-in for huishenlab/utils wzbed.h as used by /root/reference/src/{vcf2bed.c,mergecg.c}
 * (see README.md): BED line reader; contig names get ids in order of first appearance. */
#ifndef BSQ_SHIM_SRC_WZBED_H
#define BSQ_SHIM_SRC_WZBED_H
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>
#include "wvec.h"
#include "wzmisc.h"
typedef struct { char *name; int tid; } wz_target_t;
DEFINE_VECTOR(wz_target_v, wz_target_t)
static inline int wz_get_tid(wz_target_v *t, const char *name, int insert) {
  for (size_t i = 0; i < t->size; ++i)
    if (strcmp(t->buffer[i].name, name) == 0) return (int)i;
  if (!insert) return -1;
  wz_target_t *e = next_ref_wz_target_v(t);
  e->name = strdup(name);
  e->tid = (int)t->size - 1;
  return e->tid;
}
static inline char *tid2name(wz_target_v *t, int tid) { return t->buffer[tid].name; }
#define target_name(t, tid) tid2name(t, tid)
static inline void wz_free_targets(wz_target_v *t) {
  for (size_t i = 0; i < t->size; ++i) free(t->buffer[i].name);
  free_wz_target_v(t);
}
typedef struct bed1_t { int tid; int64_t beg; int64_t end; void *data; } bed1_t;
typedef void (*init_data_f)(bed1_t *b, void *aux_data);
typedef void (*free_data_f)(void *data);
typedef void (*parse_data_f)(bed1_t *b, char **fields, int nfields);
static inline bed1_t *init_bed1(init_data_f init_data, void *aux_data) {
  bed1_t *b = (bed1_t *)calloc(1, sizeof(bed1_t));
  if (init_data) init_data(b, aux_data);
  return b;
}
static inline void free_bed1(bed1_t *b, free_data_f free_data) {
  if (free_data) free_data(b->data);
  free(b);
}
/* one whole line from a (possibly gzipped) text file; returns 0 at end of file */
static inline int wz_gzreadline(gzFile fh, char **line, size_t *cap) {
  size_t l = 0;
  if (!*line) { *cap = 1 << 16; *line = (char *)malloc(*cap); }
  for (;;) {
    if (!gzgets(fh, *line + l, (int)(*cap - l))) break;
    l += strlen(*line + l);
    if (l && (*line)[l - 1] == '\n') break;
    if (l + 1 >= *cap) { *cap *= 2; *line = (char *)realloc(*line, *cap); }
  }
  if (!l) return 0;
  while (l && ((*line)[l - 1] == '\n' || (*line)[l - 1] == '\r')) (*line)[--l] = 0;
  return 1;
}
typedef struct bed_file_t { char *file_path; gzFile fh; char *line; size_t cap; wz_target_v *targets; } bed_file_t;
static inline bed_file_t *init_bed_file(char *file_path) {
  bed_file_t *bed = (bed_file_t *)calloc(1, sizeof(bed_file_t));
  bed->fh = strcmp(file_path, "-") == 0 ? gzdopen(fileno(stdin), "r") : gzopen(file_path, "r");
  if (!bed->fh) wzfatal("Could not read file: %s\n", file_path);
  bed->file_path = strdup(file_path);
  bed->targets = init_wz_target_v(16);
  return bed;
}
static inline void free_bed_file(bed_file_t *bed) {
  gzclose(bed->fh); free(bed->line); free(bed->file_path); wz_free_targets(bed->targets); free(bed);
}
static inline int bed_read1(bed_file_t *bed, bed1_t *b, parse_data_f parse_data) {
  for (;;) {
    if (!wz_gzreadline(bed->fh, &bed->line, &bed->cap)) return 0;
    if (bed->line[0] && bed->line[0] != '#') break;
  }
  char **fields; int nfields;
  line_get_fields(bed->line, "\t", &fields, &nfields);
  if (nfields < 3) wzfatal("[%s:%d] Bed file has fewer than 3 columns.\n", __func__, __LINE__);
  b->tid = wz_get_tid(bed->targets, fields[0], 1);
  b->beg = atoll(fields[1]);
  b->end = atoll(fields[2]);
  if (parse_data) parse_data(b, fields, nfields);
  free_char_array(fields, nfields);
  return 1;
}
#endif
