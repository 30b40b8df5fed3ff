/* TEST INFRASTRUCTURE ONLY -- stand-in for htslib faidx.h (see README.md). */
#ifndef BSQ_SHIM_FAIDX_H
#define BSQ_SHIM_FAIDX_H
typedef struct faidx_t faidx_t;
faidx_t *fai_load(const char *fn);
void fai_destroy(faidx_t *fai);
int faidx_seq_len(const faidx_t *fai, const char *seq);
/* bases [p_beg_i, p_end_i] (0-based, inclusive, clipped to the sequence) as a malloc()ed string */
char *faidx_fetch_seq(const faidx_t *fai, const char *c_name, int p_beg_i, int p_end_i, int *len);
#endif
