/* TEST INFRASTRUCTURE ONLY -- stand-in for htslib hts.h (see README.md). */
#ifndef BSQ_SHIM_HTS_H
#define BSQ_SHIM_HTS_H
#include <limits.h>
#include <stdint.h>
#include "kstring.h"
typedef int64_t hts_pos_t;
typedef struct htsFile htsFile;
typedef struct hts_idx_t hts_idx_t;
typedef struct hts_itr_t hts_itr_t;
extern const char seq_nt16_str[];
htsFile *hts_open(const char *fn, const char *mode);
int hts_close(htsFile *fp);
void hts_idx_destroy(hts_idx_t *idx);
void hts_itr_destroy(hts_itr_t *iter);
/* "name[:beg[-end]]" -> 0-based beg, end; returns the end of the name part, NULL when not parsable */
const char *hts_parse_reg(const char *str, int *beg, int *end);
#endif
