/* TEST INFRASTRUCTURE ONLY -- stand-in for htslib sam.h (see README.md): the BAM record layout and the calls that
 * src/pileup.c, src/bisc_utils.c make. */
#ifndef BSQ_SHIM_SAM_H
#define BSQ_SHIM_SAM_H
#include "hts.h"
typedef struct bam_hdr_t {
  int32_t n_targets;
  uint32_t *target_len;
  char **target_name;
  char *text;
  size_t l_text;
} bam_hdr_t;
typedef bam_hdr_t sam_hdr_t;
#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9
#define BAM_CIGAR_STR "MIDNSHP=XB"
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK 0xf
#define BAM_CIGAR_TYPE 0x3C1A7
#define bam_cigar_op(c) ((c)&BAM_CIGAR_MASK)
#define bam_cigar_oplen(c) ((c) >> BAM_CIGAR_SHIFT)
#define bam_cigar_type(o) (BAM_CIGAR_TYPE >> ((o) << 1) & 3)
#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048
typedef struct bam1_core_t {
  hts_pos_t pos;
  int32_t tid;
  uint16_t bin;
  uint8_t qual;
  uint8_t l_extranul;
  uint16_t flag;
  uint16_t l_qname;
  uint32_t n_cigar;
  int32_t l_qseq;
  int32_t mtid;
  hts_pos_t mpos;
  hts_pos_t isize;
} bam1_core_t;
typedef struct bam1_t {
  bam1_core_t core;
  uint64_t id;
  uint8_t *data;
  int l_data;
  uint32_t m_data;
  uint32_t mempolicy : 2, : 30;
} bam1_t;
#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_get_l_aux(b) ((b)->l_data - ((b)->core.n_cigar << 2) - (b)->core.l_qname - (b)->core.l_qseq - (((b)->core.l_qseq + 1) >> 1))
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i)&1) << 2) & 0xf)
extern const int8_t bam_cigar_table[256];
bam_hdr_t *sam_hdr_read(htsFile *fp);
void bam_hdr_destroy(bam_hdr_t *h);
int bam_name2id(bam_hdr_t *h, const char *ref);
hts_idx_t *sam_index_load(htsFile *fp, const char *fn);
hts_itr_t *sam_itr_queryi(const hts_idx_t *idx, int tid, hts_pos_t beg, hts_pos_t end);
int sam_itr_next(htsFile *fp, hts_itr_t *iter, bam1_t *b);
bam1_t *bam_init1(void);
void bam_destroy1(bam1_t *b);
hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar);
hts_pos_t bam_endpos(const bam1_t *b);
uint8_t *bam_aux_get(const bam1_t *b, const char tag[2]);
int64_t bam_aux2i(const uint8_t *s);
#endif
