/*
 * oracle/shim/htslib/sam.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for the part of htslib-1.13's <htslib/sam.h> that the
 * reference's stochasticSpike.c touches (stochasticSpike.c:26 and the call
 * sites listed in SURVEY.md 8c), so that the UNMODIFIED reference source can be
 * compiled from /root/reference into oracle/_ref/.  htslib itself is not in
 * this image; the behaviour behind these declarations (shim.c) is a
 * restatement from the htslib-1.13 API contract that reads
 * SAM *text* instead of BAM.  Written from scratch; no htslib source copied.
 */
#ifndef SSB_ORACLE_SHIM_SAM_H
#define SSB_ORACLE_SHIM_SAM_H

#include <stdint.h>
#include <stddef.h>
#include <limits.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;
#define HTS_POS_MAX ((((int64_t)INT_MAX) << 32) | INT_MAX)

typedef struct kstring_t { size_t l, m; char *s; } kstring_t;
typedef struct htsFormat { int category, format, compression; } htsFormat;

typedef struct shim_file samFile;
typedef struct shim_file htsFile;
typedef struct shim_hdr  sam_hdr_t;
typedef struct shim_idx  hts_idx_t;
typedef struct shim_itr  hts_itr_t;

#define BAM_FPAIRED        1
#define BAM_FPROPER_PAIR   2
#define BAM_FUNMAP         4
#define BAM_FMUNMAP        8
#define BAM_FREVERSE      16
#define BAM_FMREVERSE     32
#define BAM_FREAD1        64
#define BAM_FREAD2       128
#define BAM_FSECONDARY   256
#define BAM_FQCFAIL      512
#define BAM_FDUP        1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CMATCH      0
#define BAM_CINS        1
#define BAM_CDEL        2
#define BAM_CREF_SKIP   3
#define BAM_CSOFT_CLIP  4
#define BAM_CHARD_CLIP  5
#define BAM_CPAD        6
#define BAM_CEQUAL      7
#define BAM_CDIFF       8
#define BAM_CIGAR_SHIFT 4
#define BAM_CIGAR_MASK  0xf

typedef struct bam1_core_t {
    hts_pos_t pos;
    int32_t   tid;
    uint16_t  bin;
    uint8_t   qual;
    uint8_t   l_extranul;
    uint16_t  flag;
    uint16_t  l_qname;
    uint32_t  n_cigar;
    int32_t   l_qseq;
    int32_t   mtid;
    hts_pos_t mpos;
    hts_pos_t isize;
} bam1_core_t;

/* data = qname (NUL terminated) | cigar (uint32 x n_cigar, 4-byte aligned) |
 *        seq (4-bit packed, (l_qseq+1)/2 bytes) | qual (l_qseq bytes) |
 *        aux (shim-private: the SAM text of the optional fields, verbatim)   */
typedef struct bam1_t {
    bam1_core_t core;
    uint64_t    id;
    uint8_t    *data;
    int         l_data;
    uint32_t    m_data;
    int         l_aux_text;   /* shim-private */
} bam1_t;

#define bam_get_qname(b) ((char *)(b)->data)
#define bam_get_cigar(b) ((uint32_t *)((b)->data + (b)->core.l_qname))
#define bam_get_seq(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b)  ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b)   ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_seqi(s, i)   ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)
/* bam_seqi_set is deliberately NOT defined: stochasticSpike.c:4 brings its own. */

extern const char          seq_nt16_str[];
extern const unsigned char seq_nt16_table[256];

typedef struct bam_pileup1_t {
    bam1_t  *b;
    int32_t  qpos;
    int      indel, level;
    uint32_t is_del:1, is_head:1, is_tail:1, is_refskip:1, aux:28;
    int      cigar_ind;
} bam_pileup1_t;

typedef int (*bam_plp_auto_f)(void *data, bam1_t *b);
typedef struct shim_mplp *bam_mplp_t;

samFile   *sam_open(const char *fn, const char *mode);
samFile   *sam_open_format(const char *fn, const char *mode, const htsFormat *fmt);
int        sam_close(samFile *fp);
sam_hdr_t *sam_hdr_read(samFile *fp);
int        sam_hdr_write(samFile *fp, const sam_hdr_t *h);
void       sam_hdr_destroy(sam_hdr_t *h);
int        sam_hdr_count_lines(sam_hdr_t *h, const char *type);
int        sam_hdr_find_tag_pos(sam_hdr_t *h, const char *type, int pos, const char *key, kstring_t *ks);
int        sam_hdr_name2tid(sam_hdr_t *h, const char *ref);
const char *sam_hdr_tid2name(const sam_hdr_t *h, int tid);
int        sam_hdr_nref(const sam_hdr_t *h);
hts_idx_t *sam_index_load(samFile *fp, const char *fn);
hts_itr_t *sam_itr_querys(const hts_idx_t *idx, sam_hdr_t *hdr, const char *region);
int        sam_itr_next(samFile *fp, hts_itr_t *itr, bam1_t *b);
void       hts_itr_destroy(hts_itr_t *itr);
int        sam_read1(samFile *fp, sam_hdr_t *h, bam1_t *b);
int        sam_write1(samFile *fp, const sam_hdr_t *h, const bam1_t *b);
hts_pos_t  bam_cigar2rlen(int n_cigar, const uint32_t *cigar);
int64_t    bam_cigar2qlen(int n_cigar, const uint32_t *cigar);

bam_mplp_t bam_mplp_init(int n, bam_plp_auto_f func, void **data);
void       bam_mplp_set_maxcnt(bam_mplp_t iter, int maxcnt);
int        bam_mplp_auto(bam_mplp_t iter, int *_tid, int *_pos, int *n_plp, const bam_pileup1_t **plp);
void       bam_mplp_destroy(bam_mplp_t iter);

#ifdef __cplusplus
}
#endif
#endif
