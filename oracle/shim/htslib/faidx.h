/* oracle/shim/htslib/faidx.h -- TEST INFRASTRUCTURE ONLY.  Stand-in for the two
 * faidx entry points used at stochasticSpike.c:215-219,1042 (see sam.h here). */
#ifndef SSB_ORACLE_SHIM_FAIDX_H
#define SSB_ORACLE_SHIM_FAIDX_H
#include "sam.h"
typedef struct shim_fai faidx_t;
faidx_t *fai_load(const char *fn);
char    *faidx_fetch_seq64(const faidx_t *fai, const char *c_name, hts_pos_t p_beg_i, hts_pos_t p_end_i, hts_pos_t *len);
void     fai_destroy(faidx_t *fai);
#endif
