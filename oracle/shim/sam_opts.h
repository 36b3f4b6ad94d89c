/* oracle/shim/sam_opts.h -- TEST INFRASTRUCTURE ONLY.  Only SAM_GLOBAL_ARGS_INIT and the
 * .out member are used (stochasticSpike.c:977,980). */
#ifndef SSB_ORACLE_SHIM_SAM_OPTS_H
#define SSB_ORACLE_SHIM_SAM_OPTS_H
#include "htslib/sam.h"
typedef struct sam_global_args {
    htsFormat in, out;
    char *reference;
    int nthreads;
} sam_global_args;
#define SAM_GLOBAL_ARGS_INIT {{0, 0, 0}, {0, 0, 0}, NULL, 0}
#endif
