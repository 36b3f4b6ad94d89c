/* oracle/shim/samtools.h -- TEST INFRASTRUCTURE ONLY.  stochasticSpike.c:27 includes
 * samtools.h but uses nothing from it. */
#ifndef SSB_ORACLE_SHIM_SAMTOOLS_H
#define SSB_ORACLE_SHIM_SAMTOOLS_H
#endif
