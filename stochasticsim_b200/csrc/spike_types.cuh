// spike_types.cuh -- device-side data layout of the spike path (hot path 1).
//
// Everything the pipeline needs per alignment line lives in one 64-byte record produced by the
// single tokenising pass over the SAM text (sam_parse.cu).  Later stages never re-read the text
// except (a) the emit kernel, which copies whole lines, and (b) the pileup gather, which looks at
// CIGAR/SEQ/QUAL of the few reads that overlap a spike target.
#pragma once
#include <stdint.h>

struct __align__(16) SamRec {
    uint64_t line_off;     // byte offset of the line in the SAM body
    uint64_t qhash;        // FNV-1a of QNAME (mate search: hash first, bytes on a hit)
    uint32_t line_len;     // bytes including the '\n' (if the line has one)
    int32_t  tid;          // index of RNAME among the @SQ names, -1 for '*' / unknown
    int32_t  pos;          // 0-based leftmost position
    int32_t  end;          // pos + reference length of the CIGAR (M,D,N,=,X)
    uint32_t seq_off;      // offset of SEQ inside the line
    uint32_t l_seq;        // SEQ length (0 for '*')
    uint32_t qual_off;     // offset of QUAL inside the line
    uint16_t cigar_off;    // offset of CIGAR inside the line
    uint16_t cigar_len;    // bytes of CIGAR text
    uint16_t flag;
    uint16_t qname_len;
    uint8_t  mapq;
    uint8_t  bits;         // REC_*
    uint8_t  pad[6];
};
static_assert(sizeof(SamRec) == 64, "SamRec layout");

enum : uint8_t {
    REC_KEEP     = 1,      // passes read_bam (stochasticSpike.c:243-268), tid >= 0, reference length > 0
    REC_QUALSTAR = 2,      // QUAL is '*'
    REC_NO_NL    = 4,      // last line of the body without a trailing '\n'
    REC_PUSHED   = 8,      // passes read_bam and tid >= 0 (takes part in the sortedness check)
    REC_SIMPLE   = 16,     // CIGAR holds only M/=/X: query offset q aligns to reference position pos + q
    REC_EXC_DONE = 32,     // the tokeniser already listed this read's exceptional bases (see sam_parse.cuh)
    REC_AUX_F    = 64      // the line carries float-typed optional fields, judged by aux_float_kernel
};

// device error word: first error wins
struct SpikeErr { int code; unsigned long long where; };

// one pileup entry at a spike target, in pileup (= input) order
struct __align__(16) PlpEntry {
    uint32_t ord;          // kept-read ordinal
    uint32_t qpos;
    int32_t  mate;         // index of the first later entry with the same QNAME, -1 if none
    uint8_t  base;         // read base at qpos (SEQ byte)
    uint8_t  bq;           // base quality (0xff when QUAL is '*')
    uint8_t  skip;         // is_del | is_refskip (bit 0), for bookkeeping only
    uint8_t  pad;
};
static_assert(sizeof(PlpEntry) == 16, "PlpEntry layout");

struct __align__(16) Patch { uint32_t ord; uint32_t qpos; uint32_t base; uint32_t pad; };

// a spiked base of a read that the NEXT shard writes: the read is named by its line index counted from the end of the body
struct __align__(16) FwdPatch { uint32_t from_end; uint32_t qpos; uint32_t base; uint32_t order; };

// a spike target that coincides with a covered locus, in covered order
struct HitTarget {
    int64_t  locus_index;  // ordinal among covered loci
    int32_t  tid;
    int32_t  pos;
    uint32_t target;       // index into the caller's target array
    uint32_t thresh;       // coinToss threshold: rand() < thresh  (ceil((double)(float)af * 2^31))
    uint8_t  base;         // ALT byte of the record
    uint8_t  pad[7];
};

struct CovRun { int32_t tid; int32_t start; int32_t end; int32_t pad; int64_t base; };   // base = ordinal of `start`
