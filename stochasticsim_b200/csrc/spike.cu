// spike.cu -- stochastic spike-in on B200 (hot path 1).
//
// Replaces the per-locus pileup loop of stochasticSpike.c:1129-1623 for SAM text input, as a
// per-read / per-target pipeline (legal because of the facts listed in DESIGN.md section 3):
//
//   parse      one streaming pass over the SAM text -> 64-byte SamRec per line (sam_parse.cuh)
//   keep/sort  read_bam filter flags -> kept ordinals (scan), coordinate-sortedness check
//   mates      per kept read: the next kept read with the same QNAME inside its reference span
//   cover      union of [pos,end) over kept reads -> covered runs, covered-locus ordinals, stats
//   order      stable order of kept reads by (tid,end) = the reference's write order (:1272-1285,:1362-1371)
//   emit       copy every kept line to its output slot
//   targets    which .spike records hit a covered locus / are passed over (max-plus scan of :1578-1619)
//   gather     pileup entries (qpos, base, BQ, mate) of the hit targets, in pileup order
//   rng        glibc rand() stream, generated in parallel by polynomial skip-ahead (rng_glibc.cuh)
//   chain      the inherently serial part: walk the covered loci consuming selectMutantAllele draws
//              (:1197), and at each hit target run attemptToMutateBase (:526-904) over its entries
//   patch      substitute the spiked bases in the emitted text (SEQ only, QUAL untouched)
//
// No CPU fallback: every stage above is a CUDA kernel (CUB device scans/sorts are used for plumbing).
#include "common.cuh"
#include "spike_types.cuh"
#include "rng_glibc.cuh"
#include "sam_parse.cuh"
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>
#include <vector>
#include <time.h>

namespace {

constexpr int MAX_PILEUP = 10000;                 // stochasticSpike.c:38
constexpr int RNG_SEG    = 1024;                  // rand() outputs generated per thread
constexpr int RNG_TPB    = 256;                   // threads per block of the generator
constexpr uint64_t RNG_BLOCK = (uint64_t)RNG_SEG * RNG_TPB;

struct DevErr { int code; int pad; unsigned long long where; };

__device__ __forceinline__ void set_err(DevErr *e, int code, unsigned long long where)
{
    if (atomicCAS(&e->code, 0, code) == 0) e->where = where;
}

// ------------------------------------------------------------------------------------------
// keep flags / sortedness keys
// ------------------------------------------------------------------------------------------
__global__ void flags_kernel(const SamRec *__restrict__ recs, size_t n, uint32_t *__restrict__ keep, unsigned long long *__restrict__ pkey)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t bits = recs[i].bits;
    keep[i] = (bits & REC_KEEP) ? 1u : 0u;
    // bam_plp_push compares (tid, pos) of every pushed read with the running maximum
    pkey[i] = (bits & REC_PUSHED) ? (((unsigned long long)(uint32_t)(recs[i].tid + 1) << 32) | (uint32_t)recs[i].pos) : 0ull;
}

struct MaxOp { template <typename T> __device__ __forceinline__ T operator()(const T &a, const T &b) const { return a > b ? a : b; } };

__global__ void compact_kernel(const SamRec *__restrict__ recs, size_t n, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ kord,
                               const unsigned long long *__restrict__ pkey, const unsigned long long *__restrict__ pmax,
                               uint32_t *__restrict__ k_rec, unsigned long long *__restrict__ k_start, unsigned long long *__restrict__ k_end,
                               uint32_t *__restrict__ k_len, unsigned long long *__restrict__ k_hash, uint32_t *__restrict__ k_hash32, uint8_t *__restrict__ k_bits,
                               unsigned long long *__restrict__ fold, unsigned int *__restrict__ maxspan, DevErr *err)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long span = 0;
    if (i < n) {
        if (pkey[i] && pkey[i] < pmax[i]) set_err(err, SSB_E_UNSORTED, recs[i].line_off);
        if (keep[i]) {
            const SamRec r = recs[i];
            const uint32_t o = kord[i];
            k_rec[o] = (uint32_t)i;
            k_start[o] = ((unsigned long long)(uint32_t)r.tid << 32) | (uint32_t)r.pos;
            k_end[o] = ((unsigned long long)(uint32_t)r.tid << 32) | (uint32_t)r.end;
            k_len[o] = r.line_len + ((r.bits & REC_NO_NL) ? 1u : 0u);
            k_hash[o] = r.qhash; k_hash32[o] = (uint32_t)r.qhash ^ (uint32_t)(r.qhash >> 32); k_bits[o] = r.bits;
            span = (unsigned long long)(r.end - r.pos);
        }
    }
    // totalFoldCoverage = sum of reference spans of kept reads (stochasticSpike.c:1259): one atomic per block
    __shared__ unsigned long long s_sum[8]; __shared__ unsigned int s_max[8];
    unsigned long long s = span;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    unsigned int m = (unsigned int)span;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_sum[w] = s; s_max[w] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int k = 1; k < nw; k++) { s += s_sum[k]; m = max(m, s_max[k]); }
        if (s) atomicAdd(fold, s);
        if (m) atomicMax(maxspan, m);
    }
}

// ------------------------------------------------------------------------------------------
// mates: nxt[o] = first kept read after o (file order) with the same QNAME whose start lies inside
// o's reference span, NO_MATE if none; bit 31 flags "a second such read exists" (then users rescan).
// ------------------------------------------------------------------------------------------
constexpr uint32_t NO_MATE = 0x7fffffffu, MATE_MORE = 0x80000000u, PRV_NONE = 0xffffffffu;

__device__ bool same_qname(const uint8_t *sam, const SamRec &a, const SamRec &b)
{
    if (a.qhash != b.qhash || a.qname_len != b.qname_len) return false;
    const uint8_t *x = sam + a.line_off, *y = sam + b.line_off;
    for (int i = 0; i < a.qname_len; i++) if (x[i] != y[i]) return false;
    return true;
}

__global__ void mates_kernel(const uint8_t *__restrict__ sam, const SamRec *__restrict__ recs, const uint32_t *__restrict__ k_rec,
                             const unsigned long long *__restrict__ k_start, const unsigned long long *__restrict__ k_end,
                             const uint32_t *__restrict__ k_hash32, size_t K, uint32_t *__restrict__ nxt, uint32_t *__restrict__ prv,
                             uint8_t *__restrict__ cplx)
{
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= K) return;
    const unsigned long long lim = k_end[o];       // same tid, pos < end  <=>  start key < end key
    const uint32_t h = k_hash32[o];                // candidates by a 32-bit fold of the QNAME hash; same_qname() decides
    uint32_t first = NO_MATE; bool more = false;
    // reads that start inside this read's span: [o + 1, hi).  The bound is found once (galloping, then bisection), so the
    // scan itself only touches the hashes.
    size_t hi;
    {
        size_t step = 64, lo = o + 1;
        hi = lo;
        while (hi < K && k_start[hi] < lim) { lo = hi + 1; hi += step; step <<= 1; }
        if (hi > K) hi = K;
        while (lo < hi) { const size_t mid = (lo + hi) >> 1; if (k_start[mid] < lim) lo = mid + 1; else hi = mid; }
    }
    for (size_t b0 = o + 1; b0 < hi; b0 += 4) {
      // four candidates per step: the loads do not depend on each other
      uint32_t hb[4];
#pragma unroll
      for (int u = 0; u < 4; u++) hb[u] = (b0 + u < hi) ? k_hash32[b0 + u] : ~h;
      if (hb[0] != h && hb[1] != h && hb[2] != h && hb[3] != h) continue;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const size_t b = b0 + u;
        if (hb[u] != h || b >= hi) continue;
        if (!same_qname(sam, recs[k_rec[o]], recs[k_rec[b]])) continue;
        if (first == NO_MATE) { first = (uint32_t)b; prv[b] = (uint32_t)o; }    // injective: see DESIGN.md (mate links)
        else {
            // three or more same-name reads overlap: every member takes the exact brute-force path
            more = true; cplx[o] = 1; cplx[first] = 1; cplx[b] = 1;
        }
      }
    }
    nxt[o] = first | (more ? MATE_MORE : 0u);
}

// ------------------------------------------------------------------------------------------
// coverage runs
// ------------------------------------------------------------------------------------------
__global__ void runflag_kernel(const unsigned long long *__restrict__ k_start, const unsigned long long *__restrict__ pm, size_t K, uint32_t *__restrict__ flag)
{
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < K) flag[o] = (o == 0 || k_start[o] > pm[o]) ? 1u : 0u;       // new run: other contig, or a gap before this read
}

__global__ void runs_kernel(const unsigned long long *__restrict__ k_start, const unsigned long long *__restrict__ k_end,
                            const unsigned long long *__restrict__ pm, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rid_incl,
                            size_t K, CovRun *__restrict__ runs, unsigned long long *__restrict__ run_len)
{
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= K) return;
    if (flag[o]) {
        uint32_t r = rid_incl[o] - 1;
        runs[r].tid = (int32_t)(k_start[o] >> 32);
        runs[r].start = (int32_t)(uint32_t)k_start[o];
        if (o > 0) runs[r - 1].end = (int32_t)(uint32_t)pm[o];
    }
    if (o == K - 1) {
        // lexicographic max of (tid,end) over all kept reads = (last contig, end of its last run)
        const unsigned long long m = pm[o] > k_end[o] ? pm[o] : k_end[o];
        runs[rid_incl[o] - 1].end = (int32_t)(uint32_t)m;
    }
    (void)run_len;
}

__global__ void runlen_kernel(const CovRun *__restrict__ runs, size_t R, unsigned long long *__restrict__ len)
{
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) len[r] = (unsigned long long)(runs[r].end - runs[r].start);
}
__global__ void runbase_kernel(CovRun *__restrict__ runs, size_t R, const unsigned long long *__restrict__ base)
{
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) { runs[r].base = (int64_t)base[r]; runs[r].pad = 0; }
}

// run index containing covered ordinal g (runs sorted by base)
__device__ __forceinline__ size_t run_of_ordinal(const CovRun *runs, size_t R, int64_t g)
{
    size_t lo = 0, hi = R;              // last run with base <= g
    while (hi - lo > 1) { size_t mid = (lo + hi) >> 1; if (runs[mid].base <= g) lo = mid; else hi = mid; }
    return lo;
}

__device__ __forceinline__ int gcat_index(uint8_t b) { return b == 'G' ? 0 : b == 'C' ? 1 : b == 'A' ? 2 : b == 'T' ? 3 : 4; }

// reference class per covered locus: index into "GCAT" (stochasticSpike.c:340), 4 = anything else
__device__ __forceinline__ uint8_t ref_class(uint8_t c) { return c == 'G' ? 0 : c == 'C' ? 1 : c == 'A' ? 2 : c == 'T' ? 3 : 4; }

__global__ void cls_kernel(const CovRun *__restrict__ runs, size_t R, int64_t n_cov, const uint8_t *const *__restrict__ contig_seq,
                           const int64_t *__restrict__ contig_len, uint8_t *__restrict__ cls, DevErr *err)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_cov) return;
    size_t r = run_of_ordinal(runs, R, g);
    int tid = runs[r].tid; int64_t x = runs[r].start + (g - runs[r].base);
    const uint8_t *seq = contig_seq[tid];
    if (!seq || x >= contig_len[tid]) { set_err(err, SSB_E_REF, (unsigned long long)g); cls[g] = 4; return; }
    cls[g] = ref_class(seq[x]);
}

// ------------------------------------------------------------------------------------------
// output order and emit
// ------------------------------------------------------------------------------------------
__global__ void iota_kernel(uint32_t *p, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = (uint32_t)i; }

// per output line: where it comes from (one 16-byte load in the emit kernel instead of a chain of dependent loads)
struct __align__(16) EmitDesc { unsigned long long src_off; uint32_t len; uint32_t add_nl; };

__global__ void outlen_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ k_len, const uint32_t *__restrict__ k_rec,
                              const SamRec *__restrict__ recs, size_t K, unsigned long long *__restrict__ len, EmitDesc *__restrict__ desc)
{
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= K) return;
    const uint32_t ord = perm[o];
    len[o] = k_len[ord];
    const SamRec &r = recs[k_rec[ord]];
    EmitDesc d; d.src_off = r.line_off; d.len = r.line_len; d.add_nl = (r.bits & REC_NO_NL) ? 1u : 0u;
    desc[o] = d;
}

__global__ void ordoff_kernel(const uint32_t *__restrict__ perm, const unsigned long long *__restrict__ out_off, size_t K, unsigned long long *__restrict__ ord_off)
{
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o < K) ord_off[perm[o]] = out_off[o];
}

// One warp per output line, 16 bytes per lane: destination-aligned 16-byte stores; the source bytes of a chunk lie in two
// aligned 16-byte source words, shifted into place.  The shift is the same for the whole line (= for the whole warp), so the
// word part of it is a four-way switch over statically indexed registers.  Two lines are in flight per warp.
struct EmitJob { const uint8_t *src; uint8_t *dst; uint32_t len, head, nvec, mis, add_nl;
                 __device__ __forceinline__ const uint4 *S() const { return reinterpret_cast<const uint4 *>(src + head - mis); }
                 __device__ __forceinline__ uint4 *D() const { return reinterpret_cast<uint4 *>(dst + head); } };

__device__ __forceinline__ EmitJob emit_job(const uint8_t *__restrict__ sam, const EmitDesc r, unsigned long long off, uint8_t *__restrict__ out)
{
    EmitJob j;
    j.src = sam + r.src_off; j.dst = out + off; j.len = r.len; j.add_nl = r.add_nl;
    uint32_t head = (uint32_t)((16 - ((uintptr_t)j.dst & 15)) & 15);
    if (head > j.len) head = j.len;
    j.head = head;
    j.nvec = (j.len - head) >> 4;
    const uint8_t *s0 = j.src + head;
    j.mis = (uint32_t)((uintptr_t)s0 & 15);
    return j;
}
template <int WQ>
__device__ __forceinline__ uint4 emit_shift(const uint4 A, const uint4 B, uint32_t sh)
{
    const uint32_t W[8] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
    uint4 o;
    o.x = __funnelshift_r(W[WQ], W[WQ + 1], sh); o.y = __funnelshift_r(W[WQ + 1], W[WQ + 2], sh);
    o.z = __funnelshift_r(W[WQ + 2], W[WQ + 3], sh); o.w = __funnelshift_r(W[WQ + 3], W[WQ + 4], sh);
    return o;
}
__device__ __forceinline__ void emit_load(const EmitJob &j, uint32_t i, const uint8_t *sam_end, uint4 &A, uint4 &B)
{
    const uint4 *S = j.S();
    A = __ldg(S + i);
    B = make_uint4(0, 0, 0, 0);
    if (j.mis && reinterpret_cast<const uint8_t *>(S + i + 1) < sam_end) B = __ldg(S + i + 1);
}
__device__ __forceinline__ void emit_store(const EmitJob &j, uint32_t i, const uint4 A, const uint4 B)
{
    const uint32_t sh = (j.mis & 3u) * 8u;
    uint4 o;
    switch (j.mis >> 2) {
    case 0: o = emit_shift<0>(A, B, sh); break;
    case 1: o = emit_shift<1>(A, B, sh); break;
    case 2: o = emit_shift<2>(A, B, sh); break;
    default: o = emit_shift<3>(A, B, sh); break;
    }
    j.D()[i] = o;
}
__device__ __forceinline__ void emit_edges(const EmitJob &j, int lane)
{
    // < 16 bytes up to the first aligned destination address (lanes 0..15) and < 16 tail bytes (lanes 16..31) in one pass
    const uint32_t done = j.head + (j.nvec << 4);
    const uint32_t off = lane < 16 ? (uint32_t)lane : done + (uint32_t)lane - 16u;
    const bool on = lane < 16 ? (uint32_t)lane < j.head : off < j.len;
    if (on) j.dst[off] = j.src[off];
    if (lane == 0 && j.add_nl) j.dst[j.len] = '\n';
}

template <int NL, int MINB>
__global__ void __launch_bounds__(256, MINB)
emit_kernel(const uint8_t *__restrict__ sam, size_t n_sam, const EmitDesc *__restrict__ desc, const unsigned long long *__restrict__ out_off, size_t K,
            uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const uint8_t *sam_end = sam + n_sam;
    for (size_t o = (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); o < K; o += NL * warps) {
        EmitJob j[NL]; bool have[NL], in[NL]; uint4 A[NL], B[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) {
            const size_t ol = o + (size_t)l * warps;
            have[l] = ol < K;
            j[l] = emit_job(sam, desc[have[l] ? ol : o], out_off[have[l] ? ol : o], out);
            in[l] = have[l] && (uint32_t)lane < j[l].nvec;
        }
#pragma unroll
        for (int l = 0; l < NL; l++) if (in[l]) emit_load(j[l], lane, sam_end, A[l], B[l]);
#pragma unroll
        for (int l = 0; l < NL; l++) if (in[l]) emit_store(j[l], lane, A[l], B[l]);
#pragma unroll
        for (int l = 0; l < NL; l++) {
            if (!have[l]) continue;
            for (uint32_t i = lane + 32; i < j[l].nvec; i += 32) { emit_load(j[l], i, sam_end, A[l], B[l]); emit_store(j[l], i, A[l], B[l]); }
            emit_edges(j[l], lane);
        }
    }
}

__global__ void patch_kernel(const Patch *__restrict__ patches, const unsigned int *__restrict__ n_patches, const SamRec *__restrict__ recs,
                             const uint32_t *__restrict__ k_rec, const unsigned long long *__restrict__ ord_off, uint8_t *__restrict__ out)
{
    unsigned int n = *n_patches;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Patch p = patches[i];
        out[ord_off[p.ord] + recs[k_rec[p.ord]].seq_off + p.qpos] = (uint8_t)p.base;
    }
}

// Where an odd patch shares its byte with other patches, the one made at the latest locus must win
// (the reference applies them in locus order); patch_kernel wrote them in no particular order.
struct OddPatch;
__global__ void odd_fix_kernel(const OddPatch *odd_, const unsigned int *n_odd, unsigned int odd_cap, const Patch *__restrict__ patches,
                               const unsigned int *__restrict__ n_patches, const SamRec *__restrict__ recs, const uint32_t *__restrict__ k_rec,
                               const unsigned long long *__restrict__ ord_off, uint8_t *__restrict__ out);

// ------------------------------------------------------------------------------------------
// depth
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t upper_bound_u64(const unsigned long long *a, size_t n, unsigned long long key)
{
    size_t lo = 0, hi = n;                            // first index with a[i] > key
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (a[mid] <= key) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ size_t lower_bound_u64(const unsigned long long *a, size_t n, unsigned long long key)
{
    size_t lo = 0, hi = n;                            // first index with a[i] >= key
    while (lo < hi) { size_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// ------------------------------------------------------------------------------------------
// targets (stochasticSpike.c:1234-1245, :1578-1619, :1630-1646)
// ------------------------------------------------------------------------------------------
struct DevTarget { int32_t c_tid; uint32_t thresh; int64_t locus; uint8_t base; uint8_t pad[7]; };

// first covered ordinal whose (tid,pos) >= (c_tid, locus); n_cov if none
__device__ int64_t lb_ordinal(const CovRun *runs, size_t R, int64_t n_cov, int32_t c_tid, int64_t locus)
{
    if (c_tid < 0) return 0;
    size_t lo = 0, hi = R;                            // first run with (tid, end-1) >= (c_tid, locus)
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        bool before = runs[mid].tid < c_tid || (runs[mid].tid == c_tid && (int64_t)runs[mid].end - 1 < locus);
        if (before) lo = mid + 1; else hi = mid;
    }
    if (lo == R) return n_cov;
    const CovRun r = runs[lo];
    if (r.tid == c_tid && locus > r.start) return r.base + (locus - r.start);
    return r.base;
}

// Per covered locus g: cs[g] = kept reads that start at or before it, ce[g] = kept reads that end at or before it (end exclusive),
// so that the pileup depth is cs[g] - ce[g] without a search.  One pass over the reads in start order and one in end order: the
// last read of a group of equal keys fills the loci up to the next key (never more than one read span).
__global__ void cum_fill_kernel(const unsigned long long *__restrict__ keys, size_t K, const CovRun *__restrict__ runs, size_t R, int64_t n_cov,
                                uint32_t *__restrict__ out)
{
    const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= K) return;
    const unsigned long long key = keys[o];
    if (o + 1 < K && keys[o + 1] == key) return;
    const int64_t g0 = lb_ordinal(runs, R, n_cov, (int32_t)(key >> 32), (int64_t)(uint32_t)key);
    int64_t g1 = n_cov;
    if (o + 1 < K) { const unsigned long long nk = keys[o + 1]; g1 = lb_ordinal(runs, R, n_cov, (int32_t)(nk >> 32), (int64_t)(uint32_t)nk); }
    for (int64_t g = g0; g < g1; g++) out[g] = (uint32_t)(o + 1);
}

// maxDepth (stochasticSpike.c:1216)
__global__ void maxdepth_kernel(const uint32_t *__restrict__ cs, const uint32_t *__restrict__ ce, int64_t n_cov, unsigned int *__restrict__ max_depth)
{
    unsigned int d = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_cov; g += (int64_t)gridDim.x * blockDim.x) d = max(d, cs[g] - ce[g]);
    for (int s = 16; s; s >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, s));
    if ((threadIdx.x & 31) == 0 && d) atomicMax(max_depth, d);
}

__global__ void target_lb_kernel(const DevTarget *__restrict__ tg, size_t T, const CovRun *__restrict__ runs, size_t R, int64_t n_cov, long long *__restrict__ v)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) v[t] = (long long)lb_ordinal(runs, R, n_cov, tg[t].c_tid, tg[t].locus) - (long long)t;
}

// h_t = t + max_{u<=t} (lb_u - u): every target consumes exactly one covered locus (the if-not-while of :1596-1599)
__global__ void target_status_kernel(const DevTarget *__restrict__ tg, size_t T, const long long *__restrict__ vmax, const CovRun *__restrict__ runs, size_t R,
                                     int64_t n_cov, ssb_target_result *__restrict__ res, uint32_t *__restrict__ hitflag)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    ssb_target_result r;
    memset(&r, 0, sizeof r);
    long long h = vmax[t] + (long long)t;
    r.locus_index = h;
    r.rng_offset = -1;
    uint32_t hit = 0;
    if (h >= n_cov) { r.status = SSB_T_TAIL; r.at_tid = -1; r.at_pos = -1; r.locus_index = n_cov; }
    else {
        size_t ri = run_of_ordinal(runs, R, h);
        int tid = runs[ri].tid; int64_t pos = runs[ri].start + (h - runs[ri].base);
        r.at_tid = tid; r.at_pos = pos;
        if (tid == tg[t].c_tid && pos == tg[t].locus) { r.status = SSB_T_HIT; hit = 1; r.filter = SSB_F_UNDETECTED; }
        else r.status = (pos == tg[t].locus) ? SSB_T_NOCOV_SILENT : SSB_T_NOCOV;       // :1603
    }
    res[t] = r;
    hitflag[t] = hit;
}

__global__ void hits_kernel(const DevTarget *__restrict__ tg, size_t T, const uint32_t *__restrict__ hitflag, const uint32_t *__restrict__ hidx,
                            const ssb_target_result *__restrict__ res, HitTarget *__restrict__ hits)
{
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || !hitflag[t]) return;
    HitTarget h;
    memset(&h, 0, sizeof h);
    h.locus_index = res[t].locus_index; h.tid = res[t].at_tid; h.pos = (int32_t)res[t].at_pos;
    h.target = (uint32_t)t; h.thresh = tg[t].thresh; h.base = tg[t].base;
    hits[hidx[t]] = h;
}

// ------------------------------------------------------------------------------------------
// pileup gather at the hit targets
// ------------------------------------------------------------------------------------------
// column geometry of one read at reference position x (htslib resolve_cigar semantics)
__device__ void column_of(const uint8_t *cig, int cig_len, int32_t rpos, int32_t x, uint32_t &qpos, uint8_t &skip)
{
    int64_t ref = rpos; uint32_t y = 0;
    qpos = 0; skip = 0;
    uint32_t num = 0;
    for (int i = 0; i < cig_len; i++) {
        uint8_t c = cig[i];
        if (c >= '0' && c <= '9') { num = num * 10 + (c - '0'); continue; }
        uint32_t l = num; num = 0;
        if (c == 'M' || c == '=' || c == 'X') { if (x < ref + l) { qpos = y + (uint32_t)(x - ref); return; } ref += l; y += l; }
        else if (c == 'D' || c == 'N') { if (x < ref + l) { skip = 1; qpos = y; return; } ref += l; }
        else if (c == 'I' || c == 'S') y += l;
    }
}

// candidates of a locus: kept reads [lo, hi) by start; an entry is a candidate whose end lies beyond the locus
__device__ __forceinline__ void cand_range(const unsigned long long *k_start, size_t K, int32_t tid, int32_t pos, unsigned int maxspan, size_t &lo, size_t &hi)
{
    const unsigned long long key = ((unsigned long long)(uint32_t)tid << 32) | (uint32_t)pos;
    int64_t first = (int64_t)pos - (int64_t)maxspan + 1; if (first < 0) first = 0;
    lo = lower_bound_u64(k_start, K, ((unsigned long long)(uint32_t)tid << 32) | (uint32_t)first);
    hi = upper_bound_u64(k_start, K, key);
}

__global__ void gather_count_kernel(const HitTarget *__restrict__ hits, size_t H, const unsigned long long *__restrict__ k_start,
                                    const unsigned long long *__restrict__ k_end, size_t K, const unsigned int *__restrict__ maxspan,
                                    unsigned long long *__restrict__ cnt, DevErr *err)
{
    const int lane = threadIdx.x & 31;
    size_t h = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (h >= H) return;
    size_t lo, hi; cand_range(k_start, K, hits[h].tid, hits[h].pos, *maxspan, lo, hi);
    const unsigned long long lim = ((unsigned long long)(uint32_t)hits[h].tid << 32) | (uint32_t)hits[h].pos;
    unsigned int c = 0;
    for (size_t i = lo + lane; i < hi; i += 32) c += (k_end[i] > lim) ? 1u : 0u;
    for (int s = 16; s; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if (lane == 0) { cnt[h] = c; if (c > MAX_PILEUP) set_err(err, SSB_E_DEPTH, (unsigned long long)h); }
}

__global__ void gather_fill_kernel(const uint8_t *__restrict__ sam, const SamRec *__restrict__ recs, const uint32_t *__restrict__ k_rec,
                                   const HitTarget *__restrict__ hits, size_t H, const unsigned long long *__restrict__ k_start,
                                   const unsigned long long *__restrict__ k_end, size_t K, const unsigned int *__restrict__ maxspan,
                                   const unsigned long long *__restrict__ eoff, PlpEntry *__restrict__ ent)
{
    const int lane = threadIdx.x & 31;
    size_t h = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (h >= H) return;
    const int32_t x = hits[h].pos;
    size_t lo, hi; cand_range(k_start, K, hits[h].tid, x, *maxspan, lo, hi);
    const unsigned long long lim = ((unsigned long long)(uint32_t)hits[h].tid << 32) | (uint32_t)x;
    unsigned long long w = eoff[h];
    for (size_t base = lo; base < hi; base += 32) {
        size_t i = base + lane;
        bool in = i < hi && k_end[i] > lim;
        unsigned m = __ballot_sync(0xffffffffu, in);
        if (in) {
            const SamRec &r = recs[k_rec[i]];
            PlpEntry e;
            e.ord = (uint32_t)i; e.mate = -1; e.pad = 0;
            const uint8_t *line = sam + r.line_off;
            column_of(line + r.cigar_off, r.cigar_len, r.pos, x, e.qpos, e.skip);
            e.base = line[r.seq_off + e.qpos];
            e.bq = (r.bits & REC_QUALSTAR) ? (uint8_t)0xff : (uint8_t)(line[r.qual_off + e.qpos] - 33);
            ent[w + __popc(m & ((1u << lane) - 1u))] = e;
        }
        w += __popc(m);
    }
}

// mate = index, inside the same pileup, of the first later entry with the same QNAME (findOverlappingMate, :363-383)
__global__ void gather_mate_kernel(const uint8_t *__restrict__ sam, const SamRec *__restrict__ recs, const uint32_t *__restrict__ k_rec,
                                   const unsigned long long *__restrict__ k_start, const unsigned long long *__restrict__ k_end,
                                   const unsigned long long *__restrict__ k_hash, size_t K, const uint32_t *__restrict__ nxt,
                                   const HitTarget *__restrict__ hits, size_t H, const unsigned long long *__restrict__ eoff, PlpEntry *__restrict__ ent)
{
    const int lane = threadIdx.x & 31;
    size_t h = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (h >= H) return;
    const unsigned long long e0 = eoff[h], e1 = eoff[h + 1];
    const unsigned long long lim = ((unsigned long long)(uint32_t)hits[h].tid << 32) | (uint32_t)hits[h].pos;
    for (unsigned long long j = e0 + lane; j < e1; j += 32) {
        const uint32_t a = ent[j].ord;
        uint32_t n1 = nxt[a];
        uint32_t mate_ord = NO_MATE;
        if ((n1 & ~MATE_MORE) != NO_MATE) {
            uint32_t b = n1 & ~MATE_MORE;
            if (k_start[b] <= lim && k_end[b] > lim) mate_ord = b;          // the first same-name read covers the locus
            else if ((n1 & MATE_MORE) && k_start[b] <= lim) {
                // rare: several same-name reads start inside a's span; take the first one that covers the locus
                for (size_t c = b + 1; c < K && k_start[c] <= lim; c++)
                    if (k_hash[c] == k_hash[a] && k_end[c] > lim && same_qname(sam, recs[k_rec[a]], recs[k_rec[c]])) { mate_ord = (uint32_t)c; break; }
            }
        }
        if (mate_ord != NO_MATE) {
            // entries are sorted by ord: binary search inside (j, e1)
            unsigned long long lo = j + 1, hi = e1;
            while (lo < hi) { unsigned long long mid = (lo + hi) >> 1; if (ent[mid].ord < mate_ord) lo = mid + 1; else hi = mid; }
            if (lo < e1 && ent[lo].ord == mate_ord) ent[j].mate = (int32_t)(lo - e0);
        }
    }
}

// ------------------------------------------------------------------------------------------
// "odd" patches.  getBaseWithRPOcheck (:387-432) takes the mate's base at the mate's qpos without
// looking at is_del / is_refskip, and for a deleted/skipped column htslib's qpos is the NEXT
// aligned base.  attemptToMutateBase cases 3-6 can therefore overwrite a base that belongs to a
// LATER locus, and because the reference edits reads in place every later locus sees it.  Such
// patches are rare (a mate inside a D/N at a target that tosses heads); they are kept in a small
// list that the chain and the tally consult so that the result stays bit-exact.
// ------------------------------------------------------------------------------------------
struct OddPatch { uint32_t ord, qpos, base, h; int32_t tid, pos; };

__device__ __forceinline__ unsigned long long odd_bit(uint32_t ord) { return 1ull << ((ord * 2654435761u) >> 26); }

// value of (ord,qpos) as seen at locus (tid,pos): the last odd patch made at an earlier locus wins
__device__ uint8_t odd_view(const OddPatch *odd, unsigned int n_odd, uint32_t ord, uint32_t qpos, int32_t tid, int64_t pos, uint8_t base)
{
    for (unsigned int i = 0; i < n_odd; i++)          // list is in target (= locus) order
        if (odd[i].ord == ord && odd[i].qpos == qpos && (odd[i].tid < tid || (odd[i].tid == tid && odd[i].pos < pos))) base = (uint8_t)odd[i].base;
    return base;
}

__global__ void odd_fix_kernel(const OddPatch *odd, const unsigned int *n_odd, unsigned int odd_cap, const Patch *__restrict__ patches,
                               const unsigned int *__restrict__ n_patches, const SamRec *__restrict__ recs, const uint32_t *__restrict__ k_rec,
                               const unsigned long long *__restrict__ ord_off, uint8_t *__restrict__ out)
{
    const unsigned int no = min(*n_odd, odd_cap), np = *n_patches;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < no; i += gridDim.x * blockDim.x) {
        const OddPatch q = odd[i];
        uint32_t best_h = 0, best_base = q.base; bool any = false;
        for (unsigned int p = 0; p < np; p++)
            if (patches[p].ord == q.ord && patches[p].qpos == q.qpos && (!any || patches[p].pad >= best_h)) { any = true; best_h = patches[p].pad; best_base = patches[p].base; }
        out[ord_off[q.ord] + recs[k_rec[q.ord]].seq_off + q.qpos] = (uint8_t)best_base;
    }
}

// ------------------------------------------------------------------------------------------
// per-locus allele tallies at the loci that are not spike targets (stochasticSpike.c:1296-1357), for
// the SEQ_ERROR lines of truth.vcf (:1494-1557).
//
// ref(x) = depth(x) - minus(x): depth comes from the sorted start/end arrays, minus(x) counts the
// pileup entries at x that do NOT add to refAlleleCnt (deleted/skipped bases, BQ 0, N, mismatches,
// second mates already handled).  Only those exceptional bases cost an atomic.
// err64[x] packs the four error-allele counters (G,C,A,T, 16 bits each; depth <= 10000).
// ------------------------------------------------------------------------------------------
struct ReadView {
    const uint8_t *line; const uint8_t *cig; int cig_len; int32_t pos, end; uint32_t seq_off, qual_off, l_seq; bool qual_star;
    uint32_t ord; int32_t tid; const OddPatch *odd; unsigned int n_odd;      // odd != NULL only for the few reads an odd patch touched
    bool simple;                                                             // CIGAR of M/=/X only: qpos = x - pos
};
__device__ __forceinline__ ReadView view_of(const uint8_t *sam, const SamRec &r, uint32_t ord, const OddPatch *odd, unsigned int n_odd, unsigned long long bloom)
{
    ReadView v; v.line = sam + r.line_off; v.ord = ord; v.tid = r.tid; v.odd = NULL; v.n_odd = 0; v.simple = (r.bits & REC_SIMPLE) != 0;
    if (n_odd && (bloom & odd_bit(ord))) { v.odd = odd; v.n_odd = n_odd; } v.cig = v.line + r.cigar_off; v.cig_len = r.cigar_len; v.pos = r.pos; v.end = r.end;
    v.seq_off = r.seq_off; v.qual_off = r.qual_off; v.l_seq = r.l_seq; v.qual_star = (r.bits & REC_QUALSTAR) != 0;
    return v;
}
__device__ __forceinline__ void base_at(const ReadView &v, int32_t x, uint8_t &base, int &bq, bool &skip)
{
    uint32_t qpos; uint8_t sk;
    if (v.simple) { qpos = (uint32_t)(x - v.pos); sk = 0; }
    else column_of(v.cig, v.cig_len, v.pos, x, qpos, sk);
    if (qpos >= v.l_seq) qpos = v.l_seq - 1;        // malformed CIGAR tail (the reference would read out of bounds)
    base = v.line[v.seq_off + qpos];
    if (v.odd) base = odd_view(v.odd, v.n_odd, v.ord, qpos, v.tid, x, base);
    bq = v.qual_star ? 255 : (int)v.line[v.qual_off + qpos] - 33;
    skip = sk != 0;
}

struct TallyArgs {
    const uint8_t *sam; const SamRec *recs; const uint32_t *k_rec; const unsigned long long *k_start, *k_end, *k_hash; const uint8_t *k_bits; size_t K;
    const uint32_t *nxt, *prv; const uint8_t *cplx; unsigned int maxspan;
    const CovRun *runs; size_t R; const uint8_t *const *contig_seq; const int64_t *contig_len;
    unsigned long long *err64; unsigned int *minus; DevErr *err;
    const OddPatch *odd; const unsigned int *n_odd; const unsigned long long *odd_bloom;
    int listed_only;          // 1: the tokeniser's exception list is in use; fast reads it did not cover take the generic kernel
};

// what entry `self` (index into members) adds at locus x, given the same-name reads that cover x in file order:
// 0 nothing, 1 ref, 2..5 error allele G,C,A,T   (the j-loop of :1255-1357 restricted to one QNAME)
__device__ int chain_contribution(const ReadView *mem, int n, int self, int32_t x, uint8_t F)
{
    bool handled[8];
    for (int i = 0; i < n; i++) handled[i] = false;
    for (int m = 0; m <= self; m++) {
        uint8_t Rb; int rbq; bool sk;
        base_at(mem[m], x, Rb, rbq, sk);
        int contrib = 0;
        if (!(sk || rbq == 0 || handled[m])) {
            uint8_t Mb = 0; int mbq = 0;
            const int mate = m + 1 < n ? m + 1 : -1;                    // first later entry with the same QNAME
            if (mate >= 0 && !handled[mate]) { bool msk; base_at(mem[mate], x, Mb, mbq, msk); }
            if (Mb == 'N') mbq = 0;
            if (Rb == 'N') rbq = 0;
            uint8_t base = Rb;
            if (Mb && Mb != Rb && mbq > rbq) base = Mb;
            if (base == 'N') { handled[m] = true; if (Mb) handled[mate] = true; }
            else if (base == F) contrib = 1;
            else { int gi = gcat_index(base); contrib = gi < 4 ? 2 + gi : 0; handled[m] = true; if (Mb) handled[mate] = true; }
        }
        if (m == self) return contrib;
    }
    return 0;
}

__device__ __forceinline__ void tally_add(const TallyArgs &A, int64_t g, int contrib)
{
    if (contrib == 1) return;
    atomicAdd(&A.minus[g], 1u);
    if (contrib >= 2) atomicAdd(&A.err64[g], 1ull << (16 * (contrib - 2)));
}

// A read takes the fast tally when every query base aligns 1:1 to the reference (CIGAR of M/=/X only), the only reads
// that share its QNAME inside its span are its direct neighbours in the mate chain and are just as simple, and no odd
// patch touched any of them.
__device__ __forceinline__ bool tally_simple_read(const TallyArgs &A, size_t o, unsigned int n_odd, unsigned long long bloom)
{
    if (!(A.k_bits[o] & REC_SIMPLE) || A.cplx[o]) return false;
    if (n_odd && (bloom & odd_bit((uint32_t)o))) return false;
    return true;
}
__device__ __forceinline__ bool tally_is_fast(const TallyArgs &A, size_t o, const SamRec &r, unsigned int n_odd, unsigned long long bloom)
{
    if (!tally_simple_read(A, o, n_odd, bloom)) return false;
    const uint32_t p = A.prv[o], q = A.nxt[o] & ~MATE_MORE;
    if (p != PRV_NONE && !tally_simple_read(A, p, n_odd, bloom)) return false;
    if (q != NO_MATE && !tally_simple_read(A, q, n_odd, bloom)) return false;
    (void)r;
    return true;
}

// List mode: a read is settled from the tokeniser's exception list iff it is fast-eligible, the tokeniser listed its own
// exceptional bases, and it listed those of its mate-chain neighbours too (a plain base only changes its contribution when a
// neighbour's base over the same position is exceptional, and that entry is what triggers the re-evaluation).  Every other
// read is self-contained work of tally_kernel.
__device__ __forceinline__ bool tally_is_listed(const TallyArgs &A, size_t o, const SamRec &r, unsigned int n_odd, unsigned long long bloom)
{
    if (!(A.k_bits[o] & REC_EXC_DONE) || !tally_is_fast(A, o, r, n_odd, bloom)) return false;
    const uint32_t p = A.prv[o], q = A.nxt[o] & ~MATE_MORE;
    if (p != PRV_NONE && !(A.k_bits[p] & REC_EXC_DONE)) return false;
    if (q != NO_MATE && !(A.k_bits[q] & REC_EXC_DONE)) return false;
    return true;
}

__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t *p)
{
    const uint32_t a = (uint32_t)((uintptr_t)p & 3u);
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p - a);
    return __funnelshift_r(__ldg(q), __ldg(q + 1), a * 8);
}

// 0x80 flags of the bytes of a simple read that are not plain "reference base with BQ > 0" at query offset w..w+3
__device__ __forceinline__ uint32_t exc_word(const uint8_t *seq, const uint8_t *qual, bool qstar, uint32_t w, uint32_t refw)
{
    const uint32_t sq = ldg_u32_unaligned(seq + w);
    const uint32_t ql = qstar ? 0x7e7e7e7eu : ldg_u32_unaligned(qual + w);
    return (~samparse::zero_bytes(sq ^ refw) | samparse::zero_bytes(ql ^ 0x21212121u)) & 0x80808080u;
}

// the rare path of the fast tally: one exceptional base of read o at reference position x, evaluated exactly
__device__ __noinline__ void tally_exceptional_base(const TallyArgs &A, uint32_t o, int32_t x, int64_t g)
{
    const SamRec &r = A.recs[A.k_rec[o]];
    ReadView mem[3]; int n = 0, self = 0;
    const uint32_t p = A.prv[o], q = A.nxt[o] & ~MATE_MORE;
    if (p != PRV_NONE && (uint32_t)A.k_end[p] > (uint32_t)r.pos) mem[n++] = view_of(A.sam, A.recs[A.k_rec[p]], p, NULL, 0, 0);
    self = n; mem[n++] = view_of(A.sam, r, o, NULL, 0, 0);
    if (q != NO_MATE) mem[n++] = view_of(A.sam, A.recs[A.k_rec[q]], q, NULL, 0, 0);
    ReadView cov[3]; int nc = 0, sc = 0;
    for (int m = 0; m < n; m++) if (mem[m].pos <= x && x < mem[m].end) { if (m == self) sc = nc; cov[nc++] = mem[m]; }
    const uint8_t F = A.contig_seq[r.tid][x];
    tally_add(A, g, chain_contribution(cov, nc, sc, x, F));
}

// Compact per-read metadata for the fast tally (one 32-byte load instead of a chain of dependent loads)
struct __align__(16) KMeta { unsigned long long line_off; uint32_t seq_off, qual_off, l_seq; int32_t pos, end; uint32_t tid_bits; };
static_assert(sizeof(KMeta) == 32, "KMeta layout");

__global__ void kmeta_kernel(const SamRec *__restrict__ recs, const uint32_t *__restrict__ k_rec, size_t K, KMeta *__restrict__ km)
{
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= K) return;
    const SamRec &r = recs[k_rec[o]];
    KMeta m; m.line_off = r.line_off; m.seq_off = r.seq_off; m.qual_off = r.qual_off; m.l_seq = r.l_seq; m.pos = r.pos; m.end = r.end;
    m.tid_bits = ((uint32_t)r.tid << 8) | r.bits;
    km[o] = m;
}

// One warp per read: SEQ, QUAL and the reference compared four bytes per lane; only bytes that differ from the
// reference or have BQ 0 -- in this read or in an overlapping mate -- are evaluated one by one.
__global__ void __launch_bounds__(128, 8)
tally_fast_kernel(TallyArgs A, const KMeta *__restrict__ km)
{
    const int lane = threadIdx.x & 31;
    const size_t o = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= A.K) return;
    const unsigned int n_odd = *A.n_odd; const unsigned long long bloom = *A.odd_bloom;
    const KMeta me = km[o];
    const uint32_t p = A.prv[o], q = A.nxt[o] & ~MATE_MORE;
    auto simple_ok = [&](uint32_t idx, uint32_t tid_bits) {
        return (tid_bits & REC_SIMPLE) && !A.cplx[idx] && !(n_odd && (bloom & odd_bit(idx)));
    };
    if (!simple_ok((uint32_t)o, me.tid_bits)) return;
    KMeta mp, mq; mp.pos = 0; mp.end = 0; mq.pos = 0; mq.end = 0;
    bool hp = false, hq = false;
    if (p != PRV_NONE) { mp = km[p]; if (!simple_ok(p, mp.tid_bits)) return; hp = mp.end > me.pos; }
    if (q != NO_MATE) { mq = km[q]; if (!simple_ok(q, mq.tid_bits)) return; hq = true; }
    const int tid = (int)(me.tid_bits >> 8);
    const uint8_t *ref = A.contig_seq[tid];
    if (!ref || (int64_t)me.end > A.contig_len[tid]) { if (lane == 0) set_err(A.err, SSB_E_REF, me.line_off); return; }
    size_t ri; { size_t lo = 0, hi = A.R; while (hi - lo > 1) { size_t mid = (lo + hi) >> 1;
                   bool le = A.runs[mid].tid < tid || (A.runs[mid].tid == tid && A.runs[mid].start <= me.pos); if (le) lo = mid; else hi = mid; } ri = lo; }
    const int64_t gbase = A.runs[ri].base - A.runs[ri].start + me.pos;    // covered ordinal of query offset 0
    const uint8_t *seq = A.sam + me.line_off + me.seq_off, *qual = A.sam + me.line_off + me.qual_off, *rf = ref + me.pos;
    const bool qstar = (me.tid_bits & REC_QUALSTAR) != 0;
    const uint32_t L = me.l_seq;
    for (uint32_t w = lane * 4; w < L; w += 128) {
        const uint32_t rw = ldg_u32_unaligned(rf + w);
        uint32_t exc = exc_word(seq, qual, qstar, w, rw);
        const int32_t x0 = me.pos + (int32_t)w;
        if (hp && !(x0 + 3 < mp.pos || x0 >= mp.end)) {
            if (x0 >= mp.pos && x0 + 4 <= mp.end)
                exc |= exc_word(A.sam + mp.line_off + mp.seq_off, A.sam + mp.line_off + mp.qual_off, (mp.tid_bits & REC_QUALSTAR) != 0, (uint32_t)(x0 - mp.pos), rw);
            else exc = 0x80808080u;                                    // the mate starts or ends inside this word: look at every byte
        }
        if (hq && !(x0 + 3 < mq.pos || x0 >= mq.end)) {
            if (x0 >= mq.pos && x0 + 4 <= mq.end)
                exc |= exc_word(A.sam + mq.line_off + mq.seq_off, A.sam + mq.line_off + mq.qual_off, (mq.tid_bits & REC_QUALSTAR) != 0, (uint32_t)(x0 - mq.pos), rw);
            else exc = 0x80808080u;
        }
        const uint32_t rem = L - w;
        if (rem < 4) exc &= (1u << (8 * rem)) - 1u;
        while (exc) {
            const int k = (__ffs(exc) - 1) >> 3; exc &= exc - 1;
            tally_exceptional_base(A, (uint32_t)o, x0 + k, gbase + w + k);
        }
    }
}

// One thread per exceptional base listed by the tokeniser: the base itself, and any mate lying over the same position
// whose own base is plain (so it has no entry of its own) but which this base may have marked handled.
__global__ void __launch_bounds__(128)
tally_resolve_kernel(TallyArgs A, const unsigned long long *__restrict__ list, unsigned long long n_list, const uint32_t *__restrict__ keep,
                     const uint32_t *__restrict__ kord)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_list) return;
    const unsigned long long e = list[i];
    const size_t line = (size_t)(e >> 16);
    if (!keep[line]) return;
    const uint32_t o = kord[line];
    const SamRec &r = A.recs[A.k_rec[o]];
    const unsigned int n_odd = *A.n_odd; const unsigned long long bloom = *A.odd_bloom;
    const bool own = tally_is_listed(A, o, r, n_odd, bloom);           // else the read itself is tally_kernel's work
    const int32_t x = r.pos + (int32_t)(e & 0xffff);
    size_t ri; { size_t lo = 0, hi = A.R; while (hi - lo > 1) { size_t mid = (lo + hi) >> 1;
                   bool le = A.runs[mid].tid < r.tid || (A.runs[mid].tid == r.tid && A.runs[mid].start <= r.pos); if (le) lo = mid; else hi = mid; } ri = lo; }
    const int64_t g = A.runs[ri].base + (x - A.runs[ri].start);
    if (A.cplx[o]) return;                                             // three or more same-name reads overlap: all of them are tally_kernel's
    ReadView mem[3]; uint32_t ords[3]; int n = 0, self = 0;
    const uint32_t p = A.prv[o], q = A.nxt[o] & ~MATE_MORE;
    if (p != PRV_NONE && (uint32_t)A.k_end[p] > (uint32_t)r.pos) { ords[n] = p; mem[n++] = view_of(A.sam, A.recs[A.k_rec[p]], p, NULL, 0, 0); }
    self = n; ords[n] = o; mem[n++] = view_of(A.sam, r, o, NULL, 0, 0);
    if (q != NO_MATE) { ords[n] = q; mem[n++] = view_of(A.sam, A.recs[A.k_rec[q]], q, NULL, 0, 0); }
    ReadView cov[3]; uint32_t cord[3]; int nc = 0, sc = 0;
    for (int m = 0; m < n; m++) if (mem[m].pos <= x && x < mem[m].end) { if (m == self) sc = nc; cord[nc] = ords[m]; cov[nc++] = mem[m]; }
    const uint8_t F = A.contig_seq[r.tid][x];
    if (own) tally_add(A, g, chain_contribution(cov, nc, sc, x, F));
    for (int m = 0; m < nc; m++) {
        if (m == sc) continue;
        if (!(A.k_bits[cord[m]] & REC_SIMPLE)) continue;                 // a neighbour with indels is never list-settled
        uint8_t b; int bq; bool sk;
        base_at(cov[m], x, b, bq, sk);
        // a plain base of a list-settled neighbour has no entry of its own: it matters only if it was marked handled here
        if (b == F && bq != 0 && tally_is_listed(A, cord[m], A.recs[A.k_rec[cord[m]]], n_odd, bloom)) tally_add(A, g, chain_contribution(cov, nc, m, x, F));
    }
}

__global__ void __launch_bounds__(128)
tally_kernel(TallyArgs A)
{
    size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= A.K) return;
    const SamRec &r = A.recs[A.k_rec[o]];
    const unsigned int n_odd = *A.n_odd; const unsigned long long bloom = *A.odd_bloom;
    if (A.listed_only ? tally_is_listed(A, o, r, n_odd, bloom) : tally_is_fast(A, o, r, n_odd, bloom)) return;   // settled by tally_resolve_kernel / tally_fast_kernel
    const ReadView me = view_of(A.sam, r, (uint32_t)o, A.odd, n_odd, bloom);
    const int tid = r.tid;
    const uint8_t *ref = A.contig_seq[tid];
    if (!ref || (int64_t)r.end > A.contig_len[tid]) { set_err(A.err, SSB_E_REF, r.line_off); return; }   // :1181 mplp_get_ref / read past the contig
    // same-name reads that overlap this one, in file order
    ReadView mem[8]; int n = 0, self = 0; bool use_chain = false;
    if (A.cplx[o]) {
        // exact brute force: every kept read of this contig that can overlap [pos,end) and carries the same QNAME
        const unsigned long long h = A.k_hash[o];
        int64_t first = (int64_t)r.pos - (int64_t)A.maxspan + 1; if (first < 0) first = 0;
        size_t lo = lower_bound_u64(A.k_start, A.K, ((unsigned long long)(uint32_t)tid << 32) | (uint32_t)first);
        const unsigned long long lim = A.k_end[o];
        for (size_t b = lo; b < A.K && A.k_start[b] < lim && n < 8; b++) {
            if (b != o && (A.k_hash[b] != h || !same_qname(A.sam, r, A.recs[A.k_rec[b]]))) continue;
            if (b != o && (uint32_t)(A.k_end[b]) <= (uint32_t)r.pos) continue;
            if (b == o) self = n;
            mem[n++] = view_of(A.sam, A.recs[A.k_rec[b]], (uint32_t)b, A.odd, n_odd, bloom);
        }
        use_chain = n > 1;
    } else {
        const uint32_t p = A.prv[o], q = A.nxt[o] & ~MATE_MORE;
        if (p != PRV_NONE && (uint32_t)A.k_end[p] > (uint32_t)r.pos) mem[n++] = view_of(A.sam, A.recs[A.k_rec[p]], p, A.odd, n_odd, bloom);
        self = n; mem[n++] = me;
        if (q != NO_MATE) mem[n++] = view_of(A.sam, A.recs[A.k_rec[q]], q, A.odd, n_odd, bloom);
        use_chain = n > 1;
    }
    // covered ordinal of this read's first base: reads never span two runs
    size_t ri; { size_t lo = 0, hi = A.R; while (hi - lo > 1) { size_t mid = (lo + hi) >> 1;
                   bool le = A.runs[mid].tid < tid || (A.runs[mid].tid == tid && A.runs[mid].start <= r.pos); if (le) lo = mid; else hi = mid; } ri = lo; }
    const int64_t gbase = A.runs[ri].base - A.runs[ri].start;           // g = gbase + x
    // walk the CIGAR
    int64_t x = r.pos; uint32_t y = 0, num = 0;
    for (int ci = 0; ci < me.cig_len; ci++) {
        const uint8_t c = me.cig[ci];
        if (c >= '0' && c <= '9') { num = num * 10 + (c - '0'); continue; }
        const uint32_t l = num; num = 0;
        if (c == 'M' || c == '=' || c == 'X') {
            for (uint32_t t = 0; t < l; t++, x++, y++) {
                const uint8_t F = ref[x];
                int contrib;
                bool overlapped = false;
                if (use_chain) for (int m = 0; m < n; m++) if (m != self && mem[m].pos <= x && x < mem[m].end) { overlapped = true; break; }
                if (overlapped) {
                    ReadView cov[8]; int nc = 0, sc = 0;
                    for (int m = 0; m < n; m++) if (mem[m].pos <= x && x < mem[m].end) { if (m == self) sc = nc; cov[nc++] = mem[m]; }
                    contrib = chain_contribution(cov, nc, sc, (int32_t)x, F);
                } else {
                    uint8_t b = me.line[me.seq_off + y];
                    if (me.odd) b = odd_view(me.odd, me.n_odd, me.ord, y, tid, x, b);
                    const int bq = me.qual_star ? 255 : (int)me.line[me.qual_off + y] - 33;
                    if (bq == 0 || b == 'N') contrib = 0;
                    else if (b == F) contrib = 1;
                    else { int gi = gcat_index(b); contrib = gi < 4 ? 2 + gi : 0; }
                }
                tally_add(A, gbase + x, contrib);
            }
        } else if (c == 'D' || c == 'N') {
            for (uint32_t t = 0; t < l; t++, x++) atomicAdd(&A.minus[gbase + x], 1u);      // in the pileup, adds to nothing
        } else if (c == 'I' || c == 'S') y += l;
    }
}

__global__ void tally_clear_hits_kernel(const HitTarget *__restrict__ hits, size_t H, unsigned long long *__restrict__ err64)
{
    size_t h = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h < H) err64[hits[h].locus_index] = 0;          // a target locus prints its own line (:1406-1470), never SEQ_ERROR
}

__global__ void tally_flag_kernel(const unsigned long long *__restrict__ err64, int64_t n_cov, uint32_t *__restrict__ flag)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_cov) flag[g] = err64[g] ? 1u : 0u;
}

__global__ void tally_emit_kernel(const unsigned long long *__restrict__ err64, const unsigned int *__restrict__ minus, const uint32_t *__restrict__ flag,
                                  const uint32_t *__restrict__ idx, int64_t n_cov, const CovRun *__restrict__ runs, size_t R,
                                  const uint32_t *__restrict__ cum_s, const uint32_t *__restrict__ cum_e,
                                  const uint8_t *const *__restrict__ contig_seq, ssb_seq_error *__restrict__ out)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_cov || !flag[g]) return;
    size_t ri = run_of_ordinal(runs, R, g);
    const int tid = runs[ri].tid; const int64_t x = runs[ri].start + (g - runs[ri].base);
    const long long depth = (long long)cum_s[g] - (long long)cum_e[g];
    ssb_seq_error e;
    memset(&e, 0, sizeof e);
    e.tid = tid; e.pos = x; e.locus_index = g;
    e.ref_cnt = (int32_t)(depth - (long long)minus[g]);
    const unsigned long long v = err64[g];
    for (int i = 0; i < 4; i++) e.err_cnt[i] = (int32_t)((v >> (16 * i)) & 0xffff);
    e.ref_base = contig_seq[tid][x];
    out[idx[g]] = e;
}

#include "spike_chain.cuh"

// ------------------------------------------------------------------------------------------
// device arena: one stream-ordered allocation per array, all released at the end of the run
// ------------------------------------------------------------------------------------------
struct Arena {
    ssb_ctx *ctx; cudaStream_t s; std::vector<void *> ptrs; bool failed = false;
    template <typename T> T *get(size_t n)
    {
        void *p = NULL;
        if (cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), s) != cudaSuccess) { failed = true; cudaGetLastError(); return NULL; }
        ptrs.push_back(p);
        return (T *)p;
    }
    ~Arena() { for (void *p : ptrs) cudaFreeAsync(p, s); }
};

inline int grid_for(size_t n, int block) { size_t g = (n + block - 1) / block; return (int)(g ? g : 1); }

} // namespace

// ---------------------------------------------------------------------------------------------
struct ssb_spike {
    ssb_ctx *ctx;
    int n_contigs;
    char *d_names; uint32_t *d_name_off;
    uint8_t **d_seq_ptrs; int64_t *d_lens;
    std::vector<uint8_t *> d_seqs;
    RngTables *d_rng_tab;
    ssb_seq_error *d_se; size_t n_se;          // SEQ_ERROR records of the last run (device resident until asked for)
};

extern "C" int ssb_spike_create(ssb_ctx *ctx, const ssb_contig *contigs, int n_contigs, ssb_spike **out)
{
    if (!ctx || !out || n_contigs < 0 || (n_contigs && !contigs)) return SSB_E_ARG;
    *out = NULL;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    ssb_spike *sp = new ssb_spike();
    sp->ctx = ctx; sp->n_contigs = n_contigs; sp->d_se = NULL; sp->n_se = 0;
    std::vector<char> names; std::vector<uint32_t> off; std::vector<int64_t> lens; std::vector<uint8_t *> ptrs;
    off.push_back(0);
    for (int i = 0; i < n_contigs; i++) {
        const char *nm = contigs[i].name ? contigs[i].name : "";
        names.insert(names.end(), nm, nm + strlen(nm));
        off.push_back((uint32_t)names.size());
        uint8_t *d = NULL;
        if (contigs[i].seq && contigs[i].len > 0) {
            SSB_CUDA(ctx, cudaMalloc(&d, (size_t)contigs[i].len + 64));
            SSB_CUDA(ctx, cudaMemcpy(d, contigs[i].seq, (size_t)contigs[i].len, cudaMemcpyHostToDevice));
        }
        sp->d_seqs.push_back(d); ptrs.push_back(d); lens.push_back(d ? contigs[i].len : 0);
    }
    SSB_CUDA(ctx, cudaMalloc(&sp->d_names, names.size() + 1));
    SSB_CUDA(ctx, cudaMalloc(&sp->d_name_off, off.size() * sizeof(uint32_t)));
    SSB_CUDA(ctx, cudaMalloc(&sp->d_seq_ptrs, (ptrs.size() + 1) * sizeof(uint8_t *)));
    SSB_CUDA(ctx, cudaMalloc(&sp->d_lens, (lens.size() + 1) * sizeof(int64_t)));
    if (!names.empty()) SSB_CUDA(ctx, cudaMemcpy(sp->d_names, names.data(), names.size(), cudaMemcpyHostToDevice));
    SSB_CUDA(ctx, cudaMemcpy(sp->d_name_off, off.data(), off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    if (n_contigs) {
        SSB_CUDA(ctx, cudaMemcpy(sp->d_seq_ptrs, ptrs.data(), ptrs.size() * sizeof(uint8_t *), cudaMemcpyHostToDevice));
        SSB_CUDA(ctx, cudaMemcpy(sp->d_lens, lens.data(), lens.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    }
    // seed-independent skip-ahead table: x^(t*RNG_SEG) mod P, t < RNG_TPB (host integer arithmetic, a few ms)
    RngTables *tab = new RngTables();
    uint32_t step[GLIBC_DEG], cur[GLIBC_DEG], tmp[GLIBC_DEG];
    glibc_poly_xpow(RNG_SEG, step);
    glibc_poly_xpow(0, cur);
    for (int t = 0; t < RNG_TPB; t++) {
        memcpy(tab->seg[t], cur, sizeof cur);
        glibc_poly_mulmod(cur, step, tmp); memcpy(cur, tmp, sizeof cur);
    }
    SSB_CUDA(ctx, cudaMalloc(&sp->d_rng_tab, sizeof(RngTables)));
    SSB_CUDA(ctx, cudaMemcpy(sp->d_rng_tab, tab, sizeof(RngTables), cudaMemcpyHostToDevice));
    delete tab;
    *out = sp;
    return SSB_OK;
}

extern "C" void ssb_spike_destroy(ssb_spike *sp)
{
    if (!sp) return;
    cudaSetDevice(sp->ctx->device);
    cudaStreamSynchronize(sp->ctx->stream);
    for (uint8_t *d : sp->d_seqs) if (d) cudaFree(d);
    if (sp->d_se) cudaFree(sp->d_se);
    cudaFree(sp->d_names); cudaFree(sp->d_name_off); cudaFree(sp->d_seq_ptrs); cudaFree(sp->d_lens); cudaFree(sp->d_rng_tab);
    delete sp;
}

#define SPK_CHECK_ARENA(ar) do { if ((ar).failed) { snprintf(ctx->err, sizeof ctx->err, "spike: device allocation failed"); return SSB_E_NOMEM; } } while (0)

namespace {

struct Ev { cudaEvent_t e; };

// CUB wrappers: temp storage from the arena
template <typename T> int scan_sum(Arena &ar, ssb_ctx *ctx, const T *in, T *out, size_t n)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(NULL, bytes, in, out, n, ar.s);
    void *tmp = ar.get<uint8_t>(bytes); SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, ar.s));
    ctx->launches += 2;
    return SSB_OK;
}
template <typename T> int scan_sum_incl(Arena &ar, ssb_ctx *ctx, const T *in, T *out, size_t n)
{
    size_t bytes = 0;
    cub::DeviceScan::InclusiveSum(NULL, bytes, in, out, n, ar.s);
    void *tmp = ar.get<uint8_t>(bytes); SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cub::DeviceScan::InclusiveSum(tmp, bytes, in, out, n, ar.s));
    ctx->launches += 2;
    return SSB_OK;
}
template <typename T> int scan_max_excl(Arena &ar, ssb_ctx *ctx, const T *in, T *out, size_t n, T init)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveScan(NULL, bytes, in, out, MaxOp(), init, n, ar.s);
    void *tmp = ar.get<uint8_t>(bytes); SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cub::DeviceScan::ExclusiveScan(tmp, bytes, in, out, MaxOp(), init, n, ar.s));
    ctx->launches += 2;
    return SSB_OK;
}
template <typename T> int scan_max_incl(Arena &ar, ssb_ctx *ctx, const T *in, T *out, size_t n)
{
    size_t bytes = 0;
    cub::DeviceScan::InclusiveScan(NULL, bytes, in, out, MaxOp(), n, ar.s);
    void *tmp = ar.get<uint8_t>(bytes); SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cub::DeviceScan::InclusiveScan(tmp, bytes, in, out, MaxOp(), n, ar.s));
    ctx->launches += 2;
    return SSB_OK;
}

int dev_error(ssb_ctx *ctx, cudaStream_t s, DevErr *d_err, const char *stage)
{
    DevErr e;
    SSB_CUDA(ctx, cudaMemcpyAsync(&e, d_err, sizeof e, cudaMemcpyDeviceToHost, s));
    SSB_CUDA(ctx, cudaStreamSynchronize(s));
    if (e.code) { snprintf(ctx->err, sizeof ctx->err, "spike/%s: %s (at %llu)", stage, ssb_strerror(e.code), e.where); return e.code; }
    return SSB_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

} // namespace

extern "C" int ssb_spike_run_device(ssb_spike *sp, const uint8_t *d_sam, size_t n, uint8_t *d_out, size_t out_cap,
                                    const ssb_target *targets, size_t n_targets, unsigned seed,
                                    ssb_target_result *results, ssb_spike_stats *stats, size_t *out_bytes)
{
    if (!sp || (!d_sam && n) || (!d_out && n) || (n_targets && (!targets || !results)) || !stats || !out_bytes) return SSB_E_ARG;
    if (((uintptr_t)d_sam & 15) != 0) return SSB_E_ARG;
    ssb_ctx *ctx = sp->ctx;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    memset(stats, 0, sizeof *stats);
    *out_bytes = 0;
    const bool dbg_t = getenv("SSB_CHAIN_DEBUG") != NULL;
    const bool dbg_sync = dbg_t && getenv("SSB_CHAIN_DEBUG")[0] == '1';
    auto dbg_mark = [&](const char *what) { if (dbg_t) { if (dbg_sync) cudaStreamSynchronize(s); struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); fprintf(stderr, "[tally] %-12s %.3f ms\n", what, ts.tv_sec * 1e3 + ts.tv_nsec / 1e6); } };

    if (sp->d_se) { cudaFreeAsync(sp->d_se, ctx->stream); sp->d_se = NULL; }
    sp->n_se = 0;
    stats->in_bytes = (int64_t)n;
    const size_t T = n_targets;
    cudaEvent_t ev[12];
    for (int i = 0; i < 12; i++) SSB_CUDA(ctx, cudaEventCreate(&ev[i]));
    struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 12; i++) cudaEventDestroy(e[i]); } } evg{ev};
    Arena ar; ar.ctx = ctx; ar.s = s;

    DevErr *d_err = ar.get<DevErr>(1); SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(DevErr), s));
    const unsigned int odd_cap = 1u << 16;
    OddPatch *d_odd = ar.get<OddPatch>(odd_cap); unsigned int *d_nodd = ar.get<unsigned int>(1); unsigned long long *d_bloom = ar.get<unsigned long long>(1);
    SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cudaMemsetAsync(d_nodd, 0, sizeof(unsigned int), s));
    SSB_CUDA(ctx, cudaMemsetAsync(d_bloom, 0, sizeof(unsigned long long), s));
    SSB_CUDA(ctx, cudaEventRecord(ev[0], s));

    // ---------------------------------------------------------------- parse
    size_t N = 0;
    SamRec *recs = NULL;
    unsigned long long *exc_list = NULL, *d_exc_count = NULL, exc_cap = 0;
    uint32_t *keep = NULL, *kord = NULL;
    if (n) {
        const size_t n_tiles = (n + samparse::TILE - 1) / samparse::TILE;
        unsigned long long *tile_state = ar.get<unsigned long long>(n_tiles);
        unsigned int *ticket = ar.get<unsigned int>(1);
        unsigned long long *d_nlines = ar.get<unsigned long long>(1);
        SPK_CHECK_ARENA(ar);
        size_t rec_cap = n / 96 + 4096;
        exc_cap = n / 96 + 65536;
        exc_list = ar.get<unsigned long long>(exc_cap); d_exc_count = ar.get<unsigned long long>(1); SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaFuncSetAttribute(samparse::parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, samparse::SMEM_BYTES));
        for (int attempt = 0; attempt < 2; attempt++) {
            recs = ar.get<SamRec>(rec_cap); SPK_CHECK_ARENA(ar);
            SSB_CUDA(ctx, cudaMemsetAsync(tile_state, 0, n_tiles * sizeof(unsigned long long), s));
            SSB_CUDA(ctx, cudaMemsetAsync(ticket, 0, sizeof(unsigned int), s));
            SSB_CUDA(ctx, cudaMemsetAsync(d_nlines, 0, sizeof(unsigned long long), s));
            SSB_CUDA(ctx, cudaMemsetAsync(d_err, 0, sizeof(DevErr), s));
            SSB_CUDA(ctx, cudaMemsetAsync(d_exc_count, 0, sizeof(unsigned long long), s));
            samparse::ContigNames names{sp->d_names, sp->d_name_off, sp->n_contigs, (const uint8_t *const *)sp->d_seq_ptrs, sp->d_lens,
                                        getenv("SSB_NO_EXC_LIST") ? NULL : exc_list, d_exc_count, exc_cap};
            int occ = 1;
            SSB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, samparse::parse_kernel, samparse::THREADS, samparse::SMEM_BYTES));
            if (occ < 1) occ = 1;
            int grid = (int)(n_tiles < (size_t)ctx->sm_count * occ ? n_tiles : (size_t)ctx->sm_count * occ);     // persistent: exactly the resident blocks
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_PARSE, samparse::parse_kernel, grid, samparse::THREADS, samparse::SMEM_BYTES, s,
                         d_sam, n, names, recs, rec_cap, tile_state, ticket, d_nlines, reinterpret_cast<SpikeErr *>(d_err));
            unsigned long long nl = 0; DevErr e;
            SSB_CUDA(ctx, cudaMemcpyAsync(&nl, d_nlines, sizeof nl, cudaMemcpyDeviceToHost, s));
            SSB_CUDA(ctx, cudaMemcpyAsync(&e, d_err, sizeof e, cudaMemcpyDeviceToHost, s));
            SSB_CUDA(ctx, cudaStreamSynchronize(s));
            N = (size_t)nl;
            if (e.code == SSB_E_NOMEM && attempt == 0) { rec_cap = N + 16; continue; }
            if (e.code) { snprintf(ctx->err, sizeof ctx->err, "spike/parse: %s (SAM body offset %llu)", ssb_strerror(e.code), e.where); return e.code; }
            break;
        }
    }
    stats->n_lines = (int64_t)N;
    SSB_CUDA(ctx, cudaEventRecord(ev[1], s));

    // ---------------------------------------------------------------- keep / sortedness / compaction
    size_t K = 0;
    uint32_t *k_rec = NULL, *k_len = NULL, *nxt = NULL, *prv = NULL; uint8_t *cplx = NULL; unsigned long long *err64 = NULL; unsigned int *minus = NULL; unsigned long long *k_start = NULL, *k_end = NULL, *k_hash = NULL; uint32_t *k_hash32 = NULL; uint8_t *k_bits = NULL;
    unsigned long long *d_fold = ar.get<unsigned long long>(1); unsigned int *d_maxspan = ar.get<unsigned int>(1), *d_maxdepth = ar.get<unsigned int>(1);
    SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cudaMemsetAsync(d_fold, 0, sizeof(unsigned long long), s));
    SSB_CUDA(ctx, cudaMemsetAsync(d_maxspan, 0, sizeof(unsigned int), s));
    SSB_CUDA(ctx, cudaMemsetAsync(d_maxdepth, 0, sizeof(unsigned int), s));
    if (N) {
        keep = ar.get<uint32_t>(N); kord = ar.get<uint32_t>(N);
        unsigned long long *pkey = ar.get<unsigned long long>(N), *pmax = ar.get<unsigned long long>(N);
        SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, flags_kernel, grid_for(N, 256), 256, 0, s, recs, N, keep, pkey);
        int rc;
        if ((rc = scan_sum(ar, ctx, keep, kord, N))) return rc;
        if ((rc = scan_max_excl(ar, ctx, pkey, pmax, N, 0ull))) return rc;
        uint32_t last_ord = 0, last_keep = 0;
        SSB_CUDA(ctx, cudaMemcpyAsync(&last_ord, kord + N - 1, 4, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(&last_keep, keep + N - 1, 4, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        K = (size_t)last_ord + last_keep;
        k_rec = ar.get<uint32_t>(K); k_len = ar.get<uint32_t>(K); nxt = ar.get<uint32_t>(K);
        k_start = ar.get<unsigned long long>(K); k_end = ar.get<unsigned long long>(K); k_hash = ar.get<unsigned long long>(K); k_hash32 = ar.get<uint32_t>(K); k_bits = ar.get<uint8_t>(K);
        SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, compact_kernel, grid_for(N, 256), 256, 0, s, recs, N, keep, kord, pkey, pmax, k_rec, k_start, k_end, k_len,
                     k_hash, k_hash32, k_bits, d_fold, d_maxspan, d_err);
        if ((rc = dev_error(ctx, s, d_err, "sorted"))) return rc;
    }
    stats->n_kept = (int64_t)K;
    stats->alignmentCount = (int64_t)K;                                   // every kept read is written exactly once (:1275,:1365)
    SSB_CUDA(ctx, cudaEventRecord(ev[2], s));

    // ---------------------------------------------------------------- output order + emit
    size_t R = 0; int64_t n_cov = 0;
    CovRun *runs = NULL; uint8_t *cls = NULL;
    uint32_t *perm = NULL, *cum_s = NULL, *cum_e = NULL; unsigned long long *s_end = NULL, *out_off = NULL, *ord_off = NULL;
    unsigned long long total_out = 0;
    if (K) {
        int rc;
        uint32_t *iota = ar.get<uint32_t>(K); perm = ar.get<uint32_t>(K); s_end = ar.get<unsigned long long>(K);
        unsigned long long *olen = ar.get<unsigned long long>(K); out_off = ar.get<unsigned long long>(K); ord_off = ar.get<unsigned long long>(K);
        SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, iota_kernel, grid_for(K, 256), 256, 0, s, iota, K);
        int tid_bits = 1; while ((1 << tid_bits) < sp->n_contigs + 1 && tid_bits < 31) tid_bits++;
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(NULL, bytes, k_end, s_end, iota, perm, K, 0, 32 + tid_bits, s);
        void *tmp = ar.get<uint8_t>(bytes); SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(tmp, bytes, k_end, s_end, iota, perm, K, 0, 32 + tid_bits, s));   // stable: ties keep input order
        ctx->launches += 8;
        EmitDesc *edesc = ar.get<EmitDesc>(K); SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, outlen_kernel, grid_for(K, 256), 256, 0, s, perm, k_len, k_rec, recs, K, olen, edesc);
        if ((rc = scan_sum(ar, ctx, olen, out_off, K))) return rc;
        unsigned long long last_off = 0, last_len = 0;
        SSB_CUDA(ctx, cudaMemcpyAsync(&last_off, out_off + K - 1, 8, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(&last_len, olen + K - 1, 8, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        total_out = last_off + last_len;
        if (total_out > out_cap) { snprintf(ctx->err, sizeof ctx->err, "spike: output needs %llu bytes, capacity %zu", total_out, out_cap); return SSB_E_ARG; }
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, ordoff_kernel, grid_for(K, 256), 256, 0, s, perm, out_off, K, ord_off);
        SSB_CUDA(ctx, cudaEventRecord(ev[3], s));
        {
            const char *ev_ = getenv("SSB_EMIT_VARIANT"); const int v = ev_ ? atoi(ev_) : 0;
            if (v == 1) SSB_LAUNCH_P(ctx, SSB_K_SPIKE_EMIT, (emit_kernel<2, 4>), ctx->sm_count * 4, 256, 0, s, d_sam, n, edesc, out_off, K, d_out);
            else SSB_LAUNCH_P(ctx, SSB_K_SPIKE_EMIT, (emit_kernel<1, 8>), ctx->sm_count * 8, 256, 0, s, d_sam, n, edesc, out_off, K, d_out);
        }
    } else SSB_CUDA(ctx, cudaEventRecord(ev[3], s));
    *out_bytes = (size_t)total_out;
    stats->out_bytes = (int64_t)total_out;
    SSB_CUDA(ctx, cudaEventRecord(ev[4], s));

    // ---------------------------------------------------------------- mates, coverage runs, classes, depth
    if (K) {
        int rc;
        prv = ar.get<uint32_t>(K); cplx = ar.get<uint8_t>(K); SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemsetAsync(prv, 0xff, K * sizeof(uint32_t), s));
        SSB_CUDA(ctx, cudaMemsetAsync(cplx, 0, K, s));
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, mates_kernel, grid_for(K, 128), 128, 0, s, d_sam, recs, k_rec, k_start, k_end, k_hash32, K, nxt, prv, cplx);
        unsigned long long *pm = ar.get<unsigned long long>(K); uint32_t *rflag = ar.get<uint32_t>(K), *rid = ar.get<uint32_t>(K);
        SPK_CHECK_ARENA(ar);
        if ((rc = scan_max_excl(ar, ctx, k_end, pm, K, 0ull))) return rc;
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, runflag_kernel, grid_for(K, 256), 256, 0, s, k_start, pm, K, rflag);
        if ((rc = scan_sum_incl(ar, ctx, rflag, rid, K))) return rc;
        uint32_t nruns = 0;
        SSB_CUDA(ctx, cudaMemcpyAsync(&nruns, rid + K - 1, 4, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        R = nruns;
        runs = ar.get<CovRun>(R);
        unsigned long long *rlen = ar.get<unsigned long long>(R), *rbase = ar.get<unsigned long long>(R);
        SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, runs_kernel, grid_for(K, 256), 256, 0, s, k_start, k_end, pm, rflag, rid, K, runs, rlen);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, runlen_kernel, grid_for(R, 256), 256, 0, s, runs, R, rlen);
        if ((rc = scan_sum(ar, ctx, rlen, rbase, R))) return rc;
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, runbase_kernel, grid_for(R, 256), 256, 0, s, runs, R, rbase);
        unsigned long long lb = 0, ll = 0;
        SSB_CUDA(ctx, cudaMemcpyAsync(&lb, rbase + R - 1, 8, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(&ll, rlen + R - 1, 8, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        n_cov = (int64_t)(lb + ll);
        cum_s = ar.get<uint32_t>((size_t)n_cov + 1); cum_e = ar.get<uint32_t>((size_t)n_cov + 1); SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemsetAsync(cum_e, 0, ((size_t)n_cov + 1) * sizeof(uint32_t), s));       // loci before the first read end
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, cum_fill_kernel, grid_for(K, 256), 256, 0, s, k_start, K, runs, R, n_cov, cum_s);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, cum_fill_kernel, grid_for(K, 256), 256, 0, s, s_end, K, runs, R, n_cov, cum_e);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, maxdepth_kernel, ctx->sm_count * 8, 256, 0, s, cum_s, cum_e, n_cov, d_maxdepth);
    }
    stats->n_runs = (int64_t)R;
    stats->numberOfLociCovered = n_cov;
    SSB_CUDA(ctx, cudaEventRecord(ev[5], s));

    // ---------------------------------------------------------------- targets
    std::vector<DevTarget> ht(T);
    for (size_t t = 0; t < T; t++) {
        ht[t].c_tid = targets[t].c_tid; ht[t].locus = targets[t].locus; ht[t].base = targets[t].base;
        memset(ht[t].pad, 0, sizeof ht[t].pad);
        const double thr = (double)targets[t].af * 2147483648.0;          // coinToss: rand() < p * (RAND_MAX + 1.0), p a float promoted to double
        ht[t].thresh = thr > 0 ? (thr >= 2147483648.0 ? 2147483648u : (uint32_t)ceil(thr)) : 0u;
    }
    size_t H = 0;
    HitTarget *hits = NULL; ssb_target_result *d_res = NULL;
    if (T) {
        DevTarget *d_tg = ar.get<DevTarget>(T); d_res = ar.get<ssb_target_result>(T);
        SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemcpyAsync(d_tg, ht.data(), T * sizeof(DevTarget), cudaMemcpyHostToDevice, s));
        if (n_cov == 0) {
            for (size_t t = 0; t < T; t++) { memset(&results[t], 0, sizeof results[t]); results[t].status = SSB_T_TAIL; results[t].at_tid = -1; results[t].at_pos = -1; results[t].rng_offset = -1; }
            SSB_CUDA(ctx, cudaStreamSynchronize(s));
        } else {
            int rc;
            long long *v = ar.get<long long>(T), *vmax = ar.get<long long>(T);
            uint32_t *hitflag = ar.get<uint32_t>(T), *hidx = ar.get<uint32_t>(T);
            SPK_CHECK_ARENA(ar);
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, target_lb_kernel, grid_for(T, 256), 256, 0, s, d_tg, T, runs, R, n_cov, v);
            if ((rc = scan_max_incl(ar, ctx, v, vmax, T))) return rc;
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, target_status_kernel, grid_for(T, 256), 256, 0, s, d_tg, T, vmax, runs, R, n_cov, d_res, hitflag);
            if ((rc = scan_sum(ar, ctx, hitflag, hidx, T))) return rc;
            uint32_t lh = 0, lf = 0;
            SSB_CUDA(ctx, cudaMemcpyAsync(&lh, hidx + T - 1, 4, cudaMemcpyDeviceToHost, s));
            SSB_CUDA(ctx, cudaMemcpyAsync(&lf, hitflag + T - 1, 4, cudaMemcpyDeviceToHost, s));
            SSB_CUDA(ctx, cudaStreamSynchronize(s));
            H = (size_t)lh + lf;
            hits = ar.get<HitTarget>(H); SPK_CHECK_ARENA(ar);
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, hits_kernel, grid_for(T, 256), 256, 0, s, d_tg, T, hitflag, hidx, d_res, hits);
        }
    }
    stats->n_hits = (int64_t)H;
    SSB_CUDA(ctx, cudaEventRecord(ev[6], s));

    dbg_mark("targets");
    // ---------------------------------------------------------------- gather + rng + chain + patch
    if (H) {
        int rc;
        unsigned long long *cnt = ar.get<unsigned long long>(H + 1), *eoff = ar.get<unsigned long long>(H + 1);
        SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemsetAsync(cnt, 0, (H + 1) * sizeof(unsigned long long), s));
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, gather_count_kernel, grid_for(H * 32, 128), 128, 0, s, hits, H, k_start, k_end, K, d_maxspan, cnt, d_err);
        if ((rc = scan_sum(ar, ctx, cnt, eoff, H + 1))) return rc;
        unsigned long long E = 0; HitTarget last_hit;
        SSB_CUDA(ctx, cudaMemcpyAsync(&E, eoff + H, 8, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(&last_hit, hits + H - 1, sizeof last_hit, cudaMemcpyDeviceToHost, s));
        if ((rc = dev_error(ctx, s, d_err, "gather"))) return rc;
        PlpEntry *ent = ar.get<PlpEntry>(E); uint8_t *hflag = ar.get<uint8_t>(E);
        Patch *patches = ar.get<Patch>(2 * E + 16); unsigned int *n_patches = ar.get<unsigned int>(1);
        unsigned long long *d_draws = ar.get<unsigned long long>(1);
        SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, gather_fill_kernel, grid_for(H * 32, 128), 128, 0, s, d_sam, recs, k_rec, hits, H, k_start, k_end, K, d_maxspan, eoff, ent);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, gather_mate_kernel, grid_for(H * 32, 128), 128, 0, s, d_sam, recs, k_rec, k_start, k_end, k_hash, K, nxt, hits, H, eoff, ent);
        // reference classes of the covered loci the chain walks over (up to the last hit)
        const int64_t n_walk = last_hit.locus_index + 1;
        cls = ar.get<uint8_t>((size_t)n_walk + 64); SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, cls_kernel, grid_for((size_t)n_walk, 256), 256, 0, s, runs, R, n_walk, (const uint8_t *const *)sp->d_seq_ptrs, sp->d_lens, cls, d_err);
        if ((rc = dev_error(ctx, s, d_err, "reference"))) return rc;
        SSB_CUDA(ctx, cudaEventRecord(ev[7], s));

        dbg_mark("gathered");
        uint32_t seedw[61];
        glibc_seed_window(seed, seedw);
        uint32_t *d_seedw = ar.get<uint32_t>(61); SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemcpyAsync(d_seedw, seedw, sizeof seedw, cudaMemcpyHostToDevice, s));
        // reference class bit planes
        const size_t cwords = (size_t)((n_walk + 31) >> 5) + 160;
        uint32_t *pc0 = ar.get<uint32_t>(cwords), *pc1 = ar.get<uint32_t>(cwords), *pcx = ar.get<uint32_t>(cwords);
        SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemsetAsync(pc0, 0, cwords * 4, s)); SSB_CUDA(ctx, cudaMemsetAsync(pc1, 0, cwords * 4, s)); SSB_CUDA(ctx, cudaMemsetAsync(pcx, 0xff, cwords * 4, s));
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, cls_pack_kernel, grid_for(((size_t)n_walk + 31) & ~(size_t)31, 256), 256, 0, s, cls, n_walk, pc0, pc1, pcx);

        // ---- expected draws per chunk: window centres / widths, and the stream length
        const char *env_serial = getenv("SSB_CHAIN_SERIAL");
        int P = 1; int64_t Lc = n_walk;
        const char *env_chunk = getenv("SSB_CHAIN_CHUNK");                 // loci per chunk (testing / tuning)
        // the chunked formulation pays off when the walk between targets dominates; every phase-1 walker dry-runs the pileups it
        // passes, so dense deep panels (pileup entries comparable to walked loci) stay on the one-warp serial chain
        const bool sparse_targets = (unsigned long long)E * 16ull < (unsigned long long)n_walk;
        if (!(env_serial && env_serial[0] == '1') && ((n_walk >= (1 << 20) && sparse_targets) || env_chunk)) {
            Lc = n_walk / 4096; if (Lc < 8192) Lc = 8192;              // phase 3 walks one chunk per warp: short chunks keep its chain short
            if (env_chunk && atoll(env_chunk) >= 64) Lc = atoll(env_chunk);
            Lc = (Lc + 31) & ~(int64_t)31;
            P = (int)((n_walk + Lc - 1) / Lc);
            if (P < 2) { P = 1; Lc = n_walk; }
        }
        double *d_mean = ar.get<double>((size_t)P), *d_var = ar.get<double>((size_t)P);
        SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemsetAsync(d_mean, 0, P * sizeof(double), s)); SSB_CUDA(ctx, cudaMemsetAsync(d_var, 0, P * sizeof(double), s));
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, chunk_stats_kernel, grid_for((size_t)P * 32, 128), 128, 0, s, pcx, n_walk, Lc, P, d_mean, d_var);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, expect_kernel, grid_for(H * 32, 128), 128, 0, s, hits, H, eoff, ent, (const uint8_t *const *)sp->d_seq_ptrs, Lc, d_mean, d_var);
        std::vector<double> h_mean(P), h_var(P);
        SSB_CUDA(ctx, cudaMemcpyAsync(h_mean.data(), d_mean, P * sizeof(double), cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(h_var.data(), d_var, P * sizeof(double), cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        // groups of Rg consecutive chunks share one start-offset window; the window is cut into slices of ~wt offsets
        int Rg = (int)(114688 / Lc); if (Rg < 1) Rg = 1;                    // phase 1 works on groups of ~112 k loci (see spike_chain.cuh)
        uint32_t wt = 16384;
        if (const char *e = getenv("SSB_CHAIN_GROUP")) { if (atoi(e) >= 1) Rg = atoi(e); }
        if (const char *e = getenv("SSB_CHAIN_SLICE")) { if (atoi(e) >= 1) wt = (uint32_t)atoi(e); }
        const int G = (P + Rg - 1) / Rg;
        std::vector<GroupDesc> h_groups(G);
        std::vector<SliceDesc> h_slices;
        std::vector<ChunkDesc> h_chunks(P);
        double cm = 0, cv = 0; unsigned long long woff = 0;
        for (int q = 0; q < G; q++) {
            GroupDesc &gd = h_groups[q];
            gd.f0 = q * Rg; gd.nf = (gd.f0 + Rg <= P) ? Rg : P - gd.f0;
            const double half = q == 0 ? 0.0 : 5.0 * sqrt(cv) + 48.0;        // +-5 sigma: a miss (3e-7 per group) falls back to the serial chain
            double lo = cm - half; if (lo < 0) lo = 0;
            gd.klo = (unsigned long long)lo; gd.W = q == 0 ? 1u : (uint32_t)(cm + half - (double)gd.klo) + 2u;
            uint32_t S = (gd.W + wt - 1) / wt; gd.w = (gd.W + S - 1) / S; S = (gd.W + gd.w - 1) / gd.w;
            gd.S = S; gd.b0 = (uint32_t)h_slices.size();
            for (uint32_t sl = 0; sl < S; sl++) {
                SliceDesc sd; sd.q = q; sd.i0 = sl * gd.w; sd.n = (sd.i0 + gd.w <= gd.W) ? gd.w : gd.W - sd.i0; sd.off = woff; woff += sd.n;
                h_slices.push_back(sd);
            }
            for (int f = gd.f0; f < gd.f0 + gd.nf; f++) {
                h_chunks[f].g0 = (int64_t)f * Lc; h_chunks[f].g1 = (int64_t)(f + 1) * Lc < n_walk ? (int64_t)(f + 1) * Lc : n_walk;
                h_chunks[f].k_in = 0; h_chunks[f].k_out = ~0ull;
                cm += h_mean[f]; cv += h_var[f];
            }
        }
        const size_t n_slices = h_slices.size();
        const unsigned long long pool_cap = woff + (1ull << 20);
        // the stream must cover the top of the last window (and everything a lone walker can reach)
        unsigned long long M = (unsigned long long)(cm + 8.0 * sqrt(cv)) + 3 * E + (1u << 17);
        float ms_rng = 0, ms_chain = 0;
        unsigned int *d_flags = ar.get<unsigned int>(1);
        ChunkDesc *d_chunks = ar.get<ChunkDesc>((size_t)P), *d_serial = ar.get<ChunkDesc>(1);
        GroupDesc *d_groups = ar.get<GroupDesc>((size_t)G); SliceDesc *d_slices = ar.get<SliceDesc>(n_slices);
        BoundaryList *d_lists = ar.get<BoundaryList>(n_slices * (size_t)Rg);
        unsigned long long *d_gk = ar.get<unsigned long long>((size_t)G), *d_pool_used = ar.get<unsigned long long>(1);
        unsigned long long *kbuf = NULL, *pool_k = NULL; uint32_t *lobuf = NULL, *pool_lo = NULL;
        if (P > 1) { kbuf = ar.get<unsigned long long>(2 * woff); lobuf = ar.get<uint32_t>(2 * woff); pool_k = ar.get<unsigned long long>(pool_cap); pool_lo = ar.get<uint32_t>(pool_cap); }
        SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemcpyAsync(d_groups, h_groups.data(), G * sizeof(GroupDesc), cudaMemcpyHostToDevice, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(d_slices, h_slices.data(), n_slices * sizeof(SliceDesc), cudaMemcpyHostToDevice, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(d_chunks, h_chunks.data(), P * sizeof(ChunkDesc), cudaMemcpyHostToDevice, s));
        ChunkDesc serial_cd; serial_cd.g0 = 0; serial_cd.g1 = n_walk; serial_cd.k_in = 0; serial_cd.k_out = ~0ull;
        SSB_CUDA(ctx, cudaMemcpyAsync(d_serial, &serial_cd, sizeof serial_cd, cudaMemcpyHostToDevice, s));
        bool parallel = P > 1;
        stats->n_runs = (int64_t)R;
        bool chain_done = false;
        for (int attempt = 0; attempt < 8; attempt++) {
            M = (M + RNG_BLOCK - 1) / RNG_BLOCK * RNG_BLOCK;
            const size_t nblocks = (size_t)(M / RNG_BLOCK);
            // per-block polynomials x^(310 + b*RNG_BLOCK) mod P: seed independent, built on the host (31x31 products)
            std::vector<uint32_t> bp(nblocks * GLIBC_DEG);
            uint32_t stepb[GLIBC_DEG], cur[GLIBC_DEG], tmpb[GLIBC_DEG];
            glibc_poly_xpow(RNG_BLOCK, stepb);
            glibc_poly_xpow(310, cur);
            for (size_t b = 0; b < nblocks; b++) { memcpy(&bp[b * GLIBC_DEG], cur, sizeof cur); glibc_poly_mulmod(cur, stepb, tmpb); memcpy(cur, tmpb, sizeof cur); }
            const size_t ewords = (size_t)(M >> 5) + 160;
            uint32_t *d_bp = ar.get<uint32_t>(bp.size()); int32_t *Rs = ar.get<int32_t>(M + 64);
            uint32_t *pe0 = ar.get<uint32_t>(ewords), *pe1 = ar.get<uint32_t>(ewords), *pej = ar.get<uint32_t>(ewords);
            SPK_CHECK_ARENA(ar);
            SSB_CUDA(ctx, cudaMemcpyAsync(d_bp, bp.data(), bp.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            SSB_CUDA(ctx, cudaMemsetAsync(pe0 + (M >> 5), 0xAA, 160 * 4, s)); SSB_CUDA(ctx, cudaMemsetAsync(pe1 + (M >> 5), 0xCC, 160 * 4, s)); SSB_CUDA(ctx, cudaMemsetAsync(pej + (M >> 5), 0, 160 * 4, s));   // all four classes in every nibble: see walk_loci
            SSB_CUDA(ctx, cudaEventRecord(ev[8], s));
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, rng_fill_kernel, (int)nblocks, RNG_TPB, 0, s, d_bp, sp->d_rng_tab, d_seedw, Rs, M, pe0, pe1, pej);
            SSB_CUDA(ctx, cudaEventRecord(ev[9], s));
            ChainArgs A;
            A.e0 = pe0; A.e1 = pe1; A.ej = pej; A.R = Rs; A.M = M; A.c0 = pc0; A.c1 = pc1; A.cx = pcx; A.n_walk = n_walk;
            A.hits = hits; A.H = H; A.eoff = eoff; A.ent = ent; A.hflag = hflag; A.res = d_res;
            A.patches = patches; A.n_patches = n_patches; A.patch_cap = (unsigned int)(2 * E + 16);
            A.odd = d_odd; A.n_odd = d_nodd; A.odd_cap = odd_cap; A.odd_bloom = d_bloom;
            A.contig_seq = (const uint8_t *const *)sp->d_seq_ptrs; A.err = d_err;
            unsigned int flags = 0, n_odd_h = 0;
            auto reset_apply = [&]() -> int {
                SSB_CUDA(ctx, cudaMemsetAsync(hflag, 0, E, s));
                SSB_CUDA(ctx, cudaMemsetAsync(n_patches, 0, sizeof(unsigned int), s));
                SSB_CUDA(ctx, cudaMemsetAsync(d_nodd, 0, sizeof(unsigned int), s));
                SSB_CUDA(ctx, cudaMemsetAsync(d_bloom, 0, sizeof(unsigned long long), s));
                SSB_CUDA(ctx, cudaMemsetAsync(d_flags, 0, sizeof(unsigned int), s));
                return SSB_OK;
            };
            if ((rc = reset_apply())) return rc;
            if (parallel) {
                unsigned long long *d_dbg = NULL;
                if (getenv("SSB_CHAIN_DEBUG")) { d_dbg = ar.get<unsigned long long>(8 + 2 * n_slices); SPK_CHECK_ARENA(ar); SSB_CUDA(ctx, cudaMemsetAsync(d_dbg, 0, (8 + 2 * n_slices) * 8, s)); }
                SSB_CUDA(ctx, cudaMemsetAsync(d_pool_used, 0, 8, s));
                SSB_CUDA(ctx, cudaMemsetAsync(d_lists, 0, n_slices * (size_t)Rg * sizeof(BoundaryList), s));
                SSB_LAUNCH_P(ctx, SSB_K_SPIKE_CHAIN, phase1_kernel, (int)n_slices, P1_THREADS, P1_SMEM, s, A, Lc, d_groups, d_slices, kbuf, lobuf, woff,
                             d_lists, Rg, pool_k, pool_lo, d_pool_used, pool_cap, d_flags, d_dbg);
                if (d_dbg) {
                    unsigned long long h_dbg[8], used = 0;
                    SSB_CUDA(ctx, cudaMemcpyAsync(h_dbg, d_dbg, 64, cudaMemcpyDeviceToHost, s));
                    SSB_CUDA(ctx, cudaMemcpyAsync(&used, d_pool_used, 8, cudaMemcpyDeviceToHost, s));
                    SSB_CUDA(ctx, cudaStreamSynchronize(s));
                    fprintf(stderr, "[chain] chunks=%d L=%lld groups=%d slices=%zu walkers=%llu walker-loci=%llu rounds=%llu final survivors: sum %llu max %llu, recorded %llu\n",
                            P, (long long)Lc, G, n_slices, woff, h_dbg[0], h_dbg[3], h_dbg[1], h_dbg[2], used);
                    fprintf(stderr, "[chain] phase 1 block cycles: rounds with > %d walkers %.3g (avg per block), later rounds %.3g, slowest block %.3g\n", P1_THREADS,
                            (double)h_dbg[4] / (double)n_slices, (double)h_dbg[5] / (double)n_slices, (double)h_dbg[6]);
                    if (getenv("SSB_CHAIN_DEBUG")[0] == '3') {
                        std::vector<unsigned long long> pb(2 * n_slices);
                        SSB_CUDA(ctx, cudaMemcpy(pb.data(), d_dbg + 8, 2 * n_slices * 8, cudaMemcpyDeviceToHost));
                        for (size_t b = 0; b < n_slices; b += (n_slices / 60 ? n_slices / 60 : 1))
                            fprintf(stderr, "[chain]   block %zu group %d i0 %u n %u: total %.3g early %.3g\n", b, h_slices[b].q, h_slices[b].i0, h_slices[b].n, (double)pb[2 * b], (double)pb[2 * b + 1]);
                    }
                }
                SSB_LAUNCH_P(ctx, SSB_K_SPIKE_CHAIN, compose_kernel, 1, 32, 0, s, G, d_groups, d_lists, Rg, pool_k, pool_lo, d_gk, d_flags);
                SSB_LAUNCH_P(ctx, SSB_K_SPIKE_CHAIN, boundary_kernel, (P + 127) / 128, 128, 0, s, P, Rg, d_groups, d_lists, Rg, pool_k, pool_lo, d_gk, d_chunks, d_flags);
            }
            if (parallel) {
                SSB_CUDA(ctx, cudaMemcpyAsync(&flags, d_flags, 4, cudaMemcpyDeviceToHost, s));
                SSB_CUDA(ctx, cudaStreamSynchronize(s));
                if (getenv("SSB_CHAIN_DEBUG")) fprintf(stderr, "[chain] flags after phase 1/2: %u\n", flags);
                if (!flags) {
                    SSB_LAUNCH_P(ctx, SSB_K_SPIKE_CHAIN, chain_kernel, grid_for((size_t)P * 32, 128), 128, 0, s, A, d_chunks, P, d_draws, d_flags);
                    SSB_CUDA(ctx, cudaMemcpyAsync(&flags, d_flags, 4, cudaMemcpyDeviceToHost, s));
                    SSB_CUDA(ctx, cudaMemcpyAsync(&n_odd_h, d_nodd, 4, cudaMemcpyDeviceToHost, s));
                    SSB_CUDA(ctx, cudaStreamSynchronize(s));
                }
                if (getenv("SSB_CHAIN_DEBUG")) fprintf(stderr, "[chain] flags after phase 3: %u, odd patches %u\n", flags, n_odd_h);
                if (flags & CHAIN_OVERRUN) { M *= 2; continue; }
                if (flags || n_odd_h) {               // window miss / too complex / odd patches: the plain serial chain decides
                    parallel = false;
                    if ((rc = reset_apply())) return rc;
                }
            }
            if (!parallel) {
                SSB_LAUNCH_P(ctx, SSB_K_SPIKE_CHAIN, chain_kernel, 1, 32, 0, s, A, d_serial, 1, d_draws, d_flags);
                SSB_CUDA(ctx, cudaMemcpyAsync(&flags, d_flags, 4, cudaMemcpyDeviceToHost, s));
                SSB_CUDA(ctx, cudaStreamSynchronize(s));
                if (flags & CHAIN_OVERRUN) { M *= 2; continue; }
            }
            SSB_CUDA(ctx, cudaEventRecord(ev[10], s));
            DevErr e;
            SSB_CUDA(ctx, cudaMemcpyAsync(&e, d_err, sizeof e, cudaMemcpyDeviceToHost, s));
            SSB_CUDA(ctx, cudaStreamSynchronize(s));
            ms_rng += ev_ms(ev[8], ev[9]); ms_chain += ev_ms(ev[9], ev[10]);
            if (e.code) { snprintf(ctx->err, sizeof ctx->err, "spike/chain: %s", ssb_strerror(e.code)); return e.code; }
            stats->chain_mode = parallel ? P : 1;
            chain_done = true;
            break;
        }
        if (!chain_done) { snprintf(ctx->err, sizeof ctx->err, "spike/chain: rand() stream exhausted"); return SSB_E_STATE; }
        dbg_mark("chain-end");
        stats->ms_rng = ms_rng; stats->ms_chain = ms_chain;
        unsigned long long draws = 0;
        SSB_CUDA(ctx, cudaMemcpyAsync(&draws, d_draws, 8, cudaMemcpyDeviceToHost, s));
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, patch_kernel, 64, 256, 0, s, patches, n_patches, recs, k_rec, ord_off, d_out);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, odd_fix_kernel, 4, 128, 0, s, d_odd, d_nodd, odd_cap, patches, n_patches, recs, k_rec, ord_off, d_out);
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        stats->rng_draws = (int64_t)draws;
        dbg_mark("patched");
    } else {
        SSB_CUDA(ctx, cudaEventRecord(ev[7], s));
        SSB_CUDA(ctx, cudaEventRecord(ev[10], s));
    }
    dbg_mark("begin");
    if (n_cov) {
        int rc;
        // per-locus tallies of the non-target loci (SEQ_ERROR lines)
        err64 = ar.get<unsigned long long>((size_t)n_cov + 1); minus = ar.get<unsigned int>((size_t)n_cov + 1);
        SPK_CHECK_ARENA(ar);
        SSB_CUDA(ctx, cudaMemsetAsync(err64, 0, ((size_t)n_cov + 1) * sizeof(unsigned long long), s));
        SSB_CUDA(ctx, cudaMemsetAsync(minus, 0, ((size_t)n_cov + 1) * sizeof(unsigned int), s));
        dbg_mark("memset");
        unsigned int h_maxspan = 0;
        SSB_CUDA(ctx, cudaMemcpyAsync(&h_maxspan, d_maxspan, 4, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        TallyArgs TA;
        TA.sam = d_sam; TA.recs = recs; TA.k_rec = k_rec; TA.k_start = k_start; TA.k_end = k_end; TA.k_hash = k_hash; TA.k_bits = k_bits; TA.K = K;
        TA.nxt = nxt; TA.prv = prv; TA.cplx = cplx; TA.maxspan = h_maxspan; TA.runs = runs; TA.R = R;
        TA.contig_seq = (const uint8_t *const *)sp->d_seq_ptrs; TA.contig_len = sp->d_lens; TA.err64 = err64; TA.minus = minus; TA.err = d_err;
        TA.odd = d_odd; TA.n_odd = d_nodd; TA.odd_bloom = d_bloom;
        // exceptional bases of the simple reads: listed by the tokeniser (unless the list overflowed or is switched off)
        unsigned long long n_list = 0;
        if (d_exc_count) { SSB_CUDA(ctx, cudaMemcpyAsync(&n_list, d_exc_count, 8, cudaMemcpyDeviceToHost, s)); SSB_CUDA(ctx, cudaStreamSynchronize(s)); }
        const bool use_list = d_exc_count && !getenv("SSB_NO_EXC_LIST") && n_list <= exc_cap;
        TA.listed_only = use_list ? 1 : 0;
        if (use_list) {
            if (n_list) SSB_LAUNCH_P(ctx, SSB_K_SPIKE_TALLY, tally_resolve_kernel, grid_for((size_t)n_list, 128), 128, 0, s, TA, exc_list, n_list, keep, kord);
        } else {
            KMeta *kmeta = ar.get<KMeta>(K); SPK_CHECK_ARENA(ar);
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_TALLY, kmeta_kernel, grid_for(K, 256), 256, 0, s, recs, k_rec, K, kmeta);
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_TALLY, tally_fast_kernel, grid_for(K * 32, 128), 128, 0, s, TA, kmeta);
        }
        dbg_mark("resolve");
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_TALLY, tally_kernel, grid_for(K, 128), 128, 0, s, TA);
        dbg_mark("generic");
        if ((rc = dev_error(ctx, s, d_err, "reference"))) return rc;
        if (H) SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, tally_clear_hits_kernel, grid_for(H, 256), 256, 0, s, hits, H, err64);
        uint32_t *sflag = ar.get<uint32_t>((size_t)n_cov), *sidx = ar.get<uint32_t>((size_t)n_cov);
        SPK_CHECK_ARENA(ar);
        SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, tally_flag_kernel, grid_for((size_t)n_cov, 256), 256, 0, s, err64, n_cov, sflag);
        if ((rc = scan_sum(ar, ctx, sflag, sidx, (size_t)n_cov))) return rc;
        uint32_t li = 0, lf = 0;
        SSB_CUDA(ctx, cudaMemcpyAsync(&li, sidx + n_cov - 1, 4, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaMemcpyAsync(&lf, sflag + n_cov - 1, 4, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        dbg_mark("flag+scan");
        sp->n_se = (size_t)li + lf;
        if (sp->n_se) {
            SSB_CUDA(ctx, cudaMallocAsync((void **)&sp->d_se, sp->n_se * sizeof(ssb_seq_error), s));
            SSB_LAUNCH_P(ctx, SSB_K_SPIKE_OTHER, tally_emit_kernel, grid_for((size_t)n_cov, 256), 256, 0, s, err64, minus, sflag, sidx, n_cov, runs, R,
                         cum_s, cum_e, (const uint8_t *const *)sp->d_seq_ptrs, sp->d_se);
        }
    }
    dbg_mark("emit");
    SSB_CUDA(ctx, cudaEventRecord(ev[11], s));
    if (T && n_cov) SSB_CUDA(ctx, cudaMemcpyAsync(results, d_res, T * sizeof(ssb_target_result), cudaMemcpyDeviceToHost, s));
    unsigned long long fold = 0; unsigned int md = 0;
    SSB_CUDA(ctx, cudaMemcpyAsync(&fold, d_fold, 8, cudaMemcpyDeviceToHost, s));
    SSB_CUDA(ctx, cudaMemcpyAsync(&md, d_maxdepth, 4, cudaMemcpyDeviceToHost, s));
    SSB_CUDA(ctx, cudaStreamSynchronize(s));
    stats->totalFoldCoverage = (int64_t)fold;
    stats->maxDepth = (int64_t)md;
    if (md > MAX_PILEUP) { snprintf(ctx->err, sizeof ctx->err, "spike: pileup depth %u exceeds MAX_PILEUP_SIZE", md); return SSB_E_DEPTH; }
    stats->ms_parse = ev_ms(ev[0], ev[1]);
    stats->ms_sort = ev_ms(ev[1], ev[3]);
    stats->ms_emit = ev_ms(ev[3], ev[4]);
    stats->ms_cover = ev_ms(ev[4], ev[5]);
    stats->ms_gather = ev_ms(ev[5], ev[7]);
    stats->ms_patch = H ? ev_ms(ev[10], ev[11]) : 0;
    stats->ms_total = ev_ms(ev[0], ev[11]);
    return SSB_OK;
}

extern "C" int ssb_spike_run_host(ssb_spike *sp, const uint8_t *sam, size_t n, uint8_t *out, size_t out_cap,
                                  const ssb_target *targets, size_t n_targets, unsigned seed,
                                  ssb_target_result *results, ssb_spike_stats *stats, size_t *out_bytes)
{
    if (!sp || (!sam && n) || (!out && n) || !stats || !out_bytes) return SSB_E_ARG;
    ssb_ctx *ctx = sp->ctx;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t *d_in = NULL, *d_out = NULL;
    SSB_CUDA(ctx, cudaMallocAsync((void **)&d_in, n + 64, ctx->stream));
    SSB_CUDA(ctx, cudaMallocAsync((void **)&d_out, n + 64, ctx->stream));
    if (n) SSB_CUDA(ctx, cudaMemcpyAsync(d_in, sam, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = ssb_spike_run_device(sp, d_in, n, d_out, n + 1, targets, n_targets, seed, results, stats, out_bytes);
    if (rc == SSB_OK) {
        if (*out_bytes > out_cap) { snprintf(ctx->err, sizeof ctx->err, "spike: output needs %zu bytes, capacity %zu", *out_bytes, out_cap); rc = SSB_E_ARG; }
        else if (*out_bytes) {
            cudaError_t e = cudaMemcpyAsync(out, d_out, *out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { snprintf(ctx->err, sizeof ctx->err, "spike: copy back: %s", cudaGetErrorString(e)); rc = SSB_E_CUDA; }
        }
    }
    cudaFreeAsync(d_in, ctx->stream); cudaFreeAsync(d_out, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

// rand() #k0 .. #k0+n-1 of srand(seed) (glibc TYPE_3), generated by the same kernel the spike path uses.
extern "C" int ssb_spike_rand(ssb_spike *sp, unsigned seed, uint64_t k0, size_t n, int32_t *out_host)
{
    if (!sp || (!out_host && n)) return SSB_E_ARG;
    if (!n) return SSB_OK;
    ssb_ctx *ctx = sp->ctx;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    Arena ar; ar.ctx = ctx; ar.s = s;
    const unsigned long long M = (n + RNG_BLOCK - 1) / RNG_BLOCK * RNG_BLOCK;
    const size_t nblocks = (size_t)(M / RNG_BLOCK);
    std::vector<uint32_t> bp(nblocks * GLIBC_DEG);
    uint32_t stepb[GLIBC_DEG], cur[GLIBC_DEG], tmpb[GLIBC_DEG];
    glibc_poly_xpow(RNG_BLOCK, stepb);
    glibc_poly_xpow(310 + k0, cur);
    for (size_t b = 0; b < nblocks; b++) { memcpy(&bp[b * GLIBC_DEG], cur, sizeof cur); glibc_poly_mulmod(cur, stepb, tmpb); memcpy(cur, tmpb, sizeof cur); }
    uint32_t seedw[61];
    glibc_seed_window(seed, seedw);
    uint32_t *d_bp = ar.get<uint32_t>(bp.size()), *d_seedw = ar.get<uint32_t>(61); int32_t *R = ar.get<int32_t>(M);
    SPK_CHECK_ARENA(ar);
    SSB_CUDA(ctx, cudaMemcpyAsync(d_bp, bp.data(), bp.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    SSB_CUDA(ctx, cudaMemcpyAsync(d_seedw, seedw, sizeof seedw, cudaMemcpyHostToDevice, s));
    SSB_LAUNCH(ctx, rng_fill_kernel, (int)nblocks, RNG_TPB, 0, s, d_bp, sp->d_rng_tab, d_seedw, R, M, (uint32_t *)NULL, (uint32_t *)NULL, (uint32_t *)NULL);
    SSB_CUDA(ctx, cudaMemcpyAsync(out_host, R, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    SSB_CUDA(ctx, cudaStreamSynchronize(s));
    return SSB_OK;
}

extern "C" int ssb_spike_seq_error_count(ssb_spike *sp, size_t *count)
{
    if (!sp || !count) return SSB_E_ARG;
    *count = sp->n_se;
    return SSB_OK;
}

extern "C" int ssb_spike_seq_errors(ssb_spike *sp, ssb_seq_error *dst, size_t cap)
{
    if (!sp || (!dst && cap)) return SSB_E_ARG;
    ssb_ctx *ctx = sp->ctx;
    size_t n = sp->n_se < cap ? sp->n_se : cap;
    if (n) {
        SSB_CUDA(ctx, cudaSetDevice(ctx->device));
        SSB_CUDA(ctx, cudaMemcpyAsync(dst, sp->d_se, n * sizeof(ssb_seq_error), cudaMemcpyDeviceToHost, ctx->stream));
        SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return SSB_OK;
}
