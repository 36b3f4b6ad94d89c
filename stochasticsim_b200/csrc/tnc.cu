// tnc.cu -- trinucleotide-context scan on B200 (hot path 2).
//
// Replaces the getline / 3-byte-window loop of tncCountsProfile.c:391-447 and the 64-way
// strncmp chain incCtx (tncCountsProfile.c:105-363).  Results are bit-exact with the
// reference, including its three quirks:
//   * a record is KEPT unless empty, starting with '>' or free of upper-case G/C/A/T
//     (tncCountsProfile.c:398-407); non-kept records are invisible, contigs are joined;
//   * every window of three upper-case bases inside a kept record counts (:430-438);
//   * between two kept records exactly ONE straddling window (last(l), l'[0], l'[1]) counts,
//     the other one is lost to the newline slot (:409, :441-443).
//
// Decomposition (one streaming pass over the bytes, no inter-block dependency):
//   total = sum over byte positions p of
//             [b[p-2],b[p-1],b[p] upper-case bases]                         (in-line window)
//           + [b[p-2]=='\n', b[p-3],b[p-1],b[p] upper-case bases]           (straddle, optimistic)
//           - corrections at the (rare) places where the optimistic rule is wrong.
// The optimistic rule is wrong only around header records and where the record before a line
// start does not end in an upper-case base.  tnc_scan_kernel appends those positions to an
// exception list; tnc_fixup_kernel resolves each one exactly by walking back to the nearest kept
// record (one warp per exception).  In a 60-column genome FASTA that is a few thousand
// exceptions per 3 GB.
//
// tnc_scan_kernel is bit-sliced: each thread classifies 32 bytes with SWAR byte tricks (PRMT
// table lookup + zero-byte detection), transposes the per-byte flags into 32-position bit
// planes, and counts all 64 contexts with AND + POPC into 64 register counters.  No shared
// memory atomics in the hot loop.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int      TNC_BLOCK = 256;
constexpr int      TNC_BPT   = 32;            // bytes per thread per iteration
constexpr uint32_t EXC_HEADER = 0x80000000u;  // exception kind flag (positions stay < 2^31)
constexpr size_t   TNC_MAX_PIECE = (size_t)1 << 30;
// A piece this small can never overflow the exception list: at most one record per 2 bytes
// ("\n>" headers) = 65536 < tnc_exc_cap(131072) = 8192 + 65536.
constexpr size_t   TNC_SAFE_PIECE = (size_t)1 << 17;
constexpr int      TNC_SEG_SHIFT = 12;            // 4 KiB segments = 128 chunks of 32 bytes

struct TncDevState {                          // mirrors ssb_tnc_carry
    uint8_t started, prev[3], carry, frag_nonempty, frag_first, frag_has_base;
};
static_assert(sizeof(TncDevState) == sizeof(ssb_tnc_carry), "state layout");

__host__ __device__ inline bool is_base(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
// reference index order A<C<G<T (tncCountsProfile.c:14-77)
__host__ __device__ inline int ref_code(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

// ---- per-word classification -------------------------------------------------------------
// Flags come back in bit 7 of every byte of the word.
struct WordFlags { uint32_t a, c, g, t, nl, gt; };

__device__ __forceinline__ WordFlags classify(uint32_t w)
{
    // 8-entry byte table indexed by the low 3 bits of the byte: the only byte with those low
    // bits that we care about.  A=0x41(1) C=0x43(3) G=0x47(7) T=0x54(4) '\n'=0x0A(2) '>'=0x3E(6).
    // Entries 0 and 5 hold 0x01, which can never equal a byte whose low bits are 0 or 5.
    const uint32_t TLO = 0x430A4101u, THI = 0x473E0154u;
    uint32_t s  = w & 0x07070707u;
    uint32_t e0 = __byte_perm(TLO, THI, s);          // [T[b0], T[0], T[b1], T[0]]
    uint32_t e1 = __byte_perm(TLO, THI, s >> 16);    // [T[b2], T[0], T[b3], T[0]]
    uint32_t e  = __byte_perm(e0, e1, 0x6420);       // [T[b0], T[b1], T[b2], T[b3]]
    uint32_t d  = w ^ e;                             // zero byte <=> byte is one of the six
    uint32_t t2 = (d & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    uint32_t z  = ~(t2 | d | 0x7f7f7f7fu);           // 0x80 where the byte of d is zero (exact)
    uint32_t w1 = w << 1, w5 = w << 5, w6 = w << 6;  // bit6 / bit2 / bit1 of each byte -> bit 7
    uint32_t base = z & w1;
    WordFlags f;
    f.a  = base & ~w5 & ~w6;
    f.c  = base & ~w5 &  w6;
    f.g  = base &  w5 &  w6;
    f.t  = base &  w5 & ~w6;
    f.nl = z & ~w1 & ~w5;
    f.gt = z & ~w1 &  w5;
    return f;
}

// Bit planes over the 32 positions of a thread, in "residue-major" order:
// bit (8k + j) <-> byte k of word j <-> position 4j + k.
struct Planes { uint32_t a, c, g, t, nl, gt; };

// plane value at position p-1 / p-2 / p-3, given the flags of the 3 bytes before the thread's chunk
__device__ __forceinline__ uint32_t back1(uint32_t u, uint32_t h) { return (u << 8)  | ((u >> 23) & 0x000000FEu) | h; }
__device__ __forceinline__ uint32_t back2(uint32_t u, uint32_t h) { return (u << 16) | ((u >> 15) & 0x0000FEFEu) | h; }
__device__ __forceinline__ uint32_t back3(uint32_t u, uint32_t h) { return (u << 24) | ((u >> 7)  & 0x00FEFEFEu) | h; }

// halo flags (bit 7 of bytes 1..3 of the word before the chunk = positions -3,-2,-1)
__device__ __forceinline__ uint32_t halo1(uint32_t f) { return (f >> 31) & 1u; }                                   // pos -1 -> bit 0
__device__ __forceinline__ uint32_t halo2(uint32_t f) { return ((f >> 23) & 1u) | (((f >> 31) & 1u) << 8); }       // -2 -> bit0, -1 -> bit8
__device__ __forceinline__ uint32_t halo3(uint32_t f) { return ((f >> 15) & 1u) | (((f >> 23) & 1u) << 8) | (((f >> 31) & 1u) << 16); }

__device__ __forceinline__ uint32_t ld_word(const uint8_t *b, size_t n, size_t off)
{
    // aligned 4-byte load with zero fill past the end (0x00 is "other": no flag is set for it)
    if (off + 4 <= n) return *reinterpret_cast<const uint32_t *>(b + off);
    uint32_t w = 0;
    for (int k = 0; k < 4; k++) if (off + k < n) w |= (uint32_t)b[off + k] << (8 * k);
    return w;
}

__global__ void __launch_bounds__(TNC_BLOCK, 2)
tnc_scan_kernel(const uint8_t *__restrict__ b, size_t n, const TncDevState *__restrict__ st_in,
                unsigned long long *__restrict__ counts, uint32_t *__restrict__ exc_count,
                uint32_t *__restrict__ exc, uint32_t exc_cap, uint32_t *__restrict__ seg_nobase)
{
    uint32_t cnt[64];
#pragma unroll
    for (int i = 0; i < 64; i++) cnt[i] = 0;

    const size_t n_chunks = (n + TNC_BPT - 1) / TNC_BPT;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t chunk = (size_t)blockIdx.x * blockDim.x + threadIdx.x; chunk < n_chunks; chunk += stride) {
        const size_t p0 = chunk * TNC_BPT;
        uint32_t w[8];
        if (p0 + TNC_BPT <= n) {
            const uint4 v0 = *reinterpret_cast<const uint4 *>(b + p0);
            const uint4 v1 = *reinterpret_cast<const uint4 *>(b + p0 + 16);
            w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = ld_word(b, n, p0 + 4 * j);
        }
        // the 4 bytes before the chunk; for the very first chunk they come from the carried state
        uint32_t hw;
        bool first = (p0 == 0);
        if (!first) hw = *reinterpret_cast<const uint32_t *>(b + p0 - 4);
        else {
            TncDevState s = *st_in;
            hw = s.started ? ((uint32_t)s.prev[0] << 8 | (uint32_t)s.prev[1] << 16 | (uint32_t)s.prev[2] << 24)
                           : 0x0A0A0A00u;                      // start of file behaves like "\n\n\n"
        }
        Planes P = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            WordFlags f = classify(w[j]);
            P.a |= f.a >> (7 - j); P.c |= f.c >> (7 - j); P.g |= f.g >> (7 - j);
            P.t |= f.t >> (7 - j); P.nl |= f.nl >> (7 - j); P.gt |= f.gt >> (7 - j);
        }
        // the shifts above drag bit 7 of byte k+1.. into lower bits of neighbours only for j<7:
        // f >> (7-j) moves bit 7 of byte k to bit j of byte k -- never across a byte. (f has only bit-7 flags.)
        const WordFlags hf = classify(hw);
        // Straddle windows need the last byte of the previous record (position p-3).  For the
        // first chunk of a piece that byte lies in the previous piece: never trust it here, send
        // the window to the exception path instead (the carried state resolves it).
        const uint32_t keep = first ? 0u : 0xFFFFFFFFu;

        const uint32_t V   = P.a | P.c | P.g | P.t;
        // base-free chunks are counted per 4 KiB segment: the fix-up kernel jumps over base-free stretches (N blocks)
        if (V == 0u) atomicAdd(&seg_nobase[p0 >> TNC_SEG_SHIFT], 1u);
        const uint32_t hV  = hf.a | hf.c | hf.g | hf.t;
        const uint32_t NL2 = back2(P.nl, halo2(hf.nl));        // newline at p-2
        const uint32_t V1  = back1(V, halo1(hV));
        const uint32_t V3c = back3(V, halo3(hV & keep));       // base at p-3, usable as a carry
        // exceptions ---------------------------------------------------------------
        uint32_t x2 = NL2 & V1 & V & ~V3c;                     // straddle window whose carry is not the byte before the newline
        uint32_t x1 = P.gt & back1(P.nl, halo1(hf.nl));        // header record starts here
        while (x2 | x1) {
            uint32_t m = x2 ? x2 : x1;
            uint32_t kind = x2 ? 0u : EXC_HEADER;
            int bit = __ffs(m) - 1;
            if (x2) x2 &= x2 - 1; else x1 &= x1 - 1;
            size_t p = p0 + 4 * (bit & 7) + (bit >> 3);
            if (p < n) {
                uint32_t slot = atomicAdd(exc_count, 1u);
                if (slot < exc_cap) exc[slot] = (uint32_t)p | kind;
            }
        }
        // counting -----------------------------------------------------------------
        uint32_t X[4], Y[4], Z[4];
        {
            const uint32_t pl[4]  = {P.a, P.c, P.g, P.t};      // internal order A,C,G,T
            const uint32_t hpl[4] = {hf.a, hf.c, hf.g, hf.t};
#pragma unroll
            for (int x = 0; x < 4; x++) {
                Z[x] = pl[x];
                Y[x] = back1(pl[x], halo1(hpl[x]));
                X[x] = back2(pl[x], halo2(hpl[x])) | (NL2 & back3(pl[x], halo3(hpl[x] & keep)));
            }
        }
#pragma unroll
        for (int x = 0; x < 4; x++)
#pragma unroll
            for (int y = 0; y < 4; y++) {
                const uint32_t xy = X[x] & Y[y];
#pragma unroll
                for (int z = 0; z < 4; z++) cnt[16 * x + 4 * y + z] += __popc(xy & Z[z]);
            }
    }
    // fold: warp shuffle reduce, one 64-bit global atomic per warp and context
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t v = cnt[i];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&counts[i], (unsigned long long)v);
    }
}

// ---- exception resolution ----------------------------------------------------------------
// All helpers below are called by a full warp with identical arguments.

__device__ __forceinline__ uint8_t byte_at(const uint8_t *b, const TncDevState &st, long i)
{
    return i >= 0 ? b[i] : st.prev[3 + i];       // i in [-3,-1] reads the bytes before the piece
}

// Looks backwards from index e-1 for the nearest '\n'.  a = its index + 1, or -1 when the
// record began before the piece.  has_base: an upper-case base lies in [max(a,0), e).
__device__ void scan_back(const uint8_t *b, long e, long &a, bool &has_base)
{
    const int lane = threadIdx.x & 31;
    has_base = false;
    long hi = e;
    a = -1;
    while (hi > 0) {
        long idx = hi - 1 - lane;
        uint8_t c = idx >= 0 ? b[idx] : (uint8_t)0xFF;
        unsigned m_nl = __ballot_sync(0xffffffffu, c == '\n');
        unsigned m_b  = __ballot_sync(0xffffffffu, is_base(c));
        if (m_nl) {
            int l = __ffs(m_nl) - 1;
            has_base |= (m_b & ((1u << l) - 1u)) != 0;
            a = hi - l;
            return;
        }
        has_base |= m_b != 0;
        hi -= 32;
    }
}

// Looks forwards from index q for the nearest '\n'; returns n when there is none.
__device__ long scan_fwd_newline(const uint8_t *b, long n, long q)
{
    const int lane = threadIdx.x & 31;
    for (long lo = q; lo < n; lo += 32) {
        long idx = lo + lane;
        unsigned m = __ballot_sync(0xffffffffu, idx < n && b[idx] == '\n');
        if (m) return lo + __ffs(m) - 1;
    }
    return n;
}

// Last byte of the nearest kept, newline-terminated record that ends before line start q
// (0 when there is none): the reference's 1-byte carry (tncCountsProfile.c:441-443).
// every 32-byte chunk of segment s is free of upper-case bases
__device__ __forceinline__ bool seg_free(const uint32_t *seg, long n, long s)
{
    const long left = n - (s << TNC_SEG_SHIFT);
    const uint32_t chunks = left >= (1L << TNC_SEG_SHIFT) ? (1u << (TNC_SEG_SHIFT - 5)) : (uint32_t)((left + 31) >> 5);
    return seg[s] == chunks;
}

__device__ uint8_t resolve_carry(const uint8_t *b, long n, const uint32_t *seg, const TncDevState &st, long q)
{
    long e = q - 1;                               // the newline that terminates the previous record
    for (;;) {
        if (e < 0) return st.carry;               // that newline lies before this piece
        // records that lie completely inside a base-free stretch cannot be kept: jump to the record that straddles its start
        if (seg && e >= (2L << TNC_SEG_SHIFT)) {
            long s = (e - 1) >> TNC_SEG_SHIFT;
            if (seg_free(seg, n, s) && seg_free(seg, n, s - 1)) {
                while (s > 0 && seg_free(seg, n, s - 1)) s--;
                const long z0 = s << TNC_SEG_SHIFT;
                const long f = scan_fwd_newline(b, n, z0);      // first newline inside the stretch (e itself at the latest)
                if (f < e) e = f;
            }
        }
        long a; bool hb;
        scan_back(b, e, a, hb);
        if (a >= 0) {
            if (e > a && b[a] != '>' && hb) return b[e - 1];
            e = a - 1;
        } else {
            bool nonempty = st.frag_nonempty || e > 0;
            uint8_t first = st.frag_nonempty ? st.frag_first : b[0];
            hb = hb || st.frag_has_base;
            uint8_t last = e > 0 ? b[e - 1] : st.prev[2];
            if (nonempty && first != '>' && hb) return last;
            return st.carry;
        }
    }
}

__device__ __forceinline__ void bump(unsigned long long *counts, uint8_t x, uint8_t y, uint8_t z, long long delta)
{
    atomicAdd(&counts[16 * ref_code(x) + 4 * ref_code(y) + ref_code(z)], (unsigned long long)delta);
}

// A header record occupies [q, f) (q may be -1: it began before the piece).  Undo what the scan
// counted optimistically inside and right after it.
__device__ void fix_header(const uint8_t *b, long n, const uint32_t *seg, const TncDevState &st, unsigned long long *counts, long q)
{
    const int lane = threadIdx.x & 31;
    long f = scan_fwd_newline(b, n, q < 0 ? 0 : q);
    // (1) windows inside the header were counted as in-line windows
    for (long lo = (q < 0 ? 0 : q); lo < f; lo += 32) {
        long i = lo + lane;
        if (i < f) {
            uint8_t c0 = byte_at(b, st, i - 2), c1 = byte_at(b, st, i - 1), c2 = b[i];
            if (is_base(c0) && is_base(c1) && is_base(c2)) bump(counts, c0, c1, c2, -1);
        }
    }
    // (2) the record after the header: the scan used the header's last byte as the carry
    if (f + 2 < n && f - 1 >= 0 && is_base(b[f - 1]) && is_base(b[f + 1]) && is_base(b[f + 2])) {
        uint8_t carry = resolve_carry(b, n, seg, st, f + 1);
        if (lane == 0) {
            bump(counts, b[f - 1], b[f + 1], b[f + 2], -1);
            if (is_base(carry)) bump(counts, carry, b[f + 1], b[f + 2], +1);
        }
    }
}

__global__ void __launch_bounds__(128)
tnc_fixup_kernel(const uint8_t *__restrict__ b, size_t n_, const TncDevState *__restrict__ st_in,
                 TncDevState *__restrict__ st_out, unsigned long long *__restrict__ counts,
                 const uint32_t *__restrict__ exc_count, uint32_t *__restrict__ ovf,
                 const uint32_t *__restrict__ exc, uint32_t exc_cap, const uint32_t *__restrict__ seg)
{
    const long n = (long)n_;
    TncDevState st = *st_in;
    if (!st.started) { st.prev[0] = st.prev[1] = st.prev[2] = '\n'; st.carry = 0; st.frag_nonempty = 0; st.frag_first = 0; st.frag_has_base = 0; }
    const int lane = threadIdx.x & 31;
    const long warps = ((long)gridDim.x * blockDim.x) >> 5;
    uint32_t nexc = *exc_count;
    if (nexc > exc_cap || *ovf) {                 // list overflowed: flag it, the host redoes the call in smaller pieces
        if (blockIdx.x == 0 && threadIdx.x == 0) *ovf = 1;
        return;
    }
    // tasks: [0, nexc) exceptions, nexc = implicit header at the start of the piece, nexc+1 = state out
    for (long task = (((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); task < (long)nexc + 2; task += warps) {
        if (task < nexc) {
            uint32_t rec = exc[task];
            long p = (long)(rec & ~EXC_HEADER);
            if (rec & EXC_HEADER) fix_header(b, n, seg, st, counts, p);
            else {
                uint8_t carry = resolve_carry(b, n, seg, st, p - 1);
                if (lane == 0 && is_base(carry)) bump(counts, carry, byte_at(b, st, p - 1), b[p], +1);
            }
        } else if (task == nexc) {
            if (st.frag_nonempty && st.frag_first == '>') fix_header(b, n, seg, st, counts, -1);
        } else if (st_out) {
            long a; bool hb;
            scan_back(b, n, a, hb);
            TncDevState o;
            o.started = 1;
            if (a >= 0) {
                o.carry = resolve_carry(b, n, seg, st, a);
                o.frag_nonempty = n > a;
                o.frag_first = n > a ? b[a] : 0;
                o.frag_has_base = hb;
            } else {
                o.carry = st.carry;
                o.frag_nonempty = st.frag_nonempty || n > 0;
                o.frag_first = st.frag_nonempty ? st.frag_first : (n > 0 ? b[0] : 0);
                o.frag_has_base = st.frag_has_base || hb;
            }
            for (int k = 0; k < 3; k++) o.prev[k] = byte_at(b, st, n - 3 + k < -3 ? -3 : n - 3 + k);
            if (n < 3) {                           // fewer than 3 new bytes: slide the old ones
                uint8_t all[6] = {st.prev[0], st.prev[1], st.prev[2], 0, 0, 0};
                for (long k = 0; k < n; k++) all[3 + k] = b[k];
                for (int k = 0; k < 3; k++) o.prev[k] = all[n + k];
            }
            if (lane == 0) *st_out = o;
        }
    }
}

// reverse complement folding order of the reference's printf block (tncCountsProfile.c:452-483)
const char *const OUT_ORDER[32] = {
    "ACA","ACC","ACG","ACT","ATA","ATC","ATG","ATT","CCA","CCC","CCG","CCT","CTA","CTC","CTG","CTT",
    "GCA","GCC","GCG","GCT","GTA","GTC","GTG","GTT","TCA","TCC","TCG","TCT","TTA","TTC","TTG","TTT"};

// device scratch: scanner state (ping-pong), per-call accumulators, overflow flag, exception list
struct TncScratch {
    TncDevState        *st[2];
    uint32_t           *exc_count;
    uint32_t           *ovf;         // set by the fix-up kernel when the exception list overflowed
    unsigned long long *acc;         // 64 per-call accumulators
    uint32_t           *exc;
    uint32_t            exc_cap;
    uint32_t           *seg;         // per 4 KiB segment: number of 32-byte chunks without an upper-case base
    size_t              seg_words;
    uint8_t            *tail;        // first byte after the scratch arrays (256-byte aligned)
};

size_t tnc_exc_cap(size_t piece_bytes) { return piece_bytes / 16 + (1u << 16); }

int tnc_scratch(ssb_ctx *ctx, size_t piece_bytes, size_t extra_bytes, TncScratch *s)
{
    size_t cap = tnc_exc_cap(piece_bytes);
    size_t seg_words = (piece_bytes >> TNC_SEG_SHIFT) + 2;
    size_t head = 1024 + cap * sizeof(uint32_t);
    head = (head + 255) & ~(size_t)255;
    const size_t seg_off = head;
    head += seg_words * sizeof(uint32_t);
    head = (head + 255) & ~(size_t)255;
    int r = ssb_scratch_reserve(ctx, head + extra_bytes);
    if (r) return r;
    uint8_t *base = (uint8_t *)ctx->scratch;
    s->st[0] = (TncDevState *)(base);
    s->st[1] = (TncDevState *)(base + 64);
    s->exc_count = (uint32_t *)(base + 128);
    s->ovf = (uint32_t *)(base + 192);
    s->acc = (unsigned long long *)(base + 256);
    s->exc = (uint32_t *)(base + 1024);
    s->exc_cap = (uint32_t)cap;
    s->seg = (uint32_t *)(base + seg_off);
    s->seg_words = seg_words;
    s->tail = base + head;
    return SSB_OK;
}

__global__ void tnc_final_kernel(const unsigned long long *__restrict__ acc, const uint32_t *__restrict__ ovf,
                                 unsigned long long *__restrict__ counts)
{
    if (*ovf == 0) counts[threadIdx.x] += acc[threadIdx.x];
}

// One piece (n <= TNC_MAX_PIECE) already in HBM; state flows st_in -> st_out on the device.
int tnc_piece(ssb_ctx *ctx, cudaStream_t stream, const uint8_t *d, size_t n, const TncScratch &s,
              TncDevState *st_in, TncDevState *st_out)
{
    SSB_CUDA(ctx, cudaMemsetAsync(s.exc_count, 0, sizeof(uint32_t), stream));
    SSB_CUDA(ctx, cudaMemsetAsync(s.seg, 0, ((n >> TNC_SEG_SHIFT) + 2) * sizeof(uint32_t), stream));
    size_t n_chunks = (n + TNC_BPT - 1) / TNC_BPT;
    int grid = (int)((n_chunks + TNC_BLOCK - 1) / TNC_BLOCK);
    int max_grid = ctx->sm_count * 2 * 4;        // 2 resident blocks per SM, 4 block slots of work each
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    SSB_LAUNCH_P(ctx, SSB_K_TNC_SCAN, tnc_scan_kernel, grid, TNC_BLOCK, 0, stream, d, n, st_in, s.acc, s.exc_count, s.exc, s.exc_cap, s.seg);
    int fgrid = ctx->sm_count * 4;
    SSB_LAUNCH_P(ctx, SSB_K_TNC_FIXUP, tnc_fixup_kernel, fgrid, 128, 0, stream, d, n, st_in, st_out, s.acc, s.exc_count, s.ovf, s.exc, s.exc_cap, s.seg);
    return SSB_OK;
}

int tnc_begin(ssb_ctx *ctx, cudaStream_t stream, const TncScratch &s, const ssb_tnc_carry *carry_in)
{
    static const ssb_tnc_carry zero = {0, {0, 0, 0}, 0, 0, 0, 0};
    SSB_CUDA(ctx, cudaMemsetAsync(s.exc_count, 0, 128 + 64 * sizeof(unsigned long long), stream));   // count, ovf, acc
    SSB_CUDA(ctx, cudaMemcpyAsync(s.st[0], carry_in ? carry_in : &zero, sizeof(ssb_tnc_carry), cudaMemcpyHostToDevice, stream));
    return SSB_OK;
}

// Device-resident buffer of any size, cut into pieces of at most `piece` bytes.
int tnc_run_device(ssb_ctx *ctx, const uint8_t *d, size_t n, size_t piece, const TncScratch &s, int *cur_io)
{
    int cur = *cur_io;
    size_t off = 0;
    do {
        size_t len = n - off < piece ? n - off : piece;
        int r = tnc_piece(ctx, ctx->stream, d + off, len, s, s.st[cur], s.st[cur ^ 1]);
        if (r) return r;
        cur ^= 1;
        off += len;
    } while (off < n);
    *cur_io = cur;
    return SSB_OK;
}

} // namespace

// ---- C ABI ---------------------------------------------------------------------------------

extern "C" int ssb_tnc_count_device(ssb_ctx *ctx, const uint8_t *d_fasta, size_t n, const ssb_tnc_carry *carry_in,
                                    ssb_tnc_carry *carry_out, int64_t *d_counts64)
{
    if (!ctx || (!d_fasta && n) || !d_counts64) return SSB_E_ARG;
    if (((uintptr_t)d_fasta & 15) != 0) return SSB_E_ARG;          // 16-byte vector loads
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t piece = n < TNC_MAX_PIECE ? n : TNC_MAX_PIECE;
    for (int attempt = 0; attempt < 2; attempt++) {
        TncScratch s;
        int r = tnc_scratch(ctx, piece, 0, &s);
        if (r) return r;
        // attempt 1 (after an overflow): pieces so small that the list holds their worst case
        size_t use = attempt == 0 ? piece : TNC_SAFE_PIECE;
        if ((r = tnc_begin(ctx, ctx->stream, s, carry_in))) return r;
        int cur = 0;
        if ((r = tnc_run_device(ctx, d_fasta, n, use, s, &cur))) return r;
        SSB_LAUNCH(ctx, tnc_final_kernel, 1, 64, 0, ctx->stream, s.acc, s.ovf, (unsigned long long *)d_counts64);
        uint32_t ovf = 0;
        ssb_tnc_carry out;
        SSB_CUDA(ctx, cudaMemcpyAsync(&ovf, s.ovf, sizeof ovf, cudaMemcpyDeviceToHost, ctx->stream));
        SSB_CUDA(ctx, cudaMemcpyAsync(&out, s.st[cur], sizeof out, cudaMemcpyDeviceToHost, ctx->stream));
        SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!ovf) { if (carry_out) *carry_out = out; return SSB_OK; }
    }
    snprintf(ctx->err, sizeof ctx->err, "tnc: exception list overflow in safe mode");
    return SSB_E_FORMAT;
}

extern "C" int ssb_tnc_count_host(ssb_ctx *ctx, const uint8_t *fasta, size_t n, const ssb_tnc_carry *carry_in,
                                  ssb_tnc_carry *carry_out, int64_t counts64[64])
{
    if (!ctx || (!fasta && n) || !counts64) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t chunk = (size_t)64 << 20;
    if (const char *e = getenv("SSB_TNC_CHUNK")) { size_t v = strtoull(e, NULL, 10); if (v >= 64) chunk = v; }
    chunk &= ~(size_t)31;
    if (chunk > TNC_MAX_PIECE) chunk = TNC_MAX_PIECE;
    if (chunk > n) chunk = (n + 31) & ~(size_t)31;
    if (chunk < 32) chunk = 32;
    // is the caller's buffer already pinned?  then copy straight out of it
    cudaPointerAttributes attr;
    bool pinned_src = cudaPointerGetAttributes(&attr, fasta) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    for (int attempt = 0; attempt < 2; attempt++) {
        TncScratch s;
        int r = tnc_scratch(ctx, chunk, 2 * chunk, &s);
        if (r) return r;
        uint8_t *dbuf[2] = {s.tail, s.tail + chunk};
        if (!pinned_src) { r = ssb_pinned_reserve(ctx, 2 * chunk); if (r) return r; }
        uint8_t *hbuf[2] = {(uint8_t *)ctx->pinned, (uint8_t *)ctx->pinned + chunk};
        if ((r = tnc_begin(ctx, ctx->stream, s, carry_in))) return r;
        int cur = 0;
        size_t n_chunks = n ? (n + chunk - 1) / chunk : 1;
        // ev[k]: chunk copied into buffer k (copy stream -> compute); ev[2+k]: buffer k consumed (compute -> copy)
        for (size_t i = 0; i < n_chunks; i++) {
            int k = (int)(i & 1);
            size_t off = i * chunk, len = n - off < chunk ? n - off : chunk;
            if (i >= 2) SSB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[2 + k], 0));
            const uint8_t *src = fasta + off;
            if (!pinned_src && len) {
                if (i >= 2) SSB_CUDA(ctx, cudaEventSynchronize(ctx->ev[k]));      // staging buffer k is free again
                memcpy(hbuf[k], src, len);
                src = hbuf[k];
            }
            if (len) SSB_CUDA(ctx, cudaMemcpyAsync(dbuf[k], src, len, cudaMemcpyHostToDevice, ctx->copy_stream));
            SSB_CUDA(ctx, cudaEventRecord(ctx->ev[k], ctx->copy_stream));
            SSB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev[k], 0));
            if ((r = tnc_piece(ctx, ctx->stream, dbuf[k], len, s, s.st[cur], s.st[cur ^ 1]))) return r;
            cur ^= 1;
            SSB_CUDA(ctx, cudaEventRecord(ctx->ev[2 + k], ctx->stream));
        }
        uint32_t ovf = 0;
        ssb_tnc_carry out;
        SSB_CUDA(ctx, cudaMemcpyAsync(counts64, s.acc, 64 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        SSB_CUDA(ctx, cudaMemcpyAsync(&ovf, s.ovf, sizeof ovf, cudaMemcpyDeviceToHost, ctx->stream));
        SSB_CUDA(ctx, cudaMemcpyAsync(&out, s.st[cur], sizeof out, cudaMemcpyDeviceToHost, ctx->stream));
        SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!ovf) { if (carry_out) *carry_out = out; return SSB_OK; }
        // adversarial input (a header every few bytes): redo with chunks whose worst case fits the list
        chunk = TNC_SAFE_PIECE;
    }
    snprintf(ctx->err, sizeof ctx->err, "tnc: exception list overflow in safe mode");
    return SSB_E_FORMAT;
}

// Host restatement of the *state transition only* (no counting): lets a caller cut a FASTA into
// shards for several GPUs.  Mirrors the state-out task of tnc_fixup_kernel.
extern "C" int ssb_tnc_carry_after(const uint8_t *b, size_t n_, const ssb_tnc_carry *carry_in, ssb_tnc_carry *out)
{
    if (!out || (!b && n_)) return SSB_E_ARG;
    ssb_tnc_carry st;
    if (carry_in && carry_in->started) st = *carry_in;
    else { memset(&st, 0, sizeof st); st.prev[0] = st.prev[1] = st.prev[2] = '\n'; }
    long n = (long)n_;
    auto scan_back_h = [&](long e, long &a, bool &hb) {
        hb = false; a = -1;
        for (long i = e - 1; i >= 0; i--) { if (b[i] == '\n') { a = i + 1; return; } if (is_base(b[i])) hb = true; }
    };
    auto resolve = [&](long q) -> uint8_t {
        long e = q - 1;
        for (;;) {
            if (e < 0) return st.carry;
            long a; bool hb; scan_back_h(e, a, hb);
            if (a >= 0) { if (e > a && b[a] != '>' && hb) return b[e - 1]; e = a - 1; }
            else {
                bool nonempty = st.frag_nonempty || e > 0;
                uint8_t first = st.frag_nonempty ? st.frag_first : b[0];
                hb = hb || st.frag_has_base;
                uint8_t last = e > 0 ? b[e - 1] : st.prev[2];
                return (nonempty && first != '>' && hb) ? last : st.carry;
            }
        }
    };
    ssb_tnc_carry o; memset(&o, 0, sizeof o);
    o.started = 1;
    long a; bool hb; scan_back_h(n, a, hb);
    if (a >= 0) { o.carry = resolve(a); o.frag_nonempty = n > a; o.frag_first = n > a ? b[a] : 0; o.frag_has_base = hb; }
    else {
        o.carry = st.carry; o.frag_nonempty = st.frag_nonempty || n > 0;
        o.frag_first = st.frag_nonempty ? st.frag_first : (n > 0 ? b[0] : 0);
        o.frag_has_base = st.frag_has_base || hb;
    }
    uint8_t all[6] = {st.prev[0], st.prev[1], st.prev[2], 0, 0, 0};
    if (n >= 3) { o.prev[0] = b[n - 3]; o.prev[1] = b[n - 2]; o.prev[2] = b[n - 1]; }
    else { for (long k = 0; k < n; k++) all[3 + k] = b[k]; for (int k = 0; k < 3; k++) o.prev[k] = all[n + k]; }
    *out = o;
    return SSB_OK;
}

extern "C" int ssb_tnc_format(const int64_t counts64[64], char *dst, size_t cap)
{
    if (!counts64 || !dst) return SSB_E_ARG;
    size_t used = 0;
    for (int i = 0; i < 32; i++) {
        const char *s = OUT_ORDER[i];
        int fwd = 16 * ref_code(s[0]) + 4 * ref_code(s[1]) + ref_code(s[2]);
        int rev = 16 * (3 - ref_code(s[2])) + 4 * (3 - ref_code(s[1])) + (3 - ref_code(s[0]));
        int w = snprintf(dst + used, cap - used, "%s\t%ld\n", s, (long)(counts64[fwd] + counts64[rev]));
        if (w < 0 || (size_t)w >= cap - used) return SSB_E_ARG;
        used += (size_t)w;
    }
    return (int)used;
}
