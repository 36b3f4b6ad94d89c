/*
 * tools/gen_synth.c -- seeded synthetic inputs for tests and bench.
 *
 * Produces the three inputs of the spike path -- a reference FASTA, a coordinate
 * sorted paired-end SAM (ART-like: `=`/`X`/`I`/`D` CIGARs, QNAME `T_X1_<contig>-<n>`
 * as generatePhasedBams.bash:54 prefixes it, flags 99/147/83/163, RNEXT `=`) and a
 * `.spike` table (README.md:529-537) -- from one seed.  Everything is COUNTER BASED
 * (splitmix64 of (seed, stream, index)): fragment i's start, length, errors and
 * qualities depend only on i, so any coordinate range can be generated on its own
 * and the concatenation of ranges is byte-identical to a single run.  That is what
 * lets bench.py fill shards from many host threads and lets N ranks build their own
 * shard without talking to each other.
 *
 * This is test/bench infrastructure, not part of the product path.
 *
 * Library entry points (ctypes, see stochasticsim_b200/synth.py):
 *   synth_ref_contig()      reference bases of one contig
 *   synth_sam_range()       SAM records with POS in [lo, hi) of one contig
 *   synth_spike_table()     the .spike text
 * CLI: gen_synth out=<prefix> [key=value ...]   (writes <prefix>.fa/.sam/.spike)
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct {
    uint64_t seed;
    int      read_len;        /* bases per read                                   */
    double   frag_mean, frag_sd;
    double   coverage;        /* fold coverage over [win_lo, win_hi)              */
    double   sub_rate;        /* per-base substitution (CIGAR X)                  */
    double   indel_rate;      /* per-read probability of one 1-bp I or D          */
    double   n_rate;          /* per-base probability the read base is 'N'        */
    double   q0_rate;         /* per-base probability of base quality 0 ('!')     */
    double   softclip_rate;   /* per-read probability of a trailing soft clip     */
    double   refskip_rate;    /* per-read probability of one N (ref-skip) op      */
    double   filt_rate;       /* per-pair probability of a flag/MAPQ that read_bam drops */
    int      fa_width;        /* FASTA line width                                 */
    double   lower_frac;      /* fraction of reference in soft-masked lowercase runs */
    int      aux_tags;        /* 1: add NM:i / RG:Z aux fields                    */
} synth_params;

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t idx)
{
    return splitmix64(splitmix64(seed ^ (stream * 0xD1342543DE82EF95ULL)) + idx * 0x9E3779B97F4A7C15ULL);
}
static inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

enum { S_REF = 1, S_LOWER = 2, S_START = 3, S_FRAG = 4, S_READ = 5, S_SPIKE = 6 };

/* ---------------------------------------------------------------- reference */

/* Bases [0,len) of contig `cidx`.  N outside [win_lo,win_hi) when mask_outside. */
void synth_ref_contig(const synth_params *p, int cidx, int64_t len, int64_t win_lo, int64_t win_hi,
                      int mask_outside, char *out)
{
    uint64_t s = p->seed ^ ((uint64_t)(cidx + 1) << 40);
    for (int64_t i = 0; i < len; i += 32) {
        uint64_t r = rnd(s, S_REF, (uint64_t)(i >> 5));
        int64_t e = i + 32 < len ? i + 32 : len;
        for (int64_t j = i; j < e; j++, r >>= 2) out[j] = "ACGT"[r & 3];
    }
    if (p->lower_frac > 0) {
        /* soft-masked runs: blocks of 256 bases, each lowercase with prob lower_frac */
        for (int64_t b = 0; b * 256 < len; b++) {
            if (u01(rnd(s, S_LOWER, (uint64_t)b)) < p->lower_frac) {
                int64_t e = (b + 1) * 256 < len ? (b + 1) * 256 : len;
                for (int64_t j = b * 256; j < e; j++) out[j] = (char)(out[j] | 0x20);
            }
        }
    }
    if (mask_outside) {
        for (int64_t j = 0; j < win_lo && j < len; j++) out[j] = 'N';
        for (int64_t j = win_hi; j < len; j++) out[j] = 'N';
    }
}

/* ------------------------------------------------------------------- reads */

typedef struct { int64_t pos; int64_t frag; int mate; } pending_t;   /* heap of second mates */

static int pend_less(const pending_t *a, const pending_t *b)
{
    if (a->pos != b->pos) return a->pos < b->pos;
    return a->frag < b->frag;
}
static void heap_push(pending_t *h, int *n, pending_t v)
{
    int i = (*n)++;
    h[i] = v;
    while (i > 0) {
        int par = (i - 1) >> 1;
        if (!pend_less(&h[i], &h[par])) break;
        pending_t t = h[i]; h[i] = h[par]; h[par] = t;
        i = par;
    }
}
static pending_t heap_pop(pending_t *h, int *n)
{
    pending_t top = h[0];
    h[0] = h[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < *n && pend_less(&h[l], &h[m])) m = l;
        if (r < *n && pend_less(&h[r], &h[m])) m = r;
        if (m == i) break;
        pending_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
    return top;
}

typedef struct { char *buf; size_t len, cap; int overflow; } outbuf;

static inline void ob_put(outbuf *o, const char *s, size_t n)
{
    if (o->len + n > o->cap) { o->overflow = 1; o->len += n; return; }
    if (o->buf) memcpy(o->buf + o->len, s, n);
    o->len += n;
}
static inline void ob_int(outbuf *o, long long v)
{
    char t[24]; int n = 0, neg = v < 0;
    unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
    do { t[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) t[n++] = '-';
    char r[24];
    for (int i = 0; i < n; i++) r[i] = t[n - 1 - i];
    ob_put(o, r, (size_t)n);
}

typedef struct {
    const synth_params *p;
    int cidx;
    const char *cname;
    const char *ref;          /* contig bases */
    int64_t len, win_lo, win_hi;
    int64_t n_frag;           /* fragments over the window */
    int64_t maxfrag;          /* starts stay below win_hi - maxfrag, so nothing is ever clamped */
    double  step;
} contig_ctx;

static int64_t frag_start(const contig_ctx *c, int64_t i)
{
    /* monotone in i: slot i of width `step` plus a jitter inside the slot */
    uint64_t s = c->p->seed ^ ((uint64_t)(c->cidx + 1) << 40);
    double j = u01(rnd(s, S_START, (uint64_t)i));
    int64_t st = c->win_lo + (int64_t)(((double)i + j) * c->step);
    return st;
}
static int frag_len(const contig_ctx *c, int64_t i)
{
    uint64_t s = c->p->seed ^ ((uint64_t)(c->cidx + 1) << 40);
    uint64_t r = rnd(s, S_FRAG, (uint64_t)i);
    /* Box-Muller from two 32-bit halves */
    double u1 = ((double)(r >> 32) + 1.0) / 4294967297.0, u2 = (double)(r & 0xffffffffu) / 4294967296.0;
    double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    int f = (int)lround(c->p->frag_mean + c->p->frag_sd * z);
    if (f < c->p->read_len) f = c->p->read_len;
    if (f > c->maxfrag) f = (int)c->maxfrag;
    return f;
}

/* One SAM line for mate `mate` (0 = left, 1 = right) of fragment i. */
static void emit_read(const contig_ctx *c, int64_t i, int mate, outbuf *o)
{
    const synth_params *p = c->p;
    int L = p->read_len;
    int64_t st = frag_start(c, i);
    int fl = frag_len(c, i);
    int64_t pos_l = st, pos_r = st + fl - L;
    int64_t pos = mate ? pos_r : pos_l;
    uint64_t s = p->seed ^ ((uint64_t)(c->cidx + 1) << 40);
    uint64_t key = (uint64_t)i * 2 + (uint64_t)mate;
    uint64_t rp = rnd(s, S_FRAG, (uint64_t)i ^ 0x8000000000000000ULL);   /* pair-level draws */
    int left_is_r1 = (int)(rp & 1);
    int flag;
    if (mate == 0) flag = left_is_r1 ? 99 : 163;
    else           flag = left_is_r1 ? 147 : 83;
    int mapq = 99;
    double uf = u01(rnd(s, S_FRAG, (uint64_t)i ^ 0x4000000000000000ULL));
    if (uf < p->filt_rate) {
        /* pair-level property that read_bam (stochasticSpike.c:243-268) filters on */
        int kind = (int)((rp >> 8) % 5);
        if (kind == 0) mapq = 7 + (int)((rp >> 16) % 20);        /* MAPQ < 30            */
        else if (kind == 1) flag |= 1024;                         /* duplicate            */
        else if (kind == 2) flag &= ~2;                           /* paired, not proper   */
        else if (kind == 3) flag |= 512;                          /* QC fail              */
        else if (mate == 1) flag |= 256;                          /* secondary (one mate) */
    }

    /* read-level draws */
    uint64_t r0 = rnd(s, S_READ, key * 64);
    int has_indel = u01(r0) < p->indel_rate;
    uint64_t r1 = rnd(s, S_READ, key * 64 + 1);
    int indel_is_ins = (int)(r1 & 1);
    int indel_at = 5 + (int)((r1 >> 8) % (uint64_t)(L - 10));      /* query offset of the event */
    int clip = 0;
    if (u01(rnd(s, S_READ, key * 64 + 2)) < p->softclip_rate) clip = 1 + (int)((r1 >> 32) % 8);
    int has_skip = !has_indel && u01(rnd(s, S_READ, key * 64 + 3)) < p->refskip_rate;
    int skip_at = 10 + (int)((r1 >> 40) % (uint64_t)(L - 20));
    int skip_len = 20 + (int)((r1 >> 48) % 200);
    if (pos + L + skip_len + 2 >= c->len) has_skip = 0;     /* never run past the contig end */

    /* build SEQ/QUAL/CIGAR by walking query offsets */
    char seq[1024], qual[1024], cig[512];
    int ncig = 0, run = 0; char runop = 0;
    int nm = 0;
#define CIG_FLUSH() do { if (run) { ncig += sprintf(cig + ncig, "%d%c", run, runop); run = 0; } } while (0)
#define CIG_ADD(op, n) do { if (runop != (op)) { CIG_FLUSH(); runop = (op); } run += (n); } while (0)
    int64_t rpos = pos;                 /* reference cursor */
    int aligned = L - clip;
    for (int q = 0; q < L; q++) {
        uint64_t rb = rnd(s, S_READ, key * 64 + 8 + (uint64_t)(q >> 1));
        if (q & 1) rb >>= 32;
        double ub = (double)(rb & 0xffffff) / 16777216.0;
        int qv = 2 + (int)((rb >> 24) % 39);                      /* Phred 2..40 */
        if (q >= aligned) {                                        /* soft clipped tail */
            seq[q] = "ACGT"[(rb >> 8) & 3];
            qual[q] = (char)(33 + qv);
            CIG_ADD('S', 1);
            continue;
        }
        if (has_indel && q == indel_at) {
            if (indel_is_ins) {
                seq[q] = "ACGT"[(rb >> 8) & 3];
                qual[q] = (char)(33 + qv);
                CIG_ADD('I', 1); nm++;
                continue;
            } else { CIG_ADD('D', 1); rpos++; nm++; }
        }
        if (has_skip && q == skip_at) { CIG_ADD('N', skip_len); rpos += skip_len; }
        char rb_ref = (rpos < c->len) ? c->ref[rpos] : 'N';
        char up = (char)(rb_ref & ~0x20);
        char b = up;
        int is_x = 0;
        if (ub < p->sub_rate) {
            int k = (int)((rb >> 10) % 3);
            const char *alt = up == 'A' ? "CGT" : up == 'C' ? "AGT" : up == 'G' ? "ACT" : up == 'T' ? "ACG" : "ACG";
            b = alt[k]; is_x = 1;
        } else if (ub < p->sub_rate + p->n_rate) { b = 'N'; is_x = 1; }
        if (up != 'A' && up != 'C' && up != 'G' && up != 'T') is_x = (b != up);
        if (u01(splitmix64(rb + 0x5bd1e995ULL * (uint64_t)(q + 1))) < p->q0_rate) qv = 0;
        seq[q] = b;
        qual[q] = (char)(33 + qv);
        if (is_x) { CIG_ADD('X', 1); nm++; } else CIG_ADD('=', 1);
        rpos++;
    }
    CIG_FLUSH();
    cig[ncig] = 0;

    /* QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL [aux] */
    ob_put(o, "T_X1_", 5); ob_put(o, c->cname, strlen(c->cname)); ob_put(o, "-", 1); ob_int(o, (long long)i + 1);
    ob_put(o, "\t", 1); ob_int(o, flag);
    ob_put(o, "\t", 1); ob_put(o, c->cname, strlen(c->cname));
    ob_put(o, "\t", 1); ob_int(o, (long long)pos + 1);
    ob_put(o, "\t", 1); ob_int(o, mapq);
    ob_put(o, "\t", 1); ob_put(o, cig, (size_t)ncig);
    ob_put(o, "\t=\t", 3); ob_int(o, (long long)(mate ? pos_l : pos_r) + 1);
    ob_put(o, "\t", 1); ob_int(o, mate ? -(long long)fl : (long long)fl);
    ob_put(o, "\t", 1); ob_put(o, seq, (size_t)L);
    ob_put(o, "\t", 1); ob_put(o, qual, (size_t)L);
    if (p->aux_tags) { ob_put(o, "\tNM:i:", 6); ob_int(o, nm); ob_put(o, "\tRG:Z:grp1", 10); }
    ob_put(o, "\n", 1);
}

static void contig_setup(contig_ctx *c, const synth_params *p, int cidx, const char *cname, const char *ref,
                         int64_t len, int64_t win_lo, int64_t win_hi)
{
    c->p = p; c->cidx = cidx; c->cname = cname; c->ref = ref; c->len = len;
    c->win_lo = win_lo; c->win_hi = win_hi;
    c->maxfrag = (int64_t)(p->frag_mean + 8 * p->frag_sd) + p->read_len + 2;
    double span = (double)(win_hi - win_lo - c->maxfrag);
    if (span < 1) span = 1;
    c->n_frag = (int64_t)llround(p->coverage * span / (2.0 * p->read_len));
    if (c->n_frag < 1) c->n_frag = 1;
    c->step = span / (double)c->n_frag;
}

int64_t synth_num_fragments(const synth_params *p, int64_t win_lo, int64_t win_hi)
{
    contig_ctx c; contig_setup(&c, p, 0, "", NULL, win_hi, win_lo, win_hi);
    return c.n_frag;
}

/*
 * SAM records of contig `cidx` whose POS (0-based) lies in [lo, hi), in coordinate
 * order with ties broken by (fragment index, mate).  Returns bytes needed; writes
 * at most `cap` (call with out=NULL/cap=0 to size).  *n_reads gets the record count.
 */
int64_t synth_sam_range(const synth_params *p, int cidx, const char *cname, const char *ref, int64_t len,
                        int64_t win_lo, int64_t win_hi, int64_t lo, int64_t hi,
                        char *out, int64_t cap, int64_t *n_reads)
{
    contig_ctx c; contig_setup(&c, p, cidx, cname, ref, len, win_lo, win_hi);
    outbuf o = { out, 0, out ? (size_t)cap : 0, 0 };
    if (!out) o.cap = (size_t)-1;
    int64_t maxfrag = c.maxfrag;
    /* first fragment that can contribute: start >= lo - maxfrag */
    int64_t i0 = (int64_t)(((double)(lo - maxfrag - c.win_lo)) / c.step) - 2;
    if (i0 < 0) i0 = 0;
    int hcap = 1 << 16, hn = 0;
    pending_t *heap = malloc(sizeof(pending_t) * (size_t)hcap);
    int64_t cnt = 0;
    for (int64_t i = i0; i < c.n_frag; i++) {
        int64_t st = frag_start(&c, i);      /* monotone in i */
        int fl = frag_len(&c, i);
        if (st >= hi) break;
        /* flush pending right mates that sort before this left mate */
        pending_t me = { st, i, 0 };
        while (hn && pend_less(&heap[0], &me)) {
            pending_t t = heap_pop(heap, &hn);
            if (t.pos >= lo && t.pos < hi) { emit_read(&c, t.frag, t.mate, &o); cnt++; }
        }
        if (hn == hcap - 2) { hcap <<= 1; heap = realloc(heap, sizeof(pending_t) * (size_t)hcap); }
        if (st >= lo) { emit_read(&c, i, 0, &o); cnt++; }
        pending_t r = { st + fl - p->read_len, i, 1 };
        heap_push(heap, &hn, r);
    }
    while (hn) {
        pending_t t = heap_pop(heap, &hn);
        if (t.pos >= lo && t.pos < hi) { emit_read(&c, t.frag, t.mate, &o); cnt++; }
    }
    free(heap);
    if (n_reads) *n_reads = cnt;
    return (int64_t)o.len;
}

/* ------------------------------------------------------------------- spike */

/*
 * n loci uniform over [win_lo,win_hi) of one contig, sorted, distinct slots.
 * alt_mode: 0 = always '.', 1 = half explicit base / half '.'.  AF ~ U(af_lo, af_hi).
 */
int64_t synth_spike_table(const synth_params *p, int cidx, const char *cname, int64_t win_lo, int64_t win_hi,
                          int64_t n, int alt_mode, double af_lo, double af_hi, char *out, int64_t cap)
{
    outbuf o = { out, 0, out ? (size_t)cap : 0, 0 };
    if (!out) o.cap = (size_t)-1;
    uint64_t s = p->seed ^ ((uint64_t)(cidx + 1) << 40);
    double step = (double)(win_hi - win_lo) / (double)n;
    for (int64_t i = 0; i < n; i++) {
        uint64_t r = rnd(s, S_SPIKE, (uint64_t)i);
        int64_t locus = win_lo + (int64_t)(((double)i + u01(r)) * step);
        uint64_t r2 = splitmix64(r);
        double af = af_lo + (af_hi - af_lo) * u01(r2);
        char line[128];
        char alt = '.';
        if (alt_mode == 1 && (r2 & 1)) alt = "GCAT"[(r2 >> 1) & 3];
        int w = snprintf(line, sizeof line, "%s\t%lld\t%c\t%.4g\n", cname, (long long)locus + 1, alt, af);
        ob_put(&o, line, (size_t)w);
    }
    return (int64_t)o.len;
}

void synth_default_params(synth_params *p)
{
    memset(p, 0, sizeof *p);
    p->seed = 1; p->read_len = 76; p->frag_mean = 180; p->frag_sd = 10; p->coverage = 25;
    p->sub_rate = 0.001; p->indel_rate = 0.0001; p->fa_width = 60;
}

/* --------------------------------------------------------------------- CLI */
#ifdef GEN_SYNTH_MAIN
static const char *arg(int argc, char **argv, const char *key, const char *def)
{
    size_t kl = strlen(key);
    for (int i = 1; i < argc; i++) if (strncmp(argv[i], key, kl) == 0 && argv[i][kl] == '=') return argv[i] + kl + 1;
    return def;
}

int main(int argc, char **argv)
{
    synth_params p; synth_default_params(&p);
    const char *out = arg(argc, argv, "out", NULL);
    if (!out) {
        fprintf(stderr, "usage: gen_synth out=<prefix> [seed= contigs=name:len[,name:len..] win=lo:hi mask=0|1 read_len= frag_mean= frag_sd=\n"
                        "       coverage= sub= indel= nrate= q0= softclip= refskip= filt= width= lower= aux=0|1 spikes= alt_mode= af=lo:hi sm=<sample>]\n");
        return 2;
    }
    p.seed = strtoull(arg(argc, argv, "seed", "1"), NULL, 10);
    p.read_len = atoi(arg(argc, argv, "read_len", "76"));
    p.frag_mean = atof(arg(argc, argv, "frag_mean", "180"));
    p.frag_sd = atof(arg(argc, argv, "frag_sd", "10"));
    p.coverage = atof(arg(argc, argv, "coverage", "25"));
    p.sub_rate = atof(arg(argc, argv, "sub", "0.001"));
    p.indel_rate = atof(arg(argc, argv, "indel", "0.0001"));
    p.n_rate = atof(arg(argc, argv, "nrate", "0"));
    p.q0_rate = atof(arg(argc, argv, "q0", "0"));
    p.softclip_rate = atof(arg(argc, argv, "softclip", "0"));
    p.refskip_rate = atof(arg(argc, argv, "refskip", "0"));
    p.filt_rate = atof(arg(argc, argv, "filt", "0"));
    p.fa_width = atoi(arg(argc, argv, "width", "60"));
    p.lower_frac = atof(arg(argc, argv, "lower", "0"));
    p.aux_tags = atoi(arg(argc, argv, "aux", "0"));
    int mask = atoi(arg(argc, argv, "mask", "0"));
    int64_t n_spikes = atoll(arg(argc, argv, "spikes", "100"));
    int alt_mode = atoi(arg(argc, argv, "alt_mode", "1"));
    double af_lo = 0.01, af_hi = 0.5;
    sscanf(arg(argc, argv, "af", "0.01:0.5"), "%lf:%lf", &af_lo, &af_hi);
    const char *sm = arg(argc, argv, "sm", NULL);
    char *contigs = strdup(arg(argc, argv, "contigs", "chr19:100000"));
    const char *win = arg(argc, argv, "win", NULL);

    char path[4096];
    snprintf(path, sizeof path, "%s.fa", out);    FILE *ffa = fopen(path, "w");
    snprintf(path, sizeof path, "%s.sam", out);   FILE *fsam = fopen(path, "w");
    snprintf(path, sizeof path, "%s.spike", out); FILE *fsp = fopen(path, "w");
    if (!ffa || !fsam || !fsp) { perror("open"); return 1; }

    /* parse contigs */
    char *names[256]; int64_t lens[256]; int nc = 0;
    for (char *tok = strtok(contigs, ","); tok && nc < 256; tok = strtok(NULL, ",")) {
        char *colon = strchr(tok, ':');
        if (!colon) { fprintf(stderr, "bad contig spec\n"); return 2; }
        *colon = 0; names[nc] = tok; lens[nc] = atoll(colon + 1); nc++;
    }
    fprintf(fsam, "@HD\tVN:1.6\tSO:coordinate\n");
    for (int c = 0; c < nc; c++) fprintf(fsam, "@SQ\tSN:%s\tLN:%lld\n", names[c], (long long)lens[c]);
    if (sm) fprintf(fsam, "@RG\tID:grp1\tSM:%s\n", sm);
    fprintf(fsam, "@PG\tID:gen_synth\tPN:gen_synth\n");
    fprintf(fsp, "# Comment / meta lines begining with at '#' are ignored by stochasticSpike\n#CHROM\tPOS\tALT\tAF\n");

    for (int c = 0; c < nc; c++) {
        int64_t wlo = 0, whi = lens[c];
        if (win) { long long a, b; if (sscanf(win, "%lld:%lld", &a, &b) == 2) { wlo = a; whi = b < lens[c] ? b : lens[c]; } }
        char *ref = malloc((size_t)lens[c] + 1);
        synth_ref_contig(&p, c, lens[c], wlo, whi, mask, ref);
        fprintf(ffa, ">%s\n", names[c]);
        for (int64_t i = 0; i < lens[c]; i += p.fa_width) {
            int64_t e = i + p.fa_width < lens[c] ? i + p.fa_width : lens[c];
            fwrite(ref + i, 1, (size_t)(e - i), ffa);
            fputc('\n', ffa);
        }
        int64_t n_reads = 0;
        int64_t need = synth_sam_range(&p, c, names[c], ref, lens[c], wlo, whi, 0, lens[c], NULL, 0, &n_reads);
        char *buf = malloc((size_t)need + 1);
        synth_sam_range(&p, c, names[c], ref, lens[c], wlo, whi, 0, lens[c], buf, need, &n_reads);
        fwrite(buf, 1, (size_t)need, fsam);
        free(buf);
        if (n_spikes > 0) {
            int64_t sn = synth_spike_table(&p, c, names[c], wlo, whi, n_spikes, alt_mode, af_lo, af_hi, NULL, 0);
            char *sb = malloc((size_t)sn + 1);
            synth_spike_table(&p, c, names[c], wlo, whi, n_spikes, alt_mode, af_lo, af_hi, sb, sn);
            fwrite(sb, 1, (size_t)sn, fsp);
            free(sb);
        }
        fprintf(stderr, "[gen_synth] %s: %lld reads, %lld SAM bytes\n", names[c], (long long)n_reads, (long long)need);
        free(ref);
    }
    fclose(ffa); fclose(fsam); fclose(fsp);
    free(contigs);
    return 0;
}
#endif
