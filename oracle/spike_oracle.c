/*
 * oracle/spike_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, not product).
 *
 * Plain-C restatement of the reference's stochasticSpike (stochasticSpike.c:98-1671)
 * INCLUDING the htslib-1.13 behaviour it leans on (SAM text parsing/formatting,
 * read filter, pileup order, faidx) as described in SURVEY.md App. A/D.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or run it; libssb200.so never does.
 *
 * Pinning: the real binary cannot be built in this image (htslib/samtools are
 * fetched from the network by the reference's install.bash:82-101 and are absent).
 * This restatement is checked byte-for-byte (SAM, truth.vcf, stdout) against
 * oracle/_ref/stochasticSpike = the UNMODIFIED reference source compiled over
 * oracle/shim/ (tests/test_spike_oracle.py), and against the glibc rand() known
 * answers of SURVEY.md App. C.  The htslib layer itself is "parity unpinned".
 *
 * It uses the real glibc srand()/rand() (stochasticSpike.c:948,297,334) and keeps
 * the O(depth) QNAME strcmp mate search (stochasticSpike.c:363-383), so its cost
 * profile mirrors the reference's.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <stdbool.h>

#define MAX_PILEUP_SIZE 10000                 /* stochasticSpike.c:38 */
#define MIN_MAPQ 30                           /* stochasticSpike.c:916 */

static const char NT16_STR[] = "=ACMGRSVTWYHKDBN";
static unsigned char nt16_of(unsigned char c)
{
    switch (toupper(c)) {
    case '=': return 0;  case 'A': return 1;  case 'C': return 2;  case 'M': return 3;
    case 'G': return 4;  case 'R': return 5;  case 'S': return 6;  case 'V': return 7;
    case 'T': return 8;  case 'W': return 9;  case 'Y': return 10; case 'H': return 11;
    case 'K': return 12; case 'D': return 13; case 'B': return 14; default:  return 15;
    }
}

/* ------------------------------------------------------------- containers */

typedef struct {
    char   *qname;
    int     flag, tid, mapq, mtid;
    int64_t pos, mpos, isize;        /* 0-based */
    int     n_cigar;
    uint32_t *cigar;                 /* len<<4 | op, op index in "MIDNSHP=X" */
    int     l_seq;
    char   *seq;                     /* canonical upper-case IUPAC chars (what a 4-bit round trip gives) */
    uint8_t *qual;                   /* phred, 0xff = absent */
    char   *aux;                     /* verbatim optional fields (without leading tab) or NULL */
    int64_t end;                     /* pos + reference length */
} read_t;

typedef struct {
    char   *text; size_t l_text;
    int     n; char **name; int64_t *len;
} header_t;

typedef struct { int n; char **name; char **seq; int64_t *len; } fasta_t;

typedef struct {
    char contig[1024]; int c_tid; long locus; char base; float mutFreq;   /* stochasticSpike.c:88-94 */
} target_t;

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "oom\n"); exit(1); } return q; }

static int hdr_name2tid(const header_t *h, const char *s)
{
    for (int i = 0; i < h->n; i++) if (strcmp(h->name[i], s) == 0) return i;
    return -1;
}

/* ------------------------------------------------------------------ FASTA */
/* faidx_fetch_seq64(whole contig): newlines and other non-graphic bytes dropped, case preserved
 * (stochasticSpike.c:215-219). Contig name = header up to the first white space. */
static int load_fasta(const char *fn, fasta_t *fa)
{
    FILE *fp = fopen(fn, "r");
    if (!fp) return -1;
    memset(fa, 0, sizeof *fa);
    size_t cap = 0, scap = 0; char *line = NULL; size_t lcap = 0; ssize_t n; int cur = -1;
    while ((n = getline(&line, &lcap, fp)) >= 0) {
        if (line[0] == '>') {
            if ((size_t)fa->n == cap) {
                cap = cap ? cap * 2 : 16;
                fa->name = xrealloc(fa->name, cap * sizeof(char *));
                fa->seq = xrealloc(fa->seq, cap * sizeof(char *));
                fa->len = xrealloc(fa->len, cap * sizeof(int64_t));
            }
            size_t e = 1; while (e < (size_t)n && !isspace((unsigned char)line[e])) e++;
            cur = fa->n++;
            fa->name[cur] = strndup(line + 1, e - 1); fa->seq[cur] = NULL; fa->len[cur] = 0; scap = 0;
            continue;
        }
        if (cur < 0) continue;
        if ((size_t)fa->len[cur] + (size_t)n + 1 > scap) {
            scap = scap ? scap : (1u << 20);
            while ((size_t)fa->len[cur] + (size_t)n + 1 > scap) scap *= 2;
            fa->seq[cur] = xrealloc(fa->seq[cur], scap);
        }
        for (ssize_t i = 0; i < n; i++) if (isgraph((unsigned char)line[i])) fa->seq[cur][fa->len[cur]++] = line[i];
    }
    free(line); fclose(fp);
    return 0;
}

/* -------------------------------------------------------------------- SAM */

static int64_t cigar_rlen(const read_t *r)
{
    int64_t l = 0;
    for (int k = 0; k < r->n_cigar; k++) { int op = r->cigar[k] & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) l += r->cigar[k] >> 4; }
    return l;
}
static int64_t cigar_qlen(const read_t *r)
{
    int64_t l = 0;
    for (int k = 0; k < r->n_cigar; k++) { int op = r->cigar[k] & 15; if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) l += r->cigar[k] >> 4; }
    return l;
}

/* one SAM alignment line (no newline) -> read_t; returns <0 when htslib would reject the record */
static int parse_read(const header_t *h, const char *s, size_t n, read_t *r)
{
    const char *f[12]; size_t fl[12]; int nf = 0;
    const char *p = s, *end = s + n;
    while (nf < 12) {
        const char *t = (nf < 11) ? memchr(p, '\t', (size_t)(end - p)) : NULL;
        f[nf] = p; fl[nf] = t ? (size_t)(t - p) : (size_t)(end - p); nf++;
        if (!t) break;
        p = t + 1;
    }
    if (nf < 11) return -1;
    char tmp[64];
#define NUM(i) (memcpy(tmp, f[i], fl[i] < 63 ? fl[i] : 63), tmp[fl[i] < 63 ? fl[i] : 63] = 0, tmp)
    memset(r, 0, sizeof *r);
    r->qname = strndup(f[0], fl[0]);
    r->flag = (int)(strtol(NUM(1), NULL, 0) & 0xffff);
    char *rn = strndup(f[2], fl[2]);
    r->tid = strcmp(rn, "*") == 0 ? -1 : hdr_name2tid(h, rn);
    r->pos = strtoll(NUM(3), NULL, 10) - 1;
    r->mapq = (int)(strtol(NUM(4), NULL, 10) & 0xff);
    if (!(fl[5] == 1 && f[5][0] == '*')) {
        int nc = 0;
        for (size_t i = 0; i < fl[5]; i++) if (!isdigit((unsigned char)f[5][i])) nc++;
        r->cigar = xrealloc(NULL, sizeof(uint32_t) * (size_t)(nc ? nc : 1));
        const char *c = f[5], *ce = f[5] + fl[5];
        for (int k = 0; k < nc; k++) {
            unsigned long len = 0; const char *c0 = c;
            while (c < ce && isdigit((unsigned char)*c)) len = len * 10 + (unsigned long)(*c++ - '0');
            const char *ops = "MIDNSHP=X", *o = (c < ce) ? strchr(ops, *c) : NULL;
            if (c == c0 || !o || !*o) { free(rn); return -1; }
            r->cigar[k] = (uint32_t)(len << 4) | (uint32_t)(o - ops);
            c++;
        }
        r->n_cigar = nc;
    }
    char *mn = strndup(f[6], fl[6]);
    r->mtid = strcmp(mn, "=") == 0 ? r->tid : strcmp(mn, "*") == 0 ? -1 : hdr_name2tid(h, mn);
    free(mn); free(rn);
    r->mpos = strtoll(NUM(7), NULL, 10) - 1;
    r->isize = strtoll(NUM(8), NULL, 10);
    if (!(fl[9] == 1 && f[9][0] == '*')) {
        r->l_seq = (int)fl[9];
        r->seq = xrealloc(NULL, fl[9] + 1);
        for (size_t i = 0; i < fl[9]; i++) r->seq[i] = NT16_STR[nt16_of((unsigned char)f[9][i])];
        r->seq[fl[9]] = 0;
    }
    if (r->n_cigar && r->l_seq && cigar_qlen(r) != r->l_seq) return -1;
    r->qual = xrealloc(NULL, (size_t)r->l_seq + 1);
    if (fl[10] == 1 && f[10][0] == '*') memset(r->qual, 0xff, (size_t)r->l_seq);
    else {
        if ((int)fl[10] != r->l_seq) return -1;
        for (int i = 0; i < r->l_seq; i++) r->qual[i] = (uint8_t)(f[10][i] - 33);
    }
    if (nf == 12 && fl[11]) r->aux = strndup(f[11], fl[11]);
    r->end = r->pos + cigar_rlen(r);
    return 0;
}

/* sam_write1's text form (stochasticSpike.c:273) */
static void write_read(FILE *o, const header_t *h, const read_t *r)
{
    fputs(r->qname, o);
    fprintf(o, "\t%d\t", r->flag);
    if (r->tid >= 0) fputs(h->name[r->tid], o); else fputc('*', o);
    fprintf(o, "\t%lld\t%d\t", (long long)r->pos + 1, r->mapq);
    if (r->n_cigar) for (int k = 0; k < r->n_cigar; k++) fprintf(o, "%u%c", r->cigar[k] >> 4, "MIDNSHP=X"[r->cigar[k] & 15]);
    else fputc('*', o);
    fputc('\t', o);
    if (r->mtid < 0) fputc('*', o); else if (r->mtid == r->tid) fputc('=', o); else fputs(h->name[r->mtid], o);
    fprintf(o, "\t%lld\t%lld\t", (long long)r->mpos + 1, (long long)r->isize);
    if (r->l_seq) {
        fwrite(r->seq, 1, (size_t)r->l_seq, o);
        fputc('\t', o);
        if (r->qual[0] == 0xff) fputc('*', o);
        else for (int i = 0; i < r->l_seq; i++) fputc(r->qual[i] + 33, o);
    } else fputs("*\t*", o);
    if (r->aux) { fputc('\t', o); fputs(r->aux, o); }
    fputc('\n', o);
}

/* read_bam's filter (stochasticSpike.c:243-268) */
static bool read_passes(const read_t *r)
{
    if (r->flag & (4 | 256 | 512 | 1024)) return false;
    if (r->mapq < MIN_MAPQ) return false;
    if ((r->flag & 1) && !(r->flag & 2)) return false;
    return true;
}

/* column geometry of one read at reference position pos (App. D; htslib resolve_cigar semantics) */
static void column_of(const read_t *r, int64_t pos, int *qpos, int *is_del, int *is_refskip)
{
    int64_t x = r->pos; int y = 0;
    *qpos = 0; *is_del = 0; *is_refskip = 0;
    for (int k = 0; k < r->n_cigar; k++) {
        int op = r->cigar[k] & 15; int l = (int)(r->cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) { if (pos < x + l) { *qpos = y + (int)(pos - x); return; } x += l; y += l; }
        else if (op == 2 || op == 3) { if (pos < x + l) { *is_del = 1; *is_refskip = (op == 3); *qpos = y; return; } x += l; }
        else if (op == 1 || op == 4) y += l;
    }
}

/* --------------------------------------------------------- RNG helpers */
/* stochasticSpike.c:283-302 with min=0,max=3: cutoff = (RAND_MAX/4)*4 */
static int random_0_3(void)
{
    int cutoff = (RAND_MAX / 4) * 4, r;
    do { r = rand(); } while (r >= cutoff);
    return r % 4;
}
/* stochasticSpike.c:338-360 */
static char select_mutant_allele(char wild)
{
    static const char bases[] = "GCAT";
    int i;
    do { i = random_0_3(); } while (bases[i] == wild);
    return bases[i];
}
/* stochasticSpike.c:332-335; mutFreq is a float promoted to double at the call (:613) */
static bool coin_toss(double p) { return rand() < (p * ((double)RAND_MAX + 1.0)); }

/* ------------------------------------------------------------- .spike */
/* getNextTarget (stochasticSpike.c:98-158).  Returns false at EOF (contig[0] = 0). */
static bool next_target(target_t *t, FILE *fp, const header_t *h)
{
    char *line = NULL; size_t cap = 0; ssize_t n;
    t->locus = 0;
    while ((n = getline(&line, &cap, fp)) != -1) {
        if (line[0] == '#') continue;
        if (!strlen(line)) continue;
        char *tok = strtok(line, "\t");
        if (!tok) continue;
        strncpy(t->contig, tok, sizeof t->contig - 1); t->contig[sizeof t->contig - 1] = 0;
        t->c_tid = hdr_name2tid(h, t->contig);
        if (!(tok = strtok(NULL, "\t"))) continue;
        t->locus = atol(tok) - 1;
        if (!(tok = strtok(NULL, "\t"))) continue;
        t->base = *tok;
        if (!(tok = strtok(NULL, "\t"))) continue;
        t->mutFreq = (float)atof(tok);
        free(line);
        return true;
    }
    free(line);
    t->contig[0] = 0;
    return false;
}

static void print_no_coverage(FILE *vcf, const target_t *t)
{
    fprintf(vcf, "%s\t%ld\t.\t.\t.\t.\tNO_COVERAGE\t.\t.\n", t->contig, t->locus + 1);   /* :1604-1614, :1632-1642 */
}

/* ------------------------------------------------------------ the run */

typedef struct { long alignmentCount, numberOfLociCovered, totalFoldCoverage, maxDepth; } spike_stats;

enum { F_NONE = 0, F_PASS, F_MASKED, F_MASKED_OVL, F_UNDETECTED, F_NO_COVERAGE };
static const char *const FILTER_NAME[] = {"NONE", "PASS", "MASKED", "MASKED_OVL", "UNDETECTED", "NO_COVERAGE"};

typedef struct { int read; int qpos, is_del, is_refskip; } col_entry;

/* error-allele tallies in the reference's fixed G,C,A,T order (:1207) */
static void err_inc(int cnt[4], char b) { const char *o = "GCAT"; for (int i = 0; i < 4; i++) if (o[i] == b) { cnt[i]++; return; } }

/* stable bubble sort of indices by descending count (:500-523) */
static void err_order(const int cnt[4], int idx[4])
{
    for (int i = 0; i < 4; i++) idx[i] = i;
    for (int step = 0; step < 3; ++step) {
        bool sw = false;
        for (int i = 0; i < 3 - step; ++i) if (cnt[idx[i]] < cnt[idx[i + 1]]) { int t = idx[i]; idx[i] = idx[i + 1]; idx[i + 1] = t; sw = true; }
        if (!sw) break;
    }
}
static void print_err_list(FILE *vcf, const int cnt[4], const int idx[4], bool as_counts)
{
    for (int i = 0; i < 4; i++) {
        if (cnt[idx[i]] == 0) break;
        if (as_counts) fprintf(vcf, "%d", cnt[idx[i]]); else fprintf(vcf, "%c", "GCAT"[idx[i]]);
        if (i + 1 == 4 || cnt[idx[i + 1]] == 0) break;
        fprintf(vcf, ",");
    }
}

/*
 * cmd = basename(argv[0]) as printed in the VCF header (:931-936,:1020-1021);
 * a1..a5 = argv[1..5] as echoed at :1024-1030.  Returns the process exit status
 * the reference would produce.
 */
int spike_oracle_run(const char *cmd, const char *a1, const char *a2, const char *a3, const char *a4, const char *a5,
                     const char *vcf_path, FILE *statsout, spike_stats *st_out)
{
    FILE *vcf = fopen(vcf_path, "w");                                    /* :944 */
    srand((unsigned)atoi(a4));                                           /* :948 */

    FILE *in = strcmp(a1, "-") == 0 ? stdin : fopen(a1, "r");
    if (!in) { fprintf(stderr, "Couldn't open bam...\n"); return 1; }    /* :962-965 */
    header_t h; memset(&h, 0, sizeof h);
    size_t tcap = 0, hcap = 0;
    char *line = NULL; size_t lcap = 0; ssize_t n;
    bool have_line = false;
    while ((n = getline(&line, &lcap, in)) >= 0) {
        if (line[0] != '@') { have_line = true; break; }
        if (h.l_text + (size_t)n + 2 > hcap) { hcap = hcap ? hcap * 2 : 65536; while (h.l_text + (size_t)n + 2 > hcap) hcap *= 2; h.text = xrealloc(h.text, hcap); }
        memcpy(h.text + h.l_text, line, (size_t)n); h.l_text += (size_t)n;
        if (line[n - 1] != '\n') h.text[h.l_text++] = '\n';
        if (strncmp(line, "@SQ\t", 4) == 0) {
            char *copy = strndup(line, (size_t)n), *save = NULL, *sn = NULL; int64_t ln = 0;
            for (char *tok = strtok_r(copy, "\t\r\n", &save); tok; tok = strtok_r(NULL, "\t\r\n", &save)) {
                if (!strncmp(tok, "SN:", 3)) sn = tok + 3; else if (!strncmp(tok, "LN:", 3)) ln = strtoll(tok + 3, NULL, 10);
            }
            if (sn) {
                if ((size_t)h.n == tcap) { tcap = tcap ? tcap * 2 : 64; h.name = xrealloc(h.name, tcap * sizeof(char *)); h.len = xrealloc(h.len, tcap * sizeof(int64_t)); }
                h.name[h.n] = strdup(sn); h.len[h.n] = ln; h.n++;
            }
            free(copy);
        }
    }
    FILE *out = fopen(a5, "w");                                          /* :980 */
    if (!out) { fprintf(stderr, "Couldn't write out ...\n"); return 1; }
    static char obuf[1 << 22]; setvbuf(out, obuf, _IOFBF, sizeof obuf);
    if (h.l_text) fwrite(h.text, 1, h.l_text, out);                      /* :986 header verbatim */

    /* sample name = SM of the first @RG (:998-1014) */
    const char *samp = "SAMPLE"; char *smbuf = NULL;
    {
        const char *p = h.text, *e = h.text ? h.text + h.l_text : NULL;
        while (p && p < e) {
            const char *nl = memchr(p, '\n', (size_t)(e - p)); size_t l = nl ? (size_t)(nl - p) : (size_t)(e - p);
            if (l >= 4 && !strncmp(p, "@RG\t", 4)) {
                const char *q = p + 3, *le = p + l; bool found = false;
                while (q < le) {
                    if (*q == '\t') q++;
                    const char *t = memchr(q, '\t', (size_t)(le - q)); size_t fl = t ? (size_t)(t - q) : (size_t)(le - q);
                    if (fl >= 3 && q[0] == 'S' && q[1] == 'M' && q[2] == ':') { while (fl > 3 && q[fl - 1] == '\r') fl--; smbuf = strndup(q + 3, fl - 3); found = true; break; }
                    q += fl;
                }
                if (!found) { fprintf(stderr, "Couldn't read sample name from header...\n"); return 255; }   /* exit(-1), :1006-1009 */
                break;
            }
            p += l + 1;
        }
        if (smbuf) samp = smbuf;
    }
    fprintf(vcf, "##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n");   /* :1017-1018 */
    fprintf(vcf, "##%sVersion=%s\n##%sCommand=%s %s %s %s %s\n", cmd, "0.01", cmd, a1, a2, a3, a4, a5);
    fprintf(vcf, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n", samp);

    fasta_t fa;
    if (load_fasta(a2, &fa) < 0) { fprintf(stderr, "Could not load faidx: %s\n", a2); return 1; }   /* :1042-1046 */
    FILE *cfg = fopen(a3, "r");
    if (!cfg) { fprintf(stderr, "\nCan't open %s..\n\n", a3); return 255; }                          /* return -1, :1050-1053 */
    int *fa_of_tid = xrealloc(NULL, sizeof(int) * (size_t)(h.n + 1));
    for (int t = 0; t < h.n; t++) { fa_of_tid[t] = -1; for (int i = 0; i < fa.n; i++) if (!strcmp(fa.name[i], h.name[t])) { fa_of_tid[t] = i; break; } }

    target_t tgt; memset(&tgt, 0, sizeof tgt); strcpy(tgt.contig, "EMPTY");                           /* :1057 */
    next_target(&tgt, cfg, &h);                                                                         /* :1066 */

    /* pileup state: list of live reads in arrival order (App. D) */
    size_t rcap = 1024, rn = 0;
    read_t **live = xrealloc(NULL, rcap * sizeof(read_t *));
    col_entry *col = xrealloc(NULL, sizeof(col_entry) * (MAX_PILEUP_SIZE + 64));
    size_t colcap = MAX_PILEUP_SIZE + 64;
    int *handled = xrealloc(NULL, sizeof(int) * colcap);
    int cur_tid = 0, max_tid = -1; int64_t cur_pos = 0, max_pos = -1;
    bool eof = false; int status = 0;
    long alignmentCount = 0, nLoci = 0, fold = 0, maxDepth = 0;

    for (;;) {
        /* pull reads until look-ahead proves column (cur_tid,cur_pos) complete */
        while (!eof && !(max_tid > cur_tid || (max_tid == cur_tid && max_pos > cur_pos))) {
            read_t *r = NULL;
            for (;;) {
                if (!have_line) { n = getline(&line, &lcap, in); if (n < 0) { eof = true; break; } }
                have_line = false;
                size_t l = (size_t)n; while (l && (line[l - 1] == '\n' || line[l - 1] == '\r')) l--;
                if (!l) continue;
                r = xrealloc(NULL, sizeof *r);
                if (parse_read(&h, line, l, r) < 0) { fprintf(stderr, "[oracle] malformed SAM record\n"); status = -2; eof = true; r = NULL; break; }
                if (!read_passes(r)) { r = NULL; continue; }          /* leaked on purpose: tiny, test tool */
                break;
            }
            if (!r) break;
            if (r->tid < 0) continue;
            /* depth cap (bam_mplp_set_maxcnt, :1107): live nodes + the empty tail node */
            if (cur_tid == r->tid && cur_pos == r->pos && (long)rn + 1 > MAX_PILEUP_SIZE) continue;
            if (r->tid < max_tid || (r->tid == max_tid && r->pos < max_pos)) { fprintf(stderr, "[oracle] input not sorted\n"); status = -2; eof = true; break; }
            max_tid = r->tid; max_pos = r->pos;
            if (r->end > cur_pos || r->tid > cur_tid) {
                if (rn == rcap) { rcap *= 2; live = xrealloc(live, rcap * sizeof(read_t *)); }
                live[rn++] = r;
            }
        }
        if (status < 0) break;                                     /* htslib error: loop ends with ret < 0 (:1129) */
        if (eof && rn == 0) break;
        /* build column, dropping finished reads */
        size_t w = 0; int ncol = 0;
        for (size_t i = 0; i < rn; i++) {
            read_t *r = live[i];
            if (r->tid < cur_tid || (r->tid == cur_tid && r->end <= cur_pos)) continue;   /* freed by htslib; we just drop */
            live[w++] = r;
            if (r->tid == cur_tid && r->pos <= cur_pos) {
                if ((size_t)ncol == colcap) { colcap *= 2; col = xrealloc(col, sizeof(col_entry) * colcap); handled = xrealloc(handled, sizeof(int) * colcap); }
                col[ncol].read = (int)(w - 1);
                column_of(r, cur_pos, &col[ncol].qpos, &col[ncol].is_del, &col[ncol].is_refskip);
                ncol++;
            }
        }
        rn = w;
        int tid = cur_tid; int64_t pos = cur_pos;
        if (rn && cur_tid < live[0]->tid) { cur_tid = live[0]->tid; cur_pos = live[0]->pos; }
        else if (rn && cur_pos < live[0]->pos) cur_pos = live[0]->pos;
        else cur_pos++;
        if (!ncol) { if (eof && rn == 0) break; continue; }

        /* ---------------- one covered locus (stochasticSpike.c:1129-1623) ---------------- */
        if (fa_of_tid[tid] < 0) { fprintf(stderr, "[oracle] contig %s not in reference\n", h.name[tid]); return 1; }
        const char *ref = fa.seq[fa_of_tid[tid]];
        char refBase = ref[pos];
        int filter = F_NONE, refCnt = 0, mutCnt = 0, err[4] = {0, 0, 0, 0};
        char mutantAllele = select_mutant_allele(refBase);                                  /* :1197 (always) */
        if (tgt.base == 'G' || tgt.base == 'C' || tgt.base == 'A' || tgt.base == 'T') mutantAllele = tgt.base;   /* :1199-1203 */
        nLoci++;
        if (ncol > maxDepth) maxDepth = ncol;
        bool mutateHere = false;
        if (pos == tgt.locus && !strcmp(tgt.contig, h.name[tid])) { mutateHere = true; filter = F_UNDETECTED; }   /* :1234-1245 */
        for (int j = 0; j < ncol; j++) handled[j] = 0;

        for (int j = 0; j < ncol; j++) {
            read_t *r = live[col[j].read];
            fold++;
            bool skip = col[j].is_del || col[j].is_refskip || r->qual[col[j].qpos] == 0 || handled[j];   /* :1270 */
            if (!skip) {
                /* getBaseWithRPOcheck (:387-432): mate = first later entry with the same QNAME */
                char readBase = r->seq[col[j].qpos], mateBase = 0; int readBQ = r->qual[col[j].qpos], mateBQ = 0, m = -1;
                for (int x = j + 1; x < ncol; x++) if (!strcmp(r->qname, live[col[x].read]->qname)) { m = x; break; }
                read_t *mr = NULL;
                if (m >= 0 && !handled[m]) { mr = live[col[m].read]; mateBase = mr->seq[col[m].qpos]; mateBQ = mr->qual[col[m].qpos]; }
                if (mateBase == 'N') mateBQ = 0;
                if (readBase == 'N') readBQ = 0;
                char base = readBase;
                if (mateBase && mateBase != readBase && mateBQ > readBQ) base = mateBase;
                if (base == 'N') { handled[j] = 1; if (mateBase) handled[m] = 1; }                       /* :584-592, :1337-1340 */
                else if (!mutateHere || !coin_toss(tgt.mutFreq)) {                                        /* :613 */
                    if (base == refBase) refCnt++;                                                        /* :619-620, :1343-1345 */
                    else { err_inc(err, base); handled[j] = 1; if (mateBase) handled[m] = 1; }
                } else {
                    char A = mutantAllele, F = refBase, R = readBase, M = mateBase;
#define SETBASE(rd, qp, b) ((rd)->seq[(qp)] = (b))
                    if (!M && R == F) {                                            /* case 1 (:639-658) */
                        SETBASE(r, col[j].qpos, A); mutCnt++;
                        if (filter == F_UNDETECTED || filter == F_NO_COVERAGE) filter = F_PASS;
                        handled[j] = 1;
                    } else if (!M && R != F) {                                     /* case 2 (:661-699) */
                        if (filter != F_MASKED_OVL) filter = F_MASKED;
                        if (R == A) { char d = select_mutant_allele(A); SETBASE(r, col[j].qpos, d); base = d; }
                        err_inc(err, base); handled[j] = 1;
                    } else if (M && R == F && M == F) {                            /* case 3 (:702-730) */
                        SETBASE(r, col[j].qpos, A); handled[j] = 1;
                        SETBASE(mr, col[m].qpos, A); handled[m] = 1;
                        mutCnt++;
                        if (filter == F_UNDETECTED || filter == F_NO_COVERAGE) filter = F_PASS;
                    } else if (M && R == F && M != F) {                            /* case 4 (:734-784) */
                        SETBASE(r, col[j].qpos, A); handled[j] = 1;
                        if (M == A) { char d = select_mutant_allele(A); SETBASE(mr, col[m].qpos, d); if (base == M) base = d; }
                        handled[m] = 1; filter = F_MASKED_OVL;
                        if (base == F) mutCnt++; else err_inc(err, base);
                    } else if (M && R != F && M == F) {                            /* case 5 (:787-838) */
                        SETBASE(mr, col[m].qpos, A); handled[m] = 1;
                        if (R == A) { char d = select_mutant_allele(A); SETBASE(r, col[j].qpos, d); if (base == R) base = d; }
                        handled[j] = 1; filter = F_MASKED_OVL;
                        if (base == F) mutCnt++; else err_inc(err, base);
                    } else {                                                       /* case 6 (:841-898) */
                        filter = F_MASKED_OVL;
                        if (R == A) { char d = select_mutant_allele(A); SETBASE(r, col[j].qpos, d); if (base == R) base = d; }
                        if (M == A) { char d = select_mutant_allele(A); SETBASE(mr, col[m].qpos, d); if (base == M) base = d; }
                        err_inc(err, base); handled[m] = 1; handled[j] = 1;
                    }
                }
            }
            if (pos == r->end - 1) { alignmentCount++; write_read(out, &h, r); }                           /* :1272-1285, :1362-1371 */
        }

        /* truth.vcf (:1406-1557) */
        int idx[4]; int totErr = err[0] + err[1] + err[2] + err[3];
        if (filter != F_NONE && filter != F_NO_COVERAGE) {
            err_order(err, idx);
            fprintf(vcf, "%s\t%d\t.\t%c\t%c", h.name[tid], (int)pos + 1, refBase, mutantAllele);
            if (totErr) { fprintf(vcf, ","); print_err_list(vcf, err, idx, false); }
            fprintf(vcf, "\t.\t%s\tDP=%d;AF=%.6g\tAD\t%d,%d", FILTER_NAME[filter], refCnt + mutCnt + totErr, tgt.mutFreq, refCnt, mutCnt);
            if (totErr) { fprintf(vcf, ","); print_err_list(vcf, err, idx, true); }
            fprintf(vcf, "\n");
        } else if (totErr) {
            err_order(err, idx);
            fprintf(vcf, "%s\t%d\t.\t%c\t", h.name[tid], (int)pos + 1, refBase);
            print_err_list(vcf, err, idx, false);
            int dp = refCnt + totErr;
            fprintf(vcf, "\t.\tSEQ_ERROR\tDP=%d;AF=%.6g\tAD\t%d,", dp, (float)totErr / dp, refCnt);
            print_err_list(vcf, err, idx, true);
            fprintf(vcf, "\n");
        }

        /* target advance, including the documented if-not-while skip (:1578-1619) */
        if (tgt.contig[0] == 0) continue;
        if (pos == tgt.locus && !strcmp(tgt.contig, h.name[tid])) next_target(&tgt, cfg, &h);
        else if ((tid == tgt.c_tid && pos > tgt.locus) || tid > tgt.c_tid) {
            if (pos != tgt.locus) print_no_coverage(vcf, &tgt);
            next_target(&tgt, cfg, &h);
        }
    }
    while (tgt.contig[0] != 0) { print_no_coverage(vcf, &tgt); next_target(&tgt, cfg, &h); }                /* :1630-1646 */
    fclose(vcf); fclose(out); fclose(cfg);
    if (in != stdin) fclose(in);
    free(line);
    if (st_out) { st_out->alignmentCount = alignmentCount; st_out->numberOfLociCovered = nLoci; st_out->totalFoldCoverage = fold; st_out->maxDepth = maxDepth; }
    if (statsout)                                                                                           /* :1668 (SIGFPE when nothing is covered) */
        fprintf(statsout, "\nDONE...\nalignmentCount (#reads) = %ld,\nnumberOfLociCovered = %ld\ntotalFoldCoverage = %ld\nmaxDepth = %ld, Avg. coverage = %ld\n",
                alignmentCount, nLoci, fold, maxDepth, (long)(fold / nLoci));
    return 0;
}

/* ---- glibc rand() model, exported so tests can pin the device RNG model on CPU ---- */
void glibc_rand_fill(unsigned seed, uint64_t skip, int64_t n, int32_t *out)
{
    srand(seed);
    for (uint64_t i = 0; i < skip; i++) (void)rand();
    for (int64_t i = 0; i < n; i++) out[i] = rand();
}

/* RNG walk alone: one select_mutant_allele(ref[i]) per base; returns draws consumed (App. C RNG-walk KAT) */
int64_t glibc_walk(unsigned seed, const char *ref, int64_t n, int32_t *next_rand)
{
    srand(seed);
    int64_t draws = 0;
    int cutoff = (RAND_MAX / 4) * 4;
    for (int64_t i = 0; i < n; i++) {
        for (;;) {
            int r = rand(); draws++;
            if (r >= cutoff) continue;
            if ("GCAT"[r % 4] != ref[i]) break;
        }
    }
    if (next_rand) *next_rand = rand();
    return draws;
}

#ifdef SPIKE_ORACLE_MAIN
int main(int argc, char **argv)
{
    char *cmd = strrchr(argv[0], '/'); cmd = cmd ? cmd + 1 : argv[0];
    const char *name = getenv("SPIKE_ORACLE_CMDNAME");          /* lets tests match the product's basename(argv[0]) */
    if (argc != 6) {
        fprintf(stderr, "\n Usage: %s <donor BAM> <donor reference> <somatic mutation config file> <seed> <output SAM filename>\n\n", cmd);
        exit(0);                                                  /* :938-941 */
    }
    return spike_oracle_run(name ? name : cmd, argv[1], argv[2], argv[3], argv[4], argv[5], "truth.vcf", stdout, NULL);
}
#endif
