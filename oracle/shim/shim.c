/*
 * oracle/shim/shim.c -- TEST INFRASTRUCTURE ONLY (not product, never linked into
 * libssb200.so).
 *
 * Behaviour behind oracle/shim/htslib/{sam,faidx}.h: the I/O + pileup layer of
 * htslib-1.13 that the reference's stochasticSpike.c relies on, restated from
 * the htslib API contract ("[htslib-recall]") so that the
 * UNMODIFIED /root/reference/stochasticSpike.c links and runs here.  htslib is a
 * third-party dependency (samtools-1.13 bundles htslib-1.13, fetched by the
 * reference's install.bash:82-101) that is absent from this image; nothing in
 * the reference's own tree pins its behaviour, so everything in this file is
 * "parity unpinned" against real htslib and says so.
 *
 * Differences from real htslib, on purpose:
 *   - input is SAM TEXT (no BGZF/BAM); sam_index_load() returns a dummy;
 *   - optional (aux) fields are kept as their SAM text and re-emitted verbatim
 *     (htslib would parse and re-print them; identical for canonical input).
 *
 * What is restated (each cites the reference call site that needs it):
 *   sam_hdr_read/sam_hdr_write ........ stochasticSpike.c:971,986 (header verbatim)
 *   sam_hdr_count_lines/find_tag_pos .. stochasticSpike.c:998-1010 (first @RG's SM)
 *   sam_hdr_name2tid/tid2name/nref .... stochasticSpike.c:122,216,1131,1236
 *   sam_read1 ......................... stochasticSpike.c:248 (SAM line -> bam1_t, 4-bit SEQ)
 *   sam_write1 ........................ stochasticSpike.c:273 (bam1_t -> SAM line)
 *   bam_cigar2rlen/qlen ............... stochasticSpike.c:255,1272,1362
 *   bam_mplp_init/set_maxcnt/auto ..... stochasticSpike.c:1097,1107,1129 (pileup engine)
 *   fai_load/faidx_fetch_seq64 ........ stochasticSpike.c:1042,215
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>
#include <errno.h>

#include "htslib/sam.h"
#include "htslib/faidx.h"

const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

/* IUPAC -> 4-bit code; anything unknown is N (15); '=' is 0. */
const unsigned char seq_nt16_table[256] = {
#define N15 15
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    1, 2, 4, 8, N15,N15,N15,N15,N15,N15,N15,N15,N15, 0 ,N15,N15,       /* '0'..'3' are htslib's digit aliases, '=' -> 0 */
    N15, 1, 14, 2, 13,N15,N15, 4, 11,N15,N15, 12,N15, 3, 15,N15,       /* @ A B C D E F G H I J K L M N O */
    N15,N15, 5, 6, 8,N15, 7, 9,N15, 10,N15,N15,N15,N15,N15,N15,        /* P Q R S T U V W X Y Z */
    N15, 1, 14, 2, 13,N15,N15, 4, 11,N15,N15, 12,N15, 3, 15,N15,       /* ` a b c d e f g h i j k l m n o */
    N15,N15, 5, 6, 8,N15, 7, 9,N15, 10,N15,N15,N15,N15,N15,N15,        /* p q r s t u v w x y z */
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,
    N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15,N15
#undef N15
};

/* ------------------------------------------------------------------ files */

struct shim_file {
    FILE  *fp;
    int    writing;
    char  *pending;      /* first alignment line, read while scanning the header */
    size_t pending_cap;
    ssize_t pending_len; /* -1: none */
    char  *line;
    size_t line_cap;
    char  *iobuf;
};

struct shim_idx { int dummy; };
struct shim_itr { int dummy; };

static samFile *shim_open(const char *fn, int writing)
{
    samFile *f = calloc(1, sizeof *f);
    if (!f) return NULL;
    if (strcmp(fn, "-") == 0) f->fp = writing ? stdout : stdin;
    else f->fp = fopen(fn, writing ? "w" : "r");
    if (!f->fp) { free(f); return NULL; }
    f->writing = writing;
    f->pending_len = -1;
    f->iobuf = malloc(1 << 22);
    if (f->iobuf) setvbuf(f->fp, f->iobuf, _IOFBF, 1 << 22);
    return f;
}

samFile *sam_open(const char *fn, const char *mode) { return shim_open(fn, mode && mode[0] == 'w'); }
samFile *sam_open_format(const char *fn, const char *mode, const htsFormat *fmt) { (void)fmt; return sam_open(fn, mode); }

int sam_close(samFile *f)
{
    int r = 0;
    if (!f) return 0;
    if (f->fp && f->fp != stdin && f->fp != stdout) r = fclose(f->fp);
    else if (f->fp == stdout) fflush(stdout);
    free(f->pending); free(f->line); free(f->iobuf); free(f);
    return r;
}

/* ----------------------------------------------------------------- header */

struct shim_hdr {
    char   *text;
    size_t  l_text;
    int     n_targets;
    char  **target_name;
    int64_t *target_len;
    /* open-addressing name -> tid */
    int     hcap;
    int    *hslot;
};

static uint64_t str_hash(const char *s)
{
    uint64_t h = 1469598103934665603ULL;
    for (; *s; s++) { h ^= (unsigned char)*s; h *= 1099511628211ULL; }
    return h;
}

static void hdr_index_names(sam_hdr_t *h)
{
    h->hcap = 16;
    while (h->hcap < 4 * (h->n_targets + 1)) h->hcap <<= 1;
    h->hslot = malloc(sizeof(int) * h->hcap);
    for (int i = 0; i < h->hcap; i++) h->hslot[i] = -1;
    for (int t = 0; t < h->n_targets; t++) {
        uint64_t k = str_hash(h->target_name[t]) & (h->hcap - 1);
        int dup = 0;
        while (h->hslot[k] >= 0) {
            if (strcmp(h->target_name[h->hslot[k]], h->target_name[t]) == 0) { dup = 1; break; }
            k = (k + 1) & (h->hcap - 1);
        }
        if (!dup) h->hslot[k] = t;   /* first @SQ with a name wins */
    }
}

sam_hdr_t *sam_hdr_read(samFile *f)
{
    sam_hdr_t *h = calloc(1, sizeof *h);
    size_t cap = 1 << 16, tcap = 64;
    h->text = malloc(cap);
    h->target_name = malloc(sizeof(char *) * tcap);
    h->target_len = malloc(sizeof(int64_t) * tcap);
    for (;;) {
        ssize_t n = getline(&f->pending, &f->pending_cap, f->fp);
        if (n < 0) { f->pending_len = -1; break; }
        if (f->pending[0] != '@') { f->pending_len = n; break; }
        if (h->l_text + (size_t)n + 2 > cap) { while (h->l_text + (size_t)n + 2 > cap) cap <<= 1; h->text = realloc(h->text, cap); }
        memcpy(h->text + h->l_text, f->pending, (size_t)n);
        h->l_text += (size_t)n;
        if (f->pending[n - 1] != '\n') h->text[h->l_text++] = '\n';
        if (strncmp(f->pending, "@SQ\t", 4) == 0) {
            char *sn = NULL; int64_t ln = 0;
            char *copy = strndup(f->pending, (size_t)n), *save = NULL;
            for (char *tok = strtok_r(copy, "\t\r\n", &save); tok; tok = strtok_r(NULL, "\t\r\n", &save)) {
                if (strncmp(tok, "SN:", 3) == 0) sn = tok + 3;
                else if (strncmp(tok, "LN:", 3) == 0) ln = strtoll(tok + 3, NULL, 10);
            }
            if (sn) {
                if ((size_t)h->n_targets == tcap) {
                    tcap <<= 1;
                    h->target_name = realloc(h->target_name, sizeof(char *) * tcap);
                    h->target_len = realloc(h->target_len, sizeof(int64_t) * tcap);
                }
                h->target_name[h->n_targets] = strdup(sn);
                h->target_len[h->n_targets] = ln;
                h->n_targets++;
            }
            free(copy);
        }
    }
    h->text[h->l_text] = 0;
    hdr_index_names(h);
    return h;
}

int sam_hdr_write(samFile *f, const sam_hdr_t *h)
{
    /* hrecs are not parsed yet at stochasticSpike.c:986, so the text goes out as read */
    if (h->l_text && fwrite(h->text, 1, h->l_text, f->fp) != h->l_text) return -1;
    return 0;
}

void sam_hdr_destroy(sam_hdr_t *h)
{
    if (!h) return;
    for (int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text); free(h->hslot); free(h);
}

int sam_hdr_nref(const sam_hdr_t *h) { return h->n_targets; }
const char *sam_hdr_tid2name(const sam_hdr_t *h, int tid) { return (tid >= 0 && tid < h->n_targets) ? h->target_name[tid] : NULL; }

int sam_hdr_name2tid(sam_hdr_t *h, const char *ref)
{
    uint64_t k = str_hash(ref) & (h->hcap - 1);
    while (h->hslot[k] >= 0) {
        if (strcmp(h->target_name[h->hslot[k]], ref) == 0) return h->hslot[k];
        k = (k + 1) & (h->hcap - 1);
    }
    return -1;
}

/* n-th header line of a 2-letter type, or NULL */
static const char *hdr_line(const sam_hdr_t *h, const char *type, int pos, size_t *len)
{
    const char *p = h->text, *end = h->text + h->l_text;
    int seen = 0;
    while (p < end) {
        const char *nl = memchr(p, '\n', (size_t)(end - p));
        size_t l = nl ? (size_t)(nl - p) : (size_t)(end - p);
        if (l >= 3 && p[0] == '@' && p[1] == type[0] && p[2] == type[1] && (l == 3 || p[3] == '\t')) {
            if (seen == pos) { *len = l; return p; }
            seen++;
        }
        p += l + 1;
    }
    return NULL;
}

int sam_hdr_count_lines(sam_hdr_t *h, const char *type)
{
    int n = 0; size_t l;
    while (hdr_line(h, type, n, &l)) n++;
    return n;
}

int sam_hdr_find_tag_pos(sam_hdr_t *h, const char *type, int pos, const char *key, kstring_t *ks)
{
    size_t l;
    const char *line = hdr_line(h, type, pos, &l);
    if (!line) return -1;
    const char *p = line + 3, *end = line + l;
    while (p < end) {
        if (*p == '\t') p++;
        const char *q = memchr(p, '\t', (size_t)(end - p));
        size_t fl = q ? (size_t)(q - p) : (size_t)(end - p);
        if (fl >= 3 && p[0] == key[0] && p[1] == key[1] && p[2] == ':') {
            size_t vl = fl - 3;
            while (vl && (p[3 + vl - 1] == '\r')) vl--;
            free(ks->s);
            ks->s = malloc(vl + 1);
            memcpy(ks->s, p + 3, vl);
            ks->s[vl] = 0;
            ks->l = vl; ks->m = vl + 1;
            return 0;
        }
        p += fl;
    }
    return -1;
}

hts_idx_t *sam_index_load(samFile *fp, const char *fn) { (void)fp; (void)fn; static struct shim_idx d; return &d; }
hts_itr_t *sam_itr_querys(const hts_idx_t *idx, sam_hdr_t *hdr, const char *region) { (void)idx; (void)hdr; (void)region; return NULL; }
int  sam_itr_next(samFile *fp, hts_itr_t *itr, bam1_t *b) { (void)fp; (void)itr; (void)b; return -1; }
void hts_itr_destroy(hts_itr_t *itr) { (void)itr; }

/* ------------------------------------------------------------- alignments */

hts_pos_t bam_cigar2rlen(int n_cigar, const uint32_t *cigar)
{
    hts_pos_t l = 0;
    for (int k = 0; k < n_cigar; k++) {
        int op = cigar[k] & BAM_CIGAR_MASK;
        if (op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF)
            l += cigar[k] >> BAM_CIGAR_SHIFT;
    }
    return l;
}

int64_t bam_cigar2qlen(int n_cigar, const uint32_t *cigar)
{
    int64_t l = 0;
    for (int k = 0; k < n_cigar; k++) {
        int op = cigar[k] & BAM_CIGAR_MASK;
        if (op == BAM_CMATCH || op == BAM_CINS || op == BAM_CSOFT_CLIP || op == BAM_CEQUAL || op == BAM_CDIFF)
            l += cigar[k] >> BAM_CIGAR_SHIFT;
    }
    return l;
}

static int ensure_data(bam1_t *b, size_t need)
{
    if (need > b->m_data) {
        size_t m = b->m_data ? b->m_data : 256;
        while (m < need) m <<= 1;
        uint8_t *d = realloc(b->data, m);
        if (!d) return -1;
        b->data = d; b->m_data = (uint32_t)m;
    }
    return 0;
}

static int cigar_op_code(char c)
{
    switch (c) {
    case 'M': return BAM_CMATCH; case 'I': return BAM_CINS; case 'D': return BAM_CDEL;
    case 'N': return BAM_CREF_SKIP; case 'S': return BAM_CSOFT_CLIP; case 'H': return BAM_CHARD_CLIP;
    case 'P': return BAM_CPAD; case '=': return BAM_CEQUAL; case 'X': return BAM_CDIFF;
    default: return -1;
    }
}

/* Splits `s` (length n, no newline) on tabs in place; returns number of fields. */
static int split_tabs(char *s, size_t n, char **fld, size_t *flen, int maxf)
{
    int nf = 0;
    char *p = s, *end = s + n;
    while (nf < maxf) {
        char *t = memchr(p, '\t', (size_t)(end - p));
        fld[nf] = p;
        if (!t || nf == maxf - 1) { flen[nf] = (size_t)(end - p); nf++; break; }
        flen[nf] = (size_t)(t - p);
        *t = 0;
        nf++;
        p = t + 1;
    }
    return nf;
}

/* SAM text line -> bam1_t.  <-1 on malformed input (pileup then aborts, as htslib's would). */
static int parse_sam_line(sam_hdr_t *h, char *s, size_t n, bam1_t *b)
{
    char *fld[12]; size_t flen[12];
    int nf = split_tabs(s, n, fld, flen, 12);   /* 12th = all aux text, tabs kept */
    if (nf < 11) return -2;
    if (nf == 12) { /* undo nothing: split_tabs leaves the tail untouched */ }
    for (int i = 0; i < 11; i++) fld[i][flen[i]] = 0;

    bam1_core_t *c = &b->core;
    memset(c, 0, sizeof *c);
    size_t l_qname = flen[0] + 1;
    size_t pad = (4 - (l_qname & 3)) & 3;            /* keep the cigar 4-byte aligned */
    c->l_extranul = (uint8_t)pad;
    c->l_qname = (uint16_t)(l_qname + pad);
    c->flag = (uint16_t)strtol(fld[1], NULL, 0);
    if (strcmp(fld[2], "*") == 0) c->tid = -1;
    else c->tid = sam_hdr_name2tid(h, fld[2]);       /* unknown name -> -1 (treated as unmapped) */
    c->pos = strtoll(fld[3], NULL, 10) - 1;
    c->qual = (uint8_t)strtol(fld[4], NULL, 10);

    /* CIGAR */
    uint32_t n_cigar = 0;
    if (strcmp(fld[5], "*") != 0) for (char *p = fld[5]; *p; p++) if (!isdigit((unsigned char)*p)) n_cigar++;
    c->n_cigar = n_cigar;
    int32_t l_qseq = (strcmp(fld[9], "*") == 0) ? 0 : (int32_t)flen[9];
    c->l_qseq = l_qseq;
    size_t l_aux = (nf == 12) ? flen[11] : 0;
    size_t need = c->l_qname + 4u * n_cigar + (size_t)((l_qseq + 1) >> 1) + (size_t)l_qseq + l_aux + 1;
    if (ensure_data(b, need) < 0) return -2;
    memcpy(b->data, fld[0], flen[0]);
    memset(b->data + flen[0], 0, 1 + pad);
    uint32_t *cig = bam_get_cigar(b);
    if (n_cigar) {
        char *p = fld[5];
        for (uint32_t k = 0; k < n_cigar; k++) {
            char *q;
            unsigned long len = strtoul(p, &q, 10);
            int op = cigar_op_code(*q);
            if (q == p || op < 0) return -2;
            cig[k] = (uint32_t)(len << BAM_CIGAR_SHIFT) | (uint32_t)op;
            p = q + 1;
        }
    }
    if (strcmp(fld[6], "=") == 0) c->mtid = c->tid;
    else if (strcmp(fld[6], "*") == 0) c->mtid = -1;
    else c->mtid = sam_hdr_name2tid(h, fld[6]);
    c->mpos = strtoll(fld[7], NULL, 10) - 1;
    c->isize = strtoll(fld[8], NULL, 10);

    uint8_t *seq = bam_get_seq(b);
    memset(seq, 0, (size_t)((l_qseq + 1) >> 1));
    for (int32_t i = 0; i < l_qseq; i++)
        seq[i >> 1] |= (uint8_t)(seq_nt16_table[(unsigned char)fld[9][i]] << ((~i & 1) << 2));
    if (n_cigar && l_qseq && bam_cigar2qlen((int)n_cigar, cig) != l_qseq) return -2;   /* "CIGAR and query sequence are of different length" */
    uint8_t *qual = bam_get_qual(b);
    if (strcmp(fld[10], "*") == 0) memset(qual, 0xff, (size_t)l_qseq);
    else {
        if ((int32_t)flen[10] != l_qseq) return -2;                                      /* "SEQ and QUAL are of different length" */
        for (int32_t i = 0; i < l_qseq; i++) qual[i] = (uint8_t)(fld[10][i] - 33);
    }
    if (l_aux) memcpy(bam_get_aux(b), fld[11], l_aux);
    b->l_aux_text = (int)l_aux;
    b->l_data = (int)(need - 1);
    return 0;
}

int sam_read1(samFile *f, sam_hdr_t *h, bam1_t *b)
{
    char *s; ssize_t n;
    for (;;) {
        if (f->pending_len >= 0) { s = f->pending; n = f->pending_len; f->pending_len = -1; }
        else {
            n = getline(&f->line, &f->line_cap, f->fp);
            if (n < 0) return -1;
            s = f->line;
        }
        while (n > 0 && (s[n - 1] == '\n' || s[n - 1] == '\r')) n--;
        if (n == 0) continue;
        break;
    }
    int r = parse_sam_line(h, s, (size_t)n, b);
    if (r < 0) { fprintf(stderr, "[shim] malformed SAM record\n"); return -2; }
    return (int)n;
}

int sam_write1(samFile *f, const sam_hdr_t *h, const bam1_t *b)
{
    const bam1_core_t *c = &b->core;
    FILE *o = f->fp;
    fputs(bam_get_qname(b), o);
    fprintf(o, "\t%d\t", c->flag);
    if (c->tid >= 0) fputs(h->target_name[c->tid], o); else fputc('*', o);
    fprintf(o, "\t%lld\t%d\t", (long long)c->pos + 1, c->qual);
    if (c->n_cigar) {
        const uint32_t *cig = bam_get_cigar(b);
        for (uint32_t k = 0; k < c->n_cigar; k++) fprintf(o, "%u%c", cig[k] >> BAM_CIGAR_SHIFT, "MIDNSHP=XB"[cig[k] & BAM_CIGAR_MASK]);
    } else fputc('*', o);
    fputc('\t', o);
    if (c->mtid < 0) fputc('*', o);
    else if (c->mtid == c->tid) fputc('=', o);
    else fputs(h->target_name[c->mtid], o);
    fprintf(o, "\t%lld\t%lld\t", (long long)c->mpos + 1, (long long)c->isize);
    if (c->l_qseq) {
        const uint8_t *seq = bam_get_seq(b), *q = bam_get_qual(b);
        for (int32_t i = 0; i < c->l_qseq; i++) fputc(seq_nt16_str[bam_seqi(seq, i)], o);
        fputc('\t', o);
        if (q[0] == 0xff) fputc('*', o);
        else for (int32_t i = 0; i < c->l_qseq; i++) fputc(q[i] + 33, o);
    } else fputs("*\t*", o);
    if (b->l_aux_text) { fputc('\t', o); fwrite(bam_get_aux(b), 1, (size_t)b->l_aux_text, o); }
    fputc('\n', o);
    return ferror(o) ? -1 : 1;
}

/* ----------------------------------------------------------------- pileup */
/*
 * One input stream (NUM_BAMS == 1, stochasticSpike.c:37).  Reads are appended to
 * a list in arrival order; a read is reported at every position in
 * [beg, end) with end = pos + reference length; it leaves the list once
 * end <= current position; columns with no reads are never returned.
 */
typedef struct plp_node {
    bam1_t b;
    hts_pos_t beg, end;
    struct plp_node *next;
} plp_node;

struct shim_mplp {
    bam_plp_auto_f func;
    void *data;
    plp_node *head, *tail;      /* tail is always an empty node waiting to be filled */
    plp_node *free_list;
    int cnt;                    /* nodes in use, including the empty tail */
    int maxcnt;
    int32_t tid, max_tid;
    hts_pos_t pos, max_pos;
    int is_eof, error;
    bam1_t scratch;
    bam_pileup1_t *plp;
    int max_plp;
};

static plp_node *node_alloc(bam_mplp_t it)
{
    plp_node *n = it->free_list;
    if (n) { it->free_list = n->next; n->next = NULL; }
    else n = calloc(1, sizeof *n);
    it->cnt++;
    return n;
}

static void node_free(bam_mplp_t it, plp_node *n)
{
    n->next = it->free_list;
    it->free_list = n;
    it->cnt--;
}

bam_mplp_t bam_mplp_init(int n, bam_plp_auto_f func, void **data)
{
    if (n != 1) { fprintf(stderr, "[shim] only one input supported\n"); exit(1); }
    bam_mplp_t it = calloc(1, sizeof *it);
    it->func = func; it->data = data[0];
    it->head = it->tail = node_alloc(it);
    it->maxcnt = 8000;
    it->max_tid = -1; it->max_pos = -1;
    return it;
}

void bam_mplp_set_maxcnt(bam_mplp_t it, int maxcnt) { it->maxcnt = maxcnt; }

void bam_mplp_destroy(bam_mplp_t it)
{
    if (!it) return;
    for (plp_node *n = it->head; n;) { plp_node *x = n->next; free(n->b.data); free(n); n = x; }
    for (plp_node *n = it->free_list; n;) { plp_node *x = n->next; free(n->b.data); free(n); n = x; }
    free(it->scratch.data); free(it->plp); free(it);
}

static int node_copy(plp_node *dst, const bam1_t *src)
{
    uint8_t *d = dst->b.data; uint32_t m = dst->b.m_data;
    dst->b = *src;
    dst->b.data = d; dst->b.m_data = m;
    if (ensure_data(&dst->b, (size_t)src->l_data + 1) < 0) return -1;
    memcpy(dst->b.data, src->data, (size_t)src->l_data);
    return 0;
}

static int plp_push(bam_mplp_t it, const bam1_t *b)
{
    if (it->error) return -1;
    if (!b) { it->is_eof = 1; return 0; }
    if (b->core.tid < 0) return 0;
    if (b->core.flag & BAM_FUNMAP) return 0;
    /* depth cap: a read starting exactly at the column being built is dropped
       once more than maxcnt nodes are live (stochasticSpike.c:1107 sets 10000) */
    if (it->tid == b->core.tid && it->pos == b->core.pos && it->cnt > it->maxcnt) return 0;
    if (node_copy(it->tail, b) < 0) return -1;
    it->tail->beg = b->core.pos;
    it->tail->end = b->core.pos + bam_cigar2rlen((int)b->core.n_cigar, bam_get_cigar(b));
    if (b->core.tid < it->max_tid) { fprintf(stderr, "[shim] input not sorted (chromosomes out of order)\n"); it->error = 1; return -1; }
    if (b->core.tid == it->max_tid && it->tail->beg < it->max_pos) { fprintf(stderr, "[shim] input not sorted (reads out of order)\n"); it->error = 1; return -1; }
    it->max_tid = b->core.tid; it->max_pos = it->tail->beg;
    if (it->tail->end > it->pos || it->tail->b.core.tid > it->tid) {
        if (!it->tail->next) it->tail->next = node_alloc(it);
        it->tail = it->tail->next;
    }
    return 0;
}

/* position -> (qpos, is_del, is_refskip) for one read; walks the CIGAR from the start */
static void resolve_column(bam_pileup1_t *p, hts_pos_t pos)
{
    const bam1_t *b = p->b;
    const uint32_t *cig = bam_get_cigar(b);
    hts_pos_t x = b->core.pos;
    int32_t y = 0;
    p->is_del = p->is_refskip = p->is_head = p->is_tail = 0;
    p->indel = 0; p->level = 0; p->qpos = 0; p->cigar_ind = 0;
    for (uint32_t k = 0; k < b->core.n_cigar; k++) {
        int op = cig[k] & BAM_CIGAR_MASK;
        int32_t l = (int32_t)(cig[k] >> BAM_CIGAR_SHIFT);
        if (op == BAM_CMATCH || op == BAM_CEQUAL || op == BAM_CDIFF) {
            if (pos < x + l) { p->qpos = y + (int32_t)(pos - x); p->cigar_ind = (int)k; return; }
            x += l; y += l;
        } else if (op == BAM_CDEL || op == BAM_CREF_SKIP) {
            if (pos < x + l) { p->is_del = 1; p->is_refskip = (op == BAM_CREF_SKIP); p->qpos = y; p->cigar_ind = (int)k; return; }
            x += l;
        } else if (op == BAM_CINS || op == BAM_CSOFT_CLIP) {
            y += l;
        }
    }
}

static const bam_pileup1_t *plp_next(bam_mplp_t it, int *tid_out, hts_pos_t *pos_out, int *n_out)
{
    if (it->error) { *n_out = -1; return NULL; }
    *n_out = 0;
    if (it->is_eof && it->head == it->tail) return NULL;
    while (it->is_eof || it->max_tid > it->tid || (it->max_tid == it->tid && it->max_pos > it->pos)) {
        int n = 0;
        plp_node **pp = &it->head;
        while (*pp != it->tail) {
            plp_node *p = *pp;
            if (p->b.core.tid < it->tid || (p->b.core.tid == it->tid && p->end <= it->pos)) {
                *pp = p->next;
                node_free(it, p);
            } else {
                if (p->b.core.tid == it->tid && p->beg <= it->pos) {
                    if (n == it->max_plp) {
                        it->max_plp = it->max_plp ? it->max_plp << 1 : 256;
                        it->plp = realloc(it->plp, sizeof(bam_pileup1_t) * (size_t)it->max_plp);
                    }
                    it->plp[n].b = &p->b;
                    resolve_column(&it->plp[n], it->pos);
                    n++;
                }
                pp = &(*pp)->next;
            }
        }
        *n_out = n; *tid_out = it->tid; *pos_out = it->pos;
        if (it->head != it->tail && it->tid > it->head->b.core.tid) {
            fprintf(stderr, "[shim] unsorted input, pileup aborts\n");
            it->error = 1; *n_out = -1; return NULL;
        }
        if (it->head != it->tail && it->tid < it->head->b.core.tid) { it->tid = it->head->b.core.tid; it->pos = it->head->beg; }
        else if (it->head != it->tail && it->pos < it->head->beg) it->pos = it->head->beg;
        else ++it->pos;
        if (n) return it->plp;
        if (it->is_eof && it->head == it->tail) break;
    }
    return NULL;
}

int bam_mplp_auto(bam_mplp_t it, int *_tid, int *_pos, int *n_plp, const bam_pileup1_t **plp)
{
    int tid = 0; hts_pos_t pos = 0; int n = 0;
    const bam_pileup1_t *col = NULL;
    if (it->error) return -1;
    col = plp_next(it, &tid, &pos, &n);
    if (!col && !it->error && !it->is_eof) {
        int ret;
        while ((ret = it->func(it->data, &it->scratch)) >= 0) {
            if (plp_push(it, &it->scratch) < 0) return -1;
            if ((col = plp_next(it, &tid, &pos, &n)) != NULL) break;
        }
        if (!col) {
            if (ret < -1) { it->error = 1; return -1; }
            if (plp_push(it, NULL) < 0) return -1;
            col = plp_next(it, &tid, &pos, &n);
        }
    }
    if (it->error) return -1;
    if (!col) { n_plp[0] = 0; plp[0] = NULL; return 0; }
    *_tid = tid; *_pos = (int)pos; n_plp[0] = n; plp[0] = col;
    return 1;
}

/* ------------------------------------------------------------------ faidx */

struct shim_fai {
    int n;
    char **name;
    char **seq;
    int64_t *len;
};

faidx_t *fai_load(const char *fn)
{
    FILE *fp = fopen(fn, "r");
    if (!fp) return NULL;
    faidx_t *fa = calloc(1, sizeof *fa);
    size_t cap = 16, scap = 0;
    fa->name = malloc(sizeof(char *) * cap);
    fa->seq = malloc(sizeof(char *) * cap);
    fa->len = malloc(sizeof(int64_t) * cap);
    char *line = NULL; size_t lcap = 0; ssize_t n;
    int cur = -1;
    while ((n = getline(&line, &lcap, fp)) >= 0) {
        if (line[0] == '>') {
            if ((size_t)fa->n == cap) {
                cap <<= 1;
                fa->name = realloc(fa->name, sizeof(char *) * cap);
                fa->seq = realloc(fa->seq, sizeof(char *) * cap);
                fa->len = realloc(fa->len, sizeof(int64_t) * cap);
            }
            size_t e = 1;
            while (e < (size_t)n && !isspace((unsigned char)line[e])) e++;
            cur = fa->n++;
            fa->name[cur] = strndup(line + 1, e - 1);
            fa->seq[cur] = NULL; fa->len[cur] = 0; scap = 0;
            continue;
        }
        if (cur < 0) continue;
        if ((size_t)fa->len[cur] + (size_t)n + 1 > scap) {
            scap = scap ? scap : (1 << 20);
            while ((size_t)fa->len[cur] + (size_t)n + 1 > scap) scap <<= 1;
            fa->seq[cur] = realloc(fa->seq[cur], scap);
        }
        for (ssize_t i = 0; i < n; i++)
            if (isgraph((unsigned char)line[i])) fa->seq[cur][fa->len[cur]++] = line[i];   /* case preserved */
    }
    free(line);
    fclose(fp);
    return fa;
}

char *faidx_fetch_seq64(const faidx_t *fa, const char *c_name, hts_pos_t beg, hts_pos_t end, hts_pos_t *len)
{
    for (int i = 0; i < fa->n; i++) {
        if (strcmp(fa->name[i], c_name) == 0) {
            if (beg < 0) beg = 0;
            if (end >= fa->len[i]) end = fa->len[i] - 1;
            int64_t l = end >= beg ? end - beg + 1 : 0;
            char *s = malloc((size_t)l + 1);
            if (l) memcpy(s, fa->seq[i] + beg, (size_t)l);
            s[l] = 0;
            *len = l;
            return s;
        }
    }
    fprintf(stderr, "[shim] contig %s not in reference\n", c_name);
    *len = -2;
    return NULL;
}

void fai_destroy(faidx_t *fa)
{
    if (!fa) return;
    for (int i = 0; i < fa->n; i++) { free(fa->name[i]); free(fa->seq[i]); }
    free(fa->name); free(fa->seq); free(fa->len); free(fa);
}
