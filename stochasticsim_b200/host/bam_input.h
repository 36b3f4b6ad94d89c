/*
 * bam_input.h -- BGZF/BAM -> SAM text on the host, for the stochasticSpike main.
 *
 * The reference reads its first argument through htslib (sam_open / sam_hdr_read / sam_read1, stochasticSpike.c:960,971,248),
 * i.e. a BGZF-compressed BAM in the pipeline (bin/spikeIn.bash:41).  This is the decode stage in front of the GPU
 * tokeniser: BGZF blocks are independent deflate streams, so they are inflated in parallel (zlib, one range of
 * blocks per host thread), and the BAM records are then printed as the SAM lines htslib's sam_format1 would print
 * (same field order, '=' for a mate on the same contig, '*' for absent fields, integer aux types as `i`, floats
 * with %g).  The text that comes out is canonical by construction, so it satisfies the pass-through envelope of
 * the tokeniser.  No computation of the spike path happens here.
 */
#ifndef SSB_BAM_INPUT_H
#define SSB_BAM_INPUT_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#include <zlib.h>

static int bam_is_bgzf(const uint8_t *d, size_t n) { return n >= 18 && d[0] == 0x1f && d[1] == 0x8b && d[2] == 8 && (d[3] & 4); }

typedef struct { const uint8_t *src; size_t n_blocks; const size_t *off, *csize, *uoff; uint8_t *dst; size_t lo, hi; int rc; } bgzf_job;

static void *bgzf_worker(void *arg)
{
    bgzf_job *j = (bgzf_job *)arg;
    for (size_t b = j->lo; b < j->hi; b++) {
        const uint8_t *blk = j->src + j->off[b];
        const size_t xlen = blk[10] | (blk[11] << 8);
        const uint8_t *payload = blk + 12 + xlen;
        const size_t clen = j->csize[b] - 12 - xlen - 8;
        const size_t ulen = j->uoff[b + 1] - j->uoff[b];
        if (!ulen) continue;
        z_stream zs; memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) { j->rc = -1; return NULL; }
        zs.next_in = (Bytef *)payload; zs.avail_in = (uInt)clen;
        zs.next_out = j->dst + j->uoff[b]; zs.avail_out = (uInt)ulen;
        const int r = inflate(&zs, Z_FINISH);
        inflateEnd(&zs);
        if (r != Z_STREAM_END || zs.avail_out != 0) { j->rc = -1; return NULL; }
    }
    return NULL;
}

/* whole BGZF file -> uncompressed bytes (malloc'd); NULL on a malformed file */
static uint8_t *bgzf_inflate_all(const uint8_t *d, size_t n, size_t *n_out)
{
    size_t cap = 1024, nb = 0;
    size_t *off = malloc(cap * sizeof(size_t)), *cs = malloc(cap * sizeof(size_t)), *uo = malloc((cap + 1) * sizeof(size_t));
    size_t p = 0, total = 0;
    while (p + 18 <= n) {
        if (!(d[p] == 0x1f && d[p + 1] == 0x8b)) break;
        const size_t xlen = d[p + 10] | (d[p + 11] << 8);
        size_t bsize = 0, q = p + 12;
        while (q + 4 <= p + 12 + xlen) {                       /* extra subfields: 'B','C',2,BSIZE-1 */
            const size_t slen = d[q + 2] | (d[q + 3] << 8);
            if (d[q] == 'B' && d[q + 1] == 'C' && slen == 2) bsize = (size_t)(d[q + 4] | (d[q + 5] << 8)) + 1;
            q += 4 + slen;
        }
        if (!bsize || p + bsize > n) { free(off); free(cs); free(uo); return NULL; }
        if (nb == cap) { cap *= 2; off = realloc(off, cap * sizeof(size_t)); cs = realloc(cs, cap * sizeof(size_t)); uo = realloc(uo, (cap + 1) * sizeof(size_t)); }
        const uint8_t *t = d + p + bsize - 4;
        off[nb] = p; cs[nb] = bsize; uo[nb] = total;
        total += (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
        nb++; p += bsize;
    }
    uo[nb] = total;
    uint8_t *out = malloc(total + 16);
    long nt = sysconf(_SC_NPROCESSORS_ONLN); if (nt < 1) nt = 1; if (nt > 64) nt = 64; if ((size_t)nt > nb) nt = nb ? (long)nb : 1;
    pthread_t th[64]; bgzf_job job[64];
    for (long t = 0; t < nt; t++) {
        job[t] = (bgzf_job){d, nb, off, cs, uo, out, nb * (size_t)t / (size_t)nt, nb * (size_t)(t + 1) / (size_t)nt, 0};
        pthread_create(&th[t], NULL, bgzf_worker, &job[t]);
    }
    int bad = 0;
    for (long t = 0; t < nt; t++) { pthread_join(th[t], NULL); bad |= job[t].rc; }
    free(off); free(cs); free(uo);
    if (bad) { free(out); return NULL; }
    *n_out = total;
    return out;
}

/* ---- BAM records -> SAM text ---- */
typedef struct { char *s; size_t l, m; } bam_str;
static inline void bs_need(bam_str *b, size_t k) { if (b->l + k > b->m) { b->m = (b->l + k) * 2 + 256; b->s = realloc(b->s, b->m); } }
static inline void bs_putc(bam_str *b, char c) { bs_need(b, 1); b->s[b->l++] = c; }
static inline void bs_put(bam_str *b, const void *p, size_t k) { bs_need(b, k); memcpy(b->s + b->l, p, k); b->l += k; }
static inline void bs_int(bam_str *b, long long v)
{
    char t[24]; int k = 0; unsigned long long u = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
    do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) t[k++] = '-';
    bs_need(b, (size_t)k); while (k) b->s[b->l++] = t[--k];
}
static inline int32_t le32(const uint8_t *p) { return (int32_t)((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24); }
static inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | p[1] << 8); }

static int bam_record_to_sam(const uint8_t *r, size_t block, char **names, int n_ref, bam_str *o)
{
    if (block < 32) return -1;
    const int32_t refID = le32(r), pos = le32(r + 4); const uint8_t l_name = r[8], mapq = r[9];
    const uint16_t n_cig = le16(r + 12), flag = le16(r + 14); const uint32_t l_seq = (uint32_t)le32(r + 16);
    const int32_t nref = le32(r + 20), npos = le32(r + 24), tlen = le32(r + 28);
    size_t p = 32;
    if (p + l_name + 4ull * n_cig + (l_seq + 1) / 2 + l_seq > block || l_name == 0) return -1;
    bs_put(o, r + p, (size_t)l_name - 1); p += l_name;
    bs_putc(o, '\t'); bs_int(o, flag); bs_putc(o, '\t');
    if (refID >= 0 && refID < n_ref) bs_put(o, names[refID], strlen(names[refID])); else bs_putc(o, '*');
    bs_putc(o, '\t'); bs_int(o, (long long)pos + 1); bs_putc(o, '\t'); bs_int(o, mapq); bs_putc(o, '\t');
    if (n_cig == 0) bs_putc(o, '*');
    for (int k = 0; k < n_cig; k++) { const uint32_t c = (uint32_t)le32(r + p + 4 * k); bs_int(o, c >> 4); bs_putc(o, "MIDNSHP=XB??????"[c & 15]); }
    p += 4ull * n_cig;
    bs_putc(o, '\t');
    if (nref < 0) bs_putc(o, '*'); else if (nref == refID) bs_putc(o, '='); else if (nref < n_ref) bs_put(o, names[nref], strlen(names[nref])); else bs_putc(o, '*');
    bs_putc(o, '\t'); bs_int(o, (long long)npos + 1); bs_putc(o, '\t'); bs_int(o, tlen); bs_putc(o, '\t');
    if (l_seq == 0) { bs_put(o, "*\t*", 3); p += 0; }
    else {
        bs_need(o, 2 * (size_t)l_seq + 2);
        for (uint32_t i = 0; i < l_seq; i++) o->s[o->l++] = "=ACMGRSVTWYHKDBN"[(r[p + (i >> 1)] >> ((~i & 1) << 2)) & 15];
        p += (l_seq + 1) / 2;
        o->s[o->l++] = '\t';
        if (r[p] == 0xff) o->s[o->l++] = '*'; else for (uint32_t i = 0; i < l_seq; i++) o->s[o->l++] = (char)(r[p + i] + 33);
        p += l_seq;
    }
    while (p + 3 <= block) {                                     /* optional fields */
        bs_putc(o, '\t'); bs_put(o, r + p, 2); bs_putc(o, ':');
        const uint8_t ty = r[p + 2]; p += 3;
        switch (ty) {
        case 'A': bs_put(o, "A:", 2); bs_putc(o, (char)r[p]); p += 1; break;
        case 'c': bs_put(o, "i:", 2); bs_int(o, (int8_t)r[p]); p += 1; break;
        case 'C': bs_put(o, "i:", 2); bs_int(o, r[p]); p += 1; break;
        case 's': bs_put(o, "i:", 2); bs_int(o, (int16_t)le16(r + p)); p += 2; break;
        case 'S': bs_put(o, "i:", 2); bs_int(o, le16(r + p)); p += 2; break;
        case 'i': bs_put(o, "i:", 2); bs_int(o, le32(r + p)); p += 4; break;
        case 'I': bs_put(o, "i:", 2); bs_int(o, (uint32_t)le32(r + p)); p += 4; break;
        case 'f': { float f; memcpy(&f, r + p, 4); char t[32]; int k = snprintf(t, sizeof t, "%g", f); bs_put(o, "f:", 2); bs_put(o, t, (size_t)k); p += 4; break; }
        case 'Z': case 'H': { bs_putc(o, (char)ty); bs_putc(o, ':'); size_t e = p; while (e < block && r[e]) e++; bs_put(o, r + p, e - p); p = e + 1; break; }
        case 'B': {
            const uint8_t st = r[p]; const uint32_t cnt = (uint32_t)le32(r + p + 1); p += 5;
            bs_put(o, "B:", 2); bs_putc(o, (char)st);
            for (uint32_t i = 0; i < cnt; i++) {
                bs_putc(o, ',');
                switch (st) {
                case 'c': bs_int(o, (int8_t)r[p]); p += 1; break;
                case 'C': bs_int(o, r[p]); p += 1; break;
                case 's': bs_int(o, (int16_t)le16(r + p)); p += 2; break;
                case 'S': bs_int(o, le16(r + p)); p += 2; break;
                case 'i': bs_int(o, le32(r + p)); p += 4; break;
                case 'I': bs_int(o, (uint32_t)le32(r + p)); p += 4; break;
                case 'f': { float f; memcpy(&f, r + p, 4); char t[32]; int k = snprintf(t, sizeof t, "%g", f); bs_put(o, t, (size_t)k); p += 4; break; }
                default: return -1;
                }
            }
            break;
        }
        default: return -1;
        }
        if (p > block) return -1;
    }
    bs_putc(o, '\n');
    return 0;
}

typedef struct { const uint8_t *bam; const size_t *rec_off; size_t lo, hi; char **names; int n_ref; bam_str out; int rc; } bam_job;
static void *bam_worker(void *arg)
{
    bam_job *j = (bam_job *)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        const uint8_t *r = j->bam + j->rec_off[i];
        if (bam_record_to_sam(r + 4, (size_t)le32(r), j->names, j->n_ref, &j->out)) { j->rc = -1; return NULL; }
    }
    return NULL;
}

/* Uncompressed BAM -> SAM text (header + alignment lines), malloc'd.  NULL on a malformed file. */
static uint8_t *bam_to_sam_text(const uint8_t *bam, size_t n, size_t *n_out)
{
    if (n < 12 || memcmp(bam, "BAM\1", 4)) return NULL;
    const size_t l_text = (size_t)(uint32_t)le32(bam + 4);
    if (8 + l_text + 4 > n) return NULL;
    size_t p = 8 + l_text;
    const int n_ref = le32(bam + p); p += 4;
    if (n_ref < 0) return NULL;
    char **names = malloc(sizeof(char *) * (size_t)(n_ref + 1)); int64_t *lens = malloc(sizeof(int64_t) * (size_t)(n_ref + 1));
    for (int i = 0; i < n_ref; i++) {
        if (p + 4 > n) return NULL;
        const size_t ln = (size_t)(uint32_t)le32(bam + p); p += 4;
        if (p + ln + 4 > n || ln == 0) return NULL;
        names[i] = strndup((const char *)bam + p, ln - 1); p += ln;
        lens[i] = (uint32_t)le32(bam + p); p += 4;
    }
    bam_str hdr = {0, 0, 0};
    size_t tl = l_text; while (tl && bam[8 + tl - 1] == 0) tl--;
    bs_put(&hdr, bam + 8, tl);
    if (tl && hdr.s[hdr.l - 1] != '\n') bs_putc(&hdr, '\n');
    int has_sq = 0;
    for (size_t i = 0; i + 3 < hdr.l; i++) if ((i == 0 || hdr.s[i - 1] == '\n') && !memcmp(hdr.s + i, "@SQ", 3)) { has_sq = 1; break; }
    if (!has_sq) for (int i = 0; i < n_ref; i++) { bs_put(&hdr, "@SQ\tSN:", 7); bs_put(&hdr, names[i], strlen(names[i])); bs_put(&hdr, "\tLN:", 4); bs_int(&hdr, lens[i]); bs_putc(&hdr, '\n'); }
    /* record boundaries, then one range of records per thread */
    size_t cap = 1 << 16, nr = 0; size_t *ro = malloc(cap * sizeof(size_t));
    while (p + 4 <= n) {
        const size_t bl = (size_t)(uint32_t)le32(bam + p);
        if (p + 4 + bl > n) { free(ro); return NULL; }
        if (nr == cap) { cap *= 2; ro = realloc(ro, cap * sizeof(size_t)); }
        ro[nr++] = p; p += 4 + bl;
    }
    long nt = sysconf(_SC_NPROCESSORS_ONLN); if (nt < 1) nt = 1; if (nt > 64) nt = 64; if ((size_t)nt > nr) nt = nr ? (long)nr : 1;
    pthread_t th[64]; bam_job job[64];
    for (long t = 0; t < nt; t++) {
        job[t] = (bam_job){bam, ro, nr * (size_t)t / (size_t)nt, nr * (size_t)(t + 1) / (size_t)nt, names, n_ref, {0, 0, 0}, 0};
        pthread_create(&th[t], NULL, bam_worker, &job[t]);
    }
    int bad = 0; size_t total = hdr.l;
    for (long t = 0; t < nt; t++) { pthread_join(th[t], NULL); bad |= job[t].rc; total += job[t].out.l; }
    uint8_t *out = NULL;
    if (!bad) {
        out = malloc(total + 16);
        memcpy(out, hdr.s, hdr.l); size_t w = hdr.l;
        for (long t = 0; t < nt; t++) { memcpy(out + w, job[t].out.s, job[t].out.l); w += job[t].out.l; }
        *n_out = total;
    }
    for (long t = 0; t < nt; t++) free(job[t].out.s);
    for (int i = 0; i < n_ref; i++) free(names[i]);
    free(names); free(lens); free(ro); free(hdr.s);
    return out;
}
#endif
