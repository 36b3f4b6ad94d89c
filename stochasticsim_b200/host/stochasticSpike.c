/*
 * stochasticSpike -- drop-in replacement for the reference program of the same name.
 *
 *   stochasticSpike <donor BAM> <donor reference> <somatic mutation config file> <seed> <output SAM filename>
 *
 * Same argv, same files (argv[5] SAM, ./truth.vcf in the CWD, stats block on stdout) and the same exit
 * codes as main() of stochasticSpike.c:911-1671, so bin/spikeIn.bash:41,46 can call it unchanged.  This
 * file stays in C and does only what a host has to do: read the inputs, split the header, parse the
 * `.spike` table (getNextTarget, :98-158), hand everything to libssb200.so, and print what comes
 * back.  The pileup loop itself (:1129-1623) runs on the GPU; there is no CPU fallback: without a
 * B200 the program fails with exit status 3.
 *
 * Input format: argv[1] is a coordinate-sorted BAM (BGZF; its .bai/.csi must exist, as at :1035-1039, although the
 * reference never uses it, :1078-1080) or SAM text ("-" = stdin), auto-detected like htslib's sam_open.  BAM is
 * decoded on the host (bam_input.h: parallel BGZF inflate + sam_format1-style printing) into the SAM text the
 * GPU tokeniser consumes.
 *
 * Environment: SSB_DEVICE=i selects the (first) CUDA device (default 0).  SSB_GPUS=N spreads ONE input over N GPUs: the body is
 * cut into coordinate ranges (ssb_spike_plan_shards), one thread drives each shard, the shards hand the rand() offset on
 * and the outputs are concatenated in shard order (ssb_spike_run_shard_host).  SSB_SHARDS=M (default N) makes M shards and
 * places shard g on device g mod N; SSB_HALO=<bases> is the largest reference span a read may have (default 4096; the run
 * is repeated with the measured value when a read turns out longer).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <ctype.h>
#include <signal.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <pthread.h>
#include "ssb200.h"
#include "bam_input.h"

#define VERSION "0.01"                        /* stochasticSpike.c:1 */

typedef struct { char **name; int64_t *len; uint8_t **seq; int n, cap; } fasta_t;

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "out of memory\n"); exit(1); } return q; }

/* whole file into memory; *mapped says how to release it */
static uint8_t *slurp(const char *fn, size_t *n_out, int *mapped)
{
    int fd = strcmp(fn, "-") == 0 ? 0 : open(fn, O_RDONLY);
    *mapped = 0; *n_out = 0;
    if (fd < 0) return NULL;
    struct stat sb;
    if (fd != 0 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size > 0) {
        void *m = mmap(NULL, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) { madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL); *n_out = (size_t)sb.st_size; *mapped = 1; close(fd); return (uint8_t *)m; }
    }
    size_t cap = 1 << 22, len = 0; uint8_t *buf = xrealloc(NULL, cap); ssize_t r;
    while ((r = read(fd, buf + len, cap - len)) > 0) { len += (size_t)r; if (len == cap) { cap *= 2; buf = xrealloc(buf, cap); } }
    if (fd != 0) close(fd);
    *n_out = len;
    return buf;
}

/* faidx_fetch_seq64(whole contig): white space dropped, case preserved (stochasticSpike.c:215-219) */
static int load_fasta(const char *fn, fasta_t *fa)
{
    size_t n; int mapped;
    uint8_t *d = slurp(fn, &n, &mapped);
    memset(fa, 0, sizeof *fa);
    if (!d && n == 0 && access(fn, R_OK) != 0) return -1;
    size_t p = 0; int cur = -1; size_t scap = 0;
    while (p < n) {
        const uint8_t *nl = memchr(d + p, '\n', n - p);
        size_t e = nl ? (size_t)(nl - d) : n;
        if (d[p] == '>') {
            if (fa->n == fa->cap) {
                fa->cap = fa->cap ? fa->cap * 2 : 32;
                fa->name = xrealloc(fa->name, sizeof(char *) * (size_t)fa->cap);
                fa->len = xrealloc(fa->len, sizeof(int64_t) * (size_t)fa->cap);
                fa->seq = xrealloc(fa->seq, sizeof(uint8_t *) * (size_t)fa->cap);
            }
            size_t q = p + 1; while (q < e && !isspace(d[q])) q++;
            cur = fa->n++;
            fa->name[cur] = strndup((const char *)d + p + 1, q - p - 1); fa->len[cur] = 0; fa->seq[cur] = NULL; scap = 0;
        } else if (cur >= 0) {
            if ((size_t)fa->len[cur] + (e - p) + 1 > scap) {
                if (!scap) scap = 1 << 20;
                while ((size_t)fa->len[cur] + (e - p) + 1 > scap) scap *= 2;
                fa->seq[cur] = xrealloc(fa->seq[cur], scap);
            }
            for (size_t i = p; i < e; i++) if (isgraph(d[i])) fa->seq[cur][fa->len[cur]++] = d[i];
        }
        p = e + 1;
    }
    if (mapped) munmap(d, n); else free(d);
    return 0;
}

typedef struct { char *contig; ssb_target t; } target_rec;

/* getNextTarget (stochasticSpike.c:98-158) applied to the whole file: one record per valid line, FILE order */
static size_t load_targets(FILE *fp, char **names, int n_names, target_rec **out)
{
    size_t n = 0, cap = 0; target_rec *v = NULL;
    char *line = NULL; size_t lcap = 0;
    while (getline(&line, &lcap, fp) != -1) {
        if (line[0] == '#') continue;
        if (!strlen(line)) continue;
        char *tok = strtok(line, "\t");
        if (!tok) continue;
        char contig[1024];
        strncpy(contig, tok, sizeof contig - 1); contig[sizeof contig - 1] = 0;
        if (!(tok = strtok(NULL, "\t"))) continue;
        long locus = atol(tok) - 1;
        if (!(tok = strtok(NULL, "\t"))) continue;
        char base = *tok;
        if (!(tok = strtok(NULL, "\t"))) continue;
        float af = (float)atof(tok);
        if (n == cap) { cap = cap ? cap * 2 : 1024; v = xrealloc(v, cap * sizeof *v); }
        memset(&v[n], 0, sizeof v[n]);
        v[n].contig = strdup(contig);
        v[n].t.c_tid = -1;
        for (int i = 0; i < n_names; i++) if (strcmp(names[i], contig) == 0) { v[n].t.c_tid = i; break; }   /* sam_hdr_name2tid, :122 */
        v[n].t.locus = locus; v[n].t.base = (uint8_t)base; v[n].t.af = af;
        n++;
    }
    free(line);
    *out = v;
    return n;
}

static const char *const FILTER_NAME[] = {"NONE", "PASS", "MASKED", "MASKED_OVL", "UNDETECTED"};

/* error alleles with a count, in descending count, stable over the reference's G,C,A,T order (:500-523) */
static void err_order(const int32_t cnt[4], int idx[4])
{
    for (int i = 0; i < 4; i++) idx[i] = i;
    for (int step = 0; step < 3; ++step) {
        int sw = 0;
        for (int i = 0; i < 3 - step; ++i) if (cnt[idx[i]] < cnt[idx[i + 1]]) { int t = idx[i]; idx[i] = idx[i + 1]; idx[i + 1] = t; sw = 1; }
        if (!sw) break;
    }
}
static void print_err_list(FILE *vcf, const int32_t cnt[4], const int idx[4], int as_counts)
{
    for (int i = 0; i < 4; i++) {
        if (cnt[idx[i]] == 0) break;
        if (as_counts) fprintf(vcf, "%d", cnt[idx[i]]); else fputc("GCAT"[idx[i]], vcf);
        if (i + 1 == 4 || cnt[idx[i + 1]] == 0) break;
        fputc(',', vcf);
    }
}

static void print_seq_error(FILE *vcf, char **names, const ssb_seq_error *e)                 /* :1494-1557 */
{
    int idx[4]; err_order(e->err_cnt, idx);
    int tot = e->err_cnt[0] + e->err_cnt[1] + e->err_cnt[2] + e->err_cnt[3], dp = e->ref_cnt + tot;
    fprintf(vcf, "%s\t%d\t.\t%c\t", names[e->tid], (int)e->pos + 1, e->ref_base);
    print_err_list(vcf, e->err_cnt, idx, 0);
    fprintf(vcf, "\t.\tSEQ_ERROR\tDP=%d;AF=%.6g\tAD\t%d,", dp, (float)tot / dp, e->ref_cnt);
    print_err_list(vcf, e->err_cnt, idx, 1);
    fputc('\n', vcf);
}

static void print_target(FILE *vcf, char **names, const ssb_target_result *r, float af)      /* :1406-1470 */
{
    int idx[4]; err_order(r->err_cnt, idx);
    int tot = r->err_cnt[0] + r->err_cnt[1] + r->err_cnt[2] + r->err_cnt[3];
    fprintf(vcf, "%s\t%d\t.\t%c\t%c", names[r->at_tid], (int)r->at_pos + 1, r->ref_base, r->mutant_allele);
    if (tot) { fputc(',', vcf); print_err_list(vcf, r->err_cnt, idx, 0); }
    fprintf(vcf, "\t.\t%s\tDP=%d;AF=%.6g\tAD\t%d,%d", FILTER_NAME[r->filter], r->ref_cnt + r->mut_cnt + tot, af, r->ref_cnt, r->mut_cnt);
    if (tot) { fputc(',', vcf); print_err_list(vcf, r->err_cnt, idx, 1); }
    fputc('\n', vcf);
}

static void print_no_coverage(FILE *vcf, const target_rec *t)                                 /* :1604-1614, :1632-1642 */
{
    fprintf(vcf, "%s\t%ld\t.\t.\t.\t.\tNO_COVERAGE\t.\t.\n", t->contig, (long)t->t.locus + 1);
}

/* one shard of the input on one GPU (SSB_GPUS / SSB_SHARDS); a plain run is the single shard of a group of one */
#define MAX_SHARDS 64
typedef struct {
    int device; ssb_spike_shard shard; ssb_exchange *xc;
    const uint8_t *body; size_t n;
    char **names; int n_names; const fasta_t *fa;
    const ssb_target *targets; size_t T; unsigned seed;
    ssb_target_result *res; uint8_t *out; size_t out_n; ssb_spike_stats st;
    ssb_seq_error *se; size_t n_se;
    int rc; char errtext[600];
} shard_job;

static void *shard_main(void *arg)
{
    shard_job *j = (shard_job *)arg;
    ssb_ctx *ctx = NULL; ssb_spike *sp = NULL;
    j->rc = ssb_ctx_create(j->device, &ctx);
    if (j->rc) { snprintf(j->errtext, sizeof j->errtext, "device %d", j->device); goto fail; }
    {
        /* only the contigs this shard can touch are uploaded */
        ssb_contig *contigs = xrealloc(NULL, sizeof(ssb_contig) * (size_t)(j->n_names + 1));
        for (int t = 0; t < j->n_names; t++) {
            contigs[t].name = j->names[t]; contigs[t].len = 0; contigs[t].seq = NULL;
            if (j->shard.count > 1 && (t < j->shard.lo_tid || t > j->shard.hi_tid)) continue;
            for (int i = 0; i < j->fa->n; i++) if (!strcmp(j->fa->name[i], j->names[t])) { contigs[t].len = j->fa->len[i]; contigs[t].seq = j->fa->seq[i]; break; }
        }
        j->rc = ssb_spike_create(ctx, contigs, j->n_names, &sp);
        free(contigs);
    }
    if (!j->rc) j->rc = ssb_spike_run_shard_host(sp, &j->shard, j->xc, j->body, j->n, j->out, j->n + 1, j->targets, j->T, j->seed, j->res, &j->st, &j->out_n);
    if (!j->rc) {
        ssb_spike_seq_error_count(sp, &j->n_se);
        j->se = xrealloc(NULL, sizeof(ssb_seq_error) * (j->n_se + 1));
        j->rc = ssb_spike_seq_errors(sp, j->se, j->n_se);
    }
    if (j->rc) snprintf(j->errtext, sizeof j->errtext, "%s", ssb_last_error(ctx));
fail:
    if (sp) ssb_spike_destroy(sp);
    if (ctx) ssb_ctx_destroy(ctx);
    return NULL;
}

int main(int argc, char **argv)
{
    char *cmd = strrchr(argv[0], '/'); cmd = cmd ? cmd + 1 : argv[0];                       /* :931-936 */
    if (argc != 6) {                                                                          /* :938-941 */
        fprintf(stderr, "\n Usage: %s <donor BAM> <donor reference> <somatic mutation config file> <seed> <output SAM filename>\n\n", cmd);
        exit(0);
    }
    FILE *vcf = fopen("truth.vcf", "w");                                                      /* :944 */
    unsigned seed = (unsigned)atoi(argv[4]);                                                  /* :948 */

    size_t n_in; int mapped;
    uint8_t *in = slurp(argv[1], &n_in, &mapped);
    if (!in) { fprintf(stderr, "Couldn't open bam...\n"); return 1; }                         /* :962-965 */
    int bgzf_input = 0;
    if (bam_is_bgzf(in, n_in)) {
        /* BGZF: inflate the blocks in parallel; a BAM inside is printed as SAM text (htslib's sam_open auto-detects the same way) */
        size_t n_raw = 0, n_txt = 0;
        uint8_t *raw = bgzf_inflate_all(in, n_in, &n_raw);
        if (!raw) { fprintf(stderr, "Couldn't read header...\n"); return 1; }                        /* :971-975 */
        if (mapped) munmap(in, n_in); else free(in);
        mapped = 0;
        if (n_raw >= 4 && !memcmp(raw, "BAM\1", 4)) {
            uint8_t *txt = bam_to_sam_text(raw, n_raw, &n_txt);
            free(raw);
            if (!txt) { fprintf(stderr, "Couldn't read header...\n"); return 1; }
            in = txt; n_in = n_txt;
        } else { in = raw; n_in = n_raw; }                                                            /* bgzip-compressed SAM text */
        bgzf_input = 1;
    }
    /* header = leading lines that start with '@' (sam_hdr_read, :971) */
    size_t hdr_end = 0;
    while (hdr_end < n_in && in[hdr_end] == '@') {
        const uint8_t *nl = memchr(in + hdr_end, '\n', n_in - hdr_end);
        hdr_end = nl ? (size_t)(nl - in) + 1 : n_in;
    }
    char **names = NULL; int64_t *lens = NULL; int n_names = 0, cap_names = 0;
    char *sample = NULL; int seen_rg = 0;
    for (size_t p = 0; p < hdr_end;) {
        const uint8_t *nl = memchr(in + p, '\n', hdr_end - p);
        size_t e = nl ? (size_t)(nl - in) : hdr_end;
        if (e - p >= 4 && !memcmp(in + p, "@SQ\t", 4)) {
            char *copy = strndup((const char *)in + p, e - p), *save = NULL, *sn = NULL; int64_t ln = 0;
            for (char *tok = strtok_r(copy, "\t\r", &save); tok; tok = strtok_r(NULL, "\t\r", &save)) {
                if (!strncmp(tok, "SN:", 3)) sn = tok + 3; else if (!strncmp(tok, "LN:", 3)) ln = strtoll(tok + 3, NULL, 10);
            }
            if (sn) {
                if (n_names == cap_names) { cap_names = cap_names ? cap_names * 2 : 64; names = xrealloc(names, sizeof(char *) * (size_t)cap_names); lens = xrealloc(lens, sizeof(int64_t) * (size_t)cap_names); }
                names[n_names] = strdup(sn); lens[n_names] = ln; n_names++;
            }
            free(copy);
        } else if (!seen_rg && e - p >= 4 && !memcmp(in + p, "@RG\t", 4)) {                   /* first @RG's SM (:998-1014) */
            seen_rg = 1;
            size_t q = p + 4;
            while (q < e) {
                const uint8_t *t = memchr(in + q, '\t', e - q); size_t fe = t ? (size_t)(t - in) : e;
                if (fe - q >= 3 && in[q] == 'S' && in[q + 1] == 'M' && in[q + 2] == ':') {
                    size_t ve = fe; while (ve > q + 3 && in[ve - 1] == '\r') ve--;
                    sample = strndup((const char *)in + q + 3, ve - q - 3); break;
                }
                q = fe + 1;
            }
            if (!sample) { fprintf(stderr, "Couldn't read sample name from header...\n"); exit(-1); }   /* :1006-1009 */
        }
        p = e + 1;
    }
    FILE *out = fopen(argv[5], "w");                                                          /* :980 */
    if (!out) { fprintf(stderr, "Couldn't write out ...\n"); return 1; }
    if (hdr_end) {
        fwrite(in, 1, hdr_end, out);                                                          /* :986 header verbatim */
        if (in[hdr_end - 1] != '\n') fputc('\n', out);
    }
    fprintf(vcf, "##fileformat=VCFv4.2\n##FILTER=<ID=PASS,Description=\"All filters passed\">\n");        /* :1017-1032 */
    fprintf(vcf, "##%sVersion=%s\n##%sCommand=%s %s %s %s %s\n", cmd, VERSION, cmd, argv[1], argv[2], argv[3], argv[4], argv[5]);
    fprintf(vcf, "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n", sample ? sample : "SAMPLE");

    if (bgzf_input && strcmp(argv[1], "-") != 0) {                                            /* :1035-1039: the index must exist (checked where the reference checks: after both headers) */
        char ip[4200]; int have = 0;
        snprintf(ip, sizeof ip, "%s.bai", argv[1]); if (access(ip, R_OK) == 0) have = 1;
        snprintf(ip, sizeof ip, "%s.csi", argv[1]); if (access(ip, R_OK) == 0) have = 1;
        size_t al = strlen(argv[1]);
        if (al > 4 && !strcmp(argv[1] + al - 4, ".bam")) { snprintf(ip, sizeof ip, "%.*s.bai", (int)(al - 4), argv[1]); if (access(ip, R_OK) == 0) have = 1; }
        if (!have) { fflush(vcf); fflush(out); fprintf(stderr, "\nCan't load index for %s..\n\n", argv[1]); exit(EXIT_FAILURE); }
    }
    fasta_t fa;
    if (load_fasta(argv[2], &fa) < 0) { fprintf(stderr, "Could not load faidx: %s\n", argv[2]); return 1; }   /* :1042-1046 */
    FILE *cfg = fopen(argv[3], "r");
    if (!cfg) { fprintf(stderr, "\nCan't open %s..\n\n", argv[3]); return -1; }                            /* :1050-1053 */
    target_rec *tg; size_t T = load_targets(cfg, names, n_names, &tg);
    fclose(cfg);

    /* ---- GPU(s) ---- */
    int dev = getenv("SSB_DEVICE") ? atoi(getenv("SSB_DEVICE")) : 0;
    int ngpu = getenv("SSB_GPUS") ? atoi(getenv("SSB_GPUS")) : 1;
    int want = getenv("SSB_SHARDS") ? atoi(getenv("SSB_SHARDS")) : ngpu;
    int64_t halo = getenv("SSB_HALO") ? atoll(getenv("SSB_HALO")) : 4096;
    if (ngpu < 1) ngpu = 1;
    if (want < 1) want = 1;
    if (want > MAX_SHARDS) want = MAX_SHARDS;
    ssb_target *tarr = xrealloc(NULL, sizeof(ssb_target) * (T + 1));
    ssb_target_result *res = xrealloc(NULL, sizeof(ssb_target_result) * (T + 1));
    for (size_t t = 0; t < T; t++) tarr[t] = tg[t].t;
    size_t body_n = n_in - hdr_end, out_n = 0;
    uint8_t *body_out = xrealloc(NULL, body_n + 2);
    ssb_spike_stats st;
    ssb_seq_error *se = NULL; size_t n_se = 0;
    int rc;
    for (int round = 0;; round++) {
        shard_job job[MAX_SHARDS]; ssb_spike_shard plan[MAX_SHARDS]; size_t off[MAX_SHARDS], len[MAX_SHARDS]; ssb_exchange *xcs[MAX_SHARDS];
        int made = ssb_spike_plan_shards(in + hdr_end, body_n, (const char *const *)names, n_names, want, halo, plan, off, len);
        if (made < 1) { fprintf(stderr, "%s: cannot cut the input into shards\n", cmd); return 3; }
        if (made > 1 && ssb_exchange_local_create(made, xcs)) { fprintf(stderr, "%s: exchange\n", cmd); return 3; }
        memset(job, 0, sizeof job);
        for (int g = 0; g < made; g++) {
            job[g].device = dev + g % ngpu; job[g].shard = plan[g]; job[g].xc = made > 1 ? xcs[g] : NULL;
            job[g].body = in + hdr_end + off[g]; job[g].n = len[g];
            job[g].names = names; job[g].n_names = n_names; job[g].fa = &fa;
            job[g].targets = tarr; job[g].T = T; job[g].seed = seed;
            job[g].res = xrealloc(NULL, sizeof(ssb_target_result) * (T + 1));
            job[g].out = xrealloc(NULL, len[g] + 2);
        }
        pthread_t th[MAX_SHARDS];
        for (int g = 1; g < made; g++) pthread_create(&th[g], NULL, shard_main, &job[g]);
        shard_main(&job[0]);
        for (int g = 1; g < made; g++) pthread_join(th[g], NULL);
        if (made > 1) for (int g = 0; g < made; g++) ssb_exchange_destroy(xcs[g]);
        rc = 0; int shard_err = 0; unsigned maxspan_msg = 0;
        for (int g = 0; g < made; g++) if (job[g].rc && job[g].rc != SSB_E_PEER) { if (!rc) rc = job[g].rc; if (job[g].rc == SSB_E_SHARD) shard_err = 1; }
        if (!rc) for (int g = 0; g < made; g++) if (job[g].rc) rc = job[g].rc;
        if (shard_err && round == 0 && made > 1) {
            /* a read is longer than the halo: take the next power of two that surely covers it and cut again */
            halo = halo < 65536 ? 65536 : halo * 16; (void)maxspan_msg;
            for (int g = 0; g < made; g++) { free(job[g].res); free(job[g].out); free(job[g].se); }
            continue;
        }
        if (rc) {
            for (int g = 0; g < made; g++) if (job[g].rc == rc) { fprintf(stderr, "%s: %s (%s)\n", cmd, ssb_strerror(job[g].rc), job[g].errtext); break; }
            return rc == SSB_E_NODEVICE ? 3 : (rc == SSB_E_REF ? 1 : 3);
        }
        /* the shards' outputs in shard order are the reference's output; the stats block is their sum (maxDepth: maximum) */
        memset(&st, 0, sizeof st);
        for (int g = 0; g < made; g++) {
            memcpy(body_out + out_n, job[g].out, job[g].out_n); out_n += job[g].out_n;
            st.alignmentCount += job[g].st.alignmentCount; st.numberOfLociCovered += job[g].st.numberOfLociCovered;
            st.totalFoldCoverage += job[g].st.totalFoldCoverage; if (job[g].st.maxDepth > st.maxDepth) st.maxDepth = job[g].st.maxDepth;
            n_se += job[g].n_se;
        }
        se = xrealloc(NULL, sizeof(ssb_seq_error) * (n_se + 1));
        size_t k = 0;
        for (int g = 0; g < made; g++) { if (job[g].n_se) memcpy(se + k, job[g].se, job[g].n_se * sizeof *se); k += job[g].n_se; }
        for (size_t t = 0; t < T; t++) {
            int found = 0;
            for (int g = 0; g < made; g++) if (job[g].res[t].status != SSB_T_ELSEWHERE) { res[t] = job[g].res[t]; found = 1; break; }
            if (!found) { fprintf(stderr, "%s: target %zu was claimed by no shard\n", cmd, t); return 3; }
        }
        for (int g = 0; g < made; g++) { free(job[g].res); free(job[g].out); free(job[g].se); }
        break;
    }
    if (out_n && fwrite(body_out, 1, out_n, out) != out_n) { fprintf(stderr, "Couldn't write out ...\n"); return 1; }   /* :275-278 */

    /* ---- truth.vcf: per covered locus its own line, then the NO_COVERAGE line of a target passed there ---- */
    size_t si = 0, t = 0;
    for (; t < T && res[t].status != SSB_T_TAIL; t++) {
        int64_t li = res[t].locus_index;
        if (res[t].status == SSB_T_HIT) {
            while (si < n_se && se[si].locus_index < li) print_seq_error(vcf, names, &se[si++]);
            print_target(vcf, names, &res[t], tg[t].t.af);
        } else {
            while (si < n_se && se[si].locus_index <= li) print_seq_error(vcf, names, &se[si++]);
            if (res[t].status == SSB_T_NOCOV) print_no_coverage(vcf, &tg[t]);
        }
    }
    while (si < n_se) print_seq_error(vcf, names, &se[si++]);
    for (; t < T; t++) print_no_coverage(vcf, &tg[t]);                                        /* :1630-1646 */

    fclose(vcf); fclose(out);
    fflush(stderr);
    if (st.numberOfLociCovered == 0) { fflush(stdout); raise(SIGFPE); }                       /* the reference divides by zero at :1668 */
    printf("\nDONE...\nalignmentCount (#reads) = %ld,\nnumberOfLociCovered = %ld\ntotalFoldCoverage = %ld\nmaxDepth = %ld, Avg. coverage = %ld\n",
           (long)st.alignmentCount, (long)st.numberOfLociCovered, (long)st.totalFoldCoverage, (long)st.maxDepth,
           (long)(st.totalFoldCoverage / st.numberOfLociCovered));
    return 0;
}
