/*
 * oracle/tnc_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU restatement, not product).
 *
 * Plain-C restatement of the trinucleotide-context scan of the reference's
 * tncCountsProfile (reference: tncCountsProfile.c:366-483).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may build, load or run this file.  The product path (libssb200.so) never does.
 *
 * The reference walks every line with a 3-byte carry buffer
 * (tncCountsProfile.c:384-446).  That state machine is equivalent to the
 * line-oriented closed form below, which is how this file is written:
 *
 *   - a "record" is what getline() returns (tncCountsProfile.c:391); the '\n'
 *     is replaced by NUL (:394);
 *   - a record is KEPT unless it is empty or starts with '>' (:398) or holds no
 *     upper-case G/C/A/T at all (:401-407); non-kept records leave the carry
 *     buffer untouched, so contigs are joined;
 *   - inside a kept record every window of 3 upper-case bases counts (:430-438);
 *   - the per-character loop runs over `read` bytes INCLUDING the newline slot
 *     (:409), so the NUL lands in tncBuff[2], the window (l[-2],l[-1],next[0])
 *     is lost and only l[-1] is carried (:441-443).  The next kept record then
 *     contributes exactly one straddling window (carry, l'[0], l'[1]) if it has
 *     at least 2 characters; a 1-character record just replaces the carry.
 *
 * Pinned against the compiled reference (oracle/_ref/tncCountsProfile_ref) by
 * tests/test_tnc_oracle.py on the known-answer vectors of SURVEY.md App. B and
 * on seeded fuzz files.  Inputs holding NUL bytes are rejected (-2): the
 * reference's strcspn/strlen treatment of embedded NULs is not restated.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* index = 16*a + 4*b + c with A<C<G<T (tncCountsProfile.c:14-77) */
static inline int base_code(unsigned char ch)
{
    switch (ch) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default:  return -1;
    }
}

static inline void bump(int64_t *cnt, int a, int b, int c)
{
    if (a >= 0 && b >= 0 && c >= 0)          /* tncCountsProfile.c:430-433 */
        cnt[16 * a + 4 * b + c]++;           /* incCtx, tncCountsProfile.c:105-363 */
}

/* Streaming form so large files can be fed in pieces: state carried between calls. */
typedef struct {
    int carry;        /* base code (or -1) of the last byte of the previous kept,
                         newline-terminated record; -2 = no carry yet            */
    /* partial record (no '\n' seen yet) is buffered by the caller: this oracle
       only accepts whole records except for the final call.                     */
} tnc_oracle_state;

int tnc_oracle_count(const unsigned char *buf, size_t n, int64_t cnt[64])
{
    int carry = -2;
    size_t pos = 0;
    if (memchr(buf, 0, n) != NULL) return -2;
    while (pos < n) {
        const unsigned char *line = buf + pos;
        const unsigned char *nl = memchr(line, '\n', n - pos);
        size_t len = nl ? (size_t)(nl - line) : n - pos;
        pos += len + (nl ? 1 : 0);

        if (len == 0 || line[0] == '>') continue;            /* :398 */
        int has_base = 0;
        for (size_t i = 0; i < len; i++)                      /* :401-407 */
            if (base_code(line[i]) >= 0) { has_base = 1; break; }
        if (!has_base) continue;

        /* straddling window from the previous kept record (:413-421) */
        if (carry != -2 && len >= 2)
            bump(cnt, carry, base_code(line[0]), base_code(line[1]));
        /* in-line windows */
        for (size_t i = 0; i + 2 < len; i++)
            bump(cnt, base_code(line[i]), base_code(line[i + 1]), base_code(line[i + 2]));
        /* only a newline-terminated record leaves a 1-byte carry (:409, :441-443);
           an unterminated record can only be the last one.                      */
        if (nl) carry = base_code(line[len - 1]);
    }
    return 0;
}

static const char *const OUT_ORDER[32] = {           /* tncCountsProfile.c:452-483 */
    "ACA","ACC","ACG","ACT","ATA","ATC","ATG","ATT","CCA","CCC","CCG","CCT","CTA","CTC","CTG","CTT",
    "GCA","GCC","GCG","GCT","GTA","GTC","GTG","GTT","TCA","TCC","TCG","TCT","TTA","TTC","TTG","TTT"};

static int ctx_index(const char *s) { return 16 * base_code(s[0]) + 4 * base_code(s[1]) + base_code(s[2]); }
static int revcomp_index(const char *s)
{
    /* complement: A<->T (0<->3), C<->G (1<->2)  => 3-code */
    return 16 * (3 - base_code(s[2])) + 4 * (3 - base_code(s[1])) + (3 - base_code(s[0]));
}

/* 32 unstranded totals in the reference's output order. */
void tnc_oracle_fold(const int64_t cnt[64], int64_t out32[32])
{
    for (int i = 0; i < 32; i++)
        out32[i] = cnt[ctx_index(OUT_ORDER[i])] + cnt[revcomp_index(OUT_ORDER[i])];
}

int tnc_oracle_format(const int64_t cnt[64], char *dst, size_t cap)
{
    int64_t o[32];
    size_t used = 0;
    tnc_oracle_fold(cnt, o);
    for (int i = 0; i < 32; i++) {
        int w = snprintf(dst + used, cap - used, "%s\t%ld\n", OUT_ORDER[i], (long)o[i]);
        if (w < 0 || (size_t)w >= cap - used) return -1;
        used += (size_t)w;
    }
    return (int)used;
}

#ifdef TNC_ORACLE_MAIN
int main(int argc, char **argv)
{
    FILE *fp = argc > 1 ? fopen(argv[1], "rb") : NULL;
    if (!fp) exit(EXIT_FAILURE);                              /* :380-382 */
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    unsigned char *buf = malloc(n > 0 ? (size_t)n : 1);
    if (n > 0 && fread(buf, 1, (size_t)n, fp) != (size_t)n) exit(EXIT_FAILURE);
    fclose(fp);
    int64_t cnt[64] = {0};
    if (tnc_oracle_count(buf, (size_t)n, cnt) != 0) { fprintf(stderr, "NUL byte in input\n"); return 2; }
    char out[4096];
    int w = tnc_oracle_format(cnt, out, sizeof out);
    fwrite(out, 1, (size_t)w, stdout);
    free(buf);
    return 0;
}
#endif
