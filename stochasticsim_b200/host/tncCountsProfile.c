/*
 * tncCountsProfile -- drop-in replacement for the reference program of the same name.
 *
 *   tncCountsProfile <target fasta>            (same argv, stdout and exit codes as the reference)
 *
 * Replaces main() of tncCountsProfile.c:366-485.  This file stays in C and only does what a host
 * has to do -- open/map the file, hand byte ranges to the GPU(s), print the 32 lines.  All
 * counting happens in libssb200.so (CUDA, sm_100a); there is no CPU fallback: without a B200
 * the program fails with exit status 3.
 *
 * Environment (the reference reads none; these only select hardware):
 *   SSB_GPUS=N      spread the file over the first N GPUs by byte range (default 1); the per-GPU
 *                   count vectors are combined with ONE NCCL all-reduce of 64 int64.
 *   SSB_DEVICE=i    first device ordinal (default 0).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <fcntl.h>
#include <unistd.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "ssb200.h"

int ssb_nccl_init_all(ssb_ctx **ctxs, int n, void **comms_out);
int ssb_tnc_allreduce_group(ssb_ctx **ctxs, void **comms, int64_t **d_counts64, int n);
void ssb_nccl_destroy_all(void **comms, int n);

typedef struct {
    ssb_ctx *ctx;
    const uint8_t *base;
    size_t off, len;
    ssb_tnc_carry carry_in;
    int64_t counts[64];
    int rc;
} shard_t;

static void *shard_main(void *arg)
{
    shard_t *s = (shard_t *)arg;
    s->rc = ssb_tnc_count_host(s->ctx, s->base + s->off, s->len, &s->carry_in, NULL, s->counts);
    return NULL;
}

int main(int argc, char **argv)
{
    (void)argc;
    /* tncCountsProfile.c:380-382: fopen failure (or no argument) -> silent exit(EXIT_FAILURE) */
    int fd = argv[1] ? open(argv[1], O_RDONLY) : -1;
    if (fd < 0) exit(EXIT_FAILURE);
    struct stat sb;
    if (fstat(fd, &sb) != 0) exit(EXIT_FAILURE);
    size_t n = (size_t)sb.st_size;
    const uint8_t *data = (const uint8_t *)"";
    if (n) {
        data = mmap(NULL, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (data == MAP_FAILED) {
            /* not mappable (pipe, /dev/stdin): slurp it */
            size_t cap = 1 << 20, len = 0; uint8_t *buf = malloc(cap); ssize_t r;
            while ((r = read(fd, buf + len, cap - len)) > 0) { len += (size_t)r; if (len == cap) { cap *= 2; buf = realloc(buf, cap); } }
            data = buf; n = len;
        } else madvise((void *)data, n, MADV_SEQUENTIAL);
    } else {
        /* st_size 0 may still be a pipe */
        size_t cap = 1 << 20, len = 0; uint8_t *buf = malloc(cap); ssize_t r;
        while ((r = read(fd, buf + len, cap - len)) > 0) { len += (size_t)r; if (len == cap) { cap *= 2; buf = realloc(buf, cap); } }
        data = buf; n = len;
    }
    if (n && memchr(data, 0, n)) {
        fprintf(stderr, "tncCountsProfile: NUL byte in input is not supported (see DESIGN.md)\n");
        return 2;
    }

    int ngpu = getenv("SSB_GPUS") ? atoi(getenv("SSB_GPUS")) : 1;
    int dev0 = getenv("SSB_DEVICE") ? atoi(getenv("SSB_DEVICE")) : 0;
    if (ngpu < 1) ngpu = 1;
    if (ngpu > 64) ngpu = 64;
    if ((size_t)ngpu > n / 4096 + 1) ngpu = (int)(n / 4096 + 1);

    shard_t sh[64]; ssb_ctx *ctxs[64];
    memset(sh, 0, sizeof sh);
    for (int g = 0; g < ngpu; g++) {
        int rc = ssb_ctx_create(dev0 + g, &ctxs[g]);
        if (rc) { fprintf(stderr, "tncCountsProfile: device %d: %s\n", dev0 + g, ssb_strerror(rc)); return 3; }
        sh[g].ctx = ctxs[g]; sh[g].base = data;
        sh[g].off = (n / (size_t)ngpu * (size_t)g) & ~(size_t)31;
    }
    for (int g = 0; g < ngpu; g++) {
        sh[g].len = (g + 1 < ngpu ? sh[g + 1].off : n) - sh[g].off;
        /* scanner state at the shard start, from the bytes before it (no counting) */
        if (g) ssb_tnc_carry_after(data, sh[g].off, NULL, &sh[g].carry_in);
    }
    pthread_t th[64];
    for (int g = 1; g < ngpu; g++) pthread_create(&th[g], NULL, shard_main, &sh[g]);
    shard_main(&sh[0]);
    for (int g = 1; g < ngpu; g++) pthread_join(th[g], NULL);
    for (int g = 0; g < ngpu; g++)
        if (sh[g].rc) { fprintf(stderr, "tncCountsProfile: %s (%s)\n", ssb_strerror(sh[g].rc), ssb_last_error(ctxs[g])); return 3; }

    int64_t total[64];
    if (ngpu == 1) memcpy(total, sh[0].counts, sizeof total);
    else {
        /* the path's only collective: all-reduce the 64 counters over NVLink */
        void *comms[64]; int64_t *dc[64];
        int rc = ssb_nccl_init_all(ctxs, ngpu, comms);
        if (rc) { fprintf(stderr, "tncCountsProfile: %s (%s)\n", ssb_strerror(rc), ssb_last_error(ctxs[0])); return 3; }
        for (int g = 0; g < ngpu; g++) {
            if (ssb_dev_alloc(ctxs[g], sizeof total, (void **)&dc[g]) || ssb_memcpy_h2d(ctxs[g], dc[g], sh[g].counts, sizeof total)) return 3;
        }
        rc = ssb_tnc_allreduce_group(ctxs, comms, dc, ngpu);
        if (rc) { fprintf(stderr, "tncCountsProfile: %s (%s)\n", ssb_strerror(rc), ssb_last_error(ctxs[0])); return 3; }
        if (ssb_memcpy_d2h(ctxs[0], total, dc[0], sizeof total) || ssb_sync(ctxs[0])) return 3;
        for (int g = 0; g < ngpu; g++) { ssb_sync(ctxs[g]); ssb_dev_free(ctxs[g], dc[g]); }
        ssb_nccl_destroy_all(comms, ngpu);
    }
    char out[4096];
    int w = ssb_tnc_format(total, out, sizeof out);           /* tncCountsProfile.c:452-483 */
    if (w < 0) return 3;
    fwrite(out, 1, (size_t)w, stdout);
    for (int g = 0; g < ngpu; g++) ssb_ctx_destroy(ctxs[g]);
    return 0;
}
