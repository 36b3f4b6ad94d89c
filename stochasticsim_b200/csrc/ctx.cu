// ctx.cu -- context lifetime, buffers, copies, timing.  No computation lives here.
#include "common.cuh"
#include <stdlib.h>

struct ssb_prof {
    static const int CAP = 1024;
    cudaEvent_t a[SSB_K_SLOTS][CAP], b[SSB_K_SLOTS][CAP];
    int         made[SSB_K_SLOTS], used[SSB_K_SLOTS];
    double      ms[SSB_K_SLOTS];
    uint64_t    n[SSB_K_SLOTS];
};

extern "C" int ssb_abi_version(void) { return SSB_ABI_VERSION; }

extern "C" const char *ssb_strerror(int code)
{
    switch (code) {
    case SSB_OK:         return "ok";
    case SSB_E_CUDA:     return "CUDA runtime error";
    case SSB_E_NODEVICE: return "no usable sm_100 CUDA device (this library has no CPU fallback)";
    case SSB_E_ARG:      return "invalid argument";
    case SSB_E_NOMEM:    return "out of memory";
    case SSB_E_FORMAT:   return "input violates a documented precondition";
    case SSB_E_UNSORTED: return "SAM input is not coordinate sorted";
    case SSB_E_DEPTH:    return "pileup deeper than MAX_PILEUP_SIZE (10000)";
    case SSB_E_REF:      return "contig missing from the reference FASTA or read past its end";
    case SSB_E_STATE:    return "call order violated";
    case SSB_E_NCCL:     return "NCCL error";
    case SSB_E_PEER:     return "another shard of the cooperative run failed";
    case SSB_E_SHARD:    return "a shard's alignment lines do not match its coordinate range (or its halo is too small)";
    default:             return "unknown error";
    }
}

extern "C" const char *ssb_last_error(const ssb_ctx *ctx) { return ctx ? ctx->err : ""; }

extern "C" int ssb_ctx_create(int device, ssb_ctx **out)
{
    if (!out) return SSB_E_ARG;
    *out = NULL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return SSB_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SSB_E_NODEVICE;
    if (prop.major != 10) return SSB_E_NODEVICE;            // kernels are built for sm_100a only
    ssb_ctx *ctx = (ssb_ctx *)calloc(1, sizeof(ssb_ctx));
    if (!ctx) return SSB_E_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = prop.totalGlobalMem;
    snprintf(ctx->name, sizeof ctx->name, "%.127s", prop.name);
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->t0) != cudaSuccess || cudaEventCreate(&ctx->t1) != cudaSuccess) {
        free(ctx);
        return SSB_E_CUDA;
    }
    for (int i = 0; i < 4; i++)
        if (cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming) != cudaSuccess) { free(ctx); return SSB_E_CUDA; }
    // keep stream-ordered allocations cached between calls (the spike path allocates its work arrays per run)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    *out = ctx;
    return SSB_OK;
}

extern "C" void ssb_ctx_destroy(ssb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->prof) {
        for (int s = 0; s < SSB_K_SLOTS; s++)
            for (int i = 0; i < ctx->prof->made[s]; i++) { cudaEventDestroy(ctx->prof->a[s][i]); cudaEventDestroy(ctx->prof->b[s][i]); }
        free(ctx->prof);
    }
    for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->ev[i]);
    cudaEventDestroy(ctx->t0); cudaEventDestroy(ctx->t1);
    cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->copy_stream);
    free(ctx);
}

extern "C" int ssb_ctx_device_info(const ssb_ctx *ctx, char *name, size_t cap, int *sm_count, size_t *total_mem)
{
    if (!ctx) return SSB_E_ARG;
    if (name && cap) snprintf(name, cap, "%s", ctx->name);
    if (sm_count) *sm_count = ctx->sm_count;
    if (total_mem) *total_mem = ctx->total_mem;
    return SSB_OK;
}

int ssb_scratch_reserve(ssb_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return SSB_OK;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) { cudaFree(ctx->scratch); ctx->scratch = NULL; ctx->scratch_bytes = 0; }
    size_t want = bytes + (bytes >> 3) + (1u << 20);
    SSB_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return SSB_OK;
}

int ssb_pinned_reserve(ssb_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->pinned_bytes) return SSB_OK;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SSB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->pinned) { cudaFreeHost(ctx->pinned); ctx->pinned = NULL; ctx->pinned_bytes = 0; }
    SSB_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return SSB_OK;
}

extern "C" int ssb_host_alloc(ssb_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    SSB_CUDA(ctx, cudaMallocHost(out, bytes ? bytes : 1));
    return SSB_OK;
}
extern "C" void ssb_host_free(ssb_ctx *ctx, void *p) { (void)ctx; if (p) cudaFreeHost(p); }

extern "C" int ssb_dev_alloc(ssb_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    SSB_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return SSB_OK;
}
extern "C" void ssb_dev_free(ssb_ctx *ctx, void *p) { if (ctx) cudaSetDevice(ctx->device); if (p) cudaFree(p); }

extern "C" int ssb_memcpy_h2d(ssb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SSB_OK;
}
extern "C" int ssb_memcpy_d2h(ssb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return SSB_OK;
}
extern "C" int ssb_memset_dev(ssb_ctx *ctx, void *dst, int value, size_t bytes)
{
    if (!ctx) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    return SSB_OK;
}
extern "C" int ssb_sync(ssb_ctx *ctx)
{
    if (!ctx) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SSB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return SSB_OK;
}
extern "C" int ssb_timer_start(ssb_ctx *ctx)
{
    if (!ctx) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    return SSB_OK;
}
extern "C" int ssb_timer_stop(ssb_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
    SSB_CUDA(ctx, cudaEventSynchronize(ctx->t1));
    SSB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return SSB_OK;
}
// ---- per-kernel timing: event pairs around single launches, summed on read -------------------

static void prof_flush(ssb_ctx *ctx, int slot)
{
    ssb_prof *p = ctx->prof;
    for (int i = 0; i < p->used[slot]; i++) {
        float ms = 0;
        if (cudaEventSynchronize(p->b[slot][i]) == cudaSuccess && cudaEventElapsedTime(&ms, p->a[slot][i], p->b[slot][i]) == cudaSuccess) {
            p->ms[slot] += ms; p->n[slot]++;
        }
    }
    p->used[slot] = 0;
}

void ssb_prof_begin(ssb_ctx *ctx, int slot, cudaStream_t s)
{
    ssb_prof *p = ctx->prof;
    if (!p || slot < 0 || slot >= SSB_K_SLOTS) return;
    if (p->used[slot] == ssb_prof::CAP) prof_flush(ctx, slot);
    int i = p->used[slot];
    if (i >= p->made[slot]) { cudaEventCreate(&p->a[slot][i]); cudaEventCreate(&p->b[slot][i]); p->made[slot] = i + 1; }
    cudaEventRecord(p->a[slot][i], s);
}

void ssb_prof_end(ssb_ctx *ctx, int slot, cudaStream_t s)
{
    ssb_prof *p = ctx->prof;
    if (!p || slot < 0 || slot >= SSB_K_SLOTS) return;
    cudaEventRecord(p->b[slot][p->used[slot]++], s);
}

extern "C" int ssb_profile_enable(ssb_ctx *ctx, int on)
{
    if (!ctx) return SSB_E_ARG;
    if (on && !ctx->prof) { ctx->prof = (ssb_prof *)calloc(1, sizeof(ssb_prof)); if (!ctx->prof) return SSB_E_NOMEM; }
    ctx->prof_on = on ? 1 : 0;
    return SSB_OK;
}

extern "C" int ssb_profile_read(ssb_ctx *ctx, int slot, double *total_ms, uint64_t *launches, int reset)
{
    if (!ctx || !ctx->prof || slot < 0 || slot >= SSB_K_SLOTS) return SSB_E_ARG;
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    prof_flush(ctx, slot);
    if (total_ms) *total_ms = ctx->prof->ms[slot];
    if (launches) *launches = ctx->prof->n[slot];
    if (reset) { ctx->prof->ms[slot] = 0; ctx->prof->n[slot] = 0; }
    return SSB_OK;
}

extern "C" uint64_t ssb_kernel_launches(const ssb_ctx *ctx) { return ctx ? ctx->launches : 0; }
