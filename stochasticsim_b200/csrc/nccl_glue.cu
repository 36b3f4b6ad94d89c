// nccl_glue.cu -- the TNC path's only collective: one all-reduce of 64 int64 counters.
//
// NCCL is resolved at run time (dlsym on the already-loaded image first, so a process that has
// torch's bundled NCCL loaded uses that one; otherwise libnccl.so.2 from the system), so that
// libssb200.so has no link-time NCCL dependency and single-GPU users never load it.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : NULL;
    tried = true;
    void *h = RTLD_DEFAULT;
    if (!dlsym(h, "ncclAllReduce")) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return NULL;
    }
    *(void **)&api.AllReduce = dlsym(h, "ncclAllReduce");
    *(void **)&api.CommInitAll = dlsym(h, "ncclCommInitAll");
    *(void **)&api.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void **)&api.GroupStart = dlsym(h, "ncclGroupStart");
    *(void **)&api.GroupEnd = dlsym(h, "ncclGroupEnd");
    *(void **)&api.GetErrorString = dlsym(h, "ncclGetErrorString");
    api.ok = api.AllReduce && api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api.ok ? &api : NULL;
}

int nccl_fail(ssb_ctx *ctx, NcclApi *a, ncclResult_t r, const char *what)
{
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, a ? a->GetErrorString(r) : "NCCL not available");
    return SSB_E_NCCL;
}
} // namespace

extern "C" int ssb_tnc_allreduce(ssb_ctx *ctx, void *nccl_comm, int64_t *d_counts64)
{
    if (!ctx || !nccl_comm || !d_counts64) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctx, NULL, ncclSuccess, "dlopen(libnccl)");
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclResult_t r = a->AllReduce(d_counts64, d_counts64, 64, ncclInt64, ncclSum, (ncclComm_t)nccl_comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, a, r, "ncclAllReduce");
    return SSB_OK;
}

// Single-process, several GPUs (what the tncCountsProfile main does): one communicator per context.
extern "C" int ssb_nccl_init_all(ssb_ctx **ctxs, int n, void **comms_out)
{
    if (!ctxs || n <= 0 || !comms_out) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctxs[0], NULL, ncclSuccess, "dlopen(libnccl)");
    int devs[64];
    if (n > 64) return SSB_E_ARG;
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    ncclResult_t r = a->CommInitAll((ncclComm_t *)comms_out, n, devs);
    if (r != ncclSuccess) return nccl_fail(ctxs[0], a, r, "ncclCommInitAll");
    return SSB_OK;
}

extern "C" int ssb_tnc_allreduce_group(ssb_ctx **ctxs, void **comms, int64_t **d_counts64, int n)
{
    if (!ctxs || !comms || !d_counts64 || n <= 0) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctxs[0], NULL, ncclSuccess, "dlopen(libnccl)");
    ncclResult_t r = a->GroupStart();
    if (r != ncclSuccess) return nccl_fail(ctxs[0], a, r, "ncclGroupStart");
    for (int i = 0; i < n; i++) {
        cudaSetDevice(ctxs[i]->device);
        r = a->AllReduce(d_counts64[i], d_counts64[i], 64, ncclInt64, ncclSum, (ncclComm_t)comms[i], ctxs[i]->stream);
        if (r != ncclSuccess) { a->GroupEnd(); return nccl_fail(ctxs[0], a, r, "ncclAllReduce"); }
    }
    r = a->GroupEnd();
    if (r != ncclSuccess) return nccl_fail(ctxs[0], a, r, "ncclGroupEnd");
    return SSB_OK;
}

extern "C" void ssb_nccl_destroy_all(void **comms, int n)
{
    NcclApi *a = nccl_api();
    if (!a || !comms) return;
    for (int i = 0; i < n; i++) if (comms[i]) a->CommDestroy((ncclComm_t)comms[i]);
}
