// exchange.cu -- the two transports behind ssb_exchange, and the NCCL plumbing of the library.
//
//   local   n shards driven by n threads of one process (the stochasticSpike main with SSB_GPUS=N, or several logical
//           shards on one device): a mutex, a condition variable, a few mailboxes.
//   nccl    one shard per process / GPU over an NCCL communicator (bench.py under torchrun): all-gather, all-reduce(max),
//           send / recv over NVLink.  Payloads are staged through a small device buffer of the context.
//
// NCCL is resolved at run time (dlsym on the already-loaded image first, so a process that has torch's bundled NCCL loaded
// uses that one; otherwise libnccl.so.2 from the system): libssb200.so has no link-time NCCL dependency and single-GPU
// users never load it.  The TNC path's only collective (one all-reduce of 64 int64) lives here too.
#include "exchange.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <condition_variable>
#include <mutex>
#include <vector>

namespace {
struct NcclApi {
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
    bool ok;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = RTLD_DEFAULT;
        if (!dlsym(h, "ncclAllReduce")) {
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!h) { api.ok = false; return; }
        }
        *(void **)&api.AllReduce = dlsym(h, "ncclAllReduce");
        *(void **)&api.AllGather = dlsym(h, "ncclAllGather");
        *(void **)&api.Send = dlsym(h, "ncclSend");
        *(void **)&api.Recv = dlsym(h, "ncclRecv");
        *(void **)&api.CommInitAll = dlsym(h, "ncclCommInitAll");
        *(void **)&api.CommInitRank = dlsym(h, "ncclCommInitRank");
        *(void **)&api.GetUniqueId = dlsym(h, "ncclGetUniqueId");
        *(void **)&api.CommDestroy = dlsym(h, "ncclCommDestroy");
        *(void **)&api.GroupStart = dlsym(h, "ncclGroupStart");
        *(void **)&api.GroupEnd = dlsym(h, "ncclGroupEnd");
        *(void **)&api.GetErrorString = dlsym(h, "ncclGetErrorString");
        api.ok = api.AllReduce && api.AllGather && api.Send && api.Recv && api.CommInitAll && api.CommInitRank && api.GetUniqueId &&
                 api.CommDestroy && api.GroupStart && api.GroupEnd && api.GetErrorString;
    });
    return api.ok ? &api : NULL;
}

int nccl_fail(ssb_ctx *ctx, NcclApi *a, ncclResult_t r, const char *what)
{
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s: %s", what, a ? a->GetErrorString(r) : "NCCL not available");
    return SSB_E_NCCL;
}

#define SSB_NCCL(ctx, a, call, what) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return nccl_fail((ctx), (a), r_, (what)); } while (0)

// ---------------------------------------------------------------------------------------------------------------- local
struct LocalHub {
    int n;
    std::mutex mu;
    std::condition_variable cv;
    int refs;
    bool failed = false;           // a shard gave up: every waiting call returns SSB_E_PEER
    // all-gather / all-reduce: one round at a time
    std::vector<uint8_t> gather; size_t gather_bytes = 0; int arrived = 0, left = 0; unsigned long long round = 0;
    std::vector<long long> red; size_t red_count = 0;
    // mailboxes: box[i] holds messages for shard i from shard i - 1, in order
    std::vector<std::vector<std::vector<uint8_t>>> box;
    explicit LocalHub(int n_) : n(n_), refs(n_), box((size_t)n_) {}
};

struct LocalExchange : ssb_exchange {
    LocalHub *hub;
    ~LocalExchange() override
    {
        bool last;
        { std::lock_guard<std::mutex> g(hub->mu); last = --hub->refs == 0; }
        if (last) delete hub;
    }
    // generic barrier round: `fill` runs under the lock when a shard arrives, `take` when everybody has
    template <typename F1, typename F2> int round_trip(F1 fill, F2 take)
    {
        std::unique_lock<std::mutex> g(hub->mu);
        hub->cv.wait(g, [&] { return hub->left == 0 || hub->failed; });          // the previous round has been read by everyone
        if (hub->failed) return SSB_E_PEER;
        const unsigned long long my = hub->round;
        fill();
        if (++hub->arrived == hub->n) { hub->arrived = 0; hub->left = hub->n; hub->round++; hub->cv.notify_all(); }
        else hub->cv.wait(g, [&] { return hub->round != my || hub->failed; });
        if (hub->round == my) return SSB_E_PEER;
        take();
        if (--hub->left == 0) hub->cv.notify_all();
        return SSB_OK;
    }
    void abort_group() override
    {
        std::lock_guard<std::mutex> g(hub->mu);
        hub->failed = true;
        hub->cv.notify_all();
    }
    int allgather(const void *send, void *recv, size_t bytes) override
    {
        return round_trip([&] { if (hub->gather.size() < bytes * (size_t)n) hub->gather.resize(bytes * (size_t)n);
                                memcpy(hub->gather.data() + bytes * (size_t)rank, send, bytes); },
                          [&] { memcpy(recv, hub->gather.data(), bytes * (size_t)n); });
    }
    int allreduce_max_i64(long long *d_buf, size_t count, cudaStream_t s) override
    {
        std::vector<long long> mine(count);
        if (cudaMemcpyAsync(mine.data(), d_buf, count * sizeof(long long), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) return SSB_E_CUDA;
        const int rc = round_trip([&] { if (hub->arrived == 0) hub->red.assign(mine.begin(), mine.end());
                                        else for (size_t i = 0; i < count; i++) if (mine[i] > hub->red[i]) hub->red[i] = mine[i]; },
                                  [&] { memcpy(mine.data(), hub->red.data(), count * sizeof(long long)); });
        if (rc) return rc;
        if (cudaMemcpyAsync(d_buf, mine.data(), count * sizeof(long long), cudaMemcpyHostToDevice, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) return SSB_E_CUDA;
        return SSB_OK;
    }
    int send_next(const void *buf, size_t bytes) override
    {
        if (rank + 1 >= n) return SSB_OK;
        std::lock_guard<std::mutex> g(hub->mu);
        hub->box[(size_t)rank + 1].emplace_back((const uint8_t *)buf, (const uint8_t *)buf + bytes);
        hub->cv.notify_all();
        return SSB_OK;
    }
    int recv_prev(void *buf, size_t bytes) override
    {
        if (rank == 0) return SSB_E_ARG;
        std::unique_lock<std::mutex> g(hub->mu);
        auto &b = hub->box[(size_t)rank];
        hub->cv.wait(g, [&] { return !b.empty() || hub->failed; });
        if (b.empty()) return SSB_E_PEER;
        if (b.front().size() != bytes) return SSB_E_STATE;
        memcpy(buf, b.front().data(), bytes);
        b.erase(b.begin());
        return SSB_OK;
    }
    int shift(const void *send, size_t send_bytes, void *recv, size_t recv_bytes) override
    {
        int rc = SSB_OK;
        if (rank + 1 < n && send_bytes) rc = send_next(send, send_bytes);
        if (!rc && rank > 0 && recv_bytes) rc = recv_prev(recv, recv_bytes);
        return rc;
    }
};

// ----------------------------------------------------------------------------------------------------------------- nccl
struct NcclExchange : ssb_exchange {
    ssb_ctx *ctx; ncclComm_t comm; NcclApi *a;
    uint8_t *d_stage = NULL; size_t stage_bytes = 0;
    ~NcclExchange() override { if (d_stage) { cudaSetDevice(ctx->device); cudaFree(d_stage); } }
    int reserve(size_t bytes)
    {
        if (bytes <= stage_bytes) return SSB_OK;
        SSB_CUDA(ctx, cudaSetDevice(ctx->device));
        if (d_stage) { SSB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(d_stage); d_stage = NULL; stage_bytes = 0; }
        const size_t want = bytes + (bytes >> 2) + 4096;
        SSB_CUDA(ctx, cudaMalloc(&d_stage, want));
        stage_bytes = want;
        return SSB_OK;
    }
    int allgather(const void *send, void *recv, size_t bytes) override
    {
        int rc = reserve(bytes * (size_t)(n + 1)); if (rc) return rc;
        cudaStream_t s = ctx->stream;
        SSB_CUDA(ctx, cudaMemcpyAsync(d_stage, send, bytes, cudaMemcpyHostToDevice, s));
        SSB_NCCL(ctx, a, a->AllGather(d_stage, d_stage + bytes, bytes, ncclUint8, comm, s), "ncclAllGather");
        SSB_CUDA(ctx, cudaMemcpyAsync(recv, d_stage + bytes, bytes * (size_t)n, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        return SSB_OK;
    }
    int allreduce_max_i64(long long *d_buf, size_t count, cudaStream_t s) override
    {
        SSB_NCCL(ctx, a, a->AllReduce(d_buf, d_buf, count, ncclInt64, ncclMax, comm, s), "ncclAllReduce");
        return SSB_OK;
    }
    int send_next(const void *buf, size_t bytes) override
    {
        if (rank + 1 >= n) return SSB_OK;
        int rc = reserve(bytes); if (rc) return rc;
        cudaStream_t s = ctx->copy_stream;               // the hand-off must not queue behind kernels of the compute stream
        SSB_CUDA(ctx, cudaMemcpyAsync(d_stage, buf, bytes, cudaMemcpyHostToDevice, s));
        SSB_NCCL(ctx, a, a->Send(d_stage, bytes, ncclUint8, rank + 1, comm, s), "ncclSend");
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        return SSB_OK;
    }
    int recv_prev(void *buf, size_t bytes) override
    {
        if (rank == 0) return SSB_E_ARG;
        int rc = reserve(bytes); if (rc) return rc;
        cudaStream_t s = ctx->copy_stream;
        SSB_NCCL(ctx, a, a->Recv(d_stage, bytes, ncclUint8, rank - 1, comm, s), "ncclRecv");
        SSB_CUDA(ctx, cudaMemcpyAsync(buf, d_stage, bytes, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        return SSB_OK;
    }
    int shift(const void *send, size_t send_bytes, void *recv, size_t recv_bytes) override
    {
        const bool do_send = rank + 1 < n && send_bytes, do_recv = rank > 0 && recv_bytes;
        if (!do_send && !do_recv) return SSB_OK;
        int rc = reserve(send_bytes + recv_bytes + 256); if (rc) return rc;
        cudaStream_t s = ctx->copy_stream;
        uint8_t *d_send = d_stage, *d_recv = d_stage + ((send_bytes + 255) & ~(size_t)255);
        if (do_send) SSB_CUDA(ctx, cudaMemcpyAsync(d_send, send, send_bytes, cudaMemcpyHostToDevice, s));
        SSB_NCCL(ctx, a, a->GroupStart(), "ncclGroupStart");
        ncclResult_t r1 = do_send ? a->Send(d_send, send_bytes, ncclUint8, rank + 1, comm, s) : ncclSuccess;
        ncclResult_t r2 = do_recv ? a->Recv(d_recv, recv_bytes, ncclUint8, rank - 1, comm, s) : ncclSuccess;
        SSB_NCCL(ctx, a, a->GroupEnd(), "ncclGroupEnd");
        if (r1 != ncclSuccess) return nccl_fail(ctx, a, r1, "ncclSend");
        if (r2 != ncclSuccess) return nccl_fail(ctx, a, r2, "ncclRecv");
        if (do_recv) SSB_CUDA(ctx, cudaMemcpyAsync(recv, d_recv, recv_bytes, cudaMemcpyDeviceToHost, s));
        SSB_CUDA(ctx, cudaStreamSynchronize(s));
        return SSB_OK;
    }
};
} // namespace

extern "C" int ssb_exchange_local_create(int n_shards, ssb_exchange **out)
{
    if (n_shards < 1 || !out) return SSB_E_ARG;
    LocalHub *hub = new LocalHub(n_shards);
    for (int i = 0; i < n_shards; i++) { LocalExchange *x = new LocalExchange(); x->rank = i; x->n = n_shards; x->hub = hub; out[i] = x; }
    return SSB_OK;
}

extern "C" int ssb_exchange_nccl_create(ssb_ctx *ctx, void *nccl_comm, int rank, int n_ranks, ssb_exchange **out)
{
    if (!ctx || !nccl_comm || !out || rank < 0 || rank >= n_ranks) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctx, NULL, ncclSuccess, "dlopen(libnccl)");
    NcclExchange *x = new NcclExchange();
    x->rank = rank; x->n = n_ranks; x->ctx = ctx; x->comm = (ncclComm_t)nccl_comm; x->a = a;
    *out = x;
    return SSB_OK;
}

extern "C" void ssb_exchange_destroy(ssb_exchange *xc) { delete xc; }

extern "C" int ssb_nccl_unique_id(uint8_t id128[128])
{
    if (!id128) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return SSB_E_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) return SSB_E_NCCL;
    memcpy(id128, &id, 128);
    return SSB_OK;
}

extern "C" int ssb_nccl_comm_init_rank(ssb_ctx *ctx, int n_ranks, int rank, const uint8_t id128[128], void **comm_out)
{
    if (!ctx || !id128 || !comm_out) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctx, NULL, ncclSuccess, "dlopen(libnccl)");
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclComm_t comm;
    SSB_NCCL(ctx, a, a->CommInitRank(&comm, n_ranks, id, rank), "ncclCommInitRank");
    *comm_out = comm;
    return SSB_OK;
}

extern "C" void ssb_nccl_comm_destroy(void *nccl_comm)
{
    NcclApi *a = nccl_api();
    if (a && nccl_comm) a->CommDestroy((ncclComm_t)nccl_comm);
}

// ---- TNC: the path's only collective ------------------------------------------------------------------------------
extern "C" int ssb_tnc_allreduce(ssb_ctx *ctx, void *nccl_comm, int64_t *d_counts64)
{
    if (!ctx || !nccl_comm || !d_counts64) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctx, NULL, ncclSuccess, "dlopen(libnccl)");
    SSB_CUDA(ctx, cudaSetDevice(ctx->device));
    SSB_NCCL(ctx, a, a->AllReduce(d_counts64, d_counts64, 64, ncclInt64, ncclSum, (ncclComm_t)nccl_comm, ctx->stream), "ncclAllReduce");
    ctx->launches++;
    return SSB_OK;
}

// Single-process, several GPUs (what the C mains do): one communicator per context.
extern "C" int ssb_nccl_init_all(ssb_ctx **ctxs, int n, void **comms_out)
{
    if (!ctxs || n <= 0 || !comms_out) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctxs[0], NULL, ncclSuccess, "dlopen(libnccl)");
    int devs[64];
    if (n > 64) return SSB_E_ARG;
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    SSB_NCCL(ctxs[0], a, a->CommInitAll((ncclComm_t *)comms_out, n, devs), "ncclCommInitAll");
    return SSB_OK;
}

extern "C" int ssb_tnc_allreduce_group(ssb_ctx **ctxs, void **comms, int64_t **d_counts64, int n)
{
    if (!ctxs || !comms || !d_counts64 || n <= 0) return SSB_E_ARG;
    NcclApi *a = nccl_api();
    if (!a) return nccl_fail(ctxs[0], NULL, ncclSuccess, "dlopen(libnccl)");
    SSB_NCCL(ctxs[0], a, a->GroupStart(), "ncclGroupStart");
    for (int i = 0; i < n; i++) {
        cudaSetDevice(ctxs[i]->device);
        ncclResult_t r = a->AllReduce(d_counts64[i], d_counts64[i], 64, ncclInt64, ncclSum, (ncclComm_t)comms[i], ctxs[i]->stream);
        if (r != ncclSuccess) { a->GroupEnd(); return nccl_fail(ctxs[0], a, r, "ncclAllReduce"); }
    }
    SSB_NCCL(ctxs[0], a, a->GroupEnd(), "ncclGroupEnd");
    return SSB_OK;
}

extern "C" void ssb_nccl_destroy_all(void **comms, int n)
{
    NcclApi *a = nccl_api();
    if (!a || !comms) return;
    for (int i = 0; i < n; i++) if (comms[i]) a->CommDestroy((ncclComm_t)comms[i]);
}

// ---- test hooks (host only; tests/test_spike_shards.py): the transport without a GPU -----------------------------------
extern "C" int ssb_exchange_test_allgather(ssb_exchange *xc, const void *send, void *recv, size_t bytes)
{
    return xc ? xc->allgather(send, recv, bytes) : SSB_E_ARG;
}
extern "C" int ssb_exchange_test_relay(ssb_exchange *xc, uint64_t *v)
{
    if (!xc || !v) return SSB_E_ARG;
    int rc = SSB_OK;
    if (xc->rank > 0) { rc = xc->recv_prev(v, sizeof *v); *v += 1; }
    if (!rc) rc = xc->send_next(v, sizeof *v);
    return rc;
}
