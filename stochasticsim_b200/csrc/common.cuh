// common.cuh -- context object and error plumbing shared by every translation unit of libssb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/ssb200.h"

struct ssb_ctx {
    int          device;
    cudaStream_t stream;       // compute stream (all kernels + the timer)
    cudaStream_t copy_stream;  // second stream for double-buffered host<->device copies
    cudaEvent_t  t0, t1;
    cudaEvent_t  ev[4];        // scratch events for stream hand-offs
    int          sm_count;
    size_t       total_mem;
    char         name[128];
    char         err[512];
    uint64_t     launches;
    // lazily grown device scratch, reused across calls
    void        *scratch;
    size_t       scratch_bytes;
    void        *pinned;
    size_t       pinned_bytes;
};

#define SSB_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__,    \
                     #call, cudaGetErrorString(e_));                                         \
            return SSB_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

// Every kernel launch goes through this so ssb_kernel_launches() is an honest count.
#define SSB_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                              \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
        (ctx)->launches++;                                                                   \
        SSB_CUDA(ctx, cudaGetLastError());                                                   \
    } while (0)

int ssb_scratch_reserve(ssb_ctx *ctx, size_t bytes);   // ctx->scratch has at least `bytes`
int ssb_pinned_reserve(ssb_ctx *ctx, size_t bytes);    // ctx->pinned  has at least `bytes`
