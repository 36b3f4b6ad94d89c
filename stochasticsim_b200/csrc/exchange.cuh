// exchange.cuh -- how the shards of one cooperative spike run talk to each other (internal interface behind the
// opaque ssb_exchange of include/ssb200.h).  Everything that travels is small: a few scalars per shard, one int64 per
// .spike record, the 8-byte rand() offset and the spiked bases of reads that straddle a cut.
#pragma once
#include "common.cuh"

struct ssb_exchange {
    int rank = 0, n = 1;
    virtual ~ssb_exchange() {}
    // host buffers: every shard contributes `bytes`, recv gets n * bytes in shard order
    virtual int allgather(const void *send, void *recv, size_t bytes) = 0;
    // device buffer of `count` int64, element-wise maximum over the shards, in place; returns with the result visible to stream s
    virtual int allreduce_max_i64(long long *d_buf, size_t count, cudaStream_t s) = 0;
    // the chain of shards: blocking send to shard rank + 1 / receive from shard rank - 1 (host buffers)
    virtual int send_next(const void *buf, size_t bytes) = 0;
    virtual int recv_prev(void *buf, size_t bytes) = 0;
    // every shard sends send_bytes to its successor and receives recv_bytes from its predecessor in one step (the sizes were
    // agreed on before; the first shard receives nothing, the last one sends nothing; 0 bytes = no transfer on that link)
    virtual int shift(const void *send, size_t send_bytes, void *recv, size_t recv_bytes) = 0;
    // this shard failed: wake the others out of any call above (they return SSB_E_PEER).  Transports that cannot do this leave
    // it to the process launcher to end the group.
    virtual void abort_group() {}
};
