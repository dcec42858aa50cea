// Peer window of a sample set that is sharded over the GPUs of one box: every rank holds the FULL bitset
// (double-buffered), one member count per rank and one step flag per rank; the scan kernels of rank r write r's words
// straight into every rank's window through NVLink peer mappings, and a one-warp exchange kernel publishes r's count
// and flag and waits for everybody else's.  Internal to the library (C ABI: carmpc_shard_*, include/carmpc.h).
#pragma once

#include "common.cuh"

namespace carmpc {

constexpr int kShardMaxWorld = 8;
constexpr size_t kShardHeaderBytes = 512;      // [flags 8 x u64 | pad][counts 2 x 8 x i64][error word | pad]

struct ShardWindow : HandleBase {
    int rank = 0, world = 1;
    int64_t n_total = 0;         // samples of the whole set
    int64_t words_pad = 0;       // bitset words per buffer (n_total / 32 rounded up to whole 128-byte lines)
    size_t bytes = 0;
    unsigned char* base = nullptr;                       // this rank's window (cudaMalloc, IPC-exportable)
    unsigned char* peer[kShardMaxWorld] = {nullptr};     // every rank's window as mapped here (peer[rank] == base)
    bool ipc_opened[kShardMaxWorld] = {false};
    bool connected = false;
    bool pending = false;                                // a step was published and not yet waited for
    unsigned long long step = 0;                         // collective steps enqueued so far (host mirror of d_local_count[3])
    unsigned long long* d_local_count = nullptr;         // [0] member count of this rank's shard in the current step,
                                                         // [1..2] work counters of the scan kernel (self-resetting),
                                                         // [3] completed collective steps (advanced by the exchange kernel)
    ~ShardWindow() override;

    unsigned long long* flags(int r) const { return reinterpret_cast<unsigned long long*>(peer[r]); }
    long long* counts(int r) const { return reinterpret_cast<long long*>(peer[r] + 128); }
    int* error_word() const { return reinterpret_cast<int*>(base + 384); }
    uint32_t* bits(int r, int slot) const {
        return reinterpret_cast<uint32_t*>(peer[r] + kShardHeaderBytes) + (size_t)slot * words_pad;
    }
};

// publish: this rank's count + flag of the next step into every rank's window (never blocks)
int shard_publish_launch(ShardWindow* W, cudaStream_t st);
// wait: for every rank's flag; sum the counts, advance the device-side step, re-arm the local count
int shard_wait_launch(ShardWindow* W, int64_t* d_total, cudaStream_t st);

}  // namespace carmpc
