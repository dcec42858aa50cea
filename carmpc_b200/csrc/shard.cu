// Peer windows for sharded evaluation (BASELINE config 2: "a 10^8-point grid sharded across 1/2/4/8 B200").
//
// north_star gathers the membership bitsets of the ranks over NVLink.  Instead of a separate collective after the scan,
// the scan kernel of every rank stores its bitset words directly into all ranks' windows (peer-mapped device memory:
// CUDA IPC between the one-process-per-GPU ranks, plain pointers inside one process), 128 bytes per warp store; what is
// left of the "all-gather" is one flag per rank.  The two one-warp kernels below are that flag protocol (publish never
// blocks; wait does):
//     counts[slot][rank] <- my member count      (relaxed system-scope store into every window)
//     flags[rank]        <- step                 (release, system scope, after a system fence)
//     wait until flags[q] >= step for every q    (acquire, system scope, in my own window)
//     total = sum_q counts[slot][q]
// Buffers alternate between two slots; rank r can only start writing slot s of step k + 2 after its exchange of step
// k + 1 has seen every peer's flag k + 1, i.e. after every peer has finished (in stream order) whatever it enqueued
// between its exchanges of steps k and k + 1 - the consumers of slot s.
#include <string.h>

#include "shard.cuh"

namespace carmpc {

ShardWindow::~ShardWindow() {
    for (int r = 0; r < kShardMaxWorld; ++r)
        if (ipc_opened[r] && peer[r] != nullptr) cudaIpcCloseMemHandle(peer[r]);
    cudaFree(base);
    cudaFree(d_local_count);
}

namespace {

struct ExchangeArgs {
    unsigned long long* peer_flags[kShardMaxWorld];
    long long* peer_counts[kShardMaxWorld];
    const unsigned long long* local_flags;
    const long long* local_counts;
    unsigned long long* local_count;       // [0] count (re-armed here), [3] completed steps (advanced here)
    int* error_word;
    long long* total;
    int rank, world;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// publish: my member count, then my flag, into every rank's window (never blocks)
__global__ void __launch_bounds__(32) shard_publish_kernel(const ExchangeArgs a) {
    const int lane = threadIdx.x;
    const unsigned long long step = a.local_count[3] + 1ull;
    const int slot = (int)(step & 1ull);
    if (lane < a.world) {
        const long long mine = (long long)a.local_count[0];
        asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(a.peer_counts[lane] + slot * kShardMaxWorld + a.rank), "l"(mine) : "memory");
        __threadfence_system();
        st_release_sys(a.peer_flags[lane] + a.rank, step);
    }
}

// wait: every rank's flag arrives in this rank's own window; then sum the counts, advance the step, re-arm the count
__global__ void __launch_bounds__(32) shard_wait_kernel(const ExchangeArgs a) {
    const int lane = threadIdx.x;
    const unsigned long long step = a.local_count[3] + 1ull;
    const int slot = (int)(step & 1ull);
    if (lane < a.world) {
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(a.local_flags + lane) < step) {
            if (global_ns() - t0 > 5000000000ull) {          // 5 s: a peer died; fail loudly instead of hanging the GPU
                atomicExch(a.error_word, 1 + lane);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncwarp();
    long long c = 0;
    if (lane < a.world)
        asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(c) : "l"(a.local_counts + slot * kShardMaxWorld + lane) : "memory");
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
        if (a.total != nullptr) *a.total = c;
        a.local_count[0] = 0ull;               // re-armed for the next step's scan
        a.local_count[3] = step;
    }
}

}  // namespace

static ExchangeArgs exchange_args(ShardWindow* W, int64_t* d_total) {
    ExchangeArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < W->world; ++r) {
        a.peer_flags[r] = W->flags(r);
        a.peer_counts[r] = W->counts(r);
    }
    a.local_flags = W->flags(W->rank);
    a.local_counts = W->counts(W->rank);
    a.local_count = W->d_local_count;
    a.error_word = W->error_word();
    a.total = reinterpret_cast<long long*>(d_total);
    a.rank = W->rank; a.world = W->world;
    return a;
}

int shard_publish_launch(ShardWindow* W, cudaStream_t st) {
    shard_publish_kernel<<<1, 32, 0, st>>>(exchange_args(W, nullptr));
    CARMPC_CUDA(cudaGetLastError());
    W->pending = true;
    return CARMPC_OK;
}

int shard_wait_launch(ShardWindow* W, int64_t* d_total, cudaStream_t st) {
    shard_wait_kernel<<<1, 32, 0, st>>>(exchange_args(W, d_total));
    CARMPC_CUDA(cudaGetLastError());
    W->pending = false;
    return CARMPC_OK;
}

}  // namespace carmpc

using namespace carmpc;

extern "C" {

int carmpc_shard_create(int rank, int world, int64_t n_total, void** handle) {
    CARMPC_REQUIRE(handle != nullptr, "handle");
    CARMPC_REQUIRE(world >= 1 && world <= kShardMaxWorld, "world must be in [1, 8]");
    CARMPC_REQUIRE(rank >= 0 && rank < world, "rank");
    CARMPC_REQUIRE(n_total >= 0, "n_total");
    ShardWindow* W = new ShardWindow();
    W->kind = kShard;
    cudaGetDevice(&W->device);
    W->rank = rank; W->world = world; W->n_total = n_total;
    const int64_t words = (n_total + 31) / 32;
    W->words_pad = (words + 31) / 32 * 32 + 32;
    W->bytes = kShardHeaderBytes + sizeof(uint32_t) * 2 * (size_t)W->words_pad;
    if (cudaMalloc(&W->base, W->bytes) != cudaSuccess || cudaMalloc(&W->d_local_count, sizeof(unsigned long long) * 4) != cudaSuccess ||
        cudaMemset(W->d_local_count, 0, sizeof(unsigned long long) * 4) != cudaSuccess ||
        cudaMemset(W->base, 0, W->bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        set_error("carmpc_shard_create: %s", cudaGetErrorString(cudaGetLastError()));
        delete W;
        return CARMPC_ERR_CUDA;
    }
    W->peer[rank] = W->base;
    W->connected = world == 1;
    *handle = W;
    return CARMPC_OK;
}

int carmpc_shard_export(void* shard, unsigned char* h_ipc_handle64) {
    ShardWindow* W = check_handle<ShardWindow>(shard, kShard);
    CARMPC_REQUIRE(W != nullptr, "not a shard window");
    CARMPC_REQUIRE(h_ipc_handle64 != nullptr, "h_ipc_handle64");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    CARMPC_CUDA(cudaIpcGetMemHandle(&h, W->base));
    memcpy(h_ipc_handle64, &h, 64);
    return CARMPC_OK;
}

int carmpc_shard_connect(void* shard, const unsigned char* h_ipc_handles) {
    ShardWindow* W = check_handle<ShardWindow>(shard, kShard);
    CARMPC_REQUIRE(W != nullptr, "not a shard window");
    CARMPC_REQUIRE(W->world == 1 || h_ipc_handles != nullptr, "h_ipc_handles");
    for (int r = 0; r < W->world; ++r) {
        if (r == W->rank || W->peer[r] != nullptr) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, h_ipc_handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        CARMPC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        W->peer[r] = static_cast<unsigned char*>(p);
        W->ipc_opened[r] = true;
    }
    W->connected = true;
    return CARMPC_OK;
}

int carmpc_shard_connect_local(void* shard, void* const* peer_shards) {
    ShardWindow* W = check_handle<ShardWindow>(shard, kShard);
    CARMPC_REQUIRE(W != nullptr, "not a shard window");
    CARMPC_REQUIRE(W->world == 1 || peer_shards != nullptr, "peer_shards");
    for (int r = 0; r < W->world; ++r) {
        if (r == W->rank) continue;
        ShardWindow* Q = check_handle<ShardWindow>(peer_shards[r], kShard);
        CARMPC_REQUIRE(Q != nullptr && Q->rank == r && Q->world == W->world && Q->n_total == W->n_total,
                       "peer_shards[r] must be the window of rank r of the same sample set");
        if (Q->device != W->device) {
            int can = 0;
            CARMPC_CUDA(cudaDeviceCanAccessPeer(&can, W->device, Q->device));
            CARMPC_REQUIRE(can != 0, "the devices of two ranks have no peer access");
            int cur = 0;
            CARMPC_CUDA(cudaGetDevice(&cur));
            CARMPC_CUDA(cudaSetDevice(W->device));
            const cudaError_t e = cudaDeviceEnablePeerAccess(Q->device, 0);
            cudaSetDevice(cur);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CARMPC_CUDA(e);
            cudaGetLastError();
        }
        W->peer[r] = Q->base;
    }
    W->connected = true;
    return CARMPC_OK;
}

int carmpc_shard_result(void* shard, const uint32_t** d_bits, int64_t* h_steps) {
    ShardWindow* W = check_handle<ShardWindow>(shard, kShard);
    CARMPC_REQUIRE(W != nullptr, "not a shard window");
    // the step counter lives on the device (a captured graph replays steps without passing through this library)
    unsigned long long step = 0;
    CARMPC_CUDA(cudaMemcpy(&step, W->d_local_count + 3, sizeof(step), cudaMemcpyDeviceToHost));
    if (d_bits) *d_bits = W->bits(W->rank, (int)(step & 1ull));
    if (h_steps) *h_steps = (int64_t)step;
    return CARMPC_OK;
}

int carmpc_shard_wait(void* shard, int64_t* d_total_count, void* stream) {
    ShardWindow* W = check_handle<ShardWindow>(shard, kShard);
    CARMPC_REQUIRE(W != nullptr, "not a shard window");
    CARMPC_REQUIRE(W->pending, "no published step is waiting for its peers (call a *_sharded function with defer_wait first)");
    return shard_wait_launch(W, d_total_count, (cudaStream_t)stream);
}

int carmpc_shard_check(void* shard) {
    ShardWindow* W = check_handle<ShardWindow>(shard, kShard);
    CARMPC_REQUIRE(W != nullptr, "not a shard window");
    int err = 0;
    CARMPC_CUDA(cudaMemcpy(&err, W->error_word(), sizeof(int), cudaMemcpyDeviceToHost));
    if (err != 0) {
        set_error("carmpc_shard: rank %d never published its flag (waited 5 s); the collective step is incomplete", err - 1);
        return CARMPC_ERR_CUDA;
    }
    return CARMPC_OK;
}

}  // extern "C"
