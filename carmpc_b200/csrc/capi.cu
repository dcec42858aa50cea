// Error reporting, handle destruction and the device micro-benchmarks of the C ABI.
#include <stdarg.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace carmpc {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

namespace {
struct KernelConfigEntry {
    const void* fn; int device; int threads; size_t smem; int per_sm;
};
std::mutex g_kc_mutex;
std::vector<KernelConfigEntry> g_kc_entries;
struct KernelAttrEntry { const void* fn; int device; size_t max_smem; };
std::vector<KernelAttrEntry> g_ka_entries;
}  // namespace

int kernel_config(const void* fn, int threads, size_t smem, int* per_sm) {
    int dev = 0;
    CARMPC_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_kc_mutex);
    bool attr_ok = false;
    for (KernelAttrEntry& a : g_ka_entries)
        if (a.fn == fn && a.device == dev) {
            if (a.max_smem < smem) {
                CARMPC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                a.max_smem = smem;
            }
            attr_ok = true;
            break;
        }
    if (!attr_ok) {
        CARMPC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_ka_entries.push_back({fn, dev, smem});
    }
    if (per_sm == nullptr) return CARMPC_OK;
    for (const KernelConfigEntry& e : g_kc_entries)
        if (e.fn == fn && e.device == dev && e.threads == threads && e.smem == smem) { *per_sm = e.per_sm; return CARMPC_OK; }
    int occ = 0;
    CARMPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
    if (occ < 1) occ = 1;
    g_kc_entries.push_back({fn, dev, threads, smem, occ});
    *per_sm = occ;
    return CARMPC_OK;
}

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsFallback;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        return kNumSMsFallback;
    return n;
}

// ---- micro-benchmarks: dependent-chain-free FMA streams, 8 independent accumulators per thread ----------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
    T acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = T(threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = acc[k] * a + b;
        }
    }
    T s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Register-tiled outer-product loop fed from shared memory (8 x 8 accumulators, one broadcast 128-bit and one
// per-lane 128-bit load pair per 64 FFMA): the practical FFMA ceiling of a shared-memory tile product, where both
// multiplicands are vector registers (unlike fma_peak_kernel, whose multiplier is a uniform register).
__global__ void __launch_bounds__(256) tile_peak_kernel(float* out, int iters) {
    __shared__ __align__(16) float sa[64 * 8];
    __shared__ __align__(16) float sb[4 * 2048];
    for (int i = threadIdx.x; i < 64 * 8; i += 256) sa[i] = 1.0f + 1e-6f * i;
    for (int i = threadIdx.x; i < 4 * 2048; i += 256) sb[i] = 1.0f - 1e-6f * (i & 1023);
    __syncthreads();
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int lane8 = threadIdx.x * 8;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(sa + k * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(sa + k * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(sb + (((k & 3) * 2048 + lane8) & (4 * 2048 - 1)));
            const float4 b1 = *reinterpret_cast<const float4*>(sb + (((k & 3) * 2048 + lane8 + 4) & (4 * 2048 - 1)));
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int measure_tile(double* value) {
    const int blocks = sm_count() * 2, threads = 256, iters = 256;
    float* out = nullptr;
    CARMPC_CUDA(cudaMalloc(&out, sizeof(float) * blocks * threads));
    CARMPC_CUDA(cudaFuncSetAttribute(tile_peak_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cudaEvent_t e0, e1;
    CARMPC_CUDA(cudaEventCreate(&e0));
    CARMPC_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CARMPC_CUDA(cudaEventRecord(e0));
        tile_peak_kernel<<<blocks, threads>>>(out, iters);
        CARMPC_CUDA(cudaEventRecord(e1));
        CARMPC_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CARMPC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *value = best;
    return CARMPC_OK;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = src[i];
}

template <typename T>
static int measure_fma(double* value) {
    const int blocks = sm_count() * 8, threads = 256, iters = 4096;
    T* out = nullptr;
    CARMPC_CUDA(cudaMalloc(&out, sizeof(T) * blocks * threads));
    cudaEvent_t e0, e1;
    CARMPC_CUDA(cudaEventCreate(&e0));
    CARMPC_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CARMPC_CUDA(cudaEventRecord(e0));
        fma_peak_kernel<T><<<blocks, threads>>>(out, iters, T(0.999), T(1e-3));
        CARMPC_CUDA(cudaEventRecord(e1));
        CARMPC_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CARMPC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *value = best;
    return CARMPC_OK;
}

static int measure_copy(double* value) {
    const size_t bytes = (size_t)2 << 30;          // 2 GiB each way, far larger than the 126 MB L2
    double2 *src = nullptr, *dst = nullptr;
    CARMPC_CUDA(cudaMalloc(&src, bytes));
    CARMPC_CUDA(cudaMalloc(&dst, bytes));
    CARMPC_CUDA(cudaMemset(src, 1, bytes));
    cudaEvent_t e0, e1;
    CARMPC_CUDA(cudaEventCreate(&e0));
    CARMPC_CUDA(cudaEventCreate(&e1));
    double best = 0;
    const size_t n = bytes / sizeof(double2);
    for (int rep = 0; rep < 6; ++rep) {
        CARMPC_CUDA(cudaEventRecord(e0));
        copy_kernel<<<sm_count() * 16, 256>>>(src, dst, n);
        CARMPC_CUDA(cudaEventRecord(e1));
        CARMPC_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CARMPC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double gbs = 2.0 * bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best) best = gbs;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(src);
    cudaFree(dst);
    *value = best;
    return CARMPC_OK;
}

}  // namespace carmpc

extern "C" {

const char* carmpc_last_error(void) { return carmpc::g_error; }

const char* carmpc_version(void) { return "carmpc-b200 0.1 (sm_100a)"; }

void carmpc_destroy(void* handle) {
    if (handle == nullptr) return;
    delete static_cast<carmpc::HandleBase*>(handle);
}

int carmpc_measure_peak(int which, double* h_value) {
    CARMPC_REQUIRE(h_value != nullptr, "h_value");
    switch (which) {
        case 0: return carmpc::measure_fma<float>(h_value);
        case 1: return carmpc::measure_fma<double>(h_value);
        case 2: return carmpc::measure_copy(h_value);
        case 3: return carmpc::measure_tile(h_value);
        default: carmpc::set_error("carmpc_measure_peak: which must be 0..3"); return CARMPC_ERR_INVALID;
    }
}

}  // extern "C"
