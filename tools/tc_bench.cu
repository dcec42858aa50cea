// Throughput of tcgen05.mma kind::tf32 on B200 as a function of the tile width N (development tool, sm_100a only).
//
// Round 1's tools/tc_selftest.cu timed 6-19 MMAs between issue and mbarrier completion, i.e. the LATENCY of a short
// sequence on operands in the no-swizzle layout, and read a flat 150-300 cycles per MMA off it.  This tool measures
// what the ADMM decision needs: sustained cycles per MMA for long back-to-back sequences on operands in the 128-byte
// swizzled K-major layout (the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B and the tensor core reads without
// bank conflicts), for N = 32 ... 256, one CTA alone and one CTA on every SM, results checked against float64.
//
// Layout of an operand tile [rows][32 tf32] (one 128-byte swizzle span of K per row): groups of 8 rows = 1024 bytes,
// 16-byte chunk c of row r stored at chunk c ^ (r & 7).  Descriptor: start address >> 4, SBO = 1024 (>> 4), version 1,
// layout type 2 (SWIZZLE_128B); a K step of 8 elements advances the start address by 32 bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_bench tools/tc_bench.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO: 8 rows x 128 bytes
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // D f32, A / B tf32, both K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of element (r, k) of a [rows][32] tf32 tile in the SW128 K-major layout
__host__ __device__ inline int sw128_off(int r, int k) {
    return (r >> 3) * 1024 + (r & 7) * 128 + (((k >> 2) ^ (r & 7)) << 4) + (k & 3) * 4;
}

// A: [128][K], B: [N][K] row-major in global memory, K = 32 KB_; D: [128][N] (block 0 writes it).
// The MMA sequence over the KB_ resident k-blocks (4 MMAs each) is issued `rounds` times back to back before one commit.
__global__ void __launch_bounds__(128, 1) tc_bench_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                          int N, int KB_, int rounds, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, K = 32 * KB_;
    unsigned char* sA = smem;                                  // [KB_][128 x 128 B]
    unsigned char* sB = smem + (size_t)KB_ * 128 * 128;        // [KB_][N x 128 B]
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sA + (size_t)(k >> 5) * 128 * 128 + sw128_off(r, k & 31)) = A[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sB + (size_t)(k >> 5) * N * 128 + sw128_off(r, k & 31)) = B[i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int round = 0; round < rounds; ++round) {
            // two accumulator tiles alternate per round (columns [0, N) and [256, 256 + N)): no dependence between rounds
            const uint32_t dcol = tbase + (uint32_t)((round & 1) * 256);
            for (int kb = 0; kb < KB_; ++kb) {
                const uint64_t ad = desc_sw128(smem_u32(sA + (size_t)kb * 128 * 128));
                const uint64_t bd = desc_sw128(smem_u32(sB + (size_t)kb * N * 128));
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_ss(dcol, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (uint32_t)(kb | ks));
            }
        }
        mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) { t1 = clock64(); if (cycles) cycles[blockIdx.x] = t1 - t0; }
    __syncthreads();
    if (blockIdx.x == 0) {
        const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 8) {
            float v[8];
            tmem_ld8(lane_base + c0, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = v[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

// Chunked issue, as the ADMM kernel does it: `per` MMAs, then `ncommit` tcgen05.commit to scratch barriers, repeated `chunks`
// times; drain = 1 additionally waits for each chunk's completion before issuing the next (the latency of one hand-off).
__global__ void __launch_bounds__(128, 1) tc_chunk_kernel(int N, int per, int ncommit, int chunks, int drain, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bars[4];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* sA = smem;
    unsigned char* sB = smem + 4 * 128 * 128;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < (4 * 128 * 128 + 4 * 256 * 128) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t idesc = make_idesc(128, N);
    if (tid == 0) {
        const long long t0 = clock64();
        uint32_t phase = 0;
        for (int c = 0; c < chunks; ++c) {
            const uint64_t ad = desc_sw128(smem_u32(sA + (size_t)(c & 3) * 128 * 128));
            const uint64_t bd = desc_sw128(smem_u32(sB + (size_t)(c & 3) * N * 128));
            for (int i = 0; i < per; ++i) mma_ss(tbase, ad + (uint64_t)((i & 3) * 2), bd + (uint64_t)((i & 3) * 2), idesc, (uint32_t)(c | i));
            for (int k = 0; k < ncommit; ++k) mma_commit(smem_u32(&bars[drain ? 0 : 1 + k]));
            if (drain) { mbar_wait(smem_u32(&bars[0]), phase); phase ^= 1; asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
        }
        mma_commit(smem_u32(&bars[3]));
        mbar_wait(smem_u32(&bars[3]), 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

static void run_chunks(int N, int per, int ncommit, int drain) {
    const int chunks = 256;
    long long* dc; CK(cudaMalloc(&dc, 8));
    const size_t smem = 4 * 128 * 128 + 4 * 256 * 128 + 1024;
    CK(cudaFuncSetAttribute(tc_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_chunk_kernel<<<1, 128, smem>>>(N, per, ncommit, chunks, drain, dc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long c = 0;
    CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    printf("chunked issue N=%3d: %2d MMAs + %d commits per chunk%s: %7.0f cycles per chunk, %6.1f per MMA\n", N, per, ncommit,
           drain ? " + wait for completion" : "", (double)c / chunks, (double)c / chunks / per);
    cudaFree(dc);
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static int run(int N, int KB, int rounds, int blocks, double clock_ghz) {
    const int K = 32 * KB;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(17 + N + K);
    for (auto& a : A) a = trunc_tf32((float)rand() / RAND_MAX * 2.f - 1.f);
    for (auto& b : B) b = trunc_tf32((float)rand() / RAND_MAX * 2.f - 1.f);
    float *dA, *dB, *dD; long long* dc;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dc, 8 * blocks));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    const size_t smem = (size_t)KB * (128 + N) * 128 + 1024;
    CK(cudaFuncSetAttribute(tc_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_bench_kernel<<<blocks, 128, smem>>>(dA, dB, dD, N, KB, rounds, dc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<long long> cyc(blocks);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc.data(), dc, 8 * blocks, cudaMemcpyDeviceToHost));
    // the first accumulator tile holds the product of the last even round: one full pass over K
    double worst = 0, scale = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < N; ++c) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * (double)B[c * K + k];
            worst = fmax(worst, fabs(s - (double)D[r * N + c]));
            scale = fmax(scale, fabs(s));
        }
    long long cmax = 0;
    for (long long c : cyc) cmax = c > cmax ? c : cmax;
    const long long nm = (long long)rounds * KB * 4;
    const double per = (double)cmax / nm, mac = 128.0 * N * 8 / per;
    printf("N=%3d K=%3d CTAs=%3d: %6lld MMAs, %7.1f cycles/MMA, %6.0f MAC/clk/SM, %6.1f TFLOP/s on %d SMs at %.3f GHz   check: max |err| %.2e of %.1f %s\n",
           N, K, blocks, nm, per, mac, 2.0 * mac * clock_ghz * blocks / 1e3, blocks, clock_ghz, worst, scale,
           worst <= 1e-4 * fmax(scale, 1.0) ? "OK" : "MISMATCH");
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
    return worst <= 1e-4 * fmax(scale, 1.0) ? 0 : 1;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz / 1e6;
    int bad = 0;
    const int Ns[] = {16, 32, 64, 96, 128, 160, 192, 256};
    for (int N : Ns) bad += run(N, 4, 64, 1, ghz);                    // 1024 back-to-back MMAs, one CTA alone
    for (int N : Ns) bad += run(N, 4, 64, prop.multiProcessorCount, ghz);
    bad += run(160, 2, 4, 1, ghz);                                      // a short sequence (32 MMAs): latency-dominated, what round 1 timed
    bad += run(48, 4, 1, 1, ghz);
    for (int N : {48, 80, 192}) {
        run_chunks(N, 12, 0, 0); run_chunks(N, 12, 1, 0); run_chunks(N, 12, 2, 0); run_chunks(N, 12, 1, 1); run_chunks(N, 4, 1, 1);
    }
    printf(bad ? "TC BENCH FAILED (%d)\n" : "TC BENCH OK\n", bad);
    return bad != 0;
}
