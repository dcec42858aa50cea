// Stand-alone check of the tcgen05 building blocks the tensor-core ADMM kernel uses (development tool, sm_100a only):
//   1. kind::tf32 MMA, A and B from shared memory in the no-swizzle K-major canonical layout
//      (element (r, k) at (k / 4) * rows * 16 + r * 16 + (k % 4) * 4 bytes: LBO = rows * 16, SBO = 128);
//   2. the same product with A read from tensor memory (lane = row, column = k, written with tcgen05.st);
//   3. the 3xTF32 split (hi*hi + lo*hi + hi*lo) against a float64 product;
//   4. issue-to-completion time of the MMA sequences of one ADMM iteration (N = 48, K = 152 and N = 112, K = 48).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tc_selftest tools/tc_selftest.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    return d;                                     // base offset 0, layout type 0 = no swizzle
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                    "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// mode 0: A from smem, plain (inputs pre-truncated by the host).  mode 1: A from tmem, plain.
// mode 2: 3xTF32, A hi/lo in tmem, B hi/lo in smem.  mode 3: 3xTF32, A hi/lo in smem.
// A: [128][K] row-major, B: [N][K] row-major, D: [128][N].  reps > 1 repeats the MMA sequence for timing.
__global__ void __launch_bounds__(128, 1) selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                          int N, int K, int mode, int reps, long long* cycles, int nacc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ksteps = K / 8;
    float* sAh = reinterpret_cast<float*>(smem);                        // [K/4][128][4]
    float* sAl = sAh + 128 * K;
    float* sBh = sAl + 128 * K;                                         // [K/4][N][4]
    float* sBl = sBh + N * K;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // stage operands
    const bool split = mode >= 2;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        const float a = A[i], h = split ? tf32_hi(a) : a;
        const int off = (k / 4) * 128 * 4 + r * 4 + (k % 4);
        sAh[off] = h; sAl[off] = tf32_hi(a - h);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        const float b = B[i], h = split ? tf32_hi(b) : b;
        const int off = (k / 4) * N * 4 + r * 4 + (k % 4);
        sBh[off] = h; sBl[off] = tf32_hi(b - h);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    // TMEM columns: D at [0, N), A hi at [128, 128 + K), A lo at [320, 320 + K)   (K <= 184)
    const uint32_t colD = 0, colAh = 128, colAl = 320;
    if (mode == 1 || mode == 2) {
        for (int k0 = 0; k0 < K; k0 += 8) {
            float vh[8], vl[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = A[tid * K + k0 + j];
                vh[j] = split ? tf32_hi(a) : a;
                vl[j] = tf32_hi(a - vh[j]);
            }
            tmem_st8(lane_base + colAh + k0, vh);
            tmem_st8(lane_base + colAl + k0, vl);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = 0, t1 = 0;
    uint32_t parity = 0;
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (rep == 1 || reps == 1) t0 = clock64();
            uint32_t acc = 0;
            const int passes = split ? 3 : 1;
            for (int p = 0; p < passes; ++p) {
                // pass 0: Ah Bh, pass 1: Al Bh, pass 2: Ah Bl.  Descriptors advance by a constant per k-step (start-address field).
                const uint64_t bd0 = make_desc(smem_u32(p == 2 ? sBl : sBh), N * 16, 128), bstep = (uint64_t)((2 * N * 16) >> 4);
                const uint64_t ad0 = make_desc(smem_u32(p == 1 ? sAl : sAh), 128 * 16, 128), astep = (uint64_t)((2 * 128 * 16) >> 4);
                const uint32_t ta = tbase + (p == 1 ? colAl : colAh);
                const bool ts = mode == 1 || mode == 2;
#pragma unroll 4
                for (int j = 0; j < ksteps; ++j) {
                    // nacc > 1 (timing only): round-robin over independent accumulators
                    const uint32_t dcol = tbase + colD + (uint32_t)((j % nacc) * 16);
                    const uint32_t en = nacc > 1 ? (uint32_t)(j >= nacc || p > 0) : acc;
                    if (ts) mma_ts(dcol, ta + j * 8, bd0 + j * bstep, idesc, en);
                    else mma_ss(dcol, ad0 + j * astep, bd0 + j * bstep, idesc, en);
                    acc = 1;
                }
            }
            mma_commit(smem_u32(&bar));
        }
        mbar_wait(smem_u32(&bar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) t1 = clock64();
        __syncthreads();
    }
    if (tid == 0 && cycles) *cycles = reps > 1 ? (t1 - t0) / (reps - 1) : (t1 - t0);
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tmem_ld8(lane_base + colD + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static int run(int N, int K, int mode, int reps, int nacc = 1) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(17 + N + K);
    for (auto& a : A) a = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& b : B) b = (float)rand() / RAND_MAX * 2.f - 1.f;
    if (mode < 2) { for (auto& a : A) a = trunc_tf32(a); for (auto& b : B) b = trunc_tf32(b); }
    float *dA, *dB, *dD; long long* dc;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * 4;
    CK(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, mode, reps, dc, nacc);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc = 0;
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    double worst = 0, scale = 0;
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < N; ++c) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * (double)B[c * K + k];
            worst = fmax(worst, fabs(s - (double)D[r * N + c]));
            scale = fmax(scale, fabs(s));
        }
    if (nacc > 1) {
        const int nm = (mode >= 2 ? 3 : 1) * (K / 8);
        printf("  timing only, %d independent accumulators, mode %d N=%3d K=%3d: %d MMAs in %lld cycles (%.1f / MMA, %.0f MAC/clk)\n", nacc, mode, N, K, nm,
               cyc, (double)cyc / nm, 128.0 * N * 8 * nm / (double)cyc);
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
        return 0;
    }
    const char* names[] = {"SS plain", "TS plain (A in TMEM)", "TS 3xTF32", "SS 3xTF32"};
    const int nm = (mode >= 2 ? 3 : 1) * (K / 8);
    printf("%-22s M=128 N=%3d K=%3d: max |err| %.3e (max |d| %.2f)  %s   %d MMAs in %lld cycles (%.1f / MMA, %.0f MAC/clk)\n", names[mode], N, K,
           worst, scale, worst <= (mode >= 2 ? 2e-5 : 1e-4) * fmax(scale, 1.0) ? "OK" : "MISMATCH", nm, cyc, (double)cyc / nm,
           128.0 * N * 8 * nm / (double)cyc);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dc);
    return worst <= (mode >= 2 ? 2e-5 : 1e-4) * fmax(scale, 1.0) ? 0 : 1;
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int bad = 0;
    for (int mode = 0; mode < 4; ++mode) {
        if (only >= 0 && mode != only) continue;
        bad += run(48, 152, mode, 1);
        bad += run(112, 48, mode, 1);
        bad += run(48, 112, mode, 1);
        run(48, 152, mode, 20);
        run(112, 48, mode, 20);
        for (int nacc = 2; nacc <= 8; nacc *= 2) run(16, 152, mode, 20, nacc);      // N = 16 so that 8 accumulators fit in [0, 128)
        run(16, 152, mode, 20, 1);
        if (mode == 0 || mode == 3) { run(256, 64, mode, 20, 1); run(128, 64, mode, 20, 1); run(64, 64, mode, 20, 1); }
    }
    printf(bad ? "SELFTEST FAILED (%d)\n" : "SELFTEST OK\n", bad);
    return bad != 0;
}
