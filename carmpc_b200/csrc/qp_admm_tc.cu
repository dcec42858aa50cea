// Tensor-core form of the batched ADMM (same iteration as qp_admm.cu, same role: active-set / infeasibility finder in
// front of the float64 polish; replaces the cvxpy -> OSQP call of lib/mpc.py:334-335 / :477-478, one QP per state).
//
// A CTA owns a tile of 128 samples = the M dimension of tcgen05.mma.cta_group::1.kind::tf32.  Compute thread (slot, group)
// - 4 G compute warps, G column groups - owns lane `slot` of tensor memory and the 16-column blocks b = group, group + G, ...
// of its sample in every phase, so per-sample quantities are private to G threads (combined through shared memory once
// per round, at the convergence check).
//   tensor memory   columns [0, mp)            w^ of the general rows (state, written with tcgen05.st)
//                   next mp (2 mp) columns     accumulator of product 1 (z^)
//                   next np (2 np) columns     accumulator of products 0 and 2 (x~, Gs' dy)
//   registers       w of the box rows (16 per owned block)
//   shared memory   A-operand ring (chunks of 32 K-columns x 128 samples, TF32 hi image + lo residual image, written by
//                   the compute threads in the 128-byte swizzled K-major layout), B-operand chunks (resident for small
//                   problems, else streamed from L2 by cp.async.bulk through a ring), per-row tables, per-slot state.
// 3xTF32: every product is issued as A_hi B_hi + A_lo B_hi + A_hi B_lo (float32 accumulate): indistinguishable from the
// float32 FFMA kernel on this iteration (tests/tf32_study.py), which plain TF32 is not.  Where tensor memory allows
// (TcTables::merged) B_hi and B_lo are consumed as one operand of 2 N rows (their chunk images are adjacent), i.e. two MMAs
// per k-step, and the two accumulator halves are added when they are read.
// Roles: the compute warps run the elementwise phases (produce A chunks, consume accumulators); warp 4 G issues the MMAs and
// warp 4 G + 1 streams B chunks - both run their control flow warp-uniformly and one elected lane issues, so descriptors
// live in uniform registers.  The phases of an iteration are pipelined through mbarriers only:
//   S_B (box rows: x~ -> w_b, X chunks)  ->  product 1  ->  S_G (general rows: z^ -> w^_g, V chunks)  ->  product 0  -> ...
// with the MMAs of a product consuming the chunks while the phase that produces them is still running.
// A problem too large for tensor memory whose variable graph splits into independent chains (horizon 80) is solved as one
// pass of this kernel per chain (TcTables per part; AdmmBatch::combine merges the verdicts).
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "qp_internal.cuh"

namespace carmpc {

namespace {

constexpr int kTcGroups = 3;                       // column groups of four compute warps
constexpr int kAStageBytes = 128 * 128 * 2;        // hi image + lo image of a 128 x 32 chunk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (!done && ++spins > (1 << 20)) __trap();          // a broken pipeline must fail loudly, not hang the GPU
    }
}

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups of 1024 bytes (validated by tools/tc_bench.cu)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int N) {         // D f32, A / B tf32, both K-major, M = 128
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (the lane that issues the warp's tcgen05.mma / cp.async.bulk instructions)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\tselp.u32 %0, 1, 0, px;\n\t}"
                 : "=r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
// the loaded registers are valid only after the wait: tie them to it so that no use can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 consecutive entries of a shared-memory table with four 128-bit loads (a broadcast 4-byte load costs the same
// shared-memory wavefront as a broadcast 16-byte load, and the kernel is short of shared-memory bandwidth)
__device__ __forceinline__ void lds16(const float* p, float (&v)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}
__device__ __forceinline__ float clampf(float w, float lo, float hi) { return fminf(fmaxf(w, lo), hi); }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// x as three TF32-exact pieces (33 significant bits)
__device__ __forceinline__ void split3(double x, float& a, float& b, float& c) {
    a = tf32_hi((float)x);
    const double r1 = x - (double)a;
    b = tf32_hi((float)r1);
    c = tf32_hi((float)(r1 - (double)b));
}

// cycle counters (tensor mode 2): [0] MMA thread: round total, [1] waiting for A chunks, [2] waiting for B chunks,
// [3] compute thread 0: waiting for x~ / the certificate product, [4] waiting for z^, [5] waiting for a free A stage,
// [6] retire / refill, [7] round total, [8] rounds, [9] B stream thread: waiting for a free stage
__device__ __forceinline__ void mbar_wait_prof(uint64_t* bar, uint32_t parity, unsigned long long* acc) {
    if (acc == nullptr) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    *acc += (unsigned long long)(clock64() - t0);
}

struct TcSmem {
    unsigned char *a_ring, *b_res, *b_ring;
    float *nwd, *lam, *lb, *ub, *t;
    double *his, *gxs, *gcs;                   // [mp], [mp][4], [mp]: the per-sample bound h = his - gxs x0 - gcs c
    double* xs;                                // [5][128]  x0 (4) and the disturbance of each slot's sample
    int *slot_sample, *slot_state, *slot_iter, *slot_fresh;
    unsigned *red_res, *red_nrm;
    float *red_sup, *red_abs;
    uint64_t *a_full, *a_empty, *b_full, *b_empty, *bar_x, *bar_z, *bar_res;
    uint32_t* tmem_slot;
    unsigned long long* pc;                    // [16] cycle counters of the three profiled threads (tensor mode 2)
};

template <int NP>
__device__ __forceinline__ TcSmem tc_carve(unsigned char* raw, const TcTables& C) {
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    TcSmem s;
    s.a_ring = base;
    s.b_res = s.a_ring + (size_t)C.na_stages * kAStageBytes;
    s.b_ring = s.b_res + C.resident_bytes;
    s.his = reinterpret_cast<double*>(s.b_ring + (size_t)C.nb_stages * C.b_stage_bytes);
    s.gxs = s.his + C.mp;
    s.gcs = s.gxs + 4 * C.mp;
    s.xs = s.gcs + C.mp;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s.xs + 5 * 128);
    s.a_full = bars; s.a_empty = bars + 4; s.b_full = bars + 8; s.b_empty = bars + 10;
    s.bar_x = bars + 12; s.bar_z = bars + 13; s.bar_res = bars + 14;
    s.nwd = reinterpret_cast<float*>(bars + 16);
    s.lam = s.nwd + C.mp;
    s.lb = s.lam + NP;
    s.ub = s.lb + NP;
    s.t = s.ub + NP;
    s.slot_sample = reinterpret_cast<int*>(s.t + NP);
    s.slot_state = s.slot_sample + 128;
    s.slot_iter = s.slot_state + 128;
    s.slot_fresh = s.slot_iter + 128;
    s.red_res = reinterpret_cast<unsigned*>(s.slot_fresh + 128);
    s.red_nrm = s.red_res + 128;
    s.red_sup = reinterpret_cast<float*>(s.red_nrm + 128);
    s.red_abs = s.red_sup + 128;
    s.tmem_slot = reinterpret_cast<uint32_t*>(s.red_abs + 128);
    s.pc = reinterpret_cast<unsigned long long*>(s.tmem_slot + 2);
    return s;
}

__device__ __forceinline__ void atomic_max_pos(unsigned* addr, float v) {      // v >= 0 (NaN maps to a huge value)
    atomicMax(addr, __float_as_uint(v == v ? v : INFINITY));
}

// The compute threads' view of the A-operand ring.  A product's K columns are written in blocks of 16 (one tcgen05.ld
// block) by the column group that owns the block; a chunk of 32 columns is complete after 8 warp arrivals (two blocks x
// four lane quarters; the lone half chunk at the tail of a product arrives twice).  Every thread derives the chunk
// sequence arithmetically, the MMA warp counts the same sequence.
struct AStream {
    unsigned char* ring;
    uint64_t *full, *empty;
    int na;
    unsigned base;       // chunks published by the streams before the current one
    int kend;
    uint32_t row_off;    // byte offset of this thread's row inside an image
    int r7;
    unsigned long long* wait_acc;
};

__device__ __forceinline__ void a_begin(AStream& A, int kend) { A.kend = kend; }
__device__ __forceinline__ void a_end(AStream& A) { A.base += (unsigned)((A.kend + 31) >> 5); }

__device__ __forceinline__ void a_put16(const AStream& A, int k0, const float (&v)[16], int lane) {
    const unsigned chunk = A.base + (unsigned)(k0 >> 5);
    const unsigned use = A.na == 2 ? chunk >> 1 : chunk / 3u;
    const int stage = (int)(chunk - use * (unsigned)A.na);
    const int half = (k0 >> 4) & 1;
    mbar_wait_prof(A.empty + stage, (use & 1u) ^ 1u, A.wait_acc);
    unsigned char* hi_row = A.ring + (size_t)stage * kAStageBytes + A.row_off;
    unsigned char* lo_row = hi_row + 128 * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 h, l;
        h.x = tf32_hi(v[4 * q + 0]); h.y = tf32_hi(v[4 * q + 1]); h.z = tf32_hi(v[4 * q + 2]); h.w = tf32_hi(v[4 * q + 3]);
        l.x = v[4 * q + 0] - h.x; l.y = v[4 * q + 1] - h.y; l.z = v[4 * q + 2] - h.z; l.w = v[4 * q + 3] - h.w;
        const int piece = ((half << 2) + q) ^ A.r7;
        *reinterpret_cast<float4*>(hi_row + (piece << 4)) = h;
        *reinterpret_cast<float4*>(lo_row + (piece << 4)) = l;
    }
    const long long tf0 = A.wait_acc ? clock64() : 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    if (A.wait_acc) A.wait_acc[5] += (unsigned long long)(clock64() - tf0);          // [10]: the proxy fence
    __syncwarp();
    if (A.wait_acc) { A.wait_acc[6] += (unsigned long long)(clock64() - tf0); ++A.wait_acc[7]; }   // [11]: fence + warp sync, [12]: chunks written
    if (lane == 0) {
        mbar_arrive(A.full + stage);
        if (half == 0 && k0 + 16 == A.kend) mbar_arrive(A.full + stage);     // nobody writes the second half
    }
}

// G column groups of 4 warps: thread (slot = tid & 127, group = tid >> 7) owns the blocks b = group, group + G, ... of its
// sample in every phase; warps 4 G and 4 G + 1 issue the MMAs and stream the B chunks.
template <int NP, int G>
__global__ void __launch_bounds__(128 * G + 64, 1) admm_tc_kernel(const AdmmTables T, const TcTables C, const AdmmBatch Bq) {
    constexpr int NC = 128 * G;                     // compute threads
    constexpr int NPB = NP / 16;                    // blocks of box rows
    constexpr int MAXB = (NPB + G - 1) / G;         // box blocks per thread
    constexpr int EG = NPB % G;                     // the group that writes the constant columns (the block after the box blocks)
    extern __shared__ unsigned char smem_raw[];
    const TcSmem sm = tc_carve<NP>(smem_raw, C);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);        // provably warp-uniform: the roles below branch on it
    const int slot = tid & 127, cg = tid >> 7;
    const int mp = C.mp, m = C.m, mt = C.mt;
    const int MPB = mp >> 4;
    const float alpha = T.alpha;

    // ---- one-time setup ------------------------------------------------------------------------------------------------
    if (warp == 4 * G) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(sm.tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(sm.a_full + i, 8); mbar_init(sm.a_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(sm.b_full + i, 1); mbar_init(sm.b_empty + i, 1); }
        mbar_init(sm.bar_x, 1); mbar_init(sm.bar_z, 1); mbar_init(sm.bar_res, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid < NC) {
        for (int i = tid; i < mp; i += NC) {
            sm.nwd[i] = C.nwd[i]; sm.his[i] = C.his[i]; sm.gcs[i] = C.gcs[i];
#pragma unroll
            for (int c = 0; c < 4; ++c) sm.gxs[i * 4 + c] = C.gxs[(size_t)i * 4 + c];
        }
        for (int j = tid; j < NP; j += NC) {
            sm.lam[j] = C.lam[j]; sm.lb[j] = C.lb[j]; sm.ub[j] = C.ub[j];
            const double* kf = C.kfv + (size_t)j * 4;
            sm.t[j] = (float)(-(kf[0] * Bq.xref[0] + kf[1] * Bq.xref[1] + kf[2] * Bq.xref[2] + kf[3] * Bq.xref[3]));
        }
        if (tid < 16) sm.pc[tid] = 0;
        if (tid < 128) {
            sm.slot_sample[tid] = -1; sm.slot_state[tid] = kSlotIdle; sm.slot_iter[tid] = 0; sm.slot_fresh[tid] = 0;
            sm.red_res[tid] = 0; sm.red_nrm[tid] = 0; sm.red_sup[tid] = 0.f; sm.red_abs[tid] = 0.f;
#pragma unroll
            for (int c = 0; c < 5; ++c) sm.xs[c * 128 + tid] = 0.0;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *sm.tmem_slot;
    // merged: every accumulator is 2 N wide - columns [0, N) hold A_hi B_hi + A_lo B_hi, columns [N, 2 N) hold A_hi B_lo
    const bool merged = C.merged != 0;
    const uint32_t col_state = tbase, col_z = tbase + (uint32_t)mp, col_x = col_z + (uint32_t)(merged ? 2 * mp : mp);
    const uint32_t col_z2 = col_z + (uint32_t)mp, col_x2 = col_x + (uint32_t)NP;
    const int check_every = T.check_every;
    const int q_count = Bq.count_dev != nullptr ? min(*Bq.count_dev, Bq.count) : Bq.count;

    // role-private pipeline counters (they persist over the rounds)
    unsigned mma_a = 0, mma_b = 0, tma_b = 0;
    if (warp == 4 * G + 1 && C.resident && elect_one()) {
        mbar_expect_tx(sm.bar_res, (uint32_t)C.resident_bytes);
        for (int o = 0; o < C.resident_bytes; o += 16384)
            bulk_load(sm.b_res + o, C.img + C.off[0] + o, (uint32_t)min(16384, C.resident_bytes - o), sm.bar_res);
    }
    if (warp == 4 * G && C.resident) { mbar_wait(sm.bar_res, 0); tc_fence_after(); }

    // ---- compute-thread state --------------------------------------------------------------------------------------------
    float wb[MAXB][16];                // w of this thread's box rows: block lb covers variables 16 (lb G + group) ...
    bool drained = false;              // (group 0) the queue is empty
    unsigned n_x = 0, n_z = 0;         // completed waits on bar_x / bar_z
    AStream A;
    A.ring = sm.a_ring; A.full = sm.a_full; A.empty = sm.a_empty; A.na = C.na_stages; A.base = 0; A.kend = 0;
    A.row_off = (uint32_t)((slot >> 3) * 1024 + (slot & 7) * 128); A.r7 = slot & 7;
    unsigned long long* const pc = sm.pc;                            // cycle counters (each written by one thread)
    const bool prof_on = Bq.prof != nullptr;
    const bool prof = prof_on && tid == 0;
    A.wait_acc = prof ? &pc[5] : nullptr;
#pragma unroll
    for (int lb = 0; lb < MAXB; ++lb)
#pragma unroll
        for (int r = 0; r < 16; ++r) wb[lb][r] = 0.f;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;     // this warp's quarter of tensor memory
    const float eps_abs = T.eps_abs * Bq.eps_scale, eps_rel = T.eps_rel * Bq.eps_scale;
    auto sync_compute = [&]() { asm volatile("bar.sync 1, %0;" :: "r"(NC) : "memory"); };

    for (;;) {
        int running = 0;
        const long long t_round = clock64();
        if (warp < 4 * G) {
            // ================= retire finished samples, take the next ones from the queue =================
            // (tensor-memory loads / stores are warp-collective: every lane executes them, only the owners of a finished /
            //  fresh slot act on the values)
            const bool fin = sm.slot_state[slot] > 0;
            const int sample_old = sm.slot_sample[slot];
            if (__any_sync(0xffffffffu, fin)) {
                double xs[4] = {0.0, 0.0, 0.0, 0.0}, cd = 0.0;
                int8_t* sg = nullptr;
                float* wo = nullptr;
                if (fin) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) xs[c] = sm.xs[c * 128 + slot];
                    cd = sm.xs[4 * 128 + slot];
                    sg = Bq.sign + (size_t)sample_old * mt;
                    wo = Bq.warm + (size_t)sample_old * mt;
                }
                for (int b = cg; b < MPB; b += G) {
                    const int g0 = b << 4;
                    uint32_t wr[16];
                    tmem_ld16(col_state + lane_addr + (uint32_t)g0, wr);
                    tmem_ld_wait(wr);
                    if (fin) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const int rid = C.row_id[g0 + r];
                            if (rid < 0) continue;
                            const float w = __uint_as_float(wr[r]);
                            sg[rid] = (int8_t)((w > 0.f) - (w < sm.nwd[g0 + r]));
                            if (Bq.warm_out) {
                                const double* gx = sm.gxs + (g0 + r) * 4;
                                const double h = sm.his[g0 + r] - gx[0] * xs[0] - gx[1] * xs[1] - gx[2] * xs[2] - gx[3] * xs[3] - sm.gcs[g0 + r] * cd;
                                wo[rid] = (float)((double)w + h);
                            }
                        }
                    }
                }
                if (fin) {
#pragma unroll
                    for (int lb = 0; lb < MAXB; ++lb) {
                        const int j0 = (lb * G + cg) << 4;
                        if (j0 < NP) {
#pragma unroll
                            for (int r = 0; r < 16; ++r) {
                                const int j = j0 + r, vj = C.var_id[j];
                                if (vj >= 0) {
                                    sg[m + vj] = (int8_t)((wb[lb][r] > sm.ub[j]) - (wb[lb][r] < sm.lb[j]));
                                    if (Bq.warm_out) wo[m + vj] = wb[lb][r];
                                }
                            }
                        }
                    }
                }
            }
            sync_compute();                        // every group has read the slot's state and sample
            if (cg == 0) {
                int st = sm.slot_state[slot];
                if (st > 0) {
                    const int it = sm.slot_iter[slot];
                    int verdict = st == kSlotSolved ? CARMPC_QP_SOLVED : (st == kSlotMaxIter ? CARMPC_QP_MAX_ITER : CARMPC_QP_INFEASIBLE);
                    if (Bq.combine) {
                        // a later chain of a split problem: infeasible if any chain is, solved only if every chain is; the
                        // sample's iteration count is the longest chain's
                        const int before = Bq.status[sample_old], it_before = Bq.iters[sample_old];
                        if (before == CARMPC_QP_INFEASIBLE || verdict == CARMPC_QP_INFEASIBLE) verdict = CARMPC_QP_INFEASIBLE;
                        else if (before != CARMPC_QP_SOLVED || verdict != CARMPC_QP_SOLVED) verdict = CARMPC_QP_MAX_ITER;
                        Bq.status[sample_old] = verdict;
                        Bq.iters[sample_old] = max(it, it_before);
                        if (it > it_before) atomicAdd(Bq.total_iters, (unsigned long long)(it - it_before));
                    } else {
                        Bq.status[sample_old] = verdict;
                        Bq.iters[sample_old] = it + (Bq.iters_accumulate ? Bq.iters[sample_old] : 0);
                        atomicAdd(Bq.total_iters, (unsigned long long)it);
                    }
                    st = kSlotIdle;
                }
                int fresh = 0;
                while (st == kSlotIdle && !drained) {
                    const int qi = atomicAdd(Bq.next, 1);
                    if (qi >= q_count) { drained = true; break; }
                    const int sample = Bq.idx_list ? Bq.idx_list[qi] : qi;
                    double xs[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) xs[c] = Bq.x0[(size_t)c * Bq.stride + sample];
                    const double cd = Bq.cdist ? Bq.cdist[sample] : 0.0;
                    bool pre_ok = isfinite(xs[0]) && isfinite(xs[1]) && isfinite(xs[2]) && isfinite(xs[3]);
                    for (int k = 0; k < T.kpre; ++k) {
                        const double v = T.Px[k * 4 + 0] * xs[0] + T.Px[k * 4 + 1] * xs[1] + T.Px[k * 4 + 2] * xs[2] +
                                         T.Px[k * 4 + 3] * xs[3] + T.Pc[k] * cd;
                        pre_ok = pre_ok && v <= T.pre_hi[k] && v >= T.pre_lo[k];
                    }
                    if (!pre_ok) {
                        // a violated row that does not depend on u: infeasible without iterating (the FFMA kernel books one round)
                        if (!Bq.combine) {
                            Bq.status[sample] = CARMPC_QP_INFEASIBLE;
                            Bq.iters[sample] = check_every + (Bq.iters_accumulate ? Bq.iters[sample] : 0);
                            atomicAdd(Bq.total_iters, (unsigned long long)check_every);
                        }
                        continue;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) sm.xs[c * 128 + slot] = xs[c];
                    sm.xs[4 * 128 + slot] = cd;
                    sm.slot_sample[slot] = sample;
                    sm.slot_iter[slot] = 0;
                    fresh = 1;
                    st = kSlotRunning;
                }
                sm.slot_state[slot] = st;
                sm.slot_fresh[slot] = fresh;
                sm.red_res[slot] = 0; sm.red_nrm[slot] = 0; sm.red_sup[slot] = 0.f; sm.red_abs[slot] = 0.f;
            }
            sync_compute();
            const bool fresh = sm.slot_fresh[slot] != 0;
            if (__any_sync(0xffffffffu, fresh)) {
                // state of a new sample: w^ = w - h (cold: w = 0), box rows
                const int sample = sm.slot_sample[slot];
                const float* wi = (fresh && Bq.warm_in) ? Bq.warm + (size_t)sample * mt : nullptr;
                double xs[4] = {0.0, 0.0, 0.0, 0.0}, cd = 0.0;
                if (fresh) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) xs[c] = sm.xs[c * 128 + slot];
                    cd = sm.xs[4 * 128 + slot];
                }
                for (int b = cg; b < MPB; b += G) {
                    const int g0 = b << 4;
                    uint32_t wr[16];
                    tmem_ld16(col_state + lane_addr + (uint32_t)g0, wr);
                    tmem_ld_wait(wr);
                    if (fresh) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const int rid = C.row_id[g0 + r];
                            float w = 0.f;
                            if (rid >= 0) {
                                const double* gx = sm.gxs + (g0 + r) * 4;
                                const double h = sm.his[g0 + r] - gx[0] * xs[0] - gx[1] * xs[1] - gx[2] * xs[2] - gx[3] * xs[3] - sm.gcs[g0 + r] * cd;
                                w = (float)((wi ? (double)wi[rid] : 0.0) - h);
                            }
                            wr[r] = __float_as_uint(w);
                        }
                    }
                    tmem_st16(col_state + lane_addr + (uint32_t)g0, wr);
                }
                tmem_st_wait();
                if (fresh) {
#pragma unroll
                    for (int lb = 0; lb < MAXB; ++lb)
#pragma unroll
                        for (int r = 0; r < 16; ++r) {
                            const int j = ((lb * G + cg) << 4) + r;
                            const int vj = (wi && j < NP) ? C.var_id[j] : -1;
                            wb[lb][r] = vj >= 0 ? wi[m + vj] : 0.f;
                        }
                }
            }
            running = sm.slot_state[slot] == kSlotRunning;
        }
        if (prof && tid == 0) pc[6] += (unsigned long long)(clock64() - t_round);
        if (!__syncthreads_or(running)) break;
        const long long t_work = clock64();

        if (warp < 4 * G) {
            // ================= elementwise phases of one round (check_every iterations, the last one checked) =================
            // the constant columns of this slot's sample (x0 pieces, 1, disturbance pieces): exact TF32 values
            auto put_e = [&](int k0) {
                float e[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) split3(sm.xs[c * 128 + slot], e[3 * c], e[3 * c + 1], e[3 * c + 2]);
                e[12] = 1.f;
                split3(sm.xs[4 * 128 + slot], e[13], e[14], e[15]);
                a_put16(A, k0, e, lane);
            };
            auto put_box_v = [&]() {          // V_b = 2 clip(w_b) - w_b, then the constant columns: the head of product 0's K
#pragma unroll
                for (int lb = 0; lb < MAXB; ++lb) {
                    const int j0 = (lb * G + cg) << 4;
                    if (j0 < NP) {
                        float v[16], tl[16], tu[16];
                        lds16(sm.lb + j0, tl); lds16(sm.ub + j0, tu);
#pragma unroll
                        for (int r = 0; r < 16; ++r) { const float w = wb[lb][r]; v[r] = fmaf(2.f, clampf(w, tl[r], tu[r]), -w); }
                        a_put16(A, j0, v, lane);
                    }
                }
                if (cg == EG) put_e(NP);
            };
            // round start: every chunk of product 0 from the state
            a_begin(A, NP + 16 + mp);
            put_box_v();
            for (int b = cg; b < MPB; b += G) {
                const int g0 = b << 4;
                uint32_t wr[16];
                tmem_ld16(col_state + lane_addr + (uint32_t)g0, wr);
                tmem_ld_wait(wr);
                float v[16], tn[16];
                lds16(sm.nwd + g0, tn);
#pragma unroll
                for (int r = 0; r < 16; ++r) { const float w = __uint_as_float(wr[r]); v[r] = fmaf(2.f, clampf(w, tn[r], 0.f), -w); }
                a_put16(A, NP + 16 + g0, v, lane);
            }
            a_end(A);
            float p_res = 0.f, p_nrm = 0.f, p_sup = 0.f, p_abs = 0.f;
            for (int it = 0; it < check_every; ++it) {
                const bool check = it == check_every - 1;
                // ---------------- S_B: box rows (z = lam x~), X chunks of product 1 ----------------
                mbar_wait_prof(sm.bar_x, n_x & 1u, prof ? &pc[3] : nullptr); ++n_x;
                tc_fence_after();
                a_begin(A, NP + 16);
#pragma unroll
                for (int lb = 0; lb < MAXB; ++lb) {
                    const int j0 = (lb * G + cg) << 4;
                    if (j0 < NP) {
                        uint32_t xr[16];
                        tmem_ld16(col_x + lane_addr + (uint32_t)j0, xr);
                        if (merged) {
                            uint32_t x2[16];
                            tmem_ld16(col_x2 + lane_addr + (uint32_t)j0, x2);
                            tmem_ld_wait(xr);
                            tmem_ld_wait(x2);
#pragma unroll
                            for (int r = 0; r < 16; ++r) xr[r] = __float_as_uint(__uint_as_float(xr[r]) + __uint_as_float(x2[r]));
                        } else {
                            tmem_ld_wait(xr);
                        }
                        float x[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t4 = *reinterpret_cast<const float4*>(sm.t + j0 + 4 * q);
                            const float4 l4 = *reinterpret_cast<const float4*>(sm.lb + j0 + 4 * q);
                            const float4 u4 = *reinterpret_cast<const float4*>(sm.ub + j0 + 4 * q);
                            const float4 m4 = *reinterpret_cast<const float4*>(sm.lam + j0 + 4 * q);
                            const float tt[4] = {t4.x, t4.y, t4.z, t4.w}, tl[4] = {l4.x, l4.y, l4.z, l4.w};
                            const float tu[4] = {u4.x, u4.y, u4.z, u4.w}, tm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) {
                                const int r = 4 * q + rr, j = j0 + r;
                                x[r] = __uint_as_float(xr[r]) + tt[rr];
                                const float lb_ = tl[rr], ub_ = tu[rr];
                                const float w0 = wb[lb][r], z = tm[rr] * x[r];
                                const float c0 = clampf(w0, lb_, ub_);
                                const float w1 = fmaf(alpha, z - c0, w0);
                                wb[lb][r] = w1;
                                if (check) {
                                    const float c1 = clampf(w1, lb_, ub_), einv = C.einv_b[j];
                                    p_res = fmaxf(p_res, fmaxf(fabsf(z - c1), fabsf(c1 - c0)) * einv);
                                    p_nrm = fmaxf(p_nrm, fmaxf(fabsf(z), fabsf(c1)) * einv);
                                }
                            }
                        }
                        a_put16(A, j0, x, lane);
                    }
                }
                if (cg == EG) put_e(NP);
                a_end(A);
                if (!check) { a_begin(A, NP + 16 + mp); put_box_v(); }      // head of the next iteration's product 0
                else a_begin(A, mp);                                         // dy chunks of the certificate product
                const int kg0 = check ? 0 : NP + 16;
                // ---------------- S_G: general rows (w^ += alpha (z^ - c^)), V chunks of product 0 ----------------
                mbar_wait_prof(sm.bar_z, n_z & 1u, prof ? &pc[4] : nullptr); ++n_z;
                tc_fence_after();
                double xd[4] = {0.0, 0.0, 0.0, 0.0}, cdd = 0.0;
                if (check) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) xd[c] = sm.xs[c * 128 + slot];
                    cdd = sm.xs[4 * 128 + slot];
                }
                for (int b = cg; b < MPB; b += G) {
                    const int g0 = b << 4;
                    uint32_t zr[16], wr[16];
                    tmem_ld16(col_z + lane_addr + (uint32_t)g0, zr);
                    if (merged) {
                        tmem_ld16(col_z2 + lane_addr + (uint32_t)g0, wr);
                        tmem_ld_wait(zr);
                        tmem_ld_wait(wr);
#pragma unroll
                        for (int r = 0; r < 16; ++r) zr[r] = __float_as_uint(__uint_as_float(zr[r]) + __uint_as_float(wr[r]));
                    }
                    tmem_ld16(col_state + lane_addr + (uint32_t)g0, wr);
                    tmem_ld_wait(zr);
                    tmem_ld_wait(wr);
                    float v[16], tn[16];
                    lds16(sm.nwd + g0, tn);
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float nwd = tn[r];
                        const float w0 = __uint_as_float(wr[r]), z = __uint_as_float(zr[r]);
                        const float c0 = clampf(w0, nwd, 0.f);
                        const float w1 = fmaf(alpha, z - c0, w0);
                        const float c1 = clampf(w1, nwd, 0.f);
                        wr[r] = __float_as_uint(w1);
                        if (!check) {
                            v[r] = fmaf(2.f, c1, -w1);
                        } else {
                            const int i = g0 + r;
                            const float einv = C.einv_g[i];
                            const double* gx = sm.gxs + i * 4;
                            const float h = (float)(sm.his[i] - gx[0] * xd[0] - gx[1] * xd[1] - gx[2] * xd[2] - gx[3] * xd[3] - sm.gcs[i] * cdd);
                            p_res = fmaxf(p_res, fmaxf(fabsf(z - c1), fabsf(c1 - c0)) * einv);
                            p_nrm = fmaxf(p_nrm, fmaxf(fabsf(z + h), fabsf(c1 + h)) * einv);
                            float e = (w1 - c1) - (w0 - c0);
                            if (nwd == -INFINITY) e = fmaxf(e, 0.f);
                            v[r] = e;                        // the A operand carries delta-y for the certificate product
                            const float term = e > 0.f ? h * e : (e < 0.f ? (h + nwd) * e : 0.f);
                            p_sup += term;
                            p_abs += fabsf(term);
                        }
                    }
                    tmem_st16(col_state + lane_addr + (uint32_t)g0, wr);
                    a_put16(A, kg0 + g0, v, lane);
                }
                tmem_st_wait();
                a_end(A);
            }
            // ---------------- certificate: y_b = -(Gs' dy) / lam makes A'y = 0 exactly; infeasible iff the support sum < 0 ----------------
            mbar_wait_prof(sm.bar_x, n_x & 1u, prof ? &pc[3] : nullptr); ++n_x;
            tc_fence_after();
#pragma unroll
            for (int lb = 0; lb < MAXB; ++lb) {
                const int j0 = (lb * G + cg) << 4;
                if (j0 < NP) {
                    uint32_t yr[16];
                    tmem_ld16(col_x + lane_addr + (uint32_t)j0, yr);
                    if (merged) {
                        uint32_t y2[16];
                        tmem_ld16(col_x2 + lane_addr + (uint32_t)j0, y2);
                        tmem_ld_wait(yr);
                        tmem_ld_wait(y2);
#pragma unroll
                        for (int r = 0; r < 16; ++r) yr[r] = __float_as_uint(__uint_as_float(yr[r]) + __uint_as_float(y2[r]));
                    } else {
                        tmem_ld_wait(yr);
                    }
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const int j = j0 + r;
                        const float yb = __uint_as_float(yr[r]) * C.nrl[j];
                        const float term = yb > 0.f ? sm.ub[j] * yb : (yb < 0.f ? sm.lb[j] * yb : 0.f);
                        p_sup += term;
                        p_abs += fabsf(term);
                    }
                }
            }
            tc_fence_before();
            atomic_max_pos(sm.red_res + slot, p_res);
            atomic_max_pos(sm.red_nrm + slot, p_nrm);
            atomicAdd(sm.red_sup + slot, p_sup);
            atomicAdd(sm.red_abs + slot, p_abs);
            sync_compute();
            if (cg == 0 && sm.slot_state[slot] == kSlotRunning) {
                const int it = sm.slot_iter[slot] + check_every;
                sm.slot_iter[slot] = it;
                const float res = __uint_as_float(sm.red_res[slot]), nrm = __uint_as_float(sm.red_nrm[slot]);
                const float sup = sm.red_sup[slot], sabs = sm.red_abs[slot];
                int ns = kSlotRunning;
                if (sabs > 0.f && sup <= -T.eps_inf * sabs) ns = kSlotInfeasible;
                else if (res <= eps_abs + eps_rel * nrm) ns = kSlotSolved;
                else if (it >= Bq.max_iter || !(res == res)) ns = kSlotMaxIter;
                sm.slot_state[slot] = ns;
            }
            sync_compute();
        } else if (warp == 4 * G) {
            // ================= MMA issue: the same chunk sequence the compute threads publish =================
            // The whole warp runs the control flow (waits, counters, descriptors stay warp-uniform, so they live in uniform
            // registers); one elected lane issues the tcgen05 instructions.
            const bool leader = elect_one();
            unsigned long long* const pa = (prof_on && leader) ? &pc[1] : nullptr;
            unsigned long long* const pb = (prof_on && leader) ? &pc[2] : nullptr;
            auto product = [&](int p, uint32_t dcol, bool streamed, uint64_t* done) {
                const int N = C.ncols[p];
                const uint32_t idesc = make_idesc(N), idesc2 = make_idesc(2 * N);
                const int e_ks = p == 2 ? -1 : NP / 8;              // the two k-steps of the constant columns: their lo image is zero
                const unsigned char* res_base = sm.b_res + (p == 1 ? (size_t)C.nchunks[0] * C.pair_bytes[0] : 0);
                for (int c = 0; c < C.nchunks[p]; ++c) {
                    const unsigned use = C.na_stages == 2 ? mma_a >> 1 : mma_a / 3u;
                    const int sa = (int)(mma_a - use * (unsigned)C.na_stages);
                    mbar_wait_prof(sm.a_full + sa, use & 1u, pa);
                    const unsigned char* bsrc;
                    int sb = 0;
                    if (streamed) {
                        sb = C.nb_stages == 2 ? (int)(mma_b & 1u) : 0;
                        mbar_wait_prof(sm.b_full + sb, (C.nb_stages == 2 ? mma_b >> 1 : mma_b) & 1u, pb);
                        bsrc = sm.b_ring + (size_t)sb * C.b_stage_bytes;
                    } else {
                        bsrc = res_base + (size_t)c * C.pair_bytes[p];
                    }
                    tc_fence_after();
                    const uint64_t a_hi = desc_sw128(smem_u32(sm.a_ring + (size_t)sa * kAStageBytes));
                    const uint64_t a_lo = a_hi + (uint64_t)((128 * 128) >> 4);
                    const uint64_t b_hi = desc_sw128(smem_u32(bsrc));
                    const uint64_t b_lo = b_hi + (uint64_t)((N * 128) >> 4);
                    const int nks = min(4, C.ksteps[p] - 4 * c);
                    const bool e_chunk = 4 * c <= e_ks + 1 && e_ks < 4 * c + 4;       // this chunk holds constant columns
                    if (leader) {
                        if (merged) {
                            // A_hi x [B_hi; B_lo] (the lo image follows the hi image: one operand of 2 N rows), then A_lo x B_hi
                            for (int ks = 0; ks < nks; ++ks) {
                                const uint64_t o = (uint64_t)(ks * 2);
                                const int kg = 4 * c + ks;
                                mma_ss(dcol, a_hi + o, b_hi + o, idesc2, (uint32_t)(kg != 0));
                                if (kg != e_ks && kg != e_ks + 1) mma_ss(dcol, a_lo + o, b_hi + o, idesc, 1u);
                            }
                        } else if (!e_chunk && nks == 4) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t o = (uint64_t)(ks * 2);
                                mma_ss(dcol, a_hi + o, b_hi + o, idesc, (uint32_t)((c | ks) != 0));
                                mma_ss(dcol, a_lo + o, b_hi + o, idesc, 1u);
                                mma_ss(dcol, a_hi + o, b_lo + o, idesc, 1u);
                            }
                        } else {
                            for (int ks = 0; ks < nks; ++ks) {
                                const uint64_t o = (uint64_t)(ks * 2);
                                const int kg = 4 * c + ks;
                                mma_ss(dcol, a_hi + o, b_hi + o, idesc, (uint32_t)(kg != 0));
                                if (kg != e_ks && kg != e_ks + 1) mma_ss(dcol, a_lo + o, b_hi + o, idesc, 1u);
                                mma_ss(dcol, a_hi + o, b_lo + o, idesc, 1u);
                            }
                        }
                        mma_commit(sm.a_empty + sa);
                        if (streamed) mma_commit(sm.b_empty + sb);
                    }
                    __syncwarp();
                    ++mma_a;
                    if (streamed) ++mma_b;
                }
                if (leader) mma_commit(done);
                __syncwarp();
            };
            for (int it = 0; it < check_every; ++it) {
                product(0, col_x, !C.resident, sm.bar_x);
                product(1, col_z, !C.resident, sm.bar_z);
            }
            product(2, col_x, true, sm.bar_x);
            if (prof_on && leader) pc[0] += (unsigned long long)(clock64() - t_work);
        } else {
            // ================= B-operand stream (L2 -> shared memory ring) =================
            const bool leader = elect_one();
            unsigned long long* const pw = (prof_on && leader) ? &pc[9] : nullptr;
            auto stream = [&](int p) {
                for (int c = 0; c < C.nchunks[p]; ++c) {
                    const int sb = C.nb_stages == 2 ? (int)(tma_b & 1u) : 0;
                    mbar_wait_prof(sm.b_empty + sb, ((C.nb_stages == 2 ? tma_b >> 1 : tma_b) & 1u) ^ 1u, pw);
                    if (leader) {
                        const uint32_t bytes = (uint32_t)C.pair_bytes[p];
                        mbar_expect_tx(sm.b_full + sb, bytes);
                        const unsigned char* src = C.img + C.off[p] + (size_t)c * bytes;
                        unsigned char* dst = sm.b_ring + (size_t)sb * C.b_stage_bytes;
                        for (uint32_t o = 0; o < bytes; o += 16384) bulk_load(dst + o, src + o, min(16384u, bytes - o), sm.b_full + sb);
                    }
                    __syncwarp();
                    ++tma_b;
                }
            };
            for (int it = 0; it < check_every; ++it)
                if (!C.resident) { stream(0); stream(1); }
            stream(2);
        }
        if (prof) { pc[7] += (unsigned long long)(clock64() - t_work); ++pc[8]; }
        __syncwarp();
    }
    __syncthreads();
    if (prof_on && tid < 13) atomicAdd(Bq.prof + tid, pc[tid]);

    tc_fence_before();
    __syncthreads();
    if (warp == 4 * G) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

template <int NP, int G>
int tc_launch_np(QPHandle* q, const TcTables& part, const AdmmBatch& b, cudaStream_t st) {
    const size_t smem = (size_t)part.smem_bytes;
    const int64_t tiles = ((int64_t)b.count + 127) / 128;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, q->sm));
    constexpr int threads = 128 * G + 64;
    { const int rc = kernel_config(reinterpret_cast<const void*>(admm_tc_kernel<NP, G>), threads, smem, nullptr); if (rc != CARMPC_OK) return rc; }
    admm_tc_kernel<NP, G><<<blocks, threads, smem, st>>>(q->admm, part, b);
    CARMPC_CUDA(cudaGetLastError());
    return CARMPC_OK;
}

}  // namespace

bool admm_tc_usable(const QPHandle* q, const AdmmBatch& b) {
    // Large first passes only: the tile is always 128 samples wide, so a batch that cannot give every SM a full tile
    // (second passes, closed-loop steps) stays on the FFMA kernel's narrow tiles; so does a launch that must export the
    // raw iterate (write_u: the accumulator that holds x~ is reused by the next product).
    // Mode 1 (default) takes it where it measured faster than the FFMA kernel: problems whose matrices the FFMA kernel has
    // to read through L1 / L2 (horizon 40: 1.75x on the whole solve); with the matrices in shared memory (horizons 10, 20)
    // the FFMA kernel is 20 % ahead.  Modes 2 / 3 take every problem that has a tensor-core form.
    const bool wanted = q->tensor_mode >= 2 || (q->tensor_mode == 1 && !q->host.mats_in_smem);
    return wanted && !q->tc_parts.empty() && !b.narrow && !b.write_u && b.warm != nullptr && b.sign != nullptr &&
           (int64_t)b.count >= (int64_t)128 * q->sm;
}

int admm_tc_launch(QPHandle* q, const AdmmBatch& b_in, cudaStream_t st) {
    // one pass per part: the whole problem, or the independent chains one after the other (same samples, disjoint rows and
    // variables; the later passes merge their verdicts into the earlier ones')
    for (size_t k = 0; k < q->tc_parts.size(); ++k) {
        const TcTables& part = q->tc_parts[k];
        AdmmBatch b = b_in;
        b.combine = k > 0 ? 1 : 0;
        if (k > 0) CARMPC_CUDA(cudaMemsetAsync(b.next, 0, sizeof(int), st));
        int rc = CARMPC_ERR_UNSUPPORTED;
        switch (part.np) {
            case 16: rc = tc_launch_np<16, kTcGroups>(q, part, b, st); break;
            case 32: rc = tc_launch_np<32, kTcGroups>(q, part, b, st); break;
            case 48: rc = tc_launch_np<48, kTcGroups>(q, part, b, st); break;
            case 64: rc = tc_launch_np<64, kTcGroups>(q, part, b, st); break;
            case 80: rc = tc_launch_np<80, kTcGroups>(q, part, b, st); break;
            default: set_error("admm_tc_launch: no kernel variant for %d variables", part.np);
        }
        if (rc != CARMPC_OK) return rc;
    }
    return CARMPC_OK;
}

}  // namespace carmpc
