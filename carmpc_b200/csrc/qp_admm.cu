// Batched float32 ADMM over the condensed MPC QP (replaces the cvxpy -> OSQP call of lib/mpc.py:334-335 / :477-478,
// one QP per initial state) - the active-set / infeasibility finder in front of the float64 polish (qp_polish.cu).
//
// Tiling.  A CTA of 8 warps owns a tile of Bt = 32 S samples ("slots"); lane l owns slots l S .. l S + S - 1, the
// warps split the rows.  Each iteration is two FFMA tile products against matrices shared by every sample:
//     stage A   x~ (n x Bt)  = P (n x K) V (K x Bt) + x~0          thread tile kRA rows x S samples
//     stage B   z  (m x Bt)  = Gs (m x n) x~ (n x Bt)              thread tile kRB rows x S samples
// Matrix fragments are warp-broadcast 128-bit loads (shared memory when the matrices fit, else L1-cached global),
// sample fragments are conflict-free S-wide vector loads; the ADMM state w of a (row, sample) lives in the registers
// of the thread that computes its z.  Row groups carry K ranges computed on the host from the problem structure
// (independent acceleration / steering chains, causal state rows), so structural zeros are never multiplied.
//
// Scheduling.  The grid is persistent (one CTA per SM).  Samples converge after very different iteration counts, so
// slots are refilled from a global queue at every convergence check: a finished sample leaves its slot, writes its
// active-set signs / iterate / status, and the next queued sample takes the slot over (state reset in registers).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "qp_internal.cuh"

namespace carmpc {

namespace {

template <int S> struct Vec;
template <> struct Vec<1> {
    float v[1];
    __device__ __forceinline__ static Vec ld(const float* p) { Vec r; r.v[0] = *p; return r; }
    __device__ __forceinline__ void st(float* p) const { *p = v[0]; }
};
template <> struct Vec<2> {
    float v[2];
    __device__ __forceinline__ static Vec ld(const float* p) { const float2 t = *reinterpret_cast<const float2*>(p); Vec r; r.v[0] = t.x; r.v[1] = t.y; return r; }
    __device__ __forceinline__ void st(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <> struct Vec<4> {
    float v[4];
    __device__ __forceinline__ static Vec ld(const float* p) { const float4 t = *reinterpret_cast<const float4*>(p); Vec r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
    __device__ __forceinline__ void st(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

template <bool SMEM>
__device__ __forceinline__ float4 ld_mat(const float* p) {
    if (SMEM) return *reinterpret_cast<const float4*>(p);
    return __ldg(reinterpret_cast<const float4*>(p));
}

// acc[r][s] += sum_{k in [kb, ke)} M[r][k] * B[k][s0 + s]      (kb, ke multiples of 4; M row stride ld; B row stride Bt)
// S >= 2: the samples of a lane are processed in pairs with FFMA2 (matrix element broadcast, sample pair as loaded from
// shared memory); every accumulator sees the same fma sequence as the scalar loop, so results are bit-identical.
template <int R, int S, bool SMEM>
__device__ __forceinline__ void tile_product(float (&acc)[R][S], const float* __restrict__ M, int ld,
                                             const float* __restrict__ B, int Bt, int s0, int kb, int ke) {
    if constexpr (S == 1) {
#pragma unroll 2
        for (int k = kb; k < ke; k += 4) {
            const float b0 = B[(k + 0) * Bt + s0], b1 = B[(k + 1) * Bt + s0], b2 = B[(k + 2) * Bt + s0], b3 = B[(k + 3) * Bt + s0];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 a = ld_mat<SMEM>(M + r * ld + k);
                float t = acc[r][0];
                t = fmaf(a.x, b0, t);
                t = fmaf(a.y, b1, t);
                t = fmaf(a.z, b2, t);
                t = fmaf(a.w, b3, t);
                acc[r][0] = t;
            }
        }
    } else {
        constexpr int S2 = S / 2;
        f32x2 acc2[R][S2];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int s = 0; s < S2; ++s) acc2[r][s] = pack2(acc[r][2 * s], acc[r][2 * s + 1]);
#pragma unroll 2
        for (int k = kb; k < ke; k += 4) {
            const Vec<S> b0 = Vec<S>::ld(B + (k + 0) * Bt + s0);
            const Vec<S> b1 = Vec<S>::ld(B + (k + 1) * Bt + s0);
            const Vec<S> b2 = Vec<S>::ld(B + (k + 2) * Bt + s0);
            const Vec<S> b3 = Vec<S>::ld(B + (k + 3) * Bt + s0);
            f32x2 p0[S2], p1[S2], p2[S2], p3[S2];
#pragma unroll
            for (int s = 0; s < S2; ++s) {
                p0[s] = pack2(b0.v[2 * s], b0.v[2 * s + 1]); p1[s] = pack2(b1.v[2 * s], b1.v[2 * s + 1]);
                p2[s] = pack2(b2.v[2 * s], b2.v[2 * s + 1]); p3[s] = pack2(b3.v[2 * s], b3.v[2 * s + 1]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 a = ld_mat<SMEM>(M + r * ld + k);
                const f32x2 ax = pack2(a.x, a.x), ay = pack2(a.y, a.y), az = pack2(a.z, a.z), aw = pack2(a.w, a.w);
#pragma unroll
                for (int s = 0; s < S2; ++s) {
                    f32x2 t = acc2[r][s];
                    t = ffma2(ax, p0[s], t);
                    t = ffma2(ay, p1[s], t);
                    t = ffma2(az, p2[s], t);
                    t = ffma2(aw, p3[s], t);
                    acc2[r][s] = t;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int s = 0; s < S2; ++s) unpack2(acc2[r][s], acc[r][2 * s], acc[r][2 * s + 1]);
    }
}

__device__ __forceinline__ float clampf(float w, float lo, float hi) { return fminf(fmaxf(w, lo), hi); }

__device__ __forceinline__ void atomic_max_pos(unsigned* addr, float v) {      // v >= 0 (NaN maps to a huge value)
    atomicMax(addr, __float_as_uint(v == v ? v : INFINITY));
}

struct Smem {
    float *V, *Xt, *HI, *Pm, *Gm, *width, *lam, *lbs, *ubs;
    float *einv_g, *esc_g, *einv_b, *esc_b, *dinv, *dsc;     // per-row / per-variable scales (checks, outputs)
    double *his, *gxs, *gcs, *kf;                             // slot (re)initialisation tables
    int *vpos, *row_id, *var_id;
    int4* segA;
    int2* segB;
    double* x0;          // [5][Bt]  x, y, psi, v, c
    int *slot_sample, *slot_state, *slot_iter, *slot_init, *slot_pos;
    unsigned *red_res, *red_nrm;
    float *red_sup, *red_abs;
    int* misc;           // [0] n_free, [1] base
};

template <int S, int H, bool MATS>
__device__ __forceinline__ Smem carve(unsigned char* raw, const AdmmTables& T) {
    constexpr int Bt = 32 * S * H;
    Smem s;
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = raw + off; off += (bytes + 15) & ~size_t(15); return p; };
    s.x0 = reinterpret_cast<double*>(take(sizeof(double) * 5 * Bt));
    s.his = reinterpret_cast<double*>(take(sizeof(double) * T.m_phys));
    s.gxs = reinterpret_cast<double*>(take(sizeof(double) * 4 * T.m_phys));
    s.gcs = reinterpret_cast<double*>(take(sizeof(double) * T.m_phys));
    s.kf = reinterpret_cast<double*>(take(sizeof(double) * 4 * T.nA_rows));
    s.V = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T.ktot * Bt));
    s.Xt = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T.npad4 * Bt));
    s.HI = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T.m_phys * Bt));
    if (MATS) {
        s.Pm = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T.nA_rows * T.ktot));
        s.Gm = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T.m_phys * T.npad4));
    } else {
        s.Pm = nullptr; s.Gm = nullptr;
    }
    s.width = reinterpret_cast<float*>(take(sizeof(float) * T.m_phys));
    s.lam = reinterpret_cast<float*>(take(sizeof(float) * T.nA_rows));
    s.lbs = reinterpret_cast<float*>(take(sizeof(float) * T.nA_rows));
    s.ubs = reinterpret_cast<float*>(take(sizeof(float) * T.nA_rows));
    s.einv_g = reinterpret_cast<float*>(take(sizeof(float) * T.m_phys));
    s.esc_g = reinterpret_cast<float*>(take(sizeof(float) * T.m_phys));
    s.einv_b = reinterpret_cast<float*>(take(sizeof(float) * T.nA_rows));
    s.esc_b = reinterpret_cast<float*>(take(sizeof(float) * T.nA_rows));
    s.dinv = reinterpret_cast<float*>(take(sizeof(float) * T.nA_rows));
    s.dsc = reinterpret_cast<float*>(take(sizeof(float) * T.nA_rows));
    s.vpos = reinterpret_cast<int*>(take(sizeof(int) * T.m_phys));
    s.row_id = reinterpret_cast<int*>(take(sizeof(int) * T.m_phys));
    s.var_id = reinterpret_cast<int*>(take(sizeof(int) * T.nA_rows));
    s.segA = reinterpret_cast<int4*>(take(sizeof(int4) * T.nGA));
    s.segB = reinterpret_cast<int2*>(take(sizeof(int2) * T.nGB));
    s.slot_sample = reinterpret_cast<int*>(take(sizeof(int) * Bt));
    s.slot_state = reinterpret_cast<int*>(take(sizeof(int) * Bt));
    s.slot_iter = reinterpret_cast<int*>(take(sizeof(int) * Bt));
    s.slot_init = reinterpret_cast<int*>(take(sizeof(int) * Bt));
    s.slot_pos = reinterpret_cast<int*>(take(sizeof(int) * Bt));
    s.red_res = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * Bt));
    s.red_nrm = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * Bt));
    s.red_sup = reinterpret_cast<float*>(take(sizeof(float) * Bt));
    s.red_abs = reinterpret_cast<float*>(take(sizeof(float) * Bt));
    s.misc = reinterpret_cast<int*>(take(sizeof(int) * 4));
    return s;
}

// S samples per lane, H sample-halves: 8 H warps; warp w owns row groups (w & 7) of the samples of half (w >> 3).
// More warps per scheduler (H = 2) hide shared-memory latency better; fewer (H = 1, larger S) reuse each matrix
// fragment over more samples.
template <int S, int H, int GA, int GB, bool MATS>
__global__ void __launch_bounds__(kAdmmThreads * H, 1) admm_kernel(const AdmmTables T, const AdmmBatch Bq) {
    constexpr int Bt = 32 * S * H;
    constexpr int NT = kAdmmThreads * H;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem sm = carve<S, H, MATS>(smem_raw, T);
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = (tid >> 5) & (kAdmmWarps - 1);            // row-group owner index
    const int s0 = (tid >> 8) * 32 * S + lane * S;             // first slot of this lane
    const float* __restrict__ Pm = MATS ? sm.Pm : T.P;
    const float* __restrict__ Gm = MATS ? sm.Gm : T.Gs;
    const float alpha = T.alpha;

    // ---- one-time staging ------------------------------------------------------------------------------------------------
    for (int i = tid; i < T.ktot * Bt; i += NT) sm.V[i] = 0.f;
    for (int i = tid; i < T.npad4 * Bt; i += NT) sm.Xt[i] = 0.f;
    for (int i = tid; i < T.m_phys * Bt; i += NT) sm.HI[i] = 3.0e38f;
    if (MATS) {
        const float4* src = reinterpret_cast<const float4*>(T.P);
        float4* dst = reinterpret_cast<float4*>(sm.Pm);
        for (int i = tid; i < T.nA_rows * T.ktot / 4; i += NT) dst[i] = src[i];
        src = reinterpret_cast<const float4*>(T.Gs);
        dst = reinterpret_cast<float4*>(sm.Gm);
        for (int i = tid; i < T.m_phys * T.npad4 / 4; i += NT) dst[i] = src[i];
    }
    for (int i = tid; i < T.m_phys; i += NT) {
        sm.width[i] = T.width[i]; sm.vpos[i] = T.vpos[i]; sm.row_id[i] = T.row_id[i];
        sm.einv_g[i] = T.Einv_g[i]; sm.esc_g[i] = T.Esc_g[i];
        sm.his[i] = T.his[i]; sm.gcs[i] = T.Gcs[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) sm.gxs[i * 4 + c] = T.Gxs[i * 4 + c];
    }
    for (int i = tid; i < T.nA_rows; i += NT) {
        sm.lam[i] = T.lam[i]; sm.lbs[i] = T.lbs[i]; sm.ubs[i] = T.ubs[i]; sm.var_id[i] = T.var_id[i];
        sm.einv_b[i] = T.Einv_b[i]; sm.esc_b[i] = T.Esc_b[i]; sm.dinv[i] = T.Dinv[i]; sm.dsc[i] = T.Dsc[i];
#pragma unroll
        for (int c = 0; c < 4; ++c) sm.kf[i * 4 + c] = T.KF[i * 4 + c];
    }
    for (int i = tid; i < T.nGA; i += NT) sm.segA[i] = T.segA[i];
    for (int i = tid; i < T.nGB; i += NT) sm.segB[i] = T.segB[i];
    if (tid < Bt) {
        sm.slot_sample[tid] = -1; sm.slot_state[tid] = kSlotIdle; sm.slot_iter[tid] = 0; sm.slot_init[tid] = 0;
        sm.red_res[tid] = 0; sm.red_nrm[tid] = 0; sm.red_sup[tid] = 0.f; sm.red_abs[tid] = 0.f;
    }
    if (tid == 0) { sm.misc[0] = 0; sm.misc[1] = 0; }

    // ---- persistent per-thread state ----------------------------------------------------------------------------------------
    float wB[GB][kRB][S];          // general rows owned by this thread
    float wA[GA][kRA][S];          // box rows (one per variable) owned by this thread
    float x0t[GA][kRA][S];         // x~0 = -K^-1 q of the slot's sample
#pragma unroll
    for (int g = 0; g < GB; ++g)
#pragma unroll
        for (int r = 0; r < kRB; ++r)
#pragma unroll
            for (int s = 0; s < S; ++s) wB[g][r][s] = 0.f;
#pragma unroll
    for (int g = 0; g < GA; ++g)
#pragma unroll
        for (int r = 0; r < kRA; ++r)
#pragma unroll
            for (int s = 0; s < S; ++s) { wA[g][r][s] = 0.f; x0t[g][r][s] = 0.f; }
    __syncthreads();

    const float eps_abs = T.eps_abs * Bq.eps_scale, eps_rel = T.eps_rel * Bq.eps_scale;
    const int q_count = Bq.count_dev != nullptr ? min(*Bq.count_dev, Bq.count) : Bq.count;

    // One ADMM iteration.  CHECK: also accumulate the residual norms and run the infeasibility-certificate product.
    auto iteration = [&](auto check_tag) {
        constexpr bool CHECK = decltype(check_tag)::value;
        // ---------------- stage A: x~ = P V + x~0 ----------------
        float xt[GA][kRA][S];
#pragma unroll
        for (int g = 0; g < GA; ++g) {
            const int pg = g * kAdmmWarps + warp;
            if (pg < T.nGA) {
                const int4 sg = sm.segA[pg];
                float acc[kRA][S];
#pragma unroll
                for (int r = 0; r < kRA; ++r)
#pragma unroll
                    for (int s = 0; s < S; ++s) acc[r][s] = x0t[g][r][s];
                const float* Mrow = Pm + (size_t)(pg * kRA) * T.ktot;
                tile_product<kRA, S, MATS>(acc, Mrow, T.ktot, sm.V, Bt, s0, sg.x, sg.y);
                tile_product<kRA, S, MATS>(acc, Mrow, T.ktot, sm.V, Bt, s0, sg.z, sg.w);
#pragma unroll
                for (int r = 0; r < kRA; ++r) {
                    Vec<S> o;
#pragma unroll
                    for (int s = 0; s < S; ++s) { xt[g][r][s] = acc[r][s]; o.v[s] = acc[r][s]; }
                    o.st(sm.Xt + (pg * kRA + r) * Bt + s0);
                }
            }
        }
        __syncthreads();                                   // x~ complete; every read of V is done
        float p_res[S], p_nrm[S], p_abs[S], p_sup[S];
        if (CHECK) {
#pragma unroll
            for (int s = 0; s < S; ++s) { p_res[s] = 0.f; p_nrm[s] = 0.f; p_abs[s] = 0.f; p_sup[s] = 0.f; }
        }
        // ---------------- box rows: z = lam x~ ----------------
#pragma unroll
        for (int g = 0; g < GA; ++g) {
            const int pg = g * kAdmmWarps + warp;
            if (pg < T.nGA) {
#pragma unroll
                for (int r = 0; r < kRA; ++r) {
                    const int j = pg * kRA + r;
                    const float lam = sm.lam[j], lb = sm.lbs[j], ub = sm.ubs[j];
                    Vec<S> o;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const float w0 = wA[g][r][s];
                        const float z = lam * xt[g][r][s];
                        const float c0 = clampf(w0, lb, ub);
                        const float w1 = fmaf(alpha, z - c0, w0);
                        const float c1 = clampf(w1, lb, ub);
                        wA[g][r][s] = w1;
                        o.v[s] = 2.f * c1 - w1;
                        if (CHECK) {
                            const float einv = sm.einv_b[j];
                            p_res[s] = fmaxf(p_res[s], fmaxf(fabsf(z - c1), fabsf(c1 - c0)) * einv);
                            p_nrm[s] = fmaxf(p_nrm[s], fmaxf(fabsf(z), fabsf(c1)) * einv);
                        }
                    }
                    o.st(sm.V + (T.mv4 + j) * Bt + s0);
                }
            }
        }
        // ---------------- stage B: z = Gs x~, then the w update of the general rows ----------------
#pragma unroll
        for (int g = 0; g < GB; ++g) {
            const int pg = g * kAdmmWarps + warp;
            const int2 sg = sm.segB[pg];
            float acc[kRB][S];
#pragma unroll
            for (int r = 0; r < kRB; ++r)
#pragma unroll
                for (int s = 0; s < S; ++s) acc[r][s] = 0.f;
            tile_product<kRB, S, MATS>(acc, Gm + (size_t)(pg * kRB) * T.npad4, T.npad4, sm.Xt, Bt, s0, sg.x, sg.y);
#pragma unroll
            for (int r = 0; r < kRB; ++r) {
                const int i = pg * kRB + r;
                const float wd = sm.width[i];
                const int vp = sm.vpos[i];
                const Vec<S> hi = Vec<S>::ld(sm.HI + i * Bt + s0);
                Vec<S> o;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const float h = hi.v[s], lo = h - wd;
                    const float w0 = wB[g][r][s];
                    const float z = acc[r][s];
                    const float c0 = clampf(w0, lo, h);
                    const float w1 = fmaf(alpha, z - c0, w0);
                    const float c1 = clampf(w1, lo, h);
                    wB[g][r][s] = w1;
                    if (!CHECK) {
                        o.v[s] = 2.f * c1 - w1;
                    } else {
                        const float einv = sm.einv_g[i];
                        p_res[s] = fmaxf(p_res[s], fmaxf(fabsf(z - c1), fabsf(c1 - c0)) * einv);
                        p_nrm[s] = fmaxf(p_nrm[s], fmaxf(fabsf(z), fabsf(c1)) * einv);
                        float e = (w1 - c1) - (w0 - c0);
                        if (wd == INFINITY) e = fmaxf(e, 0.f);
                        o.v[s] = e;                         // V carries delta-y for the certificate product
                        const float term = e > 0.f ? h * e : (e < 0.f ? lo * e : 0.f);
                        p_sup[s] += term;
                        p_abs[s] += fabsf(term);
                    }
                }
                if (vp >= 0) o.st(sm.V + vp * Bt + s0);
            }
        }
        if (CHECK) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                atomic_max_pos(sm.red_res + s0 + s, p_res[s]);
                atomic_max_pos(sm.red_nrm + s0 + s, p_nrm[s]);
                atomicAdd(sm.red_abs + s0 + s, p_abs[s]);
                atomicAdd(sm.red_sup + s0 + s, p_sup[s]);
            }
        }
        __syncthreads();                                   // V (or delta-y) complete
        if (CHECK) {
            // ---------------- infeasibility certificate (Farkas, with the box rows absorbing the residual) ----------------
            // y_g = delta-y of the general rows (in V).  Every variable has a finite box, so the multipliers of the box rows
            // can be chosen as y_b = -(Gs' y_g) / lam, which makes A_s' y = 0 exactly; the problem is infeasible iff the
            // support function  hi' y_g+ + lo' y_g- + ub' y_b+ + lb' y_b-  is negative.  No tolerance on ||A' y|| is needed,
            // which certifies barely infeasible states long before OSQP's test (||A' dy|| <= eps ||dy||) would.
#pragma unroll
            for (int s = 0; s < S; ++s) { p_sup[s] = 0.f; p_abs[s] = 0.f; }
#pragma unroll
            for (int g = 0; g < GA; ++g) {
                const int pg = g * kAdmmWarps + warp;
                if (pg < T.nGA) {
                    const int4 sg = sm.segA[pg];
                    float acc[kRA][S];
#pragma unroll
                    for (int r = 0; r < kRA; ++r)
#pragma unroll
                        for (int s = 0; s < S; ++s) acc[r][s] = 0.f;
                    tile_product<kRA, S, false>(acc, T.GsT + (size_t)(pg * kRA) * T.mv4, T.mv4, sm.V, Bt, s0, sg.x, sg.y);
#pragma unroll
                    for (int r = 0; r < kRA; ++r) {
                        const int j = pg * kRA + r;
                        const float lam = sm.lam[j], lb = sm.lbs[j], ub = sm.ubs[j];
                        const float rl = lam > 0.f ? -1.f / lam : 0.f;            // pad variables: Gs' column is zero
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const float yb = acc[r][s] * rl;
                            const float term = yb > 0.f ? ub * yb : (yb < 0.f ? lb * yb : 0.f);   // +inf when that side is unbounded
                            p_sup[s] += term;
                            p_abs[s] += fabsf(term);
                        }
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                atomicAdd(sm.red_abs + s0 + s, p_abs[s]);
                atomicAdd(sm.red_sup + s0 + s, p_sup[s]);
            }
            __syncthreads();                               // delta-y fully consumed
            // restore V = 2 clip(w) - w for the general rows
#pragma unroll
            for (int g = 0; g < GB; ++g) {
                const int pg = g * kAdmmWarps + warp;
#pragma unroll
                for (int r = 0; r < kRB; ++r) {
                    const int i = pg * kRB + r;
                    const float wd = sm.width[i];
                    const int vp = sm.vpos[i];
                    const Vec<S> hi = Vec<S>::ld(sm.HI + i * Bt + s0);
                    Vec<S> o;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        const float w = wB[g][r][s];
                        o.v[s] = 2.f * clampf(w, hi.v[s] - wd, hi.v[s]) - w;
                    }
                    if (vp >= 0) o.st(sm.V + vp * Bt + s0);
                }
            }
        }
    };

    for (;;) {
        // ================= retire finished slots, refill from the queue =================
        // (entered at start with every slot idle, and after every checked iteration)
        int fin[S];
#pragma unroll
        for (int s = 0; s < S; ++s) { const int st = sm.slot_state[s0 + s]; fin[s] = st > 0 ? sm.slot_sample[s0 + s] : -1; }
        // outputs of the finished samples, by the owners of the rows
        bool any_fin = false;
#pragma unroll
        for (int s = 0; s < S; ++s) any_fin |= fin[s] >= 0;
        if (any_fin) {
#pragma unroll
        for (int g = 0; g < GB; ++g) {
            const int pg = g * kAdmmWarps + warp;
#pragma unroll
            for (int r = 0; r < kRB; ++r) {
                const int i = pg * kRB + r;
                const int rid = sm.row_id[i];
                if (rid < 0) continue;
                const float wd = sm.width[i];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    if (fin[s] < 0) continue;
                    const float h = sm.HI[i * Bt + s0 + s], w = wB[g][r][s];
                    Bq.sign[(size_t)fin[s] * T.mt + rid] = (int8_t)((w > h) - (w < h - wd));
                    if (Bq.warm_out) Bq.warm[(size_t)fin[s] * T.mt + rid] = w;
                }
            }
        }
#pragma unroll
        for (int g = 0; g < GA; ++g) {
            const int pg = g * kAdmmWarps + warp;
            if (pg >= T.nGA) continue;
#pragma unroll
            for (int r = 0; r < kRA; ++r) {
                const int j = pg * kRA + r;
                const int vid = sm.var_id[j];
                if (vid < 0) continue;
                const float lb = sm.lbs[j], ub = sm.ubs[j], d = sm.dsc[j];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    if (fin[s] < 0) continue;
                    const float w = wA[g][r][s];
                    Bq.sign[(size_t)fin[s] * T.mt + T.m + vid] = (int8_t)((w > ub) - (w < lb));
                    if (Bq.write_u) Bq.u_admm[(size_t)fin[s] * T.n + vid] = d * sm.Xt[j * Bt + s0 + s];
                    if (Bq.warm_out) Bq.warm[(size_t)fin[s] * T.mt + T.m + vid] = w;
                }
            }
        }
        }
        if (tid < Bt) {
            const int st = sm.slot_state[tid];
            if (st > 0) {
                const int sample = sm.slot_sample[tid];
                Bq.status[sample] = st == kSlotSolved ? CARMPC_QP_SOLVED : (st == kSlotMaxIter ? CARMPC_QP_MAX_ITER : CARMPC_QP_INFEASIBLE);
                Bq.iters[sample] = sm.slot_iter[tid] + (Bq.iters_accumulate ? Bq.iters[sample] : 0);
                atomicAdd(Bq.total_iters, (unsigned long long)sm.slot_iter[tid]);
            }
            sm.slot_pos[tid] = st != kSlotRunning ? atomicAdd(&sm.misc[0], 1) : -1;
        }
        __syncthreads();
        if (tid == 0) {
            const int nfree = sm.misc[0];
            sm.misc[1] = nfree > 0 ? atomicAdd(Bq.next, nfree) : 0;
            sm.misc[0] = 0;
        }
        __syncthreads();
        if (tid < Bt) {
            int init = 0;
            if (sm.slot_pos[tid] >= 0) {
                const int qi = sm.misc[1] + sm.slot_pos[tid];
                if (qi < q_count) {
                    const int sample = Bq.idx_list ? Bq.idx_list[qi] : qi;
                    double x[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) { x[c] = Bq.x0[(size_t)c * Bq.stride + sample]; sm.x0[c * Bt + tid] = x[c]; }
                    const double cd = Bq.cdist ? Bq.cdist[sample] : 0.0;
                    sm.x0[4 * Bt + tid] = cd;
                    bool pre_ok = isfinite(x[0]) && isfinite(x[1]) && isfinite(x[2]) && isfinite(x[3]);
                    for (int k = 0; k < T.kpre; ++k) {
                        const double v = T.Px[k * 4 + 0] * x[0] + T.Px[k * 4 + 1] * x[1] + T.Px[k * 4 + 2] * x[2] +
                                         T.Px[k * 4 + 3] * x[3] + T.Pc[k] * cd;
                        pre_ok = pre_ok && v <= T.pre_hi[k] && v >= T.pre_lo[k];
                    }
                    sm.slot_sample[tid] = sample;
                    sm.slot_state[tid] = pre_ok ? kSlotRunning : kSlotPreInfeasible;
                    sm.slot_iter[tid] = 0;
                    init = 1;
                } else {
                    init = sm.slot_state[tid] == kSlotIdle && sm.slot_sample[tid] == -2 ? 0 : 2;     // park once
                    sm.slot_sample[tid] = -2;
                    sm.slot_state[tid] = kSlotIdle;
                }
            }
            sm.slot_init[tid] = init;
            sm.red_res[tid] = 0; sm.red_nrm[tid] = 0; sm.red_sup[tid] = 0.f; sm.red_abs[tid] = 0.f;
        }
        __syncthreads();
        // (re)initialise the registers / tiles of the slots that changed hands
        int ini[S];
        bool any_init = false;
#pragma unroll
        for (int s = 0; s < S; ++s) { ini[s] = sm.slot_init[s0 + s]; any_init |= ini[s] != 0; }
        if (any_init) {
            double dx[S][4], xs[S][4], cd[S];
            int smp[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { xs[s][c] = sm.x0[c * Bt + s0 + s]; dx[s][c] = xs[s][c] - Bq.xref[c]; }
                cd[s] = sm.x0[4 * Bt + s0 + s];
                smp[s] = sm.slot_sample[s0 + s];
            }
#pragma unroll
            for (int g = 0; g < GB; ++g) {
                const int pg = g * kAdmmWarps + warp;
#pragma unroll
                for (int r = 0; r < kRB; ++r) {
                    const int i = pg * kRB + r;
                    const int rid = sm.row_id[i];
                    const float wd = sm.width[i];
                    const int vp = sm.vpos[i];
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!ini[s]) continue;
                        float h = 3.0e38f, w = 0.f;
                        if (ini[s] == 1 && rid >= 0) {
                            const double* gx = sm.gxs + i * 4;
                            h = (float)(sm.his[i] - gx[0] * xs[s][0] - gx[1] * xs[s][1] - gx[2] * xs[s][2] - gx[3] * xs[s][3] -
                                        sm.gcs[i] * cd[s]);
                            if (Bq.warm_in) w = Bq.warm[(size_t)smp[s] * T.mt + rid];
                        }
                        sm.HI[i * Bt + s0 + s] = h;
                        wB[g][r][s] = w;
                        if (vp >= 0) sm.V[vp * Bt + s0 + s] = rid >= 0 ? 2.f * clampf(w, h - wd, h) - w : 0.f;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < GA; ++g) {
                const int pg = g * kAdmmWarps + warp;
                if (pg >= T.nGA) continue;
#pragma unroll
                for (int r = 0; r < kRA; ++r) {
                    const int j = pg * kRA + r;
                    const int vid = sm.var_id[j];
                    const float lb = sm.lbs[j], ub = sm.ubs[j];
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (!ini[s]) continue;
                        float w = 0.f, x0v = 0.f;
                        if (ini[s] == 1 && vid >= 0) {
                            const double* kf = sm.kf + j * 4;
                            x0v = (float)(kf[0] * dx[s][0] + kf[1] * dx[s][1] + kf[2] * dx[s][2] + kf[3] * dx[s][3]);
                            if (Bq.warm_in) w = Bq.warm[(size_t)smp[s] * T.mt + T.m + vid];
                        }
                        wA[g][r][s] = w;
                        x0t[g][r][s] = x0v;
                        sm.V[(T.mv4 + j) * Bt + s0 + s] = (ini[s] == 1 && vid >= 0) ? 2.f * clampf(w, lb, ub) - w : 0.f;
                    }
                }
            }
        }
        int running = 0;
        if (tid < Bt) running = sm.slot_state[tid] == kSlotRunning || sm.slot_state[tid] == kSlotPreInfeasible;
        if (!__syncthreads_or(running)) break;

        // ================= check_every iterations, the last one with residuals =================
        for (int sub = 1; sub < T.check_every; ++sub) iteration(std::false_type{});
        iteration(std::true_type{});

        // ================= per-slot decisions =================
        if (tid < Bt) {
            const int st = sm.slot_state[tid];
            if (st == kSlotRunning || st == kSlotPreInfeasible) {
                const int it = sm.slot_iter[tid] + T.check_every;
                sm.slot_iter[tid] = it;
                const float res = __uint_as_float(sm.red_res[tid]), nrm = __uint_as_float(sm.red_nrm[tid]);
                const float sup = sm.red_sup[tid], sabs = sm.red_abs[tid];
                int ns = kSlotRunning;
                if (st == kSlotPreInfeasible) ns = kSlotInfeasible;
                // the margin eps_inf * sum |terms| dominates the float32 rounding of the sums (~1e-5 of it) and of the bounds
                else if (sabs > 0.f && sup <= -T.eps_inf * sabs) ns = kSlotInfeasible;
                else if (res <= eps_abs + eps_rel * nrm) ns = kSlotSolved;
                else if (it >= Bq.max_iter || !(res == res)) ns = kSlotMaxIter;
                sm.slot_state[tid] = ns;
            }
        }
        __syncthreads();
    }
}

template <int S, int H, int GA, int GB>
int launch_variant(QPHandle* q, const AdmmBatch& b, cudaStream_t st) {
    const size_t smem = admm_smem_bytes(q->host, S * H, q->host.mats_in_smem);
    const int64_t tiles = (b.count + 32 * S * H - 1) / (32 * S * H);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, q->sm));
    if (q->host.mats_in_smem) {
        if constexpr (GA >= 4) {
            set_error("admm_launch: matrices of this size class never fit shared memory");
            return CARMPC_ERR_UNSUPPORTED;
        } else {
            { const int rc = kernel_config(reinterpret_cast<const void*>(admm_kernel<S, H, GA, GB, true>), kAdmmThreads * H, smem, nullptr); if (rc != CARMPC_OK) return rc; }
            admm_kernel<S, H, GA, GB, true><<<blocks, kAdmmThreads * H, smem, st>>>(q->admm, b);
        }
    } else {
        { const int rc = kernel_config(reinterpret_cast<const void*>(admm_kernel<S, H, GA, GB, false>), kAdmmThreads * H, smem, nullptr); if (rc != CARMPC_OK) return rc; }
        admm_kernel<S, H, GA, GB, false><<<blocks, kAdmmThreads * H, smem, st>>>(q->admm, b);
    }
    CARMPC_CUDA(cudaGetLastError());
    return CARMPC_OK;
}

}  // namespace

int admm_launch(QPHandle* q, const AdmmBatch& b, cudaStream_t st) {
    if (b.count <= 0) return CARMPC_OK;
    if (admm_tc_usable(q, b)) {
        q->last_tc_samples += b.count;
        if (q->tensor_mode != 3) return admm_tc_launch(q, b, st);
        AdmmBatch bp = b;
        if (q->ws_prof == nullptr) CARMPC_CUDA(cudaMalloc(&q->ws_prof, sizeof(unsigned long long) * 16));
        CARMPC_CUDA(cudaMemsetAsync(q->ws_prof, 0, sizeof(unsigned long long) * 16, st));
        bp.prof = q->ws_prof;
        return admm_tc_launch(q, bp, st);
    }
    // Samples per tile: the size class fixes the maximum (registers / shared memory); a batch too small to give every
    // SM a full tile runs narrower tiles, whose iterations are proportionally shorter.
    const int smax = q->host.samples_per_lane;
    int S = 1;
    while (!b.narrow && S * 2 <= smax && (int64_t)b.count >= (int64_t)32 * (S * 2) * q->sm) S *= 2;
    const int key = q->host.ga_per_warp * 100 + q->host.gb_per_warp * 10 + S;
    switch (key) {
        // n <= 40: 16 warps as two sample-halves of 64 (measured equal to 8 warps x 4 samples per lane, with and without
        // FFMA2, fewer registers)
        case 124: return launch_variant<2, 2, 1, 2>(q, b, st);
        case 122: return launch_variant<2, 1, 1, 2>(q, b, st);
        case 121: return launch_variant<1, 1, 1, 2>(q, b, st);
        // n <= 80: 8 warps x 2 samples per lane (measured 20 % faster than 16 warps x 1)
        case 242: return launch_variant<2, 1, 2, 4>(q, b, st);
        case 241: return launch_variant<1, 1, 2, 4>(q, b, st);
        case 471: return launch_variant<1, 1, 4, 7>(q, b, st);
    }
    set_error("admm_launch: no kernel variant for this problem size");
    return CARMPC_ERR_UNSUPPORTED;
}

}  // namespace carmpc
