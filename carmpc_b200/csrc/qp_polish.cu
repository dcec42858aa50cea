// Float64 active-set polish of the batched QP (second kernel of a pass; see qp_internal.cuh).
//
// Given the active-set guess of the float32 ADMM (sign per constraint row: +1 at its upper bound, -1 at its lower), one
// warp per sample solves the equality-constrained QP exactly through the Schur complement on the shared matrices
//       u = u_unc - (A H^-1)'_act lambda ,   (A_act H^-1 A_act' + delta I) lambda = A_act u_unc - b_act ,   u_unc = -H^-1 q
// then checks the KKT conditions in float64 (every row within its bounds, multiplier signs right).  A wrong guess is
// repaired (drop rows with a wrong-signed multiplier, add the most violated row) a few times.  What leaves this kernel
// with status 0 and polished = 1 is a certified optimum of the reference's QP (lib/mpc.py:321-332) to ~1e-10.
#include <math.h>

#include <algorithm>

#include "qp_internal.cuh"

namespace carmpc {

namespace {

constexpr int kPolishRounds = 8;
constexpr double kFeasTol = 1e-8;
constexpr double kSignTol = 1e-9;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct PolishSmemLayout {
    int na_max, ldm;
    size_t per_warp;
};

__host__ __device__ inline PolishSmemLayout polish_layout(int n, int mt) {
    PolishSmemLayout L;
    L.na_max = n + 8 < kPolishMaxActive ? n + 8 : kPolishMaxActive;
    if (L.na_max > mt) L.na_max = mt;
    if (L.na_max < 1) L.na_max = 1;
    L.ldm = L.na_max | 1;                                    // odd leading dimension: conflict-free column walks
    size_t d = 0;
    d += 3 * (size_t)n;                                      // q, u_unc, u
    d += 3 * (size_t)mt;                                     // Au_unc, hi, lo
    d += (size_t)L.na_max * L.ldm;                           // M
    d += 3 * (size_t)L.na_max;                               // rhs / lambda, b, diag0
    size_t bytes = d * sizeof(double);
    bytes += sizeof(int) * (size_t)L.na_max;                 // act
    bytes += (size_t)((mt + 15) & ~15);                      // sgn
    L.per_warp = (bytes + 15) & ~size_t(15);
    return L;
}

__global__ void __launch_bounds__(256) polish_kernel(const PolishTables T, const PolishBatch B, int warps_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_per_cta) return;
    const int n = T.n, m = T.m, mt = T.mt;
    const PolishSmemLayout L = polish_layout(n, mt);
    unsigned char* base = smem_raw + (size_t)warp * L.per_warp;
    double* q = reinterpret_cast<double*>(base);
    double* uunc = q + n;
    double* u = uunc + n;
    double* Auu = u + n;
    double* hi = Auu + mt;
    double* lo = hi + mt;
    double* M = lo + mt;
    double* rhs = M + (size_t)L.na_max * L.ldm;
    double* bact = rhs + L.na_max;
    double* diag0 = bact + L.na_max;
    int* act = reinterpret_cast<int*>(diag0 + L.na_max);
    signed char* sgn = reinterpret_cast<signed char*>(act + L.na_max);
    const double NaN = __longlong_as_double(0x7ff8000000000000ll);

    const int gw = blockIdx.x * warps_per_cta + warp, nw = gridDim.x * warps_per_cta;
    for (int qi = gw; qi < B.count; qi += nw) {
        const int sample = B.idx_list ? B.idx_list[qi] : qi;
        const int st_in = B.status[sample];
        __syncwarp();
        if (st_in == CARMPC_QP_INFEASIBLE) {
            if (lane == 0) {
                if (B.u0) { B.u0[sample] = NaN; B.u0[B.stride + sample] = NaN; }
                if (B.objective) B.objective[sample] = INFINITY;
                if (B.polished) B.polished[sample] = 0;
            }
            if (B.u_full) for (int j = lane; j < n; j += 32) B.u_full[(size_t)sample * n + j] = NaN;
            continue;
        }
        double x0[4], dx[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { x0[c] = B.x0[(size_t)c * B.stride + sample]; dx[c] = x0[c] - B.xref[c]; }
        const double cd = B.cdist ? B.cdist[sample] : 0.0;
        for (int j = lane; j < n; j += 32) {
            const double* f = T.F + (size_t)j * 4;
            q[j] = f[0] * dx[0] + f[1] * dx[1] + f[2] * dx[2] + f[3] * dx[3];
        }
        for (int i = lane; i < mt; i += 32) {
            double shift = 0.0;
            if (i < m) {
                const double* gx = T.Gx + (size_t)i * 4;
                shift = gx[0] * x0[0] + gx[1] * x0[1] + gx[2] * x0[2] + gx[3] * x0[3] + T.Gc[i] * cd;
            }
            hi[i] = T.hi[i] - shift;
            lo[i] = T.lo[i] - shift;
            sgn[i] = B.sign[(size_t)sample * mt + i];
        }
        __syncwarp();
        for (int j = lane; j < n; j += 32) {                 // u_unc = -H^-1 q
            double s = 0;                                     // Hinv is symmetric: walk it column-wise (coalesced)
            for (int k = 0; k < n; ++k) s += T.Hinv[(size_t)k * n + j] * q[k];
            uunc[j] = -s;
        }
        __syncwarp();
        for (int i = lane; i < mt; i += 32) {                // A u_unc
            double s;
            if (i < m) {
                s = 0;
                for (int k = 0; k < n; ++k) s += T.GT[(size_t)k * m + i] * uunc[k];
            } else {
                s = uunc[i - m];
            }
            Auu[i] = s;
        }
        __syncwarp();

        bool certified = false;
        const int max_rounds = B.rounds < 0 ? kPolishRounds : B.rounds;
        for (int round = 0; round < max_rounds && !certified; ++round) {
            // ---- compact the active set ----
            int na = 0;
            for (int i0 = 0; i0 < mt; i0 += 32) {
                const int i = i0 + lane;
                const bool on = i < mt && sgn[i] != 0;
                const unsigned mask = __ballot_sync(0xffffffffu, on);
                const int p = na + __popc(mask & ((1u << lane) - 1u));
                if (on && p < L.na_max) act[p] = i;
                na += __popc(mask);
            }
            if (na > L.na_max) break;
            __syncwarp();
            // ---- M = AHA[act, act], rhs = A_act u_unc - b_act ----
            for (int e = lane; e < na * na; e += 32) {
                const int a = e / na, b = e - a * na;
                M[a * L.ldm + b] = T.AHA[(size_t)act[a] * mt + act[b]];
            }
            for (int a = lane; a < na; a += 32) {
                const int i = act[a];
                bact[a] = sgn[i] > 0 ? hi[i] : lo[i];
                rhs[a] = Auu[i] - bact[a];
            }
            __syncwarp();
            for (int a = lane; a < na; a += 32) { diag0[a] = M[a * L.ldm + a]; M[a * L.ldm + a] += 1e-13 * diag0[a]; }
            __syncwarp();
            // ---- Cholesky with pivot skipping (dependent active rows get lambda = 0) ----
            for (int j = 0; j < na; ++j) {
                const double d = M[j * L.ldm + j];
                const bool skip = !(d > 1e-11 * diag0[j]);
                const double piv = skip ? 1.0 : sqrt(d);
                __syncwarp();
                if (lane == 0) M[j * L.ldm + j] = skip ? 0.0 : piv;      // 0 on the diagonal marks a skipped pivot
                for (int i = j + 1 + lane; i < na; i += 32) M[i * L.ldm + j] = skip ? 0.0 : M[i * L.ldm + j] / piv;
                __syncwarp();
                if (!skip)
                    for (int i = j + 1 + lane; i < na; i += 32) {
                        const double lij = M[i * L.ldm + j];
                        for (int k = j + 1; k <= i; ++k) M[i * L.ldm + k] -= lij * M[k * L.ldm + j];
                    }
                __syncwarp();
            }
            // ---- L y = rhs, L' lambda = y ----
            for (int j = 0; j < na; ++j) {
                const double d = M[j * L.ldm + j];
                const double y = d > 0.0 ? rhs[j] / d : 0.0;
                __syncwarp();
                if (lane == 0) rhs[j] = y;
                for (int i = j + 1 + lane; i < na; i += 32) rhs[i] -= M[i * L.ldm + j] * y;
                __syncwarp();
            }
            for (int j = na - 1; j >= 0; --j) {
                const double d = M[j * L.ldm + j];
                const double x = d > 0.0 ? rhs[j] / d : 0.0;
                __syncwarp();
                if (lane == 0) rhs[j] = x;
                for (int i = lane; i < j; i += 32) rhs[i] -= M[j * L.ldm + i] * x;
                __syncwarp();
            }
            // rhs now holds lambda (signed: positive pushes against an upper bound)
            // ---- u = u_unc - (AH)'_act lambda ----
            for (int j = lane; j < n; j += 32) {
                double s = uunc[j];
                for (int a = 0; a < na; ++a) s -= T.AH[(size_t)act[a] * n + j] * rhs[a];
                u[j] = s;
            }
            // ---- KKT check ----
            double worst = 0.0;
            int worst_i = -1, worst_sign = 0;
            for (int i = lane; i < mt; i += 32) {
                if (isinf(hi[i]) && isinf(lo[i])) continue;
                double s = Auu[i];
                for (int a = 0; a < na; ++a) s -= T.AHA[(size_t)act[a] * mt + i] * rhs[a];
                const double vu = s - hi[i], vl = lo[i] - s;
                const double v = fmax(vu, vl);
                if (v > worst) { worst = v; worst_i = i; worst_sign = vu >= vl ? 1 : -1; }
            }
            const double wmax = warp_max(worst);
            const unsigned who = __ballot_sync(0xffffffffu, worst == wmax && worst_i >= 0);
            int add_i = -1, add_sign = 0;
            if (wmax > kFeasTol && who) {
                const int src = __ffs(who) - 1;
                add_i = __shfl_sync(0xffffffffu, worst_i, src);
                add_sign = __shfl_sync(0xffffffffu, worst_sign, src);
            }
            double lam_max = 0.0;
            for (int a = lane; a < na; a += 32) lam_max = fmax(lam_max, fabs(rhs[a]));
            lam_max = warp_max(lam_max);
            int n_bad = 0;
            for (int a = lane; a < na; a += 32) {
                const int i = act[a];
                if (rhs[a] * (double)sgn[i] < -kSignTol * (1.0 + lam_max)) { sgn[i] = 0; ++n_bad; }
            }
            n_bad = __reduce_add_sync(0xffffffffu, n_bad);
            __syncwarp();
            if (add_i < 0 && n_bad == 0) { certified = true; break; }
            if (add_i >= 0 && lane == 0) sgn[add_i] = (signed char)add_sign;
            __syncwarp();
        }

        if (!certified) {
            if (!B.final_pass) {
                if (lane == 0) {
                    B.status[sample] = kStatusNeedsMoreAdmm;
                    const int slot = atomicAdd(B.n_failed, 1);
                    B.failed_list[slot] = sample;
                }
                continue;
            }
            for (int j = lane; j < n; j += 32) u[j] = (double)B.u_admm[(size_t)sample * n + j];
            __syncwarp();
        }
        // ---- outputs: u(0), objective 1/2 u'Hu + q'u, full sequence ----
        double part = 0.0;
        for (int j = lane; j < n; j += 32) {
            double s = 0;                                     // H is symmetric
            for (int k = 0; k < n; ++k) s += T.H[(size_t)k * n + j] * u[k];
            part += u[j] * (0.5 * s + q[j]);
        }
        const double obj = warp_sum(part);
        if (lane == 0) {
            if (B.u0) { B.u0[sample] = u[0]; B.u0[B.stride + sample] = n > 1 ? u[1] : 0.0; }
            if (B.objective) B.objective[sample] = obj;
            if (B.polished) B.polished[sample] = certified ? 1 : 0;
            if (certified) B.status[sample] = CARMPC_QP_SOLVED;
        }
        if (B.u_full) for (int j = lane; j < n; j += 32) B.u_full[(size_t)sample * n + j] = u[j];
    }
}

}  // namespace

int polish_launch(QPHandle* qh, const PolishBatch& b, cudaStream_t st) {
    if (b.count <= 0) return CARMPC_OK;
    const PolishSmemLayout L = polish_layout(qh->polish.n, qh->polish.mt);
    int warps = (int)std::min<size_t>(8, (size_t)(200 * 1024) / L.per_warp);
    if (warps < 1) { set_error("polish: shared-memory budget exceeded"); return CARMPC_ERR_UNSUPPORTED; }
    const size_t smem = L.per_warp * warps;
    CARMPC_CUDA(cudaFuncSetAttribute(polish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t need = (b.count + warps - 1) / warps;
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)qh->sm * 4));
    polish_kernel<<<blocks, 256, smem, st>>>(qh->polish, b, warps);
    CARMPC_CUDA(cudaGetLastError());
    return CARMPC_OK;
}

}  // namespace carmpc
