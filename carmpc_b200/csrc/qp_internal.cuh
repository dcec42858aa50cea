// Internal layout of the batched condensed-QP solver (shared by qp_setup.cu, qp_admm.cu, qp_polish.cu, qp_api.cu and
// closed_loop.cu).  Nothing here is part of the C ABI.
//
// Problem (include/carmpc.h):   min 1/2 u'Hu + (F (x0 - xref))'u
//                               lo - Gx x0 - Gc c <= G u <= hi - Gx x0 - Gc c ,   lb <= u <= ub
// One instance per initial state x0; H, F, G, Gx and the bounds are shared by the whole batch, which is what
// makes the ADMM linear operator one dense matrix applied to a (variables x samples) tile.
//
// Device algorithm (two kernels per pass):
//   1. admm_kernel   float32 OSQP-style ADMM on the Ruiz-equilibrated problem, in the single-vector form
//                        c = clip(w) ; x~ = P (2c - w) + x~0 ; z = A_s x~ ; w += alpha (z - c)
//                    with P = rho K^-1 A_s' precomputed on the host (K = H_s + rho A_s'A_s is shared).  It only has
//                    to identify the active set / prove infeasibility.
//   2. polish_kernel float64 active-set solve through the Schur complement  (A_act H^-1 A_act') lambda = ... with
//                    the shared H^-1, A H^-1, A H^-1 A' precomputed; verifies primal feasibility and multiplier
//                    signs (a KKT certificate), repairs the set a few times, and produces u to ~1e-10.
#pragma once

#include <atomic>
#include <vector>

#include "common.cuh"

namespace carmpc {

constexpr int kRA = 5;            // rows of x~ per thread tile in stage A
constexpr int kRB = 7;            // constraint rows per thread tile in stage B
constexpr int kAdmmWarps = 8;
constexpr int kAdmmThreads = kAdmmWarps * 32;
constexpr int kMaxN = 160;        // variables (2 x horizon 80)
constexpr int kMaxM = 392;        // general rows (8 warps x 7 groups x 7 rows)
constexpr int kPolishMaxActive = 64;
constexpr int kPolishStats = 20;       // counters of carmpc_qp_polish_stats
constexpr int kFirstPassIters = 100;  // iteration cap of the first ADMM pass (stragglers continue in the second pass)

enum SlotState : int { kSlotIdle = -1, kSlotRunning = 0, kSlotSolved = 1, kSlotInfeasible = 2, kSlotMaxIter = 3,
                       kSlotPreInfeasible = 4 };

// status codes between the kernels (the ABI codes are CARMPC_QP_*)
constexpr int kStatusNeedsMoreAdmm = 3;      // polish could not certify: re-run ADMM tighter

struct AdmmTables {
    // matrices (float32)
    const float* P;        // [nA_rows][ktot]      rho K^-1 [Gs' | diag(lam)], columns in V order
    const float* Gs;       // [m_phys][npad4]      scaled general rows, physical (owner) order, permuted variables
    const float* GsT;      // [nA_rows][mv4]       Gs', columns in V order (certificate pass)
    // per physical general row
    const double* his;     // [m_phys]   Eg * hi        (pad rows: 3e38)
    const double* Gxs;     // [m_phys][4]
    const double* Gcs;     // [m_phys]
    const float* width;    // [m_phys]   Eg * (hi - lo), +inf for one-sided and pad rows
    const float* Einv_g;   // [m_phys]   1 / Eg         (pad rows: 0)
    const float* Esc_g;    // [m_phys]   Eg             (pad rows: 0)
    const int* vpos;       // [m_phys]   row of V this constraint row writes
    const int* row_id;     // [m_phys]   logical row index, -1 for pad rows
    // per permuted variable
    const float* lam;      // [nA_rows]  Eb * D (scaled box row), 0 for pad
    const float* lbs;      // [nA_rows]
    const float* ubs;      // [nA_rows]
    const float* Einv_b;   // [nA_rows]
    const float* Esc_b;    // [nA_rows]  Eb
    const float* Dinv;     // [nA_rows]
    const float* Dsc;      // [nA_rows]  D (u = D x~)
    const double* KF;      // [nA_rows][4]   x~0 = KF (x0 - xref)
    const int* var_id;     // [nA_rows]  logical variable index, -1 for pad
    // per group of rows
    const int4* segA;      // [nA_rows / kRA]   V ranges (general beg, end, box beg, end), multiples of 4
    const int2* segB;      // [m_phys / kRB]    x~ range (beg, end), multiples of 4
    // rows that do not depend on u: pre_lo <= Px x0 + Pc c <= pre_hi
    const double* Px;
    const double* Pc;
    const double* pre_lo;
    const double* pre_hi;
    int kpre;
    int n, m, mt;                          // logical sizes, mt = m + n
    int nA_rows, m_phys, npad4, mv4, ktot; // padded sizes
    int nGA, nGB;                          // number of row groups in stage A / B
    float rho, alpha, eps_abs, eps_rel, eps_inf;
    int check_every;
};

struct AdmmBatch {
    const double* x0;        // SoA: x0[c * stride + sample]
    int64_t stride;
    const double* cdist;     // nullable, per sample
    double xref[4];
    const int* idx_list;     // nullable: sample = idx_list[q]
    int count;               // number of queue entries
    const int* count_dev;    // nullable: the number of entries lives on the device (count is then an upper bound)
    int narrow;              // host side: always run the narrowest tile (device-sized batches of unknown, usually small, size)
    int* next;               // global work counter (zeroed before launch)
    int8_t* sign;            // [batch][mt]   +1 upper active, -1 lower active
    float* u_admm;           // [batch][n]    unscaled iterate (fallback when the polish cannot certify)
    int* status;
    int* iters;
    float* warm;             // nullable, [batch][mt] scaled w
    int warm_in, warm_out;
    unsigned long long* total_iters;
    float eps_scale;
    int max_iter;
    int iters_accumulate;    // second pass: add to the iteration count of the first
    int write_u;             // also write the ADMM iterate (needed when the polish may pass it through)
    int combine;             // tensor-core kernel, second and later chains of a split problem: merge status / iterations
                             // with what the earlier chains wrote (infeasible if any chain is, solved if all are)
    unsigned long long* prof; // nullable (tensor-core kernel, tensor mode 2): cycle counters summed over the CTAs
};

// Tensor-core form of the ADMM (qp_admm_tc.cu): a tile of 128 samples is the M dimension of tcgen05.mma kind::tf32, the
// shared matrices are the B operands (K-major, 128-byte swizzle, split into a TF32 "hi" image and a TF32 "lo" residual
// image for the 3xTF32 products), the ADMM state of the general rows and the accumulators live in tensor memory.
// Everything is in logical order (live general rows, then variables); pads are inert (zero matrix rows / columns).
//   product 0   x~ (np)   = [V_b (np) | e (16) | V^_g (mp)] . B0'      K0 = np + 16 + mp
//   product 1   z^ (mp)   = [x~  (np) | e (16)] . B1'                  K1 = np + 16
//   product 2   Gs'dy (np) = [dy (mp)] . B2'                           K2 = mp          (certificate, checked iterations)
// e = per-sample constants as exact TF32 pieces: x0_c as hi / mid / lo (c = 0..3), 1, the disturbance c as hi / mid / lo.
// General rows are kept shifted by their per-sample upper bound h = his - Gxs x0 - Gcs c (w^ = w - h, z^ = z - h), which
// makes their clip bounds (-width, 0) sample-independent; the shift enters both products through the e columns.
struct TcTables {
    const unsigned char* img;   // chunk images: chunk k of product p at img + off[p] + k * pair_bytes[p] (hi image, then lo)
    int off[3], pair_bytes[3], nchunks[3], ksteps[3], ncols[3];
    const float* nwd;           // [mp]   -width (scaled; -inf: one-sided or pad row)
    const float* einv_g;        // [mp]   1 / Eg (pads 0)
    const float* hisf;          // [mp]   float copies of his, Gxs, Gcs for the residual norms / support sums (pads 0)
    const float* gxsf;          // [mp][4]
    const float* gcsf;          // [mp]
    const double* his;          // [mp]   (pads 0)
    const double* gxs;          // [mp][4]
    const double* gcs;          // [mp]
    const int* row_id;          // [mp]   logical row, -1 pad
    const float* lam;           // [np]   Eb D (pads 0)
    const float* lb;            // [np]   scaled box (pads -inf / +inf)
    const float* ub;
    const float* einv_b;        // [np]
    const float* nrl;           // [np]   -1 / lam (pads 0)
    const double* kfv;          // [np][4]  x~0 = kfv (x0 - xref)
    const int* var_id;          // [np]   logical variable, -1 pad
    int n, m, mt, np, mp;
    int resident;               // products 0 and 1 stay in shared memory (loaded once); product 2 always streams
    int merged;                 // B_hi and B_lo are consumed as ONE operand of 2 N rows (A_hi fetched once per k-step: two
                                // MMAs instead of three); needs 2 N accumulator columns per product: 3 mp + 2 np <= 512
    int na_stages, nb_stages, b_stage_bytes, resident_bytes;
    int smem_bytes;
    int ok;                     // 0: this problem has no tensor-core form (too large for tensor memory / shared memory)
};

// host images of one part of the tensor-core form (the whole problem, or one independent chain of it)
struct TcPart {
    TcTables t;
    std::vector<unsigned char> img;
    std::vector<float> nwd, einv_g, hisf, gxsf, gcsf, lam, lb, ub, einv_b, nrl;
    std::vector<double> his, gxs, gcs, kfv;
    std::vector<int> row_id, var_id;
};

struct PolishTables {
    const double* H;         // [n][n]
    const double* Hinv;      // [n][n]
    const double* F;         // [n][4]
    const double* Uu;        // [n][4]      u_unc = Uu (x0 - xref),  Uu = -H^-1 F
    const double* AUu;       // [mt][4]     [G; I] Uu
    const double* AH;        // [mt][n]     [G; I] H^-1
    const double* AHA;       // [mt][mt]    [G; I] H^-1 [G; I]'
    const double* Gx;        // [m][4]
    const double* Gc;        // [m]
    const double* hi;        // [mt]  (general rows then box)
    const double* lo;        // [mt]
    const double* G;         // [m][n]      general rows (rows bounded only from below negated, as hi / lo)
    const double* Eg;        // [m]         row scaling of the ADMM (its dual iterate is in scaled units)
    int n, m, mt;
};

struct PolishBatch {
    const double* x0;
    int64_t stride;
    const double* cdist;
    double xref[4];
    const int* idx_list;
    int count;
    const int* count_dev;    // nullable: the number of list entries lives on the device (count is then an upper bound)
    const int8_t* sign;
    const float* u_admm;
    int* status;             // in: 0 / 2 from ADMM -> out: 0 solved (polished or accepted), 3 needs more ADMM
    double* u0;              // SoA 2 x stride (nullable)
    double* objective;       // nullable
    double* u_full;          // [batch][n] nullable
    int8_t* polished;        // nullable
    int* n_failed;           // device counter of samples set to kStatusNeedsMoreAdmm
    int* failed_list;        // their indices
    int final_pass;          // 1: accept the ADMM iterate when the polish cannot certify
    int rounds;              // repair rounds (-1: default; 0: no polish, pass the ADMM iterate through)
    int na_cap;              // active rows this launch has shared memory for
    int* overflow_list;      // samples with more (nullable: they count as not certified)
    int* n_overflow;
    // active-set reuse (closed loop: the previous step's certified set is tried before any ADMM iteration)
    int precheck;            // 1: also evaluate the u-independent rows / finiteness the ADMM slot refill checks
    const double* Px;        // [kpre][4]
    const double* Pc;        // [kpre]
    const double* pre_lo;
    const double* pre_hi;
    int kpre;
    int8_t* sign_out;        // nullable: the certified active set is written back ([batch][mt])
    int* iters_out;          // nullable: certified samples get 0 iterations
    // active-set seeding (region-of-attraction maps: neighbouring states mostly share their active set)
    const int* seed;         // nullable: sample i starts from the certified set of sample seed[i] (an anchor solved before)
    // multiplier maps of the anchors' critical regions: lambda(x0) = Lam [x0; 1] on the anchor's active rows
    const int* rec_of;       // [batch] record of a sample (-1: none)
    double* rec_lam;         // [records][32][5]
    int* rec_act;            // [records][33]   na (-1: no map), then the active rows in the order of Lam's rows
    int rec_write, rec_read;
    unsigned long long* stats;   // nullable [20]: [r] samples certified after r repair rounds (r = 0..9), [10] not certified,
                                 // [11] round 0 taken from a multiplier map, [12] certified by a multiplier map alone,
                                 // [13] proven infeasible by the anchor's Farkas certificate (or the u-independent rows),
                                 // [14] max_iter samples proven infeasible by the certificate of their own ADMM state,
                                 // [15] samples proven infeasible the same way before the second pass
                                 // [16] samples the final polish could not certify (handed to the float64 fallback),
                                 // [17] of those solved + certified, [18] proven infeasible, [19] left undecided (status 2)
};

// float64 images of the equilibrated problem for the fallback solver (qp_exact.cu), logical order
struct ExactTables {
    const double* Gs;        // [m][n]   Eg G D (rows without a finite bound: zero)
    const double* GsT;       // [n][m]
    const double* Kinv;      // [n][n]   (Hs + rho (Gs'Gs + lam^2))^-1
    const double* lam;       // [n]      Eb D
    const double* Eb;        // [n]
    const double* D;         // [n]
    const double* rowmin;    // [m]      min over the input box of G_i u  (-inf when a needed box side is infinite)
    const double* rowmax;    // [m]      max over the input box of G_i u
    const double* rowabs;    // [m]      sum_j |G_ij| max(|lb_j|, |ub_j|): rounding scale of the two
    double cs, rho, alpha;
};

// Host-side setup product (qp_setup.cu)
struct QPHost {
    int n = 0, m = 0, kpre = 0;
    carmpc_qp_opts opts;
    // scaling (logical order)
    std::vector<double> D, Eg, Eb, Kinv, Gs64;
    double cscale = 1.0;
    // padded device images
    AdmmTables geo;          // sizes only (pointers filled by the handle)
    std::vector<float> P, Gs, GsT, width, Einv_g, Esc_g, lam, lbs, ubs, Einv_b, Esc_b, Dinv, Dsc;
    std::vector<double> his, Gxs, Gcs, KF, Px, Pc, pre_lo, pre_hi;
    std::vector<int> vpos, row_id, var_id;
    std::vector<int4> segA;
    std::vector<int2> segB;
    // polish (logical order, unscaled)
    std::vector<double> H, Hinv, F, G, Uu, AUu, AH, AHA, Gx, Gc, hi, lo;
    // tensor-core form: one part for the whole problem, or one per independent chain (sizes in t; pointers filled by the handle)
    TcTables tc;                     // = tc_parts[0].t
    std::vector<TcPart> tc_parts;
    int ga_per_warp = 0, gb_per_warp = 0, samples_per_lane = 0;
    bool mats_in_smem = false;
    size_t smem_bytes = 0;
    double flops_per_iter = 0;      // executed FFMA * 2 per sample-iteration (structure-aware)
    double flops_per_iter_dense = 0;
};

int qp_host_setup(int n, int m, int k, const double* H, const double* F, const double* G, const double* Gx,
                  const double* Gc, const double* lo, const double* hi, const double* lb, const double* ub,
                  const double* Px, const double* Pc, const double* pre_lo, const double* pre_hi,
                  const carmpc_qp_opts& opts, QPHost* out);

struct QPHandle;
struct QPBusyGuard {
    std::atomic<bool>* flag;
    bool acquired;
    explicit QPBusyGuard(std::atomic<bool>& f) : flag(&f) { bool expect = false; acquired = f.compare_exchange_strong(expect, true); }
    ~QPBusyGuard() { if (acquired) flag->store(false); }
};
int admm_launch(QPHandle* q, const AdmmBatch& b, cudaStream_t st);
// tcgen05 form of the same iteration (qp_admm_tc.cu); admm_launch picks it for large first passes
int admm_tc_launch(QPHandle* q, const AdmmBatch& b, cudaStream_t st);
bool admm_tc_usable(const QPHandle* q, const AdmmBatch& b);
int polish_launch(QPHandle* q, const PolishBatch& b, cudaStream_t st);
// the empty active set tried first (one thread per sample); samples it does not settle are appended to d_rest
int polish_unconstrained_launch(QPHandle* q, const PolishBatch& b, int* d_rest, int* d_n_rest, cudaStream_t st);
// infeasible anchors of a seeded map: turn the ADMM dual iterate into an exact Farkas certificate, affine in x0
int farkas_export_launch(QPHandle* q, const int* d_anchors, int count, const int* d_status, const float* d_warm,
                         const double* d_x0, int64_t stride, cudaStream_t st);
// before the second pass: samples the certificate of their first-pass state proves infeasible are done, the rest is listed
int farkas_filter_launch(QPHandle* q, const int* d_list, int count, int* d_status, const float* d_warm, const double* d_x0,
                         int64_t stride, double* d_u0, double* d_objective, double* d_u_full, int8_t* d_polished,
                         int* d_survivors, int* d_n_survivors, cudaStream_t st);
// samples that ran out of ADMM iterations: a valid certificate from their final ADMM state makes them proven infeasible
int farkas_decide_launch(QPHandle* q, const int* d_list, int count, int* d_status, const float* d_warm, const double* d_x0,
                         int64_t stride, double* d_u0, double* d_objective, double* d_u_full, int8_t* d_polished,
                         cudaStream_t st, const int* d_count = nullptr);
// ADMM verdicts "infeasible" are only accepted with a float64 certificate: the others re-enter the second pass
int farkas_verify_launch(QPHandle* q, const int* d_list, int count, int* d_status, const float* d_warm, const double* d_x0,
                         int64_t stride, const double* d_c, const double* xref, int* d_failed, int* d_n_failed, cudaStream_t st,
                         const int* d_count = nullptr);
// float64 fallback for what the second pass could not prove (qp_exact.cu)
int exact_fallback(QPHandle* q, const PolishBatch& pb_final, const int* d_list, int count, float* d_warm, int* d_iters,
                   cudaStream_t st, int* h_handled, const int* d_count = nullptr);
size_t admm_smem_bytes(const QPHost& h, int samples_per_lane, bool mats_in_smem);

struct QPHandle : HandleBase {
    QPHost host;
    AdmmTables admm;
    TcTables tc;                                 // = tc_parts[0]
    std::vector<TcTables> tc_parts;              // device tables of every part (one, or one per independent chain)
    int tensor_mode = 1;                         // 0: FFMA kernel only; 1: tcgen05 kernel where it is faster; 2: wherever possible; 3: 2 + cycle counters
    int64_t last_tc_samples = 0;                 // samples the tcgen05 kernel took in the last solve
    unsigned long long* ws_prof = nullptr;       // [16] cycle counters of the tcgen05 kernel (tensor mode 2)
    PolishTables polish;
    ExactTables exact;
    std::vector<void*> allocations;
    // per-batch workspace (grown on demand)
    int64_t ws_batch = 0;
    int8_t* ws_sign = nullptr;
    float* ws_u = nullptr;
    int* ws_status = nullptr;
    int* ws_iters = nullptr;
    int* ws_failed = nullptr;
    int* ws_failed0 = nullptr;                   // samples whose reused active set did not certify
    int* ws_rest = nullptr;                      // samples the empty active set does not settle
    int* ws_anchor = nullptr;                    // seeded solve: samples solved cold (anchors) / from a seed (followers)
    int* ws_follow = nullptr;
    int* ws_rec_of = nullptr;                    // [batch] record index of an anchor
    double* ws_rec_lam = nullptr;                // multiplier maps of the anchors (grown on demand, capped)
    int* ws_rec_act = nullptr;
    int64_t rec_cap = 0;
    int use_records = 0;                         // set by solve_seeded for the two solve() calls it makes
    unsigned long long* ws_polish_stats = nullptr;   // [kPolishStats] histogram of the polish launches of the last solve
    int stats_hold = 0;
    int defer_total = 0;                         // closed loop: ws_total_iters accumulates over the steps, read once at the end
    int* ws_overflow = nullptr;
    int* ws_unproven = nullptr;                  // samples the second pass left without a proof (float64 fallback)
    int* ws_counters = nullptr;                  // [0] work counter, [1] n_failed
    unsigned long long* ws_total_iters = nullptr;
    int8_t* ws_polished = nullptr;
    float* ws_warm = nullptr;                    // ADMM state of every sample (warm start of the second pass)
    // device side of the host-buffer entry point (grown on demand, kept across calls)
    int64_t io_cap = 0, io_full_cap = 0;
    double *io_x0_aos = nullptr, *io_x0 = nullptr, *io_c = nullptr, *io_u0 = nullptr, *io_u0_aos = nullptr, *io_obj = nullptr,
           *io_full = nullptr;
    int32_t *io_status = nullptr, *io_iters = nullptr, *io_seed = nullptr;
    double* io_axes = nullptr;                   // grid axes of carmpc_qp_map_host
    int ensure_io(int64_t batch, bool want_full);
    int64_t last_total_iters = 0, last_launches = 0, last_second_pass = 0, last_fallback = 0;
    int sm = 148;
    bool host_only = false;
    // per-call state lives in the handle (workspace, counters): one call at a time.  A second call entering while one is
    // in flight (another host thread) is refused instead of corrupting both.
    std::atomic<bool> busy{false};
    ~QPHandle() override;
    int ensure_workspace(int64_t batch);
    // full solve on device buffers (all pointers device; any output may be null except status)
    int solve(const double* d_x0, int64_t stride, const double* xref, const double* d_c, const int* d_idx, int64_t count,
              double* d_u0, double* d_objective, int32_t* d_status, int32_t* d_iters, double* d_u_full, float* d_warm,
              int warm_in, int warm_out, cudaStream_t st, int reuse_active_set = 0, const int* d_seed = nullptr,
              int keep_sign = 0);
    // The same pipeline for a warm-started sequence (closed loop), with every sample count left on the device: the list
    // d_idx holds *d_count <= max_count entries, the previous step's certified active sets are tried first, and nothing
    // is read back - the call only enqueues (no host synchronisation), so consecutive steps run back to back on the GPU.
    int solve_enqueue(const double* d_x0, int64_t stride, const double* xref, const int* d_idx, const int* d_count,
                      int64_t max_count, double* d_u0, int32_t* d_status, float* d_warm, cudaStream_t st);
    // anchors (seed[i] == i or out of range) cold, then every other sample from its anchor's certified active set
    int solve_seeded(const double* d_x0, int64_t batch, const double* xref, const double* d_c, const int* d_seed,
                     double* d_u0, double* d_objective, int32_t* d_status, int32_t* d_iters, double* d_u_full,
                     cudaStream_t st);
    int64_t last_reused = 0;                     // samples certified from the reused active set in the last solve
    int64_t last_anchors = 0;                    // seeded solve: samples solved cold
};

}  // namespace carmpc
