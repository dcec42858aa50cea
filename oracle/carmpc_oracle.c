/*
 * CPU oracle, C part: terminal-set membership and its rollout form.
 *
 * TEST INFRASTRUCTURE ONLY - never linked into or loaded by the product (carmpc_b200/).  Used by tests/ as
 * the bit-exact checker of the CUDA kernels and by bench.py as the timed CPU baseline (all host threads).
 *
 * Restates lib/terminal_set.py:107-113 (np.all(A @ point <= b) for every grid point) and the sampled form
 * of lib/terminal_set.py:53-59, 198-200 (constraint rows applied along the closed-loop LQR rollout).
 * The dot product is evaluated as   fma(a3, v, fma(a2, psi, fma(a1, y, a0 * x)))   in IEEE float64, the
 * order the CUDA kernels use, so bitsets can be compared with ==.  numpy's BLAS may associate the four
 * products differently; oracle/carmpc_oracle.py keeps the numpy expression and the tests enumerate the
 * (measure-zero) samples within 1e-6 of a facet where the two could disagree.
 *
 * Pinned by: tests/golden/grid_config1.npz (434 members, 49 exact ties) and the shipped terminal_sets.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

/* ---- a minimal pthread parallel-for over 32-sample words (this gcc ships without libgomp) ------------ */
typedef int64_t (*range_fn)(void* ctx, int64_t w0, int64_t w1);
typedef struct { range_fn fn; void* ctx; int64_t w0, w1, result; } job_t;

static void* job_main(void* arg) {
    job_t* j = (job_t*)arg;
    j->result = j->fn(j->ctx, j->w0, j->w1);
    return NULL;
}

int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static int64_t parallel_words(range_fn fn, void* ctx, int64_t words, int threads) {
    if (threads <= 0) threads = oracle_max_threads();
    if (threads > 256) threads = 256;
    if ((int64_t)threads > words) threads = words > 0 ? (int)words : 1;
    if (threads == 1) return fn(ctx, 0, words);
    pthread_t tid[256];
    job_t jobs[256];
    const int64_t per = (words + threads - 1) / threads;
    int64_t total = 0;
    for (int t = 0; t < threads; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].w0 = t * per < words ? t * per : words;
        jobs[t].w1 = (t + 1) * per < words ? (t + 1) * per : words;
        jobs[t].result = 0;
        pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) { pthread_join(tid[t], NULL); total += jobs[t].result; }
    return total;
}

/* bit i of bits[] = sample i is inside {A x <= b}; returns the member count.  threads <= 0: all. */
typedef struct {
    const double *Ab, *x, *y, *psi, *v;
    int rows; int64_t n; uint32_t* bits;
} member_ctx;

static int64_t membership_range(void* vctx, int64_t w0, int64_t w1) {
    const member_ctx* c = (const member_ctx*)vctx;
    const double *Ab = c->Ab, *x = c->x, *y = c->y, *psi = c->psi, *v = c->v;
    const int rows = c->rows; const int64_t n = c->n; uint32_t* bits = c->bits;
    int64_t count = 0;
    for (int64_t w = w0; w < w1; ++w) {
        uint32_t word = 0;
        const int64_t base = w * 32;
        const int lim = (int)((n - base) < 32 ? (n - base) : 32);
        for (int k = 0; k < lim; ++k) {
            const int64_t i = base + k;
            const double X = x[i], Y = y[i], P = psi[i], V = v[i];
            int in = 1;
            for (int r = 0; r < rows && in; ++r) {
                const double* a = Ab + 5 * r;
                const double rr = fma(a[3], V, fma(a[2], P, fma(a[1], Y, a[0] * X)));
                in = rr <= a[4];
            }
            word |= (uint32_t)in << k;
            count += in;
        }
        bits[w] = word;
    }
    return count;
}

int64_t oracle_membership(const double* Ab, int rows, const double* x, const double* y, const double* psi,
                          const double* v, int64_t n, uint32_t* bits, int threads) {
    member_ctx c = {Ab, x, y, psi, v, rows, n, bits};
    return parallel_words(membership_range, &c, (n + 31) / 32, threads);
}

/* e_0 = p - goal, e_{t+1} = Ak e_t; state rows for t = 0..k_steps, input rows at t = 0 (mode 0) or always. */
typedef struct {
    const double *Ak, *Acon, *bcon, *Ain, *bin, *goal, *x, *y, *psi, *v;
    int s, rin, k_steps, input_mode; int64_t n; uint32_t* bits; int32_t* first_violation;
} rollout_ctx;

static int64_t rollout_range(void* vctx, int64_t w0, int64_t w1) {
    const rollout_ctx* c = (const rollout_ctx*)vctx;
    const double *Ak = c->Ak, *Acon = c->Acon, *bcon = c->bcon, *Ain = c->Ain, *bin = c->bin, *goal = c->goal;
    const double *x = c->x, *y = c->y, *psi = c->psi, *v = c->v;
    const int s = c->s, rin = c->rin, k_steps = c->k_steps, input_mode = c->input_mode;
    const int64_t n = c->n; uint32_t* bits = c->bits; int32_t* first_violation = c->first_violation;
    int64_t count = 0;
    for (int64_t w = w0; w < w1; ++w) {
        uint32_t word = 0;
        const int64_t base = w * 32;
        const int lim = (int)((n - base) < 32 ? (n - base) : 32);
        for (int k = 0; k < lim; ++k) {
            const int64_t i = base + k;
            double e0 = x[i] - goal[0], e1 = y[i] - goal[1], e2 = psi[i] - goal[2], e3 = v[i] - goal[3];
            int in = 1, first = -1;
            for (int t = 0; t <= k_steps && in; ++t) {
                int ok = 1;
                for (int r = 0; r < s; ++r) {
                    const double* a = Acon + 4 * r;
                    ok &= fma(a[3], e3, fma(a[2], e2, fma(a[1], e1, a[0] * e0))) <= bcon[r];
                }
                if (t == 0 || input_mode == 1) {
                    for (int r = 0; r < rin; ++r) {
                        const double* a = Ain + 4 * r;
                        ok &= fma(a[3], e3, fma(a[2], e2, fma(a[1], e1, a[0] * e0))) <= bin[r];
                    }
                }
                if (!ok) { first = t; in = 0; }
                const double n0 = fma(Ak[3], e3, fma(Ak[2], e2, fma(Ak[1], e1, Ak[0] * e0)));
                const double n1 = fma(Ak[7], e3, fma(Ak[6], e2, fma(Ak[5], e1, Ak[4] * e0)));
                const double n2 = fma(Ak[11], e3, fma(Ak[10], e2, fma(Ak[9], e1, Ak[8] * e0)));
                const double n3 = fma(Ak[15], e3, fma(Ak[14], e2, fma(Ak[13], e1, Ak[12] * e0)));
                e0 = n0; e1 = n1; e2 = n2; e3 = n3;
            }
            word |= (uint32_t)in << k;
            count += in;
            if (first_violation) first_violation[i] = first;
        }
        bits[w] = word;
    }
    return count;
}

int64_t oracle_rollout(const double* Ak, const double* Acon, const double* bcon, int s, const double* Ain,
                       const double* bin, int rin, const double* goal, int k_steps, int input_mode,
                       const double* x, const double* y, const double* psi, const double* v, int64_t n,
                       uint32_t* bits, int32_t* first_violation, int threads) {
    rollout_ctx c = {Ak, Acon, bcon, Ain, bin, goal, x, y, psi, v, s, rin, k_steps, input_mode, n, bits, first_violation};
    return parallel_words(rollout_range, &c, (n + 31) / 32, threads);
}

/* The reference's per-point form, kept scalar and unfused on purpose: sum_j a_j p_j left to right with
 * separately rounded products (what a plain loop over `A @ point` does without BLAS). */
int64_t oracle_membership_plain(const double* Ab, int rows, const double* pts /* n x 4 */, int64_t n, uint8_t* out) {
    int64_t count = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double* p = pts + 4 * i;
        int in = 1;
        for (int r = 0; r < rows; ++r) {
            const double* a = Ab + 5 * r;
            volatile double s = a[0] * p[0];
            s = s + a[1] * p[1];
            s = s + a[2] * p[2];
            s = s + a[3] * p[3];
            in &= s <= a[4];
        }
        out[i] = (uint8_t)in;
        count += in;
    }
    return count;
}
