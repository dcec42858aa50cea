// Monte-Carlo closed loops: very many independent runs of  plant step -> (observer) -> MPC QP  in lock-step.
//
// Order of operations is the reference's loop (examples/run_MPCOutputFB.py:29-41, run_MPCStateFB.py:29-39): the plant
// (lib/simulator.py:51-69, forward Euler of the kinematic bicycle) moves first with the previous input ([0, 0] at the
// first step), then the controller steps on the new output: Luenberger observer (lib/mpc.py:448) for output feedback,
// the state itself for state feedback, then one condensed QP per run (lib/mpc.py:461-478 / :318-335).  A run whose QP is
// infeasible stops there, as the reference raises OutsideTheRegionOfAttractionError.
//
// One thread per run for plant + observer (4 state and 2 input registers); the QPs of all live runs go through the
// batched solver with the previous step's ADMM state as warm start.
#include <math.h>
#include <string.h>

#include <stdlib.h>

#include "qp_internal.cuh"

namespace carmpc {

namespace {

struct LoopConst {
    double A[16], B[8], C[12], L[12];
    double dt, l1;
    int mode;            // 0 state feedback, 1 output feedback
};

// plant step with the previous input, observer update, estimate -> QP input; builds the list of live runs
__global__ void __launch_bounds__(256)
loop_advance_kernel(const LoopConst K, int64_t runs, double* __restrict__ x, double* __restrict__ xhat,
                    const double* __restrict__ u_prev, const int32_t* __restrict__ fail_step,
                    double* __restrict__ est, int* __restrict__ live_list, int* __restrict__ live_count,
                    double* __restrict__ traj_step, const int* __restrict__ step_dev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= runs) return;
    // replayed steps (CUDA graph) read the step number on the device: traj_step is then the base of the trajectory
    if (traj_step != nullptr && step_dev != nullptr) traj_step += (size_t)*step_dev * 4 * runs;
    double s[4], u[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) s[c] = x[c * runs + i];
    const bool alive = fail_step[i] < 0;
    if (alive) {
        u[0] = u_prev[i];
        u[1] = u_prev[runs + i];
        // state + dt * [v cos(psi), v sin(psi), v / l1 * tan(delta), a]      (rate * dt + state, as the reference)
        const double psi = s[2], v = s[3];
        const double r0 = v * cos(psi), r1 = v * sin(psi), r2 = v / K.l1 * tan(u[1]), r3 = u[0];
        s[0] = r0 * K.dt + s[0];
        s[1] = r1 * K.dt + s[1];
        s[2] = r2 * K.dt + s[2];
        s[3] = r3 * K.dt + s[3];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c * runs + i] = s[c];
        double e[4];
        if (K.mode == 1) {
            double xh[4], innov[3];
#pragma unroll
            for (int c = 0; c < 4; ++c) xh[c] = xhat[c * runs + i];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double y = 0, yh = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) { y += K.C[r * 4 + c] * s[c]; yh += K.C[r * 4 + c] * xh[c]; }
                innov[r] = y - yh;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                double t = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) t += K.A[r * 4 + c] * xh[c];
                t += K.B[r * 2 + 0] * u[0] + K.B[r * 2 + 1] * u[1];
#pragma unroll
                for (int c = 0; c < 3; ++c) t += K.L[r * 3 + c] * innov[c];
                e[r] = t;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) xhat[c * runs + i] = e[c];
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) e[c] = s[c];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) est[c * runs + i] = e[c];
        live_list[atomicAdd(live_count, 1)] = (int)i;
    }
    if (traj_step != nullptr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) traj_step[c * runs + i] = s[c];
    }
}

// take the QP results of the live runs: new input, or mark the run as failed at this step
__global__ void __launch_bounds__(256)
loop_apply_kernel(int64_t runs, int step, const int* __restrict__ live_list, int live, const int* __restrict__ live_dev,
                  const double* __restrict__ u0, const int32_t* __restrict__ status, double* __restrict__ u_prev,
                  int32_t* __restrict__ fail_step, const int* __restrict__ step_dev) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (live_dev != nullptr) live = min(live, *live_dev);
    if (step_dev != nullptr) step = *step_dev;
    if (q >= live) return;
    const int i = live_list[q];
    if (status[i] == CARMPC_QP_SOLVED) {
        u_prev[i] = u0[i];
        u_prev[runs + i] = u0[runs + i];
    } else {
        fail_step[i] = step;
    }
}

// replayed steps: the input log of this step, then the step number moves on (one thread, after everything that read it)
__global__ void __launch_bounds__(256)
loop_log_kernel(int64_t runs, const double* __restrict__ u_prev, double* __restrict__ u_log, const int* __restrict__ step_dev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * runs) return;
    u_log[(size_t)*step_dev * 2 * runs + i] = u_prev[i];
}
__global__ void loop_next_kernel(int* step_dev) { *step_dev += 1; }
__global__ void loop_set_kernel(int* step_dev, int v) { *step_dev = v; }

__global__ void fill_i32(int32_t* p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace
}  // namespace carmpc

using namespace carmpc;

extern "C" int carmpc_closed_loop(void* qp, int mode, const double* h_A, const double* h_B, const double* h_C,
                                  const double* h_L, const double* h_xref, double dt, double l1, int steps,
                                  int warm_start, const double* d_x_init, const double* d_xhat_init, int64_t runs,
                                  double* d_final, int32_t* d_fail_step, double* d_traj, double* d_u_log,
                                  int64_t* h_total_iters, void* stream) {
    QPHandle* q = check_handle<QPHandle>(qp, kQP);
    CARMPC_REQUIRE(q != nullptr, "not a QP handle");
    CARMPC_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (state feedback) or 1 (output feedback)");
    CARMPC_REQUIRE(h_A && h_B && h_xref, "null model pointer");
    CARMPC_REQUIRE(mode == 0 || (h_C && h_L), "output feedback needs C and L");
    CARMPC_REQUIRE(steps >= 0 && runs >= 0 && runs < (int64_t)1 << 31, "steps / runs");
    CARMPC_REQUIRE(warm_start >= 0 && warm_start <= 2, "warm_start must be 0, 1 or 2");
    CARMPC_REQUIRE(dt > 0 && l1 > 0, "dt, l1");
    if (h_total_iters) *h_total_iters = 0;
    if (runs == 0) return CARMPC_OK;
    CARMPC_REQUIRE(d_x_init && d_final && d_fail_step, "null device pointer");
    QPBusyGuard guard(q->busy);
    CARMPC_REQUIRE(guard.acquired, "this QP handle is in use by another call (one call per handle at a time)");
    // The loop runs on its own (capturable) stream, ordered after the caller's stream by an event and joined at the end by
    // the final synchronisation: the legacy default stream cannot be captured into a graph.
    cudaStream_t user_st = (cudaStream_t)stream, st = nullptr;
    cudaEvent_t ev = nullptr;
    bool graph_ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    if (graph_ok) graph_ok = cudaEventRecord(ev, user_st) == cudaSuccess && cudaStreamWaitEvent(st, ev, 0) == cudaSuccess;
    if (!graph_ok) { cudaGetLastError(); if (st) cudaStreamDestroy(st); st = user_st; }
    LoopConst K;
    memset(&K, 0, sizeof(K));
    memcpy(K.A, h_A, sizeof(K.A));
    memcpy(K.B, h_B, sizeof(K.B));
    if (h_C) memcpy(K.C, h_C, sizeof(K.C));
    if (h_L) memcpy(K.L, h_L, sizeof(K.L));
    K.dt = dt; K.l1 = l1; K.mode = mode;

    const int mt = q->admm.mt;
    double *x = d_final, *xhat = nullptr, *est = nullptr, *u_prev = nullptr, *u0 = nullptr;
    int32_t* status = nullptr;
    int *live_list = nullptr, *live_count = nullptr;
    float* warm = nullptr;
    int rc = CARMPC_OK;
    auto cleanup = [&]() { if (st != user_st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } if (ev) cudaEventDestroy(ev);
                           q->defer_total = 0; cudaFree(xhat); cudaFree(est); cudaFree(u_prev); cudaFree(u0); cudaFree(status); cudaFree(live_list); cudaFree(live_count); cudaFree(warm); };
#define TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s: %s", #call, cudaGetErrorString(e__)); cleanup(); return CARMPC_ERR_CUDA; } } while (0)
    TRY(cudaMalloc(&xhat, sizeof(double) * 4 * runs));
    TRY(cudaMalloc(&est, sizeof(double) * 4 * runs));
    TRY(cudaMalloc(&u_prev, sizeof(double) * 2 * runs));
    TRY(cudaMalloc(&u0, sizeof(double) * 2 * runs));
    TRY(cudaMalloc(&status, sizeof(int32_t) * runs));
    TRY(cudaMalloc(&live_list, sizeof(int) * runs));
    TRY(cudaMalloc(&live_count, sizeof(int) * 2));      // [0] live runs of the step, [1] step number of replayed steps
    if (warm_start) TRY(cudaMalloc(&warm, sizeof(float) * (size_t)mt * runs));
    if (x != d_x_init) TRY(cudaMemcpyAsync(x, d_x_init, sizeof(double) * 4 * runs, cudaMemcpyDeviceToDevice, st));
    TRY(cudaMemcpyAsync(xhat, d_xhat_init ? d_xhat_init : d_x_init, sizeof(double) * 4 * runs, cudaMemcpyDeviceToDevice, st));
    TRY(cudaMemsetAsync(u_prev, 0, sizeof(double) * 2 * runs, st));
    const int blocks = (int)((runs + 255) / 256);
    fill_i32<<<blocks, 256, 0, st>>>(d_fail_step, runs, -1);
    int64_t total_iters = 0;
    bool warm_valid = false;
    // the iteration total accumulates on the device over the steps (one read at the end instead of one per step)
    rc = q->ensure_workspace(runs);
    if (rc != CARMPC_OK) { cleanup(); return rc; }
    TRY(cudaMemsetAsync(q->ws_total_iters, 0, sizeof(unsigned long long), st));
    q->defer_total = 1;
    // One enqueue-only step (every count stays on the device).  step_dev == nullptr: the step number k comes from the host.
    auto enqueue_step = [&](int k, const int* step_dev) -> int {
        cudaError_t e = cudaMemsetAsync(live_count, 0, sizeof(int), st);
        if (e != cudaSuccess) { set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return CARMPC_ERR_CUDA; }
        loop_advance_kernel<<<blocks, 256, 0, st>>>(K, runs, x, xhat, u_prev, d_fail_step, est, live_list, live_count,
                                                    d_traj ? (step_dev ? d_traj : d_traj + (size_t)k * 4 * runs) : nullptr, step_dev);
        const int rc2 = q->solve_enqueue(est, runs, h_xref, live_list, live_count, runs, u0, status, warm, st);
        if (rc2 != CARMPC_OK) return rc2;
        loop_apply_kernel<<<blocks, 256, 0, st>>>(runs, k, live_list, (int)runs, live_count, u0, status, u_prev, d_fail_step, step_dev);
        if (d_u_log) {
            if (step_dev) {
                loop_log_kernel<<<(int)((2 * runs + 255) / 256), 256, 0, st>>>(runs, u_prev, d_u_log, step_dev);
            } else {
                e = cudaMemcpyAsync(d_u_log + (size_t)k * 2 * runs, u_prev, sizeof(double) * 2 * runs, cudaMemcpyDeviceToDevice, st);
                if (e != cudaSuccess) { set_error("cudaMemcpyAsync: %s", cudaGetErrorString(e)); return CARMPC_ERR_CUDA; }
            }
        }
        if (step_dev) loop_next_kernel<<<1, 1, 0, st>>>(live_count + 1);
        return CARMPC_OK;
    };
    for (int k = 0; k < steps; ++k) {
        if (warm_valid && warm_start != 2) {
            // Every later step only enqueues: the number of live runs, of runs whose active set changed, of second-pass and
            // fallback samples all stay on the device (kernels sized for the upper bound read them there), so the steps
            // run back to back on the GPU without a host round trip.  The first such step is launched kernel by kernel
            // (it also fills the kernel-attribute caches); the remaining ones are ONE captured CUDA graph replayed with
            // the step number on the device, so that a slow host (thousands of small launches otherwise) cannot stretch
            // the loop.
            if (!graph_ok || k + 2 >= steps || k < 2) {
                rc = enqueue_step(k, nullptr);
                if (rc != CARMPC_OK) { q->defer_total = 0; cleanup(); return rc; }
                continue;
            }
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            loop_set_kernel<<<1, 1, 0, st>>>(live_count + 1, k);
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                // no capture possible here (e.g. the caller is capturing itself): launch the steps one by one
                cudaGetLastError();
                graph_ok = false;
                rc = enqueue_step(k, nullptr);
                if (rc != CARMPC_OK) { q->defer_total = 0; cleanup(); return rc; }
                continue;
            }
            rc = enqueue_step(k, live_count + 1);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc == CARMPC_OK && ce != cudaSuccess) { set_error("cudaStreamEndCapture: %s", cudaGetErrorString(ce)); rc = CARMPC_ERR_CUDA; }
            if (rc == CARMPC_OK && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { set_error("cudaGraphInstantiate failed"); rc = CARMPC_ERR_CUDA; }
            for (; rc == CARMPC_OK && k < steps; ++k)
                if (cudaGraphLaunch(exec, st) != cudaSuccess) { set_error("cudaGraphLaunch failed"); rc = CARMPC_ERR_CUDA; }
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            if (rc != CARMPC_OK) { cudaGetLastError(); q->defer_total = 0; cleanup(); return rc; }
            break;
        }
        TRY(cudaMemsetAsync(live_count, 0, sizeof(int), st));
        loop_advance_kernel<<<blocks, 256, 0, st>>>(K, runs, x, xhat, u_prev, d_fail_step, est, live_list, live_count,
                                                    d_traj ? d_traj + (size_t)k * 4 * runs : nullptr, nullptr);
        int live = 0;
        TRY(cudaMemcpyAsync(&live, live_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(st));
        if (live > 0) {
            // from the second step on, the active set certified at the previous step is tried first
            rc = q->solve(est, runs, h_xref, nullptr, live_list, live, u0, nullptr, status, nullptr, nullptr, warm,
                          warm_valid ? 1 : 0, warm ? 1 : 0, st, warm_valid ? 1 : 0);
            if (rc != CARMPC_OK) { q->defer_total = 0; cleanup(); return rc; }
            warm_valid = warm != nullptr;
            loop_apply_kernel<<<(live + 255) / 256, 256, 0, st>>>(runs, k, live_list, live, nullptr, u0, status, u_prev, d_fail_step, nullptr);
        }
        if (d_u_log)      // the input each run will apply at the next plant step (unchanged for stopped runs)
            TRY(cudaMemcpyAsync(d_u_log + (size_t)k * 2 * runs, u_prev, sizeof(double) * 2 * runs, cudaMemcpyDeviceToDevice, st));
    }
    q->defer_total = 0;
    unsigned long long total_dev = 0;
    TRY(cudaMemcpyAsync(&total_dev, q->ws_total_iters, sizeof(total_dev), cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    total_iters = (int64_t)total_dev;
    TRY(cudaGetLastError());
#undef TRY
    cleanup();
    if (h_total_iters) *h_total_iters = total_iters;
    return CARMPC_OK;
}
