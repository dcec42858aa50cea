// Temporary placeholders for the QP half of the ABI (replaced by qp_admm.cu).
#include "common.cuh"
extern "C" {
void carmpc_qp_default_opts(carmpc_qp_opts* o) { if (o) { o->rho = 0.1; o->alpha = 1.6; o->eps_abs = 1e-3; o->eps_rel = 1e-3; o->eps_prim_inf = 1e-4; o->max_iter = 4000; o->check_every = 10; o->scaling_iters = 15; o->precise = 0; } }
int carmpc_qp_create(int, int, int, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const double*, const carmpc_qp_opts*, void**) { carmpc::set_error("qp: not built yet"); return CARMPC_ERR_UNSUPPORTED; }
int carmpc_qp_get_setup(void*, int, double*, int) { return CARMPC_ERR_UNSUPPORTED; }
int carmpc_qp_solve_batch(void*, const double*, const double*, const double*, int64_t, double*, double*, int32_t*, int32_t*, double*, float*, int, int, void*) { return CARMPC_ERR_UNSUPPORTED; }
int carmpc_qp_solve_host(void*, const double*, const double*, const double*, int64_t, double*, double*, int32_t*, int32_t*, double*) { return CARMPC_ERR_UNSUPPORTED; }
int carmpc_qp_last_stats(void*, int64_t*, int64_t*) { return CARMPC_ERR_UNSUPPORTED; }
int carmpc_closed_loop(void*, int, const double*, const double*, const double*, const double*, const double*, double, double, int, int, const double*, const double*, int64_t, double*, int32_t*, double*, double*, int64_t*, void*) { return CARMPC_ERR_UNSUPPORTED; }
}
