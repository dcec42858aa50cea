"""Would the two ADMM tile products hold up on the tensor pipe?  (north_star: "tensor cores only if the tolerance holds")

Re-runs the kernel's arithmetic (tests/admm_emulation.py, float32, same tables the library exports) with the operands of
the two products rounded the way tcgen05.mma kind::tf32 consumes them (10-bit mantissa, low 13 bits ignored), in three
variants: plain float32 (the FFMA kernel), 1xTF32, and the 3xTF32 split (hi*hi + lo*hi + hi*lo, float32 accumulate).
For every variant it reports, on random config-3 states after a fixed iteration budget, how many samples the float64
polish certifies from the ADMM's active set (what matters: the polish, not the ADMM, produces the reported solution)
and how far the ADMM iterate is from the float32 one.  CPU only; run from the repo root."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def tf32(a):
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


def tf32_rn(a):
    b = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    b = (b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    return b.view(np.float32)


def product(V, M, mode):
    """V (B, k) @ M (r, k)' with float32 accumulation."""
    if mode == "fp32":
        return V @ M.T
    Vh, Mh = tf32(V), tf32(M)
    if mode == "tf32":
        return Vh @ Mh.T
    if mode == "2xtf32":
        # data split exactly (hi + lo), matrix rounded to nearest TF32 once on the host: a FIXED perturbation of the operator
        Mr = tf32_rn(M)
        return Vh @ Mr.T + tf32(V - Vh) @ Mr.T
    Vl, Ml = tf32(V - Vh), tf32(M - Mh)
    return Vh @ Mh.T + (Vl @ Mh.T + Vh @ Ml.T)


def run(T, x0, xref, iters, mode):
    f32 = np.float32
    B = len(x0)
    dx = x0 - xref[None, :]
    x0t = (dx @ T.KF.T).astype(f32)
    hi = np.where(T.row_id[None, :] >= 0, T.his[None, :] - x0 @ T.Gxs.T, 3.0e38).astype(f32)
    with np.errstate(invalid="ignore", over="ignore"):
        lo = hi - T.width[None, :]
    wB = np.zeros((B, T.mp), f32); wA = np.zeros((B, T.nA), f32); V = np.zeros((B, T.ktot), f32)
    live = T.row_id >= 0
    vp = T.vpos
    V[:, vp[live]] = (2 * np.clip(wB, lo, hi) - wB)[:, live]
    V[:, T.mv4:T.mv4 + T.nA] = np.where(T.var_id[None, :] >= 0, 2 * np.clip(wA, T.lbs, T.ubs) - wA, 0)
    # dense masked operators (structural zeros stay exact zeros in every mode)
    Pm = np.zeros_like(T.P); Gm = np.zeros_like(T.Gs)
    for ga in range(T.nGA):
        r0 = ga * T.RA; gb, ge, bb, be = T.segA[ga]
        Pm[r0:r0 + T.RA, gb:ge] = T.P[r0:r0 + T.RA, gb:ge]; Pm[r0:r0 + T.RA, bb:be] = T.P[r0:r0 + T.RA, bb:be]
    for g in range(T.nGB):
        r0 = g * T.RB; kb, ke = T.segB[g]
        Gm[r0:r0 + T.RB, kb:ke] = T.Gs[r0:r0 + T.RB, kb:ke]
    res = None
    for it in range(iters):
        xt = (x0t + product(V, Pm, mode)).astype(f32)
        z = T.lam[None, :] * xt[:, :T.nA]
        c0 = np.clip(wA, T.lbs, T.ubs); wA = wA + T.alpha * (z - c0); c1 = np.clip(wA, T.lbs, T.ubs)
        V[:, T.mv4:T.mv4 + T.nA] = 2 * c1 - wA
        xpad = np.zeros((B, T.npad4), f32); xpad[:, :xt.shape[1]] = xt[:, :T.npad4]
        zB = product(xpad, Gm, mode).astype(f32)
        c0 = np.clip(wB, lo, hi); wB = wB + T.alpha * (zB - c0); c1 = np.clip(wB, lo, hi)
        V[:, vp[live]] = (2 * c1 - wB)[:, live]
        res = np.abs(zB - c1)[:, live].max(1)
    u = np.zeros((B, T.n)); sign = np.zeros((B, T.m + T.n), np.int8)
    vid = T.var_id; ok = vid >= 0
    u[:, vid[ok]] = (T.Dsc[None, :] * xt[:, :T.nA])[:, ok]
    sign[:, T.m + vid[ok]] = ((wA > T.ubs).astype(np.int8) - (wA < T.lbs).astype(np.int8))[:, ok]
    sign[:, T.row_id[live]] = ((wB > hi).astype(np.int8) - (wB < lo).astype(np.int8))[:, live]
    return u, sign, res


def main():
    from conftest import make_env, make_controller
    from admm_emulation import KernelTables, polish_reference
    from carmpc_b200.batch import BatchQP
    from oracle import carmpc_oracle as orc
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    nstates = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    c = make_controller(make_env("RoadOneCarEnv", [29.9, 1.5, 0, 0]), N)
    bq = BatchQP.from_controller(c)
    T = KernelTables(bq)
    goal = np.array(c.goal, float)
    rng = np.random.default_rng(3)
    lo = np.array([5.0, -3.0, -np.pi / 8, -1.0]); hi = np.array([30.0, 3.0, np.pi / 8, 5.0])
    x0 = lo + rng.uniform(size=(nstates * 2, 4)) * (hi - lo)
    Ab = np.load(os.path.join(ROOT, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    oq = orc.CondensedQP("RoadOneCarEnv", N, Ab)
    feas, _ = orc.qp_feasible_lp(oq, x0)
    x0 = x0[feas][:nstates]
    print(f"N = {N}: {len(x0)} feasible config-3 states, tables n={T.n} m={T.m} ktot={T.ktot}")
    for iters in (30, 60):
        base = None
        for mode in ("fp32", "tf32", "2xtf32", "3xtf32"):
            u, sign, res = run(T, x0, goal, iters, mode)
            cert0 = cert8 = 0
            for i in range(len(x0)):
                cert0 += polish_reference(bq.pq, x0[i], goal, sign[i], rounds=1)[1]
                cert8 += polish_reference(bq.pq, x0[i], goal, sign[i], rounds=8)[1]
            if base is None:
                base = (u, sign)
            du = np.abs(u - base[0]).max()
            same = (sign == base[1]).all(1).mean()
            print(f"  iters {iters:3d} {mode:7s}: residual median {np.median(res):.2e} max {res.max():.2e} | active set == fp32's "
                  f"{same:6.1%} | max |u - u_fp32| {du:.2e} | polish certifies as-is {cert0 / len(x0):6.1%}, "
                  f"after <= 8 repairs {cert8 / len(x0):6.1%}")


if __name__ == "__main__":
    main()
