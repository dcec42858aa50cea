"""Generate golden vectors from the UNMODIFIED reference (``/root/reference``).

Test infrastructure only.  Run in the authoring container (the reference tree does not exist
on the GPU box); the outputs are committed under ``tests/golden/`` and are what the tests read.

    python tests/golden/gen_golden.py

The reference imports cvxpy / polytope / matplotlib at module top (``lib/mpc.py:4``,
``lib/terminal_set.py:3-7``, ``lib/environments.py:1-3``); none is installed here.  Dummy modules
are registered for them so that every constructor and matrix builder runs unmodified
(SURVEY.md appendix B).  Nothing that *calls* cvxpy / polytope is executed.

What is recorded (all float64, straight from reference objects):
  * model.npz        A, B, P, K, Q, R, L, C, input bounds                 (lib/mpc.py:53-93, 387-404)
  * predmod_N*.npz   T, S, H, h, const for N in {1, 5, 10, 20}; for N in {40, 80} a digest
                     (selected rows / columns, traces, sums)               (lib/matrix_gen.py:6-72)
  * constraints_<env>_N*.npz  terminal / input / state constraint stacks  (lib/mpc.py:196-253)
  * simulator.npz    CarSimulator trajectories                            (lib/simulator.py:51-118)
  * observer.npz     MPCOutputFB.luenberger_observer sequence             (lib/mpc.py:439-448)
  * lqr.npz          MPC(use_LQR=True).step outputs                       (lib/mpc.py:255-276)
  * in_adm_set.npz   algorithm_1 / algorithm_2 on small inputs            (lib/in_adm_set.py:4-77)
  * terminal_sets/*.npy   byte copies of the reference's shipped H-rep fixtures (data, not source)
  * grid_config1.npz membership of the lib/terminal_set.py:96-113 grid, evaluated with the
                     reference's own expression ``np.all(A @ point <= b)``
  * qp_<env>_N*.npz  everything lib/mpc.py:318-332 (MPCStateFB.step) / :461-478 (MPCOutputFB.step) puts into
                     the QP, straight from the reference controller object: H, h, T, S, the three constraint
                     stacks, goal.  The BASELINE configs: RoadOneCarEnv N = 10/20/40/80 (configs 3 and 5),
                     RoadEnv N = 20 (config 4), RoadMultipleCarsEnv N = 20.  tests/kkt_check.py verifies
                     solver outputs against these with no oracle solver in the loop
  * disturbance.npz  ABd of MPCOutputFBWithDisturbance.step, computed by the reference's own expression
                     (lib/mpc.py:631-635) for N in {1, 5, 20}, with Bd, Cd, L1, L2 (lib/mpc.py:536-549)
"""
import os
import shutil
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return _Stub("call")


def _import_reference():
    for name in ["cvxpy", "polytope", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "matplotlib.lines", "matplotlib.colors", "matplotlib.transforms"]:
        sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)
    os.chdir(os.path.join(REF, "examples"))  # '../terminal_sets/' is relative (lib/mpc.py:98)


def main():
    _import_reference()
    from lib.mpc import MPC, MPCStateFB, MPCOutputFB
    from lib.environments import RoadEnv, RoadOneCarEnv, RoadMultipleCarsEnv
    from lib.simulator import CarSimulator
    from lib.in_adm_set import algorithm_1, algorithm_2
    from lib import configuration as cfg

    def ctrl(env, N, **kw):
        return MPCStateFB(dt=cfg.DT_CONTROL, N=N, lin_state=cfg.LINEARIZE_STATE,
                          lin_input=cfg.LINEARIZE_INPUT, env=env, **kw)

    # ---- model ------------------------------------------------------------------------------
    env = RoadOneCarEnv()
    env.set_goal([29.9, 1.5, 0, 0])
    c = ctrl(env, 20)
    ofb = MPCOutputFB(dt=cfg.DT_CONTROL, N=20, lin_state=cfg.LINEARIZE_STATE,
                      lin_input=cfg.LINEARIZE_INPUT, init_state=[5, -1.5, 0, 0], env=RoadEnv())
    np.savez(os.path.join(OUT, "model.npz"), A=c.A, B=c.B, P=c.P, K=c.K, Q=c.Q, R=c.R,
             L=ofb.L, C=ofb.C, y_goal=ofb.y_goal, input_upper=c.input_upper,
             input_lower=c.input_lower)

    # ---- prediction / cost matrices -----------------------------------------------------------
    for N in (1, 5, 10, 20):
        cN = ctrl(env, N, terminal_constraint=False)
        from lib.matrix_gen import costgen
        H, h, const = costgen(cN.Q, cN.R, cN.P, cN.T, cN.S, 4)
        np.savez_compressed(os.path.join(OUT, f"predmod_N{N}.npz"), T=cN.T, S=cN.S, H=H, h=h,
                            const=const)
    for N in (40, 80):
        cN = ctrl(env, N, terminal_constraint=False)
        np.savez_compressed(
            os.path.join(OUT, f"predmod_digest_N{N}.npz"),
            T_last=cN.T[-4:], S_last=cN.S[-4:], S_col0=cN.S[:, :2], H_row0=cN.H[0], H_diag=np.diag(cN.H),
            H_last=cN.H[-1], h_first=cN.h[:2], h_last=cN.h[-2:], H_trace=np.trace(cN.H),
            H_sum=cN.H.sum(), h_sum=cN.h.sum(), S_sum=cN.S.sum(), T_sum=cN.T.sum())

    # ---- constraint stacks ---------------------------------------------------------------------
    envs = {"RoadEnv": (RoadEnv, None), "RoadOneCarEnv": (RoadOneCarEnv, [29.9, 1.5, 0, 0]),
            "RoadOneCarEnvDefault": (RoadOneCarEnv, None),
            "RoadMultipleCarsEnv": (RoadMultipleCarsEnv, None)}
    for tag, (cls, goal) in envs.items():
        for N in (1, 3, 20):
            e = cls()
            if goal is not None:
                e.set_goal(goal)
            cN = ctrl(e, N)
            At, bt = cN.terminal_constraint()
            Ai, bi = cN.input_constraint()
            As, bs = cN.state_constraint()
            np.savez_compressed(os.path.join(OUT, f"constraints_{tag}_N{N}.npz"), At=At, bt=bt, Ai=Ai,
                                bi=bi, As=As, bs=bs, goal=np.array(e.goal, dtype=float),
                                env_A=np.array(e.constraints_A, dtype=float),
                                env_b=np.array(e.constraints_b, dtype=float))
    # missing terminal-set file -> fallback terminal constraint (lib/mpc.py:107-117)
    e = RoadEnv()
    e.set_goal([10, 0, 0, 0])
    cN = ctrl(e, 4)
    At, bt = cN.terminal_constraint()
    np.savez(os.path.join(OUT, "constraints_fallback_N4.npz"), At=At, bt=bt)

    # ---- simulator ------------------------------------------------------------------------------
    rng = np.random.default_rng(0)
    sims = {}
    for k, (dt, clip) in enumerate([(0.2, False), (0.01, True)]):
        sim = CarSimulator(dt=dt, clip=clip, C=ofb.C)
        x0 = np.array([5.0, -1.5, 0.1, 2.0])
        sim.reset(x0.copy())
        us = rng.uniform([-2.0, -0.7], [2.0, 0.7], size=(25, 2)) if not clip else \
            rng.uniform([-3.0, -1.2], [3.0, 1.2], size=(25, 2))
        states, outs = [], []
        for u in us:
            sim.step(u)
            states.append(np.array(sim.state))
            outs.append(np.array(sim.output))
        sims[f"x0_{k}"] = x0
        sims[f"u_{k}"] = us
        sims[f"states_{k}"] = np.array(states)
        sims[f"outputs_{k}"] = np.array(outs)
        sims[f"dt_{k}"] = dt
        sims[f"time_{k}"] = sim.time
    np.savez(os.path.join(OUT, "simulator.npz"), **sims)

    # ---- observer (lib/mpc.py:448), driven open-loop with recorded u --------------------------------
    xh = [np.array(ofb.x_estimate, dtype=float)]
    ys = rng.uniform([0, -2, 0], [30, 2, 4], size=(12, 3))
    us = rng.uniform([-2, -0.3], [2, 0.3], size=(12, 2))
    for y, u in zip(ys, us):
        ofb.previous_u = u
        ofb.x_estimate = ofb.luenberger_observer(y)
        xh.append(np.array(ofb.x_estimate))
    np.savez(os.path.join(OUT, "observer.npz"), y=ys, u=us, xhat=np.array(xh))

    # ---- LQR step -----------------------------------------------------------------------------------
    lq = MPC(dt=cfg.DT_CONTROL, N=20, lin_state=cfg.LINEARIZE_STATE, lin_input=cfg.LINEARIZE_INPUT,
             env=RoadMultipleCarsEnv(), use_LQR=True)
    xs = rng.uniform([0, -3, -0.4, -1], [35, 3, 0.4, 5], size=(16, 4))
    u_lqr, costs = [], []
    for x in xs:
        u_lqr.append(lq.step(x))
        costs.append(lq.stage_cost)
    np.savez(os.path.join(OUT, "lqr.npz"), x=xs, u=np.array(u_lqr), stage_cost=np.array(costs))

    # ---- Fourier-Motzkin (lib/in_adm_set.py) ---------------------------------------------------------
    G = rng.normal(size=(7, 2))
    Hm = rng.normal(size=(7, 2))
    Hm[2, :] = 0.0
    phi = -np.abs(rng.normal(size=7)) - 0.5
    P1, g1 = algorithm_1(G, Hm[:, 0].copy(), phi)
    P2, g2 = algorithm_2(G, Hm, phi)
    np.savez(os.path.join(OUT, "in_adm_set.npz"), G=G, H=Hm, phi=phi, P1=P1, g1=g1, P2=P2, g2=g2)

    # ---- shipped terminal sets (data) ------------------------------------------------------------------
    os.makedirs(os.path.join(OUT, "terminal_sets"), exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "terminal_sets"))):
        shutil.copyfile(os.path.join(REF, "terminal_sets", f), os.path.join(OUT, "terminal_sets", f))

    # ---- config-1 grid evaluated with the reference expression (lib/terminal_set.py:96-113) -----------------
    ts = np.load(os.path.join(REF, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy"))
    A, b = ts[..., :4], ts[..., 4]
    center = [29.9, 1.5, 0, 0]
    extent, steps = 25, 100
    xx, yy = np.meshgrid(np.linspace(-extent + center[0], extent + center[0], steps),
                         np.linspace(-extent + center[1], extent + center[1], steps))
    member = np.zeros((6, steps, steps), dtype=bool)
    margin = np.zeros((6, steps, steps))
    for idx, v in enumerate(np.arange(6)):
        for i in range(xx.shape[0]):
            for j in range(xx.shape[1]):
                point = [xx[i, j], yy[i, j], 0, v]
                r = A @ point
                member[idx, i, j] = np.all(r <= b).astype(bool)
                margin[idx, i, j] = np.min(b - r)
    np.savez_compressed(os.path.join(OUT, "grid_config1.npz"), member=member, margin=margin, xx=xx, yy=yy)
    print("members:", member.sum(), member.reshape(6, -1).sum(1), "ties:", (margin[member] == 0).sum())

    # ---- the QP of lib/mpc.py:318-332 / :461-478, as the reference controller object holds it --------------
    qp_cases = [("RoadOneCarEnv", RoadOneCarEnv, [29.9, 1.5, 0, 0], (10, 20, 40, 80)),
                ("RoadEnv", RoadEnv, None, (20,)),
                ("RoadMultipleCarsEnv", RoadMultipleCarsEnv, None, (20,))]
    for tag, cls, goal, horizons in qp_cases:
        for N in horizons:
            e = cls()
            if goal is not None:
                e.set_goal(goal)
            cN = ctrl(e, N)
            cN.set_goal(e.goal)
            At, bt = cN.terminal_constraint()
            Ai, bi = cN.input_constraint()
            As, bs = cN.state_constraint()
            # float32-exact structural zeros stay zeros; everything float64
            np.savez_compressed(os.path.join(OUT, f"qp_{tag}_N{N}.npz"), H=cN.H, h=cN.h, T=cN.T, S=cN.S, P=cN.P, Q=cN.Q,
                                R=cN.R, At=At, bt=bt, Ai=Ai, bi=bi, As=As, bs=bs, goal=np.array(cN.goal, dtype=float))

    # ---- ABd of the disturbance controller (lib/mpc.py:631-635), the reference's own expression ----------------
    from lib.mpc import MPCOutputFBWithDisturbance
    dist = {}
    for N in (1, 5, 20):
        dc = MPCOutputFBWithDisturbance(dt=cfg.DT_CONTROL, N=N, lin_state=cfg.LINEARIZE_STATE,
                                        lin_input=cfg.LINEARIZE_INPUT, init_state=[20, 0.5, 0, 2], env=RoadEnv())
        to_stack = [np.vstack(
            [np.linalg.matrix_power(dc.A, i - j) @ dc.Bd if (i - j) >= 0 else np.zeros(4) for i in range(dc.N)]).T
                    for j in range(dc.N + 1)]
        to_stack.reverse()
        dist[f"ABd_N{N}"] = np.vstack(to_stack).sum(axis=1)
    dist.update(Bd=dc.Bd.astype(float), Cd=dc.Cd.astype(float), L1=dc.L1, L2=dc.L2.astype(float), C=dc.C.astype(float))
    np.savez(os.path.join(OUT, "disturbance.npz"), **dist)


if __name__ == "__main__":
    main()
