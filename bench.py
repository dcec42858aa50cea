#!/usr/bin/env python
"""Benchmark of the CarMPC batch-evaluation hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline line (one JSON object on stdout, rank 0): terminal-set samples/s on BASELINE config 2
(RoadMultipleCarsEnv H-rep, 100^4 = 10^8-point float64 SoA grid per GPU, resident in HBM).  One "step" = one
membership pass over the whole grid -> bitset + count.  The same line carries the QP half of the metric under
``"qp"`` (config 3: 10^6 horizon-20 condensed QPs, RoadOneCarEnv).

``--impl reference`` times the reference's CPU path (its per-point membership test, restated in oracle/ because the
reference is pure Python with third-party solvers that are not installed) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_SAMPLE = 32.125          # 4 x float64 read + 1 bit written
TERMINAL_SET = os.path.join(ROOT, "terminal_sets", "RoadMultipleCarsEnv_30_1.5_0_0.npy")
QP_TERMINAL_SET = os.path.join(ROOT, "terminal_sets", "RoadOneCarEnv_29.9_1.5_0_0.npy")


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# ------------------------------------------------------------------------------------------------------------
def _cpu_sample(points: int):
    """Bounded sample of the config-2 grid: every (10^8 / points)-th grid point, in grid order."""
    from carmpc_b200.grids import config2_axes, grid_size
    axes = config2_axes()
    n = grid_size(axes)
    idx = np.arange(0, n, n // points, dtype=np.int64)[:points]
    dims = [len(a) for a in axes]
    cols, stride = [], n
    for a, d in zip(axes, dims):
        stride //= d
        cols.append(np.ascontiguousarray(a[(idx // stride) % d]))
    return cols


def cpu_membership_rate(min_seconds: float, points: int = 10_000_000, threads: int = 0):
    """samples/s of the C restatement of lib/terminal_set.py:107-113 on all host threads."""
    from oracle import c_oracle
    Ab = np.load(TERMINAL_SET)
    cols = _cpu_sample(points)
    threads = threads or c_oracle.max_threads()
    c_oracle.membership_bits(Ab, *[c[:100000] for c in cols], threads=threads)       # warm-up / build
    t0 = time.perf_counter()
    passes = 0
    while True:
        c_oracle.membership_bits(Ab, *cols, threads=threads)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return passes * points / dt, threads, f"{passes} passes over a {points}-point strided slice of the 10^8 grid", dt / passes


def cpu_pointwise_rate(points: int = 60000):
    """The reference's literal per-point Python loop (np.all(A @ point <= b)), one core."""
    from oracle import carmpc_oracle as orc
    Ab = np.load(TERMINAL_SET)
    cols = _cpu_sample(points)
    pts = np.stack(cols, axis=1)
    t0 = time.perf_counter()
    orc.membership_pointwise(Ab, pts)
    return points / (time.perf_counter() - t0)


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    per_step_s = min(1.0, 120.0 / max(1, steps))          # bounded sample per step: the whole run ends within minutes
    for _ in range(warmup):
        cpu_membership_rate(0.0, points=2_000_000)
    rates, step_ms = [], []
    for _ in range(steps):
        rate, threads, sample, dt = cpu_membership_rate(per_step_s)
        rates.append(rate)
        step_ms.append(dt * 1e3)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "terminal-set samples/s", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": float(np.mean(step_ms)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config 2: RoadMultipleCarsEnv terminal set (42 rows) on the 100^4 float64 grid",
                   "note": "the reference is pure Python (cvxpy/polytope not installed, no build); this arm is the "
                           "C restatement of lib/terminal_set.py:107-113 in oracle/, all host threads"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                         "pointwise_python_samples_per_s": cpu_pointwise_rate()},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from carmpc_b200.batch import TerminalSetEvaluator
    from carmpc_b200.grids import config2_axes, materialise_grid, grid_size

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed and not dist.is_initialized():
        # NCCL prints its version banner on stdout when the first communicator is created; the contract is ONE JSON
        # line on stdout, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    Ab = np.load(TERMINAL_SET)
    ev = TerminalSetEvaluator(Ab)
    axes = config2_axes()
    n = grid_size(axes)
    x, y, psi, v = materialise_grid(axes, device=dev)                 # 3.2 GB of float64 SoA per GPU
    words = (n + 31) // 32
    bits = [torch.empty(words, dtype=torch.int32, device=dev) for _ in range(2)]
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    from carmpc_b200.sharding import gather_bitset, reduce_count
    launches = 0

    def step(i):
        # samples shard with no data-path collective: a step is this rank's membership pass over its own grid
        nonlocal launches
        ev.contains_bits(x, y, psi, v, mode=args.mode, bits=bits[i & 1], count=count)
        launches += 1

    for i in range(args.warmup):
        step(i)
    barrier()
    launches = 0
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        start.record()
        for i in range(args.steps):
            step(i)
        stop.record()
        barrier()
    ms = start.elapsed_time(stop)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    members = int(count.item())
    # result collection (outside the timed steps): bitset words of every rank over NCCL / NVLink, and the member count
    gather_ms = None
    if distributed:
        gather_bitset(bits[0], world * n)                    # communicator warm-up
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = gather_bitset(bits[(args.steps - 1) & 1], world * n)
        total_members = reduce_count(count.clone())
        g1.record()
        g1.synchronize()
        gather_ms = g0.elapsed_time(g1)
        assert full.numel() == world * words and int(total_members.item()) == world * members

    # kernel-only timing for the roofline (no gather), same inputs (3.2 GB >> 126 MB L2, so no flush is needed)
    k_ms = []
    for i in range(min(args.steps, 50)):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        ev.contains_bits(x, y, psi, v, mode=args.mode, bits=bits[0], count=count)
        s1.record()
        s1.synchronize()
        k_ms.append(s0.elapsed_time(s1))
    kernel_ms = float(np.mean(k_ms))

    # ---- end to end through the host-buffer C-ABI call (pinned host SoA in, host bitset out) -----------------
    e2e = None
    if not args.skip_e2e:
        hn = args.e2e_samples
        host = [torch.empty(hn, dtype=torch.float64, pin_memory=True).copy_(t[:hn]) for t in (x, y, psi, v)]
        hx, hy, hp, hv = [h.numpy() for h in host]
        ev.contains_bits_host(hx[:1 << 20], hy[:1 << 20], hp[:1 << 20], hv[:1 << 20], mode=args.mode)   # warm-up
        ev.contains_bits_host(hx, hy, hp, hv, mode=args.mode)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = args.e2e_steps
        for _ in range(e2e_steps):
            hbits, hcount = ev.contains_bits_host(hx, hy, hp, hv, mode=args.mode)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * hn * e2e_steps / float(tt.item()), "unit": "samples/s",
               "h2d_bytes_per_step": int(hn * 32), "d2h_bytes_per_step": int((hn + 31) // 32 * 4 + 8),
               "samples_per_step": hn, "steps": e2e_steps,
               "call": "carmpc_membership_bitset_host (pinned host SoA -> chunked H2D | kernel | D2H pipeline)"}

    # the same grid through the implicit-grid entry point: host axes in (3.2 KB), host bitset out (12.5 MB)
    e2e_grid = None
    if not args.skip_e2e:
        gbits = torch.empty(words, dtype=torch.int32, device=dev)
        hbits = torch.empty(words, dtype=torch.int32, pin_memory=True)
        ev.contains_grid_bits(axes, bits=gbits, count=count)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            ev.contains_grid_bits(axes, bits=gbits, count=count)
            hbits.copy_(gbits, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        e2e_grid = {"value": world * n / dt, "unit": "samples/s", "h2d_bytes_per_step": int(sum(len(a) for a in axes) * 8),
                    "d2h_bytes_per_step": int(words * 4 + 8),
                    "call": "carmpc_membership_grid (axes on the host, coordinates generated in-kernel) + D2H of the bitset"}

    # ---- the rollout form of the same test (lib/terminal_set.py:53-59, 198-200 sampled): k* + 1 closed-loop steps, state rows
    # at every step, input rows at t = 0; float32 screen over the expanded rows + float64 step-by-step rollout ----------
    rollout = None
    if not args.skip_rollout:
        from carmpc_b200.batch import RolloutEvaluator
        from carmpc_b200.lib.environments import RoadMultipleCarsEnv
        k_star = 16
        rv = RolloutEvaluator.from_env(RoadMultipleCarsEnv(), k_star)
        rcount = torch.zeros(1, dtype=torch.int64, device=dev)
        for _ in range(3):
            rv.contains_bits(x, y, psi, v, bits=bits[0], count=rcount)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(args.rollout_steps):
            rv.contains_bits(x, y, psi, v, bits=bits[0], count=rcount)
        r1.record()
        barrier()
        r_ms = r0.elapsed_time(r1) / args.rollout_steps
        tr = torch.tensor([r_ms], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        r_ms = float(tr.item())
        r_members = int(rcount.item())
        # the plain float64 kernel (one sample per thread; selected when the first violated step is requested)
        rv.contains_bits(x, y, psi, v, want_first_violation=True, bits=bits[0], count=rcount)
        torch.cuda.synchronize()
        r0.record()
        for _ in range(3):
            rv.contains_bits(x, y, psi, v, want_first_violation=True, bits=bits[0], count=rcount)
        r1.record()
        torch.cuda.synchronize()
        exact_ms = r0.elapsed_time(r1) / 3
        peak, _ = _peaks()
        s_rows, r_in = len(rv.b_con), len(rv.b_in)
        rollout = {"metric": "rollout-form terminal-set samples/s", "value": world * n / (r_ms * 1e-3), "unit": "samples/s",
                   "ms_per_step": r_ms, "steps": args.rollout_steps, "k_steps": k_star, "state_rows": s_rows, "input_rows": r_in,
                   "screen_rows": s_rows * (k_star + 1) + r_in, "members": r_members, "members_hrep": members,
                   "float64_kernel_ms": exact_ms, "float64_kernel_samples_per_s": n / (exact_ms * 1e-3),
                   "float64_kernel_note": "rollout_kernel, also writes the first violated step (4 B/sample)",
                   "roofline": {"bound": "hbm", "achieved": n * BYTES_PER_SAMPLE / (r_ms * 1e-3) / 1e9, "peak": peak,
                                "unit": "GB/s", "frac": n * BYTES_PER_SAMPLE / (r_ms * 1e-3) / 1e9 / peak,
                                "kernel": "membership_tma_kernel<1, rollout>"},
                   "flop_per_sample_float64_form": (k_star + 1) * 8 * s_rows + 32 * k_star + 32}
        if rank == 0 and world == 1 and not args.skip_cpu:
            from oracle import c_oracle
            cols = _cpu_sample(2_000_000)
            t0 = time.perf_counter()
            passes = 0
            while time.perf_counter() - t0 < 4.0:
                c_oracle.rollout_bits(rv.A_k, rv.A_con, rv.b_con, rv.A_in, rv.b_in, rv.goal, k_star, 0, *cols)
                passes += 1
            rollout["cpu_baseline"] = {"value": passes * 2_000_000 / (time.perf_counter() - t0), "unit": "samples/s",
                                       "cores": c_oracle.max_threads(), "kind": "port",
                                       "sample": f"{passes} passes over a 2000000-point strided slice of the 10^8 grid"}

    qp = None
    if not args.skip_qp:
        try:
            from bench_qp import run_qp_bench
            qp = run_qp_bench(args, rank, world, dev, barrier)
        except Exception as exc:                                      # the QP half must never hide the headline
            qp = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        peak, peak_src = _peaks()
        # at N = 1 a step is exactly one launch of the kernel: its average duration over the timed region (back-to-back
        # launches, CUDA events on the launching stream); with the gather in the step (N > 1) the isolated timing is used
        launch_ms = ms / args.steps if not distributed else kernel_ms
        achieved = n * BYTES_PER_SAMPLE / (launch_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.skip_cpu:
            rate, threads, sample, _ = cpu_membership_rate(10.0)
            cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                   "pointwise_python_samples_per_s": cpu_pointwise_rate()}
        line = {
            "metric": "terminal-set samples/s", "value": world * n * args.steps / (ms_total * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 2: RoadMultipleCarsEnv terminal set (42 rows) on the 100^4 = 10^8-point "
                                   "float64 SoA grid per GPU; one step = one membership pass -> bitset + count",
                       "samples_per_gpu_per_step": n, "kernel_mode": "fp32 screen + fp64 re-check" if args.mode else "fp64",
                       "l2": "inputs (3.2 GB) exceed the 126 MB L2; no flush between iterations",
                       "members": members,
                       "multi_gpu": "each rank scans its own 10^8 grid, no data-path collective; the bitsets are all-gathered "
                                    "over NCCL after the timed steps (result_gather_ms)" if distributed else "single GPU",
                       "result_gather_ms": gather_ms},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "e2e_grid": e2e_grid,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": 3.2125e9, "traffic_source": "ncu --set full: dram__bytes_read.sum 3.200 GB + "
                         "dram__bytes_write.sum 12.5 MB per launch (profiles/r01_s5_membership_tma_ncu_full.txt)",
                         "peak_source": peak_src, "launch_ms": launch_ms, "isolated_launch_ms": kernel_ms,
                         "kernel": "membership_tma_kernel<1>" if args.mode == 1 else "membership_kernel<0,true>",
                         "bytes_per_sample": BYTES_PER_SAMPLE, "algorithmic_bytes_per_launch": n * BYTES_PER_SAMPLE},
            "cpu_baseline": cpu,
            "rollout": rollout,
            "qp": qp,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", type=int, default=1, help="membership kernel: 0 = float64, 1 = float32 screen + float64 re-check")
    ap.add_argument("--e2e-samples", type=int, default=100_000_000)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-qp", action="store_true")
    ap.add_argument("--skip-rollout", action="store_true")
    ap.add_argument("--rollout-steps", type=int, default=20)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--qp-states", type=int, default=1_000_000)
    ap.add_argument("--qp-steps", type=int, default=5)
    ap.add_argument("--skip-sweep", action="store_true")
    ap.add_argument("--skip-seeded", action="store_true")
    ap.add_argument("--seed-blocks", default="2x8x1x1", help="comma-separated lattice blocks (points per axis) for the seeded map")
    ap.add_argument("--skip-closed-loop", action="store_true")
    ap.add_argument("--cl-runs", type=int, default=100_000)
    ap.add_argument("--cl-steps", type=int, default=200)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: launch with torch.distributed.run for --gpus > 1")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
