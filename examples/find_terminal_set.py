#!/usr/bin/env python
"""Counterpart of the reference's examples/find_terminal_set.py: build the terminal set of an environment, compare it
with the shipped terminal_sets/*.npy and evaluate the reference's own sample grid (lib/terminal_set.py:96-113) on the GPU.

    python examples/find_terminal_set.py [--env RoadOneCarEnv] [--goal 29.9 1.5 0 0] [--save DIR]
"""
import argparse
import os
import tempfile

import numpy as np

import _common
from carmpc_b200.lib import terminal_set as ts


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="RoadOneCarEnv")
    ap.add_argument("--goal", type=float, nargs=4, default=[29.9, 1.5, 0, 0])
    ap.add_argument("--save", default=None, help="directory for the .npy (default: a temporary directory)")
    args = ap.parse_args(argv)
    env = _common.make_env(args.env, args.goal)
    shipped_dir = ts.TERMINAL_SET_DIR
    out_dir = args.save or tempfile.mkdtemp(prefix="terminal_sets_")
    ts.TERMINAL_SET_DIR = out_dir + os.sep
    try:
        A, b = ts.calc_terminal_set(env)                      # writes <out_dir>/<env.name>_<goal>.npy
    finally:
        ts.TERMINAL_SET_DIR = shipped_dir
    print(f"{args.env}: terminal set with {len(b)} rows -> {os.path.join(out_dir, ts.terminal_set_filename(env))}")
    shipped = os.path.join(shipped_dir, ts.terminal_set_filename(env))
    if os.path.isfile(shipped):
        ref = np.load(shipped)
        same = ref.shape == (len(b), 5) and np.abs(ref - np.column_stack((A, b))).max() <= 1e-9
        print(f"shipped fixture {os.path.basename(shipped)}: {'reproduced' if same else 'DIFFERS'}")
    _, inside = ts.grid_membership(A, b, env)                 # the reference's 100 x 100 x 6 grid, one kernel launch
    per_v = inside.reshape(6, -1).sum(axis=1)
    print(f"members on the reference's sample grid: {int(inside.sum())} of {inside.size}, per velocity {per_v.tolist()}")
    return A, b, inside


if __name__ == "__main__":
    main()
