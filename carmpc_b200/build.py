"""Builds ``carmpc_b200/_lib/libcarmpc_b200.so`` (the C ABI of include/carmpc.h) with nvcc for sm_100a.

One scripted command, usable from a fresh checkout:  ``python -m carmpc_b200.build [--force]``.
nvcc cross-compiles without a GPU; the built library is git-ignored and travels to the GPU box with the
``gpurun`` snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libcarmpc_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libcarmpc_b200.so (set NVCC=/path/to/nvcc)")


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def is_stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` into one shared library; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            return obj, "", 0
        proc = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", "-o", obj, src], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)
        return obj, proc.stdout, proc.returncode

    # one nvcc per translation unit, in parallel (qp_admm.cu alone is minutes of ptxas time)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(sources()), os.cpu_count() or 1))) as pool:
        results = list(pool.map(compile_one, sources()))
    log = "".join(out for _, out, _ in results)
    failed = [obj for obj, _, rc in results if rc != 0]
    if verbose or failed:
        print(log, file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed building " + ", ".join(os.path.basename(f) for f in failed) + " (output above)")
    tmp = LIB_PATH + ".tmp"
    proc = subprocess.run([nvcc, "-shared", "-o", tmp] + [obj for obj, _, _ in results], stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        print(proc.stdout, file=sys.stderr)
        raise RuntimeError("linking libcarmpc_b200.so failed (output above)")
    if log.strip():
        with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
            f.write(log)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
