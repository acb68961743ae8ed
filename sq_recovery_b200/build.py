"""Build libsqloss.so (and the pipe-throughput microbenchmark) in-tree with nvcc for sm_100a.

    python -m sq_recovery_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libsqloss.so")
LIB_COUNT = os.path.join(HERE, "libsqloss_count.so")
PEAKS = os.path.join(HERE, "sq_peaks")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-ftz=true", "-std=c++17"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out: str, srcs) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, "sqloss.cu"), os.path.join(CSRC, "sq_core.cuh"), os.path.join(ROOT, "include", "sqloss.h"),
            os.path.join(CSRC, "sq_tables.inc")]
    if force or _stale(LIB, srcs):
        cmd = [_nvcc(), *ARCH, *FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, srcs[0]]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    # the counting build (bench.py: walked fraction, issued-op counts): same sources, -DSQ_COUNT; never the timed library
    if force or _stale(LIB_COUNT, srcs):
        subprocess.check_call([_nvcc(), *ARCH, *FLAGS, "-DSQ_COUNT", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_COUNT, srcs[0]])
    peaks_src = os.path.join(CSRC, "peaks.cu")
    if force or _stale(PEAKS, [peaks_src]):
        subprocess.check_call([_nvcc(), *ARCH, "-O3", "-o", PEAKS, peaks_src])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
