"""Builds marl_soccer_b200/libmsoc.so (hand-written sm_100a kernels + C-ABI) in-tree with nvcc.

    python -m marl_soccer_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "msoc.cu")
DEPS = [SRC, os.path.join(PKG, "csrc", "step_core.cuh"), os.path.join(ROOT, "include", "msoc.h")]
OUT = os.path.join(PKG, "libmsoc.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def build(force: bool = False, verbose: bool = False) -> str:
    stale = (not os.path.exists(OUT)) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)
    if force or stale:
        nvcc = os.environ.get("NVCC", "nvcc")
        cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-o", OUT, SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libmsoc.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
