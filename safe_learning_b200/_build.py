"""Build libslb200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["gp_sweep.cu", "light.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "slb200.h")]
OUTPUT = os.path.join(HERE, "libslb200.so")

NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]


def _stale():
    if not os.path.exists(OUTPUT):
        return True
    built = os.path.getmtime(OUTPUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > built for d in deps)


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> libslb200.so.  Returns the output path."""
    if not force and not _stale():
        return OUTPUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUTPUT] + SOURCES
    proc = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libslb200.so:\n" + proc.stdout)
    return OUTPUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
