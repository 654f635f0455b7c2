"""Builds libma_b200.so in-tree with nvcc for sm_100a (explicit -gencode; cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libma_b200.so")
SOURCES = ["ma_b200.cu"]
HEADERS = ["ma_geom.cuh", "ma_cell.cuh", "ma_block.cuh", "ma_warm.cuh", "ma_amg.cuh", "ma_seg.cuh", "ma_kernels.cuh", "ma_pcg.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "ma_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    tmp = LIB_PATH + f".{os.getpid()}.tmp"
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)  # atomic: a snapshot / another process never sees a half-written library
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
