"""Build tests/kernel_emu/libkernel_emu.so: every kernel file of csrc/ (ingest, projection, radix_sort, ranges, compositor) compiled by g++ for the CPU (see cuda_shim.h).
TEST INFRASTRUCTURE; needs only the CUDA headers (no GPU, no nvcc)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libkernel_emu.so")
CXX = os.environ.get("ORC_CXX", "/usr/bin/g++")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
DEPS = [os.path.join(HERE, f) for f in ("kernel_emu.cpp", "cuda_shim.h", "build.py")] + [
    os.path.join(ROOT, "oracle", "glsl_cpu", "glsl_emu.hpp"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "compositor.cu"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "ranges.cu"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "radix_sort.cu"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "projection.cu"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "ingest.cu"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "present.cu"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "group.cu"),
    os.path.join(ROOT, "godotgaussiansplatting_b200", "csrc", "common.cuh"),
]


def build(force: bool = False) -> str:
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in DEPS):
        # -ffp-contract=off: the kernels' arithmetic contract (nvcc -fmad=false); fmaf() only where the source says so
        subprocess.run([CXX, "-std=gnu++17", "-O1", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w",
                        "-I", CUDA_INC, os.path.join(HERE, "kernel_emu.cpp"), "-o", OUT], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
