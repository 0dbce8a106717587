"""Builds tests/emu/libpc_emu.so: the SC-list kernel SOURCES compiled for the CPU (test infrastructure only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libpc_emu.so")
SRC = [os.path.join(HERE, "sclp_emu.cpp")]
DEPS = SRC + [os.path.join(HERE, "cuda_emu.h")] + [os.path.join(ROOT, "polarcub_b200", "csrc", f)
                                                    for f in ("scl_path.cu", "scl_tables.cu", "scl_tables.cuh", "scl_arith.cuh", "common.cuh")]


def build(force=False):
    if not force and os.path.isfile(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in DEPS):
        return SO
    cuda_inc = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-DPC_EMU", "-I", cuda_inc, "-shared", "-fPIC",
           "-Wno-unused-function", "-o", SO] + SRC
    subprocess.run(cmd, check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True))
