"""Builds the oracle's C kernels (oracle/csrc -> oracle/_build/liboracle_cpu.so).  TEST INFRASTRUCTURE:
called by __graft_entry__.build(); the library is loaded only by oracle/omp.py for the CPU timing arms."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "cpu_kernels.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_cpu.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # no -march=native: the library is built in one container and runs on the GPU box's host CPU
    cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
