"""Compile the oracle's C restatements (gcc + OpenMP) into oracle/_build/ (git-ignored)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libtheta_cg_omp.so")
SRC = os.path.join(HERE, "c", "theta_cg_omp.c")


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
