"""Builds libdnn_b200.so in-tree with nvcc for sm_100a only (the .so is git-ignored but travels
to the GPU box with the snapshot).  Usage: python dnn-inference-engine_b200/build.py [--force]"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdnn_b200.so")
SOURCES = ["engine.cu"]
# everything under csrc/ is compiled into the one translation unit
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + [os.path.join("..", "..", "include", "dnn_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr"]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.exists(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


def build(force=False, verbose=False, extra=()):
    deps = [os.path.join(CSRC, d) for d in DEPS]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a != "--force"]
    print(build(force="--force" in sys.argv, verbose=True, extra=extra))
