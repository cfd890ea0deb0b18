"""Builds the in-tree CUDA shared library (sm_100a) with nvcc. No JIT cache: the .so travels with the repo."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libteb_b200.so")
SOURCES = [os.path.join(CSRC, "teb_cabi.cu")]
DEPS = [os.path.join(CSRC, f) for f in ("teb_cabi.cu", "teb_kernels.cuh", "teb_device.cuh", "teb_resize.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "teb_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC] + FLAGS + ["-o", LIB] + SOURCES + ["-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libteb_b200.so")
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
