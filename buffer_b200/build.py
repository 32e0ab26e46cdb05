"""Builds buffer_b200/libbuffer_b200.so: hand-written CUDA for sm_100a behind the C ABI of include/buffer_b200.h.

Plain nvcc, no torch dependency, in-tree output (the .so travels to the GPU box with the repo snapshot).
``python -m buffer_b200.build`` or ``buffer_b200.build.build()``.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libbuffer_b200.so")
SOURCES = ["mutual_nn.cu", "mutual_nn_tc.cu", "ransac.cu", "refine.cu", "extras.cu", "api.cu"]
HEADERS = ["bfr_common.cuh", "bfr_kernels.h", "bfr_tcgen05.cuh", os.path.join("..", "..", "include", "buffer_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--fmad=false",            # no implicit contraction: every FMA in the kernels is an explicit intrinsic
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--shared", "-cudart", "shared"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libbuffer_b200.so")
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)
