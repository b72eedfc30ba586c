"""Builds libgubernator_b200.so in-tree with nvcc for sm_100a (no torch involved: the library is plain CUDA runtime)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgubernator_b200.so")
SOURCES = [os.path.join(CSRC, f) for f in ("gub_api.cu", "host_util.cpp", "host_v1.cpp")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("gub_kernels.cuh", "bucket_math.cuh")] + [os.path.join(ROOT, "include", "gubernator_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",              # Go on amd64 never fuses multiply-add; the float64 leaky-bucket math must be bit-exact
    "-Xcompiler", "-fPIC,-O2,-Wall,-ffp-contract=off",
    "-shared", "-cudart", "static", "-lpthread",
]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(s)]
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("GUB_NVCC_EXTRA", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgubernator_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
