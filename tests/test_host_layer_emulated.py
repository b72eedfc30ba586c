"""The host layer (gubernator_b200/csrc/host_v1.cpp: GetRateLimits mirror, RPC aggregator, Store plugin) on the CPU: built together
with tests/emu_abi.cpp (the C ABI over the emulated kernels) and driven through the normal Python binding in a subprocess, with
the bodies of the GPU tests.  See tests/host_layer_emulated_cases.py."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libgub_emulated_test.so")
CSRC = os.path.join(ROOT, "gubernator_b200", "csrc")
SRC = [os.path.join(HERE, f) for f in ("emu_abi.cpp", "emu_abi_stubs.cpp")] + [os.path.join(CSRC, f) for f in ("host_util.cpp", "host_v1.cpp")]
DEPS = SRC + [os.path.join(HERE, f) for f in ("kernel_emu_harness.cpp", "cuda_emu.h")] + \
    [os.path.join(CSRC, f) for f in ("gub_kernels.cuh", "gub_p2p.cuh", "bucket_math.cuh")] + \
    [os.path.join(ROOT, "include", f) for f in ("gubernator_b200.h", "gubernator_b200_host.h")]


def test_host_layer_on_the_emulated_abi():
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-msse2", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-subobject-linkage", "-I", os.path.join(ROOT, "include"), "-x", "c++"] + SRC + ["-o", SO, "-lpthread"])
    res = subprocess.run([sys.executable, os.path.join(HERE, "host_layer_emulated_cases.py"), SO], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "test bodies passed" in res.stdout
