"""Runs in a subprocess of tests/test_host_layer_emulated.py: the Python binding is pointed at tests/libgub_emulated_test.so — the
REAL host layer (host_v1.cpp, host_util.cpp) over the C ABI implemented on the CPU emulation of the kernels (tests/emu_abi.cpp) —
and the bodies of the GPU tests that exercise host logic are run unchanged: V1Instance.GetRateLimits against the reference's
functional tables, field validation and error strings, UpdatePeerGlobals, the RPC aggregator with concurrent callers, the Store
plugin's call sequences, Load / Store of items.  Test infrastructure only: the product library is never replaced outside this
process."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [ROOT, HERE, os.path.join(ROOT, "oracle")]

import gubernator_b200 as g  # noqa: E402

g.native.LIB_PATH = sys.argv[1]  # before anything has loaded the product library
assert g.native._lib is None

import test_gpu_parity as T  # noqa: E402
from golden import reference_kat as K  # noqa: E402

ran = []


def run(name, *args):
    import time
    t0 = time.time()
    getattr(T, name)(g, *args)
    ran.append(name)
    print(f"  {name}{args if args and not isinstance(args[0], dict) else ''}: {time.time() - t0:.1f}s")


for sc in K.SCENARIOS:
    run("test_functional_scenarios", sc)
run("test_missing_fields_and_batch_cap")
run("test_error_strings_and_order")
run("test_update_peer_globals_items")
run("test_add_get_scan_items")
run("test_invalid_at_of_loaded_items")
run("test_rpc_aggregator_coalesces_concurrent_calls")
run("test_aggregator_honours_the_store_plugin")
for algo in (0, 1):
    run("test_store_plugin_call_sequences", algo)
assert g.native.lib()._name == sys.argv[1]
print(f"host layer on the emulated ABI: {len(ran)} test bodies passed")
