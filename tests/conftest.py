import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# The incremental expiry sweep of the batch kernel frees expired entries earlier than the reference's lazy removal
# (lrucache.go:115); responses never depend on it, but the tests that compare the whole table with the oracle's cache item by
# item (expired ones included) need it off.  Tests of the sweep itself turn it on per table (Table.set_sweep).
os.environ.setdefault("GUB_SWEEP", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
