"""Multi-GPU parity, one process per GPU (torchrun): the routing kernels + per-shard evaluation against one oracle per shard — records by
NCCL all-to-all ("nccl"), by NVLink mailboxes through cudaIpc ("p2p"; "p2p2": routing on its own stream, all steps in flight) — and the
GLOBAL behaviour with the C-side sync tick and its NCCL all-gather against the oracle cluster model ("p2pg").
Needs >= 2 GPUs on the box (skipped otherwise; the protocol itself is covered on CPU by tests/test_sharded_gloo.py)."""
import pytest

from test_sharded_gloo import run_workers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("route", ["nccl", "p2p", "p2p2", "p2pg"])
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_gpu_matches_per_shard_oracles(nproc, route):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    res = run_workers(nproc, "gpu", route=route)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    for r in range(nproc):
        assert f"rank {r} ok" in res.stdout
