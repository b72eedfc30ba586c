"""The ring inside one box through the C ABI (gub_p2p_*): W shards in one process driven by one host thread
(gub_p2p_step_local_all), records stored by the routing kernel straight into the owners' mailboxes, evaluated out of the
mailboxes by the batch kernel, responses stored into the sources' mailboxes; checked against one oracle per shard applied in
the documented order (source rank, then source index).  On a one-GPU box all shards share the device; with >= 2 GPUs the second
test spreads them (peer memory over NVLink).  The one-process-per-GPU cudaIpc variant runs in tests/test_gpu_sharded.py."""
import numpy as np
import pytest

import oracle_py as O
from local_ring import LocalRing, per_shard_oracle_results
from workloads import T0, adversarial_batch, bench_requests, zipf_ids

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["pipeline", "fused"])
def owner_kernel_path(request, monkeypatch):
    """Owners evaluate their mailboxes with the four-kernel pipeline in ring mode (default) or with the persistent kernel k_batch
    (GUB_PATH=fused); read when a table is created."""
    monkeypatch.setenv("GUB_PATH", request.param)
    return request.param

SIZES = [[3000, 1, 0, 8192, 5], [2000, 0, 0, 8192, 700], [1, 4000, 0, 100, 8192], [8192, 0, 0, 7, 300]]


def _batches(world, step, now, cap):
    out = []
    for r in range(world):
        rng = np.random.default_rng(77 * step + r)
        n = min(SIZES[r % 4][step % 5], cap)
        if n == 0:
            out.append(np.zeros(0, dtype=O.HREQ_DTYPE))
        elif step % 2:
            out.append(adversarial_batch(rng, n, 41, now))
        else:
            out.append(bench_requests(zipf_ids(rng, n, 3000, 1.1), now))
    return out


def _oring(world):
    from gubernator_b200.sharded import shard_addresses
    oring = O.Ring(0, 512)
    for a in shard_addresses(world):
        oring.add(a)
    return oring


@pytest.mark.parametrize("spread", [False, True])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_p2p_step_matches_per_shard_oracles(world, spread):
    import torch
    if spread and (torch.cuda.device_count() < 2 or world == 1):
        pytest.skip("needs >= 2 GPUs")
    cl = LocalRing(world, spread_devices=spread)
    oring = _oring(world)
    sim = [O.Pool(workers=2, cache_size=10**7, now_ms=T0) for _ in range(world)]
    for step in range(10):
        now = T0 + step
        batches = _batches(world, step, now, cl.cap)
        got = cl.step(batches, now)
        want = per_shard_oracle_results(oring, sim, batches, now)
        for r in range(world):
            assert np.array_equal(got[r], want[r]), f"step {step} shard {r}: {int((got[r] != want[r]).sum())} differ"


@pytest.mark.parametrize("world", [2, 4])
def test_p2p_steps_back_to_back(world):
    """Eight steps enqueued back to back with no host synchronisation in between (mailbox halves and routing scratch are reused
    every second step), then every response of every step is compared with the per-shard oracles."""
    cl = LocalRing(world, cap=4096)
    oring = _oring(world)
    STEPS = 8
    batches = [_batches(world, step, T0 + step, 4096) for step in range(STEPS)]
    staged = [cl.upload(b) for b in batches]
    for step in range(STEPS):
        cl.enqueue(staged[step][0], [len(b) for b in batches[step]], T0 + step, staged[step][1])
    cl.sync()
    sim = [O.Pool(workers=2, cache_size=10**7, now_ms=T0) for _ in range(world)]
    for step in range(STEPS):
        want = per_shard_oracle_results(oring, sim, batches[step], T0 + step)
        for r in range(world):
            got = staged[step][1][r][:len(batches[step][r])].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE)
            assert np.array_equal(got, want[r]), f"step {step} shard {r}: {int((got != want[r]).sum())} differ"


def test_a_silent_peer_is_reported_not_waited_for_forever():
    """Device-side waits are bounded: a shard whose peer never steps gets GUB_ERR_PEER_TIMEOUT in-band for the requests that peer
    owns, and gub_p2p_status reports the step (the flag used to be set and never read)."""
    import torch
    import gubernator_b200 as g
    from gubernator_b200.sharded import shard_addresses
    world = 2
    ring = g.Ring(0, 512)
    for a in shard_addresses(world):
        ring.add(a)
    tabs = [g.Table(1 << 12, device=0) for _ in range(world)]
    p2ps = [g.native.P2P(tabs[r], ring, r, 1024) for r in range(world)]
    for p in p2ps:
        p.connect_local(p2ps)
    reqs = bench_requests(np.arange(600), T0)
    buf = torch.from_numpy(reqs.view(np.uint8).reshape(-1, 64).copy()).cuda()
    out = torch.zeros((600, 32), dtype=torch.uint8, device="cuda")
    p2ps[0].step(buf.data_ptr(), 600, g.clock_fill(T0), out.data_ptr(), torch.cuda.current_stream().cuda_stream)  # shard 1 never steps
    torch.cuda.synchronize()
    with pytest.raises(g.native.GubError):
        p2ps[0].status()
    got = out.cpu().numpy().reshape(-1).view(O.HRESP_DTYPE)
    oring = _oring(world)
    owner = np.array([oring.get_by_hash(int(h)) for h in reqs["key_fnv1"]])
    assert np.all(got["err_code"][owner == 1] == g.native.ERR_PEER_TIMEOUT) and (owner == 1).any()
