"""Fused NVLink-mailbox routing (gub_p2p_step) on ONE GPU: W shards in one process, each on its own stream, mailboxes
shared by pointer (gub_p2p_connect_local); responses checked against one oracle per shard applied in the documented
order (source rank, then source index).  The cross-process cudaIpc variant runs in tests/test_gpu_sharded.py."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_py as O
from workloads import T0, adversarial_batch, bench_requests, zipf_ids

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 4])
def test_p2p_step_matches_per_shard_oracles(world):
    import torch
    import gubernator_b200 as g
    from gubernator_b200.sharded import P2PStep, shard_addresses
    dev = torch.device("cuda", 0)
    ring, oring = g.Ring(0, 512), O.Ring(0, 512)
    for a in shard_addresses(world):
        ring.add(a); oring.add(a)
    tabs = [g.Table(1 << 16, max_batch=65536, device=0) for _ in range(world)]
    steppers = [P2PStep(tabs[r], ring, world, r, cap=8192) for r in range(world)]
    for s in steppers:
        s.connect_local(steppers)
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    sim = [O.Pool(workers=2, cache_size=10**7, now_ms=T0) for _ in range(world)]
    pool = ThreadPoolExecutor(world)
    sizes = [[3000, 1, 0, 8192, 5], [2000, 0, 0, 8192, 700], [1, 4000, 0, 100, 8192], [8192, 0, 0, 7, 300]]
    for step in range(10):
        now = T0 + step
        rngs = [np.random.default_rng(77 * step + r) for r in range(world)]
        batches = []
        for r in range(world):
            n = sizes[r % 4][step % 5]
            if n == 0:
                batches.append(np.zeros(0, dtype=O.HREQ_DTYPE))
            elif step % 2:
                batches.append(adversarial_batch(rngs[r], n, 41, now))
            else:
                batches.append(bench_requests(zipf_ids(rngs[r], n, 3000, 1.1), now))
        clk = g.clock_fill(now)

        def run(r):
            torch.cuda.set_device(0)
            b, n = batches[r], len(batches[r])
            with torch.cuda.stream(streams[r]):
                buf = torch.from_numpy(b.view(np.uint8).reshape(n, 64).copy()).to(dev) if n else torch.empty((1, 64), dtype=torch.uint8, device=dev)
                out = torch.zeros((max(n, 1), 32), dtype=torch.uint8, device=dev)
                steppers[r].step(buf, n, clk, out, stream=streams[r].cuda_stream)
                streams[r].synchronize()
                return out[:n].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE)
        got = [f.result() for f in [pool.submit(run, r) for r in range(world)]]
        owners = [np.array([oring.get_by_hash(int(h)) for h in b["key_fnv1"]], dtype=np.int64) for b in batches]
        want = [np.zeros(len(b), dtype=O.HRESP_DTYPE) for b in batches]
        for gi in range(world):
            sim[gi].set_now(now)
            for s in range(world):
                idx = np.nonzero(owners[s] == gi)[0]
                if len(idx):
                    want[s][idx] = sim[gi].submit_hashed(np.ascontiguousarray(batches[s][idx]))
        for r in range(world):
            assert np.array_equal(got[r], want[r]), f"step {step} shard {r}: {int((got[r] != want[r]).sum())} differ"


@pytest.mark.parametrize("world", [2, 4])
def test_p2p_two_stream_pipeline(world):
    """gub_p2p_step_streams: routing on an ingest stream, evaluation on another; eight steps enqueued back to back per shard
    with no synchronisation in between (step e+1 is being routed while step e is evaluated), then every response of every
    step is compared with the per-shard oracles."""
    import torch
    import gubernator_b200 as g
    from gubernator_b200.sharded import P2PStep, shard_addresses
    dev = torch.device("cuda", 0)
    ring, oring = g.Ring(0, 512), O.Ring(0, 512)
    for a in shard_addresses(world):
        ring.add(a); oring.add(a)
    tabs = [g.Table(1 << 16, max_batch=65536, device=0) for _ in range(world)]
    steppers = [P2PStep(tabs[r], ring, world, r, cap=4096) for r in range(world)]
    for s in steppers:
        s.connect_local(steppers)
    s_in = [torch.cuda.Stream(device=dev) for _ in range(world)]
    s_ev = [torch.cuda.Stream(device=dev) for _ in range(world)]
    STEPS = 8
    sizes = [[3000, 1, 0, 4096, 5, 2000, 4096, 17], [2000, 0, 0, 4096, 700, 1, 4096, 4096], [1, 4000, 0, 100, 4096, 9, 0, 3], [4096, 0, 0, 7, 300, 4096, 1, 1]]
    batches = [[None] * world for _ in range(STEPS)]
    for step in range(STEPS):
        for r in range(world):
            rng = np.random.default_rng(991 * step + r)
            n = sizes[r % 4][step]
            if n == 0:
                batches[step][r] = np.zeros(0, dtype=O.HREQ_DTYPE)
            elif step % 2:
                batches[step][r] = adversarial_batch(rng, n, 41, T0 + step)
            else:
                batches[step][r] = bench_requests(zipf_ids(rng, n, 3000, 1.1), T0 + step)
    bufs = [[torch.from_numpy(b.view(np.uint8).reshape(len(b), 64).copy()).to(dev) if len(b) else torch.empty((1, 64), dtype=torch.uint8, device=dev)
             for b in row] for row in batches]
    outs = [[torch.zeros((max(len(b), 1), 32), dtype=torch.uint8, device=dev) for b in row] for row in batches]
    clks = [g.clock_fill(T0 + step) for step in range(STEPS)]
    torch.cuda.synchronize()

    def run(r):
        torch.cuda.set_device(0)
        for step in range(STEPS):
            steppers[r].step(bufs[step][r], len(batches[step][r]), clks[step], outs[step][r], stream=s_ev[r].cuda_stream,
                             ingest_stream=s_in[r].cuda_stream)
    with ThreadPoolExecutor(world) as pool:
        for f in [pool.submit(run, r) for r in range(world)]:
            f.result()
    torch.cuda.synchronize()
    sim = [O.Pool(workers=2, cache_size=10**7, now_ms=T0) for _ in range(world)]
    for step in range(STEPS):
        owners = [np.array([oring.get_by_hash(int(h)) for h in b["key_fnv1"]], dtype=np.int64) for b in batches[step]]
        for gi in range(world):
            sim[gi].set_now(T0 + step)
            for s in range(world):
                idx = np.nonzero(owners[s] == gi)[0]
                if len(idx):
                    want = sim[gi].submit_hashed(np.ascontiguousarray(batches[step][s][idx]))
                    got = outs[step][s][:len(batches[step][s])].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE)[idx]
                    assert np.array_equal(got, want), f"step {step} source {s} owner {gi}: {int((got != want).sum())} differ"
