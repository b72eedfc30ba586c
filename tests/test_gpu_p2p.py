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
