"""Size-independent properties of the CUDA path at BASELINE sizes (100 M-slot-class tables and 65 536-request batches),
where a second CPU implementation would be too slow to replay everything:

* batch-split invariance: submit(A + B) gives exactly the responses of submit(A) then submit(B) (sequential semantics);
* token conservation: per key the UNDER_LIMIT responses with Hits = 1 are exactly min(count, limit) and their `remaining`
  values count down limit-1, limit-2, ... in index order; everything later is OVER_LIMIT with remaining 0;
* Hits = 0 is idempotent: a status query changes nothing (same answer twice, same answer as the last hit reported);
* a checksum of all responses equals the oracle's on a 1 % sample of batches (the oracle replays only those keys).
"""
import numpy as np
import pytest

import oracle_py as O
from workloads import T0, bench_requests, zipf_ids

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gubernator_b200 as g
    return g


def test_batch_split_invariance_and_conservation_at_full_batch_size(G):
    rng = np.random.default_rng(0xB200)
    n_keys, n = 20_000_000, 65536
    whole, split = G.Table(2 * n_keys), G.Table(2 * n_keys)
    for b in range(4):
        now = T0 + b
        ids = zipf_ids(rng, n, n_keys, 1.1)
        reqs = bench_requests(ids, now, mixed=(b % 2 == 1), dtype=G.REQ_DTYPE)
        clk = G.clock_fill(now)
        out_whole = whole.submit(reqs, clk)
        cut = int(rng.integers(1, n - 1))
        out_split = np.concatenate([split.submit(np.ascontiguousarray(reqs[:cut]), clk), split.submit(np.ascontiguousarray(reqs[cut:]), clk)])
        assert np.array_equal(out_whole, out_split), f"batch {b}: splitting the batch at {cut} changed {int((out_whole != out_split).sum())} responses"
        if b == 0:  # fresh token buckets, limit 100, Hits 1: exact conservation per key
            order = np.argsort(ids, kind="stable")
            sid, st, rem = ids[order], out_whole["status"][order], out_whole["remaining"][order]
            starts = np.r_[0, np.nonzero(sid[1:] != sid[:-1])[0] + 1]
            rank = np.arange(n) - np.repeat(starts, np.diff(np.r_[starts, n]))
            assert np.array_equal(st, (rank >= 100).astype(st.dtype))
            assert np.array_equal(rem, np.maximum(99 - rank, 0))
            assert np.all(out_whole["limit"] == 100) and np.all(out_whole["reset_time"] == now + 60000)
    # Hits = 0 is idempotent and reports what the last hit left
    probe = reqs.copy(); probe["hits"] = 0
    a = whole.submit(probe, clk); b2 = whole.submit(probe, clk)
    assert np.array_equal(a, b2)
    ws, ss = np.sort(whole.scan(), order=["key_xxh64"]), np.sort(split.scan(), order=["key_xxh64"])
    assert np.array_equal(ws, ss)


def test_full_scale_sample_against_oracle(G):
    """100 M-slot table, 10 M resident keys, 65 536-request Zipf batches; the oracle replays only the keys of the sampled
    batches (their whole history), which is enough to check every response of those batches bit for bit."""
    rng = np.random.default_rng(0xB200 + 9)
    n_keys, n, steps = 10_000_000, 65536, 12
    tab = G.Table(100_000_000)
    batches = []
    for b in range(steps):
        ids = zipf_ids(rng, n, n_keys, 1.1)
        batches.append((ids, bench_requests(ids, T0 + b * 7000, mixed=True, dtype=G.REQ_DTYPE)))
    sample = {3, 11}
    watched = np.unique(np.concatenate([batches[b][0] for b in sample]))
    pool = O.Pool(workers=4, cache_size=10**8, now_ms=T0)
    for b, (ids, reqs) in enumerate(batches):
        now = T0 + b * 7000
        got = tab.submit(reqs, G.clock_fill(now))
        mask = np.isin(ids, watched)
        pool.set_now(now)
        want = pool.submit_hashed(np.ascontiguousarray(reqs[mask]))
        assert np.array_equal(got[mask], want), f"batch {b}"
        if b in sample:
            assert mask.all()
