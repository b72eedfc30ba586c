"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical seeded batches.

Bit-exact on every response field (integer work; the leaky bucket's float64 state is compared through its int64
projections in responses and bit-for-bit in the table scan)."""
import os

import numpy as np
import pytest

import oracle_py as O
from golden import reference_kat as K
from kat_player import play_missing_fields, play_scenario
from workloads import T0, adversarial_batch, bench_batch, bench_requests, extreme_batch, key_hashes, zipf_ids

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["pipeline", "fused"])
def eval_path(request, monkeypatch):
    """Every test runs on both evaluation paths of a single table: the four-kernel pipeline (the default of gub_submit*) and the
    persistent kernel k_batch (what rings use; GUB_PATH=fused selects it for a plain table too)."""
    monkeypatch.setenv("GUB_PATH", request.param)
    return request.param


@pytest.fixture(scope="module")
def G():
    import gubernator_b200 as g
    return g


def _cmp(got, want, what=""):
    if not np.array_equal(got, want):
        bad = np.nonzero(got != want)[0]
        i = int(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {len(got)} responses differ; first at {i}: got {got[i]} want {want[i]}")


def _cmp_counters(tab, pool, base=None):
    c, oc = tab.counters(), pool.counters()
    for k in ("over_limit", "cache_hit", "cache_miss"):
        assert c[k] == oc[k], (k, c, oc)


def _check_table_equals_oracle(G, tab, pool):
    """Final-state check: every item the oracle holds is in the device table bit-for-bit, and nothing else is live."""
    items = pool.each()
    scan = tab.scan()
    dev = {(int(s["key_xxh64"]), int(s["key_fnv1"]) >> 8): s for s in scan}
    assert len(dev) == len(scan)
    live_oracle = 0
    for (kx, kf), it in items.items():
        s = dev.get((kx if kx >= 2 else kx + 2, kf >> 8))
        assert s is not None, (kx, kf)
        live_oracle += 1
        assert int(s["limit"]) == it.limit and int(s["duration"]) == it.duration and int(s["stamp"]) == it.stamp, (s, it.limit, it.duration, it.stamp)
        assert int(s["expire_at"]) == it.expire_at
        if it.value_kind == 2:
            assert int(s["algorithm"]) == 1 and int(s["burst"]) == it.burst
            assert np.float64(s["remaining_f"]).view(np.uint64) == np.float64(it.remaining_f).view(np.uint64)
        else:
            assert int(s["algorithm"]) == 0 and int(s["remaining"]) == it.remaining_i and int(s["status"]) == it.status
    assert live_oracle == len(scan)


# ---- the reference's own known-answer tables, through the service mirror ------------------------------------
@pytest.mark.parametrize("sc", K.SCENARIOS, ids=[s["name"] for s in K.SCENARIOS])
def test_functional_scenarios(G, sc):
    inst = G.V1Instance(capacity_slots=4096, now_ms=K.T0)
    play_scenario(inst, sc)


def test_missing_fields_and_batch_cap(G):
    inst = G.V1Instance(capacity_slots=4096, now_ms=K.T0)
    play_missing_fields(inst)
    reqs = [dict(name="n", unique_key=str(i), limit=1, duration=1000, hits=1) for i in range(1001)]
    with pytest.raises(ValueError, match="max size is '1000'"):
        inst.get_rate_limits(reqs)
    assert len(inst.get_rate_limits(reqs[:1000])) == 1000


def test_error_strings_and_order(G):
    inst = G.V1Instance(capacity_slots=4096, now_ms=K.T0)
    pool = O.Pool(now_ms=K.T0)
    reqs = [dict(name="n", unique_key="k", algorithm=5, limit=1, duration=1, hits=1),
            dict(name="n", unique_key="k", behavior=K.GREGORIAN, duration=K.GREG_WEEKS, limit=1, hits=1),
            dict(name="n", unique_key="k2", algorithm=1, behavior=K.GREGORIAN, duration=77, limit=1, hits=1),
            dict(name="n", unique_key="", limit=1), dict(name="", unique_key="k", limit=1)]
    reqs += [dict(name="n", unique_key="same", limit=3, duration=1000, hits=1) for _ in range(5)]  # gubernator.go:203 index order
    got, want = inst.get_rate_limits(reqs), pool.get_rate_limits(reqs)
    assert got == want


def test_update_peer_globals_items(G):
    inst = G.V1Instance(capacity_slots=4096, now_ms=K.T0)
    pool = O.Pool(now_ms=K.T0)
    for key, algo in (("a_b", 0), ("a_c", 1)):
        inst.update_peer_global(key, algo, 5000, 1 - algo, 10, 3, K.T0 + 5000)
        pool.update_peer_global(key.encode(), algo, 5000, 1 - algo, 10, 3, K.T0 + 5000)
    for uk, algo in (("b", 0), ("c", 1)):
        r = dict(name="a", unique_key=uk, algorithm=algo, limit=10, duration=5000, hits=1)
        assert inst.get_rate_limits([r]) == pool.get_rate_limits([r])


def test_random_access_probe_leaves_table_unchanged(G):
    """gub_probe_random_access (bench.py's random-access ceiling) rewrites slots with their own contents: the table a scan
    returns is bit-identical before and after, and requests still evaluate like the oracle's."""
    rng = np.random.default_rng(12)
    tab = G.Table(1 << 12)
    pool = O.Pool(now_ms=T0)
    reqs = adversarial_batch(rng, 3000, 400, T0)
    _cmp(tab.submit(reqs, G.clock_fill(T0)), pool.submit_hashed(reqs), "before the probe")
    before = np.sort(tab.scan(), order=["key_xxh64", "key_fnv1"])
    assert tab.probe_random_access(1 << 18) > 0.0
    after = np.sort(tab.scan(), order=["key_xxh64", "key_fnv1"])
    assert before.tobytes() == after.tobytes()
    reqs = adversarial_batch(rng, 3000, 400, T0 + 7)
    pool.set_now(T0 + 7)
    _cmp(tab.submit(reqs, G.clock_fill(T0 + 7)), pool.submit_hashed(reqs), "after the probe")


# ---- randomized differential tests --------------------------------------------------------------------------
@pytest.mark.parametrize("seed,n_keys,n", [(0, 3, 2000), (1, 40, 6000), (2, 400, 6000), (3, 5000, 20000), (4, 1, 3000), (5, 40, 65536)])
def test_adversarial_differential(G, seed, n_keys, n):
    rng = np.random.default_rng(1000 + seed)
    tab = G.Table(1 << 16, max_batch=65536)
    pool = O.Pool(workers=4, cache_size=10_000_000, now_ms=T0)
    now = T0
    for b in range(8):
        now += int(rng.choice([0, 1, 7, 1000, 31000, 61000, 3_700_000]))
        pool.set_now(now)
        reqs = adversarial_batch(rng, n, n_keys, now)
        _cmp(tab.submit(reqs, G.clock_fill(now)), pool.submit_hashed(reqs), f"batch {b}")
        _cmp_counters(tab, pool)
    _check_table_equals_oracle(G, tab, pool)
    c = tab.counters()
    assert c["requests"] == 8 * n and c["table_full"] == 0


@pytest.mark.parametrize("seed,n_keys", [(0, 5), (1, 300), (2, 4000)])
def test_numeric_extremes_differential(G, seed, n_keys):
    """int64 wrap-around, float64 beyond 2^52 / 2^63, Go's int64(float64) on overflow and NaN (cvt.rzi saturates on the GPU and
    is patched in f2i), negative limits and durations, created_at +-2^62."""
    rng = np.random.default_rng(7000 + seed)
    tab = G.Table(1 << 16)
    pool = O.Pool(workers=4, cache_size=10_000_000, now_ms=T0)
    now = T0
    for b in range(8):
        now += int(rng.choice([0, 1, 1000, 61000]))
        pool.set_now(now)
        reqs = extreme_batch(rng, 12000, n_keys, now)
        _cmp(tab.submit(reqs, G.clock_fill(now)), pool.submit_hashed(reqs), f"batch {b}")
        _cmp_counters(tab, pool)
    _check_table_equals_oracle(G, tab, pool)


def test_config2_uniform_token_1m_keys(G):
    """BASELINE config 2: 1 M keys, 64 k-request batches, TOKEN_BUCKET, uniform; every response diffed."""
    rng = np.random.default_rng(0xB200 + 2)
    n_keys, n = 1_000_000, 65536
    tab = G.Table(2 * n_keys)
    pool = O.Pool(workers=8, cache_size=100_000_000, now_ms=T0)
    for b in range(6):
        now = T0 + b
        pool.set_now(now)
        ids = rng.integers(0, n_keys, n)
        reqs = np.zeros(n, dtype=G.REQ_DTYPE)
        # synthetic 128-bit key fingerprints derived from the id (hashing 1 M strings per batch in Python is slow; the
        # string->hash step is covered by the host tests)
        reqs["key_xxh64"] = (ids.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ np.uint64(0xD6E8FEB86659FD93)
        reqs["key_fnv1"] = (ids.astype(np.uint64) + np.uint64(1)) * np.uint64(0xC2B2AE3D27D4EB4F)
        reqs["hits"] = 1; reqs["limit"] = 100; reqs["duration"] = 60000; reqs["created_at"] = now
        reqs["behavior"] = G.native.REQ_IS_OWNER
        _cmp(tab.submit(reqs, G.clock_fill(now)), pool.submit_hashed(reqs), f"batch {b}")
    _cmp_counters(tab, pool)
    _check_table_equals_oracle(G, tab, pool)


@pytest.mark.parametrize("mixed", [False, True])
def test_zipf_heavy_duplicates(G, mixed):
    """Config-3-shaped traffic at oracle-friendly scale: Zipf 1.1 over 200 k keys, 64 k batches, long same-key runs."""
    rng = np.random.default_rng(0xB200 + 3)
    tab = G.Table(1 << 20)
    pool = O.Pool(workers=8, cache_size=100_000_000, now_ms=T0)
    for b in range(8):
        now = T0 + b * 9000  # hot keys drain and expire across batches (duration 60 s)
        pool.set_now(now)
        reqs, _ = bench_batch(rng, 65536, 200_000, now, zipf_s=1.1, mixed=mixed)
        _cmp(tab.submit(reqs, G.clock_fill(now)), pool.submit_hashed(reqs), f"batch {b}")
    _cmp_counters(tab, pool)
    _check_table_equals_oracle(G, tab, pool)
    c = tab.counters()
    assert c["dup_groups"] > 1000 and c["mixed_groups"] == 0 and c["serial_fallbacks"] == 0


def test_many_segments_and_serial_fallback(G):
    """One hot key whose requests change parameters constantly: > MAX_SEG segments forces the serial walk."""
    rng = np.random.default_rng(77)
    tab = G.Table(4096)
    pool = O.Pool(now_ms=T0)
    n = 5000
    reqs = np.zeros(n, dtype=G.REQ_DTYPE)
    xx, fv = key_hashes([1] * n, name="hot")
    reqs["key_xxh64"], reqs["key_fnv1"] = xx, fv
    reqs["hits"] = rng.integers(0, 3, n); reqs["limit"] = rng.choice([50, 100], n); reqs["duration"] = 60000
    reqs["created_at"] = T0 + rng.integers(0, 5, n); reqs["algorithm"] = 1; reqs["behavior"] = G.native.REQ_IS_OWNER
    _cmp(tab.submit(reqs, G.clock_fill(T0)), pool.submit_hashed(reqs))
    walked = tab.counters()["serial_fallbacks"]  # chunks (<= 512 requests) that were mostly one-request segments: applied one by one
    assert walked >= 1
    # a few long uniform segments (one per simulated RPC timestamp) stay on the planned path
    reqs2 = reqs.copy()
    reqs2["hits"] = 1; reqs2["limit"] = 100
    reqs2["created_at"] = T0 + 10 + (np.arange(n) // 500)
    pool.set_now(T0 + 10)
    _cmp(tab.submit(reqs2, G.clock_fill(T0 + 10)), pool.submit_hashed(reqs2))
    assert tab.counters()["serial_fallbacks"] == walked


def test_token_reset_flipflop_in_heavy_group(G):
    """RESET_REMAINING on a token bucket alternates remove/create: no closed form, exercises the piece-buffer overflow."""
    tab = G.Table(4096)
    pool = O.Pool(now_ms=T0)
    n = 300
    reqs = np.zeros(n, dtype=G.REQ_DTYPE)
    xx, fv = key_hashes([7] * n, name="flip")
    reqs["key_xxh64"], reqs["key_fnv1"] = xx, fv
    reqs["hits"] = 1; reqs["limit"] = 10; reqs["duration"] = 60000; reqs["created_at"] = T0
    reqs["behavior"] = G.native.REQ_IS_OWNER | G.native.RESET_REMAINING
    _cmp(tab.submit(reqs, G.clock_fill(T0)), pool.submit_hashed(reqs))
    _check_table_equals_oracle(G, tab, pool)


def test_ragged_and_empty_batches(G):
    rng = np.random.default_rng(5)
    tab = G.Table(1 << 14, max_batch=4096)
    pool = O.Pool(now_ms=T0)
    assert len(tab.submit(np.zeros(0, dtype=G.REQ_DTYPE), G.clock_fill(T0))) == 0
    for n in (1, 2, 31, 32, 33, 255, 257, 4095, 4096, 4097, 10000):  # 4097 and 10000 cross the max_batch chunking
        reqs = adversarial_batch(rng, n, 50, T0)
        _cmp(tab.submit(reqs, G.clock_fill(T0)), pool.submit_hashed(reqs), f"n={n}")
    _check_table_equals_oracle(G, tab, pool)


def test_sentinel_key_hashes(G):
    """XXH64 values 0 and 1 collide with the table's empty/tombstone sentinels and are remapped."""
    tab = G.Table(4096)
    pool = O.Pool(now_ms=T0)
    reqs = np.zeros(8, dtype=G.REQ_DTYPE)
    reqs["key_xxh64"] = [0, 1, 2, 3, 0, 1, 2, 3]
    reqs["key_fnv1"] = [10 << 8, 11 << 8, 12 << 8, 13 << 8, 10 << 8, 11 << 8, 12 << 8, 13 << 8]
    reqs["hits"] = 1; reqs["limit"] = 5; reqs["duration"] = 1000; reqs["created_at"] = T0
    got = tab.submit(reqs, G.clock_fill(T0))
    # 0/2 and 1/3 are distinct keys here because their FNV tags differ
    _cmp(got, pool.submit_hashed(reqs))


def test_keys_colliding_in_the_grouping_table(G):
    """Keys that collide in the batch-wide group table (same home entry: linear probing there) and two keys that share the XXH64 but
    not the FNV-1 (one group entry, told apart by the request compare: the group takes the segment walk, key by key)."""
    rng = np.random.default_rng(2024)
    keys = rng.integers(2, 1 << 62, 4000).astype(np.uint64)
    home = ((keys * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)) & np.uint64((1 << 18) - 1)
    order = np.argsort(home, kind="stable")
    same = np.nonzero(np.diff(home[order].astype(np.int64)) == 0)[0]
    ka, kb = keys[order[same[0]]], keys[order[same[0] + 1]]
    assert ka != kb
    tab = G.Table(4096)
    pool = O.Pool(now_ms=T0)
    for step in range(3):
        now = T0 + step
        pool.set_now(now)
        n = 1500
        reqs = np.zeros(n, dtype=G.REQ_DTYPE)
        pick = rng.integers(0, 4, n)
        reqs["key_xxh64"] = np.where(pick == 0, ka, np.where(pick == 1, kb, keys[rng.integers(0, 50, n)]))
        reqs["key_fnv1"] = (reqs["key_xxh64"] * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(0x100)
        twin = pick == 3  # same XXH64 as ka, another FNV-1: a different key as far as the table is concerned
        reqs["key_xxh64"][twin] = ka
        reqs["key_fnv1"][twin] = np.uint64(0x7777700)
        reqs["hits"] = 1 if step < 2 else rng.integers(0, 3, n)
        reqs["limit"] = 500; reqs["duration"] = 60000; reqs["created_at"] = now
        reqs["algorithm"] = (reqs["key_xxh64"] & np.uint64(1)).astype(np.uint32); reqs["behavior"] = G.native.REQ_IS_OWNER
        _cmp(tab.submit(reqs, G.clock_fill(now)), pool.submit_hashed(reqs), f"step {step}")  # (the oracle keys its cache by both hashes)
    assert tab.counters()["mixed_groups"] > 0


def test_full_window_evicts_like_the_lru(G):
    """A table far too small for the batch: the reference's LRU evicts (lrucache.go:98,138-149) and still answers every request.
    So do we: keys whose probe window is full are parked and placed, with eviction, before the next batch reads the table."""
    tab = G.Table(64)
    n = 1500
    reqs = np.zeros(n, dtype=G.REQ_DTYPE)
    xx, fv = key_hashes(np.arange(n), name="full")
    reqs["key_xxh64"], reqs["key_fnv1"] = xx, fv
    reqs["hits"] = 1; reqs["limit"] = 5; reqs["duration"] = 100000; reqs["created_at"] = T0 + np.arange(n)
    big = O.Pool(now_ms=T0)
    _cmp(tab.submit(reqs, G.clock_fill(T0)), big.submit_hashed(reqs), "every request answered as if the cache had room")
    c = tab.counters()
    assert tab.size() == 64 and c["inserts"] == 64 and c["table_full"] == n - 64 - 1024
    again = reqs[-8:].copy()
    again["created_at"] = T0 + 5
    out = tab.submit(again, G.clock_fill(T0 + 5))
    c = tab.counters()
    assert c["unexpired_evictions"] == 1024 and np.all(out["err_code"] == 0) and tab.size() == 64
    # expired items are reclaimed by the sweep and the slots reused
    assert tab.sweep(T0 + 300000) == 64 and tab.size() == 0
    out = tab.submit(reqs[:64], G.clock_fill(T0 + 300000))
    assert np.all(out["err_code"] == 0)


def test_add_get_scan_items(G):
    rng = np.random.default_rng(9)
    tab = G.Table(1 << 12)
    n = 500
    items = np.zeros(n, dtype=G.ITEM_DTYPE)
    xx, fv = key_hashes(np.arange(n), name="items")
    items["key_xxh64"], items["key_fnv1"] = xx, fv
    items["algorithm"] = np.arange(n) & 1
    items["limit"] = rng.integers(1, 100, n); items["duration"] = 60000; items["remaining"] = rng.integers(0, 50, n)
    items["remaining_f"] = rng.random(n) * 50; items["stamp"] = T0; items["burst"] = items["limit"]; items["expire_at"] = T0 + 60000
    items["status"] = rng.integers(0, 2, n)
    tab.add_items(items)
    assert tab.size() == n
    got, found = tab.get_items(xx, fv, T0)
    assert found.all()
    leaky = items["algorithm"] == 1
    assert np.array_equal(got["limit"], items["limit"]) and np.array_equal(got["expire_at"], items["expire_at"])
    assert np.array_equal(got["remaining"][~leaky], items["remaining"][~leaky])
    assert np.array_equal(got["status"][~leaky], items["status"][~leaky])
    assert np.array_equal(got["remaining_f"][leaky].view(np.uint64), items["remaining_f"][leaky].view(np.uint64))
    # expiry is strict: now == ExpireAt is live, now > ExpireAt is a miss (cache.go:52)
    assert tab.get_items(xx[:4], fv[:4], T0 + 60000)[1].all() and not tab.get_items(xx[:4], fv[:4], T0 + 60001)[1].any()
    # unknown keys
    assert not tab.get_items(xx[:4] + np.uint64(12345), fv[:4], T0)[1].any()
    # overwrite: last one wins, also within one call
    again = items[:10].copy(); again["limit"] = 777
    twice = np.concatenate([items[:10], again])
    tab.add_items(twice)
    assert tab.size() == n and np.all(tab.get_items(xx[:10], fv[:10], T0)[0]["limit"] == 777)
    # requests against loaded state behave like the oracle given the same items
    pool = O.Pool(now_ms=T0)
    for i in range(n):
        it = O.Item()
        it.algorithm = int(items["algorithm"][i]); it.value_kind = 2 if leaky[i] else 1
        it.expire_at = int(items["expire_at"][i]); it.status = int(items["status"][i]) if not leaky[i] else 0
        it.limit = 777 if i < 10 else int(items["limit"][i]); it.duration = 60000; it.remaining_i = int(items["remaining"][i])
        it.remaining_f = float(items["remaining_f"][i]); it.stamp = T0; it.burst = int(items["burst"][i]) if leaky[i] else 0
        pool.add_item_hashed(xx[i], fv[i], it)
    reqs = np.zeros(n, dtype=G.REQ_DTYPE)
    reqs["key_xxh64"], reqs["key_fnv1"] = xx, fv
    reqs["hits"] = 2; reqs["limit"] = items["limit"]; reqs["limit"][:10] = 777; reqs["duration"] = 60000
    reqs["created_at"] = T0 + 5; reqs["algorithm"] = items["algorithm"]
    pool.set_now(T0 + 5)
    _cmp(tab.submit(reqs, G.clock_fill(T0 + 5)), pool.submit_hashed(reqs))


def test_async_pipeline_matches_sync(G):
    rng = np.random.default_rng(21)
    tab = G.Table(1 << 16)
    pool = O.Pool(now_ms=T0)
    depth = 4
    n = 8192
    bufs = [(G.native.PinnedArray(n, G.REQ_DTYPE), G.native.PinnedArray(n, G.RESP_DTYPE)) for _ in range(depth)]
    want, tickets = [], []
    for b in range(12):
        rq, rs = bufs[b % depth]
        if b >= depth:
            tab.wait(tickets[b - depth])
            _cmp(rs.array.copy(), want[b - depth], f"batch {b - depth}")
        reqs = adversarial_batch(rng, n, 300, T0)
        rq.array[:] = reqs
        want.append(pool.submit_hashed(reqs))
        tickets.append(tab.submit_async(rq.ptr, n, G.clock_fill(T0), rs.ptr))
    for b in range(12 - depth, 12):
        tab.wait(tickets[b])
        _cmp(bufs[b % depth][1].array.copy(), want[b], f"batch {b}")
    for rq, rs in bufs:
        rq.free(); rs.free()


def test_compact_requests_equal_full_records(G):
    """gub_submit_compact (32-byte records + a parameter table, expanded on the device) == gub_submit of the same batch."""
    rng = np.random.default_rng(61)
    a, b = G.Table(1 << 16), G.Table(1 << 16)
    pool = O.Pool(now_ms=T0)
    for step in range(4):
        now = T0 + 500 * step
        pool.set_now(now)
        reqs = adversarial_batch(rng, [5000, 1, 65536, 3000][step], 200, now)
        creqs, params, base = G.native.compact_batch(reqs)
        clk = G.clock_fill(now)
        want = pool.submit_hashed(reqs)
        _cmp(a.submit(reqs, clk), want, f"full step {step}")
        _cmp(b.submit_compact(creqs, params, base, clk), want, f"compact step {step}")
    # tables of up to 32 parameter sets travel in the kernel arguments (k_expand_inline), larger ones by a device copy: both
    # sides of the boundary, and the two-set table of the bench workload
    for n_sets in (2, 32, 33):
        now = T0 + 5000 + n_sets
        pool.set_now(now)
        ids = zipf_ids(rng, 20000, 3000, 1.1)
        reqs = bench_requests(ids, now)
        reqs["limit"] = 50 + (ids % n_sets)
        reqs["algorithm"] = (ids % n_sets) & 1
        creqs, params, base = G.native.compact_batch(reqs)
        assert len(params) == n_sets
        clk = G.clock_fill(now)
        want = pool.submit_hashed(reqs)
        _cmp(a.submit(reqs, clk), want, f"full, {n_sets} sets")
        _cmp(b.submit_compact(creqs, params, base, clk), want, f"compact, {n_sets} sets")
    # an out-of-range parameter index is an in-band error, nothing is stored
    creqs, params, base = G.native.compact_batch(adversarial_batch(rng, 10, 3, T0))
    creqs["params"][3] = 10_000
    out = b.submit_compact(creqs, params, base, G.clock_fill(T0))
    assert out["err_code"][3] == G.native.ERR_INVALID_ALGORITHM


def test_submit_device_and_route(G):
    """Device-resident buffers (torch is only the allocator) + the ring routing kernels vs the oracle ring."""
    import torch
    rng = np.random.default_rng(31)
    dev = torch.device("cuda:0")
    tab = G.Table(1 << 16)
    pool = O.Pool(now_ms=T0)
    n = 30000
    reqs = adversarial_batch(rng, n, 2000, T0)
    d_reqs = torch.from_numpy(reqs.view(np.uint8).reshape(n, 64)).to(dev)
    d_out = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    tab.submit_device(d_reqs.data_ptr(), n, G.clock_fill(T0), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    _cmp(d_out.cpu().numpy().reshape(-1).view(G.RESP_DTYPE), pool.submit_hashed(reqs))

    for nshards in (1, 2, 3, 8):
        ring, oring = G.Ring(0, 512), O.Ring(0, 512)
        for g in range(nshards):
            ring.add(f"gpu:{g}"); oring.add(f"gpu:{g}")
        d_routed = torch.zeros_like(d_reqs)
        d_perm = torch.zeros(n, dtype=torch.int32, device=dev)
        d_counts = torch.zeros(16, dtype=torch.int32, device=dev)
        tab.route_device(ring, d_reqs.data_ptr(), n, d_routed.data_ptr(), d_perm.data_ptr(), d_counts.data_ptr(), st)
        torch.cuda.synchronize()
        owner = np.array([oring.get_by_hash(int(h)) for h in reqs["key_fnv1"]])
        order = np.argsort(owner, kind="stable")  # stable partition by owner == what the reference's per-peer queues see
        assert np.array_equal(d_perm.cpu().numpy().astype(np.int64), order)
        assert np.array_equal(d_counts.cpu().numpy()[:nshards], np.bincount(owner, minlength=nshards))
        assert np.array_equal(d_routed.cpu().numpy().reshape(-1).view(G.REQ_DTYPE), reqs[order])
        # unroute restores request order
        d_back = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
        d_resp_routed = d_out[torch.from_numpy(order).to(dev)].contiguous()
        tab.unroute_device(d_resp_routed.data_ptr(), d_perm.data_ptr(), n, d_back.data_ptr(), st)
        torch.cuda.synchronize()
        assert torch.equal(d_back, d_out)


def test_epoch_wrap(G):
    """The per-batch grouping table tags entries with a 16-bit epoch; run past a wrap with tiny batches."""
    import torch
    rng = np.random.default_rng(41)
    tab = G.Table(1 << 12, max_batch=1024)
    pool = O.Pool(now_ms=T0)
    clk = G.clock_fill(T0)
    reqs = adversarial_batch(rng, 64, 5, T0)
    d_reqs = torch.from_numpy(reqs.view(np.uint8).reshape(-1, 64)).cuda()
    d_out = torch.zeros((64, 32), dtype=torch.uint8, device="cuda")
    checkpoints = {0, 1, 65530, 65533, 65534, 65535, 65536, 65599}
    st = torch.cuda.current_stream().cuda_stream
    for b in range(65600):
        tab.submit_device(d_reqs.data_ptr(), 64, clk, d_out.data_ptr(), st)
        want = pool.submit_hashed(reqs)
        if b in checkpoints:
            torch.cuda.synchronize()
            _cmp(d_out.cpu().numpy().reshape(-1).view(G.RESP_DTYPE), want, f"batch {b}")


def test_scheduled_sweep_between_batches(G):
    """The library's incremental sweep on the pipeline path (every 1024 batches 1/64 of the table, k_sweep on the batch's stream):
    with keys that expire while the clock advances, 2200 batches trigger it twice; every response still equals the oracle's (an
    expired key is a miss with or without the sweep) and entries were reclaimed."""
    rng = np.random.default_rng(23)
    tab = G.Table(1 << 10, max_batch=1024)  # 64 slices of 16 slots: sweeps 1 and 2 cover slots 0..31
    tab.set_sweep(1)
    pool = O.Pool(now_ms=T0)
    swept_before = tab.counters()["swept"]
    for b in range(2200):
        now = T0 + 40 * b
        pool.set_now(now)
        reqs = bench_requests(rng.integers(0, 300, 48), now)
        reqs["duration"] = 100  # expires after 100 ms: most resident keys are dead by the time a slice is swept
        got = tab.submit(reqs, G.clock_fill(now))
        want = pool.submit_hashed(reqs)
        if b % 97 == 0 or b > 2040:
            _cmp(got, want, f"batch {b}")
    if os.environ.get("GUB_PATH") != "fused":
        assert tab.counters()["swept"] > swept_before


def test_invalid_at_of_loaded_items(G):
    """CacheItem.InvalidAt (cache.go:40,47) — only ever set by Store / Loader plugins — travels with gub_add_items: an item past its
    InvalidAt is a miss for GetCacheItem and for the batch path (removed and re-created, like the oracle given the same items),
    one that is not yet invalid keeps its InvalidAt across updates, and Store (gub_scan) reports it."""
    n = 40
    tab, pool = G.Table(1 << 12), O.Pool(now_ms=T0)
    xx, fv = key_hashes(np.arange(n), name="inv")
    items = np.zeros(n, dtype=G.ITEM_DTYPE)
    items["key_xxh64"], items["key_fnv1"] = xx, fv
    items["algorithm"] = np.arange(n) & 1
    items["limit"] = 10; items["duration"] = 60000; items["remaining"] = 4; items["remaining_f"] = 4.0; items["stamp"] = T0
    items["burst"] = 10; items["expire_at"] = T0 + 60000
    items["invalid_at"] = np.where(np.arange(n) % 4 == 0, 0, T0 + 1000 * (np.arange(n) % 4))  # none, +1 s, +2 s, +3 s
    tab.add_items(items)
    for i in range(n):
        it = O.Item()
        it.algorithm = int(items["algorithm"][i]); it.value_kind = 2 if items["algorithm"][i] else 1
        it.expire_at = T0 + 60000; it.invalid_at = int(items["invalid_at"][i]); it.limit = 10; it.duration = 60000
        it.remaining_i = 4; it.remaining_f = 4.0; it.stamp = T0; it.burst = 10 if items["algorithm"][i] else 0
        pool.add_item_hashed(xx[i], fv[i], it)
    got, found = tab.get_items(xx, fv, T0 + 1500)
    assert np.array_equal(found.astype(bool), ~((items["invalid_at"] != 0) & (items["invalid_at"] < T0 + 1500)))
    assert np.array_equal(got["invalid_at"], items["invalid_at"])
    assert np.array_equal(np.sort(tab.scan()["invalid_at"]), np.sort(items["invalid_at"]))
    reqs = np.zeros(2 * n, dtype=G.REQ_DTYPE)
    reqs["key_xxh64"], reqs["key_fnv1"] = np.tile(xx, 2), np.tile(fv, 2)
    reqs["hits"] = 1; reqs["limit"] = 10; reqs["duration"] = 60000; reqs["algorithm"] = np.tile(items["algorithm"], 2).astype(np.uint32)
    reqs["behavior"] = G.native.REQ_IS_OWNER
    for now in (T0 + 1500, T0 + 2500, T0 + 9000):
        reqs["created_at"] = now
        pool.set_now(now)
        _cmp(tab.submit(reqs, G.clock_fill(now)), pool.submit_hashed(reqs), f"now = T0 + {now - T0}")
    _cmp_counters(tab, pool)


def test_device_key_hashing_matches_host(G):
    """XXH64 + FNV-1 of packed key strings on the device == the host implementation (itself pinned to python-xxhash and the
    reference's ring golden vector): every length 0..200 incl. the >= 32-byte XXH64 path, and the BASELINE key format."""
    import torch
    rng = np.random.default_rng(5)
    keys = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in range(0, 201)] + [f"bench_k{i:09d}".encode() for i in range(5000)]
    keys += [b"test_over_limit_account:1234", b"a" * 1000]
    offs = np.zeros(len(keys) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(k) for k in keys])
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8)
    tab = G.Table(1024)
    d_blob, d_offs = torch.from_numpy(blob.copy()).cuda(), torch.from_numpy(offs.view(np.int64).copy()).cuda()
    d_xx = torch.zeros(len(keys), dtype=torch.int64, device="cuda"); d_fv = torch.zeros_like(d_xx)
    d_reqs = torch.zeros((len(keys), 64), dtype=torch.uint8, device="cuda")
    tab.hash_keys_device(d_blob.data_ptr(), d_offs.data_ptr(), len(keys), d_xx.data_ptr(), d_fv.data_ptr(), d_reqs.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    xx, fv = G.hash_keys(keys)
    assert np.array_equal(d_xx.cpu().numpy().view(np.uint64), xx) and np.array_equal(d_fv.cpu().numpy().view(np.uint64), fv)
    r = d_reqs.cpu().numpy().reshape(-1).view(G.REQ_DTYPE)
    assert np.array_equal(r["key_xxh64"], xx) and np.array_equal(r["key_fnv1"], fv)
    # anchored on the oracle's own restatement of both hashes (pinned to python-xxhash / the reference's ring vector), every key
    assert all(int(a) == O.xxh64(k) and int(b) == O.fnv1_64(k) for a, b, k in zip(xx, fv, keys))


def test_submit_keys_hashes_on_the_device(G):
    """gub_submit_keys_async: key strings in, responses out.  Same answers as hashing with the oracle's XXH64 / FNV-1 on the host
    and evaluating the full records in the oracle; <= 32 parameter sets travel in the launch, more by pointer."""
    rng = np.random.default_rng(71)
    tab, pool = G.Table(1 << 14), O.Pool(now_ms=T0)
    for step, n_sets in enumerate((3, 40)):
        now = T0 + 50 * step
        pool.set_now(now)
        n = 6000
        ids = zipf_ids(rng, n, 1500, 1.1)
        keys = [f"bench_k{int(i):09d}".encode() if i % 7 else (b"x" * (int(i) % 90)) + f"{int(i)}".encode() for i in ids]
        reqs = bench_requests(ids, now)
        reqs["key_xxh64"] = [O.xxh64(k) for k in keys]
        reqs["key_fnv1"] = [O.fnv1_64(k) for k in keys]
        reqs["limit"] = 50 + (ids % n_sets); reqs["algorithm"] = (ids % n_sets) & 1
        reqs["created_at"] = now - (ids % 3)
        want = pool.submit_hashed(reqs)
        _, params, base = G.native.compact_batch(reqs.astype(G.REQ_DTYPE))
        creqs, _, _ = G.native.compact_batch(reqs.astype(G.REQ_DTYPE))
        packed = G.native.pack_keys(keys, reqs["hits"], creqs["params"], creqs["created_delta"])
        pin = G.native.PinnedArray(len(packed), np.uint8); pin.array[:] = packed
        ppin = G.native.PinnedArray(len(params), G.native.PARAMS_DTYPE); ppin.array[:] = params
        out = G.native.PinnedArray(n, G.RESP_DTYPE)
        tk = tab.submit_keys_async(pin.ptr, len(packed), n, ppin.ptr, len(params), base, G.clock_fill(now), out.ptr)
        tab.wait(tk)
        _cmp(out.array.copy(), want, f"{n_sets} parameter sets")
        for a in (pin, ppin, out):
            a.free()


def test_rpc_aggregator_coalesces_concurrent_calls(G):
    """Many threads call GetRateLimits (<= 1000 requests each) at once; the aggregator serves them from shared device batches.
    Per-thread key spaces are disjoint, so every thread must see exactly what the oracle gives for its own call sequence,
    whatever the interleaving; a shared key checks conservation across threads."""
    from concurrent.futures import ThreadPoolExecutor
    inst = G.V1Instance(capacity_slots=1 << 16, now_ms=K.T0)
    agg = inst.aggregator(max_batch=8192, window_us=2000)
    n_threads, calls, per_call = 16, 6, 300

    def worker(t):
        pool = O.Pool(now_ms=K.T0)
        rng = np.random.default_rng(t)
        shared_under = 0
        for c in range(calls):
            reqs = [dict(name=f"agg{t}", unique_key=f"k{int(rng.integers(0, 40))}", algorithm=int(rng.integers(0, 2)), limit=25, duration=60000,
                         hits=int(rng.integers(0, 3))) for _ in range(per_call)]
            got = agg.get_rate_limits(reqs)
            want = pool.get_rate_limits(reqs)
            assert got == want, (t, c)
            sh = agg.get_rate_limits([dict(name="shared", unique_key="one", limit=100, duration=60000, hits=1)])[0]
            shared_under += sh["status"] == 0
        return shared_under
    with ThreadPoolExecutor(n_threads) as ex:
        under = sum(ex.map(worker, range(n_threads)))
    assert under == min(100, n_threads * calls)  # 96 single hits against a limit of 100: all under the limit, none lost
    st = agg.stats()
    assert st["requests"] == n_threads * calls * (per_call + 1)
    assert st["batches"] < n_threads * calls * 2  # calls really were coalesced
    with pytest.raises(ValueError):
        agg.get_rate_limits([dict(name="n", unique_key=str(i), limit=1, duration=1, hits=1) for i in range(1001)])
    agg.close()


class _MockStore:
    """MockStore2 of the reference (mock_store_test.go:28): records the calls, returns programmed items."""

    def __init__(self):
        self.calls, self.items = [], {}

    def get(self, req, key):
        self.calls.append(("Get", key))
        return self.items.get(key)

    def on_change(self, req, key, item):
        self.calls.append(("OnChange", key, item))

    def remove(self, key):
        self.calls.append(("Remove", key))


def test_aggregator_honours_the_store_plugin(G):
    """Calls coalesced by the aggregator still reach the Store plugin (the aggregator used to submit straight to the
    table): a miss pulls from the store, every call reports OnChange for its key, in arrival order."""
    inst = G.V1Instance(capacity_slots=4096, now_ms=K.T0)
    st = _MockStore()
    inst.set_store(st)
    st.items["n_a"] = dict(algorithm=0, limit=10, duration=1000, remaining=4, remaining_f=4.0, stamp=K.T0, burst=10, expire_at=K.T0 + 1000)
    agg = inst.aggregator(max_batch=4096, window_us=200)
    r = agg.get_rate_limits([dict(name="n", unique_key="a", algorithm=0, duration=1000, limit=10, hits=1)])[0]
    assert (r["status"], r["remaining"]) == (0, 3)  # the stored item (4 left) was loaded, not a fresh bucket
    assert [c[0] for c in st.calls] == ["Get", "OnChange"] and st.calls[1][2]["remaining"] == 3
    st.calls.clear()
    r = agg.get_rate_limits([dict(name="n", unique_key="a", algorithm=0, duration=1000, limit=10, hits=1),
                             dict(name="n", unique_key="b", algorithm=0, duration=1000, limit=10, hits=2)])
    assert [x["remaining"] for x in r] == [2, 8]
    assert [c[:2] for c in st.calls] == [("Get", "n_b"), ("OnChange", "n_a"), ("OnChange", "n_b")]
    agg.close()


@pytest.mark.parametrize("algo", [0, 1])
def test_store_plugin_call_sequences(G, algo):
    """store_test.go:127-533 TestStore: which Store methods run, in which order, with which item, for a cache miss, a cache
    hit, an item found in the store, an algorithm switch, a duration change and a duration change that expires the item."""
    key = "test_over_limit_account:1234"
    req = dict(name="test_over_limit", unique_key="account:1234", algorithm=algo, duration=1000, limit=10, hits=1)

    def setup():
        inst = G.V1Instance(capacity_slots=4096, now_ms=K.T0)
        st = _MockStore()
        inst.set_store(st)
        return inst, st

    def item(duration=1000, stamp=K.T0, expire=None, algorithm=algo):
        return dict(algorithm=algorithm, limit=10, duration=duration, remaining=10, remaining_f=10.0, stamp=stamp, burst=10,
                    expire_at=expire if expire is not None else stamp + duration)

    # First rate check pulls from store (miss) -> Get, OnChange; second comes from the cache -> OnChange only (:228-271)
    inst, st = setup()
    r = inst.get_rate_limits([req])[0]
    assert (r["limit"], r["status"]) == (10, 0)
    assert [c[0] for c in st.calls] == ["Get", "OnChange"] and st.calls[1][2]["limit"] == 10 and st.calls[1][2]["duration"] == 1000
    st.calls.clear()
    inst.get_rate_limits([req])
    assert [c[0] for c in st.calls] == ["OnChange"]

    # Found in store after cache miss (:273-308)
    inst, st = setup()
    st.items[key] = item()
    r = inst.get_rate_limits([req])[0]
    assert (r["limit"], r["status"], r["remaining"]) == (10, 0, 9)
    assert [c[0] for c in st.calls] == ["Get", "OnChange"]

    # Algorithm changed: the stored item is of another type -> Remove, then OnChange with the new item (:310-348)
    inst, st = setup()
    st.items[key] = item(algorithm=1 - algo)
    r = inst.get_rate_limits([req])[0]
    assert (r["limit"], r["status"]) == (10, 0)
    assert [c[0] for c in st.calls] == ["Get", "Remove", "OnChange"] and st.calls[2][2]["algorithm"] == algo

    if algo == 0:
        # Duration changed (:353-439): ExpireAt == CreatedAt + newDuration
        inst, st = setup()
        st.items[key] = item(duration=5000)
        req2 = dict(req, duration=8000)
        r = inst.get_rate_limits([req2])[0]
        assert (r["limit"], r["status"]) == (10, 0)
        it = st.calls[-1][2]
        assert st.calls[-1][0] == "OnChange" and it["expire_at"] == it["stamp"] + 8000 and it["duration"] == 8000 and it["limit"] == 10
        # Duration changed and immediately expired (:441-531): renewed expiration, remaining reset
        inst, st = setup()
        long_ago = K.T0 - 100000
        st.items[key] = item(duration=500000, stamp=long_ago, expire=long_ago + 500000)
        r = inst.get_rate_limits([req2])[0]
        assert (r["limit"], r["status"]) == (10, 0)
        it = st.calls[-1][2]
        assert it["expire_at"] == it["stamp"] + 8000 and it["stamp"] == K.T0 and it["duration"] == 8000
