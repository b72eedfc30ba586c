"""Seeded synthetic request generators shared by the CPU and GPU parity tests and by bench.py."""
import numpy as np

import oracle_py as O

T0 = 1_700_000_000_000


def make_clock(now_ms, clock_dtype=None):
    """gub_clock for a batch, built from the ORACLE's Gregorian functions (tests only)."""
    from _host_math import CLOCK_DTYPE
    clk = np.zeros(1, dtype=clock_dtype or CLOCK_DTYPE)
    clk["now_ms"] = now_ms
    for d in range(6):
        clk["greg_expire"][0][d] = O.gregorian_expiration(now_ms, d)[0]
        clk["greg_duration"][0][d] = O.gregorian_duration(now_ms, d)[0]
    return clk


_KEY_CACHE = {}


def key_hashes(ids, name="bench"):
    """(xxh64, fnv1) of Name + "_" + "k%09d" % id for each id (BASELINE.md synthetic keys)."""
    ids = np.asarray(ids, dtype=np.int64)
    xx = np.zeros(len(ids), dtype=np.uint64)
    fv = np.zeros(len(ids), dtype=np.uint64)
    for j, i in enumerate(ids.tolist()):
        k = (name, i)
        v = _KEY_CACHE.get(k)
        if v is None:
            s = f"{name}_k{i:09d}".encode()
            v = (O.xxh64(s), O.fnv1_64(s))
            if len(_KEY_CACHE) < 2_000_000:
                _KEY_CACHE[k] = v
        xx[j], fv[j] = v
    return xx, fv


def adversarial_batch(rng, n, n_keys, now_ms, uniform_run_prob=0.3, p_weird=0.15):
    """Few keys, many duplicates, every behaviour / edge the reference's algorithms branch on.

    Runs of identical requests are mixed with fully random ones so that both the run planner and the serial paths are
    exercised.  Per key the algorithm is mostly fixed (switching resets the bucket) but sometimes switches."""
    reqs = np.zeros(n, dtype=O.HREQ_DTYPE)
    ids = rng.integers(0, n_keys, n)
    xx, fv = key_hashes(ids, name="adv")
    reqs["key_xxh64"], reqs["key_fnv1"] = xx, fv
    limits = np.array([0, 1, 2, 5, 10, 100, 2000, -3, 1 << 40])
    durs = np.array([0, 1, 5, 1000, 30000, 60000, 3600000, -50])
    hits_c = np.array([0, 1, 1, 1, 2, 3, 7, 100, -1, -2, 1 << 41])
    bursts = np.array([0, 0, 0, 5, 20, 3000])
    behs = np.array([0, 0, 0, 0, O.RESET_REMAINING, O.DRAIN_OVER_LIMIT, O.DRAIN_OVER_LIMIT, O.NO_BATCHING | O.GLOBAL,
                     O.DURATION_IS_GREGORIAN])
    i = 0
    # a per-key "usual" parameter set
    usual = {}
    while i < n:
        kid = int(ids[i])
        if kid not in usual:
            usual[kid] = dict(limit=int(rng.choice(limits[1:7])), duration=int(rng.choice(durs[3:7])),
                              burst=int(rng.choice(bursts)), algorithm=int(rng.integers(0, 2)), hits=int(rng.choice([1, 1, 1, 2, 3])),
                              behavior=int(rng.choice([0, 0, O.DRAIN_OVER_LIMIT])))
        u = usual[kid]
        r = dict(u)
        r["created_at"] = now_ms + int(rng.choice([0, 0, 0, 1, -1, 2]))
        if rng.random() < p_weird:
            r["limit"] = int(rng.choice(limits)); r["duration"] = int(rng.choice(durs)); r["hits"] = int(rng.choice(hits_c))
            r["burst"] = int(rng.choice(bursts)); r["behavior"] = int(rng.choice(behs))
            if rng.random() < 0.2:
                r["algorithm"] = int(rng.choice([0, 1, 1 - u["algorithm"], 2, 7]))
            if r["behavior"] & O.DURATION_IS_GREGORIAN:
                r["duration"] = int(rng.choice([0, 0, 1, 2, 3, 4, 5, 6, 99, -1]))
            if rng.random() < 0.15:
                r["created_at"] = now_ms + int(rng.choice([-100000, -3600001, 50000, 10**9]))
        owner = O.REQ_IS_OWNER if rng.random() < 0.8 else 0

        def put(j):
            reqs[j]["hits"] = r["hits"]; reqs[j]["limit"] = r["limit"]; reqs[j]["duration"] = r["duration"]
            reqs[j]["burst"] = r["burst"]; reqs[j]["created_at"] = r["created_at"]; reqs[j]["algorithm"] = r["algorithm"]
            reqs[j]["behavior"] = r["behavior"] | owner
        put(i)
        i += 1
        if rng.random() < uniform_run_prob:
            # repeat the identical request on later occurrences of the same key (contiguous in key order, not in index order)
            reps = int(rng.integers(1, 40))
            j = i
            while reps > 0 and j < n:
                if ids[j] == kid:
                    put(j)
                    reps -= 1
                j += 1
    return reqs


def bench_batch(rng, n, n_keys, created_at, zipf_s=None, mixed=False, perm_seed=12345, name="bench"):
    """BASELINE.md synthetic traffic: hits=1, limit=100, duration=60000, burst=0; uniform or Zipf(s) ids; algorithm fixed
    per key (odd id -> LEAKY) when `mixed`."""
    if zipf_s is None:
        ids = rng.integers(0, n_keys, n)
    else:
        ids = zipf_ids(rng, n, n_keys, zipf_s, perm_seed)
    reqs = np.zeros(n, dtype=O.HREQ_DTYPE)
    reqs["key_xxh64"], reqs["key_fnv1"] = key_hashes(ids, name=name)
    reqs["hits"] = 1; reqs["limit"] = 100; reqs["duration"] = 60000; reqs["burst"] = 0
    reqs["created_at"] = created_at
    reqs["algorithm"] = (ids & 1) if mixed else 0
    reqs["behavior"] = O.REQ_IS_OWNER
    return reqs, ids


_ZIPF_CACHE = {}


def zipf_ranks(rng, n, n_keys, s):
    """Bounded Zipf(s) over ranks 0..n_keys-1 (0 = hottest): exact inverse CDF for the first 2^20 ranks (cumulative table), the
    midpoint-rule continuous tail beyond."""
    M = min(n_keys, 1 << 20)
    key = (n_keys, s)
    if key not in _ZIPF_CACHE:
        head = np.cumsum(np.arange(1, M + 1, dtype=np.float64) ** (-s))
        a = 1.0 - s
        tail = ((n_keys + 0.5) ** a - (M + 0.5) ** a) / a if n_keys > M else 0.0
        _ZIPF_CACHE[key] = (head, tail)
    head, tail = _ZIPF_CACHE[key]
    total = head[-1] + tail
    u = rng.random(n) * total
    rank = np.searchsorted(head, u, side="right").astype(np.int64)  # 0-based rank for the head
    in_tail = u >= head[-1]
    if in_tail.any():
        a = 1.0 - s
        x = ((u[in_tail] - head[-1]) * a + (M + 0.5) ** a) ** (1.0 / a)  # invert the tail integral
        rank[in_tail] = np.clip(np.floor(x + 0.5).astype(np.int64) - 1, M, n_keys - 1)
    return np.minimum(rank, n_keys - 1)


def spread_ranks(rank, n_keys, perm_seed=12345):
    """rank -> id through a fixed multiplicative permutation (so hot keys are spread over the id space)."""
    mult = 0x9E3779B97F4A7C15
    with np.errstate(over="ignore"):
        return ((rank.astype(np.uint64) * np.uint64(mult) + np.uint64(perm_seed)) % np.uint64(n_keys)).astype(np.int64)


def zipf_ids(rng, n, n_keys, s, perm_seed=12345):
    """Bounded Zipf(s) key ids: zipf_ranks() spread over the id space by spread_ranks()."""
    return spread_ranks(zipf_ranks(rng, n, n_keys, s), n_keys, perm_seed)


# ---- vectorised hashing of the BASELINE.md synthetic keys -----------------------------------------------------
# Key string = "bench_k%09d" (16 bytes).  XXH64 (seed 0, the < 32-byte path) and FNV-1 64 over 16 bytes, in numpy
# uint64 arithmetic (wrapping), so that 10^8 keys can be prepared in seconds.  Checked against the scalar
# implementations in tests/test_workloads.py.
_P1, _P2, _P3, _P4, _P5 = (np.uint64(x) for x in (0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9,
                                                    0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5))


def _rotl(x, r):
    return (x << np.uint64(r)) | (x >> np.uint64(64 - r))


def bench_key_bytes(ids, prefix=b"bench_k"):
    """uint8 [n, 16]: prefix + 9 decimal digits."""
    ids = np.asarray(ids, dtype=np.int64)
    out = np.empty((len(ids), 16), dtype=np.uint8)
    out[:, :7] = np.frombuffer(prefix, dtype=np.uint8)
    v = ids.copy()
    for d in range(9):
        out[:, 15 - d] = (v % 10 + 48).astype(np.uint8)
        v //= 10
    return out


def bench_key_hashes(ids, prefix=b"bench_k"):
    with np.errstate(over="ignore"):
        b = bench_key_bytes(ids, prefix)
        w = np.ascontiguousarray(b).view("<u8")  # [n, 2]
        h = _P5 + np.uint64(16)
        for j in range(2):
            k = _rotl(w[:, j] * _P2, 31) * _P1
            h = _rotl(h ^ k, 27) * _P1 + _P4
        h = (h ^ (h >> np.uint64(33))) * _P2
        h = (h ^ (h >> np.uint64(29))) * _P3
        xx = h ^ (h >> np.uint64(32))
        f = np.full(len(b), 0xCBF29CE484222325, dtype=np.uint64)
        prime = np.uint64(0x100000001B3)
        for j in range(16):
            f = (f * prime) ^ b[:, j].astype(np.uint64)
    return xx, f


def bench_requests(ids, created_at, mixed=True, prefix=b"bench_k", dtype=None):
    """BASELINE.md request records for the given key ids (hits=1, limit=100, duration=60 000, burst=0)."""
    ids = np.asarray(ids, dtype=np.int64)
    reqs = np.zeros(len(ids), dtype=dtype or O.HREQ_DTYPE)
    reqs["key_xxh64"], reqs["key_fnv1"] = bench_key_hashes(ids, prefix)
    reqs["hits"] = 1; reqs["limit"] = 100; reqs["duration"] = 60000
    reqs["created_at"] = created_at
    reqs["algorithm"] = (ids & 1) if mixed else 0
    reqs["behavior"] = O.REQ_IS_OWNER
    return reqs


def extreme_batch(rng, n, n_keys, now_ms):
    """Requests whose numeric fields sit on the edges of int64 / float64: wrapping sums, 2^52..2^63 leaky values, zero and
    negative limits and durations, created_at far in the past and future.  Same-key repeats included."""
    E = np.array([0, 1, -1, 2, 7, 100, (1 << 31), (1 << 52) - 1, (1 << 52), (1 << 53) + 1, (1 << 62), (1 << 63) - 1, -(1 << 63),
                  -(1 << 62), -(1 << 53), -3, 60000, 1000], dtype=np.int64)
    reqs = np.zeros(n, dtype=O.HREQ_DTYPE)
    ids = rng.integers(0, n_keys, n)
    reqs["key_xxh64"], reqs["key_fnv1"] = key_hashes(ids, name="ext")
    reqs["hits"] = rng.choice(E, n); reqs["limit"] = rng.choice(E, n); reqs["duration"] = rng.choice(E, n)
    reqs["burst"] = rng.choice(E, n)
    reqs["created_at"] = now_ms + rng.choice(np.array([0, 0, 1, -1, 60000, -60000, 1 << 40, -(1 << 40), (1 << 62)], dtype=np.int64), n)
    reqs["algorithm"] = rng.integers(0, 2, n)
    reqs["behavior"] = rng.choice(np.array([0, 0, O.RESET_REMAINING, O.DRAIN_OVER_LIMIT, O.DURATION_IS_GREGORIAN], dtype=np.uint32), n) | np.uint32(O.REQ_IS_OWNER)
    greg = (reqs["behavior"] & O.DURATION_IS_GREGORIAN) != 0
    reqs["duration"][greg] = rng.integers(-1, 7, int(greg.sum()))
    # make a third of the traffic repeat the previous request of the same key exactly (uniform runs of extreme values)
    last = {}
    for i in range(n):
        k = int(ids[i])
        if k in last and rng.random() < 0.35:
            j = last[k]
            for f in ("hits", "limit", "duration", "burst", "created_at", "algorithm", "behavior"):
                reqs[f][i] = reqs[f][j]
        last[k] = i
    return reqs
