"""CPU-only checks of the C-ABI shared library: it loads, exports every symbol the headers declare, its host-side
helpers (hashing, ring, batch clock) agree with the oracle, and it refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_py as O
from golden import reference_kat as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G():
    from gubernator_b200 import build
    build.build()
    import gubernator_b200 as g
    return g


def test_exports_match_headers(G):
    L = G.native.lib()
    declared = set()
    for h in ("gubernator_b200.h", "gubernator_b200_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(gub_[a-z0-9_]+)\s*\(", src))
    assert len(declared) > 35
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(G.native.EXPORTS) <= declared
    assert L.gub_abi_version() == 2


def test_record_layouts_match_header(G):
    # the numpy dtypes used everywhere must be the C structs
    assert G.REQ_DTYPE.itemsize == 64 and G.RESP_DTYPE.itemsize == 32
    assert G.REQ_DTYPE == O.HREQ_DTYPE and G.RESP_DTYPE == O.HRESP_DTYPE


def test_host_hashes_match_oracle(G):
    rng = np.random.default_rng(3)
    for data, want in K.XXH64_VECTORS:
        assert G.xxh64(data) == want
    for n in list(range(0, 80)) + [255, 256, 1000]:
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert G.xxh64(b) == O.xxh64(b) and G.xxh64(b, 99) == O.xxh64(b, 99)
        assert G.fnv1_64(b) == O.fnv1_64(b) and G.fnv1a_64(b) == O.fnv1a_64(b)
    keys = [f"bench_k{i:09d}".encode() for i in range(1000)] + [b"", b"x"]
    xx, fv = G.hash_keys(keys)
    assert all(int(xx[i]) == O.xxh64(k) and int(fv[i]) == O.fnv1_64(k) for i, k in enumerate(keys))


@pytest.mark.parametrize("kind,name", [(0, "fnv1"), (1, "fnv1a")])
def test_ring_matches_reference_distribution_and_oracle(G, kind, name):
    ring, oring = G.Ring(kind, 512), O.Ring(kind, 512)
    assert ring.get("anything") == -1
    for h in K.RING_HOSTS:
        ring.add(h); oring.add(h)
    dist = {h: 0 for h in K.RING_HOSTS}
    for i in range(10000):
        ip = f"192.168.{(i >> 8) & 255}.{i & 255}"
        o = ring.get(ip)
        assert o == oring.get(ip)
        dist[K.RING_HOSTS[o]] += 1
    assert dist == K.RING_DISTRIBUTION[name]  # replicated_hash_test.go:69-83
    hs, ps = ring.points()
    ohs, ops = oring.points()
    assert np.array_equal(hs, ohs) and np.array_equal(ps, ops)


def test_clock_fill_matches_oracle_gregorian(G):
    rng = np.random.default_rng(11)
    stamps = [0, 1, 59_999, 60_000, 951_782_400_000, 1_573_430_430_000, 1_582_934_400_000, 1_709_164_800_000,
              1_735_689_599_999, 1_735_689_600_000, 4_102_444_800_000]
    stamps += [int(x) for x in rng.integers(0, 4_102_444_800_000, 500)]
    for now in stamps:
        clk = G.clock_fill(now)
        assert int(clk["now_ms"][0]) == now
        for d in (0, 1, 2, 4, 5):
            assert int(clk["greg_expire"][0][d]) == O.gregorian_expiration(now, d)[0], (now, d)
            assert int(clk["greg_duration"][0][d]) == O.gregorian_duration(now, d)[0], (now, d)
    import calendar
    for now_t, d, want in K.GREGORIAN_EXPIRATION:  # interval_test.go:47-136
        now = calendar.timegm(now_t[:6]) * 1000
        if isinstance(want, tuple):
            want = calendar.timegm(want[:6]) * 1000 + want[6]
        assert int(G.clock_fill(now)["greg_expire"][0][d]) == want


def test_no_cpu_fallback_without_gpu(G):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(G.native.GubError):
        G.Table(1024)
    with pytest.raises(G.native.GubError):
        G.V1Instance(1024)


def test_error_strings_match_oracle(G):
    L = G.native.lib()
    L.gub_format_error.argtypes = [C.c_int, C.c_char_p, C.c_int32, C.c_char_p, C.c_size_t]
    pool = O.Pool(now_ms=K.T0)
    cases = [dict(name="n", unique_key="k", algorithm=5, limit=1, duration=1, hits=1),
             dict(name="n", unique_key="k", behavior=K.GREGORIAN, duration=K.GREG_WEEKS, limit=1, hits=1),
             dict(name="n", unique_key="k", algorithm=1, behavior=K.GREGORIAN, duration=77, limit=1, hits=1),
             dict(name="n", unique_key="", limit=1), dict(name="", unique_key="k", limit=1)]
    for r in cases:
        want = pool.get_rate_limits([r])[0]
        buf = C.create_string_buffer(256)
        L.gub_format_error(want["err_code"], b"n_k", r.get("algorithm", 0), buf, 256)
        assert buf.value.decode() == want["error"]


def test_compact_batch_round_trip(G):
    """native.compact_batch (what a shim fills for gub_submit_compact*): expanding the 32-byte records with the parameter
    table, as k_expand does on the device, gives back the 64-byte records field for field."""
    from workloads import T0, adversarial_batch, bench_requests, zipf_ids
    rng = np.random.default_rng(5)
    for reqs in (adversarial_batch(rng, 5000, 300, T0), bench_requests(zipf_ids(rng, 4000, 1000, 1.1), T0 + 3), adversarial_batch(rng, 1, 1, T0)):
        reqs = reqs.astype(G.REQ_DTYPE) if reqs.dtype != G.REQ_DTYPE else reqs
        c, params, base = G.native.compact_batch(reqs)
        assert c.dtype.itemsize == 32 and params.dtype.itemsize == 32 and len(c) == len(reqs)
        back = np.zeros(len(reqs), dtype=G.REQ_DTYPE)
        back["key_xxh64"], back["key_fnv1"], back["hits"] = c["key_xxh64"], c["key_fnv1"], c["hits"]
        for f in ("limit", "duration", "burst", "algorithm", "behavior"):
            back[f] = params[f][c["params"]]
        back["created_at"] = base + c["created_delta"].astype(np.int64)
        assert back.tobytes() == reqs.tobytes()
