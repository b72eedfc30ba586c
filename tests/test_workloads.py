import numpy as np

import oracle_py as O
from workloads import bench_key_bytes, bench_key_hashes, bench_requests, zipf_ids


def test_vectorised_bench_key_hashes_match_scalar():
    ids = np.array([0, 1, 9, 10, 42, 999_999_999, 123_456_789, 100_000_000 - 1] + list(np.random.default_rng(1).integers(0, 10**8, 200)))
    b = bench_key_bytes(ids)
    xx, fv = bench_key_hashes(ids)
    for j, i in enumerate(ids.tolist()):
        s = f"bench_k{i:09d}".encode()
        assert b[j].tobytes() == s
        assert int(xx[j]) == O.xxh64(s) and int(fv[j]) == O.fnv1_64(s)
    r = bench_requests(ids, 1_700_000_000_000)
    assert np.array_equal(r["algorithm"], ids & 1) and np.all(r["limit"] == 100)


def test_zipf_ids_shape():
    rng = np.random.default_rng(0)
    ids = zipf_ids(rng, 65536, 100_000_000, 1.1)
    assert ids.min() >= 0 and ids.max() < 100_000_000
    _, counts = np.unique(ids, return_counts=True)
    top = np.sort(counts)[::-1]
    # the top key draws about 11 % of a batch at K = 1e8, s = 1.1
    assert 0.07 < top[0] / 65536 < 0.15
    assert top[:100].sum() / 65536 > 0.35
