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


def test_bench_parses_ncu_rows_of_templated_kernels(tmp_path):
    """bench.py's roofline.traffic leg: kernel names come out of ncu as `void k_rank<0>(BatchArgs)` / `gub::k_rank<1>(...)`; the parser
    keys them as k_rank etc.  Fed with the committed launch list of the final kernels, metric names swapped for the DRAM counters."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    src = open(os.path.join(root, "profiles", "r02_ncu_launches_final.csv")).read()
    both = src.replace("gpu__time_duration.sum", "dram__bytes_read.sum").replace('"ns"', '"byte"')
    extra = [ln.replace("dram__bytes_read.sum", "dram__bytes_write.sum") for ln in both.splitlines() if "dram__bytes_read.sum" in ln and "Metric Name" not in ln]
    p = tmp_path / "traffic.csv"
    p.write_text(both + "\n".join(extra) + "\n")
    out = bench.parse_traffic_csv(str(p))
    assert set(out) == {"k_group", "k_rank", "k_eval", "k_finish"}
    assert all(v["launches"] == 8 and v["dram_read_bytes_per_launch"] > 0 and v["dram_write_bytes_per_launch"] == v["dram_read_bytes_per_launch"] for v in out.values())
