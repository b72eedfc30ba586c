"""BASELINE config 1 on the CPU: 1 k keys, 10 k-request batches, TOKEN_BUCKET only, uniform keys — the stand-in for the
reference's benchmark_test.go (which cannot run here: no Go toolchain).  Times the oracle's worker-pool port
(oracle/gub_oracle.c, pre-hashed batch path) with 1 worker thread and with all host threads.  Writes one JSON object.
Run from the repo root:  python profiles/tools/cpu_config1.py > profiles/r01_cpu_config1.json"""
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import oracle_py as O
from workloads import T0, bench_requests

KEYS, BATCH, SEED = 1000, 10_000, 0xB200 + 1
rng = np.random.Generator(np.random.PCG64(SEED))
batches = [bench_requests(rng.integers(0, KEYS, BATCH), T0 + 1 + b, mixed=False) for b in range(32)]
out = {"config": "BASELINE config 1: 1k keys, 10k-request batches, TOKEN_BUCKET, uniform", "host_threads": os.cpu_count(), "legs": [],
       "note": "measured on whatever host runs this script (the committed file: the build container, not the GPU box); pre-hashed requests, so the reference's per-request string hashing, channel hops and metrics are NOT included: an upper bound for the Go path"}
for workers in (1, os.cpu_count() or 1):
    pool = O.Pool(workers=workers, cache_size=1 << 20, now_ms=T0)
    pool.submit_hashed(bench_requests(np.arange(KEYS), T0, mixed=False), threads=workers)  # every key resident
    for b in range(8):
        pool.set_now(T0 + 1 + b); pool.submit_hashed(batches[b], threads=workers)
    reps = []
    for rep in range(5):
        t_used, n = 0.0, 0
        for b in range(200):
            pool.set_now(T0 + 1 + b)
            t0 = time.perf_counter()
            pool.submit_hashed(batches[b % len(batches)], threads=workers)
            t_used += (pool.last_mt_seconds if workers > 1 else time.perf_counter() - t0)
            n += BATCH
        reps.append(n / t_used)
    out["legs"].append({"workers": workers, "decisions_per_s_median": float(np.median(reps)), "reps": reps})
print(json.dumps(out, indent=1))
