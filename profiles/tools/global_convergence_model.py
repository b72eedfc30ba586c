import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/oracle"]
import numpy as np
import oracle_py as O
from global_model import OracleCluster
from workloads import bench_requests, T0, zipf_ids

W, n_keys, hot, per_step, steps, dt = 2, 20000, 200, 2048, 500, 4
rng = [np.random.default_rng(100 + r) for r in range(W)]
cl = OracleCluster(W, T0)
def batch(r, now):
    ids = zipf_ids(rng[r], per_step, n_keys, 1.1)
    reqs = bench_requests(ids, now)
    reqs["behavior"] = np.where(ids < hot, O.GLOBAL | O.REQ_IS_OWNER, O.REQ_IS_OWNER).astype(np.uint32)
    return reqs
now = T0
for b in range(30):
    now = T0 + 1 + dt * b
    cl.step([batch(r, now) for r in range(W)], now)
cl.tick(now)
for b in range(30, 30 + steps):
    now = T0 + 1 + dt * b
    cl.step([batch(r, now) for r in range(W)], now)
cl.tick(now); cl.tick(now)
q = bench_requests(np.arange(hot, dtype=np.int64), now)
q["hits"] = 0
q["behavior"] = np.uint32(O.GLOBAL | O.REQ_IS_OWNER)
outs = cl.step([q.copy() for _ in range(W)], now)
tok = q["algorithm"] == 0
d = np.abs(outs[0]["remaining"].astype(np.int64) - outs[1]["remaining"].astype(np.int64))
print("token keys differing:", int(((outs[0]["remaining"] != outs[1]["remaining"]) | (outs[0]["status"] != outs[1]["status"]))[tok].sum()), "of", int(tok.sum()))
print("leaky: max |remaining diff|", int(d[~tok].max()), "keys differing", int((d[~tok] > 0).sum()), "of", int((~tok).sum()))
print("leaky diffs histogram", np.bincount(d[~tok]))
