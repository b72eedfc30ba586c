"""torchrun worker shared by tests/test_sharded_gloo.py (CPU, gloo, oracle backend) and tests/test_gpu_sharded.py (NCCL,
CUDA backend): runs gubernator_b200.sharded.ShardedStep for a few steps and checks this rank's responses against a
local simulation of every shard with one oracle per shard."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.environ["GUB_ROOT"]
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle_py as O  # noqa: E402
from gubernator_b200.sharded import ShardedStep, shard_addresses  # noqa: E402
from workloads import T0, adversarial_batch, bench_requests, zipf_ids  # noqa: E402

USE_GPU = os.environ.get("GUB_BACKEND", "cpu") == "gpu"
if USE_GPU:
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    dist.init_process_group("gloo")
rank, W = dist.get_rank(), dist.get_world_size()
oring = O.Ring(0, 512)
for a in shard_addresses(W):
    oring.add(a)


class CpuBackend:
    """numpy/oracle stand-in for GpuBackend (same method set)."""

    def __init__(self, pool):
        self.pool = pool

    def route(self, reqs, n):
        r = reqs[:n].numpy().reshape(-1).view(O.HREQ_DTYPE)
        owner = np.array([oring.get_by_hash(int(h)) for h in r["key_fnv1"]], dtype=np.int64)
        perm = np.argsort(owner, kind="stable")
        routed = torch.from_numpy(r[perm].view(np.uint8).reshape(n, 64).copy())
        return routed, perm, torch.from_numpy(np.bincount(owner, minlength=W).astype(np.int32))

    def empty_like(self, t):
        return torch.empty_like(t)

    def two_lists(self, a, b):
        return [int(x) for x in a.tolist()], [int(x) for x in b.tolist()]

    def req_buffer(self, m):
        return torch.empty((m, 64), dtype=torch.uint8)

    def resp_buffer(self, n):
        return torch.empty((n, 32), dtype=torch.uint8)

    def evaluate(self, inbox, m, clk):
        r = inbox[:m].numpy().reshape(-1).view(O.HREQ_DTYPE)
        self.pool.set_now(clk)
        return torch.from_numpy(self.pool.submit_hashed(np.ascontiguousarray(r)).view(np.uint8).reshape(m, 32).copy())

    def unroute(self, back, perm, n, out):
        out[perm] = back[:n]


SIZES = {0: [3000, 1, 2500, 0, 4000, 60000], 1: [2000, 700, 0, 5, 4000, 65536]}


def batch_for(src, step):
    rng = np.random.default_rng(1000 * step + src)
    n = SIZES[src % 2][step % 6]
    if n == 0:
        return np.zeros(0, dtype=O.HREQ_DTYPE)
    if step % 2 == 0:
        return adversarial_batch(rng, n, 37 if n < 10000 else 3000, T0 + step)
    return bench_requests(zipf_ids(rng, n, 5000, 1.1), T0 + step)


if USE_GPU:
    import gubernator_b200 as g
    from gubernator_b200.sharded import GpuBackend
    dev = torch.device("cuda", local)
    tab = g.Table(1 << 18, max_batch=131072, device=local)
    ring = g.Ring(0, 512)
    for a in shard_addresses(W):
        ring.add(a)
    backend = GpuBackend(tab, ring, W, dev, 131072)
else:
    backend = CpuBackend(O.Pool(workers=2, cache_size=10**7, now_ms=T0))
ROUTE = os.environ.get("GUB_ROUTE", "nccl")
if USE_GPU and ROUTE == "p2pg":
    # GLOBAL behaviour across processes, everything behind the C ABI: gub_p2p_step with gub_p2p_enable_global, gub_global_tick with
    # the NCCL all-gather of the UpdatePeerGlobal items; every rank replays the whole cluster in the oracle-side model
    from global_model import OracleCluster
    from workloads import key_hashes
    p2p = g.native.P2P(tab, ring, rank, 8192)
    handles = [None] * W
    dist.all_gather_object(handles, p2p.export())
    p2p.connect(handles)
    p2p.enable_global(4096)
    uid = [g.native.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    p2p.nccl_init(uid[0])
    dist.barrier()
    model = OracleCluster(W, T0)
    rng = np.random.default_rng(4242)  # the same stream on every rank: everybody generates everybody's batches
    now = T0
    st = torch.cuda.current_stream().cuda_stream
    for step in range(14):
        now += int(rng.choice([0, 1, 5, 400]))
        batches = []
        for r in range(W):
            n = int(rng.choice([0, 1, 700, 3000]))
            ids = rng.integers(0, 300, n)
            b = np.zeros(n, dtype=O.HREQ_DTYPE)
            xx, fv = key_hashes(ids, name="glob")
            b["key_xxh64"], b["key_fnv1"] = xx, fv
            b["limit"] = 20 + (ids % 7) * 10; b["duration"] = 30000 + (ids % 3) * 30000; b["algorithm"] = (ids >> 1) & 1
            b["hits"] = np.where(ids % 11 == 0, 0, 1 + (ids % 3)); b["created_at"] = now
            glob = (ids % 5) < 2
            b["behavior"] = np.where(glob, O.GLOBAL, 0).astype(np.uint32) | np.uint32(O.REQ_IS_OWNER)
            if step == 7:
                b["behavior"] |= np.where(glob & (ids % 13 == 0), O.RESET_REMAINING, 0).astype(np.uint32)
            batches.append(b)
        mine, n = batches[rank], len(batches[rank])
        buf = torch.from_numpy(mine.view(np.uint8).reshape(n, 64).copy()).to(dev) if n else torch.empty((1, 64), dtype=torch.uint8, device=dev)
        out = torch.zeros((max(n, 1), 32), dtype=torch.uint8, device=dev)
        p2p.step(buf.data_ptr(), n, g.clock_fill(now), out.data_ptr(), st)
        torch.cuda.synchronize()
        got = out[:n].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE)
        want = model.step(batches, now)[rank]
        if not np.array_equal(got, want):
            bad = np.nonzero(got != want)[0]
            raise AssertionError(f"rank {rank} step {step}: {len(bad)} responses differ; first {bad[0]}: {got[bad[0]]} vs {want[bad[0]]}")
        if step % 3 == 2:
            p2p.tick(g.clock_fill(now), now, st)
            torch.cuda.synchronize()
            model.tick(now)
    p2p.status()
    dist.barrier()
    print(f"rank {rank} ok")
    dist.destroy_process_group()
    sys.exit(0)
USE_P2P = USE_GPU and ROUTE in ("p2p", "p2p2")
PIPELINED = USE_GPU and ROUTE == "p2p2"  # two streams, all steps in flight before any is checked
if USE_P2P:  # records travel by NVLink stores from the routing kernels (cudaIpc mailboxes) instead of NCCL all-to-all
    from gubernator_b200.sharded import P2PStep
    stepper = P2PStep(tab, ring, W, rank, cap=65536)
    stepper.connect(dist)
else:
    stepper = ShardedStep(backend, dist, W)
sim = [O.Pool(workers=2, cache_size=10**7, now_ms=T0) for _ in range(W)]  # local simulation of every shard
got_all = None
if PIPELINED:
    s_in, s_ev = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    bufs, outs = [], []
    for step in range(8):
        reqs = batch_for(rank, step)
        n = len(reqs)
        bufs.append(torch.from_numpy(reqs.view(np.uint8).reshape(n, 64).copy()).to(dev) if n else torch.empty((1, 64), dtype=torch.uint8, device=dev))
        outs.append(torch.zeros((max(n, 1), 32), dtype=torch.uint8, device=dev))
    torch.cuda.synchronize()
    for step in range(8):
        stepper.step(bufs[step], len(batch_for(rank, step)), g.clock_fill(T0 + step), outs[step], stream=s_ev.cuda_stream, ingest_stream=s_in.cuda_stream)
    torch.cuda.synchronize()
    got_all = [outs[step][:len(batch_for(rank, step))].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE) for step in range(8)]
for step in range(8):
    now = T0 + step
    reqs = batch_for(rank, step)
    n = len(reqs)
    host = torch.from_numpy(reqs.view(np.uint8).reshape(n, 64).copy())
    if got_all is not None:
        got = got_all[step]
    elif USE_GPU:
        buf = host.to(dev) if n else torch.empty((1, 64), dtype=torch.uint8, device=dev)
        out = torch.zeros((max(n, 1), 32), dtype=torch.uint8, device=dev)
        stepper.step(buf, n, g.clock_fill(now), out)
        torch.cuda.synchronize()
        got = out[:n].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE)
    else:
        out = torch.zeros((n, 32), dtype=torch.uint8)
        stepper.step(host, n, now, out)
        got = out.numpy().reshape(-1).view(O.HRESP_DTYPE)
    # reference: every owner applies the records of source 0 (owner-sorted, index order), then source 1, ...
    all_batches = [batch_for(s, step) for s in range(W)]
    owners = [np.array([oring.get_by_hash(int(h)) for h in b["key_fnv1"]], dtype=np.int64) for b in all_batches]
    results = [np.zeros(len(b), dtype=O.HRESP_DTYPE) for b in all_batches]
    for gi in range(W):
        sim[gi].set_now(now)
        for s in range(W):
            idx = np.nonzero(owners[s] == gi)[0]
            if len(idx):
                results[s][idx] = sim[gi].submit_hashed(np.ascontiguousarray(all_batches[s][idx]))
    if not np.array_equal(got, results[rank]):
        bad = np.nonzero(got != results[rank])[0]
        raise AssertionError(f"rank {rank} step {step}: {len(bad)} responses differ; first {bad[0]}: {got[bad[0]]} vs {results[rank][bad[0]]}")
dist.barrier()
print(f"rank {rank} ok")
dist.destroy_process_group()
