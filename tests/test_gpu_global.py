"""GLOBAL behaviour on the GPU: a W-shard cluster in ONE process on one GPU (the shape of the reference's cluster.StartWith
fixture), checked against the oracle-side model in tests/global_model.py and against the reference's own GLOBAL scenarios
(functional_test.go:959-1341, :1690-2097).  Two drivers of the same device code: "c" = everything behind the C ABI
(gub_p2p_step_local_all with gub_p2p_enable_global, gub_global_tick_local_all: NVLink-mailbox routing, queues fed by the step, the
tick's broadcast read from the peers in place); "python" = the torch.distributed-shaped orchestration of
gubernator_b200.sharded.ShardedStep over the same kernels (collectives replaced by the in-process LocalExchange)."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_py as O
from global_model import OracleCluster
from workloads import T0, key_hashes

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["pipeline", "fused"])
def owner_kernel_path(request, monkeypatch):
    """Owners evaluate their mailboxes with the four-kernel pipeline in ring mode (default) or with the persistent kernel k_batch
    (GUB_PATH=fused); read when a table is created."""
    monkeypatch.setenv("GUB_PATH", request.param)
    return request.param


class GpuCluster:
    def __init__(self, world, now_ms, capacity=1 << 14):
        import torch
        import gubernator_b200 as g
        from gubernator_b200.sharded import GpuBackend, LocalExchange, ShardedStep, shard_addresses
        self.g, self.torch, self.W = g, torch, world
        self.dev = torch.device("cuda", 0)
        self.ex = LocalExchange(world)
        self.ring = g.Ring(0, 512)
        for a in shard_addresses(world):
            self.ring.add(a)
        self.tabs = [g.Table(capacity, max_batch=16384, device=0) for _ in range(world)]
        self.steppers = [ShardedStep(GpuBackend(self.tabs[r], self.ring, world, self.dev, 16384, rank=r, global_capacity=4096), self.ex.handle(r), world,
                                     global_sync=True) for r in range(world)]
        self.pool = ThreadPoolExecutor(world)

    def _all(self, fn):
        return [f.result() for f in [self.pool.submit(fn, r) for r in range(self.W)]]

    def step(self, batches, now_ms):
        clk = self.g.clock_fill(now_ms)
        torch = self.torch

        def run(r):
            torch.cuda.set_device(0)
            b = batches[r]
            n = len(b)
            buf = torch.from_numpy(b.view(np.uint8).reshape(n, 64).copy()).to(self.dev) if n else torch.empty((1, 64), dtype=torch.uint8, device=self.dev)
            out = torch.zeros((max(n, 1), 32), dtype=torch.uint8, device=self.dev)
            self.steppers[r].step(buf, n, clk, out)
            torch.cuda.synchronize()
            return out[:n].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE)
        return self._all(run)

    def tick(self, now_ms):
        clk = self.g.clock_fill(now_ms)

        def run(r):
            self.torch.cuda.set_device(0)
            res = self.steppers[r].tick(clk, now_ms)
            self.torch.cuda.synchronize()
            return res
        return self._all(run)


class CCluster:
    """The same surface over tests/local_ring.LocalRing (C ABI drivers)."""

    def __init__(self, world, now_ms, capacity=1 << 14):
        from local_ring import LocalRing
        self.lr = LocalRing(world, capacity=capacity, cap=4096, global_capacity=4096)
        self.W, self.tabs = world, self.lr.tabs

    def step(self, batches, now_ms):
        return self.lr.step(batches, now_ms)

    def tick(self, now_ms):
        return self.lr.tick(now_ms)


@pytest.fixture(params=["c", "python"])
def make_cluster(request):
    return CCluster if request.param == "c" else GpuCluster


def _req(name_id, hits, limit, duration, algorithm=0, behavior=O.GLOBAL, created_at=T0):
    r = np.zeros(1, dtype=O.HREQ_DTYPE)
    xx, fv = key_hashes([name_id], name="glob")
    r["key_xxh64"], r["key_fnv1"] = xx, fv
    r["hits"] = hits; r["limit"] = limit; r["duration"] = duration; r["algorithm"] = algorithm
    r["behavior"] = behavior | O.REQ_IS_OWNER; r["created_at"] = created_at
    return r


def _send(cl, model, shard, req, now):
    """One request through shard `shard` of both clusters; returns (gpu response, model response)."""
    empty = np.zeros(0, dtype=O.HREQ_DTYPE)
    batches = [req if r == shard else empty for r in range(cl.W)]
    a = cl.step(batches, now)[shard][0]
    b = model.step(batches, now)[shard][0]
    assert a == b, (a, b)
    return a


def test_reference_global_scenarios(make_cluster):
    """TestGlobalRateLimits (functional_test.go:959-1032) and TestGlobalRateLimitsPeerOverLimit (:1093-1142), with the
    reference's wall-clock waits for the async send/broadcast replaced by explicit ticks."""
    W = 6
    cl, model = make_cluster(W, T0), OracleCluster(W, T0)
    rq = lambda hits: _req(1, hits, 5, 180000)
    owner = int(model.owners(rq(1))[0])
    peers = [r for r in range(W) if r != owner]
    now = T0
    first = _send(cl, model, peers[0], rq(1), now)                       # created on the peer, queued for async forward
    assert (first["status"], first["remaining"], first["limit"]) == (0, 4, 5)
    r = _send(cl, model, peers[0], rq(2), now)                           # processed as if we own it
    assert (r["status"], r["remaining"]) == (0, 2) and r["reset_time"] == first["reset_time"]
    assert [x[0] for x in cl.tick(now)][peers[0]] == 1                   # one aggregated hit record sent to the owner
    model.tick(now)
    for p in (peers[1], peers[2]):                                       # they got the broadcast from the owner
        r = _send(cl, model, p, rq(0), now)
        assert (r["status"], r["remaining"]) == (0, 2) and r["reset_time"] == first["reset_time"]
    r = _send(cl, model, peers[3], rq(2), now)                           # non-owner computes remaining before forwarding
    assert (r["status"], r["remaining"]) == (0, 0)
    cl.tick(now); model.tick(now)
    r = _send(cl, model, peers[4], rq(1), now)
    assert (r["status"], r["remaining"]) == (1, 0)

    # TestGlobalRateLimitsPeerOverLimit
    rq2 = lambda hits: _req(2, hits, 2, 300000)
    owner2 = int(model.owners(rq2(1))[0])
    p0 = [r for r in range(W) if r != owner2][0]
    assert tuple(_send(cl, model, p0, rq2(1), now)[["status", "remaining"]]) == (0, 1)
    assert tuple(_send(cl, model, p0, rq2(1), now)[["status", "remaining"]]) == (0, 0)
    cl.tick(now); model.tick(now)
    assert tuple(_send(cl, model, p0, rq2(1), now)[["status", "remaining"]]) == (1, 0)
    cl.tick(now); model.tick(now)
    assert tuple(_send(cl, model, p0, rq2(0), now)[["status", "remaining"]]) == (1, 0)


def test_reference_global_reset_remaining(make_cluster):
    """TestGlobalResetRemaining (functional_test.go:1258-1341): leaky bucket, every peer takes 50, reset propagates."""
    W = 4
    cl, model = make_cluster(W, T0), OracleCluster(W, T0)
    rq = lambda hits, beh=O.GLOBAL: _req(3, hits, 100, 60000 * 1000, algorithm=1, behavior=beh)
    owner = int(model.owners(rq(1))[0])
    peers = [r for r in range(W) if r != owner]
    now = T0
    for p in peers:
        assert tuple(_send(cl, model, p, rq(50), now)[["status", "remaining"]]) == (0, 50)
    cl.tick(now); model.tick(now)
    assert tuple(_send(cl, model, peers[0], rq(1), now)[["status", "remaining"]]) == (1, 0)
    r = _send(cl, model, peers[0], rq(0, O.GLOBAL | O.RESET_REMAINING), now)
    cl.tick(now); model.tick(now)
    _send(cl, model, peers[1], rq(0), now)


@pytest.mark.parametrize("world", [2, 3])
def test_global_random_traffic_matches_model(world, make_cluster):
    rng = np.random.default_rng(500 + world)
    cl, model = make_cluster(world, T0), OracleCluster(world, T0)
    now = T0
    n_keys = 300
    for step in range(14):
        now += int(rng.choice([0, 1, 5, 400]))
        batches = []
        for r in range(world):
            n = int(rng.choice([0, 1, 700, 3000]))
            ids = rng.integers(0, n_keys, n)
            b = np.zeros(n, dtype=O.HREQ_DTYPE)
            xx, fv = key_hashes(ids, name="glob")
            b["key_xxh64"], b["key_fnv1"] = xx, fv
            # per-key fixed parameters (the queues keep the first / latest request per key: identical here, so the
            # arbitrary winner among same-step duplicates cannot matter); ~40 % of keys are GLOBAL
            b["limit"] = 20 + (ids % 7) * 10; b["duration"] = 30000 + (ids % 3) * 30000; b["algorithm"] = (ids >> 1) & 1
            b["hits"] = np.where(ids % 11 == 0, 0, 1 + (ids % 3)); b["created_at"] = now
            glob = (ids % 5) < 2
            b["behavior"] = np.where(glob, O.GLOBAL, 0).astype(np.uint32) | np.uint32(O.REQ_IS_OWNER)
            if step == 7:
                b["behavior"] |= np.where(glob & (ids % 13 == 0), O.RESET_REMAINING, 0).astype(np.uint32)
            batches.append(b)
        got, want = cl.step(batches, now), model.step(batches, now)
        for r in range(world):
            if not np.array_equal(got[r], want[r]):
                bad = np.nonzero(got[r] != want[r])[0]
                raise AssertionError(f"step {step} shard {r}: {len(bad)} differ; first {bad[0]}: {got[r][bad[0]]} vs {want[r][bad[0]]} req {batches[r][bad[0]]}")
        if step % 3 == 2:
            cl.tick(now); model.tick(now)
    # replicas and owners hold identical state on every shard
    for r in range(world):
        items = model.pools[r].each()
        scan = cl.tabs[r].scan()
        dev = {(int(s["key_xxh64"]), int(s["key_fnv1"]) >> 8): s for s in scan}
        assert len(dev) == len(items)
        for (kx, kf), it in items.items():
            s = dev[(kx, kf >> 8)]
            assert int(s["limit"]) == it.limit and int(s["expire_at"]) == it.expire_at and int(s["stamp"]) == it.stamp, (r, s, it.limit, it.expire_at, it.stamp)
            if it.value_kind == 2:
                assert np.float64(s["remaining_f"]).view(np.uint64) == np.float64(it.remaining_f).view(np.uint64)
            else:
                assert int(s["remaining"]) == it.remaining_i and int(s["status"]) == it.status


@pytest.mark.parametrize("hits", [1, 10])
@pytest.mark.parametrize("where", ["owner", "non_owner", "distributed"])
def test_reference_global_behavior(where, hits):
    """TestGlobalBehavior (functional_test.go:1690-2097), with the reference's metric counters read from the tick's statistics:
    hits on the owner -> no hit update from anybody, exactly one broadcast item from the owner, installed once on every other peer;
    hits on one non-owner -> exactly one hit-update record, from that peer; hits spread over the non-owners -> one hit-update
    record from each peer that took hits; and in every case all peers then report the same Remaining (:1816-1821)."""
    W, limit = 6, 1000
    cl, model = CCluster(W, T0), OracleCluster(W, T0)
    rq = lambda h: _req({"owner": 11, "non_owner": 12, "distributed": 13}[where] * 100 + hits, h, limit, 180000)
    owner = int(model.owners(rq(1))[0])
    peers = [r for r in range(W) if r != owner]
    now = T0
    took = set()
    for i in range(hits):
        shard = owner if where == "owner" else (peers[0] if where == "non_owner" else peers[i % len(peers)])
        r = _send(cl, model, shard, rq(1), now)
        assert r["status"] == 0
        if where != "distributed":
            assert r["remaining"] == limit - 1 - i  # sendHit(..., 999 - i)
        if shard != owner:
            took.add(shard)
    cl.tick(now); model.tick(now)
    st = cl.lr.last_tick
    assert [st[r]["hits_sent"] for r in range(W)] == [1 if r in took else 0 for r in range(W)]       # gubernator_global_send_duration_count
    assert [st[r]["updates_made"] for r in range(W)] == [1 if r == owner else 0 for r in range(W)]   # gubernator_broadcast_duration_count
    assert [st[r]["installed"] for r in range(W)] == [0 if r == owner else 1 for r in range(W)]      # UpdatePeerGlobals once per other peer
    for r in range(W):
        got = _send(cl, model, r, rq(0), now)
        assert (got["status"], got["remaining"]) == (0, limit - hits)
    cl.tick(now); model.tick(now)  # nothing queued: Hits = 0 requests are never queued (global.go:74-84)
    assert all(x["hits_sent"] == 0 and x["updates_made"] == 0 and x["installed"] == 0 for x in cl.lr.last_tick)
