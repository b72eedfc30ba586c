"""W shards of the ring inside ONE process (the shape of the reference daemon, and of its cluster.StartWith test fixture), driven by
one host thread through the C ABI: gub_p2p_step_local_all / gub_global_tick_local_all.  On a one-GPU box all shards share device
0 (the phases are enqueued shard by shard, so no kernel waits for work that has not been enqueued); with more GPUs each shard
takes its own device and the mailboxes are peer memory.  TEST INFRASTRUCTURE."""
import numpy as np

import oracle_py as O


class LocalRing:
    def __init__(self, world, capacity=1 << 16, cap=8192, global_capacity=0, spread_devices=False):
        import torch
        import gubernator_b200 as g
        from gubernator_b200.sharded import shard_addresses
        self.g, self.torch, self.W, self.cap = g, torch, world, cap
        ndev = torch.cuda.device_count() if spread_devices else 1
        self.devs = [r % ndev for r in range(world)]
        self.ring = g.Ring(0, 512)
        for a in shard_addresses(world):
            self.ring.add(a)
        self.tabs = [g.Table(capacity, device=self.devs[r]) for r in range(world)]
        self.p2ps = [g.native.P2P(self.tabs[r], self.ring, r, cap) for r in range(world)]
        for p in self.p2ps:
            p.connect_local(self.p2ps)
        if global_capacity:
            for p in self.p2ps:
                p.enable_global(global_capacity)
        self.streams = [torch.cuda.Stream(device=torch.device("cuda", d)) for d in self.devs]

    def upload(self, batches):
        torch = self.torch
        bufs, outs = [], []
        for r, b in enumerate(batches):
            dev = torch.device("cuda", self.devs[r])
            n = len(b)
            bufs.append(torch.from_numpy(np.ascontiguousarray(b).view(np.uint8).reshape(n, 64).copy()).to(dev) if n else torch.empty((1, 64), dtype=torch.uint8, device=dev))
            outs.append(torch.zeros((max(n, 1), 32), dtype=torch.uint8, device=dev))
        for d in set(self.devs):
            torch.cuda.synchronize(d)
        return bufs, outs

    def enqueue(self, bufs, ns, now_ms, outs):
        self.g.native.p2p_step_local_all(self.p2ps, [b.data_ptr() for b in bufs], ns, self.g.clock_fill(now_ms), [o.data_ptr() for o in outs],
                                         [s.cuda_stream for s in self.streams])

    def sync(self):
        for s in self.streams:
            s.synchronize()
        for p in self.p2ps:
            p.status()

    def step(self, batches, now_ms):
        """batches[r] = the ingest batch of shard r (HREQ records); returns the responses of every shard."""
        bufs, outs = self.upload(batches)
        ns = [len(b) for b in batches]
        self.enqueue(bufs, ns, now_ms, outs)
        self.sync()
        return [outs[r][:ns[r]].cpu().numpy().reshape(-1).view(O.HRESP_DTYPE) for r in range(self.W)]

    def tick(self, now_ms):
        res = self.g.native.global_tick_local_all(self.p2ps, self.g.clock_fill(now_ms), now_ms, [s.cuda_stream for s in self.streams])
        self.sync()
        self.last_tick = res
        return [(r["hits_sent"], r["installed"]) for r in res]


def per_shard_oracle_results(oring, sim, batches, now_ms):
    """Every owner applies the records of source 0 (index order), then source 1, ...: the documented evaluation order."""
    W = len(sim)
    owners = [np.array([oring.get_by_hash(int(h)) for h in b["key_fnv1"]], dtype=np.int64) for b in batches]
    want = [np.zeros(len(b), dtype=O.HRESP_DTYPE) for b in batches]
    for gi in range(W):
        sim[gi].set_now(now_ms)
        for s in range(W):
            idx = np.nonzero(owners[s] == gi)[0]
            if len(idx):
                want[s][idx] = sim[gi].submit_hashed(np.ascontiguousarray(batches[s][idx]))
    return want
