"""Key-space sharding across the GPUs of one box: the intra-box replacement for the reference's peer forwarding
(gubernator.go:257-283 -> peer_client.go:284 runBatch -> GetPeerRateLimits, gubernator.go:462).

One process per GPU.  Per step every rank
  1. routes its ingest batch by owner — owner(key) is the reference's ReplicatedConsistentHash.Get over the addresses
     "gpu:0".."gpu:N-1" (replicated_hash.go:104-119) — as a stable partition, so per-key order inside the batch survives,
  2. exchanges the per-owner counts, then the 64-byte request records (all-to-all, variable splits; NCCL over NVLink on GPUs),
  3. evaluates the records it owns: the batch it sees is the concatenation over source ranks (rank order) of their
     owner-sorted records, a deterministic order that tests reproduce with one oracle per shard,
  4. sends the 32-byte responses back along the same splits and restores request order.

The device work (route / evaluate / unroute) is supplied by a backend; GpuBackend drives the CUDA library.  The
exchange logic itself is backend-agnostic so that tests/test_sharded_gloo.py can run it on CPU with gloo.
"""
import numpy as np


def shard_addresses(n):
    return [f"gpu:{r}" for r in range(n)]


class ShardedStep:
    def __init__(self, backend, dist, world):
        self.be, self.dist, self.world = backend, dist, world

    def step(self, reqs, n, clk, out):
        """reqs/out: backend buffers holding n request / response records.  Returns the number of records evaluated here."""
        be, dist, W = self.be, self.dist, self.world
        routed, perm, counts = be.route(reqs, n)                 # counts: integer tensor [W] on the backend's device
        recv_counts = be.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts)
        send_l, recv_l = [int(x) for x in counts.tolist()], [int(x) for x in recv_counts.tolist()]
        m = sum(recv_l)
        inbox = be.req_buffer(m)
        dist.all_to_all_single(inbox[:m], routed[:n], output_split_sizes=recv_l, input_split_sizes=send_l)
        resp = be.evaluate(inbox, m, clk)
        back = be.resp_buffer(n)
        dist.all_to_all_single(back[:n], resp[:m], output_split_sizes=send_l, input_split_sizes=recv_l)
        be.unroute(back, perm, n, out)
        return m


class GpuBackend:
    """torch CUDA tensors as buffers (torch is only the allocator and the NCCL plumbing); kernels from the C ABI."""

    def __init__(self, table, ring, world, device, cap):
        import torch
        self.torch, self.tab, self.ring, self.W, self.dev = torch, table, ring, world, device
        self.cap = cap
        self.routed = torch.empty((cap, 64), dtype=torch.uint8, device=device)
        self.perm = torch.empty(cap, dtype=torch.int32, device=device)
        self.counts16 = torch.zeros(16, dtype=torch.int32, device=device)
        self.in_cap = 0
        self.inbox = self.inbox_resp = None
        self.back = torch.empty((cap, 32), dtype=torch.uint8, device=device)

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def route(self, reqs, n):
        self.tab.route_device(self.ring, reqs.data_ptr(), n, self.routed.data_ptr(), self.perm.data_ptr(), self.counts16.data_ptr(), self._stream())
        return self.routed, self.perm, self.counts16[:self.W].clone()

    def empty_like(self, t):
        return self.torch.empty_like(t)

    def req_buffer(self, m):
        if m > self.in_cap:
            self.in_cap = int(m * 1.25) + 4096
            self.inbox = self.torch.empty((self.in_cap, 64), dtype=self.torch.uint8, device=self.dev)
            self.inbox_resp = self.torch.empty((self.in_cap, 32), dtype=self.torch.uint8, device=self.dev)
        return self.inbox

    def resp_buffer(self, n):
        return self.back

    def evaluate(self, inbox, m, clk):
        self.tab.submit_device(inbox.data_ptr(), m, clk, self.inbox_resp.data_ptr(), self._stream())
        return self.inbox_resp

    def unroute(self, back, perm, n, out):
        self.tab.unroute_device(back.data_ptr(), perm.data_ptr(), n, out.data_ptr(), self._stream())
