"""Oracle-side model of a W-shard cluster with GLOBAL behaviour, restating the reference's orchestration
(gubernator.go:183-295 owner/global dispatch, gubernator.go:395-421 getGlobalRateLimit, global.go:91-283 hit
aggregation and broadcast, gubernator.go:425-459 UpdatePeerGlobals, gubernator.go:510-512 DRAIN_OVER_LIMIT on forwarded
GLOBAL hits) on top of one oracle pool per shard.  TEST INFRASTRUCTURE (uses oracle/)."""
import numpy as np

import oracle_py as O
from gubernator_b200.sharded import shard_addresses


class OracleCluster:
    def __init__(self, world, now_ms):
        self.W = world
        self.pools = [O.Pool(workers=2, cache_size=10**8, now_ms=now_ms) for _ in range(world)]
        self.ring = O.Ring(0, 512)
        for a in shard_addresses(world):
            self.ring.add(a)
        self.hits = [dict() for _ in range(world)]     # per shard: key -> [first request record, summed hits, reset flag]
        self.updates = [dict() for _ in range(world)]  # per owner: key -> latest request record

    def owners(self, reqs):
        return np.array([self.ring.get_by_hash(int(h)) for h in reqs["key_fnv1"]], dtype=np.int64)

    def step(self, batches, now_ms):
        """batches[r] = the ingest batch of shard r.  Returns the responses of every shard."""
        W = self.W
        outs = [np.zeros(len(b), dtype=O.HRESP_DTYPE) for b in batches]
        routed = []
        for r, b in enumerate(batches):
            own = self.owners(b)
            is_global = (b["behavior"] & O.GLOBAL) != 0
            dest = np.where(is_global & (own != r), r, own)
            ev = b.copy()
            local_global = is_global & (own != r)
            # gubernator.go:408-411: clone, NO_BATCHING on, GLOBAL off, IsOwner = false
            ev["behavior"][local_global] = (ev["behavior"][local_global] | O.NO_BATCHING) & ~np.uint32(O.GLOBAL | O.REQ_IS_OWNER)
            # global.go:74-78 QueueHit + 99-111 aggregation: first request kept, hits summed, RESET_REMAINING OR-ed
            for i in np.nonzero(local_global & (b["hits"] != 0))[0]:
                k = (int(b["key_xxh64"][i]), int(b["key_fnv1"][i]))
                if k in self.hits[r]:
                    e = self.hits[r][k]
                    e[1] = (e[1] + int(b["hits"][i])) & 0xFFFFFFFFFFFFFFFF
                    e[2] |= int(b["behavior"][i]) & O.RESET_REMAINING
                else:
                    self.hits[r][k] = [b[i].copy(), int(b["hits"][i]) & 0xFFFFFFFFFFFFFFFF, int(b["behavior"][i]) & O.RESET_REMAINING]
            perm = np.argsort(dest, kind="stable")
            routed.append((ev[perm], dest[perm], perm))
        for g in range(W):
            self.pools[g].set_now(now_ms)
            for s in range(W):
                ev, dest, perm = routed[s]
                sel = np.nonzero(dest == g)[0]
                if len(sel) == 0:
                    continue
                recs = np.ascontiguousarray(ev[sel])
                outs[s][perm[sel]] = self.pools[g].submit_hashed(recs)
                self._queue_updates(g, recs)
        return outs

    def _queue_updates(self, g, recs):
        # gubernator.go:604-606 + global.go:80-84: GLOBAL requests with Hits != 0 evaluated here as owner; latest wins
        for i in np.nonzero(((recs["behavior"] & O.GLOBAL) != 0) & (recs["hits"] != 0))[0]:
            self.updates[g][(int(recs["key_xxh64"][i]), int(recs["key_fnv1"][i]))] = recs[i].copy()

    def tick(self, now_ms):
        W = self.W
        # sendHits (global.go:144-190) -> GetPeerRateLimits on the owner (gubernator.go:462-539)
        outgoing = [[] for _ in range(W)]  # per owner: list of (src, record)
        for s in range(W):
            for (kx, kf), (first, total, reset) in self.hits[s].items():
                r = first.copy()
                r["hits"] = np.uint64(total).astype(np.int64)
                r["behavior"] = (int(first["behavior"]) | reset | O.DRAIN_OVER_LIMIT | O.REQ_IS_OWNER)
                outgoing[self.ring.get_by_hash(kf)].append((s, r))
            self.hits[s] = {}
        for g in range(W):
            self.pools[g].set_now(now_ms)
            for s in range(W):
                recs = [r for (src, r) in outgoing[g] if src == s]
                if recs:
                    arr = np.array(recs, dtype=O.HREQ_DTYPE)
                    self.pools[g].submit_hashed(arr)
                    self._queue_updates(g, arr)
        # broadcastPeers (global.go:234-283): status with Hits = 0, IsOwner = false; UpdatePeerGlobals on every other shard
        for g in range(W):
            ups = self.updates[g]
            self.updates[g] = {}
            for (kx, kf), rec in ups.items():
                q = np.array([rec], dtype=O.HREQ_DTYPE)
                q["hits"] = 0
                q["behavior"] = q["behavior"] & ~np.uint32(O.REQ_IS_OWNER)
                st = self.pools[g].submit_hashed(q)[0]
                if int(st["err_code"]) != 0 or int(rec["algorithm"]) > 1:
                    continue
                for p in range(W):
                    if p != g:
                        self.pools[p].set_now(now_ms)
                        self.pools[p].update_peer_global_hashed(kx, kf, int(rec["algorithm"]), int(rec["duration"]), int(st["status"]), int(st["limit"]),
                                                         int(st["remaining"]), int(st["reset_time"]))
