"""Python binding of the host service layer (include/gubernator_b200_host.h, csrc/host_v1.cpp): `V1Instance` mirrors the
reference's V1Instance.GetRateLimits (gubernator.go:183) for a one-node cluster, evaluated on the GPU."""
import ctypes as C
from dataclasses import dataclass

from . import native

MAX_BATCH_SIZE = 1000  # gubernator.go:40


class _Req(C.Structure):
    _fields_ = [("name", C.c_char_p), ("unique_key", C.c_char_p), ("hits", C.c_int64), ("limit", C.c_int64),
                ("duration", C.c_int64), ("burst", C.c_int64), ("algorithm", C.c_int32), ("behavior", C.c_int32),
                ("created_at", C.c_int64)]


class _Resp(C.Structure):
    _fields_ = [("status", C.c_int32), ("err_code", C.c_int32), ("limit", C.c_int64), ("remaining", C.c_int64),
                ("reset_time", C.c_int64), ("error", C.c_char * 256)]


class _Item(C.Structure):  # gub_item
    _fields_ = [("key_xxh64", C.c_uint64), ("key_fnv1", C.c_uint64), ("algorithm", C.c_int32), ("status", C.c_int32), ("limit", C.c_int64),
                ("duration", C.c_int64), ("remaining", C.c_int64), ("remaining_f", C.c_double), ("stamp", C.c_int64), ("burst", C.c_int64),
                ("expire_at", C.c_int64)]


_GET_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(_Req), C.c_char_p, C.POINTER(_Item))
_CHANGE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(_Req), C.c_char_p, C.POINTER(_Item))
_REMOVE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p)


class _Store(C.Structure):  # gub_store
    _fields_ = [("user", C.c_void_p), ("get", _GET_FN), ("on_change", _CHANGE_FN), ("remove", _REMOVE_FN)]


_ITEM_FIELDS = ("algorithm", "status", "limit", "duration", "remaining", "remaining_f", "stamp", "burst", "expire_at")


def _req_dict(r):
    return dict(name=(r.name or b"").decode(), unique_key=(r.unique_key or b"").decode(), hits=r.hits, limit=r.limit, duration=r.duration,
                burst=r.burst, algorithm=r.algorithm, behavior=r.behavior, created_at=r.created_at)


@dataclass
class RateLimitReq:  # gubernator.proto:137-183
    name: str = ""
    unique_key: str = ""
    hits: int = 0
    limit: int = 0
    duration: int = 0
    algorithm: int = 0
    behavior: int = 0
    burst: int = 0
    created_at: int = 0


@dataclass
class RateLimitResp:  # gubernator.proto:190-203
    status: int = 0
    limit: int = 0
    remaining: int = 0
    reset_time: int = 0
    error: str = ""


_bound = False


def _bind():
    global _bound
    L = native.lib()
    if not _bound:
        vp = C.c_void_p
        L.gub_instance_create.argtypes = [vp, C.POINTER(vp)]
        L.gub_instance_destroy.argtypes = [vp]; L.gub_instance_destroy.restype = None
        L.gub_instance_set_clock.argtypes = [vp, C.c_int64]; L.gub_instance_set_clock.restype = None
        L.gub_instance_now.argtypes = [vp]; L.gub_instance_now.restype = C.c_int64
        L.gub_instance_get_rate_limits.argtypes = [vp, C.POINTER(_Req), C.c_size_t, C.POINTER(_Resp)]
        L.gub_instance_get_rate_limits_unbounded.argtypes = [vp, C.POINTER(_Req), C.c_size_t, C.POINTER(_Resp)]
        L.gub_instance_update_peer_global.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int64]
        L.gub_instance_set_store.argtypes = [vp, C.POINTER(_Store)]; L.gub_instance_set_store.restype = None
        L.gub_aggregator_create.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(vp)]
        L.gub_aggregator_destroy.argtypes = [vp]; L.gub_aggregator_destroy.restype = None
        L.gub_aggregator_get_rate_limits.argtypes = [vp, C.POINTER(_Req), C.c_size_t, C.POINTER(_Resp)]
        L.gub_aggregator_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]; L.gub_aggregator_stats.restype = None
        _bound = True
    return L


class V1Instance:
    """V1Instance for a one-node cluster over one device table, with the frozen-clock hooks the reference's tests use."""

    def __init__(self, capacity_slots=1 << 16, max_batch=65536, device=0, now_ms=None, table=None):
        L = _bind()
        self.table = table or native.Table(capacity_slots, max_batch, device)
        h = C.c_void_p()
        if L.gub_instance_create(self.table._h, C.byref(h)) != 0:
            raise native.GubError("gub_instance_create failed")
        self._h = h
        if now_ms is not None:
            self.set_now(now_ms)

    # holster clock.Freeze / Advance
    def set_now(self, now_ms):
        _bind().gub_instance_set_clock(self._h, int(now_ms))

    def now(self):
        return _bind().gub_instance_now(self._h)

    def advance(self, ms):
        self.set_now(self.now() + int(ms))

    def set_store(self, store):
        """Config.Store (config.go:95).  `store` has get(req: dict, key: str) -> dict | None (item fields: algorithm, status,
        limit, duration, remaining, remaining_f, stamp, burst, expire_at), on_change(req, key, item: dict), remove(key)."""
        if store is None:
            _bind().gub_instance_set_store(self._h, None)
            self._store_refs = None
            return

        def _get(_u, req, key, out):
            it = store.get(_req_dict(req.contents), key.decode())
            if it is None:
                return 0
            for f in _ITEM_FIELDS:
                setattr(out.contents, f, it.get(f, 0))
            return 1

        def _change(_u, req, key, item):
            store.on_change(_req_dict(req.contents), key.decode(), {f: getattr(item.contents, f) for f in _ITEM_FIELDS})

        def _remove(_u, key):
            store.remove(key.decode())
        st = _Store(None, _GET_FN(_get), _CHANGE_FN(_change), _REMOVE_FN(_remove))
        self._store_refs = st  # keep the callbacks alive
        _bind().gub_instance_set_store(self._h, C.byref(st))

    def aggregator(self, max_batch=65536, window_us=500):
        """RPC aggregator over this instance: concurrent get_rate_limits calls share device batches."""
        return Aggregator(self, max_batch, window_us)

    def get_rate_limits(self, reqs, unbounded=False, _via=None):
        """reqs: list of RateLimitReq or dicts.  Returns list of dicts like the oracle binding (status, limit, remaining,
        reset_time, error).  Raises ValueError for more than 1000 requests (gubernator.go:189-193)."""
        L = _bind()
        n = len(reqs)
        arr = (_Req * max(n, 1))()
        for i, r in enumerate(reqs):
            if not isinstance(r, dict):
                r = r.__dict__
            arr[i].name = r.get("name", "").encode()
            arr[i].unique_key = r.get("unique_key", "").encode()
            arr[i].hits = r.get("hits", 0); arr[i].limit = r.get("limit", 0); arr[i].duration = r.get("duration", 0)
            arr[i].burst = r.get("burst", 0); arr[i].algorithm = r.get("algorithm", 0); arr[i].behavior = r.get("behavior", 0)
            arr[i].created_at = r.get("created_at", 0) or 0
        out = (_Resp * max(n, 1))()
        if _via is not None:
            rc = L.gub_aggregator_get_rate_limits(_via, arr, n, out)  # releases the GIL: other threads' calls join the batch
        else:
            fn = L.gub_instance_get_rate_limits_unbounded if unbounded else L.gub_instance_get_rate_limits
            rc = fn(self._h, arr, n, out)
        if rc == -2:
            raise ValueError("Requests.RateLimits list too large; max size is '1000'")
        if rc != 0:
            raise native.GubError("GetRateLimits: " + L.gub_last_error().decode())
        return [dict(status=o.status, limit=o.limit, remaining=o.remaining, reset_time=o.reset_time, error=o.error.decode(),
                     err_code=o.err_code) for o in out[:n]]

    def GetRateLimits(self, reqs):
        return [RateLimitResp(o["status"], o["limit"], o["remaining"], o["reset_time"], o["error"]) for o in self.get_rate_limits(reqs)]

    def update_peer_global(self, key: str, algorithm, duration, status, limit, remaining, reset_time):
        rc = _bind().gub_instance_update_peer_global(self._h, key.encode(), algorithm, duration, status, limit, remaining, reset_time)
        if rc != 0:
            raise native.GubError("UpdatePeerGlobals: " + native.lib().gub_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            _bind().gub_instance_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Aggregator:
    """Coalesces concurrent GetRateLimits calls into device batches (gub_aggregator_*): the stand-in for the batching the Go
    shim does in front of the C ABI, shaped after PeerClient.runBatch (peer_client.go:284-337)."""

    def __init__(self, instance, max_batch=65536, window_us=500):
        L = _bind()
        self.instance = instance
        h = C.c_void_p()
        if L.gub_aggregator_create(instance._h, max_batch, window_us, C.byref(h)) != 0:
            raise native.GubError("gub_aggregator_create failed")
        self._h = h

    def get_rate_limits(self, reqs):
        return self.instance.get_rate_limits(reqs, _via=self._h)

    def stats(self):
        b, r = C.c_uint64(0), C.c_uint64(0)
        _bind().gub_aggregator_stats(self._h, C.byref(b), C.byref(r))
        return dict(batches=b.value, requests=r.value)

    def close(self):
        if getattr(self, "_h", None):
            _bind().gub_aggregator_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
