"""ctypes binding of the CPU oracle (oracle/gub_oracle.c).  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgub_oracle.so")

TOKEN_BUCKET, LEAKY_BUCKET = 0, 1
UNDER_LIMIT, OVER_LIMIT = 0, 1
NO_BATCHING, GLOBAL, DURATION_IS_GREGORIAN, RESET_REMAINING, MULTI_REGION, DRAIN_OVER_LIMIT = 1, 2, 4, 8, 16, 32
REQ_IS_OWNER = 0x100

# byte-identical to include/gubernator_b200.h gub_req / gub_resp
HREQ_DTYPE = np.dtype([("key_xxh64", "<u8"), ("key_fnv1", "<u8"), ("hits", "<i8"), ("limit", "<i8"),
                       ("duration", "<i8"), ("burst", "<i8"), ("created_at", "<i8"), ("algorithm", "<u4"),
                       ("behavior", "<u4")])
HRESP_DTYPE = np.dtype([("status", "<u4"), ("err_code", "<u4"), ("limit", "<i8"), ("remaining", "<i8"),
                        ("reset_time", "<i8")])
assert HREQ_DTYPE.itemsize == 64 and HRESP_DTYPE.itemsize == 32


class Req(C.Structure):
    _fields_ = [("name", C.c_char_p), ("unique_key", C.c_char_p), ("hits", C.c_int64), ("limit", C.c_int64),
                ("duration", C.c_int64), ("burst", C.c_int64), ("algorithm", C.c_int32), ("behavior", C.c_int32),
                ("created_at", C.c_int64)]


class Resp(C.Structure):
    _fields_ = [("status", C.c_int32), ("err_code", C.c_int32), ("limit", C.c_int64), ("remaining", C.c_int64),
                ("reset_time", C.c_int64), ("error", C.c_char * 256)]


class Item(C.Structure):
    _fields_ = [("algorithm", C.c_int32), ("value_kind", C.c_int32), ("expire_at", C.c_int64),
                ("invalid_at", C.c_int64), ("status", C.c_int32), ("_pad", C.c_int32), ("limit", C.c_int64),
                ("duration", C.c_int64), ("remaining_i", C.c_int64), ("remaining_f", C.c_double),
                ("stamp", C.c_int64), ("burst", C.c_int64)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("gub_oracle.c", "gub_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u64, i64, sz = C.c_void_p, C.c_uint64, C.c_int64, C.c_size_t
        L.gubo_xxh64.restype = u64; L.gubo_xxh64.argtypes = [C.c_char_p, sz, u64]
        L.gubo_fnv1_64.restype = u64; L.gubo_fnv1_64.argtypes = [C.c_char_p, sz]
        L.gubo_fnv1a_64.restype = u64; L.gubo_fnv1a_64.argtypes = [C.c_char_p, sz]
        L.gubo_md5_hex.argtypes = [C.c_char_p, sz, C.c_char_p]
        L.gubo_gregorian_duration.argtypes = [i64, i64, C.POINTER(i64)]
        L.gubo_gregorian_expiration.argtypes = [i64, i64, C.POINTER(i64)]
        L.gubo_pool_new.restype = vp; L.gubo_pool_new.argtypes = [C.c_int, i64]
        L.gubo_pool_free.argtypes = [vp]
        L.gubo_pool_worker_index_for_hash63.argtypes = [vp, u64]
        L.gubo_pool_worker_index.argtypes = [vp, C.c_char_p, sz]
        L.gubo_pool_set_now.argtypes = [vp, i64]
        L.gubo_pool_now.restype = i64; L.gubo_pool_now.argtypes = [vp]
        L.gubo_get_rate_limits.argtypes = [vp, C.POINTER(Req), sz, C.POINTER(Resp), C.c_int, C.c_int]
        L.gubo_pool_add_item.argtypes = [vp, C.c_char_p, sz, C.POINTER(Item)]
        L.gubo_pool_get_item.argtypes = [vp, C.c_char_p, sz, C.POINTER(Item)]
        L.gubo_pool_update_peer_global.argtypes = [vp, C.c_char_p, sz, C.c_int32, i64, C.c_int32, i64, i64, i64]
        L.gubo_pool_add_item_hashed.argtypes = [vp, u64, u64, C.POINTER(Item)]
        L.gubo_pool_get_item_hashed.argtypes = [vp, u64, u64, C.POINTER(Item)]
        L.gubo_pool_update_peer_global_hashed.argtypes = [vp, u64, u64, C.c_int32, i64, C.c_int32, i64, i64, i64]
        L.gubo_pool_size.restype = i64; L.gubo_pool_size.argtypes = [vp]
        L.gubo_pool_counters.argtypes = [vp, C.POINTER(i64)]
        L.gubo_pool_each.restype = sz; L.gubo_pool_each.argtypes = [vp, C.POINTER(Item), vp, vp, sz]
        L.gubo_ring_new.restype = vp; L.gubo_ring_new.argtypes = [C.c_int, C.c_int]
        L.gubo_ring_add.argtypes = [vp, C.c_char_p]
        L.gubo_ring_get.argtypes = [vp, C.c_char_p, sz]
        L.gubo_ring_get_by_hash.argtypes = [vp, u64]
        L.gubo_ring_points.restype = sz; L.gubo_ring_points.argtypes = [vp, vp, vp, sz]
        L.gubo_ring_free.argtypes = [vp]
        L.gubo_submit_hashed.argtypes = [vp, vp, sz, vp]
        L.gubo_submit_hashed_mt.restype = C.c_double; L.gubo_submit_hashed_mt.argtypes = [vp, vp, sz, vp, C.c_int]
        L.gubo_submit_keys_mt.restype = C.c_double; L.gubo_submit_keys_mt.argtypes = [vp, vp, vp, vp, sz, vp, C.c_int]
        _lib = L
    return _lib


def xxh64(b: bytes, seed=0):
    return lib().gubo_xxh64(b, len(b), seed)


def fnv1_64(b: bytes):
    return lib().gubo_fnv1_64(b, len(b))


def fnv1a_64(b: bytes):
    return lib().gubo_fnv1a_64(b, len(b))


def md5_hex(b: bytes):
    out = C.create_string_buffer(33)
    lib().gubo_md5_hex(b, len(b), out)
    return out.value.decode()


def gregorian_duration(now_ms, d):
    out = C.c_int64(0)
    err = lib().gubo_gregorian_duration(now_ms, d, C.byref(out))
    return out.value, err


def gregorian_expiration(now_ms, d):
    out = C.c_int64(0)
    err = lib().gubo_gregorian_expiration(now_ms, d, C.byref(out))
    return out.value, err


class Pool:
    """WorkerPool + frozen clock (the holster clock.Freeze/Advance used by the reference's tests)."""

    def __init__(self, workers=1, cache_size=0, now_ms=1_700_000_000_000):
        self._p = lib().gubo_pool_new(workers, cache_size)
        self.set_now(now_ms)

    def close(self):
        if self._p:
            lib().gubo_pool_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_now(self, now_ms):
        lib().gubo_pool_set_now(self._p, int(now_ms))

    def now(self):
        return lib().gubo_pool_now(self._p)

    def advance(self, ms):
        self.set_now(self.now() + int(ms))

    def get_rate_limits(self, reqs, is_owner=True, unbounded=False):
        """reqs: list of dicts with RateLimitReq field names.  Returns list of dicts (RateLimitResp) or raises."""
        n = len(reqs)
        arr = (Req * max(n, 1))()
        for i, r in enumerate(reqs):
            arr[i].name = r.get("name", "").encode()
            arr[i].unique_key = r.get("unique_key", "").encode()
            arr[i].hits = r.get("hits", 0); arr[i].limit = r.get("limit", 0)
            arr[i].duration = r.get("duration", 0); arr[i].burst = r.get("burst", 0)
            arr[i].algorithm = r.get("algorithm", 0); arr[i].behavior = r.get("behavior", 0)
            arr[i].created_at = r.get("created_at", 0) or 0
        out = (Resp * max(n, 1))()
        rc = lib().gubo_get_rate_limits(self._p, arr, n, out, 1 if is_owner else 0, 1 if unbounded else 0)
        if rc != 0:
            raise ValueError("Requests.RateLimits list too large; max size is '1000'")
        return [dict(status=o.status, limit=o.limit, remaining=o.remaining, reset_time=o.reset_time,
                     error=o.error.decode(), err_code=o.err_code) for o in out[:n]]

    def submit_hashed(self, reqs: np.ndarray, threads=0):
        assert reqs.dtype == HREQ_DTYPE and reqs.flags.c_contiguous
        out = np.zeros(len(reqs), dtype=HRESP_DTYPE)
        if threads and threads > 0:
            self.last_mt_seconds = lib().gubo_submit_hashed_mt(self._p, reqs.ctypes.data, len(reqs), out.ctypes.data, threads)
        else:
            lib().gubo_submit_hashed(self._p, reqs.ctypes.data, len(reqs), out.ctypes.data)
        return out

    def submit_keys(self, key_bytes: np.ndarray, offsets: np.ndarray, reqs: np.ndarray, threads=1):
        """reqs: HREQ_DTYPE records whose hash fields are ignored; the keys are hashed inside the (timed) call."""
        assert reqs.dtype == HREQ_DTYPE and offsets.dtype == np.uint64 and key_bytes.dtype == np.uint8
        out = np.zeros(len(reqs), dtype=HRESP_DTYPE)
        scratch = reqs.copy()
        self.last_mt_seconds = lib().gubo_submit_keys_mt(self._p, key_bytes.ctypes.data, offsets.ctypes.data, scratch.ctypes.data, len(reqs), out.ctypes.data, max(1, threads))
        return out

    def worker_index_for_hash63(self, h):
        return lib().gubo_pool_worker_index_for_hash63(self._p, h)

    def worker_index(self, key: bytes):
        return lib().gubo_pool_worker_index(self._p, key, len(key))

    def add_item(self, key: bytes, item: Item):
        lib().gubo_pool_add_item(self._p, key, len(key), C.byref(item))

    def get_item(self, key: bytes):
        it = Item()
        ok = lib().gubo_pool_get_item(self._p, key, len(key), C.byref(it))
        return it if ok else None

    def update_peer_global(self, key: bytes, algorithm, duration, status, limit, remaining, reset_time):
        lib().gubo_pool_update_peer_global(self._p, key, len(key), algorithm, duration, status, limit, remaining, reset_time)

    def add_item_hashed(self, kx, kf, item: Item):
        lib().gubo_pool_add_item_hashed(self._p, int(kx), int(kf), C.byref(item))

    def get_item_hashed(self, kx, kf):
        it = Item()
        ok = lib().gubo_pool_get_item_hashed(self._p, int(kx), int(kf), C.byref(it))
        return it if ok else None

    def update_peer_global_hashed(self, kx, kf, algorithm, duration, status, limit, remaining, reset_time):
        lib().gubo_pool_update_peer_global_hashed(self._p, int(kx), int(kf), algorithm, duration, status, limit, remaining, reset_time)

    def size(self):
        return lib().gubo_pool_size(self._p)

    def counters(self):
        out = (C.c_int64 * 4)()
        lib().gubo_pool_counters(self._p, out)
        return dict(over_limit=out[0], cache_hit=out[1], cache_miss=out[2], unexpired_evictions=out[3])

    def each(self):
        """Returns dict (xxh64, fnv1) -> Item for pre-hashed keys."""
        n = self.size()
        items = (Item * max(n, 1))()
        kx = np.zeros(max(n, 1), dtype=np.uint64)
        kf = np.zeros(max(n, 1), dtype=np.uint64)
        m = lib().gubo_pool_each(self._p, items, kx.ctypes.data, kf.ctypes.data, n)
        assert m == n
        return {(int(kx[i]), int(kf[i])): items[i] for i in range(n)}


class Ring:
    def __init__(self, hash_kind=0, replicas=512):
        self._r = lib().gubo_ring_new(hash_kind, replicas)
        self.peers = []

    def add(self, addr: str):
        lib().gubo_ring_add(self._r, addr.encode())
        self.peers.append(addr)

    def get(self, key: str):
        b = key.encode()
        return lib().gubo_ring_get(self._r, b, len(b))

    def get_by_hash(self, h):
        return lib().gubo_ring_get_by_hash(self._r, h)

    def points(self):
        n = lib().gubo_ring_points(self._r, None, None, 0)
        hs = np.zeros(n, dtype=np.uint64)
        ps = np.zeros(n, dtype=np.int32)
        lib().gubo_ring_points(self._r, hs.ctypes.data, ps.ctypes.data, n)
        return hs, ps

    def __del__(self):
        try:
            lib().gubo_ring_free(self._r)
        except Exception:
            pass
