"""gubernator_b200 — B200-native rate-limit evaluation path behind gubernator's WorkerPool boundary.

The product is the CUDA shared library (csrc/, C ABI in include/gubernator_b200.h).  This package is the thin Python
binding used by the tests and bench.py plus a host-side mirror of the reference's service interface for the path
(`V1Instance.GetRateLimits`).  There is no CPU fallback: importing `native` fails loudly if the library is missing.
"""
from . import native  # noqa: F401
from .native import (Clock, Ring, Table, clock_fill, fnv1_64, fnv1a_64, hash_keys, xxh64, REQ_DTYPE, RESP_DTYPE,  # noqa: F401
                     ITEM_DTYPE, CLOCK_DTYPE)
from .service import V1Instance, RateLimitReq, RateLimitResp  # noqa: F401
