"""Known-answer tables transcribed from the reference's own tests (mailgun/gubernator v2.4.0).

Each scenario cites the reference test it restates (file:line in /root/reference).  A scenario is played one
request per GetRateLimits call against a frozen clock (holster clock.Freeze / clock.Advance in the reference);
`sleep` is clock.Advance in milliseconds applied AFTER the step.  Played against the oracle in
tests/test_oracle_golden.py (CPU) and against the CUDA path in tests/test_gpu_golden.py (GPU).

Fields of a step: req overrides (merged over scenario["req"]), `expect` = dict of status / remaining / limit /
error, `reset` = name of a ResetTime predicate evaluated with (resp, now_ms_at_call).
"""

TOKEN, LEAKY = 0, 1
UNDER, OVER = 0, 1
BATCHING, NO_BATCHING, GLOBAL, GREGORIAN, RESET_REMAINING, MULTI_REGION, DRAIN_OVER_LIMIT = 0, 1, 2, 4, 8, 16, 32
SECOND, MINUTE = 1000, 60000
GREG_MINUTES, GREG_HOURS, GREG_DAYS, GREG_WEEKS, GREG_MONTHS, GREG_YEARS = 0, 1, 2, 3, 4, 5

# 2023-11-14T22:13:20.123Z: mid-minute so Gregorian-minute scenarios do not straddle a boundary by accident
T0 = 1_700_000_000_123


def _leaky_reset_3s(resp, now_ms):
    # functional_test.go:598,705,846: clock.Now().Unix()+(rl.Limit-rl.Remaining)*3 == rl.ResetTime/1000
    return now_ms // 1000 + (resp["limit"] - resp["remaining"]) * 3 == resp["reset_time"] // 1000


RESET_PREDICATES = {
    "nonzero": lambda resp, now_ms: resp["reset_time"] != 0,
    "zero": lambda resp, now_ms: resp["reset_time"] == 0,
    "leaky_3s": _leaky_reset_3s,
    "gt_now_s": lambda resp, now_ms: resp["reset_time"] > now_ms // 1000,
}


def S(hits=None, remaining=None, status=UNDER, sleep=0, reset=None, error="", **req):
    if hits is not None:
        req["hits"] = hits
    exp = {"status": status, "error": error}
    if remaining is not None:
        exp["remaining"] = remaining
    return {"req": req, "expect": exp, "sleep": sleep, "reset": reset}


SCENARIOS = [
    dict(name="TestOverTheLimit", cite="functional_test.go:65-110",
         req=dict(name="test_over_limit", unique_key="account:1234", algorithm=TOKEN, duration=SECOND * 9, limit=2, hits=1),
         limit=2, steps=[S(remaining=1, reset="nonzero"), S(remaining=0, reset="nonzero"), S(remaining=0, status=OVER, reset="nonzero")]),
    dict(name="TestTokenBucket", cite="functional_test.go:160-219",
         req=dict(name="test_token_bucket", unique_key="account:1234", algorithm=TOKEN, duration=5, limit=2, hits=1),
         limit=2, steps=[S(remaining=1, reset="nonzero"), S(remaining=0, sleep=100, reset="nonzero"), S(remaining=1, reset="nonzero")]),
    dict(name="TestTokenBucketGregorian", cite="functional_test.go:221-294",
         req=dict(name="test_token_bucket_greg", unique_key="account:12345", behavior=GREGORIAN, algorithm=TOKEN,
                  duration=GREG_MINUTES, limit=60),
         limit=60, steps=[S(1, 59, reset="nonzero"), S(1, 58, reset="nonzero"), S(58, 0, reset="nonzero"),
                          S(1, 0, status=OVER, sleep=SECOND * 61, reset="nonzero"), S(0, 60, reset="nonzero")]),
    dict(name="TestTokenBucketNegativeHits", cite="functional_test.go:296-366",
         req=dict(name="test_token_bucket_negative", unique_key="account:12345", algorithm=TOKEN, duration=5, limit=2),
         limit=2, steps=[S(-1, 3, reset="nonzero"), S(-1, 4, reset="nonzero"), S(4, 0, reset="nonzero"), S(-1, 1, reset="nonzero")]),
    dict(name="TestDrainOverLimit/TOKEN_BUCKET", cite="functional_test.go:368-432",
         req=dict(name="test_drain_over_limit", unique_key="account:1234:0", algorithm=TOKEN, behavior=DRAIN_OVER_LIMIT,
                  duration=SECOND * 30, limit=10),
         limit=10, steps=[S(0, 10, reset="nonzero"), S(1, 9, reset="nonzero"), S(100, 0, status=OVER, reset="nonzero"),
                          S(0, 0, reset="nonzero")]),
    dict(name="TestDrainOverLimit/LEAKY_BUCKET", cite="functional_test.go:368-432",
         req=dict(name="test_drain_over_limit", unique_key="account:1234:1", algorithm=LEAKY, behavior=DRAIN_OVER_LIMIT,
                  duration=SECOND * 30, limit=10),
         limit=10, steps=[S(0, 10, reset="nonzero"), S(1, 9, reset="nonzero"), S(100, 0, status=OVER, reset="nonzero"),
                          S(0, 0, reset="nonzero")]),
    dict(name="TestTokenBucketRequestMoreThanAvailable", cite="functional_test.go:434-475",
         req=dict(name="test_token_more_than_available", unique_key="account:123456", algorithm=TOKEN, duration=1000, limit=2000),
         limit=2000, steps=[S(1000, 1000), S(1500, 1000, status=OVER), S(500, 500), S(400, 100), S(100, 0), S(1, 0, status=OVER)]),
    dict(name="TestLeakyBucket", cite="functional_test.go:477-602",
         req=dict(name="test_leaky_bucket", unique_key="account:1234", algorithm=LEAKY, duration=SECOND * 30, limit=10),
         limit=10, steps=[
             S(1, 9, sleep=SECOND, reset="leaky_3s"), S(1, 8, sleep=SECOND, reset="leaky_3s"),
             S(1, 7, sleep=1500, reset="leaky_3s"), S(0, 8, sleep=SECOND * 3, reset="leaky_3s"),
             S(0, 9, reset="leaky_3s"), S(9, 0, reset="leaky_3s"), S(1, 0, status=OVER, sleep=SECOND * 3, reset="leaky_3s"),
             S(0, 1, sleep=SECOND * 60, reset="leaky_3s"), S(0, 10, sleep=SECOND * 60, reset="leaky_3s"),
             S(10, 0, sleep=SECOND * 29, reset="leaky_3s"), S(9, 0, sleep=SECOND * 3, reset="leaky_3s"),
             S(1, 0, sleep=SECOND, reset="leaky_3s")]),
    dict(name="TestLeakyBucketWithBurst", cite="functional_test.go:604-709",
         req=dict(name="test_leaky_bucket_with_burst", unique_key="account:1234", algorithm=LEAKY, duration=SECOND * 30,
                  limit=10, burst=20),
         limit=10, steps=[
             S(1, 19, sleep=SECOND, reset="leaky_3s"), S(1, 18, sleep=SECOND, reset="leaky_3s"),
             S(1, 17, sleep=1500, reset="leaky_3s"), S(0, 18, sleep=SECOND * 3, reset="leaky_3s"),
             S(0, 19, reset="leaky_3s"), S(19, 0, reset="leaky_3s"), S(1, 0, status=OVER, sleep=SECOND * 3, reset="leaky_3s"),
             S(0, 1, sleep=SECOND * 60, reset="leaky_3s"), S(0, 20, sleep=SECOND, reset="leaky_3s")]),
    dict(name="TestLeakyBucketGregorian", cite="functional_test.go:711-779",
         req=dict(name="TestLeakyBucketGregorian", unique_key="Xk3mWq9ZpL", behavior=GREGORIAN, algorithm=LEAKY,
                  duration=GREG_MINUTES, limit=60),
         limit=60, steps=[S(1, 59, sleep=500, reset="gt_now_s"), S(1, 58, sleep=1200, reset="gt_now_s"), S(1, 58, reset="gt_now_s")]),
    dict(name="TestLeakyBucketNegativeHits", cite="functional_test.go:781-850",
         req=dict(name="test_leaky_bucket_negative", unique_key="account:12345", algorithm=LEAKY, duration=SECOND * 30, limit=10),
         limit=10, steps=[S(1, 9, reset="leaky_3s"), S(-1, 10, reset="leaky_3s"), S(10, 0, reset="leaky_3s"), S(-1, 1, reset="leaky_3s")]),
    dict(name="TestLeakyBucketRequestMoreThanAvailable", cite="functional_test.go:852-894",
         req=dict(name="test_leaky_more_than_available", unique_key="account:123456", algorithm=LEAKY, duration=1000, limit=2000),
         limit=2000, steps=[S(1000, 1000), S(1500, 1000, status=OVER), S(500, 500), S(400, 100), S(100, 0), S(1, 0, status=OVER)]),
    dict(name="TestChangeLimit", cite="functional_test.go:1343-1436",
         req=dict(name="test_change_limit", unique_key="account:1234", duration=9000, hits=1),
         limit=None, steps=[
             S(remaining=99, algorithm=TOKEN, limit=100, reset="nonzero"), S(remaining=98, algorithm=TOKEN, limit=100, reset="nonzero"),
             S(remaining=7, algorithm=TOKEN, limit=10, reset="nonzero"), S(remaining=6, algorithm=TOKEN, limit=10, reset="nonzero"),
             S(remaining=195, algorithm=TOKEN, limit=200, reset="nonzero"), S(remaining=99, algorithm=LEAKY, limit=100, reset="nonzero"),
             S(remaining=9, algorithm=LEAKY, limit=10, reset="nonzero"), S(remaining=8, algorithm=LEAKY, limit=10, reset="nonzero")]),
    dict(name="TestResetRemaining", cite="functional_test.go:1438-1508",
         req=dict(name="test_reset_remaining", unique_key="account:1234", algorithm=TOKEN, duration=9000, hits=1, limit=100),
         limit=100, steps=[S(remaining=99, behavior=BATCHING), S(remaining=98, behavior=BATCHING),
                           S(remaining=100, behavior=RESET_REMAINING), S(remaining=99, behavior=BATCHING)]),
    dict(name="TestLeakyBucketDivBug", cite="functional_test.go:1535-1576",
         req=dict(name="TestLeakyBucketDivBug", unique_key="qT7bN2xVaa", algorithm=LEAKY, duration=1000, limit=2000),
         limit=2000, steps=[S(1, 1999), S(100, 1899)]),
]

# TestMissingFields functional_test.go:896-957: each request is its own key / call; only error + status are pinned
MISSING_FIELDS = [
    (dict(name="test_missing_fields", unique_key="account:1234", hits=1, limit=10, duration=0), "", UNDER),
    (dict(name="test_missing_fields", unique_key="account:12345", hits=1, duration=10000, limit=0), "", OVER),
    (dict(unique_key="account:1234", hits=1, duration=10000, limit=5), "field 'namespace' cannot be empty", UNDER),
    (dict(name="test_missing_fields", hits=1, duration=10000, limit=5), "field 'unique_key' cannot be empty", UNDER),
]

# interval_test.go:47-136 — (now as (Y,M,D,h,m,s,ns) UTC, interval, expected expiration ms)
GREGORIAN_EXPIRATION = [
    ((2019, 11, 11, 0, 0, 0, 0), GREG_MINUTES, (2019, 11, 11, 0, 0, 59, 999)),
    ((2019, 11, 11, 0, 0, 30, 100), GREG_MINUTES, 1573430459999),
    ((2019, 11, 11, 0, 0, 0, 0), GREG_HOURS, (2019, 11, 11, 0, 59, 59, 999)),
    ((2019, 11, 11, 0, 20, 1, 2134), GREG_HOURS, 1573433999999),
    ((2019, 11, 11, 0, 0, 0, 0), GREG_DAYS, (2019, 11, 11, 23, 59, 59, 999)),
    ((2019, 11, 11, 12, 10, 9, 2345), GREG_DAYS, 1573516799999),
    ((2019, 11, 1, 0, 0, 0, 0), GREG_MONTHS, (2019, 11, 30, 23, 59, 59, 999)),
    ((2019, 11, 11, 22, 2, 23, 0), GREG_MONTHS, 1575158399999),
    ((2019, 1, 1, 0, 0, 0, 0), GREG_MONTHS, (2019, 1, 31, 23, 59, 59, 999)),
    ((2019, 1, 1, 0, 0, 0, 0), GREG_YEARS, (2019, 12, 31, 23, 59, 59, 999)),
    # clock.Date(2019, March, 1, 20, 30, 1231, 0): 1231 s normalises to 20:50:31
    ((2019, 3, 1, 20, 50, 31, 0), GREG_YEARS, 1577836799999),
]
GREGORIAN_INVALID_MSG = "behavior DURATION_IS_GREGORIAN is set; but `Duration` is not a valid gregorian interval"  # interval_test.go:135
GREGORIAN_WEEKS_MSG = "`Duration = GregorianWeeks` not yet supported; consider making a PR!`"  # interval.go:93

# replicated_hash_test.go:27,56-101
RING_HOSTS = ["a.svc.local", "b.svc.local", "c.svc.local"]
RING_DISTRIBUTION = {
    "fnv1": {"a.svc.local": 2948, "b.svc.local": 3592, "c.svc.local": 3460},
    "fnv1a": {"a.svc.local": 3110, "b.svc.local": 3856, "c.svc.local": 3034},
}

# workers_internal_test.go:46-55 (32 workers)
WORKER_INDEX = [(0, 0), (0x3FFF_FFFF_FFFF_FFFF, 15), (0x4000_0000_0000_0000, 16), (0x7FFF_FFFF_FFFF_FFFF, 31)]

# XXH64 published test vectors (xxHash specification; seed 0) + the value quoted in SURVEY.md §8c
XXH64_VECTORS = [(b"", 0xEF46DB3751D8E999), (b"a", 0xD24EC4F1A98C6E5B), (b"abc", 0x44BC2CF5AD770999),
                 (b"Foobar", 0x9DE0B9C33B6693DF),
                 (b"Nobody inspects the spammish repetition", 0xFBCEA83C8A378BF1)]
# FNV test vectors (Noll's reference vectors): FNV-1 64 and FNV-1a 64 of "" / "a" / "foobar"
FNV1_VECTORS = [(b"", 0xCBF29CE484222325), (b"a", 0xAF63BD4C8601B7BE), (b"foobar", 0x340D8765A4DDA9C2)]
FNV1A_VECTORS = [(b"", 0xCBF29CE484222325), (b"a", 0xAF63DC4C8601EC8C), (b"foobar", 0x85944171F73967E8)]
MD5_VECTORS = [(b"", "d41d8cd98f00b204e9800998ecf8427e"), (b"abc", "900150983cd24fb0d6963f7d28e17f72"),
               (b"The quick brown fox jumps over the lazy dog", "9e107d9d372bb6826bd81d3542a419d6")]
