/*
 * gub_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of mailgun/gubernator v2.4.0's per-request rate-limit evaluation path:
 *   V1Instance.GetRateLimits  (gubernator.go:183-295)
 *   WorkerPool.GetRateLimit   (workers.go:125-184, 261-324)
 *   LRUCache                  (lrucache.go:88-171), CacheItem.IsExpired (cache.go:43-57)
 *   tokenBucket / leakyBucket (algorithms.go:37-493)
 *   GregorianDuration / GregorianExpiration (interval.go:84-148)
 *   ReplicatedConsistentHash  (replicated_hash.go:78-119)
 *   UpdatePeerGlobals item construction (gubernator.go:425-459)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product library (gubernator_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * PARITY PINNING: the Go reference cannot be built in this image (no Go toolchain, ~100 un-vendored modules),
 * so there is no oracle/_ref.  This restatement is pinned against the reference's own known-answer tests,
 * transcribed (with file:line) in tests/golden/: functional_test.go token/leaky tables, interval_test.go
 * Gregorian timestamps, replicated_hash_test.go ring distribution, workers_internal_test.go worker-index math,
 * lrucache_test.go LRU semantics, store_test.go Loader values.  Third-party arithmetic not under
 * /root/reference: XXH64 (OneOfOne/xxhash v1.2.8, seed 0) is checked against python-xxhash and the published
 * XXH64 test vectors; FNV-1/1a 64 (segmentio/fasthash v1.0.2) and crypto/md5 are pinned end-to-end by the ring
 * distribution golden vector.  Go float->int conversions follow amd64 (cvttsd2si); the out-of-range cases
 * (leaky bucket with Limit=0) are exercised by no reference test: "parity unpinned" for those inputs only.
 */
#ifndef GUB_ORACLE_H
#define GUB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* gubernator.proto:56-135 */
enum { GUBO_TOKEN_BUCKET = 0, GUBO_LEAKY_BUCKET = 1 };
enum { GUBO_UNDER_LIMIT = 0, GUBO_OVER_LIMIT = 1 };
enum {
  GUBO_BATCHING = 0,
  GUBO_NO_BATCHING = 1,
  GUBO_GLOBAL = 2,
  GUBO_DURATION_IS_GREGORIAN = 4,
  GUBO_RESET_REMAINING = 8,
  GUBO_MULTI_REGION = 16,
  GUBO_DRAIN_OVER_LIMIT = 32
};

/* error codes carried next to the exact reference error string */
enum {
  GUBO_OK = 0,
  GUBO_ERR_UNIQUE_KEY_EMPTY = 1, /* gubernator.go:208-211 */
  GUBO_ERR_NAMESPACE_EMPTY = 2,  /* gubernator.go:213-216 */
  GUBO_ERR_INVALID_ALGORITHM = 3, /* workers.go:318 */
  GUBO_ERR_GREGORIAN_WEEKS = 4,  /* interval.go:93,134 */
  GUBO_ERR_GREGORIAN_INVALID = 5 /* interval.go:107,147 */
};

/* RateLimitReq (gubernator.proto:137-183).  created_at == 0 means "not set" (gubernator.go:218). */
typedef struct {
  const char* name;
  const char* unique_key;
  int64_t hits, limit, duration, burst;
  int32_t algorithm, behavior;
  int64_t created_at;
} gubo_req;

/* RateLimitResp (gubernator.proto:190-203) */
typedef struct {
  int32_t status;
  int32_t err_code;
  int64_t limit, remaining, reset_time;
  char error[256];
} gubo_resp;

/* CacheItem + TokenBucketItem/LeakyBucketItem flattened (cache.go:29-41, store.go:29-43) */
typedef struct {
  int32_t algorithm;  /* CacheItem.Algorithm */
  int32_t value_kind; /* dynamic type of CacheItem.Value: 0 nil, 1 *TokenBucketItem, 2 *LeakyBucketItem */
  int64_t expire_at, invalid_at;
  /* token: status, limit, duration, remaining_i, stamp=CreatedAt
     leaky: limit, duration, remaining_f, stamp=UpdatedAt, burst */
  int32_t status;
  int32_t _pad;
  int64_t limit, duration;
  int64_t remaining_i;
  double remaining_f;
  int64_t stamp;
  int64_t burst;
} gubo_item;

typedef struct gubo_pool gubo_pool;

/* hashes (third-party in the reference; see header comment) */
uint64_t gubo_xxh64(const void* data, size_t len, uint64_t seed);
uint64_t gubo_fnv1_64(const void* data, size_t len);
uint64_t gubo_fnv1a_64(const void* data, size_t len);
void gubo_md5_hex(const void* data, size_t len, char out33[33]);

/* interval.go:84-148, time zone UTC.  now_ms = clock.Now() in epoch ms.  Returns an error code. */
int gubo_gregorian_duration(int64_t now_ms, int64_t d, int64_t* out);
int gubo_gregorian_expiration(int64_t now_ms, int64_t d, int64_t* out);

/* WorkerPool (workers.go:125-184): `workers` single-threaded shards, each with an LRU of cache_size/workers. */
gubo_pool* gubo_pool_new(int workers, int64_t cache_size);
void gubo_pool_free(gubo_pool*);
/* workers.go:180-184 with an explicit 63-bit hash (workers_internal_test.go:46-55 mocks the hasher) */
int gubo_pool_worker_index_for_hash63(const gubo_pool*, uint64_t hash63);
int gubo_pool_worker_index(const gubo_pool*, const char* key, size_t len);

/* The frozen clock (holster clock.Freeze): every clock.Now()/MillisecondNow() on the path reads this. */
void gubo_pool_set_now(gubo_pool*, int64_t now_ms);
int64_t gubo_pool_now(const gubo_pool*);

/* V1Instance.GetRateLimits restricted to locally-owned keys (gubernator.go:183-295): validation, CreatedAt
 * defaulting, sequential index-order evaluation, in-band error strings.  Returns -1 if n > 1000 unless
 * `unbounded` (the device batch API is unbounded; the 1000 cap is gubernator.go:189). */
int gubo_get_rate_limits(gubo_pool*, const gubo_req* reqs, size_t n, gubo_resp* out, int is_owner, int unbounded);

/* WorkerPool.GetRateLimit for one request with an explicit key (workers.go:261-324).  created_at must be set. */
void gubo_pool_get_rate_limit(gubo_pool*, const char* key, size_t klen, gubo_req* r, int is_owner, gubo_resp* out);

/* WorkerPool.AddCacheItem / GetCacheItem (workers.go:537-626) */
void gubo_pool_add_item(gubo_pool*, const char* key, size_t klen, const gubo_item* item);
int gubo_pool_get_item(gubo_pool*, const char* key, size_t klen, gubo_item* out);
/* UpdatePeerGlobals item construction (gubernator.go:425-459) then AddCacheItem */
void gubo_pool_update_peer_global(gubo_pool*, const char* key, size_t klen, int32_t algorithm, int64_t duration,
                                  int32_t status, int64_t limit, int64_t remaining, int64_t reset_time);
/* pre-hashed keys (worker chosen from the XXH64 like gubo_submit_hashed) */
void gubo_pool_add_item_hashed(gubo_pool*, uint64_t key_xxh64, uint64_t key_fnv1, const gubo_item* item);
int gubo_pool_get_item_hashed(gubo_pool*, uint64_t key_xxh64, uint64_t key_fnv1, gubo_item* out);
void gubo_pool_update_peer_global_hashed(gubo_pool*, uint64_t key_xxh64, uint64_t key_fnv1, int32_t algorithm, int64_t duration,
                                         int32_t status, int64_t limit, int64_t remaining, int64_t reset_time);
int64_t gubo_pool_size(const gubo_pool*);
/* counters mirroring metricOverLimitCounter (gubernator.go:74), metricCacheAccess hit/miss (lrucache.go:52),
 * metricCacheUnexpiredEvictions (lrucache.go:56) */
void gubo_pool_counters(const gubo_pool*, int64_t out4[4]);
/* Each(): iterate all items (workers.go:451-534 Store path).  Returns number written (<= cap). */
size_t gubo_pool_each(gubo_pool*, gubo_item* items, uint64_t* key_xxh64, uint64_t* key_fnv1, size_t cap);

/* ReplicatedConsistentHash (replicated_hash.go:78-119).  hash_kind: 0 fnv1 (library default :33), 1 fnv1a. */
typedef struct gubo_ring gubo_ring;
gubo_ring* gubo_ring_new(int hash_kind, int replicas);
void gubo_ring_add(gubo_ring*, const char* grpc_address);
int gubo_ring_get(const gubo_ring*, const char* key, size_t len); /* index of peer in Add order; -1 if empty */
int gubo_ring_get_by_hash(const gubo_ring*, uint64_t hash);
size_t gubo_ring_points(const gubo_ring*, uint64_t* hashes, int32_t* peers, size_t cap);
void gubo_ring_free(gubo_ring*);

/* ---- pre-hashed batch form (what the device ABI consumes); used for differential tests and the CPU baseline.
 * Record layouts are byte-identical to include/gubernator_b200.h gub_req / gub_resp. */
typedef struct {
  uint64_t key_xxh64;  /* XXH64(Name+"_"+UniqueKey, 0) */
  uint64_t key_fnv1;   /* FNV-1 64 of the same string */
  int64_t hits, limit, duration, burst, created_at;
  uint32_t algorithm;  /* low 8 bits used */
  uint32_t behavior;   /* Behavior bits | GUBO_REQ_IS_OWNER */
} gubo_hreq;
#define GUBO_REQ_IS_OWNER 0x100u

typedef struct {
  uint32_t status;
  uint32_t err_code;
  int64_t limit, remaining, reset_time;
} gubo_hresp;

/* Sequential index-order evaluation of a pre-hashed batch.  Keys are the 16-byte (xxh64,fnv1) pair rendered as a
 * 32-hex-digit string so the string-keyed LRU/worker path above is exercised unchanged. */
void gubo_submit_hashed(gubo_pool*, const gubo_hreq* reqs, size_t n, gubo_hresp* out);

/* CPU baseline: the reference's worker-pool design on `threads` host threads: requests are routed by
 * XXH64>>1 / step (workers.go:180-184) to single-threaded shards and applied in index order per shard.
 * Returns elapsed seconds for the batch. */
double gubo_submit_hashed_mt(gubo_pool*, const gubo_hreq* reqs, size_t n, gubo_hresp* out, int threads);
/* the same from key strings: hashing (XXH64 + FNV-1 per key) happens inside the timed call */
double gubo_submit_keys_mt(gubo_pool* p, const char* bytes, const uint64_t* offsets, gubo_hreq* scratch, size_t n, gubo_hresp* out, int threads);

#ifdef __cplusplus
}
#endif
#endif
