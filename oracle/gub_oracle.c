/*
 * gub_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See gub_oracle.h for scope and pinning.
 *
 * Every function cites the reference file:line it restates (paths relative to mailgun/gubernator v2.4.0).
 * Build: see oracle/Makefile (gcc -O2 -fwrapv -ffp-contract=off: Go integer arithmetic wraps, amd64 Go never
 * fuses multiply-add).
 */
#define _GNU_SOURCE
#include "gub_oracle.h"

#include <emmintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------------------
 * Go numeric conversions on amd64.
 * int64(f): CVTTSD2SI — truncates toward zero; NaN and out-of-range give the "integer indefinite" INT64_MIN.
 * float64(i): CVTSI2SD — round to nearest even.
 * ---------------------------------------------------------------------------------------------------------- */
static inline int64_t go_f2i(double f) { return (int64_t)_mm_cvttsd_si64(_mm_set_sd(f)); }
static inline double go_i2f(int64_t i) { return (double)i; }

/* ------------------------------------------------------------------------------------------------------------
 * XXH64 — github.com/OneOfOne/xxhash v1.2.8 ChecksumString64S(input, 0) (workers.go:153-155) is the
 * standard XXH64; restated from the published xxHash specification.
 * ---------------------------------------------------------------------------------------------------------- */
#define XP1 11400714785074694791ULL
#define XP2 14029467366897019727ULL
#define XP3 1609587929392839161ULL
#define XP4 9650029242287828579ULL
#define XP5 2870177450012600261ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t xxh_round(uint64_t acc, uint64_t in) { acc += in * XP2; acc = rotl64(acc, 31); return acc * XP1; }
static inline uint64_t xxh_merge(uint64_t acc, uint64_t v) { acc ^= xxh_round(0, v); return acc * XP1 + XP4; }

uint64_t gubo_xxh64(const void* data, size_t len, uint64_t seed) {
  const uint8_t* p = (const uint8_t*)data;
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
    const uint8_t* lim = end - 32;
    do {
      v1 = xxh_round(v1, rd64(p)); v2 = xxh_round(v2, rd64(p + 8));
      v3 = xxh_round(v3, rd64(p + 16)); v4 = xxh_round(v4, rd64(p + 24));
      p += 32;
    } while (p <= lim);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
  } else {
    h = seed + XP5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) { h ^= xxh_round(0, rd64(p)); h = rotl64(h, 27) * XP1 + XP4; p += 8; }
  if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * XP1; h = rotl64(h, 23) * XP2 + XP3; p += 4; }
  while (p < end) { h ^= (*p) * XP5; h = rotl64(h, 11) * XP1; p++; }
  h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
  return h;
}

/* FNV-1 / FNV-1a 64 — segmentio/fasthash v1.0.2 fnv1.HashString64 / fnv1a.HashString64
 * (replicated_hash.go:33,83,108; config.go:430-433). */
uint64_t gubo_fnv1_64(const void* data, size_t len) {
  const uint8_t* p = (const uint8_t*)data;
  uint64_t h = 14695981039346656037ULL;
  for (size_t i = 0; i < len; i++) { h *= 1099511628211ULL; h ^= p[i]; }
  return h;
}
uint64_t gubo_fnv1a_64(const void* data, size_t len) {
  const uint8_t* p = (const uint8_t*)data;
  uint64_t h = 14695981039346656037ULL;
  for (size_t i = 0; i < len; i++) { h ^= p[i]; h *= 1099511628211ULL; }
  return h;
}

/* MD5 (RFC 1321) — crypto/md5 at replicated_hash.go:81; hex is lower-case "%x". */
static const uint32_t MD5_K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8,
    0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340,
    0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87,
    0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
    0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039,
    0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92,
    0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb,
    0xeb86d391};
static const uint8_t MD5_S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                                  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
static void md5_block(uint32_t st[4], const uint8_t* blk) {
  uint32_t m[16], a = st[0], b = st[1], c = st[2], d = st[3];
  for (int i = 0; i < 16; i++) m[i] = rd32(blk + 4 * i);
  for (int i = 0; i < 64; i++) {
    uint32_t f; int g;
    if (i < 16) { f = (b & c) | (~b & d); g = i; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; }
    uint32_t t = a + f + MD5_K[i] + m[g];
    a = d; d = c; c = b;
    b = b + ((t << MD5_S[i]) | (t >> (32 - MD5_S[i])));
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}
void gubo_md5_hex(const void* data, size_t len, char out33[33]) {
  uint32_t st[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
  const uint8_t* p = (const uint8_t*)data;
  size_t full = len / 64;
  for (size_t i = 0; i < full; i++) md5_block(st, p + 64 * i);
  uint8_t tail[128];
  size_t rem = len - full * 64;
  memset(tail, 0, sizeof tail);
  memcpy(tail, p + full * 64, rem);
  tail[rem] = 0x80;
  size_t tl = (rem < 56) ? 64 : 128;
  uint64_t bits = (uint64_t)len * 8;
  memcpy(tail + tl - 8, &bits, 8);
  md5_block(st, tail);
  if (tl == 128) md5_block(st, tail + 64);
  static const char* hx = "0123456789abcdef";
  for (int i = 0; i < 16; i++) {
    uint8_t byte = (uint8_t)(st[i / 4] >> (8 * (i % 4)));
    out33[2 * i] = hx[byte >> 4];
    out33[2 * i + 1] = hx[byte & 15];
  }
  out33[32] = 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * Gregorian intervals — interval.go:74-148, now.Location() = UTC.
 * ---------------------------------------------------------------------------------------------------------- */
static int64_t days_from_civil(int64_t y, int m, int d) { /* proleptic Gregorian; days since 1970-01-01 */
  y -= m <= 2;
  int64_t era = (y >= 0 ? y : y - 399) / 400;
  int64_t yoe = y - era * 400;
  int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}
static void civil_from_days(int64_t z, int64_t* y, int* m, int* d) {
  z += 719468;
  int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  int64_t doe = z - era * 146097;
  int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  int64_t yy = yoe + era * 400;
  int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  int64_t mp = (5 * doy + 2) / 153;
  *d = (int)(doy - (153 * mp + 2) / 5 + 1);
  *m = (int)(mp < 10 ? mp + 3 : mp - 9);
  *y = yy + (*m <= 2);
}
static int64_t floordiv(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
#define MS_PER_DAY 86400000LL

/* interval.go:84-109.  Months/years reproduce the operator-precedence bug
 * `end.UnixNano() - begin.UnixNano()/1000000` (interval.go:99,105). */
int gubo_gregorian_duration(int64_t now_ms, int64_t d, int64_t* out) {
  *out = 0;
  int64_t y; int m, dd;
  civil_from_days(floordiv(now_ms, MS_PER_DAY), &y, &m, &dd);
  switch (d) {
    case 0: *out = 60000; return GUBO_OK;       /* GregorianMinutes :87 */
    case 1: *out = 3600000; return GUBO_OK;     /* GregorianHours   :89 */
    case 2: *out = 86400000; return GUBO_OK;    /* GregorianDays    :91 */
    case 3: return GUBO_ERR_GREGORIAN_WEEKS;    /* :93 */
    case 4: {                                   /* GregorianMonths :94-99 */
      int64_t begin_ns = days_from_civil(y, m, 1) * MS_PER_DAY * 1000000LL;
      int ny = (m == 12) ? 1 : 0;
      int64_t end_ns = days_from_civil(y + ny, m == 12 ? 1 : m + 1, 1) * MS_PER_DAY * 1000000LL - 1;
      *out = end_ns - begin_ns / 1000000;
      return GUBO_OK;
    }
    case 5: {                                   /* GregorianYears :100-105 */
      int64_t begin_ns = days_from_civil(y, 1, 1) * MS_PER_DAY * 1000000LL;
      int64_t end_ns = days_from_civil(y + 1, 1, 1) * MS_PER_DAY * 1000000LL - 1;
      *out = end_ns - begin_ns / 1000000;
      return GUBO_OK;
    }
  }
  return GUBO_ERR_GREGORIAN_INVALID; /* :107 */
}

/* interval.go:117-148: end of the current interval in epoch ms (each is "start of next interval - 1ns", /1e6). */
int gubo_gregorian_expiration(int64_t now_ms, int64_t d, int64_t* out) {
  *out = 0;
  int64_t day = floordiv(now_ms, MS_PER_DAY);
  int64_t y; int m, dd;
  civil_from_days(day, &y, &m, &dd);
  switch (d) {
    case 0: *out = floordiv(now_ms, 60000) * 60000 + 59999; return GUBO_OK;       /* :119-122 */
    case 1: *out = floordiv(now_ms, 3600000) * 3600000 + 3599999; return GUBO_OK; /* :123-128 */
    case 2: *out = day * MS_PER_DAY + MS_PER_DAY - 1; return GUBO_OK;             /* :129-132 */
    case 3: return GUBO_ERR_GREGORIAN_WEEKS;                                      /* :133-134 */
    case 4: {                                                                     /* :135-139 */
      int ny = (m == 12) ? 1 : 0;
      *out = days_from_civil(y + ny, m == 12 ? 1 : m + 1, 1) * MS_PER_DAY - 1;
      return GUBO_OK;
    }
    case 5: *out = days_from_civil(y + 1, 1, 1) * MS_PER_DAY - 1; return GUBO_OK; /* :140-145 */
  }
  return GUBO_ERR_GREGORIAN_INVALID; /* :147 */
}

/* ------------------------------------------------------------------------------------------------------------
 * LRUCache — lrucache.go:32-178: map[string]*list.Element + container/list, lazy expiry, not thread-safe.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct node {
  struct node* hnext;       /* hash chain (Go map) */
  struct node *prev, *next; /* container/list */
  uint64_t h;
  uint32_t klen;
  gubo_item item;
  char key[];
} node;

typedef struct {
  node** buckets;
  size_t nb;        /* power of two */
  size_t len;       /* ll.Len() */
  node sentinel;    /* list root: sentinel.next = front, sentinel.prev = back */
  int64_t cache_size;
  const int64_t* now; /* frozen clock shared with the pool */
  int64_t hit, miss, unexpired_evictions;
} lru;

static void lru_init(lru* c, int64_t max_size, const int64_t* now) { /* lrucache.go:62-71 */
  if (max_size == 0) max_size = 50000; /* setter.SetDefault(&maxSize, 50_000) */
  memset(c, 0, sizeof *c);
  c->nb = 1024;
  c->buckets = (node**)calloc(c->nb, sizeof(node*));
  c->sentinel.next = c->sentinel.prev = &c->sentinel;
  c->cache_size = max_size;
  c->now = now;
}
static node* lru_find(lru* c, const char* key, size_t klen, uint64_t h) {
  for (node* n = c->buckets[h & (c->nb - 1)]; n; n = n->hnext)
    if (n->h == h && n->klen == klen && memcmp(n->key, key, klen) == 0) return n;
  return NULL;
}
static void lru_grow(lru* c) {
  size_t nnb = c->nb * 2;
  node** nbk = (node**)calloc(nnb, sizeof(node*));
  for (size_t i = 0; i < c->nb; i++) {
    node* n = c->buckets[i];
    while (n) { node* nx = n->hnext; n->hnext = nbk[n->h & (nnb - 1)]; nbk[n->h & (nnb - 1)] = n; n = nx; }
  }
  free(c->buckets);
  c->buckets = nbk;
  c->nb = nnb;
}
static void list_unlink(node* n) { n->prev->next = n->next; n->next->prev = n->prev; }
static void list_push_front(lru* c, node* n) {
  n->next = c->sentinel.next; n->prev = &c->sentinel;
  c->sentinel.next->prev = n; c->sentinel.next = n;
}
static void lru_remove_element(lru* c, node* e) { /* lrucache.go:151-156 */
  list_unlink(e);
  node** pp = &c->buckets[e->h & (c->nb - 1)];
  while (*pp != e) pp = &(*pp)->hnext;
  *pp = e->hnext;
  c->len--;
  free(e);
}
static void lru_remove_oldest(lru* c) { /* lrucache.go:138-149 */
  node* ele = c->sentinel.prev;
  if (ele != &c->sentinel) {
    if (*c->now < ele->item.expire_at) c->unexpired_evictions++;
    lru_remove_element(c, ele);
  }
}
/* lrucache.go:88-103.  Returns the stored node (the Go code stores the caller's *CacheItem). */
static node* lru_add(lru* c, const char* key, size_t klen, uint64_t h, const gubo_item* item) {
  node* ee = lru_find(c, key, klen, h);
  if (ee) { /* :90-94 */
    list_unlink(ee);
    list_push_front(c, ee);
    ee->item = *item;
    return ee;
  }
  node* n = (node*)malloc(sizeof(node) + klen);
  n->h = h; n->klen = (uint32_t)klen; n->item = *item;
  memcpy(n->key, key, klen);
  list_push_front(c, n); /* :96 */
  if (c->len + 1 > c->nb) lru_grow(c);
  n->hnext = c->buckets[h & (c->nb - 1)];
  c->buckets[h & (c->nb - 1)] = n; /* :97 */
  c->len++;
  if (c->cache_size != 0 && (int64_t)c->len > c->cache_size) { /* :98-100 */
    node* oldest = c->sentinel.prev;
    lru_remove_oldest(c);
    if (oldest == n) return NULL; /* cannot happen: n is at the front */
  }
  return n;
}
static int item_is_expired(const gubo_item* it, int64_t now) { /* cache.go:43-57 */
  if (it->invalid_at != 0 && it->invalid_at < now) return 1;
  if (it->expire_at < now) return 1;
  return 0;
}
static node* lru_get_item(lru* c, const char* key, size_t klen, uint64_t h) { /* lrucache.go:111-128 */
  node* ele = lru_find(c, key, klen, h);
  if (ele) {
    if (item_is_expired(&ele->item, *c->now)) { /* :115-119 */
      lru_remove_element(c, ele);
      c->miss++;
      return NULL;
    }
    c->hit++; /* :121 */
    list_unlink(ele);
    list_push_front(c, ele); /* :122 */
    return ele;
  }
  c->miss++; /* :126 */
  return NULL;
}
static void lru_remove(lru* c, const char* key, size_t klen, uint64_t h) { /* lrucache.go:131-135 */
  node* ele = lru_find(c, key, klen, h);
  if (ele) lru_remove_element(c, ele);
}
static void lru_free(lru* c) {
  node* n = c->sentinel.next;
  while (n != &c->sentinel) { node* nx = n->next; free(n); n = nx; }
  free(c->buckets);
}

/* ------------------------------------------------------------------------------------------------------------
 * WorkerPool — workers.go:54-184
 * ---------------------------------------------------------------------------------------------------------- */
struct mt_ctx;
struct gubo_pool {
  int workers;
  uint64_t hash_ring_step; /* workers.go:134 */
  lru* caches;             /* one per worker (workers.go:166) */
  int64_t now_ms;
  int64_t over_limit;      /* metricOverLimitCounter, gubernator.go:74 (per worker below when MT) */
  int64_t* over_limit_w;
  struct mt_ctx* mt;
};

gubo_pool* gubo_pool_new(int workers, int64_t cache_size) { /* workers.go:125-151 */
  if (workers <= 0) workers = 1;
  if (cache_size == 0) cache_size = 50000; /* :126 */
  gubo_pool* p = (gubo_pool*)calloc(1, sizeof *p);
  p->workers = workers;
  p->hash_ring_step = (1ULL << 63) / (uint64_t)workers; /* :134 */
  p->caches = (lru*)calloc((size_t)workers, sizeof(lru));
  p->over_limit_w = (int64_t*)calloc((size_t)workers, sizeof(int64_t));
  for (int i = 0; i < workers; i++) lru_init(&p->caches[i], cache_size / workers, &p->now_ms); /* :132 */
  return p;
}
static void mt_free(struct mt_ctx*);
void gubo_pool_free(gubo_pool* p) {
  if (!p) return;
  if (p->mt) mt_free(p->mt);
  for (int i = 0; i < p->workers; i++) lru_free(&p->caches[i]);
  free(p->caches);
  free(p->over_limit_w);
  free(p);
}
int gubo_pool_worker_index_for_hash63(const gubo_pool* p, uint64_t hash63) { /* workers.go:180-184 */
  uint64_t idx = hash63 / p->hash_ring_step;
  /* For worker counts that do not divide 2^63 the top (2^63 mod W) hashes index one past the end and the Go code
   * would panic; clamp (probability < 2^-60). */
  if (idx >= (uint64_t)p->workers) idx = (uint64_t)p->workers - 1;
  return (int)idx;
}
int gubo_pool_worker_index(const gubo_pool* p, const char* key, size_t len) {
  return gubo_pool_worker_index_for_hash63(p, gubo_xxh64(key, len, 0) >> 1); /* workers.go:153-155 */
}
void gubo_pool_set_now(gubo_pool* p, int64_t now_ms) { p->now_ms = now_ms; }
int64_t gubo_pool_now(const gubo_pool* p) { return p->now_ms; }
int64_t gubo_pool_size(const gubo_pool* p) {
  int64_t s = 0;
  for (int i = 0; i < p->workers; i++) s += (int64_t)p->caches[i].len;
  return s;
}
void gubo_pool_counters(const gubo_pool* p, int64_t out4[4]) {
  out4[0] = 0; out4[1] = 0; out4[2] = 0; out4[3] = 0;
  for (int i = 0; i < p->workers; i++) {
    out4[0] += p->over_limit_w[i];
    out4[1] += p->caches[i].hit;
    out4[2] += p->caches[i].miss;
    out4[3] += p->caches[i].unexpired_evictions;
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * algorithms.go
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { int32_t status; int64_t limit, remaining, reset_time; } rl_t;
#define HAS(b, f) (((b) & (f)) != 0) /* gubernator.go:776-779 */

typedef struct {
  lru* c;
  int64_t now;
  int64_t* over_limit;
  const char* key; size_t klen; uint64_t kh;
} actx;

/* algorithms.go:206-257 */
static int token_bucket_new_item(actx* a, const gubo_req* r, int is_owner, rl_t* rl) {
  int64_t created_at = r->created_at;
  int64_t expire = created_at + r->duration; /* :208 */
  gubo_item it;
  memset(&it, 0, sizeof it);
  it.value_kind = 1;
  it.limit = r->limit; it.duration = r->duration; it.remaining_i = r->limit - r->hits; it.stamp = created_at; /* :210-215 */
  if (HAS(r->behavior, GUBO_DURATION_IS_GREGORIAN)) { /* :218-223 */
    int err = gubo_gregorian_expiration(a->now, r->duration, &expire);
    if (err) return err;
  }
  it.algorithm = GUBO_TOKEN_BUCKET; it.expire_at = expire; /* :225-230 */
  rl->status = GUBO_UNDER_LIMIT; rl->limit = r->limit; rl->remaining = it.remaining_i; rl->reset_time = expire; /* :232-237 */
  if (r->hits > r->limit) { /* :240-248 */
    if (is_owner) (*a->over_limit)++;
    rl->status = GUBO_OVER_LIMIT;
    rl->remaining = r->limit;
    it.remaining_i = r->limit;
  }
  lru_add(a->c, a->key, a->klen, a->kh, &it); /* :250 */
  return GUBO_OK;
}

/* algorithms.go:37-203 (Store == nil) */
static int token_bucket(actx* a, const gubo_req* r, int is_owner, rl_t* rl) {
  node* item = lru_get_item(a->c, a->key, a->klen, a->kh); /* :43 */
  int ok = item != NULL;
  if (ok && item->item.value_kind == 0) ok = 0; /* :55-63 "Value is nil" */
  if (ok) {
    if (HAS(r->behavior, GUBO_RESET_REMAINING)) { /* :78-90 */
      lru_remove(a->c, a->key, a->klen, a->kh);
      rl->status = GUBO_UNDER_LIMIT; rl->limit = r->limit; rl->remaining = r->limit; rl->reset_time = 0;
      return GUBO_OK;
    }
    gubo_item* t = &item->item;
    if (t->value_kind != 1) { /* :91-103 client switched algorithms */
      lru_remove(a->c, a->key, a->klen, a->kh);
      return token_bucket_new_item(a, r, is_owner, rl);
    }
    if (t->limit != r->limit) { /* :106-113 */
      t->remaining_i += r->limit - t->limit;
      if (t->remaining_i < 0) t->remaining_i = 0;
      t->limit = r->limit;
    }
    rl->status = t->status; rl->limit = r->limit; rl->remaining = t->remaining_i; rl->reset_time = t->expire_at; /* :115-120 */
    if (t->duration != r->duration) { /* :123-147 */
      int64_t expire = t->stamp + r->duration;
      if (HAS(r->behavior, GUBO_DURATION_IS_GREGORIAN)) {
        int err = gubo_gregorian_expiration(a->now, r->duration, &expire);
        if (err) return err;
      }
      int64_t created_at = r->created_at;
      if (expire <= created_at) { /* :136-142 renew; rl.Remaining intentionally NOT refreshed */
        expire = created_at + r->duration;
        t->stamp = created_at;
        t->remaining_i = t->limit;
      }
      t->expire_at = expire;
      t->duration = r->duration;
      rl->reset_time = expire;
    }
    if (r->hits == 0) return GUBO_OK; /* :157-159 */
    if (rl->remaining == 0 && r->hits > 0) { /* :162-170 */
      if (is_owner) (*a->over_limit)++;
      rl->status = GUBO_OVER_LIMIT;
      t->status = rl->status;
      return GUBO_OK;
    }
    if (t->remaining_i == r->hits) { /* :173-178 */
      t->remaining_i = 0;
      rl->remaining = 0;
      return GUBO_OK;
    }
    if (r->hits > t->remaining_i) { /* :182-194 */
      if (is_owner) (*a->over_limit)++;
      rl->status = GUBO_OVER_LIMIT;
      if (HAS(r->behavior, GUBO_DRAIN_OVER_LIMIT)) { t->remaining_i = 0; rl->remaining = 0; }
      return GUBO_OK;
    }
    t->remaining_i -= r->hits; /* :196-198 */
    rl->remaining = t->remaining_i;
    return GUBO_OK;
  }
  return token_bucket_new_item(a, r, is_owner, rl); /* :202 */
}

/* algorithms.go:437-493 */
static int leaky_bucket_new_item(actx* a, const gubo_req* r, int is_owner, rl_t* rl) {
  int64_t created_at = r->created_at;
  int64_t duration = r->duration;
  double rate = go_i2f(duration) / go_i2f(r->limit); /* :440 raw duration even under Gregorian */
  if (HAS(r->behavior, GUBO_DURATION_IS_GREGORIAN)) { /* :441-450 */
    int64_t expire;
    int err = gubo_gregorian_expiration(a->now, r->duration, &expire);
    if (err) return err;
    duration = expire - a->now; /* n.UnixNano()/1000000 */
  }
  gubo_item b;
  memset(&b, 0, sizeof b);
  b.value_kind = 2;
  b.remaining_f = go_i2f(r->burst - r->hits); b.limit = r->limit; b.duration = duration; b.stamp = created_at; b.burst = r->burst; /* :453-459 */
  rl->status = GUBO_UNDER_LIMIT; rl->limit = b.limit; rl->remaining = r->burst - r->hits;
  rl->reset_time = created_at + (b.limit - (r->burst - r->hits)) * go_f2i(rate); /* :461-466 */
  if (r->hits > r->burst) { /* :469-477 */
    if (is_owner) (*a->over_limit)++;
    rl->status = GUBO_OVER_LIMIT;
    rl->remaining = 0;
    rl->reset_time = created_at + (rl->limit - rl->remaining) * go_f2i(rate);
    b.remaining_f = 0;
  }
  b.expire_at = created_at + duration; b.algorithm = r->algorithm; /* :479-484 */
  lru_add(a->c, a->key, a->klen, a->kh, &b); /* :486 */
  return GUBO_OK;
}

/* algorithms.go:260-434 (Store == nil).  Mutates r->burst like the reference (:264-266). */
static int leaky_bucket(actx* a, gubo_req* r, int is_owner, rl_t* rl) {
  if (r->burst == 0) r->burst = r->limit; /* :264-266 */
  int64_t created_at = r->created_at;     /* :268 */
  node* item = lru_get_item(a->c, a->key, a->klen, a->kh); /* :272 */
  int ok = item != NULL;
  if (ok && item->item.value_kind == 0) ok = 0; /* :284-292 */
  if (ok) {
    gubo_item* b = &item->item;
    if (b->value_kind != 2) { /* :308-318 */
      lru_remove(a->c, a->key, a->klen, a->kh);
      return leaky_bucket_new_item(a, r, is_owner, rl);
    }
    if (HAS(r->behavior, GUBO_RESET_REMAINING)) b->remaining_f = go_i2f(r->burst); /* :320-322 */
    if (b->burst != r->burst) { /* :325-330 */
      if (r->burst > go_f2i(b->remaining_f)) b->remaining_f = go_i2f(r->burst);
      b->burst = r->burst;
    }
    b->limit = r->limit;       /* :332 */
    b->duration = r->duration; /* :333 */
    int64_t duration = r->duration;
    double rate = go_i2f(duration) / go_i2f(r->limit); /* :336 */
    if (HAS(r->behavior, GUBO_DURATION_IS_GREGORIAN)) { /* :338-354 */
      int64_t d, expire;
      int err = gubo_gregorian_duration(a->now, r->duration, &d);
      if (err) return err;
      err = gubo_gregorian_expiration(a->now, r->duration, &expire);
      if (err) return err;
      rate = go_i2f(d) / go_i2f(r->limit);
      duration = expire - a->now;
    }
    if (r->hits != 0) b->expire_at = created_at + duration; /* :356-358 UpdateExpiration, lrucache.go:164 */
    int64_t elapsed = created_at - b->stamp; /* :361 */
    double leak = go_i2f(elapsed) / rate;    /* :362 */
    if (go_f2i(leak) > 0) { /* :364-367 */
      b->remaining_f += leak;
      b->stamp = created_at;
    }
    if (go_f2i(b->remaining_f) > b->burst) b->remaining_f = go_i2f(b->burst); /* :369-371 */
    rl->limit = b->limit; rl->remaining = go_f2i(b->remaining_f); rl->status = GUBO_UNDER_LIMIT;
    rl->reset_time = created_at + (b->limit - go_f2i(b->remaining_f)) * go_f2i(rate); /* :373-378 */
    if (go_f2i(b->remaining_f) == 0 && r->hits > 0) { /* :389-395 */
      if (is_owner) (*a->over_limit)++;
      rl->status = GUBO_OVER_LIMIT;
      return GUBO_OK;
    }
    if (go_f2i(b->remaining_f) == r->hits) { /* :398-403 */
      b->remaining_f = 0;
      rl->remaining = go_f2i(b->remaining_f);
      rl->reset_time = created_at + (rl->limit - rl->remaining) * go_f2i(rate);
      return GUBO_OK;
    }
    if (r->hits > go_f2i(b->remaining_f)) { /* :407-420 */
      if (is_owner) (*a->over_limit)++;
      rl->status = GUBO_OVER_LIMIT;
      if (HAS(r->behavior, GUBO_DRAIN_OVER_LIMIT)) { b->remaining_f = 0; rl->remaining = 0; }
      return GUBO_OK;
    }
    if (r->hits == 0) return GUBO_OK; /* :423-425 */
    b->remaining_f -= go_i2f(r->hits); /* :427-430 */
    rl->remaining = go_f2i(b->remaining_f);
    rl->reset_time = created_at + (rl->limit - rl->remaining) * go_f2i(rate);
    return GUBO_OK;
  }
  return leaky_bucket_new_item(a, r, is_owner, rl); /* :433 */
}

/* Worker.handleGetRateLimit — workers.go:293-324 */
static int handle_get_rate_limit(gubo_pool* p, int widx, const char* key, size_t klen, gubo_req* r, int is_owner, rl_t* rl) {
  actx a;
  a.c = &p->caches[widx]; a.now = p->now_ms; a.over_limit = &p->over_limit_w[widx];
  a.key = key; a.klen = klen; a.kh = gubo_xxh64(key, klen, 0x9E3779B97F4A7C15ULL); /* the Go map's own hash */
  memset(rl, 0, sizeof *rl);
  int err;
  switch (r->algorithm) {
    case GUBO_TOKEN_BUCKET: err = token_bucket(&a, r, is_owner, rl); break;
    case GUBO_LEAKY_BUCKET: err = leaky_bucket(&a, r, is_owner, rl); break;
    default: err = GUBO_ERR_INVALID_ALGORITHM; break; /* :317-320 */
  }
  if (err) memset(rl, 0, sizeof *rl); /* rlResponse is nil on error */
  return err;
}

static const char* algo_wrap(int32_t algorithm) { return algorithm == GUBO_LEAKY_BUCKET ? "Error in leakyBucket" : "Error in tokenBucket"; }

/* The error chain as it reaches RateLimitResp.Error through GetRateLimits:
 * gubernator.go:252 "Error while apply rate limit for '<key>'" : gubernator.go:600 "during workerPool.GetRateLimit"
 * : workers.go:304,313 "Error in tokenBucket|leakyBucket" : interval.go:93,107 ; or workers.go:318. */
static void format_error(char* out, size_t cap, int err, const char* key, const gubo_req* r) {
  switch (err) {
    case GUBO_ERR_INVALID_ALGORITHM:
      snprintf(out, cap, "Error while apply rate limit for '%s': during workerPool.GetRateLimit: Invalid rate limit algorithm '%d'", key, r->algorithm);
      break;
    case GUBO_ERR_GREGORIAN_WEEKS:
      snprintf(out, cap, "Error while apply rate limit for '%s': during workerPool.GetRateLimit: %s: `Duration = GregorianWeeks` not yet supported; consider making a PR!`", key, algo_wrap(r->algorithm));
      break;
    case GUBO_ERR_GREGORIAN_INVALID:
      snprintf(out, cap, "Error while apply rate limit for '%s': during workerPool.GetRateLimit: %s: behavior DURATION_IS_GREGORIAN is set; but `Duration` is not a valid gregorian interval", key, algo_wrap(r->algorithm));
      break;
    default: out[0] = 0;
  }
}

void gubo_pool_get_rate_limit(gubo_pool* p, const char* key, size_t klen, gubo_req* r, int is_owner, gubo_resp* out) {
  rl_t rl;
  int widx = gubo_pool_worker_index(p, key, klen); /* workers.go:263 */
  int err = handle_get_rate_limit(p, widx, key, klen, r, is_owner, &rl);
  memset(out, 0, sizeof *out);
  out->err_code = err;
  if (err) { format_error(out->error, sizeof out->error, err, key, r); return; }
  out->status = rl.status; out->limit = rl.limit; out->remaining = rl.remaining; out->reset_time = rl.reset_time;
}

/* V1Instance.GetRateLimits — gubernator.go:183-295, every key locally owned */
int gubo_get_rate_limits(gubo_pool* p, const gubo_req* reqs, size_t n, gubo_resp* out, int is_owner, int unbounded) {
  if (!unbounded && n > 1000) return -1; /* :189-193 maxBatchSize (gubernator.go:40) */
  int64_t created_at = p->now_ms;         /* :195 */
  for (size_t i = 0; i < n; i++) {        /* :203 */
    gubo_req r = reqs[i];
    gubo_resp* o = &out[i];
    memset(o, 0, sizeof *o);
    const char* name = r.name ? r.name : "";
    const char* uk = r.unique_key ? r.unique_key : "";
    if (uk[0] == 0) { /* :208-212 */
      o->err_code = GUBO_ERR_UNIQUE_KEY_EMPTY;
      snprintf(o->error, sizeof o->error, "field 'unique_key' cannot be empty");
      continue;
    }
    if (name[0] == 0) { /* :213-217 */
      o->err_code = GUBO_ERR_NAMESPACE_EMPTY;
      snprintf(o->error, sizeof o->error, "field 'namespace' cannot be empty");
      continue;
    }
    if (r.created_at == 0) r.created_at = created_at; /* :218-220 */
    size_t ln = strlen(name), lu = strlen(uk);
    char stackbuf[512];
    char* key = (ln + lu + 2 <= sizeof stackbuf) ? stackbuf : (char*)malloc(ln + lu + 2);
    memcpy(key, name, ln); key[ln] = '_'; memcpy(key + ln + 1, uk, lu); key[ln + 1 + lu] = 0; /* :204, client.go:39-41 */
    gubo_pool_get_rate_limit(p, key, ln + 1 + lu, &r, is_owner, o); /* :250 -> :598 */
    if (key != stackbuf) free(key);
  }
  return 0;
}

/* workers.go:537-581 AddCacheItem -> cache.Add */
void gubo_pool_add_item(gubo_pool* p, const char* key, size_t klen, const gubo_item* item) {
  int widx = gubo_pool_worker_index(p, key, klen);
  lru_add(&p->caches[widx], key, klen, gubo_xxh64(key, klen, 0x9E3779B97F4A7C15ULL), item);
}
/* workers.go:583-626 GetCacheItem -> cache.GetItem */
int gubo_pool_get_item(gubo_pool* p, const char* key, size_t klen, gubo_item* out) {
  int widx = gubo_pool_worker_index(p, key, klen);
  node* n = lru_get_item(&p->caches[widx], key, klen, gubo_xxh64(key, klen, 0x9E3779B97F4A7C15ULL));
  if (!n) return 0;
  *out = n->item;
  return 1;
}
/* gubernator.go:425-459 */
void gubo_pool_update_peer_global(gubo_pool* p, const char* key, size_t klen, int32_t algorithm, int64_t duration,
                                  int32_t status, int64_t limit, int64_t remaining, int64_t reset_time) {
  int64_t now = p->now_ms; /* :427 */
  gubo_item it;
  memset(&it, 0, sizeof it);
  it.expire_at = reset_time; it.algorithm = algorithm; /* :429-433 */
  switch (algorithm) {
    case GUBO_LEAKY_BUCKET: /* :435-442 */
      it.value_kind = 2; it.remaining_f = go_i2f(remaining); it.limit = limit; it.duration = duration; it.burst = limit; it.stamp = now;
      break;
    case GUBO_TOKEN_BUCKET: /* :443-450 */
      it.value_kind = 1; it.status = status; it.limit = limit; it.duration = duration; it.remaining_i = remaining; it.stamp = now;
      break;
  }
  gubo_pool_add_item(p, key, klen, &it); /* :452 */
}

/* ------------------------------------------------------------------------------------------------------------
 * ReplicatedConsistentHash — replicated_hash.go:36-119
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { uint64_t hash; int32_t peer; } ring_pt;
struct gubo_ring { int hash_kind, replicas, npeers; ring_pt* pts; size_t npts; };
gubo_ring* gubo_ring_new(int hash_kind, int replicas) {
  gubo_ring* r = (gubo_ring*)calloc(1, sizeof *r);
  r->hash_kind = hash_kind; r->replicas = replicas > 0 ? replicas : 512; /* defaultReplicas :29 */
  return r;
}
static uint64_t ring_hash(const gubo_ring* r, const char* s, size_t n) { return r->hash_kind == 1 ? gubo_fnv1a_64(s, n) : gubo_fnv1_64(s, n); }
static int pt_cmp(const void* a, const void* b) {
  uint64_t x = ((const ring_pt*)a)->hash, y = ((const ring_pt*)b)->hash;
  return x < y ? -1 : (x > y ? 1 : 0);
}
void gubo_ring_add(gubo_ring* r, const char* addr) { /* :78-91 */
  char hex[33], buf[64];
  gubo_md5_hex(addr, strlen(addr), hex); /* :81 */
  r->pts = (ring_pt*)realloc(r->pts, (r->npts + (size_t)r->replicas) * sizeof(ring_pt));
  for (int i = 0; i < r->replicas; i++) { /* :82-88 */
    int n = snprintf(buf, sizeof buf, "%d%s", i, hex); /* strconv.Itoa(i) + key */
    r->pts[r->npts].hash = ring_hash(r, buf, (size_t)n);
    r->pts[r->npts].peer = r->npeers;
    r->npts++;
  }
  r->npeers++;
  qsort(r->pts, r->npts, sizeof(ring_pt), pt_cmp); /* :90 */
}
int gubo_ring_get_by_hash(const gubo_ring* r, uint64_t hash) { /* :104-119 */
  if (r->npeers == 0) return -1;
  size_t lo = 0, hi = r->npts; /* sort.Search: first i with pts[i].hash >= hash */
  while (lo < hi) { size_t mid = lo + (hi - lo) / 2; if (r->pts[mid].hash >= hash) hi = mid; else lo = mid + 1; }
  if (lo == r->npts) lo = 0; /* :114-116 */
  return r->pts[lo].peer;
}
int gubo_ring_get(const gubo_ring* r, const char* key, size_t len) { return gubo_ring_get_by_hash(r, ring_hash(r, key, len)); }
size_t gubo_ring_points(const gubo_ring* r, uint64_t* hashes, int32_t* peers, size_t cap) {
  size_t n = r->npts < cap ? r->npts : cap;
  for (size_t i = 0; i < n; i++) { hashes[i] = r->pts[i].hash; peers[i] = r->pts[i].peer; }
  return r->npts;
}
void gubo_ring_free(gubo_ring* r) { if (r) { free(r->pts); free(r); } }

/* ------------------------------------------------------------------------------------------------------------
 * Pre-hashed batches.  The cache key is the 16 raw bytes (xxh64, fnv1); the worker is chosen from the request's
 * own XXH64 exactly as workers.go:180-184 would from the string.
 * ---------------------------------------------------------------------------------------------------------- */
static inline void apply_hashed(gubo_pool* p, int widx, const gubo_hreq* h, gubo_hresp* o) {
  gubo_req r;
  rl_t rl;
  uint64_t key[2] = {h->key_xxh64, h->key_fnv1};
  r.name = r.unique_key = NULL;
  r.hits = h->hits; r.limit = h->limit; r.duration = h->duration; r.burst = h->burst; r.created_at = h->created_at;
  r.algorithm = (int32_t)(h->algorithm & 0xff); r.behavior = (int32_t)(h->behavior & 0xff);
  int err = handle_get_rate_limit(p, widx, (const char*)key, 16, &r, (h->behavior & GUBO_REQ_IS_OWNER) != 0, &rl);
  o->status = (uint32_t)rl.status; o->err_code = (uint32_t)err; o->limit = rl.limit; o->remaining = rl.remaining; o->reset_time = rl.reset_time;
}
void gubo_submit_hashed(gubo_pool* p, const gubo_hreq* reqs, size_t n, gubo_hresp* out) {
  for (size_t i = 0; i < n; i++)
    apply_hashed(p, gubo_pool_worker_index_for_hash63(p, reqs[i].key_xxh64 >> 1), &reqs[i], &out[i]);
}

/* Pre-hashed variants: the cache key is the 16 raw bytes (xxh64, fnv1) and the worker comes from the XXH64 itself, exactly
 * like gubo_submit_hashed, so items added here are the ones later requests see. */
static int hashed_widx(const gubo_pool* p, uint64_t kx) { return gubo_pool_worker_index_for_hash63(p, kx >> 1); }
void gubo_pool_add_item_hashed(gubo_pool* p, uint64_t kx, uint64_t kf, const gubo_item* item) {
  uint64_t key[2] = {kx, kf};
  lru_add(&p->caches[hashed_widx(p, kx)], (const char*)key, 16, gubo_xxh64(key, 16, 0x9E3779B97F4A7C15ULL), item);
}
int gubo_pool_get_item_hashed(gubo_pool* p, uint64_t kx, uint64_t kf, gubo_item* out) {
  uint64_t key[2] = {kx, kf};
  node* n = lru_get_item(&p->caches[hashed_widx(p, kx)], (const char*)key, 16, gubo_xxh64(key, 16, 0x9E3779B97F4A7C15ULL));
  if (!n) return 0;
  *out = n->item;
  return 1;
}
/* gubernator.go:425-459 for a pre-hashed key */
void gubo_pool_update_peer_global_hashed(gubo_pool* p, uint64_t kx, uint64_t kf, int32_t algorithm, int64_t duration, int32_t status,
                                         int64_t limit, int64_t remaining, int64_t reset_time) {
  int64_t now = p->now_ms;
  gubo_item it;
  memset(&it, 0, sizeof it);
  it.expire_at = reset_time; it.algorithm = algorithm;
  switch (algorithm) {
    case GUBO_LEAKY_BUCKET:
      it.value_kind = 2; it.remaining_f = go_i2f(remaining); it.limit = limit; it.duration = duration; it.burst = limit; it.stamp = now;
      break;
    case GUBO_TOKEN_BUCKET:
      it.value_kind = 1; it.status = status; it.limit = limit; it.duration = duration; it.remaining_i = remaining; it.stamp = now;
      break;
  }
  gubo_pool_add_item_hashed(p, kx, kf, &it);
}

size_t gubo_pool_each(gubo_pool* p, gubo_item* items, uint64_t* kx, uint64_t* kf, size_t cap) {
  size_t k = 0;
  for (int w = 0; w < p->workers; w++)
    for (node* n = p->caches[w].sentinel.next; n != &p->caches[w].sentinel; n = n->next) {
      if (k < cap) {
        items[k] = n->item;
        if (n->klen == 16) { memcpy(&kx[k], n->key, 8); memcpy(&kf[k], n->key + 8, 8); }
        else { kx[k] = gubo_xxh64(n->key, n->klen, 0); kf[k] = gubo_fnv1_64(n->key, n->klen); }
      }
      k++;
    }
  return k;
}

/* ------------------------------------------------------------------------------------------------------------
 * Multi-threaded worker-pool baseline.  T threads; phase 1 each thread routes a contiguous slice of the batch
 * (hash63 / step) into per-(thread, worker) index lists; phase 2 worker w (on thread w % T) drains the lists for
 * w in (thread, index) order == batch index order, i.e. exactly the per-shard serial order a WorkerPool sees when
 * one caller submits the batch in order.  This replaces the Go channel hand-off (workers.go:276,284) with a
 * barrier, which is strictly cheaper than the reference: a generous baseline.
 * ---------------------------------------------------------------------------------------------------------- */
struct mt_ctx {
  gubo_pool* p;
  int T;
  pthread_t* th;
  pthread_barrier_t bar;
  volatile int quit;
  const gubo_hreq* reqs; size_t n; gubo_hresp* out;
  uint32_t** lists;   /* [t*W + w] -> indices */
  size_t* counts;     /* [t*W + w] */
  size_t* caps;
};
typedef struct { struct mt_ctx* m; int tid; } mt_arg;

static void* mt_main(void* vp) {
  mt_arg* a = (mt_arg*)vp;
  struct mt_ctx* m = a->m;
  int t = a->tid, T = m->T, W = m->p->workers;
  for (;;) {
    pthread_barrier_wait(&m->bar); /* start */
    if (m->quit) break;
    size_t lo = m->n * (size_t)t / (size_t)T, hi = m->n * (size_t)(t + 1) / (size_t)T;
    for (int w = 0; w < W; w++) m->counts[(size_t)t * W + w] = 0;
    for (size_t i = lo; i < hi; i++) {
      int w = gubo_pool_worker_index_for_hash63(m->p, m->reqs[i].key_xxh64 >> 1);
      size_t k = (size_t)t * W + w;
      if (m->counts[k] == m->caps[k]) {
        m->caps[k] = m->caps[k] ? m->caps[k] * 2 : 1024;
        m->lists[k] = (uint32_t*)realloc(m->lists[k], m->caps[k] * sizeof(uint32_t));
      }
      m->lists[k][m->counts[k]++] = (uint32_t)i;
    }
    pthread_barrier_wait(&m->bar); /* routed */
    for (int w = t; w < W; w += T)
      for (int s = 0; s < T; s++) {
        size_t k = (size_t)s * W + w;
        for (size_t j = 0; j < m->counts[k]; j++) {
          uint32_t i = m->lists[k][j];
          apply_hashed(m->p, w, &m->reqs[i], &m->out[i]);
        }
      }
    pthread_barrier_wait(&m->bar); /* done */
  }
  free(a);
  return NULL;
}
static struct mt_ctx* mt_new(gubo_pool* p, int T) {
  struct mt_ctx* m = (struct mt_ctx*)calloc(1, sizeof *m);
  m->p = p; m->T = T;
  size_t k = (size_t)T * (size_t)p->workers;
  m->lists = (uint32_t**)calloc(k, sizeof(uint32_t*));
  m->counts = (size_t*)calloc(k, sizeof(size_t));
  m->caps = (size_t*)calloc(k, sizeof(size_t));
  pthread_barrier_init(&m->bar, NULL, (unsigned)T + 1);
  m->th = (pthread_t*)calloc((size_t)T, sizeof(pthread_t));
  for (int t = 0; t < T; t++) {
    mt_arg* a = (mt_arg*)malloc(sizeof *a);
    a->m = m; a->tid = t;
    pthread_create(&m->th[t], NULL, mt_main, a);
  }
  return m;
}
static void mt_free(struct mt_ctx* m) {
  m->quit = 1;
  pthread_barrier_wait(&m->bar);
  for (int t = 0; t < m->T; t++) pthread_join(m->th[t], NULL);
  pthread_barrier_destroy(&m->bar);
  size_t k = (size_t)m->T * (size_t)m->p->workers;
  for (size_t i = 0; i < k; i++) free(m->lists[i]);
  free(m->lists); free(m->counts); free(m->caps); free(m->th); free(m);
}
/* Same, starting from the key strings: every thread first hashes the keys of its slice — XXH64 as the worker pool does
 * (workers.go:153-155) and FNV-1 as the peer picker does (replicated_hash.go:108) — into `scratch` (a copy of reqs whose hash
 * fields are filled in), then the batch runs as above.  key i = bytes[offsets[i] .. offsets[i+1]). */
struct key_job { const char* bytes; const uint64_t* offsets; gubo_hreq* scratch; size_t n; int T; };
static void* key_hash_main(void* vp) {
  void** a = (void**)vp;
  struct key_job* j = (struct key_job*)a[0];
  int t = (int)(intptr_t)a[1];
  size_t lo = j->n * (size_t)t / (size_t)j->T, hi = j->n * (size_t)(t + 1) / (size_t)j->T;
  for (size_t i = lo; i < hi; i++) {
    const char* k = j->bytes + j->offsets[i];
    size_t len = (size_t)(j->offsets[i + 1] - j->offsets[i]);
    j->scratch[i].key_xxh64 = gubo_xxh64(k, len, 0);
    j->scratch[i].key_fnv1 = gubo_fnv1_64(k, len);
  }
  return NULL;
}
double gubo_submit_hashed_mt(gubo_pool* p, const gubo_hreq* reqs, size_t n, gubo_hresp* out, int threads);
double gubo_submit_keys_mt(gubo_pool* p, const char* bytes, const uint64_t* offsets, gubo_hreq* scratch, size_t n, gubo_hresp* out, int threads) {
  if (threads < 1) threads = 1;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  struct key_job job = {bytes, offsets, scratch, n, threads};
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  void** args = (void**)calloc((size_t)threads * 2, sizeof(void*));
  for (int t = 0; t < threads; t++) { args[2 * t] = &job; args[2 * t + 1] = (void*)(intptr_t)t; pthread_create(&th[t], NULL, key_hash_main, &args[2 * t]); }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th); free(args);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double hash_s = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  return hash_s + gubo_submit_hashed_mt(p, scratch, n, out, threads);
}

double gubo_submit_hashed_mt(gubo_pool* p, const gubo_hreq* reqs, size_t n, gubo_hresp* out, int threads) {
  if (threads < 1) threads = 1;
  if (p->mt && p->mt->T != threads) { mt_free(p->mt); p->mt = NULL; }
  if (!p->mt) p->mt = mt_new(p, threads);
  struct mt_ctx* m = p->mt;
  m->reqs = reqs; m->n = n; m->out = out;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&m->bar);
  pthread_barrier_wait(&m->bar);
  pthread_barrier_wait(&m->bar);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
