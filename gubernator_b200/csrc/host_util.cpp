// host_util.cpp — host-side pieces of the boundary that are not device work: key hashing, the batch clock's Gregorian
// tables, and the replicated consistent-hash ring.  (Reference: workers.go:153, replicated_hash.go:78-119,
// interval.go:84-148 of mailgun/gubernator v2.4.0.)
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/gubernator_b200.h"

namespace {

// ---- XXH64 (what OneOfOne/xxhash ChecksumString64S computes) -------------------------------------------------
constexpr uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL,
                   P4 = 0x85EBCA77C2B2AE63ULL, P5 = 0x27D4EB2F165667C5ULL;
inline uint64_t rol(uint64_t v, int s) { return (v << s) | (v >> (64 - s)); }
template <typename T>
inline T load_le(const unsigned char* p) { T v; std::memcpy(&v, p, sizeof v); return v; }
inline uint64_t lane_step(uint64_t acc, uint64_t word) { return rol(acc + word * P2, 31) * P1; }
inline uint64_t fold_lane(uint64_t h, uint64_t lane) { return (h ^ lane_step(0, lane)) * P1 + P4; }

uint64_t xxh64_impl(const unsigned char* p, size_t len, uint64_t seed) {
  const unsigned char* const end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t lanes[4] = {seed + P1 + P2, seed + P2, seed, seed - P1};
    for (; end - p >= 32; p += 32)
      for (int l = 0; l < 4; l++) lanes[l] = lane_step(lanes[l], load_le<uint64_t>(p + 8 * l));
    h = rol(lanes[0], 1) + rol(lanes[1], 7) + rol(lanes[2], 12) + rol(lanes[3], 18);
    for (int l = 0; l < 4; l++) h = fold_lane(h, lanes[l]);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)len;
  for (; end - p >= 8; p += 8) h = rol(h ^ lane_step(0, load_le<uint64_t>(p)), 27) * P1 + P4;
  if (end - p >= 4) { h = rol(h ^ (load_le<uint32_t>(p) * P1), 23) * P2 + P3; p += 4; }
  for (; p < end; p++) h = rol(h ^ (*p * P5), 11) * P1;
  h = (h ^ (h >> 33)) * P2;
  h = (h ^ (h >> 29)) * P3;
  return h ^ (h >> 32);
}

constexpr uint64_t FNV_OFFSET = 0xCBF29CE484222325ULL, FNV_PRIME = 0x100000001B3ULL;

// ---- MD5 (crypto/md5 at replicated_hash.go:81) ---------------------------------------------------------------
struct Md5 {
  uint32_t s[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  static uint32_t k(int i) {
    static uint32_t table[64];
    static bool init = false;
    if (!init) {
      // K[i] = floor(2^32 * |sin(i + 1)|) — RFC 1321 §3.4
      static const double two32 = 4294967296.0;
      for (int j = 0; j < 64; j++) { double v = __builtin_fabs(__builtin_sin((double)(j + 1))); table[j] = (uint32_t)(uint64_t)(v * two32); }
      init = true;
    }
    return table[i];
  }
  void block(const unsigned char* b) {
    static const int rot[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
    uint32_t w[16];
    for (int i = 0; i < 16; i++) w[i] = load_le<uint32_t>(b + 4 * i);
    uint32_t a = s[0], bb = s[1], c = s[2], d = s[3];
    for (int i = 0; i < 64; i++) {
      const int round = i >> 4;
      uint32_t f; int g;
      switch (round) {
        case 0: f = d ^ (bb & (c ^ d)); g = i; break;
        case 1: f = c ^ (d & (bb ^ c)); g = (5 * i + 1) & 15; break;
        case 2: f = bb ^ c ^ d; g = (3 * i + 5) & 15; break;
        default: f = c ^ (bb | ~d); g = (7 * i) & 15; break;
      }
      const uint32_t x = a + f + k(i) + w[g];
      const int r = rot[round][i & 3];
      a = d; d = c; c = bb; bb += (x << r) | (x >> (32 - r));
    }
    s[0] += a; s[1] += bb; s[2] += c; s[3] += d;
  }
  std::string hex(const std::string& in) {
    std::string m = in;
    const uint64_t bits = (uint64_t)in.size() * 8;
    m.push_back((char)0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    for (int i = 0; i < 8; i++) m.push_back((char)((bits >> (8 * i)) & 0xFF));
    for (size_t o = 0; o < m.size(); o += 64) block(reinterpret_cast<const unsigned char*>(m.data()) + o);
    char out[33];
    for (int i = 0; i < 16; i++) std::snprintf(out + 2 * i, 3, "%02x", (s[i >> 2] >> (8 * (i & 3))) & 0xFF);
    return std::string(out, 32);
  }
};

// ---- UTC calendar helpers for the Gregorian tables ------------------------------------------------------------
int64_t utc_ms(int year, int mon /*1-12*/, int day) {
  std::tm tmv{};
  tmv.tm_year = year - 1900; tmv.tm_mon = mon - 1; tmv.tm_mday = day;
  return (int64_t)timegm(&tmv) * 1000;  // timegm normalises mon == 12 into the next year
}

}  // namespace

struct gub_ring {
  int hash_kind = 0;
  int replicas = 512;
  std::vector<std::string> peers;
  struct Pt { uint64_t h; int32_t peer; };
  std::vector<Pt> pts;
  uint64_t version = 0;
  uint64_t hash(const std::string& s) const { return hash_kind == 1 ? gub_fnv1a_64(s.data(), s.size()) : gub_fnv1_64(s.data(), s.size()); }
};

extern "C" {

uint64_t gub_xxh64(const void* data, size_t len, uint64_t seed) { return xxh64_impl(static_cast<const unsigned char*>(data), len, seed); }

uint64_t gub_fnv1_64(const void* data, size_t len) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t h = FNV_OFFSET;
  while (len--) h = (h * FNV_PRIME) ^ *p++;
  return h;
}
uint64_t gub_fnv1a_64(const void* data, size_t len) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t h = FNV_OFFSET;
  while (len--) h = (h ^ *p++) * FNV_PRIME;
  return h;
}

int gub_hash_keys(const char* bytes, const uint64_t* offsets, size_t n, uint64_t* xxh64_out, uint64_t* fnv1_out) {
  if (n && (!bytes || !offsets || !xxh64_out || !fnv1_out)) return -1;
  for (size_t i = 0; i < n; i++) {
    const char* p = bytes + offsets[i];
    const size_t len = (size_t)(offsets[i + 1] - offsets[i]);
    xxh64_out[i] = gub_xxh64(p, len, 0);
    fnv1_out[i] = gub_fnv1_64(p, len);
  }
  return 0;
}

// GregorianExpiration / GregorianDuration for d = 0..5 (interval.go:84-148); entry 3 (weeks) is unused: the kernels
// answer GUB_ERR_GREGORIAN_WEEKS for it.  Months/years durations keep the reference's precedence slip
// `end.UnixNano() - begin.UnixNano()/1000000` (interval.go:99,105).
int gub_clock_fill(int64_t now_ms, gub_clock* out) {
  if (!out) return -1;
  std::memset(out, 0, sizeof *out);
  out->now_ms = now_ms;
  int64_t secs = now_ms / 1000;
  if (now_ms % 1000 < 0) secs -= 1;
  std::time_t tt = (std::time_t)secs;
  std::tm g{};
  gmtime_r(&tt, &g);
  const int Y = g.tm_year + 1900, M = g.tm_mon + 1, D = g.tm_mday;
  const int64_t day0 = utc_ms(Y, M, D);
  const int64_t min0 = day0 + ((int64_t)g.tm_hour * 60 + g.tm_min) * 60000;
  const int64_t hour0 = day0 + (int64_t)g.tm_hour * 3600000;
  const int64_t month0 = utc_ms(Y, M, 1), month1 = utc_ms(Y, M + 1, 1);
  const int64_t year0 = utc_ms(Y, 1, 1), year1 = utc_ms(Y + 1, 1, 1);
  out->greg_expire[0] = min0 + 60000 - 1;
  out->greg_expire[1] = hour0 + 3600000 - 1;
  out->greg_expire[2] = day0 + 86400000 - 1;
  out->greg_expire[4] = month1 - 1;
  out->greg_expire[5] = year1 - 1;
  out->greg_duration[0] = 60000;
  out->greg_duration[1] = 3600000;
  out->greg_duration[2] = 86400000;
  out->greg_duration[4] = (month1 * 1000000 - 1) - (month0 * 1000000) / 1000000;
  out->greg_duration[5] = (year1 * 1000000 - 1) - (year0 * 1000000) / 1000000;
  return 0;
}

gub_ring* gub_ring_create(int hash_kind, int replicas) {
  gub_ring* r = new gub_ring();
  r->hash_kind = hash_kind;
  r->replicas = replicas > 0 ? replicas : 512;  // defaultReplicas, replicated_hash.go:29
  return r;
}
void gub_ring_destroy(gub_ring* r) { delete r; }

int gub_ring_add(gub_ring* r, const char* grpc_address) {  // replicated_hash.go:78-91
  if (!r || !grpc_address) return -1;
  const int32_t id = (int32_t)r->peers.size();
  r->peers.emplace_back(grpc_address);
  const std::string digest = Md5().hex(r->peers.back());
  for (int i = 0; i < r->replicas; i++) r->pts.push_back({r->hash(std::to_string(i) + digest), id});
  std::sort(r->pts.begin(), r->pts.end(), [](const gub_ring::Pt& a, const gub_ring::Pt& b) { return a.h < b.h; });
  r->version++;
  return id;
}
int gub_ring_size(const gub_ring* r) { return r ? (int)r->peers.size() : 0; }

int gub_ring_get_by_hash(const gub_ring* r, uint64_t h) {  // replicated_hash.go:104-119
  if (!r || r->peers.empty()) return -1;
  auto it = std::lower_bound(r->pts.begin(), r->pts.end(), h, [](const gub_ring::Pt& p, uint64_t v) { return p.h < v; });
  if (it == r->pts.end()) it = r->pts.begin();
  return it->peer;
}
int gub_ring_get(const gub_ring* r, const char* key, size_t len) {
  if (!r || r->peers.empty()) return -1;
  return gub_ring_get_by_hash(r, r->hash(std::string(key, len)));
}
size_t gub_ring_points(const gub_ring* r, uint64_t* hashes, int32_t* peers, size_t cap) {
  if (!r) return 0;
  const size_t n = std::min(cap, r->pts.size());
  for (size_t i = 0; i < n; i++) { if (hashes) hashes[i] = r->pts[i].h; if (peers) peers[i] = r->pts[i].peer; }
  return r->pts.size();
}

}  // extern "C"

// used by gub_api.cu
extern "C" uint64_t gub_ring_version_(const gub_ring* r) { return r ? r->version : 0; }
