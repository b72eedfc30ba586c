// host_v1.cpp — C++ stand-in for the Go side of the boundary (include/gubernator_b200_host.h): the request loop of
// V1Instance.GetRateLimits (gubernator.go:183-295) re-expressed as "validate, hash, one device batch, map errors".
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gubernator_b200_host.h"

struct gub_instance {
  gub_table* table = nullptr;
  int64_t frozen_now = -1;
};

namespace {
int64_t wall_ms() {
  using namespace std::chrono;
  return duration_cast<milliseconds>(system_clock::now().time_since_epoch()).count();
}
const char* algo_scope(int32_t algorithm) { return algorithm == GUB_LEAKY_BUCKET ? "Error in leakyBucket" : "Error in tokenBucket"; }
}  // namespace

extern "C" {

void gub_format_error(int err_code, const char* key, int32_t algorithm, char* out, size_t cap) {
  // gubernator.go:252 wraps gubernator.go:600 wraps workers.go:304/313 wraps interval.go:93/107 (pkg/errors "a: b")
  static const char* pre = "Error while apply rate limit for '%s': during workerPool.GetRateLimit: ";
  char head[512];
  std::snprintf(head, sizeof head, pre, key);
  switch (err_code) {
    case GUB_ERR_UNIQUE_KEY_EMPTY: std::snprintf(out, cap, "field 'unique_key' cannot be empty"); break;
    case GUB_ERR_NAMESPACE_EMPTY: std::snprintf(out, cap, "field 'namespace' cannot be empty"); break;
    case GUB_ERR_INVALID_ALGORITHM: std::snprintf(out, cap, "%sInvalid rate limit algorithm '%d'", head, algorithm); break;
    case GUB_ERR_GREGORIAN_WEEKS:
      std::snprintf(out, cap, "%s%s: `Duration = GregorianWeeks` not yet supported; consider making a PR!`", head, algo_scope(algorithm));
      break;
    case GUB_ERR_GREGORIAN_INVALID:
      std::snprintf(out, cap, "%s%s: behavior DURATION_IS_GREGORIAN is set; but `Duration` is not a valid gregorian interval", head,
                    algo_scope(algorithm));
      break;
    case GUB_ERR_TABLE_FULL: std::snprintf(out, cap, "%srate limit table is full", head); break;
    default: if (cap) out[0] = 0;
  }
}

int gub_instance_create(gub_table* table, gub_instance** out) {
  if (!table || !out) return -1;
  *out = new gub_instance();
  (*out)->table = table;
  return 0;
}
void gub_instance_destroy(gub_instance* s) { delete s; }
void gub_instance_set_clock(gub_instance* s, int64_t frozen_now_ms) { s->frozen_now = frozen_now_ms; }
int64_t gub_instance_now(gub_instance* s) { return s->frozen_now >= 0 ? s->frozen_now : wall_ms(); }

int gub_instance_get_rate_limits_unbounded(gub_instance* s, const gub_rate_limit_req* reqs, size_t n, gub_rate_limit_resp* out) {
  if (!s || (n && (!reqs || !out))) return -1;
  const int64_t now = gub_instance_now(s);  // gubernator.go:195: one timestamp per call
  std::vector<gub_req> batch;
  std::vector<size_t> where;      // batch position -> request index
  std::vector<std::string> keys;  // for error text
  batch.reserve(n); where.reserve(n); keys.reserve(n);
  for (size_t i = 0; i < n; i++) {
    gub_rate_limit_resp& o = out[i];
    std::memset(&o, 0, sizeof o);
    const char* name = reqs[i].name ? reqs[i].name : "";
    const char* uk = reqs[i].unique_key ? reqs[i].unique_key : "";
    if (!uk[0]) {  // :208-212, checked before the namespace
      o.err_code = GUB_ERR_UNIQUE_KEY_EMPTY;
      gub_format_error(o.err_code, "", 0, o.error, sizeof o.error);
      continue;
    }
    if (!name[0]) {  // :213-217
      o.err_code = GUB_ERR_NAMESPACE_EMPTY;
      gub_format_error(o.err_code, "", 0, o.error, sizeof o.error);
      continue;
    }
    std::string key = std::string(name) + "_" + uk;  // client.go:39-41
    gub_req r;
    r.key_xxh64 = gub_xxh64(key.data(), key.size(), 0);
    r.key_fnv1 = gub_fnv1_64(key.data(), key.size());
    r.hits = reqs[i].hits; r.limit = reqs[i].limit; r.duration = reqs[i].duration; r.burst = reqs[i].burst;
    r.created_at = reqs[i].created_at != 0 ? reqs[i].created_at : now;  // :218-220
    r.algorithm = (uint32_t)reqs[i].algorithm;
    r.behavior = ((uint32_t)reqs[i].behavior & 0xFFu) | GUB_REQ_IS_OWNER;  // one-node cluster: every key is ours (:247-250)
    batch.push_back(r); where.push_back(i); keys.push_back(std::move(key));
  }
  if (batch.empty()) return 0;
  gub_clock clk;
  gub_clock_fill(now, &clk);
  std::vector<gub_resp> resp(batch.size());
  if (gub_submit(s->table, batch.data(), batch.size(), &clk, resp.data()) != 0) return -1;
  for (size_t j = 0; j < batch.size(); j++) {
    gub_rate_limit_resp& o = out[where[j]];
    o.err_code = (int32_t)resp[j].err_code;
    if (resp[j].err_code) {
      gub_format_error((int)resp[j].err_code, keys[j].c_str(), reqs[where[j]].algorithm, o.error, sizeof o.error);
    } else {
      o.status = (int32_t)resp[j].status; o.limit = resp[j].limit; o.remaining = resp[j].remaining; o.reset_time = resp[j].reset_time;
    }
  }
  return 0;
}

int gub_instance_get_rate_limits(gub_instance* s, const gub_rate_limit_req* reqs, size_t n, gub_rate_limit_resp* out) {
  if (n > GUB_MAX_BATCH_SIZE) return GUB_E_TOO_LARGE;  // gubernator.go:189-193
  return gub_instance_get_rate_limits_unbounded(s, reqs, n, out);
}

int gub_instance_update_peer_global(gub_instance* s, const char* key, int32_t algorithm, int64_t duration, int32_t status, int64_t limit,
                                    int64_t remaining, int64_t reset_time) {
  if (!s || !key) return -1;
  const int64_t now = gub_instance_now(s);  // gubernator.go:427
  gub_item it;
  std::memset(&it, 0, sizeof it);
  const size_t len = std::strlen(key);
  it.key_xxh64 = gub_xxh64(key, len, 0);
  it.key_fnv1 = gub_fnv1_64(key, len);
  it.algorithm = algorithm; it.expire_at = reset_time;  // :429-433
  it.limit = limit; it.duration = duration; it.stamp = now;
  if (algorithm == GUB_LEAKY_BUCKET) { it.remaining_f = (double)remaining; it.burst = limit; }  // :435-442
  else { it.status = status; it.remaining = remaining; }                                       // :443-450
  return gub_add_items(s->table, &it, 1);
}

}  // extern "C"
