/*
 * gubernator_b200_host.h — host-side mirror of the reference's service entry for the path, above the device C ABI.
 *
 * In production this layer is gubernator's own Go code (V1Instance in gubernator.go) calling the device ABI through
 * the cgo shim in go/workerpool_b200.go.  No Go toolchain exists in this build image, so the same logic is provided
 * here in C++ (the reference is compiled code) with the reference's names, argument meaning and error behaviour, so
 * that the parity tests can drive the CUDA path exactly like the reference's functional tests drive a daemon.
 */
#ifndef GUBERNATOR_B200_HOST_H
#define GUBERNATOR_B200_HOST_H

#include "gubernator_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* RateLimitReq (gubernator.proto:137-183).  created_at == 0 means "not set" (gubernator.go:218). */
typedef struct {
  const char* name;
  const char* unique_key;
  int64_t hits, limit, duration, burst;
  int32_t algorithm, behavior;
  int64_t created_at;
} gub_rate_limit_req;

/* RateLimitResp (gubernator.proto:190-203); `error` carries the exact reference string. */
typedef struct {
  int32_t status;
  int32_t err_code;
  int64_t limit, remaining, reset_time;
  char error[256];
} gub_rate_limit_resp;

#define GUB_MAX_BATCH_SIZE 1000 /* maxBatchSize, gubernator.go:40 */
#define GUB_E_TOO_LARGE (-2)    /* codes.OutOfRange "Requests.RateLimits list too large; max size is '1000'" (gubernator.go:189-193) */

typedef struct gub_instance gub_instance;

/* NewV1Instance (gubernator.go:115) over one device table (every key locally owned: a one-node cluster). */
int gub_instance_create(gub_table* table, gub_instance** out);
void gub_instance_destroy(gub_instance* s);
/* holster clock.Freeze(now)/Advance/Unfreeze: frozen_now_ms >= 0 freezes the instance clock, < 0 returns to wall time */
void gub_instance_set_clock(gub_instance* s, int64_t frozen_now_ms);
int64_t gub_instance_now(gub_instance* s);
/* V1Instance.GetRateLimits (gubernator.go:183-295): validation, HashKey, CreatedAt defaulting, one device batch,
 * responses in request order with in-band error strings.  Returns GUB_E_TOO_LARGE when n > 1000. */
int gub_instance_get_rate_limits(gub_instance* s, const gub_rate_limit_req* reqs, size_t n, gub_rate_limit_resp* out);
/* Same, without the 1000-item cap (what an RPC aggregator submits after coalescing many RPCs). */
int gub_instance_get_rate_limits_unbounded(gub_instance* s, const gub_rate_limit_req* reqs, size_t n, gub_rate_limit_resp* out);
/* V1Instance.UpdatePeerGlobals (gubernator.go:425-459) for one entry: builds the replica item and upserts it. */
int gub_instance_update_peer_global(gub_instance* s, const char* key, int32_t algorithm, int64_t duration, int32_t status,
                                    int64_t limit, int64_t remaining, int64_t reset_time);
/* ---- Store plugin (store.go:49-65) honoured at batch granularity -----------------------------------
 * The reference calls Store.Get on every cache miss, Store.OnChange after every owner-side mutation and Store.Remove when an
 * item is deleted — synchronously, per request (algorithms.go:45-51,149-153,250-254,81-83,98-100).  Behind a device batch
 * that becomes, per GetRateLimits call: one lookup of the call's distinct keys; Store.Get for each key the table does not hold
 * (the item it returns is installed before the batch runs, like c.Add at algorithms.go:49); the batch; then Store.Remove for
 * keys whose item was deleted (token RESET_REMAINING, or an algorithm switch) and ONE Store.OnChange per key with the item as
 * the batch left it (the reference would have called it once per request, with the intermediate states).
 * Keys are the reference's HashKey strings (Name + "_" + UniqueKey).  Callbacks run on the calling thread. */
typedef struct {
  void* user;
  int (*get)(void* user, const gub_rate_limit_req* req, const char* key, gub_item* item_out); /* 1 = found (item_out filled; key hashes are set by the caller) */
  void (*on_change)(void* user, const gub_rate_limit_req* req, const char* key, const gub_item* item);
  void (*remove)(void* user, const char* key);
} gub_store;
/* Config.Store (config.go:95).  NULL detaches.  The struct is copied. */
void gub_instance_set_store(gub_instance* s, const gub_store* store);

/* ---- RPC aggregator: concurrent GetRateLimits calls (<= 1000 requests each, gubernator.go:40) are
 * coalesced into one device batch, the way PeerClient.runBatch coalesces peer requests (peer_client.go:284-337): a flush
 * happens when `max_batch` requests are queued or `window_us` after the first queued call (BatchWait = 500 us,
 * config.go:128).  Calls are applied in arrival order, each call's requests in index order, so every caller sees exactly
 * what it would see if the calls had been served one after another.  Thread-safe; blocks the caller until its responses
 * are filled.  Returns GUB_E_TOO_LARGE for more than 1000 requests. */
typedef struct gub_aggregator gub_aggregator;
int gub_aggregator_create(gub_instance* s, uint32_t max_batch, uint32_t window_us, gub_aggregator** out);
void gub_aggregator_destroy(gub_aggregator* a);
int gub_aggregator_get_rate_limits(gub_aggregator* a, const gub_rate_limit_req* reqs, size_t n, gub_rate_limit_resp* out);
/* device batches flushed / requests carried so far */
void gub_aggregator_stats(gub_aggregator* a, uint64_t* batches, uint64_t* requests);

/* RateLimitResp.Error text for an in-band error code on `key` (workers.go:304-318, gubernator.go:252,600). */
void gub_format_error(int err_code, const char* key, int32_t algorithm, char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
