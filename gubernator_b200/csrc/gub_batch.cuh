// gub_batch.cuh — the rate-limit evaluation path as ONE persistent kernel per batch (sm_100a).
//
// Replaces, for a whole batch at a time, WorkerPool.GetRateLimit -> Worker.handleGetRateLimit -> LRUCache.GetItem ->
// tokenBucket/leakyBucket (workers.go:261-324, lrucache.go:111, algorithms.go:37-493 of mailgun/gubernator v2.4.0) with the
// results of applying the batch's requests one after another in index order (gubernator.go:203) under a frozen clock.
//
// One CTA per SM, 512 threads, one request per thread.  A CTA owns one TILE of consecutive requests per round:
//
//   phase 1  the tile's request records are staged into shared memory with one bulk-async copy (TMA, cp.async.bulk +
//            mbarrier); the tile is grouped by key in a shared-memory table with index-ordered local ranks (per-warp counts,
//            one barrier); the first member of each (tile, key) FRAGMENT joins the batch-wide group entry (64-bit CAS on the
//            key, count += members, fragments += 1), sets the tile's bit in the group's presence bitmap and stores the
//            fragment size; every key's home slot is prefetched into L2.
//   ---- one grid-wide barrier: every fragment of the batch is registered ----
//   phase 2  fragments read their group entry (total members, fragments); a fragment of a key that also occurs in other tiles
//            gets its base rank = sum of the earlier tiles' fragment sizes.  The table is probed WARP-COOPERATIVELY: four
//            lanes per fragment, one 16-byte load each = one coalesced 64-byte slot per quad, tag compare broadcast inside
//            the quad, linear probing until every quad of the warp is done; the slot lands in a shared-memory snapshot.
//            Every request then evaluates run_to_rank(snapshot, request, base + local rank) by itself (closed forms make that
//            O(1)) and stores its 32-byte response.  Nobody writes a slot other tiles may still read: a key confined to one
//            tile is written back by its last member; for a key spread over tiles every fragment checks in on the group
//            entry (atomicAdd) after reading, and the LAST fragment to arrive writes the final state (or, when the group's
//            requests differ, redoes the group in order, segment by segment) and hands the entry back clean.
//
// The request tile is read from HBM exactly once; there is one launch per batch instead of four; the only grid-wide
// synchronisation is the one barrier (plus one per extra round when a batch exceeds 148 x 512 requests).
#pragma once
#if !defined(GUB_EMULATE)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "gub_kernels.cuh"

namespace gub {

constexpr int FB_THREADS = 512;      // threads per CTA = most requests per tile
constexpr int FB_WARPS = FB_THREADS / 32;
constexpr int FB_HT = 1024;          // tile-local key table (load factor <= 0.5)
constexpr int FB_PRES_WORDS = 8;     // presence bitmap: one bit per tile of a round (grid <= 256)
constexpr int FB_MAX_GRID = FB_PRES_WORDS * 32;
constexpr int FB_ROW = FB_MAX_GRID;  // fragment-size row (uint16) per group entry
constexpr int FB_MAX_SEGS = MAX_SHARDS;
constexpr uint32_t FB_AUX_ENTRIES = 1u << 18;  // batch-wide group table (a round holds <= 256 x 512 keys)
constexpr uint32_t G_NONUNIFORM = 1u;
constexpr int FB_OVF_CAP = 1024;     // items whose insert found the probe window full; placed (with eviction) at the next batch

// Batch-wide group entry.  All zero between batches: whoever finishes a group hands the entry back clean.
struct __align__(32) GEntry {
  unsigned long long key;   // remapped XXH64 (0 = free), claimed with atomicCAS
  unsigned long long cnt;   // [63:32] fragments  [31:0] members
  uint32_t rep;             // round-local index of one member: every fragment's first member is compared with it
  uint32_t flags;           // G_NONUNIFORM
  uint32_t arrived;         // fragments that have read the slot and answered
  uint32_t _pad;
};

struct FSeg {                           // one run of request records, evaluated in order after the previous segment
  const gub_req* reqs;
  gub_resp* out;                        // response j of the segment -> out[j] (may be peer memory)
  const unsigned long long* flag;       // optional: (epoch << 32 | count), published by the producer of the segment
  const uint32_t* n_dev;                // optional: count on the device
  uint32_t n;                           // count when neither is given
  uint32_t _pad;
};

struct OvfItem { uint64_t key, tag; uint64_t w[6]; uint32_t flags, _pad; };  // 72 bytes

struct FCtl {
  uint32_t bar_cnt, bar_gen;            // grid barrier
  uint32_t ord_bump;                    // allocator of `ordbuf` (member lists of non-uniform groups), reset every round
  uint32_t done_ctr;                    // CTAs that have stored all their responses (multi-GPU: the last one publishes the flags)
  uint32_t error;                       // a flag wait timed out (a peer died)
  uint32_t ovf_count;                   // pending items in `ovf`
  unsigned long long sweep_cursor;      // next slot of the incremental expiry sweep
};

struct FArgs {
  Slot* table;
  uint64_t capacity;
  FSeg seg[FB_MAX_SEGS];
  uint32_t nseg;
  uint32_t flag_epoch;                  // epoch the segments' flags must show
  GEntry* aux;
  uint32_t* presence;                   // [FB_AUX_ENTRIES][FB_PRES_WORDS]
  uint16_t* fragsize;                   // [FB_AUX_ENTRIES][FB_ROW]
  uint32_t* gpos;                       // [FB_MAX_GRID * FB_THREADS] group entry of every request of the round
  uint32_t* ordbuf;                     // [FB_MAX_GRID * FB_THREADS]
  FCtl* ctl;
  OvfItem* ovf;                         // [FB_OVF_CAP]
  unsigned long long* counters;
  // multi-GPU: when every response of this launch has been stored, resp_flag[k] (k < n_resp_flags) receives flag_epoch << 32
  unsigned long long* resp_flag[FB_MAX_SEGS];
  uint32_t n_resp_flags;
  uint32_t sweep_chunk;                 // slots each CTA sweeps per round (0 = off)
  gub_clock clk;
};

// What the rarely taken, out-of-line parts of the kernel need of FArgs, kept in shared memory (a non-inlined function taking the
// kernel's parameter struct by reference would force a 1 KB local copy of it).  Same member names as FArgs.
struct FCtx {
  Slot* table;
  uint64_t capacity;
  uint32_t* gpos;
  FCtl* ctl;
  OvfItem* ovf;
  unsigned long long* counters;
  uint32_t nseg, sweep_chunk;
  gub_clock clk;
};

// ---- PTX: bulk-async copy (TMA) + mbarrier, acquire/release, grid barrier ---------------------------------------------
#if defined(GUB_EMULATE)
struct MBar { uint64_t w; };
__device__ __forceinline__ void mbar_init(MBar*) {}
__device__ __forceinline__ void tile_load(void* dst, const void* src, uint32_t bytes, MBar*) { std::memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(MBar*, uint32_t) {}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) { return *p; }
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { *p = v; }
__device__ __forceinline__ void spin_pause() { emu::spin_yield(); }
#else
struct __align__(8) MBar { uint64_t w; };
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(MBar* b) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(b)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// One thread: arm the barrier with the byte count and start the copy; the bytes land in shared memory through the async proxy.
__device__ __forceinline__ void tile_load(void* dst, const void* src, uint32_t bytes, MBar* b) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes),
               "r"(smem_addr(b))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(MBar* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_addr(b)), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void spin_pause() { __nanosleep(20); }
#endif

// Every CTA of the grid is resident (one per SM, cooperative launch), so a counter + generation barrier is safe.  The
// generation is read before arriving: it cannot advance until this CTA has arrived.
__device__ __forceinline__ void grid_barrier(FCtl* ctl) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t gen = ld_acquire_gpu(&ctl->bar_gen);
    __threadfence();
    if (atomicAdd(&ctl->bar_cnt, 1u) == gridDim.x - 1) {
      ctl->bar_cnt = 0;
      __threadfence();
      st_release_gpu(&ctl->bar_gen, gen + 1);
    } else {
      // bounded (seconds): a CTA that never arrives (a fault, a grid that is not co-resident) must not hang the device
      uint32_t it = 0;
      while (ld_acquire_gpu(&ctl->bar_gen) == gen && ++it < (1u << 26)) spin_pause();
      if (it >= (1u << 26)) atomicExch(&ctl->error, 2u);
    }
    __threadfence();
  }
  __syncthreads();
}

// ---- shared memory of one CTA --------------------------------------------------------------------------------------
struct FMixed {                      // scratch of the segment walk (non-uniform groups)
  uint32_t np, covered, nseg, serial;
  uint32_t wsum[FB_WARPS];
  uint16_t seg[FB_THREADS + 1];
  Piece pieces[MAX_PIECES];
};

struct __align__(128) FSmem {
  gub_req req[FB_THREADS];               // the tile (TMA destination); later: staging of a non-uniform group's requests
  ulonglong2 snap[FB_THREADS][4];        // per fragment: the slot as found; later: staging of a non-uniform group's responses
  unsigned long long key[FB_HT];         // tile-local key table
  uint8_t wcnt[FB_HT][FB_WARPS];         // [key slot][warp]: members of the key among the warp's lanes
  long long fslot[FB_THREADS];           // per fragment: slot index when found, else first reusable slot of the window (or -1)
  uint32_t fpos[FB_THREADS];             // per fragment: batch-wide group entry
  uint32_t fbase[FB_THREADS];            // per fragment: rank of its first member within the group
  uint32_t ftotal[FB_THREADS];           // per fragment: members of the whole group
  uint16_t fnfrag[FB_THREADS];           // per fragment: fragments of the whole group
  uint16_t flead[FB_THREADS];            // per fragment: its first member (thread / tile-local request index)
  uint16_t fcnt[FB_THREADS];             // per fragment: members in this tile
  uint16_t f_of_sp[FB_HT];               // key slot -> fragment
  uint16_t sp[FB_THREADS];               // per request: key slot
  uint16_t local[FB_THREADS];            // per request: rank within the fragment
  uint16_t fin[FB_THREADS];              // fragments whose group this CTA finishes
  uint8_t ffound[FB_THREADS];            // per fragment: key is in the table
  uint8_t fmixed[FB_THREADS];            // per fragment: its members differ (or can never settle): the group takes the segment walk
  FMixed mx;
  FCtx cx;
  // the batch's segments
  const gub_req* seg_reqs[FB_MAX_SEGS];
  gub_resp* seg_out[FB_MAX_SEGS];
  uint32_t seg_cnt[FB_MAX_SEGS];
  uint32_t seg_tile0[FB_MAX_SEGS + 1];
  uint32_t ts, ntiles, total;
  uint32_t nfrag, nfin;
  uint32_t tally[8];
  MBar mbar;
};

struct TileInfo { uint32_t seg, off, n; };
__device__ __forceinline__ TileInfo tile_info(const FSmem& S, uint32_t nseg, uint32_t tl) {
  TileInfo ti;
  uint32_t s = 0;
  while (s + 1 < nseg && tl >= S.seg_tile0[s + 1]) s++;
  ti.seg = s;
  ti.off = (tl - S.seg_tile0[s]) * S.ts;
  ti.n = min(S.ts, S.seg_cnt[s] - ti.off);
  return ti;
}
// Round-local request index i = tile-in-round * FB_THREADS + position in the tile.
__device__ __forceinline__ const gub_req* req_at(const FSmem& S, uint32_t nseg, uint32_t round, uint32_t i) {
  const TileInfo ti = tile_info(S, nseg, round * gridDim.x + i / FB_THREADS);
  return S.seg_reqs[ti.seg] + ti.off + (i % FB_THREADS);
}
__device__ __forceinline__ gub_resp* resp_at(const FSmem& S, uint32_t nseg, uint32_t round, uint32_t i) {
  const TileInfo ti = tile_info(S, nseg, round * gridDim.x + i / FB_THREADS);
  return S.seg_out[ti.seg] + ti.off + (i % FB_THREADS);
}

__device__ __forceinline__ gub_req smem_req(const gub_req* p) {
  const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
  const ulonglong2 a = q[0], b = q[1], c = q[2], d = q[3];
  gub_req r;
  r.key_xxh64 = a.x; r.key_fnv1 = a.y; r.hits = (int64_t)b.x; r.limit = (int64_t)b.y; r.duration = (int64_t)c.x;
  r.burst = (int64_t)c.y; r.created_at = (int64_t)d.x; r.algorithm = (uint32_t)(d.y & 0xFFFFFFFFull); r.behavior = (uint32_t)(d.y >> 32);
  return r;
}
__device__ __forceinline__ gub_req global_req(const gub_req* p) {  // written earlier in this launch's lifetime by other devices / kernels: L2
  const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
  const ulonglong2 a = __ldcg(q), b = __ldcg(q + 1), c = __ldcg(q + 2), d = __ldcg(q + 3);
  gub_req r;
  r.key_xxh64 = a.x; r.key_fnv1 = a.y; r.hits = (int64_t)b.x; r.limit = (int64_t)b.y; r.duration = (int64_t)c.x;
  r.burst = (int64_t)c.y; r.created_at = (int64_t)d.x; r.algorithm = (uint32_t)(d.y & 0xFFFFFFFFull); r.behavior = (uint32_t)(d.y >> 32);
  return r;
}

// ---- the batch-wide group table ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gentry_join(const FArgs& A, uint64_t key, bool* claimed) {
  uint32_t pos = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 40) & (FB_AUX_ENTRIES - 1);
  *claimed = false;
#pragma unroll 1
  for (;;) {
    const unsigned long long old = atomicCAS(&A.aux[pos].key, 0ull, (unsigned long long)key);
    if (old == 0ull) { *claimed = true; break; }
    if (old == key) break;
    pos = (pos + 1) & (FB_AUX_ENTRIES - 1);
  }
  return pos;
}
__device__ __forceinline__ void gentry_clear(GEntry* e) {
  ulonglong2* p = reinterpret_cast<ulonglong2*>(e);
  __stcg(p, make_ulonglong2(0ull, 0ull));
  __stcg(p + 1, make_ulonglong2(0ull, 0ull));
}

// Sum of the fragment sizes of the tiles before `tt` that hold members of group `pos`.
__device__ __forceinline__ uint32_t fragment_base2(const FArgs& A, uint32_t pos, uint32_t tt) {
  const uint4* pres = reinterpret_cast<const uint4*>(A.presence + (size_t)pos * FB_PRES_WORDS);
  const uint16_t* row = A.fragsize + (size_t)pos * FB_ROW;
  const uint4 p0 = __ldcg(pres), p1 = __ldcg(pres + 1);
  uint32_t bits[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
  const uint32_t last = tt >> 5, keep = (1u << (tt & 31)) - 1u;
  uint32_t base = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    uint32_t x = ((uint32_t)w < last) ? bits[w] : ((uint32_t)w == last ? (bits[w] & keep) : 0u);
    if (!x) continue;
    if (__popc(x) <= 4) {
      while (x) { const uint32_t k = __ffs(x) - 1; x &= x - 1; base += (uint32_t)__ldcg(row + w * 32 + k); }
    } else {  // all 32 sizes of the word's tiles (64 bytes), masked
      const uint4* r4 = reinterpret_cast<const uint4*>(row + w * 32);
      const uint4 v0 = __ldcg(r4), v1 = __ldcg(r4 + 1), v2 = __ldcg(r4 + 2), v3 = __ldcg(r4 + 3);
      const uint32_t v[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const uint32_t m = (x >> (2 * k)) & 3u;
        base += ((m & 1u) ? (v[k] & 0xFFFFu) : 0u) + ((m & 2u) ? (v[k] >> 16) : 0u);
      }
    }
  }
  return base;
}

// ---- warp-cooperative table probe ---------------------------------------------------------------------------------------
// Fragments [0, F) of the tile are looked up four lanes at a time: lane q of a quad loads bytes [16q, 16q+16) of the probed
// slot (one coalesced 64-byte transaction per quad, eight slots per warp instruction), lane 0's words (key, tag|flags) are
// broadcast inside the quad, and the quads of a warp probe linearly until every one of them has hit, or met an empty slot.
__device__ __forceinline__ void probe_fragments(const FArgs& A, FSmem& S, uint32_t F) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane & 3, quad0 = lane & ~3u;
#pragma unroll 1
  for (uint32_t j0 = warp * 8; j0 < F; j0 += FB_WARPS * 8) {
    const uint32_t j = j0 + (lane >> 2);
    bool done = j >= F;
    uint64_t key = 0, tag = 0, idx = 0;
    long long reuse = -1;
    bool found = false;
    if (!done) {
      const gub_req* lr = &S.req[S.flead[j]];
      key = remap_key(lr->key_xxh64); tag = lr->key_fnv1 >> 8;
      idx = __umul64hi(key, A.capacity);
    }
#pragma unroll 1
    for (int p = 0; p < MAX_PROBE; p++) {
      ulonglong2 v = make_ulonglong2(0ull, 0ull);
      if (!done) v = __ldcg(reinterpret_cast<const ulonglong2*>(A.table + idx) + q);
      const unsigned long long w0 = __shfl_sync(0xFFFFFFFFu, v.x, quad0), w1 = __shfl_sync(0xFFFFFFFFu, v.y, quad0);
      if (!done) {
        if (w0 == key && (w1 >> 8) == tag) {
          S.snap[j][q] = v;
          found = true; done = true;
        } else if (w0 == KEY_EMPTY) {
          if (reuse < 0) reuse = (long long)idx;
          done = true;
        } else {
          if (w0 == KEY_TOMB && reuse < 0) reuse = (long long)idx;
          idx = (idx + 1 == A.capacity) ? 0 : idx + 1;
        }
      }
      if (__all_sync(0xFFFFFFFFu, done)) break;
    }
    if (j < F && q == 0) {
      S.fslot[j] = found ? (long long)idx : reuse;
      S.ffound[j] = found ? 1 : 0;
    }
  }
}

__device__ __forceinline__ void cursor_from_snapshot(const FArgs& A, const FSmem& S, uint32_t f, uint64_t key, uint64_t tag, Cursor& c) {
  c.home = __umul64hi(key, A.capacity);
  c.found = S.ffound[f] != 0;
  c.slot = S.fslot[f];
  if (c.found) {
    bucket_from(c.b, S.snap[f][0], S.snap[f][1], S.snap[f][2], S.snap[f][3]);
  } else {
    c.b.key = key; c.b.tag = tag; c.b.flags = 0; c.b.limit = 0; c.b.duration = 0; c.b.rem = 0; c.b.stamp = 0; c.b.burst = 0; c.b.expire = 0;
  }
  c.old = c.b;
}

// Writes a key's final state.  A new key whose probe window has no free slot is parked in the overflow list: the next batch
// places it before anything reads the table, evicting the entry of the window that expires first (the reference's LRU would
// have evicted as well: lrucache.go:98,138-149).
template <class Ctx>
__device__ __forceinline__ void close_or_park(const Ctx& A, Cursor& cur, Tally& t) {
  if (cursor_close(cur, A.table, A.capacity, t.inserts)) return;
  const uint32_t k = atomicAdd(&A.ctl->ovf_count, 1u);
  if (k < (uint32_t)FB_OVF_CAP) {
    OvfItem it;
    it.key = cur.b.key; it.tag = cur.b.tag; it.flags = cur.b.flags; it._pad = 0;
    it.w[0] = (uint64_t)cur.b.limit; it.w[1] = (uint64_t)cur.b.duration; it.w[2] = cur.b.rem; it.w[3] = (uint64_t)cur.b.stamp;
    it.w[4] = (uint64_t)cur.b.burst; it.w[5] = (uint64_t)cur.b.expire;
    A.ovf[k] = it;
  } else {
    t.full++;  // more than FB_OVF_CAP keys without a slot in one batch: the state of this one is dropped (counted)
  }
}

// ---- groups whose requests differ: walked in index order, segment by segment ---------------------------------------------
// ord[0..cnt) = round-local indices of the group's members in index order.  The group is taken in chunks of FB_THREADS members:
// the chunk's requests are staged in shared memory, runs of identical requests (segments) are found in parallel, and then either
// thread 0 plans every segment with plan_run() (closed forms) and all threads evaluate and store the responses, or — when the
// chunk is mostly one-request segments, where planning buys nothing — thread 0 simply applies the chunk's requests one after
// another.  The slot is opened once and written once.
__device__ __noinline__ void mixed_walk(const FCtx& A, FSmem& S, const uint32_t* ord, uint32_t cnt, uint32_t round, Tally& t) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  gub_resp* stage = reinterpret_cast<gub_resp*>(&S.snap[0][0]);
  Cursor cur;
  bool open = false;
  uint64_t ck = 0, ct = 0;
#pragma unroll 1
  for (uint32_t c0 = 0; c0 < cnt; c0 += FB_THREADS) {
    const uint32_t m = min((uint32_t)FB_THREADS, cnt - c0);
    __syncthreads();  // the staging areas are free again
    uint32_t my_i = 0;
    if (tid < m) {
      my_i = __ldcg(ord + c0 + tid);
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(req_at(S, A.nseg, round, my_i));
      ulonglong2* dst = reinterpret_cast<ulonglong2*>(&S.req[tid]);
      dst[0] = __ldcg(src); dst[1] = __ldcg(src + 1); dst[2] = __ldcg(src + 2); dst[3] = __ldcg(src + 3);
    }
    __syncthreads();
    // segment starts, in order
    bool boundary = false;
    if (tid < m) boundary = tid == 0 || !req_same(smem_req(&S.req[tid]), smem_req(&S.req[tid - 1]));
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, boundary);
    if (lane == 0) S.mx.wsum[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = 0, nseg = 0;
#pragma unroll
    for (int w = 0; w < FB_WARPS; w++) { const uint32_t c = S.mx.wsum[w]; nseg += c; before += ((uint32_t)w < warp) ? c : 0u; }
    if (boundary) S.mx.seg[before + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)tid;
    if (tid == 0) S.mx.seg[nseg] = (uint16_t)m;
    __syncthreads();
    const bool serial = nseg * 4 > m && m > 8;  // short segments: planning costs more than applying
    if (serial) {
      if (tid == 0) {
#pragma unroll 1
        for (uint32_t k = 0; k < m; k++) {
          const gub_req rq = smem_req(&S.req[k]);
          const uint64_t key = remap_key(rq.key_xxh64), tag = rq.key_fnv1 >> 8;
          if (!open || key != ck || tag != ct) {
            if (open) close_or_park(A, cur, t);
            cursor_open(cur, A.table, A.capacity, key, tag);
            open = true; ck = key; ct = tag;
          }
          Delta d = {0, 0, 0};
          const gub_resp r = apply_one(cur.b, rq, A.clk, d);
          t.over += d.over; t.hit += d.hit; t.miss += d.miss;
          stage[k] = r;
        }
        atomicAdd(A.counters + C_SERIAL, 1ull);
      }
      __syncthreads();
      if (tid < m) store_resp(resp_at(S, A.nseg, round, my_i), stage[tid]);
      continue;
    }
#pragma unroll 1
    for (uint32_t s = 0; s < nseg; s++) {
      const uint32_t lo = S.mx.seg[s], hi = S.mx.seg[s + 1], len = hi - lo;
      if (tid == 0) {
        const gub_req rq = smem_req(&S.req[lo]);
        const uint64_t key = remap_key(rq.key_xxh64), tag = rq.key_fnv1 >> 8;
        if (!open || key != ck || tag != ct) {
          if (open) close_or_park(A, cur, t);
          cursor_open(cur, A.table, A.capacity, key, tag);
          open = true; ck = key; ct = tag;
        }
        Delta d = {0, 0, 0};
        uint32_t np = 0;
        const uint32_t covered = plan_run(cur.b, rq, len, A.clk, d, S.mx.pieces, MAX_PIECES, &np);
        t.over += d.over; t.hit += d.hit; t.miss += d.miss;
        // ranks the piece buffer could not hold (no regular regime, e.g. RESET_REMAINING flip-flops): applied one by one
        for (uint32_t k = covered; k < len; k++) {
          Delta d2 = {0, 0, 0};
          stage[lo + k] = apply_one(cur.b, rq, A.clk, d2);
          t.over += d2.over; t.hit += d2.hit; t.miss += d2.miss;
        }
        S.mx.np = np; S.mx.covered = covered;
      }
      __syncthreads();
      const uint32_t np = S.mx.np, covered = S.mx.covered;
      for (uint32_t k = tid; k < covered; k += FB_THREADS) {
        uint32_t pi = 0;
        while (pi + 1 < np && S.mx.pieces[pi + 1].start <= k) pi++;
        stage[lo + k] = eval_piece(S.mx.pieces[pi], k);
      }
      __syncthreads();
    }
    if (tid < m) store_resp(resp_at(S, A.nseg, round, my_i), stage[tid]);
  }
  if (tid == 0 && open) close_or_park(A, cur, t);
  __syncthreads();
}

// Members of group `pos` in the tiles `bits` marks (round-local tile numbers), listed in index order into `ord`.
// Returns the number listed (== the group's member count).
__device__ __noinline__ uint32_t list_members(const FCtx& A, FSmem& S, uint32_t pos, const uint32_t bits[FB_PRES_WORDS], uint32_t round, uint32_t* ord) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t listed = 0;
#pragma unroll 1
  for (uint32_t w = 0; w < (uint32_t)FB_PRES_WORDS; w++) {
    uint32_t x = bits[w];
#pragma unroll 1
    while (x) {
      const uint32_t tt = w * 32 + (__ffs(x) - 1);
      x &= x - 1;
      const TileInfo ti = tile_info(S, A.nseg, round * gridDim.x + tt);
      const bool mine = tid < ti.n && __ldcg(A.gpos + tt * FB_THREADS + tid) == pos;
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mine);
      __syncthreads();
      if (lane == 0) S.mx.wsum[warp] = __popc(bal);
      __syncthreads();
      uint32_t before = 0, tot = 0;
#pragma unroll
      for (int k = 0; k < FB_WARPS; k++) { const uint32_t c = S.mx.wsum[k]; tot += c; before += ((uint32_t)k < warp) ? c : 0u; }
      if (mine) __stcg(ord + listed + before + __popc(bal & ((1u << lane) - 1u)), tt * FB_THREADS + tid);
      listed += tot;
    }
  }
  __syncthreads();
  return listed;
}

// ---- maintenance inside the batch kernel (both run before the grid barrier, when nothing reads the table) -------------------
// Items parked by close_or_park(): one warp places them, evicting when the window is still full.
__device__ __noinline__ void drain_overflow(const FCtx& A, Tally& t) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n = min(__ldcg(&A.ctl->ovf_count), (uint32_t)FB_OVF_CAP);
#pragma unroll 1
  for (uint32_t k = 0; k < n; k++) {
    const OvfItem it = A.ovf[k];
    const uint64_t home = __umul64hi(it.key, A.capacity);
    // every lane inspects 16 slots of the window: best = reusable, else dead / expired, else the smallest ExpireAt
    uint64_t best_rank = ~0ull, best_idx = 0;
#pragma unroll 1
    for (uint32_t p = lane; p < (uint32_t)MAX_PROBE; p += 32) {
      const uint64_t idx = (home + p) % A.capacity;  // (a table smaller than the window wraps more than once)
      const ulonglong2 a = __ldcg(reinterpret_cast<const ulonglong2*>(A.table + idx));
      const int64_t exp = (int64_t)__ldcg(&A.table[idx].w[7]);
      uint64_t rank;
      if (a.x == it.key && (a.y >> 8) == it.tag) rank = 0;                                   // the key itself (re-created meanwhile)
      else if (a.x <= KEY_TOMB) rank = 1;                                                    // free
      else if (!(a.y & F_LIVE) || exp < A.clk.now_ms) rank = 2;                              // removed or expired
      else rank = 3 + ((uint64_t)exp ^ 0x8000000000000000ull) / 4;                           // live: earliest ExpireAt first
      rank = (rank << 9 | (uint64_t)p) & ~0ull;                                              // ties: nearest to home
      if (rank < best_rank) { best_rank = rank; best_idx = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t r2 = __shfl_sync(0xFFFFFFFFu, (unsigned long long)best_rank, lane ^ o), i2 = __shfl_sync(0xFFFFFFFFu, (unsigned long long)best_idx, lane ^ o);
      if (r2 < best_rank) { best_rank = r2; best_idx = i2; }
    }
    if (lane == 0) {
      if ((best_rank >> 9) >= 3) atomicAdd(A.counters + C_EVICT_UNEXPIRED, 1ull);
      if ((best_rank >> 9) >= 1) t.inserts++;
      ulonglong2* p = reinterpret_cast<ulonglong2*>(A.table + best_idx);
      __stcg(p, make_ulonglong2(it.key, (it.tag << 8) | (uint64_t)(it.flags & 0xFF)));
      __stcg(p + 1, make_ulonglong2(it.w[0], it.w[1]));
      __stcg(p + 2, make_ulonglong2(it.w[2], it.w[3]));
      __stcg(p + 3, make_ulonglong2(it.w[4], it.w[5]));
    }
    __syncwarp();
  }
  if (lane == 0 && n) A.ctl->ovf_count = 0;
}

// Incremental expiry sweep: every CTA frees the removed / expired entries of a few slots per round (tombstones; a tombstone
// run that ends at an empty slot becomes empty again), so that a long-running service does not fill its probe windows with
// dead keys.  The reference frees them lazily on access (lrucache.go:115) and by LRU eviction (lrucache.go:138).
__device__ __noinline__ void sweep_slice(const FCtx& A, uint32_t* swept) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t start = (__ldcg(&A.ctl->sweep_cursor) + (uint64_t)blockIdx.x * A.sweep_chunk) % A.capacity;
#pragma unroll 1
  for (uint32_t c0 = 0; c0 < A.sweep_chunk; c0 += 32) {
    const uint32_t k = c0 + lane;
    const bool in = k < A.sweep_chunk;
    const uint64_t idx = (start + k) % A.capacity;
    uint32_t cls = 2;  // 0 empty, 1 tombstone (or about to be), 2 live
    if (in) {
      const ulonglong2 a = __ldcs(reinterpret_cast<const ulonglong2*>(A.table + idx));
      if (a.x == KEY_EMPTY) cls = 0;
      else if (a.x == KEY_TOMB) cls = 1;
      else if (!(a.y & F_LIVE) || (int64_t)__ldcs(&A.table[idx].w[7]) < A.clk.now_ms) { cls = 1; A.table[idx].w[0] = KEY_TOMB; (*swept)++; }
    }
    // the slot after the warp's 32: decides whether a trailing tombstone run may become empty
    uint64_t nxt = start + c0 + 32; nxt %= A.capacity;
    const unsigned long long after = __ldcs(&A.table[nxt].w[0]);
    const uint64_t E = (uint64_t)__ballot_sync(0xFFFFFFFFu, in && cls == 0) | ((after == KEY_EMPTY && c0 + 32 <= A.sweep_chunk) ? (1ull << 32) : 0ull);
    const uint64_t T = (uint64_t)__ballot_sync(0xFFFFFFFFu, in && cls == 1);
    if (in && cls == 1) {
      const uint64_t notT = ~T >> lane;                 // first slot at or above mine that is not a tombstone
      const uint32_t m = lane + (uint32_t)__ffsll((long long)notT) - 1;
      if (m <= 32 && ((E >> m) & 1ull)) A.table[idx].w[0] = KEY_EMPTY;
    }
  }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------
#if defined(GUB_EMULATE)
#define GUB_DYN_SMEM() (emu::dyn_smem())
#else
extern __shared__ __align__(128) unsigned char gub_dyn_smem[];
#define GUB_DYN_SMEM() (gub_dyn_smem)
#endif

__device__ __forceinline__ uint32_t segment_count(const FArgs& A, uint32_t s) {
  const FSeg& g = A.seg[s];
  if (g.flag) {
    for (uint32_t it = 0; it < 20000000u; it++) {
      const unsigned long long v = ld_acquire_sys(g.flag);
      if ((uint32_t)(v >> 32) == A.flag_epoch) return (uint32_t)(v & 0xFFFFFFFFull);
      __nanosleep(100);
    }
    atomicExch(&A.ctl->error, 1u);  // a peer died: its segment counts as empty; the host reads the flag after the step
    return 0;
  }
  if (g.n_dev) return min(__ldcg(g.n_dev), g.n);
  return g.n;
}

__global__ void __launch_bounds__(FB_THREADS, 1) k_batch(const FArgs A) {
  FSmem& S = *reinterpret_cast<FSmem*>(GUB_DYN_SMEM());
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&S.mbar); S.nfrag = 0; S.nfin = 0;
    S.cx.table = A.table; S.cx.capacity = A.capacity; S.cx.gpos = A.gpos; S.cx.ctl = A.ctl; S.cx.ovf = A.ovf; S.cx.counters = A.counters;
    S.cx.nseg = A.nseg; S.cx.sweep_chunk = A.sweep_chunk; S.cx.clk = A.clk;
  }
  if (tid < 8) S.tally[tid] = 0;
  pdl_wait();     // everything earlier in the stream (the producer of the records, the previous batch) is complete and visible
  pdl_release();
  if (tid < A.nseg) {
    S.seg_cnt[tid] = segment_count(A, tid);
    S.seg_reqs[tid] = A.seg[tid].reqs;
    S.seg_out[tid] = A.seg[tid].out;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t total = 0;
    for (uint32_t s = 0; s < A.nseg; s++) total += S.seg_cnt[s];
    // tile size: spread the batch over the grid (segments end with a partial tile each)
    const uint32_t usable = gridDim.x > A.nseg ? gridDim.x - A.nseg : 1u;
    uint32_t ts = (total + usable - 1) / usable;
    ts = (ts + 31u) & ~31u;
    if (ts < 32u) ts = 32u;
    if (ts > (uint32_t)FB_THREADS) ts = FB_THREADS;
    uint32_t tiles = 0;
    for (uint32_t s = 0; s < A.nseg; s++) { S.seg_tile0[s] = tiles; tiles += (S.seg_cnt[s] + ts - 1) / ts; }
    S.seg_tile0[A.nseg] = tiles;
    S.ts = ts; S.ntiles = tiles; S.total = total;
  }
  __syncthreads();
  const uint32_t ntiles = S.ntiles;
  const uint32_t rounds = (ntiles + gridDim.x - 1) / gridDim.x;
  Tally t = {0, 0, 0, 0, 0};
  uint32_t dup = 0, mixed_groups = 0, swept = 0;
  uint32_t parity = 0;
  if (blockIdx.x == 0 && tid == 0) {
    atomicAdd(A.counters + C_REQUESTS, (unsigned long long)S.total);
    atomicAdd(A.counters + C_BATCHES, 1ull);
  }

#pragma unroll 1
  for (uint32_t round = 0; round < max(rounds, 1u); round++) {
    // =============================== phase 1: stage, group, register ===============================
    const uint32_t tl = round * gridDim.x + blockIdx.x;
    TileInfo ti = {0, 0, 0};
    if (tl < ntiles) ti = tile_info(S, A.nseg, tl);
    const uint32_t n_t = ti.n;
    if (tid == 0 && n_t) tile_load(&S.req[0], S.seg_reqs[ti.seg] + ti.off, n_t * (uint32_t)sizeof(gub_req), &S.mbar);
    {
      ulonglong2* kz = reinterpret_cast<ulonglong2*>(&S.key[0]);  // 8 KB
      kz[tid] = make_ulonglong2(0ull, 0ull);
      ulonglong2* wz = reinterpret_cast<ulonglong2*>(&S.wcnt[0][0]);  // 16 KB
      wz[tid] = make_ulonglong2(0ull, 0ull);
      wz[tid + FB_THREADS] = make_ulonglong2(0ull, 0ull);
      S.fmixed[tid] = 0;
    }
    if (blockIdx.x == 0 && tid == 0) A.ctl->ord_bump = 0;
    __syncthreads();
    if (n_t) { mbar_wait(&S.mbar, parity); parity ^= 1u; }
    const bool valid = tid < n_t;
    uint64_t key = 0;
    uint32_t sp = 0xFFFFu;
    if (valid) {
      key = remap_key(S.req[tid].key_xxh64);
      const uint64_t home = __umul64hi(key, A.capacity);
      prefetch_l2(A.table + home);
      prefetch_l2(A.table + (home + 1 == A.capacity ? 0 : home + 1));
      sp = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 54);  // top 10 bits -> FB_HT
#pragma unroll 1
      for (;;) {
        const unsigned long long old = atomicCAS(&S.key[sp], 0ull, (unsigned long long)key);
        if (old == 0ull || old == key) break;
        sp = (sp + 1) & (FB_HT - 1);
      }
    }
    __syncthreads();
    // index-ordered rank inside the tile: members in earlier warps + earlier lanes of my warp
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, valid ? sp : (0x10000u | lane));
    if (valid && lane == (uint32_t)(__ffs(peers) - 1)) S.wcnt[sp][warp] = (uint8_t)__popc(peers);
    __syncthreads();
    uint32_t local = 0;
    if (valid) {
      const uint4 wc = *reinterpret_cast<const uint4*>(&S.wcnt[sp][0]);
      const uint32_t wv[4] = {wc.x, wc.y, wc.z, wc.w};
      uint32_t before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < FB_WARPS; w++) {
        const uint32_t c = (wv[w >> 2] >> (8 * (w & 3))) & 0xFFu;
        total += c;
        before += ((uint32_t)w < warp) ? c : 0u;
      }
      local = before + __popc(peers & ((1u << lane) - 1u));
      S.sp[tid] = (uint16_t)sp; S.local[tid] = (uint16_t)local;
      if (local == 0) {  // the fragment's first member registers it
        const uint32_t f = atomicAdd(&S.nfrag, 1u);
        S.f_of_sp[sp] = (uint16_t)f; S.flead[f] = (uint16_t)tid; S.fcnt[f] = (uint16_t)total;
        bool claimed;
        const uint32_t pos = gentry_join(A, key, &claimed);
        atomicAdd(&A.aux[pos].cnt, (1ull << 32) | (unsigned long long)total);
        if (claimed) A.aux[pos].rep = blockIdx.x * FB_THREADS + tid;
        atomicOr(&A.presence[(size_t)pos * FB_PRES_WORDS + (blockIdx.x >> 5)], 1u << (blockIdx.x & 31));
        A.fragsize[(size_t)pos * FB_ROW + blockIdx.x] = (uint16_t)total;
        S.fpos[f] = pos;
      }
    }
    __syncthreads();
    if (valid) A.gpos[blockIdx.x * FB_THREADS + tid] = S.fpos[S.f_of_sp[sp]];
    // maintenance: nothing reads the table before the barrier
    if (blockIdx.x == 0 && warp == 0 && __ldcg(&A.ctl->ovf_count)) drain_overflow(S.cx, t);
    if (A.sweep_chunk && warp == FB_WARPS - 1 && !__ldcg(&A.ctl->ovf_count)) sweep_slice(S.cx, &swept);  // (a pending placement may pick a slot the sweep is freeing)

    grid_barrier(A.ctl);

    // =============================== phase 2: probe, evaluate, finish ===============================
    if (A.sweep_chunk && blockIdx.x == 0 && tid == 0) A.ctl->sweep_cursor = (A.ctl->sweep_cursor + (uint64_t)gridDim.x * A.sweep_chunk) % A.capacity;
    const uint32_t F = S.nfrag;
    uint32_t e_pos = 0, e_total = 0, e_nfrag = 0, e_rep = 0;
    if (tid < F) {  // the fragment's group entry: issued ahead of the probe, consumed after it
      e_pos = S.fpos[tid];
      const ulonglong2 e0 = __ldcg(reinterpret_cast<const ulonglong2*>(&A.aux[e_pos]));
      e_rep = __ldcg(&A.aux[e_pos].rep);
      e_total = (uint32_t)(e0.y & 0xFFFFFFFFull); e_nfrag = (uint32_t)(e0.y >> 32);
    }
    probe_fragments(A, S, F);
    if (tid < F) {
      uint32_t base = 0;
      if (e_nfrag > 1) {
        base = fragment_base2(A, e_pos, blockIdx.x);
        const uint32_t lead = S.flead[tid];
        if (e_rep != blockIdx.x * FB_THREADS + lead) {  // uniformity across tiles: every fragment's first member == the representative
          const gub_req rr = global_req(req_at(S, A.nseg, round, e_rep));
          if (!req_same(smem_req(&S.req[lead]), rr)) atomicOr(&A.aux[e_pos].flags, G_NONUNIFORM);
        }
      } else {  // the key lives in this tile only: hand the entry and the bitmap word back now
        A.presence[(size_t)e_pos * FB_PRES_WORDS + (blockIdx.x >> 5)] = 0;
        gentry_clear(&A.aux[e_pos]);
      }
      S.fbase[tid] = base; S.ftotal[tid] = e_total; S.fnfrag[tid] = (uint16_t)e_nfrag;
    }
    __syncthreads();
    // uniformity inside the fragment: every member == the fragment's first
    gub_req rq;
    uint32_t f = 0, lead = 0;
    if (valid) {
      rq = smem_req(&S.req[tid]);
      f = S.f_of_sp[sp]; lead = S.flead[f];
      if (S.ftotal[f] > 1) {
        bool irregular = !req_regular(rq);
        if (tid != lead && !irregular) irregular = !req_same(rq, smem_req(&S.req[lead]));
        if (irregular) S.fmixed[f] = 1;
      }
    }
    __syncthreads();
    if (valid) {
      const uint32_t nfrag = S.fnfrag[f];
      if (!S.fmixed[f]) {
        Cursor cur;
        cursor_from_snapshot(A, S, f, key, rq.key_fnv1 >> 8, cur);
        Delta d = {0, 0, 0};
        const gub_resp r = run_to_rank(cur.b, rq, S.fbase[f] + local, A.clk, d);
        store_resp(S.seg_out[ti.seg] + ti.off + tid, r);
        if (nfrag == 1 && local + 1 == S.fcnt[f]) {  // the key's last request of the batch: I hold its final state and counter totals
          t.over += d.over; t.hit += d.hit; t.miss += d.miss;
          close_or_park(A, cur, t);
          if (S.fcnt[f] > 1) dup++;
        }
      } else if (nfrag > 1 && tid == lead) {
        atomicOr(&A.aux[S.fpos[f]].flags, G_NONUNIFORM);
      }
    }
    __threadfence_system();  // responses (possibly in peer memory) before the check-in below
    __syncthreads();
    // check in: the last fragment of a group to arrive finishes it
    if (tid < F) {
      const uint32_t nfrag = S.fnfrag[tid];
      bool finish = false;
      if (nfrag > 1) {
        finish = atomicAdd(&A.aux[S.fpos[tid]].arrived, 1u) == nfrag - 1;
        if (finish) __threadfence();
      } else {
        finish = S.fmixed[tid] != 0;
      }
      if (finish) S.fin[atomicAdd(&S.nfin, 1u)] = (uint16_t)tid;
    }
    __syncthreads();
    const uint32_t nfin = S.nfin;
    // (a) uniform groups spread over tiles: one thread each computes the run's final state from its own snapshot and writes it
    for (uint32_t k = tid; k < nfin; k += FB_THREADS) {
      const uint32_t ff = S.fin[k];
      if (S.fnfrag[ff] <= 1) continue;
      const uint32_t pos = S.fpos[ff];
      if (__ldcg(&A.aux[pos].flags) & G_NONUNIFORM) continue;
      const gub_req lr = smem_req(&S.req[S.flead[ff]]);
      Cursor cur;
      cursor_from_snapshot(A, S, ff, remap_key(lr.key_xxh64), lr.key_fnv1 >> 8, cur);
      Delta d = {0, 0, 0};
      run_to_rank(cur.b, lr, S.ftotal[ff] - 1, A.clk, d);
      t.over += d.over; t.hit += d.hit; t.miss += d.miss;
      close_or_park(A, cur, t);
      dup++;
      ulonglong2* pz = reinterpret_cast<ulonglong2*>(A.presence + (size_t)pos * FB_PRES_WORDS);
      __stcg(pz, make_ulonglong2(0ull, 0ull)); __stcg(pz + 1, make_ulonglong2(0ull, 0ull));
      gentry_clear(&A.aux[pos]);
      S.fin[k] = 0xFFFFu;
    }
    __syncthreads();
    // (b) groups whose requests differ: the whole CTA walks each in index order
#pragma unroll 1
    for (uint32_t k = 0; k < nfin; k++) {
      const uint32_t ff = S.fin[k];
      if (ff == 0xFFFFu) continue;
      const uint32_t pos = S.fpos[ff], total = S.ftotal[ff];
      const bool spread = S.fnfrag[ff] > 1;
      uint32_t bits[FB_PRES_WORDS];
#pragma unroll
      for (int w = 0; w < FB_PRES_WORDS; w++) bits[w] = 0;
      if (spread) {
        const uint4* pres = reinterpret_cast<const uint4*>(A.presence + (size_t)pos * FB_PRES_WORDS);
        const uint4 p0 = __ldcg(pres), p1 = __ldcg(pres + 1);
        bits[0] = p0.x; bits[1] = p0.y; bits[2] = p0.z; bits[3] = p0.w; bits[4] = p1.x; bits[5] = p1.y; bits[6] = p1.z; bits[7] = p1.w;
      } else {
        bits[blockIdx.x >> 5] = 1u << (blockIdx.x & 31);
      }
      __syncthreads();
      if (tid == 0) S.mx.np = atomicAdd(&A.ctl->ord_bump, total);
      __syncthreads();
      uint32_t* ord = A.ordbuf + S.mx.np;
      list_members(S.cx, S, pos, bits, round, ord);
      mixed_walk(S.cx, S, ord, total, round, t);
      if (tid == 0) {
        dup++; mixed_groups++;
        if (spread) {
          ulonglong2* pz = reinterpret_cast<ulonglong2*>(A.presence + (size_t)pos * FB_PRES_WORDS);
          __stcg(pz, make_ulonglong2(0ull, 0ull)); __stcg(pz + 1, make_ulonglong2(0ull, 0ull));
          gentry_clear(&A.aux[pos]);
        }
      }
    }
    __syncthreads();
    if (tid == 0) { S.nfrag = 0; S.nfin = 0; }
    if (round + 1 < rounds) grid_barrier(A.ctl);  // the next round's groups start from a clean group table and the updated slots
  }

  // ---- counters: summed per CTA first ----
  {
    const uint32_t v[8] = {t.over, t.hit, t.miss, t.inserts, t.full, dup, mixed_groups, swept};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, v[k]);
      if (lane == 0 && s) atomicAdd(&S.tally[k], s);
    }
    __syncthreads();
    if (tid < 8 && S.tally[tid]) {
      const int slot[8] = {C_OVER, C_HIT, C_MISS, C_INSERTS, C_FULL, C_DUP_GROUPS, C_MIXED_GROUPS, C_SWEPT};
      atomicAdd(A.counters + slot[tid], (unsigned long long)S.tally[tid]);
    }
  }
  // ---- multi-GPU: the last CTA to finish tells every source that its responses are in place ----
  if (A.n_resp_flags) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) S.mx.serial = atomicAdd(&A.ctl->done_ctr, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (S.mx.serial) {
      if (tid < A.n_resp_flags) {
        __threadfence_system();
        st_release_sys(A.resp_flag[tid], (unsigned long long)A.flag_epoch << 32);
      }
      if (tid == 0) A.ctl->done_ctr = 0;
    }
  }
}

// Places parked items ahead of a maintenance call that reads the table (scan, get, add).
__global__ void k_drain_overflow(const FArgs A) {
  Tally t = {0, 0, 0, 0, 0};
  __shared__ FCtx cx;
  if (threadIdx.x == 0) {
    cx.table = A.table; cx.capacity = A.capacity; cx.gpos = A.gpos; cx.ctl = A.ctl; cx.ovf = A.ovf; cx.counters = A.counters;
    cx.nseg = A.nseg; cx.sweep_chunk = 0; cx.clk = A.clk;
  }
  __syncthreads();
  if (__ldcg(&A.ctl->ovf_count)) drain_overflow(cx, t);
  if (threadIdx.x == 0 && t.inserts) atomicAdd(A.counters + C_INSERTS, (unsigned long long)t.inserts);
}

}  // namespace gub
