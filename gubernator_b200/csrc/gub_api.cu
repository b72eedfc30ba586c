// gub_api.cu — the C ABI (include/gubernator_b200.h) over the kernels in gub_kernels.cuh.
//
// This is the layer that sits where gubernator's WorkerPool sits (workers.go:54-61): gub_create = NewWorkerPool,
// gub_submit* = GetRateLimit for a whole batch, gub_add_items = AddCacheItem/Load, gub_scan = Store, gub_destroy = Close.
// There is no CPU fallback: every entry point that evaluates requests launches CUDA kernels or fails.
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gubernator_b200.h"
#include "gub_kernels.cuh"
#include "gub_batch.cuh"
#include "gub_global.cuh"
#include "gub_p2p.cuh"

extern "C" uint64_t gub_ring_version_(const gub_ring* r);

namespace {

thread_local std::string g_err;
int fail(const std::string& msg) { g_err = msg; return -1; }

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    }                                                                                                    \
  } while (0)

constexpr int PIPE_DEPTH = 4;

struct PipeSlot {
  gub_req* d_req = nullptr;
  gub_resp* d_resp = nullptr;
  gub_creq* d_creq = nullptr;       // compact submissions: what arrives over PCIe
  uint8_t* d_packed = nullptr;      // key-string submissions: [gub_kreq x n][offsets x (n+1)][key bytes]
  size_t packed_cap = 0;
  gub_params* d_params = nullptr;
  size_t cap = 0, params_cap = 0;
  cudaEvent_t in_done = nullptr, out_done = nullptr;
  bool busy = false;
};

uint32_t next_pow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return (uint32_t)p; }

}  // namespace

struct gub_table {
  int device = 0;
  std::mutex mu;
  gub::Slot* table = nullptr;
  uint64_t capacity = 0;
  uint32_t max_batch = 0;
  // per-batch scratch of the four-kernel pipeline (gub_kernels.cuh)
  struct Scratch {
    gub::AuxEntry* aux = nullptr;
    uint32_t *ent = nullptr, *meta = nullptr, *rank = nullptr, *order = nullptr, *mixed_ent = nullptr, *presence = nullptr;
    uint8_t* fragsize = nullptr;
    ulonglong2* commit = nullptr;
    gub::BatchCtr* ctr = nullptr;
    uint32_t epoch = 0;
  } scr;
  uint32_t aux_entries = 0, pres_words = 0, max_blocks = 0;
  // the fused batch kernel (gub_batch.cuh): one scratch set, one cooperative launch per batch
  // Which kernels evaluate a single table's batches (gub_submit*): the four-kernel pipeline (default: consecutive batches overlap,
  // 31.7 us per 65 536-request step on B200) or the single persistent kernel (GUB_PATH=fused: ~65 us latency, no overlap between
  // batches; profiles/README.md).  Rings (gub_p2p_*) always evaluate out of the mailboxes with the persistent kernel.
  bool fused = false;
  bool coop = true;                   // cooperative launch (co-residency of the grid guaranteed by the driver)
  int num_sms = 0;
  uint32_t sweep_chunk = 0;           // slots every CTA sweeps per batch (incremental expiry sweep), 0 = off
  // the four-kernel pipeline's incremental sweep: every SWEEP_EVERY batches one slice of the table is swept between two batches
  // (k_sweep on the batch's stream: nothing else touches the table then); the whole table once every ~65 536 batches
  uint32_t since_sweep = 0;
  uint64_t sweep_cursor = 0;
  gub::GEntry* gaux = nullptr;
  uint32_t* gpres = nullptr;
  unsigned long long* gfrag = nullptr;
  uint16_t* gmembers = nullptr;
  gub::FCtl* ctl = nullptr;
  gub::OvfItem* ovf = nullptr;
  gub::InvIndex inv{};                // CacheItem.InvalidAt side index
  unsigned long long* trace = nullptr; // per-CTA phase timestamps of the last k_batch launch (gub_set_trace)
  unsigned long long* ktrace = nullptr; // pipeline kernels: [4][KT_BLOCKS][KT_MARKS] latest time stamps (gub_set_trace / gub_get_ktrace)
  unsigned long long* counters = nullptr;
  // ordering between streams that touch the shared scratch
  cudaEvent_t last_done = nullptr;
  cudaStream_t last_stream = nullptr; // stream of the most recent table-touching work
  bool last_pending = false;          // work was enqueued on last_stream after (or without) the last event record
  // host path
  cudaStream_t s_h2d = nullptr, s_compute = nullptr, s_d2h = nullptr;
  cudaEvent_t compute_done[PIPE_DEPTH] = {};
  PipeSlot pipe[PIPE_DEPTH];
  int next_slot = 0;
  // maintenance scratch
  unsigned long long* d_scalar = nullptr;
  // route
  uint64_t* d_ring_pts = nullptr; int32_t* d_ring_peers = nullptr; uint16_t* d_ring_lut = nullptr; uint32_t ring_npts = 0; const gub_ring* ring_cached = nullptr; uint64_t ring_version = 0;
  uint8_t* d_owner = nullptr; uint32_t* d_tile_counts = nullptr;
  size_t owner_cap = 0, tiles_cap = 0;
  // optional per-kernel timing (bench.py's roofline leg): events bracket every kernel of the batch path
  bool pdl = true;                    // programmatic dependent launch between the batch kernels (GUB_PDL=0 disables)
  // GUB_PIPE_TRACE=<file>: stage timestamps of the first host-to-host submissions (diagnostic; see profiles/README.md)
  std::string trace_path;
  cudaEvent_t trace_base = nullptr;
  std::vector<cudaEvent_t> trace_ev;  // 6 per traced submission
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;   // 5 events per pending chunk
  size_t prof_pending = 0;            // chunks recorded and not yet accumulated
  double prof_ms[4] = {0, 0, 0, 0};
  uint64_t prof_launches = 0;
};

namespace {

// Folds finished per-kernel event triples into the running totals.  force: wait for everything pending.
int prof_flush(gub_table* t, bool force) {
  if (!force && t->prof_pending < 1024) return 0;
  for (size_t c = 0; c < t->prof_pending; c++) {
    cudaEvent_t* pe = &t->prof_ev[c * 5];
    CK(cudaEventSynchronize(pe[4]));
    for (int k = 0; k < 4; k++) { float ms = 0; CK(cudaEventElapsedTime(&ms, pe[k], pe[k + 1])); t->prof_ms[k] += ms; }
    t->prof_launches++;
  }
  t->prof_pending = 0;
  return 0;
}

constexpr size_t TRACE_MAX = 400;  // submissions
inline bool tracing(gub_table* t) { return t->trace_base && t->trace_ev.size() < TRACE_MAX * 6; }
int trace_mark(gub_table* t, cudaStream_t st) {
  cudaEvent_t e;
  CK(cudaEventCreate(&e));
  CK(cudaEventRecord(e, st));
  t->trace_ev.push_back(e);
  return 0;
}
void trace_dump(gub_table* t) {
  if (!t->trace_base || t->trace_ev.empty()) return;
  cudaDeviceSynchronize();
  if (FILE* f = std::fopen(t->trace_path.c_str(), "w")) {
    std::fprintf(f, "# us since table creation: h2d_start h2d_done compute_start compute_done d2h_start d2h_done\n");
    for (size_t i = 0; i + 6 <= t->trace_ev.size(); i += 6) {
      for (int k = 0; k < 6; k++) { float ms = 0; cudaEventElapsedTime(&ms, t->trace_base, t->trace_ev[i + k]); std::fprintf(f, "%.1f%c", ms * 1e3, k == 5 ? '\n' : ' '); }
    }
    std::fclose(f);
  }
  for (auto e : t->trace_ev) cudaEventDestroy(e);
  t->trace_ev.clear();
  cudaEventDestroy(t->trace_base);
  t->trace_base = nullptr;  // one dump per table
}

// Launches one batch kernel, with programmatic stream serialization when enabled (the kernels call griddepcontrol.wait
// before touching anything an earlier kernel produced; without the attribute that instruction is a no-op).
template <typename K>
cudaError_t launch_k(gub_table* t, K kernel, uint32_t grid, uint32_t block, cudaStream_t st, const gub::BatchArgs& A, bool plain = false) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  // With per-kernel profiling on, an event record sits between consecutive kernels: they are launched plainly then (full stream
  // order; griddepcontrol.* are no-ops without the attribute) rather than as programmatic dependents of "the previous kernel".
  // `plain`: the stream's previous operation is not a kernel of this chain (the memsets at the epoch wrap): full stream order.
  cfg.attrs = attr; cfg.numAttrs = (t->pdl && !t->prof && !plain) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, A);
}

void fused_base_args(gub_table* t, const gub_clock* clk, gub::FArgs& A) {
  std::memset(&A, 0, sizeof A);
  A.table = t->table; A.capacity = t->capacity; A.aux = t->gaux; A.presence = t->gpres; A.fragrow = t->gfrag; A.members = t->gmembers;
  A.ctl = t->ctl; A.ovf = t->ovf; A.counters = t->counters; A.sweep_chunk = t->sweep_chunk; A.trace = t->trace; A.inv = t->inv; A.clk = *clk;
}

// One launch of k_batch over the segments in A (A.seg / A.nseg / flags filled by the caller).  `total_hint` = number of requests
// when the host knows it (sizes the grid for small batches), 0 = only the device knows: one CTA per SM.
int launch_fused(gub_table* t, const gub::FArgs& A, uint64_t total_hint, cudaStream_t st) {
  uint32_t grid = (uint32_t)t->num_sms;
  if (total_hint) grid = (uint32_t)std::min<uint64_t>(grid, std::max<uint64_t>(1, (total_hint + 127) / 128 + A.nseg));
  cudaEvent_t* pe = nullptr;
  if (t->prof) {
    if (prof_flush(t, false)) return -1;
    if (t->prof_ev.size() < (t->prof_pending + 1) * 5) {
      for (int k = 0; k < 5; k++) { cudaEvent_t e; CK(cudaEventCreate(&e)); t->prof_ev.push_back(e); }
    }
    pe = &t->prof_ev[t->prof_pending * 5];
    t->prof_pending++;
    CK(cudaEventRecord(pe[0], st));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(gub::FB_THREADS); cfg.dynamicSmemBytes = sizeof(gub::FSmem); cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (t->coop) { attr[na].id = cudaLaunchAttributeCooperative; attr[na].val.cooperative = 1; na++; }
  if (t->pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
  cfg.attrs = attr; cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gub::k_batch, A);
  if (e != cudaSuccess && t->coop && t->pdl) {  // the two attributes do not combine on this driver: keep the co-residency guarantee
    cudaGetLastError();
    t->pdl = false;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, gub::k_batch, A);
  }
  if (e != cudaSuccess) return fail(std::string("k_batch launch: ") + cudaGetErrorString(e));
  if (pe) { for (int k = 1; k < 5; k++) CK(cudaEventRecord(pe[k], st)); }
  return 0;
}

// Ring mode: the batch is the concatenation of mailbox segments (see BatchArgs::nseg).
struct SegDesc {
  uint32_t nseg = 0;
  const uint32_t* seg_off = nullptr;
  const gub_req* reqs[gub::MAX_SHARDS] = {};
  gub_resp* out[gub::MAX_SHARDS] = {};
};

// One batch (<= max_batch requests) through the four-kernel pipeline, chained with programmatic dependent launch.
int launch_chunk(gub_table* t, const gub_req* d_reqs, uint32_t n, const gub_clock* clk, gub_resp* d_out, cudaStream_t st,
                 const uint32_t* n_dev = nullptr, uint32_t n_off = 0, const SegDesc* seg = nullptr) {
  gub_table::Scratch& sc = t->scr;
  bool wrapped = false;
  if (sc.epoch >= 65535u) {  // 16-bit epoch tags wrapped: clear the grouping table so stale tags cannot alias
    wrapped = true;
    CK(cudaMemsetAsync(sc.aux, 0, (size_t)t->aux_entries * sizeof(gub::AuxEntry), st));
    // 65535 -> 1 keeps the parity: the batch after the wrap reuses ctr[1], which k_rank (it only resets the OTHER parity) left
    // holding batch 65535's allocators.  Clear both.
    CK(cudaMemsetAsync(sc.ctr, 0, 2 * sizeof(gub::BatchCtr), st));
    sc.epoch = 0;
  }
  sc.epoch++;
  gub::BatchArgs A;
  std::memset(&A, 0, sizeof A);
  if (seg) {
    A.nseg = seg->nseg; A.seg_off = seg->seg_off;
    for (uint32_t k = 0; k < seg->nseg; k++) { A.seg_reqs[k] = seg->reqs[k]; A.seg_out[k] = seg->out[k]; }
  }
  A.table = t->table; A.capacity = t->capacity; A.reqs = d_reqs; A.out = d_out; A.n = n; A.n_dev = n_dev; A.n_off = n_off; A.epoch = sc.epoch;
  A.aux = sc.aux; A.aux_mask = t->aux_entries - 1; A.presence = sc.presence; A.fragsize = sc.fragsize;
  A.pres_words = t->pres_words; A.max_blocks = t->max_blocks; A.ent = sc.ent; A.meta = sc.meta; A.rank = sc.rank;
  A.commit = sc.commit; A.order = sc.order; A.mixed_ent = sc.mixed_ent; A.ctr = sc.ctr;
  A.counters = t->counters; A.ovf = t->ovf; A.ovf_count = &t->ctl->ovf_count; A.inv = t->inv;
  A.ktrace = t->ktrace;
  A.clk = *clk;
  const uint32_t blocks = (n + 255) / 256;
  cudaEvent_t* pe = nullptr;
  if (t->prof) {
    if (prof_flush(t, false)) return -1;
    if (t->prof_ev.size() < (t->prof_pending + 1) * 5) {
      for (int k = 0; k < 5; k++) { cudaEvent_t e; CK(cudaEventCreate(&e)); t->prof_ev.push_back(e); }
    }
    pe = &t->prof_ev[t->prof_pending * 5];
    t->prof_pending++;
    CK(cudaEventRecord(pe[0], st));
  }
  const uint32_t fin_blocks = std::min<uint32_t>(148u, std::max<uint32_t>(1u, n / 2));
  if (seg) {
    CK(launch_k(t, gub::k_group<true>, blocks, gub::GROUP_THREADS, st, A, wrapped));
    if (pe) CK(cudaEventRecord(pe[1], st));
    CK(launch_k(t, gub::k_rank<true>, blocks, gub::GROUP_THREADS, st, A));
    if (pe) CK(cudaEventRecord(pe[2], st));
    CK(launch_k(t, gub::k_eval<true>, blocks, gub::GROUP_THREADS, st, A));
    if (pe) CK(cudaEventRecord(pe[3], st));
    CK(launch_k(t, gub::k_finish<true>, fin_blocks, gub::MIXED_THREADS, st, A));
  } else {
    CK(launch_k(t, gub::k_group<false>, blocks, gub::GROUP_THREADS, st, A, wrapped));
    if (pe) CK(cudaEventRecord(pe[1], st));
    CK(launch_k(t, gub::k_rank<false>, blocks, gub::GROUP_THREADS, st, A));
    if (pe) CK(cudaEventRecord(pe[2], st));
    CK(launch_k(t, gub::k_eval<false>, blocks, gub::GROUP_THREADS, st, A));
    if (pe) CK(cudaEventRecord(pe[3], st));
    // non-uniform groups, one block each (grid-stride; normally none: the blocks return at once)
    CK(launch_k(t, gub::k_finish<false>, fin_blocks, gub::MIXED_THREADS, st, A));
  }
  if (pe) CK(cudaEventRecord(pe[4], st));
  CK(cudaGetLastError());
  return 0;
}

// The pipeline's share of the capacity policy (lrucache.go:115 frees expired items lazily on access; a table slot whose key never
// comes back would stay occupied for ever): every SWEEP_EVERY batches, 1/64 of the table is swept on the batch's stream, after the
// batch — removed / expired entries become tombstones, which inserts reuse.  (The persistent kernel sweeps a few slots per batch
// itself.)  Off with gub_set_sweep(t, 0) / GUB_SWEEP=0.
constexpr uint32_t SWEEP_EVERY = 1024;
int maybe_sweep(gub_table* t, const gub_clock* clk, cudaStream_t st) {
  if (!t->sweep_chunk || ++t->since_sweep < SWEEP_EVERY) return 0;
  t->since_sweep = 0;
  const uint64_t slice = (t->capacity + 63) / 64;
  const uint64_t lo = t->sweep_cursor, hi = std::min<uint64_t>(t->capacity, lo + slice);
  t->sweep_cursor = hi >= t->capacity ? 0 : hi;
  gub::k_sweep<<<148 * 4, 256, 0, st>>>(t->table, lo, hi, clk->now_ms, t->counters + gub::C_SWEPT);
  CK(cudaGetLastError());
  return 0;
}

// Table-touching work from different caller streams must be ordered.  The common case (same stream as last time) costs
// nothing; only a change of stream records an event on the old stream and makes the new one wait for it.
int order_after_last(gub_table* t, cudaStream_t st) {
  if (t->last_pending && t->last_stream != st) {
    CK(cudaEventRecord(t->last_done, t->last_stream));
    CK(cudaStreamWaitEvent(st, t->last_done, 0));
    t->last_pending = false;
  }
  return 0;
}

// n_dev != nullptr: the real batch size is *n_dev (<= n) on the device; launches are sized for n and trim themselves.
int launch_batch(gub_table* t, const gub_req* d_reqs, size_t n, const gub_clock* clk, gub_resp* d_out, cudaStream_t st,
                 const uint32_t* n_dev = nullptr) {
  if (order_after_last(t, st)) return -1;
  if (t->fused) {  // any size in one launch: the kernel takes the batch in rounds of one tile per CTA
    if (n > 0xFFFFFFFFull) return fail("batch too large");
    gub::FArgs A;
    fused_base_args(t, clk, A);
    A.seg[0].reqs = d_reqs; A.seg[0].out = d_out; A.seg[0].n = (uint32_t)n; A.seg[0].n_dev = n_dev;
    A.nseg = 1;
    if (((uintptr_t)d_reqs & 15u) || ((uintptr_t)d_out & 15u)) return fail("request / response buffers must be 16-byte aligned");
    if (launch_fused(t, A, n_dev ? 0 : n, st)) return -1;
    t->last_stream = st; t->last_pending = true;
    return 0;
  }
  for (size_t off = 0; off < n; off += t->max_batch) {
    const uint32_t m = (uint32_t)std::min<size_t>(t->max_batch, n - off);
    if (launch_chunk(t, d_reqs + off, m, clk, d_out + off, st, n_dev, (uint32_t)off)) return -1;
  }
  if (maybe_sweep(t, clk, st)) return -1;
  t->last_stream = st; t->last_pending = true;
  return 0;
}

int ensure_slot(PipeSlot& s, size_t n) {
  if (s.cap >= n) return 0;
  if (s.d_req) cudaFree(s.d_req);
  if (s.d_resp) cudaFree(s.d_resp);
  if (s.d_creq) cudaFree(s.d_creq);
  s.d_req = nullptr; s.d_resp = nullptr; s.d_creq = nullptr; s.cap = 0;
  CK(cudaMalloc(&s.d_req, n * sizeof(gub_req)));
  CK(cudaMalloc(&s.d_resp, n * sizeof(gub_resp)));
  s.cap = n;
  return 0;
}

gub::DevItem to_dev(const gub_item& it) {
  gub::DevItem d;
  std::memset(&d, 0, sizeof d);
  d.key = it.key_xxh64 < 2 ? it.key_xxh64 + 2 : it.key_xxh64;
  d.tag = it.key_fnv1 >> 8;
  const bool leaky = it.algorithm == GUB_LEAKY_BUCKET;
  d.flags = gub::F_LIVE | (leaky ? gub::F_LEAKY : 0u) | ((!leaky && it.status == GUB_OVER_LIMIT) ? gub::F_OVER : 0u);
  d.w[0] = (uint64_t)it.limit; d.w[1] = (uint64_t)it.duration;
  if (leaky) std::memcpy(&d.w[2], &it.remaining_f, 8); else d.w[2] = (uint64_t)it.remaining;
  d.w[3] = (uint64_t)it.stamp; d.w[4] = leaky ? (uint64_t)it.burst : 0; d.w[5] = (uint64_t)it.expire_at;
  d.invalid_at = it.invalid_at;
  return d;
}
gub_item from_dev(const gub::DevItem& d) {
  gub_item it;
  std::memset(&it, 0, sizeof it);
  it.key_xxh64 = d.key; it.key_fnv1 = d.tag << 8;
  const bool leaky = (d.flags & gub::F_LEAKY) != 0;
  it.algorithm = leaky ? GUB_LEAKY_BUCKET : GUB_TOKEN_BUCKET;
  it.status = (d.flags & gub::F_OVER) ? GUB_OVER_LIMIT : GUB_UNDER_LIMIT;
  it.limit = (int64_t)d.w[0]; it.duration = (int64_t)d.w[1];
  if (leaky) std::memcpy(&it.remaining_f, &d.w[2], 8); else it.remaining = (int64_t)d.w[2];
  it.stamp = (int64_t)d.w[3]; it.burst = (int64_t)d.w[4]; it.expire_at = (int64_t)d.w[5];
  it.invalid_at = d.invalid_at;
  return it;
}

// Per-batch scratch of the four-kernel pipeline for batches of up to B requests (replaces what was there: the device must be idle).
int alloc_scratch(gub_table* t, uint32_t B) {
  auto& sc = t->scr;
  void* old[] = {sc.aux, sc.ent, sc.meta, sc.rank, sc.order, sc.mixed_ent, sc.presence, sc.fragsize, sc.commit, sc.ctr};
  for (void* p : old) if (p) cudaFree(p);
  sc = gub_table::Scratch();
  B = std::max<uint32_t>(1024u, std::min<uint32_t>(B, 262144u));
  B = (B + 255u) & ~255u;
  t->max_batch = B;
  t->aux_entries = next_pow2((uint64_t)B * 4);
  t->max_blocks = (B / gub::GROUP_THREADS + 127u) & ~127u;  // fragment-size row per group entry; a multiple of 128 blocks = 4 presence words (16-byte loads)
  t->pres_words = t->max_blocks / 32;
  struct { void** p; size_t bytes; } want[] = {
      {(void**)&sc.aux, (size_t)t->aux_entries * sizeof(gub::AuxEntry)},
      {(void**)&sc.presence, (size_t)t->aux_entries * t->pres_words * 4},
      {(void**)&sc.fragsize, (size_t)t->aux_entries * t->max_blocks},
      {(void**)&sc.commit, (size_t)t->aux_entries * 6 * sizeof(ulonglong2)},
      {(void**)&sc.ent, (size_t)B * 4}, {(void**)&sc.meta, (size_t)B * 4}, {(void**)&sc.rank, (size_t)B * 4}, {(void**)&sc.order, (size_t)B * 4},
      {(void**)&sc.mixed_ent, ((size_t)B / 2 + 1) * 4}, {(void**)&sc.ctr, 2 * sizeof(gub::BatchCtr)},
  };
  for (auto& w : want) {
    cudaError_t e = cudaMalloc(w.p, w.bytes);
    if (e != cudaSuccess) return fail(std::string("cudaMalloc(batch scratch): ") + cudaGetErrorString(e));
    CK(cudaMemset(*w.p, 0, w.bytes));
  }
  return 0;
}

}  // namespace

extern "C" {

const char* gub_last_error(void) { return g_err.c_str(); }
int gub_abi_version(void) { return GUB_ABI_VERSION; }

void gub_destroy(gub_table* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  cudaDeviceSynchronize();
  trace_dump(t);
  void* ptrs[] = {t->table, t->counters, t->d_scalar, t->d_ring_pts, t->d_ring_peers, t->d_ring_lut, t->d_owner, t->d_tile_counts,
                  t->gaux, t->gpres, t->gfrag, t->gmembers, t->ctl, t->ovf, t->trace, t->ktrace, t->inv.e};
  for (void* p : ptrs) if (p) cudaFree(p);
  {
    auto& sc = t->scr;
    void* sp[] = {sc.aux, sc.ent, sc.meta, sc.rank, sc.order, sc.mixed_ent, sc.presence, sc.fragsize, sc.commit, sc.ctr};
    for (void* p : sp) if (p) cudaFree(p);
  }
  for (auto& s : t->pipe) {
    if (s.d_req) cudaFree(s.d_req);
    if (s.d_resp) cudaFree(s.d_resp);
    if (s.d_creq) cudaFree(s.d_creq);
    if (s.d_packed) cudaFree(s.d_packed);
    if (s.d_params) cudaFree(s.d_params);
    if (s.in_done) cudaEventDestroy(s.in_done);
    if (s.out_done) cudaEventDestroy(s.out_done);
  }
  for (auto& e : t->compute_done) if (e) cudaEventDestroy(e);
  for (auto& e : t->prof_ev) cudaEventDestroy(e);
  if (t->last_done) cudaEventDestroy(t->last_done);
  if (t->s_h2d) cudaStreamDestroy(t->s_h2d);
  if (t->s_compute) cudaStreamDestroy(t->s_compute);
  if (t->s_d2h) cudaStreamDestroy(t->s_d2h);
  delete t;
}

int gub_create(const gub_config* cfg, gub_table** out) {
  if (!cfg || !out) return fail("gub_create: null argument");
  *out = nullptr;
  if (cfg->capacity_slots < 64) return fail("gub_create: capacity_slots must be >= 64");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail("gub_create: no such CUDA device");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) return fail("gub_create: this library is built for sm_100a (B200) only");
  gub_table* t = new gub_table();
  t->device = cfg->device;
  if (const char* e = getenv("GUB_PDL")) t->pdl = std::atoi(e) != 0;
  t->capacity = cfg->capacity_slots;
  t->max_batch = cfg->max_batch ? cfg->max_batch : 65536u;
  if (t->max_batch < 1024) t->max_batch = 1024;
  if (t->max_batch > 262144u) t->max_batch = 262144u;
  t->max_batch = (t->max_batch + 255u) & ~255u;
#define ALLOC(ptr, bytes)                                                    \
  do {                                                                       \
    cudaError_t e__ = cudaMalloc((void**)&(ptr), (bytes));                   \
    if (e__ != cudaSuccess) {                                                \
      fail(std::string("cudaMalloc(" #ptr "): ") + cudaGetErrorString(e__)); \
      gub_destroy(t);                                                        \
      return -1;                                                             \
    }                                                                        \
    cudaMemset((ptr), 0, (bytes));                                           \
  } while (0)
  ALLOC(t->table, t->capacity * sizeof(gub::Slot));
  if (alloc_scratch(t, t->max_batch)) { gub_destroy(t); return -1; }
  t->num_sms = std::min<int>(prop.multiProcessorCount, gub::FB_MAX_GRID);
  if (const char* e = getenv("GUB_PATH")) t->fused = std::string(e) == "fused";
  if (const char* e = getenv("GUB_COOP")) t->coop = std::atoi(e) != 0;
  ALLOC(t->gaux, (size_t)gub::FB_AUX_ENTRIES * sizeof(gub::GEntry));
  ALLOC(t->gpres, (size_t)gub::FB_AUX_ENTRIES * gub::FB_PRES_WORDS * 4);
  ALLOC(t->gfrag, (size_t)gub::FB_AUX_ENTRIES * gub::FB_ROW * 8);
  ALLOC(t->gmembers, (size_t)gub::FB_MAX_GRID * gub::FB_THREADS * 2);
  ALLOC(t->ctl, sizeof(gub::FCtl));
  ALLOC(t->ovf, (size_t)gub::FB_OVF_CAP * sizeof(gub::OvfItem));
  ALLOC(t->inv.e, (size_t)65536 * sizeof(gub::InvEntry));
  t->inv.mask = 65535;
  {
    // incremental expiry sweep: the whole table once every ~65536 batches (GUB_SWEEP=<slots per CTA per batch>, 0 = off)
    uint64_t chunk = (t->capacity + (uint64_t)t->num_sms * 65536 - 1) / ((uint64_t)t->num_sms * 65536);
    if (const char* e = getenv("GUB_SWEEP")) chunk = (uint64_t)std::atoll(e);
    t->sweep_chunk = (uint32_t)std::min<uint64_t>(chunk, 1024);
    cudaError_t e2 = cudaFuncSetAttribute(gub::k_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(gub::FSmem));
    if (e2 != cudaSuccess) { fail(std::string("cudaFuncSetAttribute(k_batch): ") + cudaGetErrorString(e2)); gub_destroy(t); return -1; }
  }
  ALLOC(t->counters, gub::C_COUNT * sizeof(unsigned long long));
  ALLOC(t->d_scalar, 4 * sizeof(unsigned long long));
#undef ALLOC
  CK(cudaStreamCreateWithFlags(&t->s_h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&t->s_compute, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&t->s_d2h, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&t->last_done, cudaEventDisableTiming));
  for (int i = 0; i < PIPE_DEPTH; i++) {
    CK(cudaEventCreateWithFlags(&t->pipe[i].in_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&t->pipe[i].out_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&t->compute_done[i], cudaEventDisableTiming));
  }
  if (const char* e = getenv("GUB_PIPE_TRACE")) {
    t->trace_path = e;
    CK(cudaEventCreate(&t->trace_base));
    CK(cudaEventRecord(t->trace_base, t->s_h2d));
  }
  CK(cudaDeviceSynchronize());
  *out = t;
  return 0;
}

int gub_submit_device(gub_table* t, const gub_req* d_reqs, size_t n, const gub_clock* clk, gub_resp* d_out, void* stream) {
  if (!t || !clk || (n && (!d_reqs || !d_out))) return fail("gub_submit_device: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  return launch_batch(t, d_reqs, n, clk, d_out, (cudaStream_t)stream);
}

int gub_submit_device_n(gub_table* t, const gub_req* d_reqs, size_t n_cap, const uint32_t* d_n, const gub_clock* clk, gub_resp* d_out, void* stream) {
  if (!t || !clk || !d_n || (n_cap && (!d_reqs || !d_out))) return fail("gub_submit_device_n: null argument");
  if (n_cap == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  return launch_batch(t, d_reqs, n_cap, clk, d_out, (cudaStream_t)stream, d_n);
}

int gub_pipeline_depth(gub_table*) { return PIPE_DEPTH; }

int gub_submit_async(gub_table* t, const gub_req* reqs, size_t n, const gub_clock* clk, gub_resp* out, int* ticket) {
  if (!t || !clk || !ticket || (n && (!reqs || !out))) return fail("gub_submit_async: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  const int si = t->next_slot;
  t->next_slot = (t->next_slot + 1) % PIPE_DEPTH;
  PipeSlot& s = t->pipe[si];
  if (s.busy) { CK(cudaEventSynchronize(s.out_done)); s.busy = false; }
  *ticket = si;
  if (n == 0) { CK(cudaEventRecord(s.out_done, t->s_d2h)); s.busy = true; return 0; }
  if (ensure_slot(s, n)) return -1;
  CK(cudaMemcpyAsync(s.d_req, reqs, n * sizeof(gub_req), cudaMemcpyHostToDevice, t->s_h2d));
  CK(cudaEventRecord(s.in_done, t->s_h2d));
  CK(cudaStreamWaitEvent(t->s_compute, s.in_done, 0));
  if (launch_batch(t, s.d_req, n, clk, s.d_resp, t->s_compute)) return -1;
  CK(cudaEventRecord(t->compute_done[si], t->s_compute));
  CK(cudaStreamWaitEvent(t->s_d2h, t->compute_done[si], 0));
  CK(cudaMemcpyAsync(out, s.d_resp, n * sizeof(gub_resp), cudaMemcpyDeviceToHost, t->s_d2h));
  CK(cudaEventRecord(s.out_done, t->s_d2h));
  s.busy = true;
  return 0;
}

int gub_submit_compact_async(gub_table* t, const gub_creq* reqs, size_t n, const gub_params* params, size_t n_params, int64_t created_base,
                             const gub_clock* clk, gub_resp* out, int* ticket) {
  if (!t || !clk || !ticket || (n && (!reqs || !out || !params))) return fail("gub_submit_compact_async: null argument");
  if (n > 0xFFFFFFFFull || n_params > 0xFFFFFFFFull) return fail("gub_submit_compact_async: too large");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  const int si = t->next_slot;
  t->next_slot = (t->next_slot + 1) % PIPE_DEPTH;
  PipeSlot& s = t->pipe[si];
  if (s.busy) { CK(cudaEventSynchronize(s.out_done)); s.busy = false; }
  *ticket = si;
  if (n == 0) { CK(cudaEventRecord(s.out_done, t->s_d2h)); s.busy = true; return 0; }
  if (ensure_slot(s, n)) return -1;
  if (!s.d_creq) CK(cudaMalloc(&s.d_creq, s.cap * sizeof(gub_creq)));
  const bool inl = n_params <= gub::INLINE_PARAMS;  // the table travels in the kernel arguments: no second copy
  if (!inl && (s.params_cap < n_params || !s.d_params)) {
    if (s.d_params) cudaFree(s.d_params);
    s.d_params = nullptr;
    s.params_cap = std::max<size_t>(n_params, 1024);
    CK(cudaMalloc(&s.d_params, s.params_cap * sizeof(gub_params)));
  }
  const bool tr = tracing(t);
  if (tr && trace_mark(t, t->s_h2d)) return -1;
  CK(cudaMemcpyAsync(s.d_creq, reqs, n * sizeof(gub_creq), cudaMemcpyHostToDevice, t->s_h2d));
  if (!inl) CK(cudaMemcpyAsync(s.d_params, params, n_params * sizeof(gub_params), cudaMemcpyHostToDevice, t->s_h2d));
  CK(cudaEventRecord(s.in_done, t->s_h2d));
  if (tr && trace_mark(t, t->s_h2d)) return -1;
  CK(cudaStreamWaitEvent(t->s_compute, s.in_done, 0));
  if (tr && trace_mark(t, t->s_compute)) return -1;
  // The expansion runs on the compute stream: a kernel behind a copy on the copy stream costs an engine switch per batch that
  // stalls the ingest stage by ~45 us under PCIe load (measured, profiles/r01_e2e_pipeline.md).
  if (inl) {
    gub::InlineParams P;
    static_assert(sizeof(gub_params) == 32, "gub_params is two 16-byte words");
    std::memset(&P, 0, sizeof P);
    std::memcpy(&P, params, n_params * sizeof(gub_params));
    gub::k_expand_inline<<<(unsigned)((n + 255) / 256), 256, 0, t->s_compute>>>(s.d_creq, (uint32_t)n, P, (uint32_t)n_params, created_base, s.d_req);
  } else {
    gub::k_expand<<<(unsigned)((n + 255) / 256), 256, 0, t->s_compute>>>(s.d_creq, (uint32_t)n, s.d_params, (uint32_t)n_params, created_base, s.d_req);
  }
  if (launch_batch(t, s.d_req, n, clk, s.d_resp, t->s_compute)) return -1;
  CK(cudaEventRecord(t->compute_done[si], t->s_compute));
  if (tr && trace_mark(t, t->s_compute)) return -1;
  CK(cudaStreamWaitEvent(t->s_d2h, t->compute_done[si], 0));
  if (tr && trace_mark(t, t->s_d2h)) return -1;
  CK(cudaMemcpyAsync(out, s.d_resp, n * sizeof(gub_resp), cudaMemcpyDeviceToHost, t->s_d2h));
  CK(cudaEventRecord(s.out_done, t->s_d2h));
  if (tr && trace_mark(t, t->s_d2h)) return -1;
  if (tr && t->trace_ev.size() >= TRACE_MAX * 6) trace_dump(t);
  s.busy = true;
  return 0;
}

/* Key strings in, responses out: the host ships raw key bytes and 16-byte request records; hashing (XXH64 + FNV-1), the
 * expansion of the per-limit parameters and the evaluation all run on the device.  `packed` = [gub_kreq x n][uint32 offsets x
 * (n + 1)][key bytes] in one buffer (one host-to-device copy per batch: small separate copies cost the copy stream far more than
 * their size, profiles/r01_e2e_pipeline.md); gub_keys_layout() gives the offsets of the three parts. */
int gub_keys_layout(size_t n, size_t key_bytes, size_t* offsets_at, size_t* bytes_at, size_t* total) {
  if (!offsets_at || !bytes_at || !total) return fail("gub_keys_layout: null argument");
  *offsets_at = n * sizeof(gub_kreq);
  *bytes_at = *offsets_at + (n + 1) * 4;
  *total = (*bytes_at + key_bytes + 15) & ~(size_t)15;
  return 0;
}

int gub_submit_keys_async(gub_table* t, const void* packed, size_t packed_bytes, size_t n, const gub_params* params, size_t n_params, int64_t created_base,
                          const gub_clock* clk, gub_resp* out, int* ticket) {
  if (!t || !clk || !ticket || (n && (!packed || !out || !params))) return fail("gub_submit_keys_async: null argument");
  if (n > 0x7FFFFFFFull || n_params > 0xFFFFFFFFull) return fail("gub_submit_keys_async: too large");
  if (n && packed_bytes < n * sizeof(gub_kreq) + (n + 1) * 4) return fail("gub_submit_keys_async: packed buffer too small");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  const int si = t->next_slot;
  t->next_slot = (t->next_slot + 1) % PIPE_DEPTH;
  PipeSlot& s = t->pipe[si];
  if (s.busy) { CK(cudaEventSynchronize(s.out_done)); s.busy = false; }
  *ticket = si;
  if (n == 0) { CK(cudaEventRecord(s.out_done, t->s_d2h)); s.busy = true; return 0; }
  if (ensure_slot(s, n)) return -1;
  if (s.packed_cap < packed_bytes) {
    if (s.d_packed) cudaFree(s.d_packed);
    s.d_packed = nullptr;
    s.packed_cap = std::max<size_t>(packed_bytes + packed_bytes / 4, (size_t)1 << 20);
    CK(cudaMalloc(&s.d_packed, s.packed_cap));
  }
  const bool inl = n_params <= gub::INLINE_PARAMS;
  if (!inl && (s.params_cap < n_params || !s.d_params)) {
    if (s.d_params) cudaFree(s.d_params);
    s.d_params = nullptr;
    s.params_cap = std::max<size_t>(n_params, 1024);
    CK(cudaMalloc(&s.d_params, s.params_cap * sizeof(gub_params)));
  }
  CK(cudaMemcpyAsync(s.d_packed, packed, packed_bytes, cudaMemcpyHostToDevice, t->s_h2d));
  if (!inl) CK(cudaMemcpyAsync(s.d_params, params, n_params * sizeof(gub_params), cudaMemcpyHostToDevice, t->s_h2d));
  CK(cudaEventRecord(s.in_done, t->s_h2d));
  CK(cudaStreamWaitEvent(t->s_compute, s.in_done, 0));
  gub::InlineParams P;
  std::memset(&P, 0, sizeof P);
  if (inl) std::memcpy(&P, params, n_params * sizeof(gub_params));
  gub::k_hash_expand<<<(unsigned)((n + 255) / 256), 256, 0, t->s_compute>>>(s.d_packed, (uint32_t)n, P, (uint32_t)n_params, inl ? nullptr : s.d_params, created_base, s.d_req);
  if (launch_batch(t, s.d_req, n, clk, s.d_resp, t->s_compute)) return -1;
  CK(cudaEventRecord(t->compute_done[si], t->s_compute));
  CK(cudaStreamWaitEvent(t->s_d2h, t->compute_done[si], 0));
  CK(cudaMemcpyAsync(out, s.d_resp, n * sizeof(gub_resp), cudaMemcpyDeviceToHost, t->s_d2h));
  CK(cudaEventRecord(s.out_done, t->s_d2h));
  s.busy = true;
  return 0;
}

int gub_wait(gub_table* t, int ticket) {
  if (!t || ticket < 0 || ticket >= PIPE_DEPTH) return fail("gub_wait: bad ticket");
  cudaEvent_t ev;
  {
    std::lock_guard<std::mutex> lk(t->mu);
    if (!t->pipe[ticket].busy) return 0;
    ev = t->pipe[ticket].out_done;
  }
  CK(cudaSetDevice(t->device));
  CK(cudaEventSynchronize(ev));
  std::lock_guard<std::mutex> lk(t->mu);
  t->pipe[ticket].busy = false;
  return 0;
}

int gub_submit_compact(gub_table* t, const gub_creq* reqs, size_t n, const gub_params* params, size_t n_params, int64_t created_base,
                       const gub_clock* clk, gub_resp* out) {
  int ticket = -1;
  if (gub_submit_compact_async(t, reqs, n, params, n_params, created_base, clk, out, &ticket)) return -1;
  return gub_wait(t, ticket);
}

int gub_submit(gub_table* t, const gub_req* reqs, size_t n, const gub_clock* clk, gub_resp* out) {
  int ticket = -1;
  if (gub_submit_async(t, reqs, n, clk, out, &ticket)) return -1;
  return gub_wait(t, ticket);
}

void* gub_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
  return p;
}
void gub_host_free(void* p) { if (p) cudaFreeHost(p); }

namespace {
// Items parked by a batch whose probe window was full are placed before anything else reads or writes the table.
int drain_pending(gub_table* t) {
  gub_clock clk;
  std::memset(&clk, 0, sizeof clk);
  clk.now_ms = INT64_MIN;  // nothing counts as expired here: only free, removed and then the earliest-expiring entries give way
  gub::FArgs A;
  fused_base_args(t, &clk, A);
  gub::k_drain_overflow<<<1, 32>>>(A);
  CK(cudaGetLastError());
  return 0;
}
}  // namespace

int gub_add_items(gub_table* t, const gub_item* items, size_t n) {
  if (!t || (n && !items)) return fail("gub_add_items: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  // last one wins for duplicate keys (sequential AddCacheItem calls would overwrite in order)
  std::vector<gub::DevItem> dev;
  dev.reserve(n);
  std::unordered_map<uint64_t, size_t> seen;
  seen.reserve(n * 2);
  for (size_t i = 0; i < n; i++) {
    if (items[i].algorithm != GUB_TOKEN_BUCKET && items[i].algorithm != GUB_LEAKY_BUCKET) continue;  // Value would be nil (gubernator.go:434-451)
    gub::DevItem d = to_dev(items[i]);
    const uint64_t h = d.key ^ (d.tag * 0x9E3779B97F4A7C15ULL);
    auto it = seen.find(h);
    if (it != seen.end() && dev[it->second].key == d.key && dev[it->second].tag == d.tag) dev[it->second] = d;
    else { seen[h] = dev.size(); dev.push_back(d); }
  }
  if (dev.empty()) return 0;
  CK(cudaDeviceSynchronize());
  if (drain_pending(t)) return -1;
  gub::DevItem* d_items = nullptr;
  CK(cudaMalloc(&d_items, dev.size() * sizeof(gub::DevItem)));
  cudaError_t e = cudaMemcpy(d_items, dev.data(), dev.size() * sizeof(gub::DevItem), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(t->d_scalar, 0, 8);
  uint32_t failed = 0;
  if (e == cudaSuccess) {
    gub::k_add_items<<<(unsigned)((dev.size() + 255) / 256), 256>>>(t->table, t->capacity, d_items, (uint32_t)dev.size(), t->counters,
                                                                    reinterpret_cast<uint32_t*>(t->d_scalar), t->inv);
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(&failed, t->d_scalar, 4, cudaMemcpyDeviceToHost);
  cudaFree(d_items);
  if (e != cudaSuccess) return fail(std::string("gub_add_items: ") + cudaGetErrorString(e));
  if (failed) return fail("gub_add_items: table full for " + std::to_string(failed) + " items");
  return 0;
}

int gub_get_items(gub_table* t, const uint64_t* kx, const uint64_t* kf, size_t n, int64_t now_ms, gub_item* out, uint8_t* found) {
  if (!t || (n && (!kx || !kf || !out || !found))) return fail("gub_get_items: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  if (drain_pending(t)) return -1;
  uint64_t *d_kx = nullptr, *d_kf = nullptr; gub::DevItem* d_out = nullptr; uint8_t* d_found = nullptr;
  std::vector<gub::DevItem> host(n);
  cudaError_t e = cudaMalloc(&d_kx, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_kf, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, n * sizeof(gub::DevItem));
  if (e == cudaSuccess) e = cudaMalloc(&d_found, n);
  if (e == cudaSuccess) e = cudaMemcpy(d_kx, kx, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_kf, kf, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    gub::k_get_items<<<(unsigned)((n + 255) / 256), 256>>>(t->table, t->capacity, d_kx, d_kf, (uint32_t)n, now_ms, d_out, d_found, t->inv);
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(host.data(), d_out, n * sizeof(gub::DevItem), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(found, d_found, n, cudaMemcpyDeviceToHost);
  cudaFree(d_kx); cudaFree(d_kf); cudaFree(d_out); cudaFree(d_found);
  if (e != cudaSuccess) return fail(std::string("gub_get_items: ") + cudaGetErrorString(e));
  for (size_t i = 0; i < n; i++) { out[i] = from_dev(host[i]); out[i].key_xxh64 = kx[i]; out[i].key_fnv1 = kf[i]; }
  return 0;
}

int gub_scan(gub_table* t, gub_item* out, size_t cap, size_t* n_out) {
  if (!t || !n_out || (cap && !out)) return fail("gub_scan: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  if (drain_pending(t)) return -1;
  gub::DevItem* d_out = nullptr;
  if (cap) CK(cudaMalloc(&d_out, cap * sizeof(gub::DevItem)));
  unsigned long long total = 0;
  cudaError_t e = cudaMemset(t->d_scalar, 0, 8);
  if (e == cudaSuccess) {
    gub::k_scan<<<148 * 8, 256>>>(t->table, t->capacity, d_out, (unsigned long long)cap, t->d_scalar, t->inv);
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(&total, t->d_scalar, 8, cudaMemcpyDeviceToHost);
  std::vector<gub::DevItem> host(std::min<size_t>(cap, (size_t)total));
  if (e == cudaSuccess && !host.empty()) e = cudaMemcpy(host.data(), d_out, host.size() * sizeof(gub::DevItem), cudaMemcpyDeviceToHost);
  if (d_out) cudaFree(d_out);
  if (e != cudaSuccess) return fail(std::string("gub_scan: ") + cudaGetErrorString(e));
  for (size_t i = 0; i < host.size(); i++) out[i] = from_dev(host[i]);
  *n_out = (size_t)total;
  return 0;
}

int gub_size(gub_table* t, size_t* n_out) { return gub_scan(t, nullptr, 0, n_out); }

int gub_sweep(gub_table* t, int64_t now_ms, size_t* removed) {
  if (!t) return fail("gub_sweep: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(t->d_scalar, 0, 8));
  gub::k_sweep<<<148 * 8, 256>>>(t->table, (uint64_t)0, t->capacity, now_ms, t->d_scalar);
  CK(cudaDeviceSynchronize());
  unsigned long long r = 0;
  CK(cudaMemcpy(&r, t->d_scalar, 8, cudaMemcpyDeviceToHost));
  if (removed) *removed = (size_t)r;
  return 0;
}

int gub_get_counters(gub_table* t, gub_counters* out) {
  if (!t || !out) return fail("gub_get_counters: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  unsigned long long c[gub::C_COUNT];
  CK(cudaMemcpy(c, t->counters, sizeof c, cudaMemcpyDeviceToHost));
  out->over_limit = c[gub::C_OVER]; out->cache_hit = c[gub::C_HIT]; out->cache_miss = c[gub::C_MISS];
  out->inserts = c[gub::C_INSERTS]; out->table_full = c[gub::C_FULL]; out->requests = c[gub::C_REQUESTS];
  out->batches = c[gub::C_BATCHES]; out->dup_groups = c[gub::C_DUP_GROUPS]; out->mixed_groups = c[gub::C_MIXED_GROUPS];
  out->serial_fallbacks = c[gub::C_SERIAL];
  out->unexpired_evictions = c[gub::C_EVICT_UNEXPIRED]; out->swept = c[gub::C_SWEPT]; out->gq_dropped = c[gub::C_GQ_DROPPED];
  return 0;
}

int gub_hash_keys_device(gub_table* t, const char* d_bytes, const uint64_t* d_offsets, size_t n, uint64_t* d_xxh64_out, uint64_t* d_fnv1_out,
                         gub_req* d_reqs_out, void* stream) {
  if (!t || (n && (!d_bytes || !d_offsets))) return fail("gub_hash_keys_device: null argument");
  if (n == 0) return 0;
  CK(cudaSetDevice(t->device));
  gub::k_hash_keys<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint8_t*>(d_bytes), d_offsets, (uint32_t)n,
                                                                                 d_xxh64_out, d_fnv1_out, d_reqs_out);
  CK(cudaGetLastError());
  return 0;
}

/* Incremental expiry sweep inside the batch kernel: every CTA frees the removed / expired entries of `slots_per_cta` slots per batch
 * (0 = off).  The default visits the whole table once every ~65536 batches. */
int gub_set_sweep(gub_table* t, uint32_t slots_per_cta) {
  if (!t) return fail("gub_set_sweep: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  t->sweep_chunk = std::min<uint32_t>(slots_per_cta, 1024u);
  return 0;
}

/* Diagnostic: per-CTA %globaltimer stamps at the phase boundaries of k_batch.  gub_get_trace reports, for the LAST launch, the
 * time (us since the earliest CTA entered the kernel) at which the last CTA passed each of the 12 marks, and the mean over CTAs. */
int gub_set_trace(gub_table* t, int on) {
  if (!t) return fail("gub_set_trace: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  if (on && !t->trace) {
    CK(cudaMalloc(&t->trace, (size_t)gub::FB_MAX_GRID * gub::FB_TRACE_MARKS * 8));
    CK(cudaMemset(t->trace, 0, (size_t)gub::FB_MAX_GRID * gub::FB_TRACE_MARKS * 8));
    CK(cudaMalloc(&t->ktrace, (size_t)4 * gub::KT_BLOCKS * gub::KT_MARKS * 8));
    CK(cudaMemset(t->ktrace, 0, (size_t)4 * gub::KT_BLOCKS * gub::KT_MARKS * 8));
  } else if (!on && t->trace) {
    cudaFree(t->trace);
    t->trace = nullptr;
    if (t->ktrace) cudaFree(t->ktrace);
    t->ktrace = nullptr;
  }
  return 0;
}

/* Diagnostic: the pipeline kernels' time stamps since the last reset: out[(kernel * 1024 + block) * 8 + mark] = latest %globaltimer
 * (ns) at which a warp of the block — or, for the role marks, a thread in that role — passed the mark; 0 = never.  kernel:
 * 0 k_group, 1 k_rank, 2 k_eval, 3 k_finish. */
int gub_get_ktrace(gub_table* t, uint64_t* out /* 4 x 1024 x 8 */, int reset) {
  if (!t || !out) return fail("gub_get_ktrace: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  if (!t->ktrace) return fail("gub_get_ktrace: tracing is off");
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  const size_t bytes = (size_t)4 * gub::KT_BLOCKS * gub::KT_MARKS * 8;
  CK(cudaMemcpy(out, t->ktrace, bytes, cudaMemcpyDeviceToHost));
  if (reset) CK(cudaMemset(t->ktrace, 0, bytes));
  return 0;
}
int gub_get_trace(gub_table* t, double* max_us /* 12 */, double* mean_us /* 12 */) {
  if (!t || !max_us || !mean_us) return fail("gub_get_trace: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  if (!t->trace) return fail("gub_get_trace: tracing is off");
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> h((size_t)gub::FB_MAX_GRID * gub::FB_TRACE_MARKS);
  CK(cudaMemcpy(h.data(), t->trace, h.size() * 8, cudaMemcpyDeviceToHost));
  unsigned long long t0 = ~0ull;
  int ctas = 0;
  for (int b = 0; b < gub::FB_MAX_GRID; b++) if (h[(size_t)b * gub::FB_TRACE_MARKS]) { t0 = std::min(t0, h[(size_t)b * gub::FB_TRACE_MARKS]); ctas++; }
  for (int k = 0; k < gub::FB_TRACE_MARKS; k++) {
    double mx = 0, sum = 0;
    for (int b = 0; b < gub::FB_MAX_GRID; b++) {
      const unsigned long long v = h[(size_t)b * gub::FB_TRACE_MARKS + k];
      if (!h[(size_t)b * gub::FB_TRACE_MARKS] || v < t0) continue;
      const double us = (double)(v - t0) * 1e-3;
      mx = std::max(mx, us); sum += us;
    }
    max_us[k] = mx; mean_us[k] = ctas ? sum / ctas : 0;
  }
  return 0;
}

/* Raw stamps of the last launch: out[cta * 12 + mark] in ns (0 = CTA not launched). */
int gub_get_trace_raw(gub_table* t, uint64_t* out /* 256 x 12 */) {
  if (!t || !out) return fail("gub_get_trace_raw: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  if (!t->trace) return fail("gub_get_trace_raw: tracing is off");
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out, t->trace, (size_t)gub::FB_MAX_GRID * gub::FB_TRACE_MARKS * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int gub_set_profiling(gub_table* t, int on) {
  if (!t) return fail("gub_set_profiling: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  if (prof_flush(t, true)) return -1;
  t->prof = on != 0;
  return 0;
}

int gub_get_profile(gub_table* t, double kernel_ms[4], uint64_t* launches, int reset) {
  if (!t || !kernel_ms || !launches) return fail("gub_get_profile: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  if (prof_flush(t, true)) return -1;
  for (int k = 0; k < 4; k++) kernel_ms[k] = t->prof_ms[k];
  *launches = t->prof_launches;
  if (reset) { for (int k = 0; k < 4; k++) t->prof_ms[k] = 0; t->prof_launches = 0; }
  return 0;
}

// Measures the random 64-byte read-modify-write rate of this device over the table itself (contents unchanged): the
// "HBM random access" ceiling the batch path is compared with.  accesses are spread over all slots; returns GB/s moved
// (64 B read + 64 B written per access) in *gbs.
int gub_probe_random_access(gub_table* t, uint64_t accesses, double* gbs) {
  if (!t || !gbs || accesses == 0) return fail("gub_probe_random_access: bad argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const uint64_t zero = (uint64_t)(getenv("GUB_PROBE_NONZERO") != nullptr);  // always 0 in practice; opaque to the compiler
  gub::k_random_rmw<<<148 * 16, 256>>>(t->table, t->capacity, accesses / 8 + 1, 1, zero);  // warm-up
  CK(cudaEventRecord(e0, 0));
  gub::k_random_rmw<<<148 * 16, 256>>>(t->table, t->capacity, accesses, 12345, zero);
  CK(cudaEventRecord(e1, 0));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  CK(cudaGetLastError());
  *gbs = ms > 0 ? (double)accesses * 128.0 / (ms * 1e-3) * 1e-9 : 0.0;
  return 0;
}

// ---- multi-GPU routing ----------------------------------------------------------------------------------------
static int ensure_ring(gub_table* t, const gub_ring* ring) {
  const uint64_t ver = gub_ring_version_(ring);
  if (t->ring_cached == ring && t->ring_version == ver && t->d_ring_pts) return 0;
  const size_t npts = gub_ring_points(ring, nullptr, nullptr, 0);
  if (npts == 0) return fail("gub_route_device: ring is empty");
  if (gub_ring_size(ring) > gub::MAX_SHARDS) return fail("gub_route_device: at most 16 shards");
  std::vector<uint64_t> hs(npts); std::vector<int32_t> ps(npts);
  gub_ring_points(ring, hs.data(), ps.data(), npts);
  CK(cudaDeviceSynchronize());
  if (t->d_ring_pts) cudaFree(t->d_ring_pts);
  if (t->d_ring_peers) cudaFree(t->d_ring_peers);
  t->d_ring_pts = nullptr; t->d_ring_peers = nullptr;
  CK(cudaMalloc(&t->d_ring_pts, npts * 8));
  CK(cudaMalloc(&t->d_ring_peers, npts * 4));
  CK(cudaMemcpy(t->d_ring_pts, hs.data(), npts * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(t->d_ring_peers, ps.data(), npts * 4, cudaMemcpyHostToDevice));
  {  // direct-mapped entry into the sorted points by the hash's top 16 bits: first point >= (bucket << 48)
    if (npts > 0xFFFFu) return fail("gub_route_device: ring has too many points");
    std::vector<uint16_t> lut(65536);
    size_t k = 0;
    for (uint32_t b = 0; b < 65536; b++) {
      const uint64_t lo = (uint64_t)b << 48;
      while (k < npts && hs[k] < lo) k++;
      lut[b] = (uint16_t)k;
    }
    if (t->d_ring_lut) cudaFree(t->d_ring_lut);
    t->d_ring_lut = nullptr;
    CK(cudaMalloc(&t->d_ring_lut, 65536 * 2));
    CK(cudaMemcpy(t->d_ring_lut, lut.data(), 65536 * 2, cudaMemcpyHostToDevice));
  }
  t->ring_npts = (uint32_t)npts; t->ring_cached = ring; t->ring_version = ver;
  return 0;
}

static int route_impl(gub_table* t, const gub_ring* ring, int32_t self, const gub_req* d_reqs, size_t n, gub_req* d_out_reqs, uint32_t* d_perm,
                      uint32_t* d_counts, uint8_t* d_true_owner, void* stream) {
  if (!t || !ring || !d_counts || (n && (!d_reqs || !d_out_reqs || !d_perm))) return fail("gub_route_device: null argument");
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  if (ensure_ring(t, ring)) return -1;
  const uint32_t nshards = (uint32_t)gub_ring_size(ring);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { CK(cudaMemsetAsync(d_counts, 0, nshards * 4, st)); return 0; }
  if (n > 0x7FFFFFFFull) return fail("gub_route_device: batch too large");
  const uint32_t ntiles = (uint32_t)((n + gub::ROUTE_TILE - 1) / gub::ROUTE_TILE);
  if (!t->d_owner || t->owner_cap < n) { if (t->d_owner) { CK(cudaDeviceSynchronize()); cudaFree(t->d_owner); } CK(cudaMalloc(&t->d_owner, n)); t->owner_cap = n; }
  const size_t tc = (size_t)ntiles * gub::MAX_SHARDS;
  if (!t->d_tile_counts || t->tiles_cap < tc) { if (t->d_tile_counts) { CK(cudaDeviceSynchronize()); cudaFree(t->d_tile_counts); } CK(cudaMalloc(&t->d_tile_counts, tc * 4)); t->tiles_cap = tc; }
  gub::k_route_count<<<ntiles, 256, 0, st>>>(d_reqs, (uint32_t)n, t->d_ring_pts, t->d_ring_peers, t->ring_npts, nshards, t->d_owner,
                                            t->d_tile_counts, ntiles, self, d_true_owner);
  gub::k_route_scan<<<1, 1024, 0, st>>>(t->d_tile_counts, nshards * ntiles, nshards, ntiles, d_counts);
  gub::k_route_scatter<<<ntiles, 256, 0, st>>>(d_reqs, (uint32_t)n, t->d_owner, t->d_tile_counts, ntiles, nshards, d_out_reqs, d_perm);
  CK(cudaGetLastError());
  return 0;
}

int gub_route_device(gub_table* t, const gub_ring* ring, const gub_req* d_reqs, size_t n, gub_req* d_out_reqs, uint32_t* d_perm,
                     uint32_t* d_counts, void* stream) {
  return route_impl(t, ring, -1, d_reqs, n, d_out_reqs, d_perm, d_counts, nullptr, stream);
}

int gub_route_global_device(gub_table* t, const gub_ring* ring, uint32_t self, const gub_req* d_reqs, size_t n, gub_req* d_out_reqs,
                            uint32_t* d_perm, uint32_t* d_counts, uint8_t* d_owner_out, void* stream) {
  if (ring && self >= (uint32_t)gub_ring_size(ring)) return fail("gub_route_global_device: self is not a shard of the ring");
  return route_impl(t, ring, (int32_t)self, d_reqs, n, d_out_reqs, d_perm, d_counts, d_owner_out, stream);
}

__global__ void k_owner_only(const gub_req* reqs, uint32_t n, const uint64_t* pts, const int32_t* peers, uint32_t npts, uint8_t* owner) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) owner[i] = (uint8_t)gub::ring_owner(pts, peers, npts, __ldg(&reqs[i].key_fnv1));
}

int gub_route_owner_device(gub_table* t, const gub_ring* ring, const gub_req* d_reqs, size_t n, uint8_t* d_owner, void* stream) {
  if (!t || !ring || (n && (!d_reqs || !d_owner))) return fail("gub_route_owner_device: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  if (ensure_ring(t, ring)) return -1;
  k_owner_only<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_reqs, (uint32_t)n, t->d_ring_pts, t->d_ring_peers, t->ring_npts, d_owner);
  CK(cudaGetLastError());
  return 0;
}

// ---- GLOBAL behaviour: device queues ---------------------------------------------------------------------------
}  // extern "C"

struct gub_gq {
  int device = 0;
  gub::Gq q{};
  uint32_t* slot_of = nullptr;
  size_t slot_cap = 0;
};

namespace {
int gq_reserve(gub_gq* g, size_t n, cudaStream_t st) {
  if (g->slot_cap >= n) return 0;
  CK(cudaStreamSynchronize(st));
  if (g->slot_of) cudaFree(g->slot_of);
  g->slot_of = nullptr;
  CK(cudaMalloc(&g->slot_of, n * 4));
  g->slot_cap = n;
  return 0;
}
// n_dev (optional): the batch size lives on the device (<= n)
int gq_accumulate(gub_gq* g, const gub_req* d_reqs, size_t n, const uint32_t* n_dev, const uint8_t* d_owner, uint32_t self, uint64_t seq_base, cudaStream_t st) {
  if (n == 0) return 0;
  if (gq_reserve(g, n, st)) return -1;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  gub::k_gq_claim<<<blocks, 256, 0, st>>>(g->q, d_reqs, (uint32_t)n, n_dev, d_owner, self, d_owner ? 1u : 0u, (unsigned long long)seq_base, g->slot_of);
  gub::k_gq_fill<<<blocks, 256, 0, st>>>(g->q, d_reqs, (uint32_t)n, n_dev, (unsigned long long)seq_base, g->slot_of);
  CK(cudaGetLastError());
  return 0;
}
}  // namespace

extern "C" {

int gub_gq_create(int device, uint32_t capacity, int keep_latest, gub_gq** out) {
  if (!out || capacity < 64) return fail("gub_gq_create: bad argument");
  CK(cudaSetDevice(device));
  gub_gq* g = new gub_gq();
  g->device = device;
  const uint32_t cap = next_pow2(capacity);
  g->q.capacity_mask = cap - 1;
  g->q.mode = keep_latest ? gub::GQ_KEEP_LAST : gub::GQ_KEEP_FIRST;
  cudaError_t e = cudaMalloc(&g->q.slots, (size_t)cap * sizeof(gub_req));
  if (e == cudaSuccess) e = cudaMalloc(&g->q.seq, (size_t)cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&g->q.count, 16);
  if (e == cudaSuccess) e = cudaMemset(g->q.count, 0, 16);
  if (e == cudaSuccess) g->q.dropped = g->q.count + 1;
  if (e == cudaSuccess) e = cudaMemset(g->q.slots, 0, (size_t)cap * sizeof(gub_req));
  if (e == cudaSuccess) e = cudaMemset(g->q.seq, 0, (size_t)cap * 8);
  if (e == cudaSuccess) e = cudaMemset(g->q.count, 0, 8);
  if (e != cudaSuccess) { gub_gq_destroy(g); return fail(std::string("gub_gq_create: ") + cudaGetErrorString(e)); }
  *out = g;
  return 0;
}

void gub_gq_destroy(gub_gq* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  if (g->q.slots) cudaFree(g->q.slots);
  if (g->q.seq) cudaFree(g->q.seq);
  if (g->q.count) cudaFree(g->q.count);
  if (g->slot_of) cudaFree(g->slot_of);
  delete g;
}

int gub_gq_accumulate_device(gub_gq* g, const gub_req* d_reqs, size_t n, const uint8_t* d_owner, uint32_t self, uint64_t seq_base, void* stream) {
  if (!g || (n && !d_reqs)) return fail("gub_gq_accumulate_device: null argument");
  if (n == 0) return 0;
  CK(cudaSetDevice(g->device));
  return gq_accumulate(g, d_reqs, n, nullptr, d_owner, self, seq_base, (cudaStream_t)stream);
}

int gub_gq_dropped(gub_gq* g, uint64_t* dropped) {
  if (!g || !dropped) return fail("gub_gq_dropped: null argument");
  CK(cudaSetDevice(g->device));
  unsigned long long v = 0;
  CK(cudaMemcpy(&v, g->q.dropped, 8, cudaMemcpyDeviceToHost));
  *dropped = v;
  return 0;
}

int gub_gq_drain_device(gub_gq* g, gub_req* d_out, size_t cap, uint32_t* d_count, int as_status_query, void* stream) {
  if (!g || !d_count || (cap && !d_out)) return fail("gub_gq_drain_device: null argument");
  CK(cudaSetDevice(g->device));
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(d_count, 0, 4, st));
  gub::k_gq_drain<<<148, 256, 0, st>>>(g->q, d_out, (uint32_t)std::min<size_t>(cap, 0xFFFFFFFFu), d_count, as_status_query ? 1u : 0u);
  CK(cudaGetLastError());
  return 0;
}

int gub_make_updates_device(gub_table* t, const gub_req* d_queries, const gub_resp* d_resps, size_t n, gub_item* d_items, uint32_t* d_count,
                            void* stream) {
  if (!t || !d_count || (n && (!d_queries || !d_resps || !d_items))) return fail("gub_make_updates_device: null argument");
  CK(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(d_count, 0, 4, st));
  if (n == 0) return 0;
  gub::k_make_updates<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_queries, d_resps, (uint32_t)n, nullptr, d_items, d_count);
  CK(cudaGetLastError());
  return 0;
}

int gub_add_items_device(gub_table* t, const gub_item* d_items, size_t n, int64_t now_ms, void* stream) {
  if (!t || (n && !d_items)) return fail("gub_add_items_device: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (order_after_last(t, st)) return -1;
  gub::k_add_items_pub<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->table, t->capacity, d_items, (uint32_t)n, now_ms, t->counters, t->inv);
  CK(cudaGetLastError());
  t->last_stream = st; t->last_pending = true;
  return 0;
}

int gub_unroute_device(gub_table* t, const gub_resp* d_resp_in, const uint32_t* d_perm, size_t n, gub_resp* d_resp_out, void* stream) {
  if (!t || (n && (!d_resp_in || !d_perm || !d_resp_out))) return fail("gub_unroute_device: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  gub::k_unroute<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_resp_in, d_perm, (uint32_t)n, d_resp_out);
  CK(cudaGetLastError());
  return 0;
}


// ---- fused routing over peer memory -----------------------------------------------------------------------------
}  // extern "C"

extern "C" void gub_nccl_comm_destroy_(void* comm);
struct gub_p2p;
namespace {
int gq_accumulate_segments(gub_p2p* p, const gub::P2PArgs& A, cudaStream_t st);
}

struct gub_p2p {
  gub_table* t = nullptr;
  const gub_ring* ring = nullptr;
  uint32_t world = 0, rank = 0, cap = 0, epoch = 0;
  void* block = nullptr;           // our peer-visible allocation
  size_t block_bytes = 0;
  gub::P2PView views[gub::MAX_SHARDS];
  void* opened[gub::MAX_SHARDS] = {};
  bool connected = false;
  uint32_t* error = nullptr;       // device flag: a bounded wait gave up (a peer died)
  uint32_t* ticket = nullptr;      // [2] tile ticket / tiles done of the routing kernel
  // gub_p2p_step_streams: evaluation and collect run on streams of our own, so that neither the caller's stream nor the next step's
  // evaluation ever queues behind a wait for the slowest peer
  cudaStream_t s_eval = nullptr, s_collect = nullptr;
  cudaEvent_t ev_fork = nullptr;   // scratch event for one-off stream-to-stream ordering
  cudaStream_t last_ingest = nullptr;
  bool own_streams_used = false;
  uint32_t* seg_off = nullptr;     // [MAX_SHARDS + 1] prefix sums of this step's mailbox segments (k_seg_wait -> the batch kernels)
  // Routing scratch, preallocated (nothing is allocated or freed inside a step) and double-buffered by step parity: with a
  // separate ingest stream the routing of step e+1 runs while step e is still being evaluated and collected.
  struct Route {
    unsigned long long* tile_agg = nullptr;  // [(cap / RT_THREADS + 1) * MAX_SHARDS]
    uint32_t* counts = nullptr;    // [MAX_SHARDS]
    uint32_t* perm = nullptr;      // [cap]
    uint8_t* true_owner = nullptr; // [cap] (GLOBAL mode)
    cudaEvent_t routed = nullptr;     // this parity's routing has been issued (ingest stream)
    cudaEvent_t step_done = nullptr;  // this parity's collect has completed (evaluation stream)
    bool step_done_valid = false;
  } rt[2];
  // GLOBAL behaviour (gub_p2p_enable_global): device queues + tick buffers
  gub_gq *hits_q = nullptr, *updates_q = nullptr;
  uint32_t gcap = 0;
  gub_req *g_reqs = nullptr; gub_resp* g_resps = nullptr; gub_item *g_items = nullptr, *g_gather = nullptr;
  uint32_t* g_count = nullptr;     // [4] device counters (hits drained, queries drained, items made, spare)
  uint32_t* g_counts_all = nullptr;  // [MAX_SHARDS] device: items per rank
  uint64_t seq = 0;
  // all shards in one process (gub_p2p_connect_local): peers by pointer + a host-side rendezvous for the GLOBAL tick
  struct LocalGroup { std::mutex mu; std::condition_variable cv; uint32_t arrived = 0, gen = 0; } own_group;
  LocalGroup* group = nullptr;
  gub_p2p* local_peers[gub::MAX_SHARDS] = {};
  cudaEvent_t phase_ev = nullptr;  // single-thread drivers: 'this shard's phase has been enqueued and completed' (see local_phase_sync)
  uint32_t* h_counts = nullptr;    // pinned: {hits drained, queries drained, items made, spare} of the tick in flight
  void* nccl = nullptr;            // ncclComm_t
  uint64_t tick_bytes = 0;         // bytes all-gathered by the last tick
};

namespace {
size_t p2p_req_bytes(uint32_t W, uint32_t cap) { return (size_t)2 * W * cap * sizeof(gub_req); }
size_t p2p_resp_bytes(uint32_t W, uint32_t cap) { return (size_t)2 * W * cap * sizeof(gub_resp); }
gub::P2PView p2p_view(void* base, uint32_t W, uint32_t cap) {
  gub::P2PView v;
  char* b = static_cast<char*>(base);
  v.req_mb = reinterpret_cast<gub_req*>(b);
  v.resp_mb = reinterpret_cast<gub_resp*>(b + p2p_req_bytes(W, cap));
  v.req_flag = reinterpret_cast<unsigned long long*>(b + p2p_req_bytes(W, cap) + p2p_resp_bytes(W, cap));
  v.resp_flag = v.req_flag + 2 * W;
  return v;
}

// A bounded device-side wait gave up during an earlier step (a peer died or hung): surfaced here, at the next call, without
// adding a host round trip to the step itself.  The flag is host-mapped... it is device memory; reading it costs a sync, so
// callers poll it with gub_p2p_status() at their own cadence.
}  // namespace

extern "C" {

void gub_p2p_destroy(gub_p2p* p) {
  if (!p) return;
  cudaSetDevice(p->t->device);
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < p->world; r++) if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
  void* ptrs[] = {p->block, p->error, p->ticket, p->seg_off, p->g_reqs, p->g_resps, p->g_items, p->g_gather, p->g_count, p->g_counts_all};
  for (void* q : ptrs) if (q) cudaFree(q);
  if (p->h_counts) cudaFreeHost(p->h_counts);
  if (p->phase_ev) cudaEventDestroy(p->phase_ev);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->s_eval) cudaStreamDestroy(p->s_eval);
  if (p->s_collect) cudaStreamDestroy(p->s_collect);
  for (auto& r : p->rt) {
    void* rp[] = {r.tile_agg, r.counts, r.perm, r.true_owner};
    for (void* q : rp) if (q) cudaFree(q);
    if (r.routed) cudaEventDestroy(r.routed);
    if (r.step_done) cudaEventDestroy(r.step_done);
  }
  if (p->hits_q) gub_gq_destroy(p->hits_q);
  if (p->updates_q) gub_gq_destroy(p->updates_q);
  gub_nccl_comm_destroy_(p->nccl);
  delete p;
}

int gub_p2p_create(gub_table* t, const gub_ring* ring, uint32_t rank, uint32_t cap, gub_p2p** out) {
  const uint32_t world = ring ? (uint32_t)gub_ring_size(ring) : 0;
  if (!t || !out || world == 0 || world > (uint32_t)gub::MAX_SHARDS || rank >= world || cap == 0 || cap >= (1u << 24))
    return fail("gub_p2p_create: bad argument");
  CK(cudaSetDevice(t->device));
  {
    std::lock_guard<std::mutex> lk(t->mu);
    if (ensure_ring(t, ring)) return -1;  // upload the ring now: a step never allocates or synchronises the device
  }
  if (!t->fused) {
    // An owner evaluates what all `world` sources send it in a step: up to world x cap records, normally about cap.  Size the
    // pipeline's scratch so that a step is one pass where it can be (<= 262 144 requests per pass; more passes beyond that).
    const uint32_t want = (uint32_t)std::min<uint64_t>((uint64_t)world * cap, 262144u);
    std::lock_guard<std::mutex> lk(t->mu);
    if (t->max_batch < want) {
      CK(cudaDeviceSynchronize());
      if (alloc_scratch(t, want)) return -1;
    }
  }
  gub_p2p* p = new gub_p2p();
  p->t = t; p->ring = ring; p->world = world; p->rank = rank; p->cap = cap;
  p->block_bytes = p2p_req_bytes(world, cap) + p2p_resp_bytes(world, cap) + (size_t)4 * world * 8;
  cudaError_t e = cudaMalloc(&p->block, p->block_bytes);
  if (e == cudaSuccess) e = cudaMemset(p->block, 0, p->block_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->error, 4);
  if (e == cudaSuccess) e = cudaMemset(p->error, 0, 4);
  if (e == cudaSuccess) e = cudaMalloc(&p->ticket, 8);
  if (e == cudaSuccess) e = cudaMemset(p->ticket, 0, 8);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_eval, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_collect, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&p->seg_off, (gub::MAX_SHARDS + 1) * 4);
  if (e == cudaSuccess) e = cudaMemset(p->seg_off, 0, (gub::MAX_SHARDS + 1) * 4);
  const size_t agg = ((size_t)cap / gub::RT_THREADS + 1) * gub::MAX_SHARDS;
  for (auto& r : p->rt) {
    if (e == cudaSuccess) e = cudaMalloc(&r.tile_agg, agg * 8);
    if (e == cudaSuccess) e = cudaMemset(r.tile_agg, 0, agg * 8);
    if (e == cudaSuccess) e = cudaMalloc(&r.counts, gub::MAX_SHARDS * 4);
    if (e == cudaSuccess) e = cudaMemset(r.counts, 0, gub::MAX_SHARDS * 4);
    if (e == cudaSuccess) e = cudaMalloc(&r.perm, (size_t)cap * 4);
    if (e == cudaSuccess) e = cudaMalloc(&r.true_owner, (size_t)cap);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.routed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.step_done, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) { gub_p2p_destroy(p); return fail(std::string("gub_p2p_create: ") + cudaGetErrorString(e)); }
  for (uint32_t r = 0; r < world; r++) p->views[r] = p2p_view(p->block, world, cap);  // until connected: everything loops back
  CK(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

int gub_p2p_export(gub_p2p* p, void* handle_out) {
  if (!p || !handle_out) return fail("gub_p2p_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == GUB_P2P_HANDLE_BYTES, "handle size");
  CK(cudaSetDevice(p->t->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, p->block));
  std::memcpy(handle_out, &h, sizeof h);
  return 0;
}

int gub_p2p_connect(gub_p2p* p, const void* handles) {
  if (!p || !handles) return fail("gub_p2p_connect: null argument");
  CK(cudaSetDevice(p->t->device));
  for (uint32_t r = 0; r < p->world; r++) {
    if (r == p->rank) { p->views[r] = p2p_view(p->block, p->world, p->cap); continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * GUB_P2P_HANDLE_BYTES, sizeof h);
    void* base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    p->opened[r] = base;
    p->views[r] = p2p_view(base, p->world, p->cap);
  }
  p->connected = true;
  return 0;
}

// All shards in this process (the reference daemon is one process: daemon.go:73): the mailboxes are shared by pointer, and
// shards on different devices get peer access to each other's memory enabled here.
int gub_p2p_connect_local(gub_p2p* p, gub_p2p* const* peers) {
  if (!p || !peers) return fail("gub_p2p_connect_local: null argument");
  CK(cudaSetDevice(p->t->device));
  for (uint32_t r = 0; r < p->world; r++) {
    if (!peers[r] || peers[r]->world != p->world || peers[r]->cap != p->cap) return fail("gub_p2p_connect_local: mismatched peer");
    const int dev = peers[r]->t->device;
    if (dev != p->t->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, p->t->device, dev));
      if (!can) return fail("gub_p2p_connect_local: device " + std::to_string(p->t->device) + " cannot access device " + std::to_string(dev));
      cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    }
    p->views[r] = p2p_view(peers[r]->block, p->world, p->cap);
    p->local_peers[r] = peers[r];
  }
  p->group = &peers[0]->own_group;
  p->connected = true;
  return 0;
}

int gub_p2p_status(gub_p2p* p, int* error_out) {
  if (!p || !error_out) return fail("gub_p2p_status: null argument");
  CK(cudaSetDevice(p->t->device));
  uint32_t e1 = 0, e2 = 0;
  CK(cudaMemcpy(&e1, p->error, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&e2, &p->t->ctl->error, 4, cudaMemcpyDeviceToHost));
  *error_out = (int)(e1 | e2);
  if (e1 | e2) return fail("gub_p2p: a bounded device-side wait gave up (peer flag / grid barrier): a shard of the ring is not answering");
  return 0;
}

}  // extern "C"

namespace {

void p2p_args(gub_p2p* p, gub::P2PArgs& A) {
  for (uint32_t r = 0; r < p->world; r++) A.peers[r] = p->views[r];
  A.world = p->world; A.rank = p->rank; A.cap = p->cap; A.epoch = p->epoch; A.done_ctr = nullptr; A.error = p->error;
}

// The three launches of a step.  d_n (optional): the ingest batch's size lives on the device (<= n).  self_global >= 0: GLOBAL
// requests this shard does not own are answered here (and their true owners recorded for the hits queue).
int p2p_route(gub_p2p* p, gub_p2p::Route* rt, const gub::P2PArgs& A, const gub_req* d_reqs, uint32_t n, const uint32_t* d_n, int self_global, cudaStream_t si) {
  gub_table* t = p->t;
  gub::RouteArgs R;
  R.P = A; R.reqs = d_reqs; R.n = n; R.n_dev = d_n; R.pts = t->d_ring_pts; R.pt_peer = t->d_ring_peers; R.lut = t->d_ring_lut; R.npts = t->ring_npts;
  R.self_global = self_global; R.true_owner = self_global >= 0 ? rt->true_owner : nullptr; R.tile_agg = rt->tile_agg; R.counts = rt->counts; R.perm = rt->perm;
  R.ticket = p->ticket;
  const uint32_t tiles = std::max<uint32_t>(1u, (n + gub::RT_THREADS - 1) / gub::RT_THREADS);
  gub::k_p2p_route<<<tiles, gub::RT_THREADS, 0, si>>>(R);
  CK(cudaGetLastError());
  return 0;
}

// The owner side of a step.  Default: the four-kernel pipeline reading the mailbox segments in place — k_seg_wait (one warp waits
// for the sources' flags and writes the segment offsets), the batch kernels in ring mode (responses stored straight into the
// sources' response mailboxes), k_seg_publish (response flags).  GUB_PATH=fused: one launch of the persistent kernel instead.
int p2p_evaluate(gub_p2p* p, const gub::P2PArgs& A, const gub_clock* clk, cudaStream_t st) {
  gub_table* t = p->t;
  if (!t->fused) {
    const uint32_t par = p->epoch & 1u;
    SegDesc sd;
    sd.nseg = p->world; sd.seg_off = p->seg_off;
    for (uint32_t s = 0; s < p->world; s++) {
      sd.reqs[s] = A.peers[p->rank].req_mb + ((size_t)par * p->world + s) * p->cap;   // what source s stored into our mailbox
      sd.out[s] = A.peers[s].resp_mb + ((size_t)par * p->world + p->rank) * p->cap;   // NVLink stores into source s's response mailbox
    }
    gub::k_seg_wait<<<1, 32, 0, st>>>(A, p->seg_off);
    // GLOBAL requests evaluated here as owner go to the updates queue (gubernator.go:604-606, global.go:80-84).  The queue kernels
    // read the request mailboxes, so they run BEFORE the response flags go out: once a source has its answers it may route the
    // step after next into the same mailbox half.
    if (p->updates_q && gq_accumulate_segments(p, A, st)) return -1;
    const uint64_t total = (uint64_t)p->world * p->cap;
    for (uint64_t off = 0; off < total; off += t->max_batch) {
      const uint32_t m = (uint32_t)std::min<uint64_t>(t->max_batch, total - off);
      if (launch_chunk(t, nullptr, m, clk, nullptr, st, p->seg_off + p->world, (uint32_t)off, &sd)) return -1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (t->pdl && !t->prof) ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, gub::k_seg_publish, A));
    return maybe_sweep(t, clk, st);
  }
  gub::FArgs F;
  fused_base_args(t, clk, F);
  const uint32_t par = p->epoch & 1u;
  for (uint32_t s = 0; s < p->world; s++) {
    F.seg[s].reqs = A.peers[p->rank].req_mb + ((size_t)par * p->world + s) * p->cap;         // what source s stored into our mailbox
    F.seg[s].flag = &A.peers[p->rank].req_flag[(size_t)par * p->world + s];
    F.seg[s].out = A.peers[s].resp_mb + ((size_t)par * p->world + p->rank) * p->cap;         // NVLink stores into source s's response mailbox
    F.seg[s].n = p->cap;
    F.resp_flag[s] = &A.peers[s].resp_flag[(size_t)par * p->world + p->rank];
  }
  F.nseg = p->world; F.n_resp_flags = p->world; F.flag_epoch = p->epoch;
  return launch_fused(t, F, 0, st);
}

}  // namespace

extern "C" {

}  // extern "C"

namespace {

int p2p_step_check(gub_p2p* p, const gub_req* d_reqs, size_t n, const gub_clock* clk, gub_resp* d_out) {
  if (!p || !clk || (n && (!d_reqs || !d_out))) return fail("gub_p2p_step: null argument");
  if (n > p->cap) return fail("gub_p2p_step: n exceeds the mailbox capacity");
  if (p->t->ring_cached != p->ring || p->t->ring_version != gub_ring_version_(p->ring)) return fail("gub_p2p_step: the table's ring changed since gub_p2p_create");
  return 0;
}

// Phase 1 of a step, on the ingest stream: partition by owner and store the records into the owners' mailboxes.  This parity's
// scratch and mailbox halves were last used two steps ago; that step's collect (on the evaluation stream) must have finished,
// which also means every peer has drained what we sent it then.
int p2p_phase_route(gub_p2p* p, const gub_req* d_reqs, size_t n, cudaStream_t si, cudaStream_t st) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  const bool two = si != st;
  p->epoch++;
  gub_p2p::Route* rt = &p->rt[p->epoch & 1u];
  gub::P2PArgs A;
  p2p_args(p, A);
  if (two && rt->step_done_valid) CK(cudaStreamWaitEvent(si, rt->step_done, 0));
  const int self_global = p->hits_q ? (int)p->rank : -1;
  if (p2p_route(p, rt, A, d_reqs, (uint32_t)n, nullptr, self_global, si)) return -1;
  if (p->hits_q && n) {  // a non-owner queues the hits of its GLOBAL requests (gubernator.go:402-404, global.go:74-78)
    if (gq_accumulate(p->hits_q, d_reqs, n, nullptr, rt->true_owner, p->rank, p->seq, si)) return -1;
  }
  if (two) CK(cudaEventRecord(rt->routed, si));
  return 0;
}

// Phase 2, on the evaluation stream: the batch kernel synchronises with every source (ourselves included) through the mailbox
// flags, evaluates straight out of the mailboxes and stores the responses into the sources' response mailboxes.
int p2p_phase_evaluate(gub_p2p* p, const gub_clock* clk, cudaStream_t st) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  gub::P2PArgs A;
  p2p_args(p, A);
  if (order_after_last(t, st)) return -1;
  if (p2p_evaluate(p, A, clk, st)) return -1;
  t->last_stream = st; t->last_pending = true;
  if (p->updates_q && t->fused) {  // GLOBAL requests just evaluated as owner (gubernator.go:604-606, global.go:80-84); the pipeline path queues them itself
    if (gq_accumulate_segments(p, A, st)) return -1;
  }
  return 0;
}

// Phase 3: wait for the owners' flags, responses back in request order.
int p2p_phase_collect(gub_p2p* p, size_t n, gub_resp* d_out, cudaStream_t si, cudaStream_t st) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  const bool two = si != st;
  gub_p2p::Route* rt = &p->rt[p->epoch & 1u];
  gub::P2PArgs A;
  p2p_args(p, A);
  if (two) CK(cudaStreamWaitEvent(st, rt->routed, 0));  // the collect reads this parity's permutation
  gub::k_p2p_collect<<<std::max<unsigned>(1u, std::min<unsigned>(148u, (unsigned)((n + 255) / 256))), 256, 0, st>>>(A, rt->perm, (uint32_t)n, nullptr, d_out);
  if (two) { CK(cudaEventRecord(rt->step_done, st)); rt->step_done_valid = true; }
  CK(cudaGetLastError());
  p->seq += (uint64_t)1 << 32;
  return 0;
}

}  // namespace

extern "C" {

int gub_p2p_step_streams(gub_p2p* p, const gub_req* d_reqs, size_t n, const gub_clock* clk, gub_resp* d_out, void* ingest_stream, void* stream) {
  if (p2p_step_check(p, d_reqs, n, clk, d_out)) return -1;
  cudaStream_t st = (cudaStream_t)stream, si = (cudaStream_t)ingest_stream;
  // Default: routing on the ingest stream, evaluation and collect on `stream`.  GUB_RING_STREAMS=3 (experimental) moves the evaluation
  // and the collect — which waits for the slowest owner — onto streams of the ring's own, so that step e+1 is evaluated while step e's
  // answers travel; on 2 x B200 that variant ran the timed loop but stalled under the per-kernel profiling events (profiles/README.md).
  static const int own_streams = [] { const char* e = getenv("GUB_RING_STREAMS"); return e ? std::atoi(e) : 2; }();
  if (si != st && own_streams < 3) {
    if (p2p_phase_route(p, d_reqs, n, si, st)) return -1;
    if (p2p_phase_evaluate(p, clk, st)) return -1;
    return p2p_phase_collect(p, n, d_out, si, st);
  }
  if (si == st) {  // one stream: the three phases in stream order
    if (p2p_phase_route(p, d_reqs, n, si, st)) return -1;
    if (p2p_phase_evaluate(p, clk, st)) return -1;
    return p2p_phase_collect(p, n, d_out, si, st);
  }
  // Two caller streams: routing on the ingest stream; the owner-side evaluation on our own evaluation stream (it synchronises with
  // the peers through the mailbox flags, not through the caller's streams); the collect — which waits for the slowest owner — on our
  // collect stream; `stream` then waits for the collect only.  Step e+1 is routed and evaluated while step e's answers travel.
  if (p2p_phase_route(p, d_reqs, n, si, p->s_collect)) return -1;
  if (p2p_phase_evaluate(p, clk, p->s_eval)) return -1;
  CK(cudaEventRecord(p->ev_fork, st));                 // d_out is ours to write once everything queued on `stream` so far is done
  CK(cudaStreamWaitEvent(p->s_collect, p->ev_fork, 0));
  if (p2p_phase_collect(p, n, d_out, si, p->s_collect)) return -1;
  CK(cudaStreamWaitEvent(st, p->rt[p->epoch & 1u].step_done, 0));
  p->last_ingest = si; p->own_streams_used = true;
  return 0;
}

}  // extern "C"

namespace {
// Between the phases of the single-thread drivers: every shard's stream waits for the phase just enqueued on every other shard's
// stream.  The kernels of the next phase then find their mailbox flags already published and never spin: on a device shared by
// several shards a spinning kernel could keep another shard's (cooperative, all-SM) batch kernel from ever becoming resident.
int local_phase_sync(gub_p2p* const* ps, uint32_t world, void* const* streams) {
  if (world < 2) return 0;
  for (uint32_t r = 0; r < world; r++) {
    CK(cudaSetDevice(ps[r]->t->device));
    if (!ps[r]->phase_ev) CK(cudaEventCreateWithFlags(&ps[r]->phase_ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ps[r]->phase_ev, (cudaStream_t)streams[r]));
  }
  for (uint32_t r = 0; r < world; r++) {
    CK(cudaSetDevice(ps[r]->t->device));
    for (uint32_t q = 0; q < world; q++) if (q != r) CK(cudaStreamWaitEvent((cudaStream_t)streams[r], ps[q]->phase_ev, 0));
  }
  return 0;
}
}  // namespace

extern "C" {

/* One step of every shard of this process, driven by ONE host thread (the shape of the reference daemon: one process): the
 * phases are enqueued shard by shard — all routings, then all evaluations, then all collects — so that no shard's kernels wait
 * for work the host has not enqueued yet (on ONE device that would deadlock: the batch kernel occupies every SM). */
int gub_p2p_step_local_all(gub_p2p* const* ps, uint32_t world, const gub_req* const* d_reqs, const size_t* n, const gub_clock* clk,
                           gub_resp* const* d_out, void* const* streams) {
  if (!ps || !d_reqs || !n || !d_out || !streams || world == 0) return fail("gub_p2p_step_local_all: bad argument");
  for (uint32_t r = 0; r < world; r++) if (p2p_step_check(ps[r], d_reqs[r], n[r], clk, d_out[r])) return -1;
  for (uint32_t r = 0; r < world; r++) if (p2p_phase_route(ps[r], d_reqs[r], n[r], (cudaStream_t)streams[r], (cudaStream_t)streams[r])) return -1;
  if (local_phase_sync(ps, world, streams)) return -1;
  for (uint32_t r = 0; r < world; r++) if (p2p_phase_evaluate(ps[r], clk, (cudaStream_t)streams[r])) return -1;
  if (local_phase_sync(ps, world, streams)) return -1;
  for (uint32_t r = 0; r < world; r++) if (p2p_phase_collect(ps[r], n[r], d_out[r], (cudaStream_t)streams[r], (cudaStream_t)streams[r])) return -1;
  return 0;
}

int gub_p2p_step(gub_p2p* p, const gub_req* d_reqs, size_t n, const gub_clock* clk, gub_resp* d_out, void* stream) {
  return gub_p2p_step_streams(p, d_reqs, n, clk, d_out, stream, stream);
}

}  // extern "C"


// ---- GLOBAL behaviour behind the C ABI: queues fed by the step, and the sync tick (global.go:91-283) with NCCL -----------------
namespace {

int gq_accumulate_segments(gub_p2p* p, const gub::P2PArgs& A, cudaStream_t st) {
  gub_gq* g = p->updates_q;
  const size_t total = (size_t)p->world * p->cap;
  if (gq_reserve(g, total, st)) return -1;
  gub::GqSegs G;
  const uint32_t par = p->epoch & 1u;
  for (uint32_t s = 0; s < p->world; s++) {
    G.reqs[s] = A.peers[p->rank].req_mb + ((size_t)par * p->world + s) * p->cap;
    G.flag[s] = &A.peers[p->rank].req_flag[(size_t)par * p->world + s];
  }
  G.nseg = p->world; G.cap = p->cap;
  gub::k_gq_claim_segs<<<148, 256, 0, st>>>(g->q, G, (unsigned long long)p->seq, g->slot_of);
  gub::k_gq_fill_segs<<<148, 256, 0, st>>>(g->q, G, (unsigned long long)p->seq, g->slot_of);
  CK(cudaGetLastError());
  return 0;
}

// NCCL is loaded at run time (libnccl.so.2, the one torch bundles when the host process is Python): a single-GPU deployment
// needs no NCCL at all.  Minimal prototypes of nccl.h (2.27): ncclResult_t is an int enum, ncclComm_t an opaque pointer.
struct Id128 { char b[128]; };
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /* ncclUniqueId by value: 128 bytes */ Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.ok) return 0;
  const char* names[] = {getenv("GUB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return fail("NCCL not found (dlopen libnccl.so.2): GLOBAL sync across GPUs needs it");
#define SYM(field, name)                                                              \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(g_nccl.h, name);                   \
  if (!g_nccl.field) return fail(std::string("NCCL symbol missing: ") + name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.ok = true;
  return 0;
}
#define NCK(call)                                                                                                  \
  do {                                                                                                             \
    int r__ = (call);                                                                                              \
    if (r__ != 0) return fail(std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error")); \
  } while (0)
}  // namespace

extern "C" {

void gub_nccl_comm_destroy_(void* comm) {
  if (comm && g_nccl.ok) g_nccl.CommDestroy(comm);
}

/* ncclGetUniqueId: 128 bytes the host application hands to every rank (its own channel: the Go daemon's peer discovery, torch.distributed, a file) */
int gub_nccl_unique_id(void* out128) {
  if (!out128) return fail("gub_nccl_unique_id: null argument");
  if (nccl_load()) return -1;
  NCK(g_nccl.GetUniqueId(out128));
  return 0;
}

/* One communicator per shard (rank = the shard's index in the ring).  Collective across the ring's processes. */
int gub_p2p_nccl_init(gub_p2p* p, const void* id128) {
  if (!p || !id128) return fail("gub_p2p_nccl_init: null argument");
  if (nccl_load()) return -1;
  CK(cudaSetDevice(p->t->device));
  Id128 id;
  std::memcpy(&id, id128, sizeof id);
  NCK(g_nccl.CommInitRank(&p->nccl, (int)p->world, id, (int)p->rank));
  return 0;
}

/* All shards in one process: the communicators are created inside one NCCL group. */
int gub_p2p_nccl_init_local(gub_p2p* const* ps, uint32_t world) {
  if (!ps || world == 0) return fail("gub_p2p_nccl_init_local: bad argument");
  if (nccl_load()) return -1;
  Id128 id;
  NCK(g_nccl.GetUniqueId(&id));
  NCK(g_nccl.GroupStart());
  for (uint32_t r = 0; r < world; r++) {
    CK(cudaSetDevice(ps[r]->t->device));
    NCK(g_nccl.CommInitRank(&ps[r]->nccl, (int)world, id, (int)ps[r]->rank));
  }
  NCK(g_nccl.GroupEnd());
  return 0;
}

/* Turns GLOBAL handling on for this shard's steps: GLOBAL requests it does not own are answered from the local replica
 * (gubernator.go:257-269,408-411) and their hits queued (global.go:74-111); GLOBAL requests it evaluates as owner are queued for
 * the broadcast (global.go:80-84,201).  capacity = most distinct GLOBAL keys per sync window (drops are counted in gq_dropped). */
int gub_p2p_enable_global(gub_p2p* p, uint32_t capacity) {
  if (!p || capacity < 64) return fail("gub_p2p_enable_global: bad argument");
  if (p->hits_q) return 0;
  CK(cudaSetDevice(p->t->device));
  if (gub_gq_create(p->t->device, capacity, 0, &p->hits_q)) return -1;
  if (gub_gq_create(p->t->device, capacity, 1, &p->updates_q)) return -1;
  p->hits_q->q.dropped = p->t->counters + gub::C_GQ_DROPPED;
  p->updates_q->q.dropped = p->t->counters + gub::C_GQ_DROPPED;
  p->gcap = std::min<uint32_t>(next_pow2(capacity), p->cap);
  CK(cudaMalloc(&p->g_reqs, (size_t)p->gcap * sizeof(gub_req)));
  CK(cudaMalloc(&p->g_resps, (size_t)p->gcap * sizeof(gub_resp)));
  CK(cudaMalloc(&p->g_items, (size_t)p->gcap * sizeof(gub_item)));
  CK(cudaMalloc(&p->g_gather, (size_t)p->gcap * p->world * sizeof(gub_item)));
  CK(cudaMalloc(&p->g_count, 16));
  CK(cudaMemset(p->g_count, 0, 16));
  CK(cudaMalloc(&p->g_counts_all, gub::MAX_SHARDS * 4));
  CK(cudaHostAlloc(&p->h_counts, 16, cudaHostAllocDefault));
  std::memset(p->h_counts, 0, 16);
  if (gq_reserve(p->updates_q, (size_t)p->world * p->cap, 0)) return -1;
  if (gq_reserve(p->hits_q, p->cap, 0)) return -1;
  return 0;
}

/* One GLOBAL sync: sendHits (global.go:144-190), then broadcastPeers (global.go:234-283): the owners' UpdatePeerGlobal items reach
 * every other shard, which installs them (UpdatePeerGlobals, gubernator.go:425-459).  Transport of the broadcast: an NCCL
 * all-gather when the shard has a communicator (one process per GPU); direct peer reads when all shards live in this process
 * (gub_p2p_connect_local), where the calling threads — one per shard — rendezvous on the host.  Collective: every shard of the
 * ring calls it the same number of times, interleaved the same way with its steps.  now_ms = MillisecondNow() of the receivers
 * (CreatedAt / UpdatedAt of the installed items).  The one host round trip is the per-shard item counts (a tick runs every
 * GlobalSyncWait = 500 ms, not per batch).
 * stats (optional, 4 x uint64): hit records sent to owners, update items broadcast by this shard, items installed here, bytes moved. */
}  // extern "C"

namespace {

// A tick in three enqueue-only phases (no host synchronisation inside), so that one host thread can drive all the shards of a
// process phase by phase (see gub_p2p_step_local_all).
// A: the window's aggregated hits -> their owners' mailboxes (plain routing: these records go to the owner).
int tick_phase_a(gub_p2p* p, cudaStream_t st) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  if (p->own_streams_used) {  // the steps' queue updates and evaluations ran on other streams: the tick comes after all of them
    cudaStream_t others[3] = {p->last_ingest, p->s_eval, p->s_collect};
    for (cudaStream_t o : others) {
      if (o == st) continue;
      CK(cudaEventRecord(p->ev_fork, o));
      CK(cudaStreamWaitEvent(st, p->ev_fork, 0));
    }
  }
  CK(cudaMemsetAsync(p->g_count, 0, 16, st));
  gub::k_gq_drain<<<148, 256, 0, st>>>(p->hits_q->q, p->g_reqs, p->gcap, p->g_count + 0, 0u);
  p->epoch++;
  gub_p2p::Route* rt = &p->rt[p->epoch & 1u];
  if (rt->step_done_valid) CK(cudaStreamWaitEvent(st, rt->step_done, 0));
  gub::P2PArgs A;
  p2p_args(p, A);
  return p2p_route(p, rt, A, p->g_reqs, p->gcap, p->g_count + 0, -1, st);
}
// B: owners apply the hits with DRAIN_OVER_LIMIT as owner (gubernator.go:510-512).
int tick_phase_b(gub_p2p* p, const gub_clock* clk, cudaStream_t st) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  gub::P2PArgs A;
  p2p_args(p, A);
  if (order_after_last(t, st)) return -1;
  if (p2p_evaluate(p, A, clk, st)) return -1;
  t->last_stream = st; t->last_pending = true;
  if (t->fused && gq_accumulate_segments(p, A, st)) return -1;
  CK(cudaGetLastError());
  return 0;
}
// C: owners re-read every key touched by GLOBAL traffic with Hits = 0 (global.go:243-245) and build the UpdatePeerGlobal items.
int tick_phase_c(gub_p2p* p, const gub_clock* clk, cudaStream_t st) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  {  // the owners' answers to our hit records are dropped (sendHits ignores them), but the flags are consumed: epochs stay aligned.
     // (Enqueued here, after EVERY shard's evaluation: on a device shared by several shards a collect spinning ahead of another
     // shard's batch kernel would keep that cooperative kernel from ever becoming resident.)
    gub_p2p::Route* rt = &p->rt[p->epoch & 1u];
    gub::P2PArgs A;
    p2p_args(p, A);
    gub::k_p2p_collect<<<64, 256, 0, st>>>(A, rt->perm, p->gcap, p->g_count + 0, p->g_resps);
    CK(cudaEventRecord(rt->step_done, st)); rt->step_done_valid = true;
    p->seq += (uint64_t)1 << 32;
  }
  gub::k_gq_drain<<<148, 256, 0, st>>>(p->updates_q->q, p->g_reqs, p->gcap, p->g_count + 1, 1u);
  if (launch_batch(t, p->g_reqs, p->gcap, clk, p->g_resps, st, p->g_count + 1)) return -1;  // (either kernel path; the count lives on the device)
  gub::k_make_updates<<<(p->gcap + 255) / 256, 256, 0, st>>>(p->g_reqs, p->g_resps, p->gcap, p->g_count + 1, p->g_items, p->g_count + 2);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(p->h_counts, p->g_count, 16, cudaMemcpyDeviceToHost, st));
  return 0;
}
int tick_enqueue(gub_p2p* p, const gub_clock* clk, cudaStream_t st) {
  if (tick_phase_a(p, st)) return -1;
  if (tick_phase_b(p, clk, st)) return -1;
  return tick_phase_c(p, clk, st);
}

void group_barrier(gub_p2p::LocalGroup* g, uint32_t world) {
  std::unique_lock<std::mutex> lk(g->mu);
  const uint32_t my = g->gen;
  if (++g->arrived == world) { g->arrived = 0; g->gen++; g->cv.notify_all(); }
  else g->cv.wait(lk, [&] { return g->gen != my; });
}

// Installs the items every other shard of this process has made (their g_items, read in place: same device or peer memory).
int tick_install_local(gub_p2p* p, int64_t now_ms, cudaStream_t st, uint64_t* installed) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  p->tick_bytes = 0;
  for (uint32_t r = 0; r < p->world; r++) {
    if (r == p->rank) continue;  // "Exclude ourselves from the update" (global.go:263-265)
    const gub_p2p* q = p->local_peers[r];
    const uint32_t k = std::min(q->h_counts[2], q->gcap);
    if (!k) continue;
    gub::k_add_items_pub<<<(k + 255) / 256, 256, 0, st>>>(t->table, t->capacity, q->g_items, k, now_ms, t->counters, t->inv);
    *installed += k;
    p->tick_bytes += (uint64_t)k * sizeof(gub_item);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));  // the peers may overwrite their items at their next tick
  return 0;
}

int tick_install_nccl(gub_p2p* p, int64_t now_ms, cudaStream_t st, uint64_t* installed) {
  gub_table* t = p->t;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  // all-gather: counts first (the host needs them to size the item gather), then the items, padded to the largest count
  std::vector<uint32_t> all(p->world, 0);
  NCK(g_nccl.AllGather(p->g_count + 2, p->g_counts_all, 1, /* ncclUint32 */ 3, p->nccl, st));
  CK(cudaMemcpyAsync(all.data(), p->g_counts_all, p->world * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  uint32_t pad = 0;
  for (uint32_t r = 0; r < p->world; r++) pad = std::max(pad, std::min(all[r], p->gcap));
  p->tick_bytes = 0;
  if (pad) {
    NCK(g_nccl.AllGather(p->g_items, p->g_gather, (size_t)pad * sizeof(gub_item), /* ncclUint8 */ 1, p->nccl, st));
    p->tick_bytes = (uint64_t)pad * sizeof(gub_item) * p->world;
    for (uint32_t r = 0; r < p->world; r++) {
      const uint32_t k = std::min(all[r], p->gcap);
      if (r == p->rank || !k) continue;  // "Exclude ourselves from the update" (global.go:263-265)
      gub::k_add_items_pub<<<(k + 255) / 256, 256, 0, st>>>(t->table, t->capacity, p->g_gather + (size_t)r * pad, k, now_ms, t->counters, t->inv);
      *installed += k;
    }
    CK(cudaGetLastError());
  }
  return 0;
}

void tick_stats(gub_p2p* p, uint64_t installed, uint64_t* stats) {
  if (!stats) return;
  stats[0] = std::min(p->h_counts[0], p->gcap); stats[1] = std::min(p->h_counts[2], p->gcap); stats[2] = installed; stats[3] = p->tick_bytes;
}

}  // namespace

extern "C" {

int gub_global_tick(gub_p2p* p, const gub_clock* clk, int64_t now_ms, void* stream, uint64_t* stats) {
  if (!p || !clk) return fail("gub_global_tick: null argument");
  if (!p->hits_q) return fail("gub_global_tick: gub_p2p_enable_global has not been called");
  if (p->world > 1 && !p->nccl && !p->group) return fail("gub_global_tick: no NCCL communicator (gub_p2p_nccl_init) and the shards are not local to this process");
  cudaStream_t st = (cudaStream_t)stream;
  if (tick_enqueue(p, clk, st)) return -1;
  uint64_t installed = 0;
  if (p->world > 1 && p->nccl) {
    if (tick_install_nccl(p, now_ms, st, &installed)) return -1;
    CK(cudaStreamSynchronize(st));
  } else {
    CK(cudaSetDevice(p->t->device));
    CK(cudaStreamSynchronize(st));  // my items and their count are final
    if (p->world > 1) {
      group_barrier(p->group, p->world);  // ... and so are everybody else's
      const int rc = tick_install_local(p, now_ms, st, &installed);
      group_barrier(p->group, p->world);  // nobody still reads my items
      if (rc) return -1;
    }
  }
  tick_stats(p, installed, stats);
  return 0;
}

/* The same for a single host thread driving all the shards of this process (gub_p2p_connect_local): phases run shard by shard. */
int gub_global_tick_local_all(gub_p2p* const* ps, uint32_t world, const gub_clock* clk, int64_t now_ms, void* const* streams, uint64_t* stats /* world x 4, optional */) {
  if (!ps || !clk || !streams || world == 0) return fail("gub_global_tick_local_all: bad argument");
  for (uint32_t r = 0; r < world; r++) {
    if (!ps[r] || !ps[r]->hits_q || ps[r]->world != world || (world > 1 && !ps[r]->group)) return fail("gub_global_tick_local_all: shards must be local, connected and GLOBAL-enabled");
  }
  for (uint32_t r = 0; r < world; r++) if (tick_phase_a(ps[r], (cudaStream_t)streams[r])) return -1;
  if (local_phase_sync(ps, world, streams)) return -1;
  for (uint32_t r = 0; r < world; r++) if (tick_phase_b(ps[r], clk, (cudaStream_t)streams[r])) return -1;
  if (local_phase_sync(ps, world, streams)) return -1;
  for (uint32_t r = 0; r < world; r++) if (tick_phase_c(ps[r], clk, (cudaStream_t)streams[r])) return -1;
  for (uint32_t r = 0; r < world; r++) { CK(cudaSetDevice(ps[r]->t->device)); CK(cudaStreamSynchronize((cudaStream_t)streams[r])); }
  for (uint32_t r = 0; r < world; r++) {
    uint64_t installed = 0;
    if (world > 1 && tick_install_local(ps[r], now_ms, (cudaStream_t)streams[r], &installed)) return -1;
    tick_stats(ps[r], installed, stats ? stats + 4 * r : nullptr);
  }
  return 0;
}

}  // extern "C"
