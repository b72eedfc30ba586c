// gub_api.cu — the C ABI (include/gubernator_b200.h) over the kernels in gub_kernels.cuh.
//
// This is the layer that sits where gubernator's WorkerPool sits (workers.go:54-61): gub_create = NewWorkerPool,
// gub_submit* = GetRateLimit for a whole batch, gub_add_items = AddCacheItem/Load, gub_scan = Store, gub_destroy = Close.
// There is no CPU fallback: every entry point that evaluates requests launches CUDA kernels or fails.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
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
  // per-batch scratch, two sets: k_group/k_rank of batch b+1 (prep stream) overlap k_eval/k_finish of batch b
  struct Scratch {
    gub::AuxEntry* aux = nullptr;
    uint32_t *ent = nullptr, *meta = nullptr, *rank = nullptr, *order = nullptr, *mixed_ent = nullptr, *presence = nullptr;
    uint8_t* fragsize = nullptr;
    ulonglong2* commit = nullptr;
    uint32_t* commit_ent = nullptr;
    gub::BatchCtr* ctr = nullptr;
    uint32_t epoch = 0;
    cudaEvent_t prep_done = nullptr;   // stage 1 (k_group, k_rank) finished on the prep stream
    cudaEvent_t eval_done = nullptr;   // stage 2 (k_eval, k_finish) finished: the set may be reused
    bool used = false;
  } scr[2];
  uint32_t next_set = 0;
  uint32_t aux_entries = 0, pres_words = 0, max_blocks = 0;
  // the fused batch kernel (gub_batch.cuh): one scratch set, one cooperative launch per batch
  bool fused = true;                  // GUB_FUSED=0: the four-kernel path (kept for A/B measurements)
  bool coop = true;                   // cooperative launch (co-residency of the grid guaranteed by the driver)
  int num_sms = 0;
  uint32_t sweep_chunk = 0;           // slots every CTA sweeps per batch (incremental expiry sweep), 0 = off
  gub::GEntry* gaux = nullptr;
  uint32_t *gpres = nullptr, *gpos = nullptr, *ordbuf = nullptr;
  uint16_t* gfrag = nullptr;
  gub::FCtl* ctl = nullptr;
  gub::OvfItem* ovf = nullptr;
  unsigned long long* counters = nullptr;
  cudaStream_t s_prep = nullptr;
  cudaEvent_t inputs_ready = nullptr;
  bool overlap = true;                // GUB_OVERLAP=0: everything on the caller's stream
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
  uint64_t* d_ring_pts = nullptr; int32_t* d_ring_peers = nullptr; uint32_t ring_npts = 0; const gub_ring* ring_cached = nullptr; uint64_t ring_version = 0;
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
cudaError_t launch_k(gub_table* t, K kernel, uint32_t grid, uint32_t block, cudaStream_t st, const gub::BatchArgs& A) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = t->pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, A);
}

int launch_finish(gub_table* t, const gub::BatchArgs& A, uint32_t n, cudaStream_t st) {
  // commit records (one thread each, at most n/2) + non-uniform groups (one block each, grid-stride; normally none)
  const uint32_t mixed_blocks = std::min<uint32_t>(148u, std::max<uint32_t>(1u, n / 2));
  const uint32_t commit_blocks = std::max<uint32_t>(1u, std::min<uint32_t>(148u, (n / 2 + gub::MIXED_THREADS - 1) / gub::MIXED_THREADS));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(mixed_blocks + commit_blocks); cfg.blockDim = dim3(gub::MIXED_THREADS); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = t->pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, gub::k_finish, A, mixed_blocks));
  return 0;
}

void fused_base_args(gub_table* t, const gub_clock* clk, gub::FArgs& A) {
  std::memset(&A, 0, sizeof A);
  A.table = t->table; A.capacity = t->capacity; A.aux = t->gaux; A.presence = t->gpres; A.fragsize = t->gfrag; A.gpos = t->gpos; A.ordbuf = t->ordbuf;
  A.ctl = t->ctl; A.ovf = t->ovf; A.counters = t->counters; A.sweep_chunk = t->sweep_chunk; A.clk = *clk;
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

// One batch (<= max_batch requests).  Stage 1 (k_group, k_rank) never touches bucket state, so it runs on the prep
// stream and overlaps stage 2 (k_eval, k_finish) of the previous batch, which runs on the caller's stream.
int launch_chunk(gub_table* t, const gub_req* d_reqs, uint32_t n, const gub_clock* clk, gub_resp* d_out, cudaStream_t st,
                 const uint32_t* n_dev = nullptr, uint32_t n_off = 0) {
  const bool overlap = t->overlap && !t->prof && !gub::EARLY_SINGLES;  // k_rank touches the table in the early-singles build
  gub_table::Scratch& sc = t->scr[overlap ? t->next_set : 0u];  // one set is enough when batches do not overlap
  if (overlap) t->next_set ^= 1u;
  cudaStream_t sp = overlap ? t->s_prep : st;
  if (overlap) {
    CK(cudaEventRecord(t->inputs_ready, st));          // whatever produced d_reqs on the caller's stream
    CK(cudaStreamWaitEvent(sp, t->inputs_ready, 0));
    if (sc.used) CK(cudaStreamWaitEvent(sp, sc.eval_done, 0));  // the batch that last used this scratch set is done with it
  }
  if (sc.epoch >= 65535u) {  // 16-bit epoch tags wrapped: clear the grouping table so stale tags cannot alias
    CK(cudaMemsetAsync(sc.aux, 0, (size_t)t->aux_entries * sizeof(gub::AuxEntry), sp));
    // 65535 -> 1 keeps the parity: the batch after the wrap reuses ctr[1], which k_rank (it only resets the OTHER parity) left
    // holding batch 65535's allocators.  Clear both.
    CK(cudaMemsetAsync(sc.ctr, 0, 2 * sizeof(gub::BatchCtr), sp));
    sc.epoch = 0;
  }
  sc.epoch++;
  gub::BatchArgs A;
  A.table = t->table; A.capacity = t->capacity; A.reqs = d_reqs; A.out = d_out; A.n = n; A.n_dev = n_dev; A.n_off = n_off; A.epoch = sc.epoch;
  A.aux = sc.aux; A.aux_mask = t->aux_entries - 1; A.presence = sc.presence; A.fragsize = sc.fragsize;
  A.pres_words = t->pres_words; A.max_blocks = t->max_blocks; A.ent = sc.ent; A.meta = sc.meta; A.rank = sc.rank;
  A.commit = sc.commit; A.commit_ent = sc.commit_ent; A.order = sc.order; A.mixed_ent = sc.mixed_ent; A.ctr = sc.ctr;
  A.counters = t->counters;
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
  CK(launch_k(t, gub::k_group, blocks, gub::GROUP_THREADS, sp, A));
  if (pe) CK(cudaEventRecord(pe[1], st));
  CK(launch_k(t, gub::k_rank, blocks, gub::GROUP_THREADS, sp, A));
  if (pe) CK(cudaEventRecord(pe[2], st));
  if (overlap) {
    CK(cudaEventRecord(sc.prep_done, sp));
    CK(cudaStreamWaitEvent(st, sc.prep_done, 0));
  }
  CK(launch_k(t, gub::k_eval, blocks, gub::GROUP_THREADS, st, A));
  if (pe) CK(cudaEventRecord(pe[3], st));
  if (launch_finish(t, A, n, st)) return -1;
  if (pe) CK(cudaEventRecord(pe[4], st));
  if (overlap) { CK(cudaEventRecord(sc.eval_done, st)); sc.used = true; }
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
  return it;
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
  void* ptrs[] = {t->table, t->counters, t->d_scalar, t->d_ring_pts, t->d_ring_peers, t->d_owner, t->d_tile_counts,
                  t->gaux, t->gpres, t->gfrag, t->gpos, t->ordbuf, t->ctl, t->ovf};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& sc : t->scr) {
    void* sp[] = {sc.aux, sc.ent, sc.meta, sc.rank, sc.order, sc.mixed_ent, sc.presence, sc.fragsize, sc.commit, sc.commit_ent, sc.ctr};
    for (void* p : sp) if (p) cudaFree(p);
    if (sc.prep_done) cudaEventDestroy(sc.prep_done);
    if (sc.eval_done) cudaEventDestroy(sc.eval_done);
  }
  if (t->s_prep) cudaStreamDestroy(t->s_prep);
  if (t->inputs_ready) cudaEventDestroy(t->inputs_ready);
  for (auto& s : t->pipe) {
    if (s.d_req) cudaFree(s.d_req);
    if (s.d_resp) cudaFree(s.d_resp);
    if (s.d_creq) cudaFree(s.d_creq);
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
  const uint32_t B = t->max_batch;
  t->aux_entries = next_pow2((uint64_t)B * 4);
  t->max_blocks = (B / gub::GROUP_THREADS + 31u) & ~31u;  // fragment-size row per group entry (multiple of 32 bytes)
  t->pres_words = t->max_blocks / 32;
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
  for (auto& sc : t->scr) {
    ALLOC(sc.aux, (size_t)t->aux_entries * sizeof(gub::AuxEntry));
    ALLOC(sc.presence, (size_t)t->aux_entries * t->pres_words * 4);
    ALLOC(sc.fragsize, (size_t)t->aux_entries * t->max_blocks);
    ALLOC(sc.commit, (size_t)t->aux_entries * 6 * sizeof(ulonglong2));
    ALLOC(sc.commit_ent, ((size_t)B / 2 + 1) * 4);
    ALLOC(sc.ent, (size_t)B * 4);
    ALLOC(sc.meta, (size_t)B * 4);
    ALLOC(sc.rank, (size_t)B * 4);
    ALLOC(sc.order, (size_t)B * 4);
    ALLOC(sc.mixed_ent, ((size_t)B / 2 + 1) * 4);
    ALLOC(sc.ctr, 2 * sizeof(gub::BatchCtr));
  }
  t->num_sms = std::min<int>(prop.multiProcessorCount, gub::FB_MAX_GRID);
  if (const char* e = getenv("GUB_FUSED")) t->fused = std::atoi(e) != 0;
  if (const char* e = getenv("GUB_COOP")) t->coop = std::atoi(e) != 0;
  ALLOC(t->gaux, (size_t)gub::FB_AUX_ENTRIES * sizeof(gub::GEntry));
  ALLOC(t->gpres, (size_t)gub::FB_AUX_ENTRIES * gub::FB_PRES_WORDS * 4);
  ALLOC(t->gfrag, (size_t)gub::FB_AUX_ENTRIES * gub::FB_ROW * 2);
  ALLOC(t->gpos, (size_t)gub::FB_MAX_GRID * gub::FB_THREADS * 4);
  ALLOC(t->ordbuf, (size_t)gub::FB_MAX_GRID * gub::FB_THREADS * 4);
  ALLOC(t->ctl, sizeof(gub::FCtl));
  ALLOC(t->ovf, (size_t)gub::FB_OVF_CAP * sizeof(gub::OvfItem));
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
  CK(cudaStreamCreateWithFlags(&t->s_prep, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&t->inputs_ready, cudaEventDisableTiming));
  for (auto& sc : t->scr) {
    CK(cudaEventCreateWithFlags(&sc.prep_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sc.eval_done, cudaEventDisableTiming));
  }
  if (const char* e = getenv("GUB_OVERLAP")) t->overlap = std::atoi(e) != 0;
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
  gub::DevItem* d_items = nullptr;
  CK(cudaMalloc(&d_items, dev.size() * sizeof(gub::DevItem)));
  cudaError_t e = cudaMemcpy(d_items, dev.data(), dev.size() * sizeof(gub::DevItem), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(t->d_scalar, 0, 8);
  uint32_t failed = 0;
  if (e == cudaSuccess) {
    gub::k_add_items<<<(unsigned)((dev.size() + 255) / 256), 256>>>(t->table, t->capacity, d_items, (uint32_t)dev.size(), t->counters,
                                                                    reinterpret_cast<uint32_t*>(t->d_scalar));
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
  uint64_t *d_kx = nullptr, *d_kf = nullptr; gub::DevItem* d_out = nullptr; uint8_t* d_found = nullptr;
  std::vector<gub::DevItem> host(n);
  cudaError_t e = cudaMalloc(&d_kx, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_kf, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, n * sizeof(gub::DevItem));
  if (e == cudaSuccess) e = cudaMalloc(&d_found, n);
  if (e == cudaSuccess) e = cudaMemcpy(d_kx, kx, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_kf, kf, n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    gub::k_get_items<<<(unsigned)((n + 255) / 256), 256>>>(t->table, t->capacity, d_kx, d_kf, (uint32_t)n, now_ms, d_out, d_found);
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
  gub::DevItem* d_out = nullptr;
  if (cap) CK(cudaMalloc(&d_out, cap * sizeof(gub::DevItem)));
  unsigned long long total = 0;
  cudaError_t e = cudaMemset(t->d_scalar, 0, 8);
  if (e == cudaSuccess) {
    gub::k_scan<<<148 * 8, 256>>>(t->table, t->capacity, d_out, (unsigned long long)cap, t->d_scalar);
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
  gub::k_sweep<<<148 * 8, 256>>>(t->table, t->capacity, now_ms, t->d_scalar);
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
  if (e == cudaSuccess) e = cudaMalloc(&g->q.count, 8);
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
  cudaStream_t st = (cudaStream_t)stream;
  if (g->slot_cap < n) {
    CK(cudaStreamSynchronize(st));
    if (g->slot_of) cudaFree(g->slot_of);
    g->slot_of = nullptr;
    CK(cudaMalloc(&g->slot_of, n * 4));
    g->slot_cap = n;
  }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  gub::k_gq_claim<<<blocks, 256, 0, st>>>(g->q, d_reqs, (uint32_t)n, d_owner, self, d_owner ? 1u : 0u, (unsigned long long)seq_base, g->slot_of);
  gub::k_gq_fill<<<blocks, 256, 0, st>>>(g->q, d_reqs, (uint32_t)n, (unsigned long long)seq_base, g->slot_of);
  CK(cudaGetLastError());
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
  gub::k_make_updates<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_queries, d_resps, (uint32_t)n, d_items, d_count);
  CK(cudaGetLastError());
  return 0;
}

__global__ void k_add_items_pub(gub::Slot* table, uint64_t cap, const gub_item* items, uint32_t n, int64_t now_ms, unsigned long long* counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const gub_item it = items[i];
  if (it.algorithm != GUB_TOKEN_BUCKET && it.algorithm != GUB_LEAKY_BUCKET) return;
  gub::Cursor cur;
  const uint64_t key = gub::remap_key(it.key_xxh64), tag = it.key_fnv1 >> 8;
  gub::cursor_open(cur, table, cap, key, tag);
  const bool leaky = it.algorithm == GUB_LEAKY_BUCKET;
  cur.b.key = key; cur.b.tag = tag;
  cur.b.flags = gub::F_LIVE | (leaky ? gub::F_LEAKY : 0u) | ((!leaky && it.status == GUB_OVER_LIMIT) ? gub::F_OVER : 0u);
  cur.b.limit = it.limit; cur.b.duration = it.duration;
  cur.b.rem = leaky ? gub::f2bits(it.remaining_f) : (uint64_t)it.remaining;
  cur.b.stamp = now_ms; cur.b.burst = leaky ? it.burst : 0; cur.b.expire = it.expire_at;
  uint32_t ins = 0;
  if (!gub::cursor_close(cur, table, cap, ins)) atomicAdd(counters + gub::C_FULL, 1ull);
  if (ins) atomicAdd(counters + gub::C_INSERTS, (unsigned long long)ins);
}

int gub_add_items_device(gub_table* t, const gub_item* d_items, size_t n, int64_t now_ms, void* stream) {
  if (!t || (n && !d_items)) return fail("gub_add_items_device: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(t->mu);
  CK(cudaSetDevice(t->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (order_after_last(t, st)) return -1;
  k_add_items_pub<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->table, t->capacity, d_items, (uint32_t)n, now_ms, t->counters);
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

struct gub_p2p {
  gub_table* t = nullptr;
  const gub_ring* ring = nullptr;
  uint32_t world = 0, rank = 0, cap = 0, epoch = 0;
  void* block = nullptr;           // our peer-visible allocation
  size_t block_bytes = 0;
  gub::P2PView views[gub::MAX_SHARDS];
  void* opened[gub::MAX_SHARDS] = {};
  bool connected = false;
  gub_req* inbox = nullptr; gub_resp* inbox_resp = nullptr; size_t inbox_cap = 0;
  uint32_t *seg_off = nullptr, *m_dev = nullptr, *done_ctr = nullptr, *error = nullptr;
  // Routing scratch, preallocated (nothing is allocated or freed inside a step) and double-buffered by step parity: with a
  // separate ingest stream the routing of step e+1 runs while step e is still being evaluated and un-routed.
  struct Route {
    uint8_t* owner = nullptr;      // [cap]
    uint32_t* tile_off = nullptr;  // [(cap / ROUTE_TILE + 1) * MAX_SHARDS]
    uint32_t* counts = nullptr;    // [MAX_SHARDS]
    uint32_t* perm = nullptr;      // [cap]
    cudaEvent_t routed = nullptr;     // this parity's scatter has been issued and completed (ingest stream)
    cudaEvent_t step_done = nullptr;  // this parity's un-route has completed (evaluation stream)
    bool step_done_valid = false;
  } rt[2];
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
}  // namespace

extern "C" {

void gub_p2p_destroy(gub_p2p* p) {
  if (!p) return;
  cudaSetDevice(p->t->device);
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < p->world; r++) if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
  void* ptrs[] = {p->block, p->inbox, p->inbox_resp, p->seg_off, p->m_dev, p->done_ctr, p->error};
  for (void* q : ptrs) if (q) cudaFree(q);
  for (auto& r : p->rt) {
    void* rp[] = {r.owner, r.tile_off, r.counts, r.perm};
    for (void* q : rp) if (q) cudaFree(q);
    if (r.routed) cudaEventDestroy(r.routed);
    if (r.step_done) cudaEventDestroy(r.step_done);
  }
  delete p;
}

int gub_p2p_create(gub_table* t, const gub_ring* ring, uint32_t rank, uint32_t cap, gub_p2p** out) {
  const uint32_t world = ring ? (uint32_t)gub_ring_size(ring) : 0;
  if (!t || !out || world == 0 || world > (uint32_t)gub::MAX_SHARDS || rank >= world || cap == 0) return fail("gub_p2p_create: bad argument");
  CK(cudaSetDevice(t->device));
  {
    std::lock_guard<std::mutex> lk(t->mu);
    if (ensure_ring(t, ring)) return -1;  // upload the ring now: a step never allocates or synchronises the device
  }
  gub_p2p* p = new gub_p2p();
  p->t = t; p->ring = ring; p->world = world; p->rank = rank; p->cap = cap;
  p->block_bytes = p2p_req_bytes(world, cap) + p2p_resp_bytes(world, cap) + (size_t)4 * world * 8;
  cudaError_t e = cudaMalloc(&p->block, p->block_bytes);
  if (e == cudaSuccess) e = cudaMemset(p->block, 0, p->block_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->seg_off, (gub::MAX_SHARDS + 1) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&p->m_dev, 4);
  if (e == cudaSuccess) e = cudaMalloc(&p->done_ctr, 8);
  if (e == cudaSuccess) e = cudaMalloc(&p->error, 4);
  if (e == cudaSuccess) e = cudaMemset(p->done_ctr, 0, 8);
  if (e == cudaSuccess) e = cudaMemset(p->error, 0, 4);
  for (auto& r : p->rt) {
    if (e == cudaSuccess) e = cudaMalloc(&r.owner, cap);
    if (e == cudaSuccess) e = cudaMalloc(&r.tile_off, ((size_t)cap / gub::ROUTE_TILE + 1) * gub::MAX_SHARDS * 4);
    if (e == cudaSuccess) e = cudaMalloc(&r.counts, gub::MAX_SHARDS * 4);
    if (e == cudaSuccess) e = cudaMalloc(&r.perm, (size_t)cap * 4);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.routed, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.step_done, cudaEventDisableTiming);
  }
  p->inbox_cap = (size_t)world * cap;  // worst case every shard sends us its whole batch
  if (e == cudaSuccess) e = cudaMalloc(&p->inbox, p->inbox_cap * sizeof(gub_req));
  if (e == cudaSuccess) e = cudaMalloc(&p->inbox_resp, p->inbox_cap * sizeof(gub_resp));
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

int gub_p2p_connect_local(gub_p2p* p, gub_p2p* const* peers) {
  if (!p || !peers) return fail("gub_p2p_connect_local: null argument");
  for (uint32_t r = 0; r < p->world; r++) {
    if (!peers[r] || peers[r]->world != p->world || peers[r]->cap != p->cap) return fail("gub_p2p_connect_local: mismatched peer");
    p->views[r] = p2p_view(peers[r]->block, p->world, p->cap);
  }
  p->connected = true;
  return 0;
}

int gub_p2p_step_streams(gub_p2p* p, const gub_req* d_reqs, size_t n, const gub_clock* clk, gub_resp* d_out, void* ingest_stream, void* stream) {
  if (!p || !clk || (n && (!d_reqs || !d_out))) return fail("gub_p2p_step: null argument");
  if (n > p->cap) return fail("gub_p2p_step: n exceeds the mailbox capacity");
  const gub_ring* ring = p->ring;
  if (p->t->ring_cached != ring || p->t->ring_version != gub_ring_version_(ring)) return fail("gub_p2p_step: the table's ring changed since gub_p2p_create");
  gub_table* t = p->t;
  cudaStream_t st = (cudaStream_t)stream, si = (cudaStream_t)ingest_stream;
  const bool two = si != st;
  uint32_t ntiles = 0;
  gub::P2PArgs A;
  gub_p2p::Route* rt = nullptr;
  {
    std::lock_guard<std::mutex> lk(t->mu);
    CK(cudaSetDevice(t->device));
    p->epoch++;
    rt = &p->rt[p->epoch & 1u];
    for (uint32_t r = 0; r < p->world; r++) A.peers[r] = p->views[r];
    A.world = p->world; A.rank = p->rank; A.cap = p->cap; A.epoch = p->epoch; A.done_ctr = p->done_ctr; A.error = p->error;
    // ---- ingest stream: partition by owner and store the records into the owners' mailboxes.  This parity's scratch and
    // mailbox halves were last used two steps ago; that step's un-route (on the evaluation stream) must have finished, which
    // also means every peer has drained what we sent it then.
    if (two && rt->step_done_valid) CK(cudaStreamWaitEvent(si, rt->step_done, 0));
    if (n) {
      ntiles = (uint32_t)((n + gub::ROUTE_TILE - 1) / gub::ROUTE_TILE);
      gub::k_route_count<<<ntiles, 256, 0, si>>>(d_reqs, (uint32_t)n, t->d_ring_pts, t->d_ring_peers, t->ring_npts, p->world, rt->owner,
                                                rt->tile_off, ntiles, -1, nullptr);
      gub::k_route_scan<<<1, 1024, 0, si>>>(rt->tile_off, p->world * ntiles, p->world, ntiles, rt->counts);
      gub::k_p2p_scatter<<<ntiles, 256, 0, si>>>(A, d_reqs, (uint32_t)n, rt->owner, rt->tile_off, ntiles, rt->counts, rt->perm);
    } else {
      gub::k_p2p_publish_empty<<<1, 32, 0, si>>>(A);
    }
    if (two) CK(cudaEventRecord(rt->routed, si));
    // ---- evaluation stream: the gather synchronises with every source (ourselves included) through the mailbox flags
    gub::k_p2p_gather<<<148, 256, 0, st>>>(A, p->inbox, p->seg_off, p->m_dev);
    CK(cudaGetLastError());
  }
  // How many records we own is only known on the device (p->m_dev): the evaluation launches are sized for what a shard may
  // receive at most and trim themselves, so the step needs no host round trip at all.
  if (gub_submit_device_n(t, p->inbox, (size_t)p->world * p->cap, p->m_dev, clk, p->inbox_resp, stream)) return -1;
  {
    std::lock_guard<std::mutex> lk(t->mu);
    gub::k_p2p_push_resp<<<148, 256, 0, st>>>(A, p->inbox_resp, p->seg_off);
    if (two) CK(cudaStreamWaitEvent(st, rt->routed, 0));  // the un-route reads this parity's offsets and permutation
    if (n) gub::k_p2p_unroute<<<148, 256, 0, st>>>(A, rt->tile_off, ntiles, rt->perm, (uint32_t)n, d_out);
    else gub::k_p2p_wait_resp_only<<<1, 32, 0, st>>>(A);
    if (two) { CK(cudaEventRecord(rt->step_done, st)); rt->step_done_valid = true; }
    CK(cudaGetLastError());
  }
  return 0;
}

int gub_p2p_step(gub_p2p* p, const gub_req* d_reqs, size_t n, const gub_clock* clk, gub_resp* d_out, void* stream) {
  return gub_p2p_step_streams(p, d_reqs, n, clk, d_out, stream, stream);
}

}  // extern "C"
