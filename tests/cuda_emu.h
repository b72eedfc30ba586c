// cuda_emu.h — TEST-ONLY: just enough of the CUDA execution model to run gubernator_b200/csrc/gub_kernels.cuh on a CPU.
//
// Why: the parity tests proper need a B200 (`-m gpu`); this lets the very same kernel source — grouping, ranks, snapshots,
// segment planning, routing, hashing — be checked against the oracle in the CPU suite, and lets a kernel change be tried
// without spending GPU minutes.  It is an emulation of the *programming model*, not of the hardware: no timing, one
// interleaving.
//
// Model: a launch runs its blocks one after another (launch_coop: all blocks of the grid at once, for kernels with grid-wide
// barriers; those must keep their shared memory in the dynamic allocation, emu::dyn_smem()); the threads of a block are ucontext fibers on one OS thread, run
// round-robin and switched only at synchronisation points (__syncthreads and the warp collectives).  Atomics are therefore
// plain read-modify-writes, `__shared__` is a function-local static (one block is alive at a time), and a thread that
// returns early leaves the barriers it would have joined, as on the device.  What this cannot show: data races, memory
// ordering, anything that needs two blocks or two launches to run concurrently (the peer-memory mailbox kernels spin on flags
// written by other launches: the harness runs them phase by phase, so the flags are already there).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __noinline__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 v; v.x = x; v.y = y; return v; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
using std::max;
using std::min;

namespace emu {

struct Sync { unsigned live = 0, arrived = 0, gen = 0; };
struct Warp {
  Sync sync;
  unsigned long long val[32];
  bool present[32];
};
struct Block {
  Sync sync;
  std::vector<Warp> warps;
  uint3 idx{0, 0, 0};
  void* dyn_smem = nullptr;
  unsigned or_acc = 0;
};
struct Fiber {
  ucontext_t ctx;
  uint3 tid;
  Block* blk = nullptr;
  bool done = false;
};

struct State {
  ucontext_t sched;
  Fiber* cur = nullptr;
  std::vector<Fiber> fibers;
  std::vector<void*> stacks;
  std::vector<Block> blocks;   // the blocks alive right now (one, or the whole grid for launch_coop)
  dim3 block_dim, grid_dim;
  std::function<void()> body;
  unsigned long progress = 0;  // bumped on every barrier arrival / opening and when a fiber finishes: a whole pass without any is a deadlock
  unsigned long spins = 0;     // yields from spin loops (grid barriers, flag waits): many passes with nothing else = deadlock
};
inline State& st() { static State s; return s; }
inline void* dyn_smem() { return st().cur->blk->dyn_smem; }

inline void yield() { State& s = st(); swapcontext(&s.cur->ctx, &s.sched); }
inline void open_if_complete(Sync& y) {
  if (y.live > 0 && y.arrived == y.live) { y.arrived = 0; y.gen++; st().progress++; }
}
inline void arrive_and_wait(Sync& y) {
  const unsigned my = y.gen;
  y.arrived++;
  st().progress++;
  open_if_complete(y);
  while (y.gen == my) yield();
}
inline Warp& my_warp() { State& s = st(); return s.cur->blk->warps[s.cur->tid.x >> 5]; }
// a spin loop's body: let the other fibers run (the thing waited for is produced by one of them)
inline void spin_yield() { st().spins++; yield(); }

// deposit -> everybody has deposited -> compute -> everybody has read
template <class F>
inline auto warp_collective(unsigned long long mine, F compute) {
  Warp& w = my_warp();
  const unsigned lane = st().cur->tid.x & 31u;
  w.val[lane] = mine;
  arrive_and_wait(w.sync);
  auto r = compute(w, lane);
  arrive_and_wait(w.sync);
  return r;
}

inline void trampoline() {
  State& s = st();
  s.body();
  Fiber* f = s.cur;
  f->done = true;
  s.progress++;
  Warp& w = f->blk->warps[f->tid.x >> 5];
  w.present[f->tid.x & 31u] = false;
  f->blk->sync.live--; open_if_complete(f->blk->sync);
  w.sync.live--; open_if_complete(w.sync);
  swapcontext(&f->ctx, &s.sched);  // never resumed
}

constexpr size_t STACK_BYTES = 256 * 1024;

// Runs blocks [b0, b1) of the current launch concurrently (fibers round-robin).
inline void run_blocks(unsigned b0, unsigned b1, unsigned block, size_t smem_bytes) {
  State& s = st();
  const unsigned nb = b1 - b0;
  while (s.stacks.size() < (size_t)nb * block) {
    void* p = nullptr;
    if (posix_memalign(&p, 64, STACK_BYTES)) std::abort();
    s.stacks.push_back(p);
  }
  s.blocks.assign(nb, Block());
  s.fibers.assign((size_t)nb * block, Fiber());
  for (unsigned b = 0; b < nb; b++) {
    Block& B = s.blocks[b];
    B.idx = uint3{b0 + b, 0, 0};
    B.warps.assign(block / 32, Warp());
    B.sync.live = block;
    if (smem_bytes) { if (posix_memalign(&B.dyn_smem, 128, smem_bytes)) std::abort(); std::memset(B.dyn_smem, 0xA5, smem_bytes); }
    for (unsigned t = 0; t < block; t++) {
      Fiber& f = s.fibers[(size_t)b * block + t];
      f.tid = uint3{t, 0, 0};
      f.blk = &B;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = s.stacks[(size_t)b * block + t];
      f.ctx.uc_stack.ss_size = STACK_BYTES;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, (void (*)())trampoline, 0);
      Warp& w = B.warps[t >> 5];
      w.sync.live++;
      w.present[t & 31u] = true;
    }
  }
  size_t alive = s.fibers.size();
  unsigned idle_passes = 0;
  while (alive) {
    const unsigned long before = s.progress;
    alive = 0;
    for (Fiber& f : s.fibers) {
      if (f.done) continue;
      s.cur = &f;
      swapcontext(&s.sched, &f.ctx);
      if (!f.done) alive++;
    }
    if (alive && s.progress == before) {
      // spinning fibers may legitimately need a few passes (the value they wait for is written without a barrier in between)
      if (++idle_passes > 1000) {
        std::fprintf(stderr, "cuda_emu: deadlock in blocks [%u, %u) (%zu threads wait at a barrier / flag nobody will reach)\n", b0, b1, alive);
        std::abort();
      }
    } else idle_passes = 0;
  }
  for (Block& B : s.blocks) if (B.dyn_smem) std::free(B.dyn_smem);
  s.blocks.clear();
  s.cur = nullptr;
}

// Runs kernel(args...) for grid x block threads, one block after another.
template <class K, class... Args>
void launch(K kernel, unsigned grid, unsigned block, Args... args) {
  State& s = st();
  if (block == 0 || grid == 0) return;
  if (block % 32 != 0) { std::fprintf(stderr, "cuda_emu: block size %u is not a multiple of 32\n", block); std::abort(); }
  s.block_dim = dim3(block);
  s.grid_dim = dim3(grid);
  s.body = [&]() { kernel(args...); };
  for (unsigned b = 0; b < grid; b++) run_blocks(b, b + 1, block, 0);
}

// Cooperative launch: every block of the grid is alive at once (grid-wide barriers, spin waits between blocks).
template <class K, class... Args>
void launch_coop(K kernel, unsigned grid, unsigned block, size_t smem_bytes, Args... args) {
  State& s = st();
  if (block == 0 || grid == 0) return;
  if (block % 32 != 0) { std::fprintf(stderr, "cuda_emu: block size %u is not a multiple of 32\n", block); std::abort(); }
  s.block_dim = dim3(block);
  s.grid_dim = dim3(grid);
  s.body = [&]() { kernel(args...); };
  run_blocks(0, grid, block, smem_bytes);
}

}  // namespace emu

#define threadIdx (emu::st().cur->tid)
#define blockIdx (emu::st().cur->blk->idx)
#define blockDim (emu::st().block_dim)
#define gridDim (emu::st().grid_dim)

inline void __syncthreads() { emu::arrive_and_wait(emu::st().cur->blk->sync); }
inline int __syncthreads_or(int pred) {
  emu::Block* b = emu::st().cur->blk;
  if (pred) b->or_acc = 1;
  __syncthreads();
  const int r = b->or_acc != 0;
  __syncthreads();
  b->or_acc = 0;
  __syncthreads();
  return r;
}
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emu::arrive_and_wait(emu::my_warp().sync); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

inline unsigned __match_any_sync(unsigned, unsigned long long v) {
  return emu::warp_collective(v, [v](emu::Warp& w, unsigned) {
    unsigned m = 0;
    for (unsigned l = 0; l < 32; l++) if (w.present[l] && w.val[l] == v) m |= 1u << l;
    return m;
  });
}
inline unsigned __shfl_sync(unsigned, unsigned v, unsigned src) {
  return emu::warp_collective(v, [src](emu::Warp& w, unsigned) { return (unsigned)w.val[src & 31u]; });
}
inline unsigned __shfl_up_sync(unsigned, unsigned v, unsigned delta) {
  return emu::warp_collective(v, [delta, v](emu::Warp& w, unsigned lane) { return lane >= delta ? (unsigned)w.val[lane - delta] : v; });
}
inline unsigned long long __shfl_up_sync(unsigned, unsigned long long v, unsigned delta) {
  return emu::warp_collective(v, [delta, v](emu::Warp& w, unsigned lane) { return lane >= delta ? w.val[lane - delta] : v; });
}
inline unsigned __ballot_sync(unsigned, bool pred) {
  return emu::warp_collective(pred ? 1ull : 0ull, [](emu::Warp& w, unsigned) {
    unsigned m = 0;
    for (unsigned l = 0; l < 32; l++) if (w.present[l] && w.val[l]) m |= 1u << l;
    return m;
  });
}
inline unsigned long long __shfl_sync(unsigned, unsigned long long v, unsigned src) {
  return emu::warp_collective(v, [src](emu::Warp& w, unsigned) { return w.val[src & 31u]; });
}
inline bool __all_sync(unsigned, bool pred) {
  return emu::warp_collective(pred ? 1ull : 0ull, [](emu::Warp& w, unsigned) {
    for (unsigned l = 0; l < 32; l++) if (w.present[l] && !w.val[l]) return false;
    return true;
  });
}
inline bool __any_sync(unsigned, bool pred) {
  return emu::warp_collective(pred ? 1ull : 0ull, [](emu::Warp& w, unsigned) {
    for (unsigned l = 0; l < 32; l++) if (w.present[l] && w.val[l]) return true;
    return false;
  });
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  return emu::warp_collective(v, [](emu::Warp& w, unsigned) {
    unsigned sum = 0;
    for (unsigned l = 0; l < 32; l++) if (w.present[l]) sum += (unsigned)w.val[l];
    return sum;
  });
}

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, const T& v) { *p = v; }
template <class T> inline void __stcs(T* p, const T& v) { *p = v; }

inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicAnd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o & v; return o; }
inline unsigned atomicOr(unsigned* p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
inline unsigned atomicExch(unsigned* p, unsigned v) { const unsigned o = *p; *p = v; return o; }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { const unsigned o = *p; if (o == cmp) *p = v; return o; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
  const unsigned long long o = *p;
  if (o == cmp) *p = v;
  return o;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int k = 0; k < 4; k++) c += ((a >> (8 * k)) & 0xFFu) * ((b >> (8 * k)) & 0xFFu);
  return c;
}
