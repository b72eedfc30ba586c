// gub_p2p.cuh — request routing between the GPUs of one box over NVLink peer memory, fused into the routing kernels.
//
// Replaces the reference's peer forwarding (gubernator.go:257-283 -> peer_client.go:284 runBatch -> GetPeerRateLimits,
// gubernator.go:462) inside one NVSwitch domain.  Instead of "partition locally, then three NCCL all-to-alls and a
// host round trip for the split sizes", the kernel that partitions a batch by owning shard stores every 64-byte
// request record straight into the owner's mailbox (a peer-mapped buffer) and publishes the per-owner counts with a
// release store; the owner's gather kernel acquires those flags, compacts the W mailbox segments into its dense inbox
// (source-rank order, then source index order: the deterministic order the tests reproduce) and, after evaluation, the
// response kernel stores every 32-byte response back into the source's response mailbox the same way.
//
// Per rank, one peer-visible allocation (exported with cudaIpcGetMemHandle, or shared by pointer inside one process):
//   req_mb  [2][W][cap] gub_req    written by sources     resp_mb [2][W][cap] gub_resp   written by owners
//   req_flag[2][W] u64 (epoch << 32 | count)              resp_flag[2][W] u64 (epoch)
// indexed by step parity so that step e+1 may start filling while a slow peer still drains step e.
#pragma once
#if !defined(GUB_EMULATE)  // tests/kernel_emu_harness.cpp compiles this header for the CPU (see gub_kernels.cuh)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "gub_kernels.cuh"

namespace gub {

struct P2PView {        // one rank's mailbox block as seen from a given process
  gub_req* req_mb;
  gub_resp* resp_mb;
  unsigned long long* req_flag;
  unsigned long long* resp_flag;
};

struct P2PArgs {
  P2PView peers[MAX_SHARDS];  // peers[r] = rank r's block (peers[rank] is our own)
  uint32_t world, rank, cap, epoch;
  uint32_t* done_ctr;         // [2] last-block-done counters (scatter, push_resp)
  uint32_t* error;            // set when a spin timed out
};

__device__ __forceinline__ size_t mb_index(const P2PArgs& P, uint32_t src, uint32_t p) {
  return ((size_t)(P.epoch & 1u) * P.world + src) * P.cap + p;
}
// Spins until flag's epoch field equals `epoch`; gives up after ~2 s (a peer died): sets *error so the host can tell.
__device__ __forceinline__ unsigned long long wait_flag(const unsigned long long* flag, uint32_t epoch, uint32_t* error) {
  for (uint32_t it = 0; it < 20000000u; it++) {
    const unsigned long long v = ld_acquire_sys(flag);
    if ((uint32_t)(v >> 32) == epoch) return v;
    __nanosleep(100);
  }
  atomicExch(error, 1u);
  return ((unsigned long long)epoch << 32);  // count 0
}

// Stable partition by owner, written directly into the owners' mailboxes (compare k_route_scatter, which writes a local
// buffer).  tile_offsets = exclusive scan of per-(owner, tile) counts in owner-major order; owner o's dense segment starts
// at tile_offsets[o * ntiles].  perm[dense position] = original index (kept locally for the way back).
__global__ void __launch_bounds__(256) k_p2p_scatter(const P2PArgs P, const gub_req* reqs, uint32_t n, const uint8_t* owner,
                                                     const uint32_t* tile_offsets, uint32_t ntiles, const uint32_t* counts, uint32_t* perm) {
  __shared__ uint32_t run[MAX_SHARDS], seg0[MAX_SHARDS];
  __shared__ bool is_last;
  if (threadIdx.x < MAX_SHARDS) {
    run[threadIdx.x] = (threadIdx.x < P.world) ? tile_offsets[threadIdx.x * ntiles + blockIdx.x] : 0;
    seg0[threadIdx.x] = (threadIdx.x < P.world) ? tile_offsets[threadIdx.x * ntiles] : 0;
  }
  __syncthreads();
  const uint32_t base = blockIdx.x * ROUTE_TILE;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (uint32_t chunk = 0; chunk < ROUTE_TILE; chunk += blockDim.x) {
    for (uint32_t w = 0; w < nwarps; w++) {
      if (w == warp) {
        const uint32_t i = base + chunk + threadIdx.x;
        const bool valid = i < n;
        const uint32_t o = valid ? (uint32_t)(owner[i] & 0x7Fu) : 0xFFu;
        const uint32_t mates = __match_any_sync(0xFFFFFFFFu, o);
        const uint32_t rank_in = __popc(mates & ((1u << lane) - 1u));
        const uint32_t leader = __ffs(mates) - 1;
        uint32_t start = 0;
        if (valid && lane == leader) { start = run[o]; run[o] = start + __popc(mates); }
        start = __shfl_sync(0xFFFFFFFFu, start, leader);
        if (valid) {
          const uint32_t dense = start + rank_in;
          const ulonglong2* src = reinterpret_cast<const ulonglong2*>(reqs + i);
          ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.peers[o].req_mb + mb_index(P, P.rank, dense - seg0[o]));  // NVLink store
          dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2); dst[3] = __ldg(src + 3);
          perm[dense] = i;
        }
      }
      __syncthreads();
    }
  }
  // the last block to finish publishes "rank P.rank sent counts[o] records for this epoch" to every owner
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&P.done_ctr[0], 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < P.world) {
      __threadfence_system();
      st_release_sys(&P.peers[threadIdx.x].req_flag[(size_t)(P.epoch & 1u) * P.world + P.rank],
                     ((unsigned long long)P.epoch << 32) | (unsigned long long)counts[threadIdx.x]);
    }
    if (threadIdx.x == 0) P.done_ctr[0] = 0;
  }
}

// With n == 0 nothing is scattered, but the owners still wait for our flags.
__global__ void k_p2p_publish_empty(const P2PArgs P) {
  if (threadIdx.x < P.world)
    st_release_sys(&P.peers[threadIdx.x].req_flag[(size_t)(P.epoch & 1u) * P.world + P.rank], (unsigned long long)P.epoch << 32);
}

// Owner side: wait for every source's flag, then compact the W mailbox segments into the dense inbox.
// seg_off[0..W] (device) and *m_out receive the segment starts and the total.
__global__ void __launch_bounds__(256) k_p2p_gather(const P2PArgs P, gub_req* inbox, uint32_t* seg_off, uint32_t* m_out) {
  __shared__ uint32_t off[MAX_SHARDS + 1];
  if (threadIdx.x < P.world) {
    const unsigned long long v = wait_flag(&P.peers[P.rank].req_flag[(size_t)(P.epoch & 1u) * P.world + threadIdx.x], P.epoch, P.error);
    off[threadIdx.x + 1] = (uint32_t)(v & 0xFFFFFFFFull);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    off[0] = 0;
    for (uint32_t s = 0; s < P.world; s++) off[s + 1] += off[s];
    if (blockIdx.x == 0) { for (uint32_t s = 0; s <= P.world; s++) seg_off[s] = off[s]; *m_out = off[P.world]; }
  }
  __syncthreads();
  const uint32_t m = off[P.world];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
    uint32_t s = 0;
    while (j >= off[s + 1]) s++;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.peers[P.rank].req_mb + mb_index(P, s, j - off[s]));
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(inbox + j);
    dst[0] = __ldcg(src); dst[1] = __ldcg(src + 1); dst[2] = __ldcg(src + 2); dst[3] = __ldcg(src + 3);
  }
}

// Owner side, after evaluation: every response goes back to the source's response mailbox, at the record's position.
__global__ void __launch_bounds__(256) k_p2p_push_resp(const P2PArgs P, const gub_resp* resp, const uint32_t* seg_off) {
  __shared__ uint32_t off[MAX_SHARDS + 1];
  __shared__ bool is_last;
  if (threadIdx.x <= P.world) off[threadIdx.x] = seg_off[threadIdx.x];
  __syncthreads();
  const uint32_t m = off[P.world];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
    uint32_t s = 0;
    while (j >= off[s + 1]) s++;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(resp + j);
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(P.peers[s].resp_mb + mb_index(P, P.rank, j - off[s]));  // NVLink store
    dst[0] = src[0]; dst[1] = src[1];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&P.done_ctr[1], 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < P.world) {
      __threadfence_system();
      st_release_sys(&P.peers[threadIdx.x].resp_flag[(size_t)(P.epoch & 1u) * P.world + P.rank], (unsigned long long)P.epoch << 32);
    }
    if (threadIdx.x == 0) P.done_ctr[1] = 0;
  }
}

// Source side: wait for every owner's response flag, then restore request order: out[perm[dense]] = resp_mb[owner][pos].
__global__ void __launch_bounds__(256) k_p2p_unroute(const P2PArgs P, const uint32_t* tile_offsets, uint32_t ntiles, const uint32_t* perm, uint32_t n,
                                                     gub_resp* out) {
  __shared__ uint32_t seg0[MAX_SHARDS + 1];
  if (threadIdx.x < P.world) {
    wait_flag(&P.peers[P.rank].resp_flag[(size_t)(P.epoch & 1u) * P.world + threadIdx.x], P.epoch, P.error);
    seg0[threadIdx.x] = tile_offsets[threadIdx.x * ntiles];
  }
  if (threadIdx.x == 0) seg0[P.world] = n;
  __syncthreads();
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    uint32_t o = 0;
    while (j >= seg0[o + 1]) o++;
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.peers[P.rank].resp_mb + mb_index(P, o, j - seg0[o]));
    ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + perm[j]);
    dst[0] = __ldcg(src); dst[1] = __ldcg(src + 1);
  }
}

// n == 0: still consume the owners' flags so epochs stay aligned.
__global__ void k_p2p_wait_resp_only(const P2PArgs P) {
  if (threadIdx.x < P.world) wait_flag(&P.peers[P.rank].resp_flag[(size_t)(P.epoch & 1u) * P.world + threadIdx.x], P.epoch, P.error);
}

}  // namespace gub
