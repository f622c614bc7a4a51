// compact.cuh — single-pass ORDERED stream compaction (decoupled look-back over per-tile counts) and the
// one single-CTA exclusive scan of the library.
//
// Three passes of the library are "keep the flagged positions, in order, densely numbered": the segment
// heads of a sorted slot array (update.cu), the occupied cells of the dedup scratch table (shard.cu) and
// the live slots of the table (io.cu). Each used to be count -> single-CTA scan -> fill (three launches,
// the middle one a 1-CTA kernel); here it is ONE kernel: a CTA takes the next tile from a device ticket
// (so tiles start in order and a tile only ever waits for tiles that are already running), counts its
// flags, publishes the count, learns the number of flagged positions before it from the tiles behind it
// (decoupled look-back) and emits its own. The CTA of the last tile also receives the grand total.
//
// Usage inside a __global__ function with blockDim.x == kCompactThreads:
//     CompactTile ct = compact_begin(cs, n);                       // claims a tile
//     unsigned flags = 0;  for k < kCompactItems: if (pred(ct.pos(k))) flags |= 1u << k;   (pos(k) < n)
//     compact_rank(ct, flags, cs);                                 // look-back; fills ct.base / ct.total
//     for k: if (flags >> k & 1) emit(ct.pos(k), ct.rank(k));      // rank = dense index, in position order
//     if (ct.last) { total = ct.base + ct.tile_total; ... }
// cs.state (ntiles x u64) and cs.ticket (u32) must be zero before the launch (one cudaMemsetAsync).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace meepo {

constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 8;
constexpr uint32_t kCompactTile = kCompactThreads * kCompactItems;  // positions per CTA
constexpr unsigned long long kCfPartial = 1ull << 62, kCfInclusive = 2ull << 62, kCfMask = 3ull << 62;

struct CompactState {
  unsigned long long* state;  // [ntiles] flag | count
  uint32_t* ticket;           // next tile to hand out
  uint32_t* error;            // sticky: a predecessor never showed up (results are wrong), may be null
};
static inline size_t compact_state_bytes(uint64_t n) {
  return ((n + kCompactTile - 1) / kCompactTile + 2) * 8 + 64;
}
static inline uint32_t compact_tiles(uint64_t n) { return (uint32_t)((n + kCompactTile - 1) / kCompactTile); }

#ifdef __CUDACC__
struct CompactTile {
  uint32_t tile, tid, lane, warp;
  uint64_t n;
  unsigned long long base;  // flagged positions before this tile
  uint32_t tile_total;      // flagged positions of this tile
  bool last;                // this is the last tile: base + tile_total is the grand total
  unsigned ballots[kCompactItems];
  uint32_t warp_off[kCompactItems];  // flagged positions of this tile before (round k, this warp)
  // position of this thread's item k: rounds are contiguous runs of 256 positions, so position order
  // == (round, thread) order
  __device__ __forceinline__ uint64_t pos(int k) const {
    return (uint64_t)tile * kCompactTile + (uint32_t)k * kCompactThreads + tid;
  }
  __device__ __forceinline__ unsigned long long rank(int k) const {
    return base + warp_off[k] + __popc(ballots[k] & ((1u << lane) - 1u));
  }
};

__device__ __forceinline__ unsigned long long compact_ld(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void compact_st(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ CompactTile compact_begin(const CompactState& cs, uint64_t n) {
  __shared__ uint32_t s_tile;
  CompactTile ct;
  ct.tid = threadIdx.x;
  ct.lane = threadIdx.x & 31u;
  ct.warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(cs.ticket, 1u);
  __syncthreads();
  ct.tile = s_tile;
  ct.n = n;
  ct.base = 0;
  ct.tile_total = 0;
  ct.last = ((uint64_t)ct.tile + 1) * kCompactTile >= n;
  return ct;
}

// flags: bit k set if this thread's item k is kept. Every thread of the CTA must call this.
__device__ __forceinline__ void compact_rank(CompactTile& ct, unsigned flags, const CompactState& cs) {
  constexpr int W = kCompactThreads / 32;
  __shared__ uint32_t s_cnt[kCompactItems * W];  // (round, warp) counts, then exclusive offsets
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_total;
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) {
    ct.ballots[k] = __ballot_sync(0xFFFFFFFFu, (flags >> k) & 1u);
    if (ct.lane == 0) s_cnt[k * W + ct.warp] = __popc(ct.ballots[k]);
  }
  __syncthreads();
  if (ct.warp == 0) {
    // exclusive scan of the kCompactItems * W (= 64) counts: two per lane
    static_assert(kCompactItems * W == 64, "two counts per lane");
    const uint32_t a = s_cnt[2 * ct.lane], b = s_cnt[2 * ct.lane + 1];
    uint32_t x = a + b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (ct.lane >= (uint32_t)d) x += y;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, x, 31);
    s_cnt[2 * ct.lane] = x - a - b;
    s_cnt[2 * ct.lane + 1] = x - b;
    // publish, then look back over the tiles behind this one, 32 at a time (lane 0 = nearest)
    unsigned long long* mine = cs.state + ct.tile;
    if (ct.lane == 0) compact_st(mine, (unsigned long long)total | (ct.tile == 0 ? kCfInclusive : kCfPartial));
    unsigned long long excl = 0;
    long long prev = (long long)ct.tile - 1;
    while (prev >= 0) {
      const long long j = prev - (long long)ct.lane;
      unsigned long long v = j >= 0 ? compact_ld(cs.state + j) : kCfInclusive;
      for (uint32_t spin = 0; __any_sync(0xFFFFFFFFu, (v & kCfMask) == 0); spin++) {
        if (spin > (1u << 22)) {  // a tile that never shows up: give up rather than hang the GPU
          if (cs.error) *reinterpret_cast<volatile uint32_t*>(cs.error) = 1u;
          if ((v & kCfMask) == 0) v = kCfInclusive;
          break;
        }
        if ((v & kCfMask) == 0) {
          __nanosleep(20);
          v = compact_ld(cs.state + j);
        }
      }
      const unsigned inc = __ballot_sync(0xFFFFFFFFu, (v & kCfMask) == kCfInclusive);
      const int first = inc ? __ffs(inc) - 1 : 31;
      unsigned long long c = (int)ct.lane <= first ? (v & ~kCfMask) : 0ull;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
      excl += c;
      if (inc) break;
      prev -= 32;
    }
    if (ct.lane == 0) {
      if (ct.tile != 0) compact_st(mine, (excl + total) | kCfInclusive);
      s_base = excl;
      s_total = total;
    }
  }
  __syncthreads();
  ct.base = s_base;
  ct.tile_total = s_total;
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) ct.warp_off[k] = s_cnt[k * W + ct.warp];
}

// ---------------------------------------------------------------------------------------------
// The single-CTA exclusive scan (n is small: tiles x shards of the owner partition). out[n] = total.
// `skip` (optional, device): the whole pass is a no-op when *skip != 0.
__device__ __forceinline__ void block_excl_scan_1024(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                     uint32_t n, unsigned long long* total64) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[w] = x;
    __syncthreads();
    if (w == 0) {
      uint32_t s = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, d);
        if (lane >= d) s += y;
      }
      warp_sum[lane] = s;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    const uint32_t incl = x + (w ? warp_sum[w - 1] : 0);
    if (i < n) out[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[n] = carry_s;
    if (total64) *total64 = carry_s;
  }
}
#endif  // __CUDACC__

}  // namespace meepo
