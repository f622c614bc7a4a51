// probe_gather.cuh — the per-tile body of find_or_insert / lookup, shared by the
// single-table kernel (lookup.cu) and the owner side of the sharded path (peer.cu), where the keys
// come from a peer's push region and the rows are stored straight into the requester's memory.
//
// One warp owns a tile of up to 32 keys.
//   probe   lane i resolves key i on its own: one 128-byte bucket line, SIMD tag compare in
//           registers, one 8-byte key compare per tag hit (same line); on a miss (find_or_insert
//           only) a 64-bit CAS on the first free slot of the bucket. 32 independent probes in flight.
//   gather  the warp then streams the 32 rows as one flat array of 16-byte chunks: chunk c of the
//           tile belongs to key c / CPR, so every warp instruction moves 512 contiguous-per-row
//           bytes, fully coalesced on both the arena and the output side, UNROLL loads in flight
//           per lane before the first store.
// Rows of keys that are new in this batch are never read from the arena: every duplicate
// computes init_chunk(key) itself (pure function), only the CAS winner writes it back.
#pragma once
#include "table.h"

namespace meepo {

template <int CPR>
__device__ __forceinline__ void gather_tile_fast(const TableView& t, uint32_t slot, uint32_t tile_keys,
                                                 uint4* __restrict__ out_tile, uint32_t lane) {
  constexpr int UNROLL = CPR >= 8 ? 8 : CPR;
  // chunk c = it*32 + lane; key j = c / CPR; offset = c % CPR
#pragma unroll 1
  for (int it0 = 0; it0 < CPR; it0 += UNROLL) {
    uint4 v[UNROLL];
    uint32_t jj[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint32_t c = (uint32_t)(it0 + u) * 32u + lane;
      const uint32_t j = c / CPR, off = c % CPR;
      const uint32_t s = __shfl_sync(0xFFFFFFFFu, slot, j);
      jj[u] = j;
      v[u] = make_uint4(0, 0, 0, 0);
      if (s != kNil) v[u] = ld_stream(t.rows + (size_t)s * CPR + off);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint32_t c = (uint32_t)(it0 + u) * 32u + lane;
      if (jj[u] < tile_keys) st_stream(out_tile + c, v[u]);
    }
  }
}

// Generic-width version (any cpr), also the path for tiles that contain freshly inserted keys.
__device__ __forceinline__ void gather_tile_slow(const TableView& t, uint64_t key, const Probe& pr,
                                                 uint32_t tile_keys, uint4* __restrict__ out_tile,
                                                 uint32_t lane) {
  const uint32_t cpr = t.cpr;
  for (uint32_t j = 0; j < tile_keys; j++) {
    const uint32_t s = __shfl_sync(0xFFFFFFFFu, pr.slot, j);
    const uint32_t st = __shfl_sync(0xFFFFFFFFu, pr.status, j);
    const bool win = __shfl_sync(0xFFFFFFFFu, (int)pr.winner, j);
    const uint64_t kj = __shfl_sync(0xFFFFFFFFu, key, j);
    for (uint32_t off = lane; off < cpr; off += 32) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (st == MEEPO_KEY_INSERTED) {
        v = init_chunk(t, kj, off);
        if (win) t.rows[(size_t)s * cpr + off] = v;
      } else if (s != kNil) {
        v = ld_stream(t.rows + (size_t)s * cpr + off);
      }
      st_stream(out_tile + (size_t)j * cpr + off, v);
    }
    if (win) {
      const uint4 sv = init_state_chunk(t);
      for (uint32_t off = lane; off < t.scpr; off += 32) t.state[(size_t)s * t.scpr + off] = sv;
    }
  }
}

struct TileCounts {
  uint32_t hit = 0, miss = 0, full = 0;
};

// Probe + gather of one tile. Lane `lane` holds `key` (valid == false for the lanes past the end of
// the tile, whose key must be MEEPO_KEY_EMPTY). status_out / slot_out / key_out are this lane's own
// output cells (may be null). occurrences = how many batch occurrences this key stands for (1, or
// the sender's duplicate count on the sharded path): added to the hit/miss counters and to the key's
// LFU score.
template <int CPR, bool INSERT>
__device__ __forceinline__ void probe_gather_tile(const TableView& t, uint64_t key, bool valid, uint32_t tile_keys,
                                                  uint4* __restrict__ out_tile, uint8_t* status_out,
                                                  uint32_t* slot_out, uint64_t* key_out, uint32_t occurrences,
                                                  const NewList& nl, TileCounts& cnt, uint32_t lane) {
  Probe pr{kNil, MEEPO_KEY_INVALID, false};
  if (INSERT) {
    pr = probe_find_or_insert(t, key);
  } else if (key_valid(key)) {
    pr.slot = probe_find<kReadOnly>(t, key);
    pr.status = pr.slot != kNil ? MEEPO_KEY_FOUND : MEEPO_KEY_MISS;
  }
  if (valid) {
    if (status_out) *status_out = (uint8_t)pr.status;
    if (slot_out) {
      *slot_out = pr.slot;
      *key_out = key;
    }
    cnt.hit += pr.status == MEEPO_KEY_FOUND ? occurrences : 0u;
    cnt.miss += pr.status == MEEPO_KEY_MISS ? occurrences : 0u;
    cnt.full += pr.status == MEEPO_KEY_FULL ? occurrences : 0u;
    if (t.scores && pr.slot != kNil) {  // meepo.h "Evict": freq += occurrences, last_epoch = epoch
      atomicAdd(&t.scores[pr.slot].x, occurrences);
      t.scores[pr.slot].y = t.epoch;
    }
  }
  bool fresh = false;
  if (INSERT) {
    const unsigned wm = __ballot_sync(0xFFFFFFFFu, pr.winner);
    if (wm) {  // list the claimed slots for publish_kernel (one atomic per warp)
      const int leader = __ffs(wm) - 1;
      uint32_t base = 0;
      if ((int)lane == leader) base = atomicAdd(nl.count, (uint32_t)__popc(wm));
      base = __shfl_sync(0xFFFFFFFFu, base, leader);
      if (pr.winner) nl.slots[base + __popc(wm & ((1u << lane) - 1u))] = pr.slot;
    }
    fresh = __any_sync(0xFFFFFFFFu, pr.status == MEEPO_KEY_INSERTED);
  }
  if (CPR > 0 && !fresh)
    gather_tile_fast<(CPR > 0 ? CPR : 1)>(t, pr.slot, tile_keys, out_tile, lane);
  else
    gather_tile_slow(t, key, pr, tile_keys, out_tile, lane);
}

// stats: one atomic per warp per counter for the whole launch
__device__ __forceinline__ void flush_tile_counts(const TableView& t, TileCounts c, uint32_t lane) {
  c.hit = __reduce_add_sync(0xFFFFFFFFu, c.hit);
  c.miss = __reduce_add_sync(0xFFFFFFFFu, c.miss);
  c.full = __reduce_add_sync(0xFFFFFFFFu, c.full);
  if (lane == 0) {
    if (c.hit) atomicAdd(t.counters + C_HITS, (unsigned long long)c.hit);
    if (c.miss) atomicAdd(t.counters + C_MISSES, (unsigned long long)c.miss);
    if (c.full) atomicAdd(t.counters + C_FULL, (unsigned long long)c.full);
  }
}

}  // namespace meepo
