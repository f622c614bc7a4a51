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

// SCATTER: the rows of a tile do not go to one contiguous block but each to its own destination: lane j
// holds the address of key j's output row in `dst` (the sharded owner stores rows straight into the
// requester's output tensor, one row per unique key, wherever that key first occurs in the batch).
template <bool SCATTER>
__device__ __forceinline__ uint4* tile_dst(uint4* __restrict__ out_tile, unsigned long long dst, uint32_t j,
                                           uint32_t off, uint32_t c) {
  if constexpr (SCATTER)
    return reinterpret_cast<uint4*>(__shfl_sync(0xFFFFFFFFu, dst, j)) + off;
  else
    return out_tile + c;
}

template <int CPR, bool SCATTER = false>
__device__ __forceinline__ void gather_tile_fast(const TableView& t, uint32_t slot, uint32_t tile_keys,
                                                 uint4* __restrict__ out_tile, uint32_t lane,
                                                 unsigned long long dst = 0) {
  constexpr int UNROLL = CPR >= 8 ? 8 : CPR;
  // chunk c = it*32 + lane; key j = c / CPR; offset = c % CPR
#pragma unroll 1
  for (int it0 = 0; it0 < CPR; it0 += UNROLL) {
    uint4 v[UNROLL];
    uint32_t jj[UNROLL];
    uint4* dd[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint32_t c = (uint32_t)(it0 + u) * 32u + lane;
      const uint32_t j = c / CPR, off = c % CPR;
      const uint32_t s = __shfl_sync(0xFFFFFFFFu, slot, j);
      dd[u] = tile_dst<SCATTER>(out_tile, dst, j, off, c);
      jj[u] = j;
      v[u] = make_uint4(0, 0, 0, 0);
      if (s != kNil) v[u] = ld_stream(t.rows + (size_t)s * CPR + off);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
      if (jj[u] < tile_keys) st_stream(dd[u], v[u]);
  }
}

// Freshly inserted keys of a tile, after the fast gather has moved every other row (their slots are
// masked out there): one key at a time, the lanes cover its chunks. A chunk of a new row is computed
// (init_chunk is a pure function of key and column), never loaded; the CAS winner also writes it —
// and the initial optimizer state — to the arenas.
template <int CPR, bool SCATTER = false>
__device__ __forceinline__ void fixup_fresh(const TableView& t, uint64_t key, const Probe& pr, uint32_t tile_keys,
                                            uint4* __restrict__ out_tile, uint32_t lane, unsigned long long dst = 0) {
  unsigned m = __ballot_sync(0xFFFFFFFFu, pr.status == MEEPO_KEY_INSERTED);
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t s = __shfl_sync(0xFFFFFFFFu, pr.slot, j);
    const bool win = __shfl_sync(0xFFFFFFFFu, (int)pr.winner, j);
    const uint64_t kj = __shfl_sync(0xFFFFFFFFu, key, j);
    uint4* orow = out_tile + (size_t)j * CPR;
    if constexpr (SCATTER) orow = reinterpret_cast<uint4*>(__shfl_sync(0xFFFFFFFFu, dst, j));
    if ((uint32_t)j >= tile_keys) continue;
    for (uint32_t off = lane; off < (uint32_t)CPR; off += 32) {
      const uint4 v = init_chunk(t, kj, off);
      if (win) t.rows[(size_t)s * CPR + off] = v;
      st_stream(orow + off, v);
    }
    if (win) {
      const uint4 sv = init_state_chunk(t);
      for (uint32_t off = lane; off < t.scpr; off += 32) t.state[(size_t)s * t.scpr + off] = sv;
    }
  }
}

// Keys served from the host tier (meepo.h "Host tier"): promoted by find_or_insert (the CAS winner restores row,
// state, step and the tier's freq into its slot) or read through by lookup. One key at a time, the lanes cover
// its chunks; the tuple is read from the staging buffer (HBM) or the pinned ring (zero-copy over PCIe).
template <bool SCATTER = false>
__device__ __forceinline__ void fixup_tier(const TableView& t, const Probe& pr, uint32_t tslab, uint32_t tile_keys,
                                           uint4* __restrict__ out_tile, uint32_t lane, unsigned long long dst = 0) {
  unsigned m = __ballot_sync(0xFFFFFFFFu, tslab != kNil);
  const uint32_t cpr = t.cpr;
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t s = __shfl_sync(0xFFFFFFFFu, pr.slot, j);
    const bool win = __shfl_sync(0xFFFFFFFFu, (int)pr.winner, j);
    const uint32_t slab = __shfl_sync(0xFFFFFFFFu, tslab, j);
    uint4* orow = out_tile + (size_t)j * cpr;
    if constexpr (SCATTER) orow = reinterpret_cast<uint4*>(__shfl_sync(0xFFFFFFFFu, dst, j));
    if ((uint32_t)j >= tile_keys) continue;
    const TierTuple tt = tier_tuple(t, slab);
    for (uint32_t off = lane; off < cpr; off += 32) {
      const uint4 v = tt.rows[off];
      if (win) t.rows[(size_t)s * cpr + off] = v;
      st_stream(orow + off, v);
    }
    if (win) {
      for (uint32_t off = lane; off < t.scpr; off += 32) t.state[(size_t)s * t.scpr + off] = tt.state[off];
      if (lane == 0) {
        // the slot's scores were zero: the occurrences of this batch are added by the score path, here the past
        if (t.scores) atomicAdd(&t.scores[s].x, tt.meta->z);
        if (t.steps) t.steps[s] = *tt.steps;
      }
    }
  }
}

// Generic-width version (any cpr): the fallback for row widths without a specialised kernel.
template <bool SCATTER = false>
__device__ __forceinline__ void gather_tile_slow(const TableView& t, uint64_t key, const Probe& pr, uint32_t tslab,
                                                 uint32_t tile_keys, uint4* __restrict__ out_tile,
                                                 uint32_t lane, unsigned long long dst = 0) {
  const uint32_t cpr = t.cpr;
  for (uint32_t j = 0; j < tile_keys; j++) {
    if (__shfl_sync(0xFFFFFFFFu, tslab, j) != kNil) continue;  // fixup_tier moves it
    const uint32_t s = __shfl_sync(0xFFFFFFFFu, pr.slot, j);
    const uint32_t st = __shfl_sync(0xFFFFFFFFu, pr.status, j);
    const bool win = __shfl_sync(0xFFFFFFFFu, (int)pr.winner, j);
    const uint64_t kj = __shfl_sync(0xFFFFFFFFu, key, j);
    uint4* orow = out_tile + (size_t)j * cpr;
    if constexpr (SCATTER) orow = reinterpret_cast<uint4*>(__shfl_sync(0xFFFFFFFFu, dst, j));
    for (uint32_t off = lane; off < cpr; off += 32) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (st == MEEPO_KEY_INSERTED) {
        v = init_chunk(t, kj, off);
        if (win) t.rows[(size_t)s * cpr + off] = v;
      } else if (s != kNil) {
        v = ld_stream(t.rows + (size_t)s * cpr + off);
      }
      st_stream(orow + off, v);
    }
    if (win) {
      const uint4 sv = init_state_chunk(t);
      for (uint32_t off = lane; off < t.scpr; off += 32) t.state[(size_t)s * t.scpr + off] = sv;
    }
  }
}

// NOOUT tiles (the pooled verbs: rows are not returned per key, a second kernel reads the arena by slot): only
// the CAS winners have work — the new row + initial state, or the tuple coming back from the host tier, is
// written to the arena. One key at a time, the lanes cover its chunks.
__device__ __forceinline__ void init_winners(const TableView& t, uint64_t key, const Probe& pr, uint32_t tslab,
                                             uint32_t lane) {
  unsigned m = __ballot_sync(0xFFFFFFFFu, pr.winner);
  const uint32_t cpr = t.cpr;
  while (m) {
    const int j = __ffs(m) - 1;
    m &= m - 1;
    const uint32_t s = __shfl_sync(0xFFFFFFFFu, pr.slot, j);
    const uint64_t kj = __shfl_sync(0xFFFFFFFFu, key, j);
    const uint32_t slab = __shfl_sync(0xFFFFFFFFu, tslab, j);
    if (slab != kNil) {
      const TierTuple tt = tier_tuple(t, slab);
      for (uint32_t off = lane; off < cpr; off += 32) t.rows[(size_t)s * cpr + off] = tt.rows[off];
      for (uint32_t off = lane; off < t.scpr; off += 32) t.state[(size_t)s * t.scpr + off] = tt.state[off];
      if (lane == 0) {
        if (t.scores) atomicAdd(&t.scores[s].x, tt.meta->z);
        if (t.steps) t.steps[s] = *tt.steps;
      }
    } else {
      for (uint32_t off = lane; off < cpr; off += 32) t.rows[(size_t)s * cpr + off] = init_chunk(t, kj, off);
      const uint4 sv = init_state_chunk(t);
      for (uint32_t off = lane; off < t.scpr; off += 32) t.state[(size_t)s * t.scpr + off] = sv;
    }
  }
}

struct TileCounts {
  uint32_t hit = 0, miss = 0, full = 0, tier = 0;
};

// LFU / LRU score updates (meepo.h "Evict": freq += occurrences, last_epoch = epoch) go through a
// small per-CTA table in shared memory and reach the score array once per CTA and slot: the hottest
// Zipf key is ~8% of a batch, and reductions on ONE global address serialise in L2 (measured: +0.2 ms
// on a 1M-key batch without this). A slot whose cell is taken by another slot updates global memory
// directly. The kernel calls score_cache_init before and score_cache_flush after its tile loop.
constexpr uint32_t kScoreCells = 512;
struct ScoreCache {
  uint32_t* slot;  // [kScoreCells], kNil = free
  uint32_t* freq;  // [kScoreCells]
};
__device__ __forceinline__ void score_cache_init(const TableView& t, const ScoreCache& sc) {
  if (!t.scores) return;
  for (uint32_t i = threadIdx.x; i < kScoreCells; i += blockDim.x) {
    sc.slot[i] = kNil;
    sc.freq[i] = 0;
  }
  __syncthreads();
}
__device__ __forceinline__ void score_cache_add(const TableView& t, const ScoreCache& sc, uint32_t s, uint32_t n) {
  const uint32_t h = (s * 0x9E3779B1u) >> (32 - 9);
  const uint32_t old = atomicCAS(&sc.slot[h], kNil, s);
  if (old == kNil || old == s) {
    atomicAdd(&sc.freq[h], n);
  } else {
    atomicAdd(&t.scores[s].x, n);
    t.scores[s].y = t.epoch;
  }
}
__device__ __forceinline__ void score_cache_flush(const TableView& t, const ScoreCache& sc) {
  if (!t.scores) return;
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < kScoreCells; i += blockDim.x) {
    const uint32_t s = sc.slot[i];
    if (s != kNil) {
      atomicAdd(&t.scores[s].x, sc.freq[i]);
      t.scores[s].y = t.epoch;
    }
  }
}
static_assert(kScoreCells == 1u << 9, "score_cache_add hashes to 9 bits");

// Probe + gather of one tile. Lane `lane` holds `key` (valid == false for the lanes past the end of
// the tile, whose key must be MEEPO_KEY_EMPTY). status_out / slot_out / key_out are this lane's own
// output cells (may be null; new_out = this element's cell of the NewList). occurrences = how many batch occurrences this key stands for (1, or
// the sender's duplicate count on the sharded path): added to the hit/miss counters and to the key's
// LFU score.
// SCATTER: out_tile is ignored, this lane's row goes to (uint4*)dst (see tile_dst).
// TIER: the table has a host tier (compiled out otherwise: the tier code costs registers in the hot kernels).
// NOOUT: no rows are returned (slot_out / tslab_out say where each key's row is: arena slot, else tier slab).
template <int CPR, bool INSERT, bool SCATTER = false, bool TIER = false, bool NOOUT = false>
__device__ __forceinline__ void probe_gather_tile(const TableView& t, uint64_t key, bool valid, uint32_t tile_keys,
                                                  uint4* __restrict__ out_tile, uint8_t* status_out,
                                                  uint32_t* slot_out, uint64_t* key_out, uint32_t occurrences,
                                                  uint32_t* new_out, TileCounts& cnt, const ScoreCache& sc,
                                                  uint32_t lane, unsigned long long dst = 0,
                                                  uint32_t* tslab_out = nullptr) {
  Probe pr{kNil, MEEPO_KEY_INVALID, false};
  if (INSERT) {
    pr = probe_find_or_insert(t, key);
  } else if (key_valid(key)) {
    pr.slot = probe_find<kReadOnly>(t, key);
    pr.status = pr.slot != kNil ? MEEPO_KEY_FOUND : MEEPO_KEY_MISS;
  }
  // second level (meepo.h "Host tier"): a key that is not in HBM but in the tier is FOUND — promoted by
  // find_or_insert (unless the table is full), read through by lookup
  uint32_t tslab = kNil;
  if constexpr (TIER) {
    if (valid && (pr.status == MEEPO_KEY_INSERTED || pr.status == MEEPO_KEY_MISS)) tslab = tier_slab(t.tier, key);
    if (tslab != kNil) {
      pr.status = MEEPO_KEY_FOUND;
      if (!INSERT) cnt.tier += occurrences;
    }
  }
  if (valid) {
    if (status_out) *status_out = (uint8_t)pr.status;
    if (slot_out) *slot_out = pr.slot;
    if (key_out) *key_out = key;
    cnt.hit += pr.status == MEEPO_KEY_FOUND ? occurrences : 0u;
    cnt.miss += pr.status == MEEPO_KEY_MISS ? occurrences : 0u;
    cnt.full += pr.status == MEEPO_KEY_FULL ? occurrences : 0u;
  }
  if (t.scores) {  // meepo.h "Evict": freq += occurrences, last_epoch = epoch
    // duplicates of a key inside the tile are folded into one update first
    const uint32_t s = valid ? pr.slot : kNil;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, s);
    // (a reduction over a partial mask is a loop over the groups: skip it when every key counts once)
    const uint32_t total = __all_sync(0xFFFFFFFFu, occurrences == 1u) ? (uint32_t)__popc(peers)
                                                                      : __reduce_add_sync(peers, occurrences);
    if (s != kNil && lane == (uint32_t)(__ffs(peers) - 1)) score_cache_add(t, sc, s, total);
  }
  bool fresh = false;
  if (INSERT) {
    if (valid && new_out) *new_out = pr.winner ? pr.slot : kNil;  // for publish_kernel
    fresh = __any_sync(0xFFFFFFFFu, pr.status == MEEPO_KEY_INSERTED);
  }
  if constexpr (NOOUT) {
    if (valid && tslab_out) *tslab_out = tslab;
    if (INSERT && __any_sync(0xFFFFFFFFu, pr.winner)) init_winners(t, key, pr, tslab, lane);
    return;
  }
  if (CPR > 0) {
    gather_tile_fast<(CPR > 0 ? CPR : 1), SCATTER>(t, (pr.status == MEEPO_KEY_INSERTED || tslab != kNil) ? kNil : pr.slot,
                                                   tile_keys, out_tile, lane, dst);
    if (fresh) fixup_fresh<(CPR > 0 ? CPR : 1), SCATTER>(t, key, pr, tile_keys, out_tile, lane, dst);
  } else {
    gather_tile_slow<SCATTER>(t, key, pr, tslab, tile_keys, out_tile, lane, dst);
  }
  if constexpr (TIER)
    if (__any_sync(0xFFFFFFFFu, tslab != kNil)) fixup_tier<SCATTER>(t, pr, tslab, tile_keys, out_tile, lane, dst);
}

// stats: one atomic per warp per counter for the whole launch
__device__ __forceinline__ void flush_tile_counts(const TableView& t, TileCounts c, uint32_t lane) {
  c.hit = __reduce_add_sync(0xFFFFFFFFu, c.hit);
  c.miss = __reduce_add_sync(0xFFFFFFFFu, c.miss);
  c.full = __reduce_add_sync(0xFFFFFFFFu, c.full);
  c.tier = __reduce_add_sync(0xFFFFFFFFu, c.tier);
  if (lane == 0) {
    if (c.tier) atomicAdd(t.counters + C_TIER_HITS, (unsigned long long)c.tier);
    if (c.hit) atomicAdd(t.counters + C_HITS, (unsigned long long)c.hit);
    if (c.miss) atomicAdd(t.counters + C_MISSES, (unsigned long long)c.miss);
    if (c.full) atomicAdd(t.counters + C_FULL, (unsigned long long)c.full);
  }
}

}  // namespace meepo
