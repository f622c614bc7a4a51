// lookup.cu — find_or_insert / lookup: fused probe + row gather.
//
// One warp owns a tile of 32 consecutive batch keys.
//   probe   lane i resolves key i on its own: one 32-byte tag sector, SIMD compare in registers,
//           one 8-byte key compare per tag hit; on a miss (find_or_insert only) a 64-bit CAS on
//           the first free slot of the bucket. 32 independent probes in flight per warp.
//   gather  the warp then streams the 32 rows as one flat array of 16-byte chunks: chunk c of the
//           tile belongs to key c / CPR, so every warp instruction moves 512 contiguous-per-row
//           bytes, fully coalesced on both the arena and the output side, UNROLL loads in flight
//           per lane before the first store.
// Rows of keys that are new in this batch are never read from the arena: every duplicate
// computes init_chunk(key) itself (pure function), only the CAS winner writes it back.
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

template <int CPR, bool INSERT>
__global__ void __launch_bounds__(256) probe_gather_kernel(TableView t, const uint64_t* __restrict__ keys,
                                                           uint32_t n, uint4* __restrict__ out,
                                                           uint8_t* __restrict__ status, NewList nl, SlotCache sc) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t ntiles = (n + 31u) >> 5;
  const uint32_t cpr = CPR > 0 ? (uint32_t)CPR : t.cpr;
  uint32_t c_hit = 0, c_miss = 0, c_full = 0;
  for (uint32_t tile = warp; tile < ntiles; tile += nwarps) {
    const uint32_t i = tile * 32u + lane;
    const uint32_t tile_keys = min(32u, n - tile * 32u);
    const uint64_t key = i < n ? __ldg(keys + i) : MEEPO_KEY_EMPTY;
    Probe pr{kNil, MEEPO_KEY_INVALID, false};
    if (INSERT) {
      pr = probe_find_or_insert(t, key);
    } else if (key_valid(key)) {
      pr.slot = probe_find<kReadOnly>(t, key);
      pr.status = pr.slot != kNil ? MEEPO_KEY_FOUND : MEEPO_KEY_MISS;
    }
    if (i < n) {
      if (status) status[i] = (uint8_t)pr.status;
      if (sc.slots) {
        sc.slots[i] = pr.slot;
        sc.keys[i] = key;
      }
      c_hit += pr.status == MEEPO_KEY_FOUND;
      c_miss += pr.status == MEEPO_KEY_MISS;
      c_full += pr.status == MEEPO_KEY_FULL;
      if (t.scores && pr.slot != kNil) {  // meepo.h "Evict": freq += 1, last_epoch = epoch
        atomicAdd(&t.scores[pr.slot].x, 1u);
        t.scores[pr.slot].y = t.epoch;
      }
    }
    uint4* out_tile = out + (size_t)tile * 32u * cpr;
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
  c_hit = __reduce_add_sync(0xFFFFFFFFu, c_hit);
  c_miss = __reduce_add_sync(0xFFFFFFFFu, c_miss);
  c_full = __reduce_add_sync(0xFFFFFFFFu, c_full);
  if (lane == 0) {
    if (c_hit) atomicAdd(t.counters + C_HITS, (unsigned long long)c_hit);
    if (c_miss) atomicAdd(t.counters + C_MISSES, (unsigned long long)c_miss);
    if (c_full) atomicAdd(t.counters + C_FULL, (unsigned long long)c_full);
  }
}

// Writes the tags of the slots claimed by the preceding probe_gather_kernel<.., true> and folds the
// count into the table size. `cur` is this call's counter, `next` the other one (zeroed here so the
// next find_or_insert starts from 0 without a memset node).
__global__ void publish_kernel(TableView t, const uint32_t* __restrict__ new_slots, const uint32_t* cur,
                               uint32_t* next) {
  const uint32_t n = *cur;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t s = new_slots[i];
    *tag_ptr(t, s) = (uint8_t)digest_of(mix64(*key_ptr(t, s)));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *next = 0;
    if (n) {
      atomicAdd(t.counters + C_SIZE, (unsigned long long)n);
      atomicAdd(t.counters + C_INSERTS, (unsigned long long)n);
    }
  }
}

template <bool INSERT>
static const void* pick_kernel(uint32_t cpr) {
  switch (cpr) {
    case 1: return (const void*)probe_gather_kernel<1, INSERT>;
    case 2: return (const void*)probe_gather_kernel<2, INSERT>;
    case 4: return (const void*)probe_gather_kernel<4, INSERT>;
    case 8: return (const void*)probe_gather_kernel<8, INSERT>;
    case 16: return (const void*)probe_gather_kernel<16, INSERT>;
    case 32: return (const void*)probe_gather_kernel<32, INSERT>;
    case 64: return (const void*)probe_gather_kernel<64, INSERT>;
    default: return (const void*)probe_gather_kernel<0, INSERT>;
  }
}

// A find_or_insert / lookup call = begin, one or more chunks, end. Chunks of one call share the
// epoch and (find_or_insert) the claimed-slot list; tags are published once, in end, so a key that
// is new in the call reports INSERTED in every chunk (meepo.h "Status").
meepo_status probe_gather_begin(meepo_table* t, uint64_t n_total, bool insert, cudaStream_t stream) {
  t->epoch++;
  t->v.epoch = (uint32_t)t->epoch;
  t->cur_new.slots = nullptr;
  t->cur_new.count = nullptr;
  t->cur_new_next = nullptr;
  t->cache_valid = false;
  t->cache_off = 0;
  t->cache_n = 0;
  if (t->cache_enabled && n_total) {
    if (n_total > t->cache_cap) {
      MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
      cudaFree(t->cache.keys);
      cudaFree(t->cache.slots);
      t->cache = SlotCache{nullptr, nullptr};
      t->cache_cap = 0;
      const uint64_t cap = n_total + n_total / 8;
      MEEPO_CUDA_TRY(cudaMalloc(&t->cache.keys, cap * 8));
      MEEPO_CUDA_TRY(cudaMalloc(&t->cache.slots, cap * 4));
      t->cache_cap = cap;
    }
    t->cache_n = n_total;
  }
  if (insert && n_total) {
    MEEPO_TRY(t->ws.reserve(Workspace::pad(n_total * 4), stream));
    t->cur_new.slots = t->ws.take<uint32_t>(n_total);
    t->cur_new.count = &t->dstate->new_count[t->foi_parity];
    t->cur_new_next = &t->dstate->new_count[t->foi_parity ^ 1];
    t->foi_parity ^= 1;
  }
  return MEEPO_OK;
}

meepo_status probe_gather_chunk(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                uint8_t* status_out, bool insert, cudaStream_t stream) {
  if (n == 0) return MEEPO_OK;
  const void* kern = insert ? pick_kernel<true>(t->v.cpr) : pick_kernel<false>(t->v.cpr);
  const uint64_t tiles = (n + 31) / 32;
  const int grid = grid_for(t, kern, 256, 0, (tiles + 7) / 8);
  uint32_t n32 = (uint32_t)n;
  uint4* out = reinterpret_cast<uint4*>(rows_out);
  NewList nl{t->cur_new.slots, t->cur_new.count};
  SlotCache sc{nullptr, nullptr};
  if (t->cache_n) {  // chunks of one call fill consecutive ranges
    sc.keys = t->cache.keys + t->cache_off;
    sc.slots = t->cache.slots + t->cache_off;
    t->cache_off += n;
    t->cache_valid = t->cache_off == t->cache_n;
  }
  void* args[] = {&t->v, &keys, &n32, &out, &status_out, &nl, &sc};
  ProfScope ps(t, insert ? "find_or_insert.probe_gather" : "lookup.probe_gather", stream);
  MEEPO_CUDA_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(256), args, 0, stream));
  return MEEPO_OK;
}

meepo_status publish_slots(meepo_table* t, const uint32_t* slots, const uint32_t* cur, uint32_t* next,
                           uint64_t n_max, cudaStream_t stream) {
  const int pgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n_max + 255) / 256, (uint64_t)t->num_sms * 8));
  publish_kernel<<<pgrid, 256, 0, stream>>>(t->v, slots, cur, next);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

meepo_status probe_gather_end(meepo_table* t, uint64_t n_total, bool insert, cudaStream_t stream) {
  if (!insert || !n_total) return MEEPO_OK;
  ProfScope ps(t, "find_or_insert.publish", stream);
  return publish_slots(t, t->cur_new.slots, t->cur_new.count, t->cur_new_next, n_total, stream);
}

meepo_status launch_probe_gather(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                 uint8_t* status_out, bool insert, cudaStream_t stream) {
  MEEPO_TRY(probe_gather_begin(t, n, insert, stream));
  MEEPO_TRY(probe_gather_chunk(t, keys, n, rows_out, status_out, insert, stream));
  return probe_gather_end(t, n, insert, stream);
}

}  // namespace meepo

using namespace meepo;

static meepo_status check_batch(meepo_table* t, const void* keys, uint64_t n, const void* rows) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!keys || !rows)) return fail(MEEPO_EINVAL, "null buffer");
  return MEEPO_OK;
}

extern "C" {

MEEPO_API meepo_status meepo_find_or_insert(meepo_table* t, const uint64_t* keys, uint64_t n,
                                            void* rows_out, uint8_t* status_out, void* stream) {
  MEEPO_TRY(check_batch(t, keys, n, rows_out));
  DeviceGuard guard(t->device);
  return launch_probe_gather(t, keys, n, rows_out, status_out, true, (cudaStream_t)stream);
}

MEEPO_API meepo_status meepo_lookup(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                    uint8_t* found_out, void* stream) {
  MEEPO_TRY(check_batch(t, keys, n, rows_out));
  DeviceGuard guard(t->device);
  return launch_probe_gather(t, keys, n, rows_out, found_out, false, (cudaStream_t)stream);
}

}  // extern "C"
