// lookup.cu — find_or_insert / lookup: fused probe + row gather. The per-tile body
// lives in probe_gather.cuh; this file is the single-table kernel, the tag publish pass and the verbs.
#include "probe_gather.cuh"

namespace meepo {

// Tiles are handed out statically (warp-strided) except the last eighth, which the warps take from a ticket
// counter as they run dry: a tile with a new key or a displaced key is several dependent memory round trips
// slower than a plain hit, and with static shares alone the warps of a CTA wait at the closing barrier for their
// unluckiest sibling (ncu at 90% load: 23% of the stall samples). The last warp to finish resets the counters.
template <int CPR, bool INSERT, bool TIER>
__global__ void __launch_bounds__(256, 4) probe_gather_kernel(TableView t, const uint64_t* __restrict__ keys,
                                                           uint32_t n, uint4* __restrict__ out,
                                                           uint8_t* __restrict__ status, NewList nl, SlotCache sc,
                                                           uint32_t* __restrict__ sched) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t ntiles = (n + 31u) >> 5;
#ifdef MEEPO_AB_STATIC_TILES
  const uint32_t static_end = 0xFFFFFF00u / nwarps * nwarps;  // A/B: static shares only
#else
  const uint32_t static_end = (ntiles - ntiles / 8u) / nwarps * nwarps;
#endif
  const uint32_t cpr = CPR > 0 ? (uint32_t)CPR : t.cpr;
  TileCounts cnt;
  __shared__ uint32_t sc_slot[kScoreCells], sc_freq[kScoreCells];
  const ScoreCache scache{sc_slot, sc_freq};
  score_cache_init(t, scache);
  uint32_t tile = warp;
  while (true) {
#ifdef MEEPO_AB_STATIC_TILES
    if (tile >= ntiles) break;
#endif
    if (tile >= static_end) {
      uint32_t tk = 0;
      if (lane == 0) tk = atomicAdd(sched, 1u);
      tile = static_end + __shfl_sync(0xFFFFFFFFu, tk, 0);
      if (tile >= ntiles) break;
    }
    const uint32_t i = tile * 32u + lane;
    const uint32_t tile_keys = min(32u, n - tile * 32u);
    const uint64_t key = i < n ? __ldg(keys + i) : MEEPO_KEY_EMPTY;
    probe_gather_tile<CPR, INSERT, false, TIER>(t, key, i < n, tile_keys, out + (size_t)tile * 32u * cpr,
                                   status ? status + i : nullptr, sc.slots ? sc.slots + i : nullptr,
                                   sc.keys ? sc.keys + i : nullptr, 1u, nl.slots ? nl.slots + i : nullptr, cnt,
                                   scache, lane);
    tile += nwarps;
  }
  if (lane == 0 && atomicAdd(sched + 1, 1u) + 1u == nwarps) {  // every warp has drawn its last ticket
    sched[0] = 0;
    sched[1] = 0;
  }
  score_cache_flush(t, scache);
  flush_tile_counts(t, cnt, lane);
}

// Writes the tags of the slots claimed by the preceding probe_gather_kernel<.., true> (new_slots[i]
// != kNil) and folds their number into the table size.
__global__ void __launch_bounds__(256) publish_kernel(TableView t, const uint32_t* __restrict__ new_slots, uint32_t n) {
  uint32_t mine = 0, promoted = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t s = __ldg(new_slots + i);
    if (s == kNil) continue;
    const uint64_t key = *key_ptr(t, s);
    *tag_ptr(t, s) = (uint8_t)digest_of(mix64(key));
    mark_dirty(t, s);
    mine++;
    // a key that came back from the host tier leaves it now: the probe kernels of this call had to see it there
    if (t.tier.slabs && tier_erase(t, key)) promoted++;
  }
  mine = __reduce_add_sync(0xFFFFFFFFu, mine);
  promoted = __reduce_add_sync(0xFFFFFFFFu, promoted);
  if ((threadIdx.x & 31u) == 0 && mine) {
    atomicAdd(t.counters + C_SIZE, (unsigned long long)mine);
    if (mine > promoted) atomicAdd(t.counters + C_INSERTS, (unsigned long long)(mine - promoted));
    if (promoted) atomicAdd(t.counters + C_PROMOTIONS, (unsigned long long)promoted);
  }
}

template <bool INSERT, bool TIER>
static const void* pick_kernel(uint32_t cpr) {
  switch (cpr) {
    case 1: return (const void*)probe_gather_kernel<1, INSERT, TIER>;
    case 2: return (const void*)probe_gather_kernel<2, INSERT, TIER>;
    case 4: return (const void*)probe_gather_kernel<4, INSERT, TIER>;
    case 8: return (const void*)probe_gather_kernel<8, INSERT, TIER>;
    case 16: return (const void*)probe_gather_kernel<16, INSERT, TIER>;
    case 32: return (const void*)probe_gather_kernel<32, INSERT, TIER>;
    case 64: return (const void*)probe_gather_kernel<64, INSERT, TIER>;
    default: return (const void*)probe_gather_kernel<0, INSERT, TIER>;
  }
}
static const void* pick_kernel(bool insert, bool tier, uint32_t cpr) {
  if (tier) return insert ? pick_kernel<true, true>(cpr) : pick_kernel<false, true>(cpr);
  return insert ? pick_kernel<true, false>(cpr) : pick_kernel<false, false>(cpr);
}

// A find_or_insert / lookup call = begin, one or more chunks, end. Chunks of one call share the
// epoch and (find_or_insert) the claimed-slot list; tags are published once, in end, so a key that
// is new in the call reports INSERTED in every chunk (meepo.h "Status").
meepo_status probe_gather_begin(meepo_table* t, uint64_t n_total, bool insert, cudaStream_t stream, size_t extra_bytes) {
  t->epoch++;
  t->v.epoch = (uint32_t)t->epoch;
  t->cur_new.slots = nullptr;
  t->cur_new_off = 0;
  t->cache_valid = false;
  t->slot_gen++;
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
  if ((insert || extra_bytes) && n_total) {  // the caller takes its `extra_bytes` from the workspace afterwards
    MEEPO_TRY(t->ws.reserve(Workspace::pad(n_total * 4) + extra_bytes, stream));
    t->cur_new.slots = insert ? t->ws.take<uint32_t>(n_total) : nullptr;
  }
  return MEEPO_OK;
}

meepo_status probe_gather_chunk(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                uint8_t* status_out, bool insert, cudaStream_t stream) {
  if (n == 0) return MEEPO_OK;
  const void* kern = pick_kernel(insert, t->v.tier.slabs != 0, t->v.cpr);
  const uint64_t tiles = (n + 31) / 32;
  const int grid = grid_for(t, kern, 256, 0, (tiles + 7) / 8);
  uint32_t n32 = (uint32_t)n;
  uint4* out = reinterpret_cast<uint4*>(rows_out);
  NewList nl{insert ? t->cur_new.slots + t->cur_new_off : nullptr};  // chunks fill consecutive ranges
  t->cur_new_off += n;
  SlotCache sc{nullptr, nullptr};
  if (t->cache_n) {  // chunks of one call fill consecutive ranges
    sc.keys = t->cache.keys + t->cache_off;
    sc.slots = t->cache.slots + t->cache_off;
    t->cache_off += n;
    t->cache_valid = t->cache_off == t->cache_n;
  }
  uint32_t* sched = t->dstate->sched;
  void* args[] = {&t->v, &keys, &n32, &out, &status_out, &nl, &sc, &sched};
  ProfScope ps(t, insert ? "find_or_insert.probe_gather" : "lookup.probe_gather", stream);
  MEEPO_CUDA_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(256), args, 0, stream));
  return MEEPO_OK;
}

meepo_status publish_slots(meepo_table* t, const uint32_t* slots, uint64_t n, cudaStream_t stream) {
  if (n == 0) return MEEPO_OK;
  const int pgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)t->num_sms * 8));
  publish_kernel<<<pgrid, 256, 0, stream>>>(t->v, slots, (uint32_t)n);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

meepo_status probe_gather_end(meepo_table* t, uint64_t n_total, bool insert, cudaStream_t stream) {
  if (!insert || !n_total) return MEEPO_OK;
  ProfScope ps(t, "find_or_insert.publish", stream);
  return publish_slots(t, t->cur_new.slots, n_total, stream);
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
  VerbScope vs(t, (cudaStream_t)stream);
  MEEPO_TRY(vs.rc);
  return launch_probe_gather(t, keys, n, rows_out, status_out, true, (cudaStream_t)stream);
}

MEEPO_API meepo_status meepo_lookup(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                    uint8_t* found_out, void* stream) {
  MEEPO_TRY(check_batch(t, keys, n, rows_out));
  DeviceGuard guard(t->device);
  VerbScope vs(t, (cudaStream_t)stream);
  MEEPO_TRY(vs.rc);
  return launch_probe_gather(t, keys, n, rows_out, found_out, false, (cudaStream_t)stream);
}

}  // extern "C"
