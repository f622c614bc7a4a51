// evict.cu — capacity management (include/meepo.h "Evict"): exact selection of the
// lowest-(score,key) victims, zero-copy spill of their tuples into the pinned host tier, slot
// release without tombstones (the per-bucket overflow bits keep lookups correct), and re-admission.
//
// Selection is a 4-pass MSB-first radix select on the 32-bit score (one 256-bin histogram pass
// over the score array each, 8 B/slot), then only the candidates (score <= threshold) are
// compacted and ordered by (score, key) with two stable radix sorts. The spill copy is a kernel
// that writes 16-byte chunks straight into mapped pinned host memory over PCIe; the host only
// keeps the key -> slab index. Runs between batches, never concurrently with the probe kernels.
#include <cub/device/device_radix_sort.cuh>

#include <chrono>
#include <thread>
#include <cmath>
#include <cstring>

#include "table.h"

namespace meepo {

__device__ __forceinline__ uint32_t score_of(const TableView& t, uint32_t s, int policy) {
  const uint2 sc = t.scores[s];
  return policy == MEEPO_LFU ? sc.x : sc.y;
}

// Radix select over the composite (score, key), most significant byte first. Score passes
// (key_pass == 0): histogram of byte `shift/8` of the score over live slots whose higher score bytes
// equal `prefix`. Key passes: among the slots whose score IS the threshold `prefix`, histogram of
// byte `shift/8` of the key over those whose higher key bytes equal `kprefix` — the tie-break
// (smaller key first) is selected exactly instead of sorting tens of millions of tied candidates.
__global__ void __launch_bounds__(256) score_hist_kernel(TableView t, int policy, uint32_t prefix, uint32_t mask,
                                                         int shift, int key_pass, uint64_t kprefix, uint64_t kmask,
                                                         unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.slots; s += gridDim.x * blockDim.x) {
    const uint64_t key = *key_ptr(t, s);
    if (key == MEEPO_KEY_EMPTY) continue;
    const uint32_t sc = score_of(t, s, policy);
    if ((sc & mask) != prefix) continue;
    if (!key_pass)
      atomicAdd(&sh[(sc >> shift) & 0xFFu], 1u);
    else if ((key & kmask) == kprefix)
      atomicAdd(&sh[(uint32_t)(key >> shift) & 0xFFu], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}

// candidates = live slots with (score, key) <= (threshold, key_threshold) — exactly the victims
// (order arbitrary; sorted afterwards)
__global__ void __launch_bounds__(256) candidates_kernel(TableView t, int policy, uint32_t threshold,
                                                         uint64_t key_threshold, uint64_t* __restrict__ ckey,
                                                         uint32_t* __restrict__ cslot, uint32_t* __restrict__ count) {
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t s0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; s0 < t.slots; s0 += gridDim.x * blockDim.x) {
    const uint32_t s = s0 + lane;
    uint64_t key = MEEPO_KEY_EMPTY;
    bool take = false;
    if (s < t.slots) {
      key = *key_ptr(t, s);
      if (key != MEEPO_KEY_EMPTY) {
        const uint32_t sc = score_of(t, s, policy);
        take = sc < threshold || (sc == threshold && key <= key_threshold);
      }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, take);
    if (!m) continue;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (take) {
      const uint32_t p = base + __popc(m & ((1u << lane) - 1u));
      ckey[p] = key;
      cslot[p] = s;
    }
  }
}

__global__ void iota_kernel(uint32_t* __restrict__ v, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = i;
}
__global__ void gather_scores_kernel(TableView t, int policy, const uint32_t* __restrict__ order,
                                     const uint32_t* __restrict__ cslot, uint32_t n, uint32_t* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = score_of(t, cslot[order[i]], policy);
}
// victims in eviction order: vslot[j], vkey[j] for j < k
__global__ void victims_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ cslot,
                               const uint64_t* __restrict__ ckey, uint32_t k, uint32_t* __restrict__ vslot,
                               uint64_t* __restrict__ vkey) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    vslot[j] = cslot[order[j]];
    vkey[j] = ckey[order[j]];
  }
}

struct SpillView {  // structure-of-arrays slab in mapped pinned host memory
  uint4* rows;      // [cap][cpr]
  uint4* state;     // [cap][scpr]
  uint4* meta;      // [cap] {key lo, key hi, freq, epoch} ; step lives in steps[]
  uint32_t* steps;  // [cap]
};

// one warp per spilled victim: tuple -> host slab (zero-copy stores over PCIe)
__global__ void __launch_bounds__(256) spill_copy_kernel(TableView t, SpillView sp, const uint32_t* __restrict__ vslot,
                                                         const uint32_t* __restrict__ slab, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < n; j += nwarps) {
    const uint32_t s = vslot[j], d = slab[j];
    for (uint32_t q = lane; q < t.cpr; q += 32) sp.rows[(size_t)d * t.cpr + q] = t.rows[(size_t)s * t.cpr + q];
    for (uint32_t q = lane; q < t.scpr; q += 32) sp.state[(size_t)d * t.scpr + q] = t.state[(size_t)s * t.scpr + q];
    if (lane == 0) {
      const uint64_t key = *key_ptr(t, s);
      const uint2 sc = t.scores[s];
      sp.meta[d] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), sc.x, sc.y);
      sp.steps[d] = t.steps ? t.steps[s] : 0u;
    }
  }
}

// release the victims' slots: key -> EMPTY, tag -> 0, scores/steps -> 0 (overflow bits stay)
__global__ void release_kernel(TableView t, const uint32_t* __restrict__ vslot, uint32_t k) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    const uint32_t s = vslot[j];
    *key_ptr(t, s) = MEEPO_KEY_EMPTY;
    *tag_ptr(t, s) = 0;
    t.scores[s] = make_uint2(0, 0);
    if (t.steps) t.steps[s] = 0;
    mark_clean(t, s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(t.counters + C_SIZE, (unsigned long long)(-(long long)k));
    atomicAdd(t.counters + C_EVICTIONS, (unsigned long long)k);
  }
}

// Overflow bits after slots were released (meepo_evict): bit(b) must be set iff some live key sits past b on
// the probe path from its home bucket. Insertion only ever sets the bits, so without this pass a table that
// is filled, evicted and refilled for long enough ends up with every bit set and every miss / new key walks
// ever longer chains. Two passes over the bucket array: clear, then every displaced key re-marks its path.
__global__ void __launch_bounds__(256) overflow_clear_kernel(TableView t) {
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < t.num_buckets; b += gridDim.x * blockDim.x) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&t.buckets[b]) + 3;  // tags 12, 13 + metadata
    const uint32_t v = *w;
    if (v & (1u << 16)) *w = v & ~(1u << 16);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) t.counters[C_OVERFLOW] = 0ull;
}
__global__ void __launch_bounds__(256) overflow_mark_kernel(TableView t) {
  uint32_t fresh = 0;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.slots; s += gridDim.x * blockDim.x) {
    const uint64_t key = *key_ptr(t, s);
    if (key == MEEPO_KEY_EMPTY) continue;
    const uint32_t b = s / kBucket;
    for (uint32_t x = bucket_of(mix64(key), t.num_buckets); x != b; x = (x + 1 == t.num_buckets) ? 0 : x + 1) {
      uint32_t* w = reinterpret_cast<uint32_t*>(&t.buckets[x]) + 3;
      if (!(*reinterpret_cast<volatile uint32_t*>(w) & (1u << 16)) && !(atomicOr(w, 1u << 16) & (1u << 16))) fresh++;
    }
  }
  fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
  if ((threadIdx.x & 31u) == 0 && fresh) atomicAdd(t.counters + C_OVERFLOW, (unsigned long long)fresh);
}

// one warp per re-admitted tuple: host slab -> arena slot (zero-copy loads over PCIe)
__global__ void __launch_bounds__(256) readmit_copy_kernel(TableView t, SpillView sp, const uint32_t* __restrict__ slot,
                                                           const uint32_t* __restrict__ slab, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < n; j += nwarps) {
    const uint32_t s = slot[j], d = slab[j];
    if (s == kNil) continue;
    for (uint32_t q = lane; q < t.cpr; q += 32) t.rows[(size_t)s * t.cpr + q] = sp.rows[(size_t)d * t.cpr + q];
    for (uint32_t q = lane; q < t.scpr; q += 32) t.state[(size_t)s * t.scpr + q] = sp.state[(size_t)d * t.scpr + q];
    if (lane == 0) {
      const uint4 m = sp.meta[d];
      if (t.scores) t.scores[s] = make_uint2(m.z, m.w);
      if (t.steps) t.steps[s] = sp.steps[d];
      mark_dirty(t, s);
    }
  }
}

__global__ void found_flags_kernel(TableView t, const uint64_t* __restrict__ keys, uint32_t n,
                                   uint8_t* __restrict__ found) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    found[i] = key_valid(k) && probe_find<kReadOnly>(t, k) != kNil;
  }
}

static SpillView spill_view(const meepo_table* t) {
  SpillView sp;
  char* p = t->spill_ring;
  const uint64_t cap = t->spill_cap_tuples;
  sp.rows = reinterpret_cast<uint4*>(p);
  p += cap * t->row_bytes;
  sp.state = reinterpret_cast<uint4*>(p);
  p += cap * t->state_bytes;
  sp.meta = reinterpret_cast<uint4*>(p);
  p += cap * 16;
  sp.steps = reinterpret_cast<uint32_t*>(p);
  return sp;
}

// host bookkeeping of the spill tier; mirrors the FIFO + newest-copy-wins rule of meepo.h "Evict"
static void spill_drop(meepo_table* t, uint64_t key) {
  SpillTuple* it = t->spill_index.find(key);
  if (!it) return;
  t->spill_free.push_back((uint32_t)it->ring_index);
  t->spill_index.erase(key);
}
static uint32_t spill_push(meepo_table* t, uint64_t key) {
  spill_drop(t, key);
  while (t->spill_index.size() >= t->spill_cap_tuples) {
    auto f = t->spill_fifo.front();
    t->spill_fifo.pop_front();
    SpillTuple* it = t->spill_index.find(f.second);
    if (it && it->seq == f.first) {
      t->spill_free.push_back((uint32_t)it->ring_index);
      t->spill_index.erase(f.second);
    }
  }
  const uint32_t slab = t->spill_free.back();
  t->spill_free.pop_back();
  const uint64_t seq = t->spill_seq++;
  t->spill_index.put(key, SpillTuple{seq, slab});
  t->spill_fifo.emplace_back(seq, key);
  return slab;
}

static int grid1d(const meepo_table* t, uint64_t n) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)t->num_sms * 8));
}

}  // namespace meepo

using namespace meepo;

extern "C" {

MEEPO_API meepo_status meepo_evict(meepo_table* t, int32_t policy, double target_load, uint64_t* n_evicted,
                                   void* stream_) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (policy != MEEPO_LRU && policy != MEEPO_LFU) return fail(MEEPO_EINVAL, "bad policy");
  if (!(target_load >= 0.0 && target_load <= 1.0)) return fail(MEEPO_EINVAL, "bad target_load");
  if (!t->v.scores) return fail(MEEPO_EINVAL, "evict needs MEEPO_FLAG_TRACK_SCORES");
  DeviceGuard guard(t->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  if (n_evicted) *n_evicted = 0;
  uint64_t size = 0;
  MEEPO_TRY(live_size(t, &size));
  const uint64_t target = (uint64_t)std::floor(target_load * (double)t->v.slots);
  if (size <= target) return MEEPO_OK;
  const uint64_t k = size - target;
  t->cache_valid = false;
  t->slot_gen++;  // slots are about to change owners

  // --- radix select: threshold T = score of the k-th smallest, need `remaining` of the ties
  unsigned long long* d_hist = t->dstate->hist;
  unsigned long long h_hist[256];
  uint32_t prefix = 0, mask = 0;
  uint64_t remaining = k, ties = 0;
  const int sgrid = grid1d(t, t->v.slots);
  for (int shift = 24; shift >= 0; shift -= 8) {
    ProfScope ps(t, "evict.select(radix pass)", stream);
    MEEPO_CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof h_hist, stream));
    score_hist_kernel<<<sgrid, 256, 0, stream>>>(t->v, policy, prefix, mask, shift, 0, 0ull, 0ull, d_hist);
    MEEPO_CUDA_TRY(cudaMemcpyAsync(h_hist, d_hist, sizeof h_hist, cudaMemcpyDeviceToHost, stream));
    MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
    uint64_t cum = 0;
    int b = 0;
    for (; b < 256; b++) {
      if (cum + h_hist[b] >= remaining) break;
      cum += h_hist[b];
    }
    if (b == 256) return fail(MEEPO_ECUDA, "evict: inconsistent score histogram");
    remaining -= cum;
    ties = h_hist[b];
    prefix |= (uint32_t)b << shift;
    mask |= 0xFFu << shift;
  }
  const uint32_t T = prefix;
  // `remaining` of the `ties` slots with score == T go too: the ones with the smallest keys
  uint64_t Tkey = MEEPO_KEY_EMPTY;
  if (remaining < ties) {
    uint64_t kprefix = 0, kmask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
      ProfScope ps(t, "evict.select(radix pass)", stream);
      MEEPO_CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof h_hist, stream));
      score_hist_kernel<<<sgrid, 256, 0, stream>>>(t->v, policy, T, 0xFFFFFFFFu, shift, 1, kprefix, kmask, d_hist);
      MEEPO_CUDA_TRY(cudaMemcpyAsync(h_hist, d_hist, sizeof h_hist, cudaMemcpyDeviceToHost, stream));
      MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
      uint64_t cum = 0;
      int b = 0;
      for (; b < 256; b++) {
        if (cum + h_hist[b] >= remaining) break;
        cum += h_hist[b];
      }
      if (b == 256) return fail(MEEPO_ECUDA, "evict: inconsistent key histogram");
      remaining -= cum;
      kprefix |= (uint64_t)b << shift;
      kmask |= 0xFFull << shift;
    }
    Tkey = kprefix;
  }
  const uint64_t ncand = k;  // exactly the victims

  // --- candidates ordered by (score, key): sort by key, then stable sort by score
  size_t cub1 = 0, cub2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub1, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)ncand, 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, cub2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)ncand, 0, 32);
  const size_t need = 2 * Workspace::pad(ncand * 8) + 6 * Workspace::pad(ncand * 4) + Workspace::pad(std::max(cub1, cub2)) +
                      Workspace::pad(k * 8) + 2 * Workspace::pad(k * 4) + 4096;
  MEEPO_TRY(t->ws.reserve(need, stream));
  uint64_t* ckey = t->ws.take<uint64_t>(ncand);
  uint64_t* ckey_sorted = t->ws.take<uint64_t>(ncand);
  uint32_t* cslot = t->ws.take<uint32_t>(ncand);
  uint32_t* ord_a = t->ws.take<uint32_t>(ncand);
  uint32_t* ord_b = t->ws.take<uint32_t>(ncand);
  uint32_t* ord_c = t->ws.take<uint32_t>(ncand);
  uint32_t* sc_a = t->ws.take<uint32_t>(ncand);
  uint32_t* sc_b = t->ws.take<uint32_t>(ncand);
  char* tmp = t->ws.take<char>(std::max(cub1, cub2));
  uint64_t* vkey = t->ws.take<uint64_t>(k);
  uint32_t* vslot = t->ws.take<uint32_t>(k);
  uint32_t* vslab = t->ws.take<uint32_t>(k);
  uint32_t* d_count = &t->dstate->evict_count;
  {
  ProfScope ps(t, "evict.order_candidates(4 kernels + 2 cub sorts)", stream);
  MEEPO_CUDA_TRY(cudaMemsetAsync(d_count, 0, 4, stream));
  candidates_kernel<<<sgrid, 256, 0, stream>>>(t->v, policy, T, Tkey, ckey, cslot, d_count);
  const int cgrid = grid1d(t, ncand);
  iota_kernel<<<cgrid, 256, 0, stream>>>(ord_a, (uint32_t)ncand);
  MEEPO_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, cub1, (const uint64_t*)ckey, ckey_sorted, (const uint32_t*)ord_a,
                                                 ord_b, (int)ncand, 0, 64, stream));
  gather_scores_kernel<<<cgrid, 256, 0, stream>>>(t->v, policy, ord_b, cslot, (uint32_t)ncand, sc_a);
  MEEPO_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, cub2, (const uint32_t*)sc_a, sc_b, (const uint32_t*)ord_b, ord_c,
                                                 (int)ncand, 0, 32, stream));
  victims_kernel<<<grid1d(t, k), 256, 0, stream>>>(ord_c, cslot, ckey, (uint32_t)k, vslot, vkey);
  MEEPO_CUDA_TRY(cudaGetLastError());
  }

  // --- spill the last min(k, cap) victims (earlier ones would be pushed out by the FIFO anyway)
  if (t->spill_cap_tuples) {
    const uint64_t m = std::min<uint64_t>(k, t->spill_cap_tuples);
    std::vector<uint64_t> hk(m);
    std::vector<uint32_t> hs(m);
    MEEPO_CUDA_TRY(cudaMemcpyAsync(hk.data(), vkey + (k - m), m * 8, cudaMemcpyDeviceToHost, stream));
    MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
    if (k > m) {  // everything older is pushed out by m == cap fresh tuples
      t->spill_index.clear();
      t->spill_fifo.clear();
      t->spill_free.clear();
      for (uint64_t i = t->spill_cap_tuples; i-- > 0;) t->spill_free.push_back((uint32_t)i);
    }
    const int wgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((m + 7) / 8, (uint64_t)t->num_sms * 4));
    t->spill_index.reserve(std::min<uint64_t>(t->spill_index.size() + m, t->spill_cap_tuples));
    const bool roomy = t->spill_free.size() >= m;  // no tuple has to be pushed out to make room
    auto h0 = std::chrono::steady_clock::now();
    if (roomy) {
      // Slabs first (pops only), so that the PCIe copy can start; the index is filed underneath it. A key that
      // already has an older copy gets a fresh slab and the old one goes back to the free list afterwards —
      // which slab a tuple sits in is not observable, what the tier holds is the same as in the loop below.
      for (uint64_t j = 0; j < m; j++) {
        hs[j] = t->spill_free.back();
        t->spill_free.pop_back();
      }
    } else {
      for (uint64_t j = 0; j < m; j++) {  // the index is far larger than the host caches: fetch ahead
        if (j + 16 < m) t->spill_index.prefetch(hk[j + 16]);
        hs[j] = spill_push(t, hk[j]);
      }
    }
    MEEPO_CUDA_TRY(cudaMemcpyAsync(vslab, hs.data(), m * 4, cudaMemcpyHostToDevice, stream));
    {
      ProfScope ps(t, "evict.spill_copy(pcie)", stream);
      spill_copy_kernel<<<wgrid, 256, 0, stream>>>(t->v, spill_view(t), vslot + (k - m), vslab, (uint32_t)m);
    }
    if (roomy) {  // several host threads file the index while this one appends the FIFO entries
      const uint64_t seq0 = t->spill_seq;
      t->spill_seq += m;
      std::vector<uint32_t> freed;
      std::thread filer([&] {
        t->spill_index.replace_all(hk.data(), hs.data(), seq0, m, freed,
                                   (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2)));
      });
      for (uint64_t j = 0; j < m; j++) t->spill_fifo.emplace_back(seq0 + j, hk[j]);
      filer.join();
      t->spill_free.insert(t->spill_free.end(), freed.begin(), freed.end());
    }
    prof_add_host(t, roomy ? "evict.host_index(wall, under the spill copy)" : "evict.host_index(wall)",
                  std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count());
    MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));  // hs must outlive the copy
  }
  {
    ProfScope ps(t, "evict.release", stream);
    release_kernel<<<grid1d(t, k), 256, 0, stream>>>(t->v, vslot, (uint32_t)k);
  }
  {
    ProfScope ps(t, "evict.rebuild_overflow(2 kernels)", stream);
    overflow_clear_kernel<<<grid1d(t, t->v.num_buckets), 256, 0, stream>>>(t->v);
    overflow_mark_kernel<<<grid1d(t, t->v.slots), 256, 0, stream>>>(t->v);
  }
  MEEPO_CUDA_TRY(cudaGetLastError());
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  if (n_evicted) *n_evicted = k;
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_spill_readmit(meepo_table* t, const uint64_t* keys, uint64_t n, uint8_t* status_out) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && !keys) return fail(MEEPO_EINVAL, "null buffer");
  if (n == 0) return MEEPO_OK;
  DeviceGuard guard(t->device);
  cudaStream_t stream = nullptr;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  MEEPO_TRY(t->ws.reserve(Workspace::pad(n * 8) + Workspace::pad(n) + 4 * Workspace::pad(n * 4) + 4096, stream));
  uint64_t* d_keys = t->ws.take<uint64_t>(n);
  uint8_t* d_found = t->ws.take<uint8_t>(n);
  uint32_t* d_slot = t->ws.take<uint32_t>(n);
  uint32_t* d_slab = t->ws.take<uint32_t>(n);
  uint32_t* d_new = t->ws.take<uint32_t>(n);
  MEEPO_CUDA_TRY(cudaMemcpyAsync(d_keys, keys, n * 8, cudaMemcpyHostToDevice, stream));
  found_flags_kernel<<<grid1d(t, n), 256, 0, stream>>>(t->v, d_keys, (uint32_t)n, d_found);
  std::vector<uint8_t> found(n);
  MEEPO_CUDA_TRY(cudaMemcpyAsync(found.data(), d_found, n, cudaMemcpyDeviceToHost, stream));
  uint64_t size = 0;
  MEEPO_TRY(live_size(t, &size));  // synchronises
  std::vector<uint64_t> ins_keys;
  std::vector<uint32_t> ins_slab;
  std::unordered_map<uint64_t, int> admitted;  // keys restored earlier in this call count as present
  for (uint64_t i = 0; i < n; i++) {
    const uint64_t k = keys[i];
    uint8_t st;
    if (!key_valid(k))
      st = MEEPO_KEY_INVALID;
    else if (found[i] || admitted.count(k)) {
      st = MEEPO_KEY_FOUND;
      if (!admitted.count(k)) spill_drop(t, k);
    } else {
      SpillTuple* it = t->spill_index.find(k);
      if (!it)
        st = MEEPO_KEY_MISS;
      else if (size + ins_keys.size() >= t->v.slots)
        st = MEEPO_KEY_FULL;
      else {
        ins_keys.push_back(k);
        ins_slab.push_back((uint32_t)it->ring_index);
        t->spill_index.erase(k);  // the slab is recycled after the copy below
        admitted[k] = 1;
        st = MEEPO_KEY_INSERTED;
      }
    }
    if (status_out) status_out[i] = st;
  }
  const uint64_t m = ins_keys.size();
  if (m) {
    t->cache_valid = false;
    t->slot_gen++;
    MEEPO_CUDA_TRY(cudaMemcpyAsync(d_keys, ins_keys.data(), m * 8, cudaMemcpyHostToDevice, stream));
    MEEPO_CUDA_TRY(cudaMemcpyAsync(d_slab, ins_slab.data(), m * 4, cudaMemcpyHostToDevice, stream));
    NewList nl{d_new};
    MEEPO_TRY(import_probe_launch(t, d_keys, m, d_slot, nullptr, nl, stream));
    MEEPO_TRY(publish_slots(t, nl.slots, m, stream));
    const int wgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((m + 7) / 8, (uint64_t)t->num_sms * 4));
    readmit_copy_kernel<<<wgrid, 256, 0, stream>>>(t->v, spill_view(t), d_slot, d_slab, (uint32_t)m);
    MEEPO_CUDA_TRY(cudaGetLastError());
    MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
    for (uint32_t s : ins_slab) t->spill_free.push_back(s);
  }
  return MEEPO_OK;
}

}  // extern "C"
