// evict.cu — capacity management (include/meepo.h "Evict" / "Host tier"): exact selection of
// the lowest-(score,key) victims, hand-over of their tuples to the host tier, slot release without tombstones
// (the per-home displacement bounds keep lookups correct), and explicit re-admission.
//
// Everything data-dependent stays on the device. The host reads ONE number (the table size, which fixes the
// victim count k and with it every grid and scratch size), then enqueues:
//   select   4 histogram passes over the table on the score bytes (most significant first); a 1-CTA step kernel
//            turns each histogram into the next byte of the threshold T — no copy to the host per pass
//   split    one more pass: slots with score < T are victims (list A), slots with score == T are candidates
//            (list B: the ties — under LFU millions of keys share freq = 1)
//   tie      8 histogram passes over the KEYS OF LIST B select the key threshold, so that exactly k victims
//            remain (smaller key goes first) without sorting the ties
//   order    the k victims sorted by (score, key) with the library's own radix sort (3 x 32-bit keys)
//   tier     the last min(k, slabs) victims go to the ring: retire the slabs they overwrite from the device
//            index, file the new keys, gather the tuples into a staging buffer in HBM; a private stream then
//            drains the staging buffer into the pinned ring with plain DMA copies (the slabs of one eviction
//            are consecutive) underneath whatever the caller enqueues next
//   release  keys -> EMPTY, tags -> 0, scores / steps -> 0; the displacement bounds of the victims' homes lowered
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cstdio>
#include <memory>

#include "table.h"

namespace meepo {

__device__ __forceinline__ uint32_t score_of(const TableView& t, uint32_t s, int policy) {
  const uint2 sc = t.scores[s];
  return policy == MEEPO_LFU ? sc.x : sc.y;
}

struct SelState {  // device-side state of the selection
  uint32_t prefix, mask;           // decided bytes of the score threshold T
  uint32_t nA, nB, nV, bad;        // list lengths; victims taken from B; a histogram that did not add up
  uint32_t nC, c_over;             // list C = the candidates whose top key byte is the threshold's; it overflowed
  uint32_t ormask, done;           // OR of every live score (first pass): a byte that is 0 everywhere needs no pass;
                                   // done: the first pass already found T (it lies below 255)
  unsigned long long remaining;    // victims still to be found among the undecided slots
  unsigned long long ties;         // slots in the bin the threshold fell into
  unsigned long long kprefix, kmask;  // decided bytes of the key threshold
};

__global__ void sel_init_kernel(SelState* sel, unsigned long long k) {
  SelState s{};
  s.remaining = k;
  *sel = s;
}

// Histogram of byte `shift/8` of the score over the live slots whose higher score bytes equal the prefix. The
// first pass (shift 24) also counts min(score, 255) into a second histogram and ORs all live scores together:
// when the threshold lies below 255 — the usual LFU case, the victims are the keys seen once or twice — that one
// pass decides it and the other three return at once; otherwise a pass whose byte is zero in every score has
// nothing to count either (LRU epochs are small numbers).
__global__ void __launch_bounds__(256) score_hist_kernel(TableView t, int policy, SelState* __restrict__ sel,
                                                         int shift, unsigned long long* __restrict__ hist,
                                                         unsigned long long* __restrict__ hist_low) {
  __shared__ uint32_t sh[256], sl[256];
  const uint32_t prefix = sel->prefix, mask = sel->mask;
  const bool first = shift == 24;
  if (!first && (sel->done || ((sel->ormask >> shift) & 0xFFu) == 0)) return;
  sh[threadIdx.x] = 0;
  sl[threadIdx.x] = 0;
  __syncthreads();
  uint32_t acc_or = 0;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t s0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; s0 < t.slots; s0 += gridDim.x * blockDim.x) {
    const uint32_t s = s0 + lane;
    bool in = false;
    uint32_t bin = 0, low = 0;
    if (s < t.slots && *key_ptr(t, s) != MEEPO_KEY_EMPTY) {
      const uint32_t sc = score_of(t, s, policy);
      acc_or |= sc;
      in = (sc & mask) == prefix;
      bin = (sc >> shift) & 0xFFu;
      low = min(sc, 255u);
    }
    // scores cluster: most warps hold one bin only, and 32 shared-memory atomics on one address serialise
    const unsigned m = __ballot_sync(0xFFFFFFFFu, in);
    if (!m) continue;
    const int l0 = __ffs(m) - 1;
    const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, bin, l0);
    if (__all_sync(0xFFFFFFFFu, !in || bin == b0)) {
      if (lane == 0) atomicAdd(&sh[b0], (uint32_t)__popc(m));
    } else if (in) {
      atomicAdd(&sh[bin], 1u);
    }
    if (first) {
      const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, low, l0);
      if (__all_sync(0xFFFFFFFFu, !in || low == w0)) {
        if (lane == 0) atomicAdd(&sl[w0], (uint32_t)__popc(m));
      } else if (in) {
        atomicAdd(&sl[low], 1u);
      }
    }
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
  if (first) {
    if (sl[threadIdx.x]) atomicAdd(hist_low + threadIdx.x, (unsigned long long)sl[threadIdx.x]);
    acc_or = __reduce_or_sync(0xFFFFFFFFu, acc_or);
    if (lane == 0 && acc_or) atomicOr(&sel->ormask, acc_or);
  }
}
// The same over candidate keys (all of them have score == T) on byte `shift/8` of the key: list C (the candidates
// that share the threshold's top byte, ~1/256 of list B), or list B itself if C overflowed its allotment.
__global__ void __launch_bounds__(256) key_hist_kernel(const uint64_t* __restrict__ bkey, const uint64_t* __restrict__ ckey,
                                                       const SelState* __restrict__ sel, int shift,
                                                       unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const bool over = sel->c_over != 0;
  const uint64_t* __restrict__ src = over ? bkey : ckey;
  const uint32_t n = over ? sel->nB : sel->nC;
  const unsigned long long kprefix = sel->kprefix, kmask = sel->kmask;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = src[i];
    if ((key & kmask) == kprefix) atomicAdd(&sh[(uint32_t)(key >> shift) & 0xFFu], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}
// One radix-select step: the bin holding the `remaining`-th smallest undecided item becomes the next byte.
// hist_low (first score step only): counts of min(score, 255); if the k-th smallest score is below 255 the
// threshold is final.
__global__ void __launch_bounds__(256) sel_step_kernel(SelState* sel, unsigned long long* __restrict__ hist, int shift,
                                                       int key_pass, unsigned long long* __restrict__ hist_low) {
  if (!key_pass && shift != 24 && (sel->done || ((sel->ormask >> shift) & 0xFFu) == 0)) {  // the pass was skipped
    if (threadIdx.x == 0 && !sel->done) sel->mask |= 0xFFu << shift;                       // ... its byte is 0
    return;
  }
  if (threadIdx.x == 0) {
    unsigned long long cum = 0, rem = sel->remaining;
    int b = 0;
    bool decided = false;
    if (hist_low) {
      for (; b < 255; b++) {
        if (cum + hist_low[b] >= rem) break;
        cum += hist_low[b];
      }
      if (b < 255) {  // T = b exactly
        sel->prefix = (uint32_t)b;
        sel->mask = 0xFFFFFFFFu;
        sel->remaining = rem - cum;
        sel->ties = hist_low[b];
        sel->done = 1;
        decided = true;
      }
    }
    if (!decided) {
      cum = 0;
      for (b = 0; b < 256; b++) {
        if (cum + hist[b] >= rem) break;
        cum += hist[b];
      }
      if (b == 256) {  // cannot happen while the table is not mutated underneath
        b = 255;
        cum -= hist[255];
        sel->bad = 1;
      }
      sel->remaining = rem - cum;
      sel->ties = hist[b];
      if (key_pass) {
        sel->kprefix |= (unsigned long long)b << shift;
        sel->kmask |= 0xFFull << shift;
      } else {
        sel->prefix |= (uint32_t)b << shift;
        sel->mask |= 0xFFu << shift;
      }
    }
  }
  __syncthreads();
  hist[threadIdx.x] = 0;
  if (hist_low) hist_low[threadIdx.x] = 0;
}

// CTA-wide reservation of list space. Every thread calls it with the warp ballots of its ITEMS items in up to two
// classes; one thread adds the CTA totals to the global counters — ONE atomic per class and per 2048 items: the list
// counters are single addresses, and atomics on one address serialise in L2 (a warp-level atomicAdd per 32 slots
// cost 2.8 ms on a 60M-slot pass). Returns the list index of the first item of this warp in each class.
constexpr int kSelItems = 8;
constexpr uint32_t kSelTile = 256 * kSelItems;
struct Reserve2 {
  uint32_t a, b;
};
__device__ __forceinline__ Reserve2 block_reserve2(const unsigned (&ma)[kSelItems], const unsigned (&mb)[kSelItems],
                                                   uint32_t* counter_a, uint32_t* counter_b) {
  __shared__ uint32_t s_a[8], s_b[8], s_base[2];
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  uint32_t wa = 0, wb = 0;
#pragma unroll
  for (int k = 0; k < kSelItems; k++) {
    wa += __popc(ma[k]);
    wb += __popc(mb[k]);
  }
  __syncthreads();  // the previous round's readers are done with the shared cells
  if (lane == 0) {
    s_a[w] = wa;
    s_b[w] = wb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ta = 0, tb = 0;
    for (int i = 0; i < 8; i++) {
      const uint32_t xa = s_a[i], xb = s_b[i];
      s_a[i] = ta;
      s_b[i] = tb;
      ta += xa;
      tb += xb;
    }
    s_base[0] = ta ? atomicAdd(counter_a, ta) : 0u;
    s_base[1] = (tb && counter_b) ? atomicAdd(counter_b, tb) : 0u;
  }
  __syncthreads();
  return Reserve2{s_base[0] + s_a[w], s_base[1] + s_b[w]};
}

// Slots below the threshold -> list A (victims), slots at the threshold -> list B (candidates).
// The histogram of the candidates' top key byte (the first tie-break pass) is taken on the way.
__global__ void __launch_bounds__(256) split_kernel(TableView t, int policy, SelState* sel, uint64_t* __restrict__ akey,
                                                    uint32_t* __restrict__ aslot, uint32_t* __restrict__ ascore,
                                                    uint64_t* __restrict__ bkey, uint32_t* __restrict__ bslot,
                                                    unsigned long long* __restrict__ hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t T = sel->prefix;
  const unsigned below = (1u << lane) - 1u;
  for (uint64_t base = (uint64_t)blockIdx.x * kSelTile; base < t.slots; base += (uint64_t)gridDim.x * kSelTile) {
    uint64_t key[kSelItems];
    uint32_t sc[kSelItems];
    unsigned ma[kSelItems], mb[kSelItems];
    int cls[kSelItems];  // 1: below, 2: at the threshold
#pragma unroll
    for (int k = 0; k < kSelItems; k++) {
      const uint64_t s = base + (uint32_t)k * 256u + threadIdx.x;
      key[k] = s < t.slots ? *key_ptr(t, (uint32_t)s) : MEEPO_KEY_EMPTY;
      sc[k] = 0;
      cls[k] = 0;
      if (key[k] != MEEPO_KEY_EMPTY) {
        sc[k] = score_of(t, (uint32_t)s, policy);
        cls[k] = sc[k] < T ? 1 : (sc[k] == T ? 2 : 0);
      }
      ma[k] = __ballot_sync(0xFFFFFFFFu, cls[k] == 1);
      mb[k] = __ballot_sync(0xFFFFFFFFu, cls[k] == 2);
    }
    Reserve2 r = block_reserve2(ma, mb, &sel->nA, &sel->nB);
#pragma unroll
    for (int k = 0; k < kSelItems; k++) {
      const uint32_t s = (uint32_t)(base + (uint32_t)k * 256u + threadIdx.x);
      if (cls[k] == 1) {
        const uint32_t p = r.a + __popc(ma[k] & below);
        akey[p] = key[k];
        aslot[p] = s;
        ascore[p] = sc[k];
      } else if (cls[k] == 2) {
        const uint32_t p = r.b + __popc(mb[k] & below);
        bkey[p] = key[k];
        bslot[p] = s;
        atomicAdd(&sh[(uint32_t)(key[k] >> 56)], 1u);
      }
      r.a += __popc(ma[k]);
      r.b += __popc(mb[k]);
    }
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, (unsigned long long)sh[threadIdx.x]);
}
// List C: the candidates whose top key byte equals the threshold's (decided by the step after split_kernel).
__global__ void __launch_bounds__(256) narrow_kernel(SelState* sel, const uint64_t* __restrict__ bkey,
                                                     uint64_t* __restrict__ ckey, uint32_t ccap) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n = sel->nB;
  const uint32_t top = (uint32_t)(sel->kprefix >> 56);
  const unsigned below = (1u << lane) - 1u;
  for (uint64_t base = (uint64_t)blockIdx.x * kSelTile; base < n; base += (uint64_t)gridDim.x * kSelTile) {
    uint64_t key[kSelItems];
    unsigned m[kSelItems], none[kSelItems];
    bool take[kSelItems];
#pragma unroll
    for (int k = 0; k < kSelItems; k++) {
      const uint64_t i = base + (uint32_t)k * 256u + threadIdx.x;
      key[k] = i < n ? bkey[i] : 0ull;
      take[k] = i < n && (uint32_t)(key[k] >> 56) == top;
      m[k] = __ballot_sync(0xFFFFFFFFu, take[k]);
      none[k] = 0;
    }
    Reserve2 r = block_reserve2(m, none, &sel->nC, nullptr);
#pragma unroll
    for (int k = 0; k < kSelItems; k++) {
      if (take[k]) {
        const uint32_t p = r.a + __popc(m[k] & below);
        if (p < ccap)
          ckey[p] = key[k];
        else
          sel->c_over = 1;
      }
      r.a += __popc(m[k]);
    }
  }
}
// The candidates with key <= the key threshold join the victims (appended after list A).
__global__ void __launch_bounds__(256) take_ties_kernel(SelState* sel, const uint64_t* __restrict__ bkey,
                                                        const uint32_t* __restrict__ bslot, uint64_t* __restrict__ akey,
                                                        uint32_t* __restrict__ aslot, uint32_t* __restrict__ ascore,
                                                        uint32_t k_total) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n = sel->nB, nA = sel->nA, T = sel->prefix;
  const unsigned long long Tkey = sel->kprefix;
  const unsigned below = (1u << lane) - 1u;
  for (uint64_t base = (uint64_t)blockIdx.x * kSelTile; base < n; base += (uint64_t)gridDim.x * kSelTile) {
    uint64_t key[kSelItems];
    unsigned m[kSelItems], none[kSelItems];
    bool take[kSelItems];
#pragma unroll
    for (int k = 0; k < kSelItems; k++) {
      const uint64_t i = base + (uint32_t)k * 256u + threadIdx.x;
      key[k] = i < n ? bkey[i] : MEEPO_KEY_EMPTY;
      take[k] = i < n && key[k] <= Tkey;
      m[k] = __ballot_sync(0xFFFFFFFFu, take[k]);
      none[k] = 0;
    }
    Reserve2 r = block_reserve2(m, none, &sel->nV, nullptr);
#pragma unroll
    for (int k = 0; k < kSelItems; k++) {
      if (take[k]) {
        const uint32_t p = nA + r.a + __popc(m[k] & below);
        if (p < k_total) {
          akey[p] = key[k];
          aslot[p] = bslot[base + (uint32_t)k * 256u + threadIdx.x];
          ascore[p] = T;
        }
      }
      r.a += __popc(m[k]);
    }
  }
}

// sort keys of the three passes that order the victims by (score, key): key low word, key high word, score
__global__ void sort_key_kernel(int what, const uint64_t* __restrict__ vkey, const uint32_t* __restrict__ vscore,
                                const uint32_t* __restrict__ ord, uint32_t n, uint32_t* __restrict__ out,
                                uint32_t* __restrict__ iota) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t j = ord ? ord[i] : i;
    out[i] = what == 0 ? (uint32_t)vkey[j] : what == 1 ? (uint32_t)(vkey[j] >> 32) : vscore[j];
    if (iota) iota[i] = i;
  }
}
// victims in eviction order
__global__ void victims_kernel(const uint32_t* __restrict__ ord, const uint32_t* __restrict__ aslot,
                               const uint64_t* __restrict__ akey, uint32_t k, uint32_t* __restrict__ vslot,
                               uint64_t* __restrict__ vkey) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < k; j += gridDim.x * blockDim.x) {
    vslot[j] = aslot[ord[j]];
    vkey[j] = akey[ord[j]];
  }
}

// --- host tier: device-side index maintenance -----------------------------------------------------------
__device__ __forceinline__ uint64_t ld_key_volatile(const uint64_t* p) {
  return *reinterpret_cast<const volatile uint64_t*>(p);
}
// File `key` (known not to be in the index) -> slab: the first tombstone / unused cell of its probe path.
__device__ __forceinline__ void tier_index_put(const TierView& tv, uint64_t key, uint32_t slab) {
  uint32_t c = tier_home(tv, key);
  for (uint32_t p = 0; p <= tv.idx_mask; p++, c = (c + 1) & tv.idx_mask) {
    const uint64_t k = ld_key_volatile(tv.idx_key + c);
    if (k != MEEPO_KEY_EMPTY && k != kTomb) continue;
    if (atomicCAS(reinterpret_cast<unsigned long long*>(tv.idx_key + c), (unsigned long long)k,
                  (unsigned long long)key) == k) {
      tv.idx_val[c] = slab;
      return;
    }
    // another key took the cell: move on
  }
}
// The slabs [d0, d0 + m) mod slabs are about to be overwritten: whatever they hold leaves the index.
__global__ void __launch_bounds__(256) tier_retire_kernel(TableView t, uint32_t d0, uint32_t m) {
  const TierView& tv = t.tier;
  uint32_t gone = 0;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
    uint32_t d = d0 + j;
    if (d >= tv.slabs) d -= tv.slabs;
    const uint64_t old = tv.ring_key[d];
    if (old == MEEPO_KEY_EMPTY) continue;
    const uint32_t c = tier_cell(tv, old);
    if (c != kNil) tv.idx_key[c] = kTomb;
    tv.ring_key[d] = MEEPO_KEY_EMPTY;
    gone++;
  }
  gone = __reduce_add_sync(0xFFFFFFFFu, gone);
  if ((threadIdx.x & 31u) == 0 && gone) atomicAdd(t.counters + C_TIER_LIVE, (unsigned long long)(-(long long)gone));
}
// File the m new tuples: keys[j] -> slab (d0 + j) mod slabs. An older copy of a key gives way (its slab empties).
__global__ void __launch_bounds__(256) tier_insert_kernel(TableView t, const uint64_t* __restrict__ keys, uint32_t d0,
                                                          uint32_t m) {
  const TierView& tv = t.tier;
  uint32_t fresh = 0;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
    uint32_t d = d0 + j;
    if (d >= tv.slabs) d -= tv.slabs;
    const uint64_t key = keys[j];
    // an older copy? (cells may be claimed by other threads of this kernel meanwhile, never released)
    uint32_t c = tier_home(tv, key), hit = kNil;
    for (uint32_t p = 0; p <= tv.idx_mask; p++, c = (c + 1) & tv.idx_mask) {
      const uint64_t k = ld_key_volatile(tv.idx_key + c);
      if (k == key) {
        hit = c;
        break;
      }
      if (k == MEEPO_KEY_EMPTY) break;
    }
    if (hit != kNil) {
      tv.ring_key[tv.idx_val[hit]] = MEEPO_KEY_EMPTY;
      tv.idx_val[hit] = d;
    } else {
      tier_index_put(tv, key, d);
      fresh++;
    }
    tv.ring_key[d] = key;
  }
  fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
  if ((threadIdx.x & 31u) == 0 && fresh) atomicAdd(t.counters + C_TIER_LIVE, (unsigned long long)fresh);
}
// Index from scratch out of ring_key (tombstones gone). idx_key must be all EMPTY.
__global__ void __launch_bounds__(256) tier_rebuild_kernel(TableView t) {
  const TierView& tv = t.tier;
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < tv.slabs; d += gridDim.x * blockDim.x) {
    const uint64_t key = tv.ring_key[d];
    if (key != MEEPO_KEY_EMPTY) tier_index_put(tv, key, d);
  }
}

struct TierDst {  // where the tuples of an eviction are written: staging (j) or the ring itself (slab)
  uint4 *rows, *state, *meta;
  uint32_t* steps;
  uint32_t d0, slabs;  // slabs != 0: destination index = (d0 + j) mod slabs, else j
};
struct TierSrc {  // where the tuples come from: arena slots (eviction) or caller buffers (meepo_tier_import_buffers)
  const uint32_t* vslot;  // != null: tuple j = arena slot vslot[j]
  const uint64_t* keys;   // else: tuple j = keys[j], rows[j], state[j] (may be null: initial state), scores, steps
  const uint4 *rows, *state;
  const uint64_t* scores;
  const uint32_t* steps;
};
// one warp per tuple: source -> destination
__global__ void __launch_bounds__(256) tier_gather_kernel(TableView t, TierDst dst, TierSrc src, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < n; j += nwarps) {
    uint32_t d = j;
    if (dst.slabs) {
      d = dst.d0 + j;
      if (d >= dst.slabs) d -= dst.slabs;
    }
    if (src.vslot) {
      const uint32_t s = src.vslot[j];
      for (uint32_t q = lane; q < t.cpr; q += 32) dst.rows[(size_t)d * t.cpr + q] = ld_stream(t.rows + (size_t)s * t.cpr + q);
      for (uint32_t q = lane; q < t.scpr; q += 32)
        dst.state[(size_t)d * t.scpr + q] = ld_stream(t.state + (size_t)s * t.scpr + q);
      if (lane == 0) {
        const uint64_t key = *key_ptr(t, s);
        const uint2 sc = t.scores[s];
        dst.meta[d] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), sc.x, sc.y);
        dst.steps[d] = t.steps ? t.steps[s] : 0u;
      }
    } else {
      for (uint32_t q = lane; q < t.cpr; q += 32) dst.rows[(size_t)d * t.cpr + q] = src.rows[(size_t)j * t.cpr + q];
      const uint4 s0 = init_state_chunk(t);
      for (uint32_t q = lane; q < t.scpr; q += 32)
        dst.state[(size_t)d * t.scpr + q] = src.state ? src.state[(size_t)j * t.scpr + q] : s0;
      if (lane == 0) {
        const uint64_t key = src.keys[j];
        const uint64_t sc = src.scores ? src.scores[j] : 0ull;
        dst.meta[d] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)sc, (uint32_t)(sc >> 32));
        dst.steps[d] = src.steps ? src.steps[j] : 0u;
      }
    }
  }
}

// release the victims' slots: key -> EMPTY, tag -> 0, scores/steps -> 0 (displacement bounds: overflow_fix_kernel)
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

// Displacement metadata after slots were released (meepo_evict): disp(h) must stay the largest displacement of a
// LIVE key with home h — insertion only ever raises it, so without maintenance a table that is filled, evicted
// and refilled for long enough walks ever longer on every miss. Only the homes of victims that sat past their
// home bucket can change, and only downwards: one thread per such victim rescans the buckets h .. h + disp(h)
// for the keys whose home is h and lowers the bound to what it finds (every thread of one home computes the same
// value; the CAS winner that takes a home to 0 counts it out of overflow_buckets). Round 1 and the first half of
// round 2 rebuilt the metadata of the WHOLE table after every eviction (clear pass + mark pass, 0.46 ms at 67M
// slots); this is ~40K local scans for 400K victims.
__global__ void __launch_bounds__(256) overflow_fix_kernel(TableView t, const uint64_t* __restrict__ vkey,
                                                           const uint32_t* __restrict__ vslot, uint32_t k) {
  // one WARP per victim: lane l scans buckets home + 1 + l, home + 33 + l, ... (a thread per victim left the kernel
  // waiting for the few homes whose bound is in the hundreds: 0.32 ms)
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  uint32_t gone = 0;
  for (uint32_t j = warp; j < k; j += nwarps) {
    const uint32_t b = vslot[j] / kBucket;
    const uint32_t home = bucket_of(mix64(vkey[j]), t.num_buckets);
    if (b == home) continue;
    uint32_t* w = reinterpret_cast<uint32_t*>(&t.buckets[home]) + 3;
    const uint32_t cur = __shfl_sync(0xFFFFFFFFu, *reinterpret_cast<volatile uint32_t*>(w), 0);
    const uint32_t d_old = cur >> 16;
    if (d_old == 0) continue;  // a sibling already settled this home
    const uint32_t last = d_old == kDispUnknown ? t.num_buckets - 1 : min(d_old, t.num_buckets - 1);
    uint32_t dmax = 0;
    for (uint32_t d = 1 + lane; d <= last; d += 32) {
      const uint32_t x = (uint32_t)(((uint64_t)home + d) % t.num_buckets);
      for (uint32_t i = 0; i < kBucket; i++) {
        const uint64_t key = t.buckets[x].key[i];
        if (key != MEEPO_KEY_EMPTY && bucket_of(mix64(key), t.num_buckets) == home) dmax = d;
      }
    }
    dmax = __reduce_max_sync(0xFFFFFFFFu, dmax);
    if (lane == 0 && dmax < d_old && atomicCAS(w, cur, (cur & 0xFFFFu) | (dmax << 16)) == cur && dmax == 0) gone++;
  }
  gone = __reduce_add_sync(0xFFFFFFFFu, gone);
  if (lane == 0 && gone) atomicAdd(t.counters + C_OVERFLOW, (unsigned long long)(-(long long)gone));
}

// --- meepo_spill_readmit ----------------------------------------------------------------------------------
// Thread per key: FOUND (in HBM), MISS (in neither level), or claim a slot for a key of the tier (INSERTED for
// every duplicate, FULL). slot_out[i] != kNil only for the thread that has to copy the tuple.
__global__ void __launch_bounds__(256) readmit_probe_kernel(TableView t, const uint64_t* __restrict__ keys, uint32_t n,
                                                            uint8_t* __restrict__ status, uint32_t* __restrict__ slot_out,
                                                            uint32_t* __restrict__ slab_out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    uint8_t st = MEEPO_KEY_INVALID;
    uint32_t slot = kNil, slab = kNil;
    if (key_valid(k)) {
      if (probe_find<kCoherent>(t, k) != kNil) {
        st = MEEPO_KEY_FOUND;
      } else if ((slab = tier_slab(t.tier, k)) == kNil) {
        st = MEEPO_KEY_MISS;
      } else {
        const Probe pr = probe_find_or_insert(t, k);
        if (pr.status == MEEPO_KEY_FULL) {
          st = MEEPO_KEY_FULL;
          atomicAdd(t.counters + C_FULL, 1ull);
        } else {
          st = MEEPO_KEY_INSERTED;
          if (pr.winner) slot = pr.slot;
        }
      }
    }
    status[i] = st;
    slot_out[i] = slot;
    slab_out[i] = slab;
  }
}
// one warp per restored tuple: tier (staging or ring, zero-copy over PCIe) -> arena slot, scores as they were
__global__ void __launch_bounds__(256) readmit_copy_kernel(TableView t, const uint32_t* __restrict__ slot,
                                                           const uint32_t* __restrict__ slab, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < n; j += nwarps) {
    const uint32_t s = slot[j];
    if (s == kNil) continue;
    const TierTuple tt = tier_tuple(t, slab[j]);
    for (uint32_t q = lane; q < t.cpr; q += 32) t.rows[(size_t)s * t.cpr + q] = tt.rows[q];
    for (uint32_t q = lane; q < t.scpr; q += 32) t.state[(size_t)s * t.scpr + q] = tt.state[q];
    if (lane == 0) {
      const uint4 m = *tt.meta;
      if (t.scores) t.scores[s] = make_uint2(m.z, m.w);
      if (t.steps) t.steps[s] = *tt.steps;
    }
  }
}
// keys that were already in HBM: a tier copy, if there is one, is dropped
__global__ void __launch_bounds__(256) readmit_drop_kernel(TableView t, const uint64_t* __restrict__ keys,
                                                           const uint8_t* __restrict__ status, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (status[i] == MEEPO_KEY_FOUND) tier_erase(t, keys[i]);
}

// --- the tier's side of a checkpoint -----------------------------------------------------------------------
// live slabs of the ring -> (key, slab), any order (sorted by key afterwards)
__global__ void __launch_bounds__(256) tier_live_kernel(TableView t, uint64_t* __restrict__ keys, uint32_t* __restrict__ slabs,
                                                        uint32_t* __restrict__ count) {
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t d0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; d0 < t.tier.slabs; d0 += gridDim.x * blockDim.x) {
    const uint32_t d = d0 + lane;
    const uint64_t key = d < t.tier.slabs ? t.tier.ring_key[d] : MEEPO_KEY_EMPTY;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, key != MEEPO_KEY_EMPTY);
    if (!m) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (key != MEEPO_KEY_EMPTY) {
      const uint32_t p = base + __popc(m & ((1u << lane) - 1u));
      keys[p] = key;
      slabs[p] = d;
    }
  }
}
// one warp per exported tuple: tier (staging or ring) -> caller buffers (each may be null)
__global__ void __launch_bounds__(256) tier_export_kernel(TableView t, const uint32_t* __restrict__ slabs, uint32_t n,
                                                          uint4* __restrict__ rows, uint4* __restrict__ state,
                                                          uint64_t* __restrict__ scores, uint32_t* __restrict__ steps) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < n; j += nwarps) {
    const TierTuple tt = tier_tuple(t, slabs[j]);
    if (rows)
      for (uint32_t q = lane; q < t.cpr; q += 32) rows[(size_t)j * t.cpr + q] = tt.rows[q];
    if (state)
      for (uint32_t q = lane; q < t.scpr; q += 32) state[(size_t)j * t.scpr + q] = tt.state[q];
    if (lane == 0) {
      const uint4 m = *tt.meta;
      if (scores) scores[j] = ((uint64_t)m.w << 32) | m.z;
      if (steps) steps[j] = *tt.steps;
    }
  }
}

static int grid1d(const meepo_table* t, uint64_t n) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)t->num_sms * 8));
}
static int gridwarp(const meepo_table* t, uint64_t rows) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((rows + 7) / 8, (uint64_t)t->num_sms * 8));
}
static int gridsel(const meepo_table* t, uint64_t n) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((n + kSelTile - 1) / kSelTile, (uint64_t)t->num_sms * 4));
}

// ring layout: rows[slabs], state[slabs], meta[slabs], steps[slabs] (structure of arrays, so that the slabs of
// one eviction are four contiguous runs)
static void tier_carve(const meepo_table* t, char* p, uint64_t cap, uint4** rows, uint4** state, uint4** meta,
                       uint32_t** steps) {
  *rows = reinterpret_cast<uint4*>(p);
  p += cap * t->row_bytes;
  *state = reinterpret_cast<uint4*>(p);
  p += cap * t->state_bytes;
  *meta = reinterpret_cast<uint4*>(p);
  p += cap * 16;
  *steps = reinterpret_cast<uint32_t*>(p);
}

meepo_status tier_create(meepo_table* t) {
  const uint64_t slabs = t->cfg.host_spill_bytes / t->tuple_bytes();
  if (slabs == 0) return MEEPO_OK;
  if (slabs > (1ull << 30)) return fail(MEEPO_EINVAL, "host tier: more than 2^30 slabs");
  TierView& tv = t->v.tier;
  MEEPO_CUDA_TRY(cudaHostAlloc(&t->spill_ring, slabs * t->tuple_bytes(), cudaHostAllocMapped | cudaHostAllocPortable));
  t->spill_cap_tuples = slabs;
  tier_carve(t, t->spill_ring, slabs, &tv.h_rows, &tv.h_state, &tv.h_meta, &tv.h_steps);
  uint64_t cells = 1024;
  while (cells < 2 * slabs) cells <<= 1;
  MEEPO_CUDA_TRY(cudaMalloc(&tv.idx_key, cells * 8));
  MEEPO_CUDA_TRY(cudaMalloc(&tv.idx_val, cells * 4));
  MEEPO_CUDA_TRY(cudaMalloc(&tv.ring_key, slabs * 8));
  MEEPO_CUDA_TRY(cudaMemset(tv.idx_key, 0xFF, cells * 8));
  MEEPO_CUDA_TRY(cudaMemset(tv.ring_key, 0xFF, slabs * 8));
  tv.idx_mask = (uint32_t)(cells - 1);
  tv.slabs = (uint32_t)slabs;
  tv.stage_d0 = tv.stage_m = 0;
  MEEPO_CUDA_TRY(cudaStreamCreateWithFlags(&t->tier_stream, cudaStreamNonBlocking));
  MEEPO_CUDA_TRY(cudaEventCreateWithFlags(&t->tier_staged, cudaEventDisableTiming));
  MEEPO_CUDA_TRY(cudaEventCreateWithFlags(&t->tier_drained, cudaEventDisableTiming));
  return MEEPO_OK;
}

void tier_destroy(meepo_table* t) {
  if (t->tier_stream) {
    cudaStreamSynchronize(t->tier_stream);
    cudaStreamDestroy(t->tier_stream);
  }
  if (t->tier_staged) cudaEventDestroy(t->tier_staged);
  if (t->tier_drained) cudaEventDestroy(t->tier_drained);
  cudaFree(t->v.tier.idx_key);
  cudaFree(t->v.tier.idx_val);
  cudaFree(t->v.tier.ring_key);
  cudaFree(t->tier_stage);
  if (t->spill_ring) cudaFreeHost(t->spill_ring);
  t->spill_ring = nullptr;
}

// Hands the victims vslot[j0 .. j0 + m) / vkey[..] (eviction order) to the ring. head = tuples appended before
// this eviction, k = victims of this eviction (the first k - m would be overwritten by the later ones anyway).
static meepo_status tier_append(meepo_table* t, TierSrc src, const uint64_t* vkey, uint64_t k, cudaStream_t stream) {
  TierView& tv = t->v.tier;
  const uint64_t slabs = tv.slabs;
  const uint64_t m = std::min<uint64_t>(k, slabs), j0 = k - m;
  const uint32_t d0 = (uint32_t)((t->tier_head + j0) % slabs);
  // tombstones pile up in the index: rebuild it from ring_key before it could run out of unused cells
  const uint64_t cells = (uint64_t)tv.idx_mask + 1;
  if (t->tier_nonempty_ub + m > cells / 4 * 3) {
    ProfScope ps(t, "evict.tier_index_rebuild", stream);
    MEEPO_CUDA_TRY(cudaMemsetAsync(tv.idx_key, 0xFF, cells * 8, stream));
    tier_rebuild_kernel<<<grid1d(t, slabs), 256, 0, stream>>>(t->v);
    t->tier_nonempty_ub = std::min<uint64_t>(slabs, t->tier_head);
  }
  {
    ProfScope ps(t, "evict.tier_index(2 kernels)", stream);
    tier_retire_kernel<<<grid1d(t, m), 256, 0, stream>>>(t->v, d0, (uint32_t)m);
    tier_insert_kernel<<<grid1d(t, m), 256, 0, stream>>>(t->v, vkey + j0, d0, (uint32_t)m);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  t->tier_nonempty_ub += m;
  // the previous eviction's drain must be over before its staging buffer is reused
  if (t->tier_draining) MEEPO_CUDA_TRY(cudaStreamWaitEvent(stream, t->tier_drained, 0));
  static const uint64_t stage_limit = [] {
    const char* e = getenv("MEEPO_TIER_STAGE_MAX_BYTES");
    return e ? strtoull(e, nullptr, 10) : (8ull << 30);
  }();
  const bool staged = m * t->tuple_bytes() <= stage_limit;
  if (staged && m > t->tier_stage_cap) {
    if (t->tier_draining) MEEPO_CUDA_TRY(cudaEventSynchronize(t->tier_drained));
    MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
    cudaFree(t->tier_stage);
    t->tier_stage = nullptr;
    t->tier_stage_cap = 0;
    const uint64_t cap = m + m / 4;
    MEEPO_CUDA_TRY(cudaMalloc(&t->tier_stage, cap * t->tuple_bytes()));
    t->tier_stage_cap = cap;
    tier_carve(t, t->tier_stage, cap, &tv.s_rows, &tv.s_state, &tv.s_meta, &tv.s_steps);
  }
  TierDst dst;
  if (staged) {
    dst = TierDst{tv.s_rows, tv.s_state, tv.s_meta, tv.s_steps, 0, 0};
  } else {
    dst = TierDst{tv.h_rows, tv.h_state, tv.h_meta, tv.h_steps, d0, (uint32_t)slabs};
  }
  {
    ProfScope ps(t, staged ? "evict.gather_to_staging" : "evict.spill_copy(pcie, zero-copy)", stream);
    if (src.vslot) {
      src.vslot += j0;
    } else {
      src.keys += j0;
      src.rows += j0 * t->v.cpr;
      if (src.state) src.state += j0 * t->v.scpr;
      if (src.scores) src.scores += j0;
      if (src.steps) src.steps += j0;
    }
    tier_gather_kernel<<<gridwarp(t, m), 256, 0, stream>>>(t->v, dst, src, (uint32_t)m);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  tv.stage_d0 = staged ? d0 : 0;
  tv.stage_m = staged ? (uint32_t)m : 0;
  if (staged) {  // drain: the slabs are consecutive (mod slabs), so each array is one or two plain copies
    MEEPO_CUDA_TRY(cudaEventRecord(t->tier_staged, stream));
    MEEPO_CUDA_TRY(cudaStreamWaitEvent(t->tier_stream, t->tier_staged, 0));
    const uint64_t first = std::min<uint64_t>(m, slabs - d0);
    auto copy = [&](char* ring, const char* stage, uint64_t width) -> cudaError_t {
      if (!width) return cudaSuccess;
      cudaError_t e = cudaMemcpyAsync(ring + (uint64_t)d0 * width, stage, first * width, cudaMemcpyDeviceToHost, t->tier_stream);
      if (e == cudaSuccess && first < m)
        e = cudaMemcpyAsync(ring, stage + first * width, (m - first) * width, cudaMemcpyDeviceToHost, t->tier_stream);
      return e;
    };
    MEEPO_CUDA_TRY(copy((char*)tv.h_rows, (const char*)tv.s_rows, t->row_bytes));
    MEEPO_CUDA_TRY(copy((char*)tv.h_state, (const char*)tv.s_state, t->state_bytes));
    MEEPO_CUDA_TRY(copy((char*)tv.h_meta, (const char*)tv.s_meta, 16));
    MEEPO_CUDA_TRY(copy((char*)tv.h_steps, (const char*)tv.s_steps, 4));
    MEEPO_CUDA_TRY(cudaEventRecord(t->tier_drained, t->tier_stream));
    t->tier_draining = true;
  }
  t->tier_head += k;
  return MEEPO_OK;
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
  {  // the one host synchronisation of the call: with `stream` only (a drain of the previous eviction goes on)
    unsigned long long v = 0;
    MEEPO_CUDA_TRY(cudaMemcpyAsync(&v, t->dstate->counters + C_SIZE, 8, cudaMemcpyDeviceToHost, stream));
    MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
    size = v;
  }
  const uint64_t target = (uint64_t)std::floor(target_load * (double)t->v.slots);
  if (size <= target) return MEEPO_OK;
  const uint64_t k = size - target;
  t->cache_valid = false;
  t->slot_gen++;  // slots are about to change owners

  const size_t sort_tmp = radix_sort_temp_bytes(k, 32);
  const size_t need = Workspace::pad(sizeof(SelState)) + Workspace::pad(k * 8) + 2 * Workspace::pad(k * 4) +  // list A
                      Workspace::pad(size * 8) + Workspace::pad(size * 4) + Workspace::pad((size / 8 + 4096) * 8) +  // lists B, C
                      4 * Workspace::pad(k * 4) + Workspace::pad(sort_tmp) +                                 // sort
                      Workspace::pad(k * 8) + Workspace::pad(k * 4) + Workspace::pad(256 * 8) + 8192;
  MEEPO_TRY(t->ws.reserve(need, stream));
  SelState* sel = t->ws.take<SelState>(1);
  uint64_t* akey = t->ws.take<uint64_t>(k);
  uint32_t* aslot = t->ws.take<uint32_t>(k);
  uint32_t* ascore = t->ws.take<uint32_t>(k);
  uint64_t* bkey = t->ws.take<uint64_t>(size);
  uint32_t* bslot = t->ws.take<uint32_t>(size);
  const uint64_t ccap = size / 8 + 4096;
  uint64_t* ckey = t->ws.take<uint64_t>(ccap);
  uint32_t* sk_a = t->ws.take<uint32_t>(k);
  uint32_t* sk_b = t->ws.take<uint32_t>(k);
  uint32_t* ord_a = t->ws.take<uint32_t>(k);
  uint32_t* ord_b = t->ws.take<uint32_t>(k);
  char* tmp = t->ws.take<char>(sort_tmp);
  uint64_t* vkey = t->ws.take<uint64_t>(k);
  uint32_t* vslot = t->ws.take<uint32_t>(k);
  unsigned long long* d_hist = t->dstate->hist;
  unsigned long long* d_hist_low = t->ws.take<unsigned long long>(256);
  const int sgrid = grid1d(t, t->v.slots);

  {  // --- threshold T = score of the k-th smallest (score, key)
    ProfScope ps(t, "evict.select(4 table passes)", stream);
    sel_init_kernel<<<1, 1, 0, stream>>>(sel, k);
    MEEPO_CUDA_TRY(cudaMemsetAsync(d_hist, 0, 256 * 8, stream));
    MEEPO_CUDA_TRY(cudaMemsetAsync(d_hist_low, 0, 256 * 8, stream));
    for (int shift = 24; shift >= 0; shift -= 8) {
      score_hist_kernel<<<sgrid, 256, 0, stream>>>(t->v, policy, sel, shift, d_hist, d_hist_low);
      sel_step_kernel<<<1, 256, 0, stream>>>(sel, d_hist, shift, 0, shift == 24 ? d_hist_low : nullptr);
    }
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  {  // --- victims below T, candidates at T; the key threshold among the candidates
    {
      ProfScope ps(t, "evict.split(1 table pass)", stream);
      split_kernel<<<gridsel(t, t->v.slots), 256, 0, stream>>>(t->v, policy, sel, akey, aslot, ascore, bkey, bslot, d_hist);
    }
    ProfScope ps(t, "evict.ties(narrow + 7 list passes + take)", stream);
    sel_step_kernel<<<1, 256, 0, stream>>>(sel, d_hist, 56, 1, nullptr);
    narrow_kernel<<<gridsel(t, size), 256, 0, stream>>>(sel, bkey, ckey, (uint32_t)ccap);
    const int cgrid = grid1d(t, ccap);
    for (int shift = 48; shift >= 0; shift -= 8) {
      key_hist_kernel<<<cgrid, 256, 0, stream>>>(bkey, ckey, sel, shift, d_hist);
      sel_step_kernel<<<1, 256, 0, stream>>>(sel, d_hist, shift, 1, nullptr);
    }
    take_ties_kernel<<<gridsel(t, size), 256, 0, stream>>>(sel, bkey, bslot, akey, aslot, ascore, (uint32_t)k);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  {  // --- eviction order: ascending (score, key)
    ProfScope ps(t, "evict.order(3 radix sorts)", stream);
    const int kgrid = grid1d(t, k);
    if (!radix_sort_supported(k, 32)) return fail(MEEPO_EINVAL, "evict: too many victims for one call");
    sort_key_kernel<<<kgrid, 256, 0, stream>>>(0, akey, ascore, nullptr, (uint32_t)k, sk_a, ord_a);
    MEEPO_TRY(radix_sort_pairs(t, tmp, sk_a, sk_b, ord_a, ord_b, (uint32_t)k, 32, stream));
    sort_key_kernel<<<kgrid, 256, 0, stream>>>(1, akey, ascore, ord_b, (uint32_t)k, sk_a, nullptr);
    MEEPO_TRY(radix_sort_pairs(t, tmp, sk_a, sk_b, ord_b, ord_a, (uint32_t)k, 32, stream));
    sort_key_kernel<<<kgrid, 256, 0, stream>>>(2, akey, ascore, ord_a, (uint32_t)k, sk_a, nullptr);
    MEEPO_TRY(radix_sort_pairs(t, tmp, sk_a, sk_b, ord_a, ord_b, (uint32_t)k, 32, stream));
    victims_kernel<<<kgrid, 256, 0, stream>>>(ord_b, aslot, akey, (uint32_t)k, vslot, vkey);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  if (t->v.tier.slabs) MEEPO_TRY(tier_append(t, TierSrc{vslot, nullptr, nullptr, nullptr, nullptr, nullptr}, vkey, k, stream));
  {
    ProfScope ps(t, "evict.release", stream);
    release_kernel<<<grid1d(t, k), 256, 0, stream>>>(t->v, vslot, (uint32_t)k);
  }
  {
    ProfScope ps(t, "evict.fix_displacements", stream);
    overflow_fix_kernel<<<gridwarp(t, k), 256, 0, stream>>>(t->v, vkey, vslot, (uint32_t)k);
  }
  MEEPO_CUDA_TRY(cudaGetLastError());
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
  MEEPO_TRY(t->ws.reserve(Workspace::pad(n * 8) + Workspace::pad(n) + 2 * Workspace::pad(n * 4) + 4096, stream));
  uint64_t* d_keys = t->ws.take<uint64_t>(n);
  uint8_t* d_status = t->ws.take<uint8_t>(n);
  uint32_t* d_slot = t->ws.take<uint32_t>(n);
  uint32_t* d_slab = t->ws.take<uint32_t>(n);
  MEEPO_CUDA_TRY(cudaMemcpyAsync(d_keys, keys, n * 8, cudaMemcpyHostToDevice, stream));
  t->cache_valid = false;
  t->slot_gen++;
  readmit_probe_kernel<<<grid1d(t, n), 256, 0, stream>>>(t->v, d_keys, (uint32_t)n, d_status, d_slot, d_slab);
  if (t->v.tier.slabs) {  // without a tier every key is FOUND, MISS or INVALID and nothing moves
    readmit_copy_kernel<<<gridwarp(t, n), 256, 0, stream>>>(t->v, d_slot, d_slab, (uint32_t)n);
    readmit_drop_kernel<<<grid1d(t, n), 256, 0, stream>>>(t->v, d_keys, d_status, (uint32_t)n);
    MEEPO_CUDA_TRY(cudaGetLastError());
    MEEPO_TRY(publish_slots(t, d_slot, n, stream));  // tags, size, and the restored keys leave the tier
  }
  MEEPO_CUDA_TRY(cudaGetLastError());
  if (status_out) MEEPO_CUDA_TRY(cudaMemcpyAsync(status_out, d_status, n, cudaMemcpyDeviceToHost, stream));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  return MEEPO_OK;
}


// --- the host tier's side of a checkpoint (include/meepo.h "tier dump / load") -----------------------------
}  // extern "C"

namespace {
struct TierFileHeader {  // the MEEPOTB1 header (io.cu)
  char magic[8];
  uint32_t version, dim, dtype, opt;
  uint64_t n, row_bytes, state_bytes, epoch;
};
static_assert(sizeof(TierFileHeader) == 56, "header layout");

// Live tier tuples as (key, slab) sorted by key, left in the workspace. Synchronises.
meepo_status tier_sorted_live(meepo_table* t, uint64_t* n_out, uint64_t** keys_sorted, uint32_t** slabs_sorted,
                              cudaStream_t stream) {
  unsigned long long live = 0;
  MEEPO_CUDA_TRY(cudaDeviceSynchronize());
  MEEPO_CUDA_TRY(cudaMemcpy(&live, t->dstate->counters + C_TIER_LIVE, 8, cudaMemcpyDeviceToHost));
  *n_out = live;
  if (live == 0) return MEEPO_OK;
  const uint64_t n = live;
  if (!radix_sort_supported(n, 32)) return fail(MEEPO_EINVAL, "tier export: too many tuples");
  const size_t tmp_bytes = radix_sort_temp_bytes(n, 32);
  MEEPO_TRY(t->ws.reserve(2 * Workspace::pad(n * 8) + 6 * Workspace::pad(n * 4) + Workspace::pad(tmp_bytes) + 4096, stream));
  uint64_t* k_in = t->ws.take<uint64_t>(n);
  uint64_t* k_out = t->ws.take<uint64_t>(n);
  uint32_t* s_in = t->ws.take<uint32_t>(n);
  uint32_t* s_out = t->ws.take<uint32_t>(n);
  uint32_t* sk_a = t->ws.take<uint32_t>(n);
  uint32_t* sk_b = t->ws.take<uint32_t>(n);
  uint32_t* ord_a = t->ws.take<uint32_t>(n);
  uint32_t* ord_b = t->ws.take<uint32_t>(n);
  char* tmp = t->ws.take<char>(tmp_bytes);
  uint32_t* count = &t->dstate->evict_count;
  MEEPO_CUDA_TRY(cudaMemsetAsync(count, 0, 4, stream));
  tier_live_kernel<<<grid1d(t, t->v.tier.slabs), 256, 0, stream>>>(t->v, k_in, s_in, count);
  const int g = grid1d(t, n);
  sort_key_kernel<<<g, 256, 0, stream>>>(0, k_in, nullptr, nullptr, (uint32_t)n, sk_a, ord_a);
  MEEPO_TRY(radix_sort_pairs(t, tmp, sk_a, sk_b, ord_a, ord_b, (uint32_t)n, 32, stream));
  sort_key_kernel<<<g, 256, 0, stream>>>(1, k_in, nullptr, ord_b, (uint32_t)n, sk_a, nullptr);
  MEEPO_TRY(radix_sort_pairs(t, tmp, sk_a, sk_b, ord_b, ord_a, (uint32_t)n, 32, stream));
  victims_kernel<<<g, 256, 0, stream>>>(ord_a, s_in, k_in, (uint32_t)n, s_out, k_out);
  MEEPO_CUDA_TRY(cudaGetLastError());
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  *keys_sorted = k_out;
  *slabs_sorted = s_out;
  return MEEPO_OK;
}
}  // namespace

extern "C" {

MEEPO_API meepo_status meepo_tier_export_buffers(meepo_table* t, uint64_t* keys, void* rows, void* state, uint64_t* scores,
                                                 uint32_t* steps, uint64_t max_n, uint64_t* n_out) {
  if (!t || !n_out) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  VerbScope vs(t, nullptr);
  MEEPO_TRY(vs.rc);
  *n_out = 0;
  if (!t->v.tier.slabs) return MEEPO_OK;
  cudaStream_t stream = nullptr;
  uint64_t n = 0;
  uint64_t* ks = nullptr;
  uint32_t* ss = nullptr;
  if (!keys) {  // size query
    unsigned long long live = 0;
    MEEPO_CUDA_TRY(cudaDeviceSynchronize());
    MEEPO_CUDA_TRY(cudaMemcpy(&live, t->dstate->counters + C_TIER_LIVE, 8, cudaMemcpyDeviceToHost));
    *n_out = live;
    return MEEPO_OK;
  }
  MEEPO_TRY(tier_sorted_live(t, &n, &ks, &ss, stream));
  *n_out = n;
  if (n == 0) return MEEPO_OK;
  if (max_n < n) return fail(MEEPO_EINVAL, "export buffers too small");
  MEEPO_CUDA_TRY(cudaMemcpyAsync(keys, ks, n * 8, cudaMemcpyDeviceToDevice, stream));
  tier_export_kernel<<<gridwarp(t, n), 256, 0, stream>>>(t->v, ss, (uint32_t)n, reinterpret_cast<uint4*>(rows),
                                                         t->v.scpr ? reinterpret_cast<uint4*>(state) : nullptr, scores, steps);
  MEEPO_CUDA_TRY(cudaGetLastError());
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_tier_import_buffers(meepo_table* t, const uint64_t* keys, const void* rows, const void* state,
                                                 const uint64_t* scores, const uint32_t* steps, uint64_t n) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n && (!keys || !rows)) return fail(MEEPO_EINVAL, "null buffer");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n == 0) return MEEPO_OK;
  if (!t->v.tier.slabs) return fail(MEEPO_EINVAL, "the table has no host tier (host_spill_bytes == 0)");
  DeviceGuard guard(t->device);
  cudaStream_t stream = nullptr;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  TierSrc src{nullptr, keys, reinterpret_cast<const uint4*>(rows), reinterpret_cast<const uint4*>(state), scores, steps};
  MEEPO_TRY(tier_append(t, src, keys, n, stream));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_tier_export(meepo_table* t, const char* path) {
  if (!t || !path) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  VerbScope vs(t, nullptr);
  MEEPO_TRY(vs.rc);
  cudaStream_t stream = nullptr;
  uint64_t n = 0;
  uint64_t* ks = nullptr;
  uint32_t* ss = nullptr;
  if (t->v.tier.slabs) MEEPO_TRY(tier_sorted_live(t, &n, &ks, &ss, stream));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(MEEPO_EIO, std::string("cannot open ") + path);
  std::unique_ptr<FILE, int (*)(FILE*)> closer(f, fclose);
  TierFileHeader h{};
  memcpy(h.magic, "MEEPOTB1", 8);
  h.version = 1;
  h.dim = t->cfg.dim;
  h.dtype = (uint32_t)t->cfg.dtype;
  h.opt = (uint32_t)t->cfg.opt;
  h.n = n;
  h.row_bytes = t->row_bytes;
  h.state_bytes = t->state_bytes;
  h.epoch = t->epoch;
  if (fwrite(&h, sizeof h, 1, f) != 1) return fail(MEEPO_EIO, "short write");
  if (n == 0) return MEEPO_OK;
  const uint64_t chunk = 1u << 16;
  const size_t widest = std::max<size_t>({(size_t)t->row_bytes, (size_t)t->state_bytes, 8});
  char *d_stage = nullptr, *h_stage = nullptr;
  MEEPO_CUDA_TRY(cudaMalloc(&d_stage, chunk * widest));
  if (cudaHostAlloc(&h_stage, chunk * widest, cudaHostAllocDefault) != cudaSuccess) {
    cudaFree(d_stage);
    return fail(MEEPO_ENOMEM, "cudaHostAlloc(export bounce)");
  }
  meepo_status rc = MEEPO_OK;
  auto section = [&](int what, size_t width) {  // sections in file order: keys, rows, state, scores, steps
    for (uint64_t lo = 0; lo < n && rc == MEEPO_OK; lo += chunk) {
      const uint64_t m = std::min(chunk, n - lo);
      const void* src = d_stage;
      uint4* st = reinterpret_cast<uint4*>(d_stage);
      if (what == 0)
        src = ks + lo;
      else
        tier_export_kernel<<<gridwarp(t, m), 256, 0, stream>>>(
            t->v, ss + lo, (uint32_t)m, what == 1 ? st : nullptr, what == 2 ? st : nullptr,
            what == 3 ? reinterpret_cast<uint64_t*>(d_stage) : nullptr, what == 4 ? reinterpret_cast<uint32_t*>(d_stage) : nullptr);
      if (cudaMemcpyAsync(h_stage, src, m * width, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
          cudaStreamSynchronize(stream) != cudaSuccess)
        rc = fail(MEEPO_ECUDA, "tier export copy failed");
      else if (fwrite(h_stage, width, m, f) != m)
        rc = fail(MEEPO_EIO, "short write");
    }
  };
  section(0, 8);
  section(1, t->row_bytes);
  if (t->state_bytes) section(2, t->state_bytes);
  section(3, 8);
  section(4, 4);
  cudaFree(d_stage);
  cudaFreeHost(h_stage);
  if (rc == MEEPO_OK && fflush(f) != 0) rc = fail(MEEPO_EIO, "flush failed");
  return rc;
}

MEEPO_API meepo_status meepo_tier_import(meepo_table* t, const char* path) {
  if (!t || !path) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  FILE* f = fopen(path, "rb");
  if (!f) return fail(MEEPO_EIO, std::string("cannot open ") + path);
  std::unique_ptr<FILE, int (*)(FILE*)> closer(f, fclose);
  TierFileHeader h{};
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "MEEPOTB1", 8) != 0 || h.version != 1)
    return fail(MEEPO_EIO, "bad header");
  if (h.dim != t->cfg.dim || (int32_t)h.dtype != t->cfg.dtype || h.row_bytes != t->row_bytes ||
      h.state_bytes != t->state_bytes || (int32_t)h.opt != t->cfg.opt)
    return fail(MEEPO_EINVAL, "file does not match table configuration");
  const uint64_t n = h.n, R = t->row_bytes, S = t->state_bytes;
  const uint64_t tuple_file = 8 + R + S + 8 + 4;
  if (fseeko(f, 0, SEEK_END) != 0) return fail(MEEPO_EIO, "seek failed");
  const off_t fsz = ftello(f);
  if (fsz < 0 || (uint64_t)fsz != sizeof h + n * tuple_file || n > 0xFFFFFFFFull)
    return fail(MEEPO_EIO, "file size does not match the tuple count in the header");
  if (n == 0) return MEEPO_OK;
  if (!t->v.tier.slabs) return fail(MEEPO_EINVAL, "the table has no host tier (host_spill_bytes == 0)");
  const uint64_t off_keys = sizeof h, off_rows = off_keys + n * 8, off_state = off_rows + n * R,
                 off_scores = off_state + n * S, off_steps = off_scores + n * 8;
  const uint64_t chunk = 1u << 16;
  char *d_stage = nullptr, *h_stage = nullptr;
  MEEPO_CUDA_TRY(cudaMalloc(&d_stage, chunk * tuple_file + 1024));
  if (cudaHostAlloc(&h_stage, chunk * tuple_file + 1024, cudaHostAllocDefault) != cudaSuccess) {
    cudaFree(d_stage);
    return fail(MEEPO_ENOMEM, "cudaHostAlloc(import bounce)");
  }
  meepo_status rc = MEEPO_OK;
  for (uint64_t lo = 0; lo < n && rc == MEEPO_OK; lo += chunk) {
    const uint64_t m = std::min(chunk, n - lo);
    const size_t o_k = 0, o_r = o_k + chunk * 8, o_s = o_r + chunk * R, o_c = o_s + chunk * S, o_t = o_c + chunk * 8;
    auto rd = [&](uint64_t file_off, size_t width, size_t stage_off) {
      if (width == 0) return true;
      return fseeko(f, (off_t)(file_off + lo * width), SEEK_SET) == 0 && fread(h_stage + stage_off, width, m, f) == m;
    };
    if (!rd(off_keys, 8, o_k) || !rd(off_rows, R, o_r) || !rd(off_state, S, o_s) || !rd(off_scores, 8, o_c) ||
        !rd(off_steps, 4, o_t)) {
      rc = fail(MEEPO_EIO, "short read");
      break;
    }
    if (cudaMemcpy(d_stage, h_stage, o_t + chunk * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      rc = fail(MEEPO_ECUDA, "import copy failed");
      break;
    }
    rc = meepo_tier_import_buffers(t, reinterpret_cast<uint64_t*>(d_stage + o_k), d_stage + o_r, S ? d_stage + o_s : nullptr,
                                   reinterpret_cast<uint64_t*>(d_stage + o_c), reinterpret_cast<uint32_t*>(d_stage + o_t), m);
  }
  cudaFree(d_stage);
  cudaFreeHost(h_stage);
  return rc;
}

}  // extern "C"
