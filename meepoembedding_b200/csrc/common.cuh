// common.cuh — table layout in HBM, hashing, probing and row helpers shared by every kernel.
//
// Layout (DESIGN.md "Data layout"): open addressing over BUCKETS that are exactly one 128-byte line
// — the DRAM fetch granule measured on B200 (a random 8-byte read costs 128 B of HBM traffic, ncu
// dram__bytes_read; cudaLimitMaxL2FetchGranularity has no effect). One probe = one line:
//   bytes   0..13   14 one-byte tags of the bucket's 14 slots: 0 = free slot, else an 8-bit tag of
//                   the key's hash (1..255). The probe loads these 16 bytes, SIMD-compares all tags
//                   in registers and then touches only the 8-byte key(s) whose tag matched — which
//                   sit in the SAME line, i.e. an L2 hit, not a second HBM access.
//   bytes  14..15   per-bucket metadata: the largest DISPLACEMENT (in buckets) of any key whose home
//                   bucket this is (0xFFFF: unknown, walk everything). A probe for a key with home h
//                   visits buckets h .. h + disp(h) and stops: a miss costs 1 + disp(h) lines — at 90%
//                   load 90% of the homes have disp 0 — instead of a walk to the end of the run of
//                   full buckets, and eviction frees slots with no tombstones (disp is recomputed)
//   bytes 16..127   14 keys (u64, EMPTY = ~0), claimed with one 64-bit CAS
// Slot s = 14*bucket + index addresses the other arrays:
//   rows     16-byte chunks [slots * cpr]     value arena
//   state    16-byte chunks [slots * scpr]    optimizer-state arena (fp32)
//   scores   uint2[slots] {freq, last_epoch}  only with MEEPO_FLAG_TRACK_SCORES; 0 for free slots
//   steps    u32[slots]                       Adam only; 0 for free slots
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/meepo.h"

namespace meepo {

constexpr uint32_t kBucket = MEEPO_BUCKET_SLOTS;  // 14
constexpr uint32_t kNil = 0xFFFFFFFFu;

enum Counter : int {
  C_SIZE = 0, C_INSERTS, C_HITS, C_MISSES, C_FULL, C_EVICTIONS, C_UPDATES, C_DROPPED, C_OVERFLOW, C_PEER_KEYS, C_PEER_GRADS,
  C_PROMOTIONS, C_TIER_HITS, C_TIER_LIVE, C_COUNT
};

struct __align__(128) BucketLine {
  uint8_t tag[kBucket];
  uint16_t meta;  // largest displacement of a key whose home this bucket is
  uint64_t key[kBucket];
};
static_assert(sizeof(BucketLine) == 128, "a bucket is one 128-byte line");

// Host tier (meepo.h "Host tier"): a ring of `slabs` tuple slabs in mapped pinned host memory, managed entirely
// from the device. The device keeps which key each slab holds (ring_key) and an open-addressing index key -> slab
// (linear probing; erased cells become tombstones, the index is rebuilt from ring_key when they pile up). The
// slabs [stage_d0, stage_d0 + stage_m) mod slabs — the tuples of the latest meepo_evict — also sit in a staging
// buffer in HBM, from which they are served while (and after) they drain to the host ring by DMA.
struct TierView {
  uint64_t* idx_key;   // [idx_mask + 1] MEEPO_KEY_EMPTY = never used, MEEPO_KEY_RESERVED = tombstone
  uint32_t* idx_val;   // slab of idx_key[c]
  uint64_t* ring_key;  // [slabs] MEEPO_KEY_EMPTY = the slab holds nothing
  uint32_t idx_mask, slabs;  // slabs == 0: the table has no host tier
  uint4 *h_rows, *h_state, *h_meta;  // the ring: [slabs][cpr], [slabs][scpr], [slabs] {key lo, key hi, freq, epoch}
  uint32_t* h_steps;                 // [slabs]
  uint4 *s_rows, *s_state, *s_meta;  // staging copy of stage_m slabs, same layout
  uint32_t* s_steps;
  uint32_t stage_d0, stage_m;
};

struct TableView {
  BucketLine* buckets;
  uint4* rows;
  uint4* state;
  uint2* scores;
  uint32_t* steps;
  uint32_t* dirty;  // one bit per slot, only with MEEPO_FLAG_TRACK_DIRTY: inserted / updated since the last delta export
  unsigned long long* counters;
  uint32_t num_buckets, slots;
  uint32_t cpr, scpr;  // 16-byte chunks per row / per state row
  uint32_t dim;
  int32_t dtype, opt;
  float lr, eps, beta1, beta2, init_accum, init_scale;
  uint64_t init_seed;
  uint32_t epoch;
  TierView tier;
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}
__host__ __device__ __forceinline__ bool key_valid(uint64_t k) { return k < MEEPO_KEY_RESERVED; }

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t bucket_of(uint64_t h, uint32_t nb) {
  return (uint32_t)__umul64hi(h, (uint64_t)nb);
}
__device__ __forceinline__ uint32_t digest_of(uint64_t h) {
  uint32_t d = (uint32_t)h & 0xFFu;
  return d + (d == 0);
}
__device__ __forceinline__ uint32_t owner_of(uint64_t key, uint32_t g) {
  return (uint32_t)__umul64hi(mix64(key ^ MEEPO_OWNER_SALT), (uint64_t)g);
}

// --- memory helpers ----------------------------------------------------------------------------
// Streaming 16-byte accesses: rows are touched once per batch, keep them out of L1.
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// --- bucket line -------------------------------------------------------------------------------
// TB = anything with {BucketLine* buckets; uint32_t num_buckets; unsigned long long* counters;}: the
// local TableView, or a peer shard reached over NVLink (peer.cu).
template <typename TB>
__device__ __forceinline__ uint64_t* key_ptr(const TB& t, uint32_t slot) {
  return &t.buckets[slot / kBucket].key[slot % kBucket];
}
template <typename TB>
__device__ __forceinline__ uint8_t* tag_ptr(const TB& t, uint32_t slot) {
  return &t.buckets[slot / kBucket].tag[slot % kBucket];
}
// 4 tag bytes -> 4 mask bits (bit i set if byte i of `w` equals the tag replicated in `pat`).
__device__ __forceinline__ uint32_t bytes_eq4(uint32_t w, uint32_t pat) {
  uint32_t m = __vcmpeq4(w, pat) & 0x80808080u;  // bit 7 of each matching byte
  return ((m >> 7) * 0x10204080u) >> 28;           // gather bits 0,8,16,24 into a nibble
}
// Load policy of the probe. kCoherent (ld.global.cg, L2): find_or_insert, where other threads CAS
// keys / raise the displacement bound in the same line (tags and published keys never change in a kernel).
// kReadOnly (ld.global.nc, L1-cached): lookup / apply_gradients / evict probes — nothing in the
// bucket array is written while they run, and hot Zipf keys are served by L1.
enum : int { kCoherent = 0, kReadOnly = 1 };
template <int LD, typename TB>
__device__ __forceinline__ uint4 load_header(const TB& t, uint32_t b) {
  const uint4* p = reinterpret_cast<const uint4*>(&t.buckets[b]);
  return LD == kReadOnly ? __ldg(p) : __ldcg(p);
}
template <int LD, typename TB>
__device__ __forceinline__ uint64_t load_key(const TB& t, uint32_t b, uint32_t i) {
  const uint64_t* p = &t.buckets[b].key[i];
  return LD == kReadOnly ? __ldg(p) : __ldcg(p);
}
__device__ __forceinline__ uint32_t match_mask(const uint4& h, uint32_t tag) {
  const uint32_t pat = tag * 0x01010101u;
  return bytes_eq4(h.x, pat) | (bytes_eq4(h.y, pat) << 4) | (bytes_eq4(h.z, pat) << 8) |
         ((bytes_eq4(h.w, pat) & 3u) << 12);
}
constexpr uint32_t kDispUnknown = 0xFFFFu;
__device__ __forceinline__ uint32_t home_disp(const uint4& h) { return h.w >> 16; }

// Slot of `key` or kNil, starting at its home bucket b with the header already in registers: the buckets
// b .. b + disp(b).
template <int LD, typename TB>
__device__ __forceinline__ uint32_t match_in_bucket(const TB& t, uint64_t key, uint32_t tag, uint32_t b, const uint4& hdr) {
  uint32_t m = match_mask(hdr, tag);
  while (m) {
    const uint32_t i = __ffs(m) - 1;
    if (load_key<LD>(t, b, i) == key) return b * kBucket + i;
    m &= m - 1;
  }
  return kNil;
}
// The walk past the home bucket (10% of the keys at 90% load, 3% at 75%): kept out of line so that its registers
// do not count against the hot kernels. It is known to be `last` buckets long: the first two lines are fetched
// together (independent addresses) instead of one dependent HBM round trip after the other.
#ifdef MEEPO_AB_INLINE_WALK
#define MEEPO_WALK_INLINE __forceinline__
#else
#define MEEPO_WALK_INLINE __noinline__
#endif
template <int LD, typename TB>
__device__ MEEPO_WALK_INLINE uint32_t probe_walk(const TB& t, uint64_t key, uint32_t tag, uint32_t b, uint32_t last) {
  const uint32_t b1 = (b + 1 == t.num_buckets) ? 0 : b + 1;
  const uint32_t b2 = (b1 + 1 == t.num_buckets) ? 0 : b1 + 1;
  const uint4 h1 = load_header<LD>(t, b1);
  uint4 h2 = make_uint4(0, 0, 0, 0);
  if (last >= 2) h2 = load_header<LD>(t, b2);
  uint32_t s = match_in_bucket<LD>(t, key, tag, b1, h1);
  if (s != kNil || last < 2) return s;
  s = match_in_bucket<LD>(t, key, tag, b2, h2);
  if (s != kNil) return s;
  b = b2;
  for (uint32_t d = 3; d <= last; ++d) {
    b = (b + 1 == t.num_buckets) ? 0 : b + 1;
    s = match_in_bucket<LD>(t, key, tag, b, load_header<LD>(t, b));
    if (s != kNil) return s;
  }
  return kNil;
}
template <int LD, typename TB>
__device__ __forceinline__ uint32_t probe_from(const TB& t, uint64_t key, uint32_t tag, uint32_t b, uint4 hdr) {
  const uint32_t s = match_in_bucket<LD>(t, key, tag, b, hdr);
  if (s != kNil) return s;
  const uint32_t disp = home_disp(hdr);
  if (disp == 0) return kNil;
  return probe_walk<LD>(t, key, tag, b, disp == kDispUnknown ? t.num_buckets - 1 : min(disp, t.num_buckets - 1));
}
// Probe without insertion. One HBM line per bucket visited.
template <int LD = kCoherent, typename TB>
__device__ __forceinline__ uint32_t probe_find(const TB& t, uint64_t key) {
  const uint64_t h = mix64(key);
  const uint32_t b = bucket_of(h, t.num_buckets);
  return probe_from<LD>(t, key, digest_of(h), b, load_header<LD>(t, b));
}

// disp(home) = max(disp(home), d); returns true if disp(home) was 0 before (the caller counts such homes). The
// metadata is the high half of a 32-bit word whose low half holds tags 12 and 13, which no kernel that raises
// displacements writes (tags are published by their own kernel): with the low half copied from the current
// value, atomicMax orders the words by displacement alone.
template <typename TB>
__device__ __forceinline__ bool raise_disp(const TB& t, uint32_t home, uint32_t d) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&t.buckets[home]) + 3;
  d = min(d, kDispUnknown);
  const uint32_t cur = *reinterpret_cast<volatile uint32_t*>(w);
  if ((cur >> 16) >= d) return false;
  const uint32_t old = atomicMax(w, (cur & 0xFFFFu) | (d << 16));
  return (old >> 16) == 0;
}

struct Probe {
  uint32_t slot;    // kNil if the key has no row
  uint32_t status;  // MEEPO_KEY_*
  bool winner;      // this thread claimed the slot (must initialise row/state and list the slot)
};

// find, else claim a free slot with a 64-bit CAS. Tags of slots claimed in this launch are NOT
// written here (publish_kernel does it afterwards), so every thread sees the same free-slot set
// per bucket and walks it in the same order: a key can only ever land in one slot, and "found by
// tag" == "was present when the call started" (meepo.h "Status").
template <typename TB>
__device__ __forceinline__ Probe probe_find_or_insert(const TB& t, uint64_t key) {
  Probe r{kNil, MEEPO_KEY_INVALID, false};
  if (!key_valid(key)) return r;
  uint32_t s = probe_find<kCoherent>(t, key);
  if (s != kNil) {
    r.slot = s;
    r.status = MEEPO_KEY_FOUND;
    return r;
  }
  const uint64_t h = mix64(key);
  const uint32_t home = bucket_of(h, t.num_buckets);
  uint32_t b = home;
  for (uint32_t d = 0; d < t.num_buckets; ++d) {
    const uint4 hdr = load_header<kCoherent>(t, b);
    uint32_t free_m = match_mask(hdr, 0);
    while (free_m) {
      const uint32_t i = __ffs(free_m) - 1;
      unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&t.buckets[b].key[i]),
                                         (unsigned long long)MEEPO_KEY_EMPTY, (unsigned long long)key);
      if (old == MEEPO_KEY_EMPTY || old == key) {
        r.slot = b * kBucket + i;
        r.status = MEEPO_KEY_INSERTED;
        r.winner = (old == MEEPO_KEY_EMPTY);
        if (d && r.winner && raise_disp(t, home, d)) atomicAdd(t.counters + C_OVERFLOW, 1ull);
        return r;
      }
      free_m &= free_m - 1;
    }
    b = (b + 1 == t.num_buckets) ? 0 : b + 1;
  }
  r.status = MEEPO_KEY_FULL;
  return r;
}

__device__ __forceinline__ void mark_dirty(const TableView& t, uint32_t slot) {
  if (t.dirty) atomicOr(t.dirty + (slot >> 5), 1u << (slot & 31u));
}
__device__ __forceinline__ void mark_clean(const TableView& t, uint32_t slot) {
  if (t.dirty) atomicAnd(t.dirty + (slot >> 5), ~(1u << (slot & 31u)));
}

// --- host tier -----------------------------------------------------------------------------------
constexpr uint64_t kTierSalt = 0x7A3C91E5B4D2F680ull;  // the index hashes differently from the buckets
constexpr uint64_t kTomb = MEEPO_KEY_RESERVED;
__device__ __forceinline__ uint32_t tier_home(const TierView& tv, uint64_t key) {
  return (uint32_t)mix64(key ^ kTierSalt) & tv.idx_mask;
}
// Index cell holding `key`, or kNil. The index only changes in publish / evict / readmit kernels, never while
// probe kernels run.
__device__ __forceinline__ uint32_t tier_cell(const TierView& tv, uint64_t key) {
  if (!tv.slabs) return kNil;
  uint32_t c = tier_home(tv, key);
  for (uint32_t p = 0; p <= tv.idx_mask; p++, c = (c + 1) & tv.idx_mask) {
    const uint64_t k = tv.idx_key[c];
    if (k == key) return c;
    if (k == MEEPO_KEY_EMPTY) return kNil;
  }
  return kNil;
}
__device__ __forceinline__ uint32_t tier_slab(const TierView& tv, uint64_t key) {
  const uint32_t c = tier_cell(tv, key);
  return c == kNil ? kNil : tv.idx_val[c];
}
// Remove `key` from the tier (its slab becomes empty). Safe when several threads erase the same key: one wins.
__device__ __forceinline__ bool tier_erase(const TableView& t, uint64_t key) {
  const TierView& tv = t.tier;
  const uint32_t c = tier_cell(tv, key);
  if (c == kNil) return false;
  const uint32_t d = tv.idx_val[c];
  if (atomicCAS(reinterpret_cast<unsigned long long*>(tv.ring_key + d), (unsigned long long)key,
                (unsigned long long)MEEPO_KEY_EMPTY) != key)
    return false;
  tv.idx_key[c] = kTomb;
  atomicAdd(t.counters + C_TIER_LIVE, ~0ull);
  return true;
}
struct TierTuple {  // where the tuple of a slab can be read
  const uint4 *rows, *state, *meta;
  const uint32_t* steps;
};
__device__ __forceinline__ TierTuple tier_tuple(const TableView& t, uint32_t slab) {
  const TierView& tv = t.tier;
  uint32_t j = slab - tv.stage_d0;
  if (slab < tv.stage_d0) j += tv.slabs;
  TierTuple r;
  if (j < tv.stage_m) {
    r.rows = tv.s_rows + (size_t)j * t.cpr;
    r.state = tv.s_state + (size_t)j * t.scpr;
    r.meta = tv.s_meta + j;
    r.steps = tv.s_steps + j;
  } else {
    r.rows = tv.h_rows + (size_t)slab * t.cpr;
    r.state = tv.h_state + (size_t)slab * t.scpr;
    r.meta = tv.h_meta + slab;
    r.steps = tv.h_steps + slab;
  }
  return r;
}

// --- row init (meepo.h "Init") -----------------------------------------------------------------
__device__ __forceinline__ float init_from_bits(uint32_t u, float scale) {
  float a = __fmul_rn((float)(u >> 8), 1.1920928955078125e-07f);
  return __fmul_rn(__fadd_rn(a, -1.0f), scale);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-byte chunk `q` of the initial row of `key`.
__device__ __forceinline__ uint4 init_chunk(const TableView& t, uint64_t key, uint32_t q) {
  uint4 out;
  if (t.dtype == MEEPO_F32) {  // columns 4q..4q+3 -> pairs 2q, 2q+1
    uint64_t x0 = mix64(key + (t.init_seed ^ ((uint64_t)(2 * q + 1) * 0x9E3779B97F4A7C15ull)));
    uint64_t x1 = mix64(key + (t.init_seed ^ ((uint64_t)(2 * q + 2) * 0x9E3779B97F4A7C15ull)));
    out.x = __float_as_uint(init_from_bits((uint32_t)x0, t.init_scale));
    out.y = __float_as_uint(init_from_bits((uint32_t)(x0 >> 32), t.init_scale));
    out.z = __float_as_uint(init_from_bits((uint32_t)x1, t.init_scale));
    out.w = __float_as_uint(init_from_bits((uint32_t)(x1 >> 32), t.init_scale));
  } else {  // columns 8q..8q+7 -> pairs 4q..4q+3
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint64_t x = mix64(key + (t.init_seed ^ ((uint64_t)(4 * q + i + 1) * 0x9E3779B97F4A7C15ull)));
      w[i] = pack_bf16x2(init_from_bits((uint32_t)x, t.init_scale),
                         init_from_bits((uint32_t)(x >> 32), t.init_scale));
    }
    out = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return out;
}
__device__ __forceinline__ uint4 init_state_chunk(const TableView& t) {
  if (t.opt == MEEPO_ADAGRAD_ROWWISE) return make_uint4(__float_as_uint(t.init_accum), 0u, 0u, 0u);
  uint32_t a = __float_as_uint(t.opt == MEEPO_ADAGRAD ? t.init_accum : 0.0f);
  return make_uint4(a, a, a, a);
}
#endif  // __CUDACC__

}  // namespace meepo
