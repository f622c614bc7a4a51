// common.cuh — table layout in HBM, hashing, probing and row helpers shared by every kernel.
//
// Layout (DESIGN.md "Data layout"): open addressing over BUCKETS of 32 slots.
//   keys     u64[slots]           EMPTY = ~0; bucket b owns slots [32b, 32b+32)
//   digests  u8[slots]            one 32-byte sector per bucket; 0 = free slot, else an 8-bit tag
//                                 of the key's hash (1..255). A probe reads ONE sector of tags,
//                                 SIMD-compares them in registers and then touches only the
//                                 8-byte key(s) whose tag matched.
//   overflow u32 bitmap[buckets]  bit b set once any insertion has skipped past bucket b because
//                                 it was full; lookups follow the chain only while it is set, so
//                                 eviction can free slots without tombstones.
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

constexpr uint32_t kBucket = 32;
constexpr uint32_t kNil = 0xFFFFFFFFu;

enum Counter : int {
  C_SIZE = 0, C_INSERTS, C_HITS, C_MISSES, C_FULL, C_EVICTIONS, C_UPDATES, C_DROPPED, C_OVERFLOW, C_COUNT
};

struct TableView {
  uint64_t* keys;
  uint8_t* digests;
  uint32_t* overflow;
  uint4* rows;
  uint4* state;
  uint2* scores;
  uint32_t* steps;
  unsigned long long* counters;
  uint32_t num_buckets, slots;
  uint32_t cpr, scpr;  // 16-byte chunks per row / per state row
  uint32_t dim;
  int32_t dtype, opt;
  float lr, eps, beta1, beta2, init_accum, init_scale;
  uint64_t init_seed;
  uint32_t epoch;
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

// --- digest line -------------------------------------------------------------------------------
// 4 tag bytes -> 4 mask bits (bit i set if byte i of `w` equals the tag replicated in `pat`).
__device__ __forceinline__ uint32_t bytes_eq4(uint32_t w, uint32_t pat) {
  uint32_t m = __vcmpeq4(w, pat) & 0x80808080u;  // bit 7 of each matching byte
  return ((m >> 7) * 0x10204080u) >> 28;           // gather bits 0,8,16,24 into a nibble
}
struct DigestLine {
  uint4 lo, hi;
};
__device__ __forceinline__ DigestLine load_digests(const TableView& t, uint32_t b) {
  const uint4* p = reinterpret_cast<const uint4*>(t.digests + (size_t)b * kBucket);
  DigestLine d;
  d.lo = __ldg(p);
  d.hi = __ldg(p + 1);
  return d;
}
__device__ __forceinline__ uint32_t match_mask(const DigestLine& d, uint32_t tag) {
  uint32_t pat = tag * 0x01010101u;
  return bytes_eq4(d.lo.x, pat) | (bytes_eq4(d.lo.y, pat) << 4) | (bytes_eq4(d.lo.z, pat) << 8) |
         (bytes_eq4(d.lo.w, pat) << 12) | (bytes_eq4(d.hi.x, pat) << 16) | (bytes_eq4(d.hi.y, pat) << 20) |
         (bytes_eq4(d.hi.z, pat) << 24) | (bytes_eq4(d.hi.w, pat) << 28);
}
__device__ __forceinline__ bool overflowed(const TableView& t, uint32_t b) {
  return (__ldcg(t.overflow + (b >> 5)) >> (b & 31)) & 1u;
}

// Read-only probe: slot of `key` or kNil. Keys whose tag is published never move, so plain loads.
__device__ __forceinline__ uint32_t probe_find(const TableView& t, uint64_t key) {
  const uint64_t h = mix64(key);
  uint32_t b = bucket_of(h, t.num_buckets);
  const uint32_t tag = digest_of(h);
  for (uint32_t p = 0; p < t.num_buckets; ++p) {
    DigestLine d = load_digests(t, b);
    uint32_t m = match_mask(d, tag);
    while (m) {
      uint32_t s = b * kBucket + (__ffs(m) - 1);
      if (__ldg(t.keys + s) == key) return s;
      m &= m - 1;
    }
    if (!overflowed(t, b)) return kNil;
    b = (b + 1 == t.num_buckets) ? 0 : b + 1;
  }
  return kNil;
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
__device__ __forceinline__ Probe probe_find_or_insert(const TableView& t, uint64_t key) {
  Probe r{kNil, MEEPO_KEY_INVALID, false};
  if (!key_valid(key)) return r;
  uint32_t s = probe_find(t, key);
  if (s != kNil) {
    r.slot = s;
    r.status = MEEPO_KEY_FOUND;
    return r;
  }
  const uint64_t h = mix64(key);
  uint32_t b = bucket_of(h, t.num_buckets);
  for (uint32_t p = 0; p < t.num_buckets; ++p) {
    DigestLine d = load_digests(t, b);
    uint32_t free_m = match_mask(d, 0);
    while (free_m) {
      uint32_t slot = b * kBucket + (__ffs(free_m) - 1);
      unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + slot),
                                         (unsigned long long)MEEPO_KEY_EMPTY, (unsigned long long)key);
      if (old == MEEPO_KEY_EMPTY || old == key) {
        r.slot = slot;
        r.status = MEEPO_KEY_INSERTED;
        r.winner = (old == MEEPO_KEY_EMPTY);
        return r;
      }
      free_m &= free_m - 1;
    }
    if (!overflowed(t, b)) {
      const uint32_t bit = 1u << (b & 31);
      if (!(atomicOr(t.overflow + (b >> 5), bit) & bit)) atomicAdd(t.counters + C_OVERFLOW, 1ull);
    }
    b = (b + 1 == t.num_buckets) ? 0 : b + 1;
  }
  r.status = MEEPO_KEY_FULL;
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
  uint32_t a = __float_as_uint(t.opt == MEEPO_ADAGRAD ? t.init_accum : 0.0f);
  return make_uint4(a, a, a, a);
}
#endif  // __CUDACC__

}  // namespace meepo
