// update.cu — apply_gradients: probe -> sort by slot -> segment heads -> fused reduce + optimizer
// (semantics in include/meepo.h "Update").
//
// Duplicates are grouped by sorting (slot, batch index) pairs on the SLOT (<= 32 bits, and only
// ceil(log2(slots+1)) of them) instead of the 64-bit key: the slot identifies the key, the sort is
// stable so each segment lists its gradients by increasing batch index, and the per-key result
// does not depend on where the slot happens to be. One sub-warp group of lanes then owns one
// unique key: it sums the key's gradient rows in the normative order (leaves of 256, then leaf
// partials) with 16-byte loads, and applies the optimizer to row + state in the same registers.
// Segments longer than one leaf (hot Zipf keys, up to ~8% of the batch) are split into leaves that
// run in parallel (leaf_kernel) and are finished by long_finish_kernel.
#include <cub/device/device_radix_sort.cuh>

#include <map>

#include "compact.cuh"
#include "optimizer.cuh"
#include "table.h"

namespace meepo {

constexpr uint32_t kLeaf = MEEPO_REDUCE_LEAF;
// Segments with more duplicates than this leave the per-segment kernel for the leaf kernels (the sum
// has the same normative order either way): a 256-term chain in one group of lanes is a long
// latency-bound tail, whereas leaves run one per group with 8 row loads in flight.
constexpr uint32_t kLongSeg = 32;

struct LongSeg {
  uint32_t seg, base, nleaf, pad;
};

// ---------------------------------------------------------------------------------------------
// A1: slot of every key (sort key) + iota (sort value). Keys that the preceding find_or_insert /
// lookup already resolved (same batch: keys[i] == cache.keys[i]) reuse the cached slot; everything
// else is probed (one 128-byte bucket line per key).
// Pooled backward (bag_offsets != null): the sort value is the BAG of occurrence i instead of i — the reduce
// kernels read "gradient row of the sort value", which is then the bag's gradient row; the sort is stable, so
// the occurrences of a key stay in batch order either way.
__global__ void __launch_bounds__(256) grad_slots_kernel(TableView t, const uint64_t* __restrict__ keys,
                                                         uint32_t n, uint32_t* __restrict__ sort_key,
                                                         uint32_t* __restrict__ sort_val, SlotCache sc,
                                                         uint32_t cache_n, const uint32_t* __restrict__ bag_offsets,
                                                         uint32_t n_bags) {
  uint32_t dropped = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = __ldg(keys + i);
    uint32_t s;
    if (i < cache_n && __ldg(sc.keys + i) == key)
      s = __ldg(sc.slots + i);
    else
      s = key_valid(key) ? probe_find<kReadOnly>(t, key) : kNil;
    if (s == kNil) {
      s = t.slots;  // sorts after every real slot
      dropped++;
    }
    sort_key[i] = s;
    uint32_t v = i;
    if (bag_offsets) {  // last bag b with offsets[b] <= i (empty bags share their offset with the next one)
      uint32_t lo = 0, hi = n_bags;
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(bag_offsets + mid) <= i) lo = mid; else hi = mid;
      }
      v = lo;
    }
    sort_val[i] = v;
  }
  dropped = __reduce_add_sync(0xFFFFFFFFu, dropped);
  if ((threadIdx.x & 31) == 0 && dropped) atomicAdd(t.counters + C_DROPPED, (unsigned long long)dropped);
}

// ---------------------------------------------------------------------------------------------
// A3: segment heads of the sorted slot array -> seg_start[0..U], seg_desc[0..U], U: one ordered
// compaction pass (compact.cuh). The CTA of the last tile closes the lists, publishes U and resets the
// long-segment counters.
__global__ void __launch_bounds__(kCompactThreads) segments_kernel(const uint32_t* __restrict__ sk,
                                                                   const uint32_t* __restrict__ sv, uint32_t n,
                                                                   uint32_t* __restrict__ seg_start,
                                                                   uint4* __restrict__ seg_desc, uint32_t miss_key,
                                                                   SegScratch* sc, unsigned long long* counters,
                                                                   int count_updates, const uint32_t* __restrict__ n_dev,
                                                                   int chunk_shift, CompactState cs) {
  if (n_dev) n = min(n, __ldg(n_dev));
  CompactTile ct = compact_begin(cs, n);
  if ((uint64_t)ct.tile * kCompactTile >= n && !(n == 0 && ct.tile == 0)) return;  // tiles past a device-side n
  unsigned flags = 0;
  uint32_t key[kCompactItems];
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) {
    const uint64_t i = ct.pos(k);
    key[k] = 0;
    if (i < n) {
      key[k] = sk[i];
      if (i == 0 || sk[i - 1] != key[k]) flags |= 1u << k;
    }
  }
  compact_rank(ct, flags, cs);
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) {
    if ((flags >> k) & 1u) {
      const uint32_t i = (uint32_t)ct.pos(k), u = (uint32_t)ct.rank(k);
      seg_start[u] = i;
      seg_desc[u] = make_uint4(i, key[k], sv[i], 0);  // {first sorted position, sort key, first batch index}
      if (chunk_shift >= 0 && (i == 0 || (sk[i - 1] >> chunk_shift) != (key[k] >> chunk_shift)))
        sc->chunk_first[min(key[k] >> chunk_shift, 7u)] = u;
    }
  }
  if (ct.last && threadIdx.x == 0) {
    const uint32_t U = (uint32_t)ct.base + ct.tile_total;
    seg_start[U] = n;
    seg_desc[U] = make_uint4(n, kNil, 0, 0);
    sc->num_segments = U;
    sc->num_long = 0;
    sc->num_leaves = 0;
    const uint32_t applied = U - (n && sk[n - 1] >= miss_key ? 1u : 0u);
    if (applied && count_updates) atomicAdd(counters + C_UPDATES, (unsigned long long)applied);
  }
}

template <bool BF16, int UNROLL = 4>
__device__ __forceinline__ void reduce_tail(const uint4* __restrict__ grads, const uint32_t* __restrict__ sidx,
                                            uint32_t j0, uint32_t s1, uint32_t cpr, uint32_t q,
                                            float (&acc)[Chunk<BF16>::E]);

// acc = ((g[s0] + g[s0+1]) + ...) over sorted positions [s0, s1), chunk q of every row.
template <bool BF16, int UNROLL = 4>
__device__ __forceinline__ void reduce_positions(const uint4* __restrict__ grads,
                                                 const uint32_t* __restrict__ sidx, uint32_t s0, uint32_t s1,
                                                 uint32_t cpr, uint32_t q, float (&acc)[Chunk<BF16>::E]) {
  widen<BF16>(ld_nc(grads + (size_t)sidx[s0] * cpr + q), acc);
  reduce_tail<BF16, UNROLL>(grads, sidx, s0 + 1, s1, cpr, q, acc);
}

// acc = ((acc + g[j0]) + g[j0+1]) + ... over sorted positions [j0, s1)
template <bool BF16, int UNROLL>
__device__ __forceinline__ void reduce_tail(const uint4* __restrict__ grads, const uint32_t* __restrict__ sidx,
                                            uint32_t j0, uint32_t s1, uint32_t cpr, uint32_t q,
                                            float (&acc)[Chunk<BF16>::E]) {
  constexpr int E = Chunk<BF16>::E;
  uint32_t j = j0;
  for (; j + UNROLL <= s1; j += UNROLL) {
    uint4 raw[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) raw[u] = ld_nc(grads + (size_t)sidx[j + u] * cpr + q);
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      float g[E];
      widen<BF16>(raw[u], g);
#pragma unroll
      for (int e = 0; e < E; e++) acc[e] = __fadd_rn(acc[e], g[e]);
    }
  }
  for (; j < s1; j++) {
    float g[E];
    widen<BF16>(ld_nc(grads + (size_t)sidx[j] * cpr + q), g);
#pragma unroll
    for (int e = 0; e < E; e++) acc[e] = __fadd_rn(acc[e], g[e]);
  }
}


struct ApplyArgs {
  const uint4* grads;
  const uint32_t* sorted_slot;
  const uint32_t* sorted_idx;
  const uint32_t* seg_start;
  const uint4* seg_desc;  // [U+1] {first sorted position, sort key (slot), first batch index, -}
  SegScratch* sc;
  LongSeg* long_seg;
  uint2* leaf_desc;
  float* partial;  // [leaves][dim]
  uint32_t group_lanes;  // power of two <= min(cpr, 32)
  uint32_t key_lo, key_hi;  // only segments whose sort key lies in [key_lo, key_hi) are processed: keys >= the table's
                            // slot count / the batch size are the "absent key" segment, and a sharded sender reduces
                            // its unique keys chunk by chunk (the chunk index sits above the unique id)
  uint32_t key_mask;        // kStoreOnly: sort key & key_mask = index of the destination row
  int chunk;                // >= 0: visit only the segments [first segment of this chunk, first of the next)
  uint4* reduce_out;         // kStoreOnly: [unique][cpr] ...
  uint4* const* reduce_rows;  // ... or one destination row pointer per sort key (may point into a peer's window)
  int bulk;                   // kStoreOnly, pipelined kernel: rows leave through shared memory as ONE bulk async store each
};

// Segment index range of this pass: everything, or one chunk of a chunked sort key.
__device__ __forceinline__ void segment_range(const ApplyArgs& a, uint32_t U, uint32_t& lo, uint32_t& hi) {
  lo = 0;
  hi = U;
  if (a.chunk < 0) return;
  lo = U;
  for (int c = 7; c >= a.chunk; c--) {  // first non-empty chunk at or after a.chunk
    const uint32_t f = a.sc->chunk_first[c];
    if (f != kNil) lo = f;
  }
  for (int c = 7; c > a.chunk; c--) {
    const uint32_t f = a.sc->chunk_first[c];
    if (f != kNil) hi = f;
  }
  hi = max(hi, lo);
}

// kStoreOnly: where the summed row of sort key `key` goes.
template <int OPT>
__device__ __forceinline__ uint4* reduce_row_of(const ApplyArgs& a, uint32_t key, uint32_t cpr) {
  if constexpr (OPT != kStoreOnly) return nullptr;
  key &= a.key_mask;
  if (a.reduce_rows)
    return reinterpret_cast<uint4*>(__ldg(reinterpret_cast<const unsigned long long*>(a.reduce_rows) + key));
  return a.reduce_out + (size_t)key * cpr;
}

// A4: one group of lanes per segment (unique key)
template <bool BF16, int OPT>
__global__ void __launch_bounds__(256) apply_kernel(TableView t, ApplyArgs a) {
  constexpr int E = Chunk<BF16>::E;
  const uint32_t GL = a.group_lanes;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gl = lane & (GL - 1);
  const unsigned gmask = (GL == 32 ? 0xFFFFFFFFu : ((1u << GL) - 1u) << (lane & ~(GL - 1)));
  const uint32_t groups_per_block = blockDim.x / GL;
  const uint32_t group = blockIdx.x * groups_per_block + threadIdx.x / GL;
  const uint32_t ngroups = gridDim.x * groups_per_block;
  const uint32_t U = a.sc->num_segments;
  uint32_t u_lo, u_hi;
  segment_range(a, U, u_lo, u_hi);
  for (uint32_t u = u_lo + group; u < u_hi; u += ngroups) {
    const uint32_t s0 = a.seg_start[u], s1 = a.seg_start[u + 1];
    const uint32_t slot = a.sorted_slot[s0];
    if (slot < a.key_lo || slot >= a.key_hi) continue;  // absent / invalid keys, or another chunk
    const uint32_t cnt = s1 - s0;
    if (cnt > kLongSeg) {  // hand over to the leaf kernels
      const uint32_t nleaf = (cnt + kLeaf - 1) / kLeaf;
      uint32_t base = 0;
      if (gl == 0) {
        const uint32_t li = atomicAdd(&a.sc->num_long, 1u);
        base = atomicAdd(&a.sc->num_leaves, nleaf);
        a.long_seg[li] = LongSeg{u, base, nleaf, 0};
      }
      base = __shfl_sync(gmask, base, lane & ~(GL - 1));
      for (uint32_t c = gl; c < nleaf; c += GL) a.leaf_desc[base + c] = make_uint2(u, c);
      continue;
    }
    float alpha = adam_alpha<OPT>(t, slot, gmask, gl == 0);
    if (OPT != kStoreOnly && gl == 0) mark_dirty(t, slot);
    uint32_t accum = 0;  // row-wise Adagrad: the accumulator as it was before this step
    if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) {
      // meepo.h: s_r over this lane's chunk classes r = gl + m * GL, the halving tree inside the lane for the
      // distances >= GL, then across the lanes of the group. A first pass over the gradients (the update
      // below reduces them again: this is the fallback kernel for rows that are not <= 32 chunks, a power of two)
      const uint32_t M = 32u / GL;
      float cls[32];
      for (uint32_t m = 0; m < M; m++) cls[m] = 0.0f;
      uint32_t j = 0;
      for (uint32_t q = gl; q < t.cpr; q += GL, j++) {
        float acc[E];
        reduce_positions<BF16>(a.grads, a.sorted_idx, s0, s1, t.cpr, q, acc);
        cls[j % M] = __fadd_rn(cls[j % M], chunk_sumsq<BF16>(acc));
      }
      for (uint32_t h = M >> 1; h >= 1; h >>= 1)
        for (uint32_t m = 0; m < h; m++) cls[m] = __fadd_rn(cls[m], cls[m + h]);
      alpha = __fdiv_rn(group_tree_sum(cls[0], GL, gmask), (float)t.dim);
      accum = __shfl_sync(gmask, t.state[slot].x, lane & ~(GL - 1));  // read before any lane of the group stores it
    }
    for (uint32_t q = gl; q < t.cpr; q += GL) {
      float acc[E];
      reduce_positions<BF16>(a.grads, a.sorted_idx, s0, s1, t.cpr, q, acc);
      OptIn<BF16, OPT> in;
      opt_issue<BF16, OPT>(t, slot, q, in);
      if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) in.st[0].x = accum;
      opt_finish<BF16, OPT>(t, slot, q, in, acc, alpha, reduce_row_of<OPT>(a, slot, t.cpr));
    }
  }
}

// Bulk asynchronous store shared -> global (cp.async.bulk, the TMA unit; SASS UBLKCP): one 16-byte-aligned run of
// `bytes` (a multiple of 16) per instruction. On a peer-mapped destination the row crosses NVLink as one request
// of the copy engine's packet size instead of four 128-byte write requests of the load/store unit.
__device__ __forceinline__ void bulk_store_row(void* dst, const void* smem_src, uint32_t bytes) {
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// A4': the same work for rows of <= 512 B (one 16-byte chunk per lane, group_lanes == cpr), software
// pipelined: each group handles two segments per trip, puts the gradient / row / state loads of
// both in flight before it consumes the first, and fetches the next trip's descriptors underneath.
template <bool BF16, int OPT>
__global__ void __launch_bounds__(256) apply_pipelined_kernel(TableView t, ApplyArgs a) {
  constexpr int E = Chunk<BF16>::E;
  const uint32_t GL = a.group_lanes;  // == t.cpr
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t q = lane & (GL - 1);
  const unsigned gmask = (GL == 32 ? 0xFFFFFFFFu : ((1u << GL) - 1u) << (lane & ~(GL - 1)));
  const uint32_t groups_per_block = blockDim.x / GL;
  const uint32_t group = blockIdx.x * groups_per_block + threadIdx.x / GL;
  const uint32_t stride = gridDim.x * groups_per_block * 2;
  const uint32_t U = a.sc->num_segments;
  const uint32_t cpr = t.cpr;
  const uint4 nil = make_uint4(0, kNil, 0, 0);
  auto desc = [&](uint32_t i) { return i <= U ? __ldg(a.seg_desc + i) : nil; };

  // kStoreOnly with a.bulk: two staging rows per group (segment A / B of a trip)
  __shared__ __align__(128) uint4 bulk_stage[OPT == kStoreOnly ? 2 * 256 : 1];
  const bool bulk = OPT == kStoreOnly && a.bulk;
  uint32_t u_lo, u_hi;
  segment_range(a, U, u_lo, u_hi);
  uint32_t u = u_lo + group * 2;
  uint4 da = desc(u), db = desc(u + 1);
  uint32_t ec = desc(u + 2).x;
  while (u < u_hi) {
    if (bulk) {  // the previous trip's bulk stores have read their staging rows
      if (q == 0) bulk_wait_read();
      __syncwarp(gmask);
    }
    const uint32_t un = u + stride;
    const uint4 na = desc(un), nb = desc(un + 1);
    const uint32_t nc = desc(un + 2).x;
    const uint32_t cntA = db.x - da.x, cntB = ec - db.x;
    const bool liveA = da.y >= a.key_lo && da.y < a.key_hi;
    const bool liveB = (u + 1 < U) && db.y >= a.key_lo && db.y < a.key_hi;
    const bool okA = liveA && cntA <= kLongSeg, okB = liveB && cntB <= kLongSeg;
    uint4 gA = make_uint4(0, 0, 0, 0), gB = gA;
    OptIn<BF16, OPT> inA, inB;
    if (okA) {
      gA = ld_nc(a.grads + (size_t)da.z * cpr + q);
      opt_issue<BF16, OPT>(t, da.y, q, inA);
    }
    if (okB) {
      gB = ld_nc(a.grads + (size_t)db.z * cpr + q);
      opt_issue<BF16, OPT>(t, db.y, q, inB);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {  // segments longer than a leaf go to the leaf kernels
      const bool live = h ? liveB : liveA;
      const uint32_t cnt = h ? cntB : cntA;
      if (live && cnt > kLongSeg) {
        const uint32_t nleaf = (cnt + kLeaf - 1) / kLeaf;
        uint32_t base = 0;
        if (q == 0) {
          const uint32_t li = atomicAdd(&a.sc->num_long, 1u);
          base = atomicAdd(&a.sc->num_leaves, nleaf);
          a.long_seg[li] = LongSeg{u + h, base, nleaf, 0};
        }
        base = __shfl_sync(gmask, base, lane & ~(GL - 1));
        for (uint32_t c = q; c < nleaf; c += GL) a.leaf_desc[base + c] = make_uint2(u + h, c);
      }
    }
    if (OPT != kStoreOnly && q == 0) {
      if (okA) mark_dirty(t, da.y);
      if (okB) mark_dirty(t, db.y);
    }
    if (okA) {
      float alpha = adam_alpha<OPT>(t, da.y, gmask, q == 0);
      float acc[E];
      widen<BF16>(gA, acc);
      if (cntA > 1) reduce_tail<BF16>(a.grads, a.sorted_idx, da.x + 1, da.x + cntA, cpr, q, acc);
      if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) {  // lane q holds chunk q: the tree is a butterfly over the group
        alpha = __fdiv_rn(group_tree_sum(chunk_sumsq<BF16>(acc), GL, gmask), (float)t.dim);
        inA.st[0].x = __shfl_sync(gmask, inA.st[0].x, lane & ~(GL - 1));  // the value lane 0 (the writer) loaded
      }
      if (bulk) {
        bulk_stage[threadIdx.x] = narrow<BF16>(acc);
        fence_proxy_async_smem();
        __syncwarp(gmask);
        if (q == 0) bulk_store_row(reduce_row_of<OPT>(a, da.y, cpr), &bulk_stage[threadIdx.x], cpr * 16u);
      } else {
        opt_finish<BF16, OPT>(t, da.y, q, inA, acc, alpha, reduce_row_of<OPT>(a, da.y, cpr));
      }
    }
    if (okB) {
      float alpha = adam_alpha<OPT>(t, db.y, gmask, q == 0);
      float acc[E];
      widen<BF16>(gB, acc);
      if (cntB > 1) reduce_tail<BF16>(a.grads, a.sorted_idx, db.x + 1, db.x + cntB, cpr, q, acc);
      if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) {
        alpha = __fdiv_rn(group_tree_sum(chunk_sumsq<BF16>(acc), GL, gmask), (float)t.dim);
        inB.st[0].x = __shfl_sync(gmask, inB.st[0].x, lane & ~(GL - 1));
      }
      if (bulk) {
        bulk_stage[256 + threadIdx.x] = narrow<BF16>(acc);
        fence_proxy_async_smem();
        __syncwarp(gmask);
        if (q == 0) bulk_store_row(reduce_row_of<OPT>(a, db.y, cpr), &bulk_stage[256 + threadIdx.x], cpr * 16u);
      } else {
        opt_finish<BF16, OPT>(t, db.y, q, inB, acc, alpha, reduce_row_of<OPT>(a, db.y, cpr));
      }
    }
    da = na;
    db = nb;
    ec = nc;
    u = un;
  }
  if (bulk && q == 0) bulk_wait_all();  // the rows have landed before the kernel — and the barrier after it — ends
}

// A5: leaves of long segments -> partial[leaf][dim]. A leaf is a chain of up to 256 adds in batch
// order per column; what can be hidden is the latency under it. One thread owns one piece of the row
// of V 4-byte words, so a batch of 16 independent row loads costs 16*V registers and a warp still
// reads 128*V contiguous bytes per row; the indices of the next batch are fetched (broadcast loads,
// L1 hits) while the current batch is in flight. Full batches run without predicates — an earlier
// per-element-predicated version issued 17 instructions per load (ncu). V = 2 for rows above 256 B
// (fewer instructions per byte: 0.46 -> 0.34 ms on cfg3/Zipf), V = 1 below (more threads per leaf:
// 0.125 -> 0.065 ms on cfg4, where a 1M batch has too few leaves to fill the GPU).
template <int V>
struct Piece;
template <>
struct Piece<1> {
  using T = uint32_t;
  static __device__ __forceinline__ T ld(const T* p) {
    T r;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ uint32_t word(const T& v, int) { return v; }
};
template <>
struct Piece<2> {
  using T = uint2;
  static __device__ __forceinline__ T ld(const T* p) {
    T r;
    asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ uint32_t word(const T& v, int i) { return i ? v.y : v.x; }
};

template <bool BF16, int V>
struct LeafAcc {
  static constexpr int N = (BF16 ? 2 : 1) * V;  // columns per piece
  float a[N];
  template <bool FIRST>
  __device__ __forceinline__ void take(const typename Piece<V>::T& x) {
#pragma unroll
    for (int i = 0; i < V; i++) {
      const uint32_t w = Piece<V>::word(x, i);
      if constexpr (BF16) {
        const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xFFFF0000u);
        a[2 * i] = FIRST ? lo : __fadd_rn(a[2 * i], lo);
        a[2 * i + 1] = FIRST ? hi : __fadd_rn(a[2 * i + 1], hi);
      } else {
        a[i] = FIRST ? __uint_as_float(w) : __fadd_rn(a[i], __uint_as_float(w));
      }
    }
  }
};

template <bool BF16, int V>
__global__ void __launch_bounds__(256) leaf_kernel(TableView t, ApplyArgs a, uint32_t tpl /* threads per leaf */) {
  constexpr uint32_t B = 16;
  using P = Piece<V>;
  constexpr int N = LeafAcc<BF16, V>::N;
  const uint32_t W = t.cpr * 4 / V;  // pieces per row
  const uint32_t sub = threadIdx.x / tpl, wi = threadIdx.x % tpl;
  const uint32_t lpc = blockDim.x / tpl;  // leaves per CTA
  const uint32_t nleaves = a.sc->num_leaves;
  const typename P::T* g = reinterpret_cast<const typename P::T*>(a.grads);
  for (uint32_t l = blockIdx.x * lpc + sub; l < nleaves; l += gridDim.x * lpc) {
    const uint2 d = a.leaf_desc[l];
    const uint32_t s0 = a.seg_start[d.x] + d.y * kLeaf;
    const uint32_t cnt = min(a.seg_start[d.x + 1], s0 + kLeaf) - s0;
    const uint32_t* sidx = a.sorted_idx + s0;
    for (uint32_t wb = 0; wb < W; wb += tpl) {
      const uint32_t w = wb + wi;
      if (w >= W) continue;
      const typename P::T* col = g + w;
      LeafAcc<BF16, V> acc;
      uint32_t ia[B];  // indices of the batch about to be loaded, always fetched one batch ahead
#pragma unroll
      for (uint32_t u = 0; u < B; u++) ia[u] = u < cnt ? __ldg(sidx + u) : 0u;
      bool first = true;
      for (uint32_t j = 0; j < cnt; j += B) {
        typename P::T x[B];
        const bool full = j + B <= cnt;
        if (full) {
#pragma unroll
          for (uint32_t u = 0; u < B; u++) x[u] = P::ld(col + (size_t)ia[u] * W);
        } else {
#pragma unroll
          for (uint32_t u = 0; u < B; u++)
            if (j + u < cnt) x[u] = P::ld(col + (size_t)ia[u] * W);
        }
        if (j + B < cnt) {
#pragma unroll
          for (uint32_t u = 0; u < B; u++) ia[u] = j + B + u < cnt ? __ldg(sidx + j + B + u) : 0u;
        }
        if (first)
          acc.template take<true>(x[0]);
        else
          acc.template take<false>(x[0]);
        first = false;
        if (full) {
#pragma unroll
          for (uint32_t u = 1; u < B; u++) acc.template take<false>(x[u]);
        } else {
#pragma unroll
          for (uint32_t u = 1; u < B; u++)
            if (j + u < cnt) acc.template take<false>(x[u]);
        }
      }
      float* out = a.partial + (size_t)l * t.dim + (size_t)w * N;
#pragma unroll
      for (int i = 0; i < N; i++) out[i] = acc.a[i];
    }
  }
}

// A6: one CTA per long segment: G = ((leaf_0 + leaf_1) + leaf_2) + ..., then the optimizer.
// The hottest Zipf key of a 4M batch has ~1300 leaves, and the order of the sum is normative, so the
// chain cannot be split — but its loads can: one thread per COLUMN keeps 32 independent, coalesced
// 4-byte loads in flight and adds them in leaf order; the sums are transposed through shared memory
// into the 16-byte-chunk layout of the optimizer step.
template <bool BF16, int OPT>
__global__ void __launch_bounds__(256) long_finish_kernel(TableView t, ApplyArgs a) {
  constexpr int E = Chunk<BF16>::E;
  constexpr int D = 32;
  extern __shared__ float sum_s[];  // [dim]
  __shared__ float alpha_s;
  __shared__ uint32_t accum_s;  // row-wise Adagrad: the accumulator before this step
  const uint32_t nlong = a.sc->num_long;
  for (uint32_t li = blockIdx.x; li < nlong; li += gridDim.x) {
    const LongSeg ls = a.long_seg[li];
    const uint32_t slot = a.sorted_slot[a.seg_start[ls.seg]];
    for (uint32_t col = threadIdx.x; col < t.dim; col += blockDim.x) {
      const float* p = a.partial + (size_t)ls.base * t.dim + col;
      float acc = __ldcg(p);
      uint32_t c = 1;
      for (; c + D <= ls.nleaf; c += D) {  // full batches: all D loads issued before the first add
        float x[D];
#pragma unroll
        for (int k = 0; k < D; k++) x[k] = __ldcg(p + (size_t)(c + k) * t.dim);
#pragma unroll
        for (int k = 0; k < D; k++) acc = __fadd_rn(acc, x[k]);
      }
      for (; c < ls.nleaf; c += 8) {
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = c + k < ls.nleaf ? __ldcg(p + (size_t)(c + k) * t.dim) : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (c + k < ls.nleaf) acc = __fadd_rn(acc, x[k]);
      }
      sum_s[col] = acc;
    }
    if (threadIdx.x == 0) {
      alpha_s = 0.0f;
      if (OPT != kStoreOnly) mark_dirty(t, slot);
      if constexpr (OPT == MEEPO_ADAM) {  // per-row step count -> scalar step size (meepo.h "Update")
        const uint32_t tt = t.steps[slot] + 1;
        t.steps[slot] = tt;
        const double bc1 = 1.0 - pow((double)t.beta1, (double)tt);
        const double bc2 = 1.0 - pow((double)t.beta2, (double)tt);
        alpha_s = (float)((double)t.lr * sqrt(bc2) / bc1);
      }
    }
    __syncthreads();
    if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) {
      if (threadIdx.x == 0) {  // meepo.h: chunk sums, classes q mod 32 in increasing q, halving tree (a few hundred flops)
        float cls[32];
        for (int r = 0; r < 32; r++) cls[r] = 0.0f;
        for (uint32_t q = 0; q < t.cpr; q++) {
          float c = __fmul_rn(sum_s[q * E], sum_s[q * E]);
          for (int e = 1; e < E; e++) c = __fadd_rn(c, __fmul_rn(sum_s[q * E + e], sum_s[q * E + e]));
          cls[q & 31u] = __fadd_rn(cls[q & 31u], c);
        }
        for (int d = 16; d >= 1; d >>= 1)
          for (int r = 0; r < d; r++) cls[r] = __fadd_rn(cls[r], cls[r + d]);
        alpha_s = __fdiv_rn(cls[0], (float)t.dim);
        accum_s = t.state[slot].x;
      }
      __syncthreads();
    }
    for (uint32_t q = threadIdx.x; q < t.cpr; q += blockDim.x) {
      float acc[E];
#pragma unroll
      for (int e = 0; e < E; e++) acc[e] = sum_s[q * E + e];
      OptIn<BF16, OPT> in;
      opt_issue<BF16, OPT>(t, slot, q, in);
      if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) in.st[0].x = accum_s;
      opt_finish<BF16, OPT>(t, slot, q, in, acc, alpha_s, reduce_row_of<OPT>(a, slot, t.cpr));
    }
    __syncthreads();
  }
}

template <bool BF16>
static void pick_apply(int opt, bool pipelined, const void*& apply, const void*& finish) {
  if (pipelined) {
    switch (opt) {
      case MEEPO_SGD: apply = (const void*)apply_pipelined_kernel<BF16, MEEPO_SGD>; break;
      case MEEPO_ADAGRAD: apply = (const void*)apply_pipelined_kernel<BF16, MEEPO_ADAGRAD>; break;
      case MEEPO_ADAM: apply = (const void*)apply_pipelined_kernel<BF16, MEEPO_ADAM>; break;
      case MEEPO_ADAGRAD_ROWWISE: apply = (const void*)apply_pipelined_kernel<BF16, MEEPO_ADAGRAD_ROWWISE>; break;
      default: apply = (const void*)apply_pipelined_kernel<BF16, kStoreOnly>;
    }
  }
  const void* plain = nullptr;
  switch (opt) {
    case MEEPO_SGD:
      plain = (const void*)apply_kernel<BF16, MEEPO_SGD>;
      finish = (const void*)long_finish_kernel<BF16, MEEPO_SGD>;
      break;
    case MEEPO_ADAGRAD:
      plain = (const void*)apply_kernel<BF16, MEEPO_ADAGRAD>;
      finish = (const void*)long_finish_kernel<BF16, MEEPO_ADAGRAD>;
      break;
    case MEEPO_ADAM:
      plain = (const void*)apply_kernel<BF16, MEEPO_ADAM>;
      finish = (const void*)long_finish_kernel<BF16, MEEPO_ADAM>;
      break;
    case MEEPO_ADAGRAD_ROWWISE:
      plain = (const void*)apply_kernel<BF16, MEEPO_ADAGRAD_ROWWISE>;
      finish = (const void*)long_finish_kernel<BF16, MEEPO_ADAGRAD_ROWWISE>;
      break;
    default:
      plain = (const void*)apply_kernel<BF16, kStoreOnly>;
      finish = (const void*)long_finish_kernel<BF16, kStoreOnly>;
  }
  if (!pipelined) apply = plain;
}

// Scratch of one sort + segment + reduce pipeline over n (sort key, batch index) pairs.
size_t SegWork::bytes(uint64_t n, uint32_t dim, int end_bit_) {
  size_t cub = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, end_bit_);
  if (radix_sort_supported(n, end_bit_)) cub = std::max(cub, radix_sort_temp_bytes(n, end_bit_));
  const size_t max_long_ = n / (kLongSeg + 1) + 1, max_leaves_ = n / kLongSeg + 2;
  return 4 * Workspace::pad(n * 4) + Workspace::pad(cub) + Workspace::pad(compact_state_bytes(n)) +
         Workspace::pad((n + 2) * 4) + Workspace::pad((n + 2) * 16) + Workspace::pad(max_long_ * sizeof(LongSeg)) +
         Workspace::pad(max_leaves_ * 8) + Workspace::pad(max_leaves_ * dim * 4) + Workspace::pad(sizeof(SegScratch)) +
         4096;
}

void SegWork::take(Workspace& ws, uint64_t n_, uint32_t dim, int end_bit_) {
  n = (uint32_t)n_;
  end_bit = end_bit_;
  cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, end_bit);
  if (radix_sort_supported(n_, end_bit)) cub_bytes = std::max(cub_bytes, radix_sort_temp_bytes(n_, end_bit));
  ntiles = compact_tiles(n_);
  max_long = n_ / (kLongSeg + 1) + 1;
  max_leaves = n_ / kLongSeg + 2;
  sk_in = ws.take<uint32_t>(n_);
  sk_out = ws.take<uint32_t>(n_);
  sv_in = ws.take<uint32_t>(n_);
  sv_out = ws.take<uint32_t>(n_);
  cub_tmp = ws.take<char>(cub_bytes);
  cstate_bytes = compact_state_bytes(n_);
  cstate = ws.take<char>(cstate_bytes);
  seg_start = ws.take<uint32_t>(n_ + 2);
  seg_desc = ws.take<uint4>(n_ + 2);
  long_seg = ws.take<char>(max_long * sizeof(LongSeg));
  leaf_desc = ws.take<uint2>(max_leaves);
  partial = ws.take<float>(max_leaves * dim);
  sc = ws.take<SegScratch>(1);
}

// "<prefix>(N kernels)" for the hand-written sort (histogram/scan + one kernel per digit), "<prefix>(cub)" else;
// the strings live for the life of the process (the profiler keeps the pointers).
static const char* sort_scope_name(const char* prefix, int kernels) {
  static std::map<std::pair<const char*, int>, std::string> names;
  auto& s = names[{prefix, kernels}];
  if (s.empty()) s = std::string(prefix) + (kernels ? "(" + std::to_string(kernels) + " kernels)" : "(cub)");
  return s.c_str();
}

int bits_for(uint32_t max_value) {
  int b = 1;
  while (b < 32 && (max_value >> b)) b++;
  return b;
}

// Phase 1: stable sort of (sk_in, sv_in) by key, then the segment heads. n_dev (optional, device memory): the
// true number of pairs (<= w.n, which sizes grids and scratch). Keys >= miss_key are "absent".
meepo_status seg_sort_heads(meepo_table* t, SegWork& w, uint32_t miss_key, bool count_updates, const uint32_t* n_dev,
                            cudaStream_t stream, const char* const* names, uint32_t expect_n, int chunk_shift) {
  const uint32_t n32 = w.n;
  {
    ProfScope ps(t, sort_scope_name(names[0], radix_sort_supported(n32, w.end_bit) ? (w.end_bit + 7) / 8 + 1 : 0), stream);
    if (radix_sort_supported(n32, w.end_bit)) {  // hand-written onesweep (radix_sort.cu); CUB only beyond 2^30 pairs
      MEEPO_TRY(radix_sort_pairs(t, w.cub_tmp, w.sk_in, w.sk_out, w.sv_in, w.sv_out, n32, w.end_bit, stream, n_dev,
                                 expect_n));
    } else {
      if (n_dev) return fail(MEEPO_EINVAL, "a device-side pair count needs the hand-written sort (< 2^30 pairs)");
      MEEPO_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, w.cub_bytes, (const uint32_t*)w.sk_in, w.sk_out,
                                                   (const uint32_t*)w.sv_in, w.sv_out, (int)n32, 0, w.end_bit,
                                                   stream));
    }
  }
  {
    ProfScope ps(t, names[1], stream);
    MEEPO_CUDA_TRY(cudaMemsetAsync(w.cstate, 0, w.cstate_bytes, stream));
    if (chunk_shift >= 0) MEEPO_CUDA_TRY(cudaMemsetAsync(w.sc->chunk_first, 0xFF, sizeof w.sc->chunk_first, stream));
    segments_kernel<<<w.ntiles, kCompactThreads, 0, stream>>>(w.sk_out, w.sv_out, n32, w.seg_start, w.seg_desc, miss_key,
                                                              w.sc, t->v.counters, count_updates ? 1 : 0, n_dev, chunk_shift,
                                                              compact_carve(w.cstate, t->err_word + kErrLookback));
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  return MEEPO_OK;
}

// Phase 2: reduce the gradient rows of every segment whose sort key lies in [r.key_lo, r.key_hi) in the normative
// order and hand the sum to the optimizer `mode` (MEEPO_SGD/ADAGRAD/ADAM) or store it (kStoreOnly) to
// reduce_out[key & mask] / *reduce_rows[key & mask]. May be called several times over one phase-1 result with
// disjoint key ranges (`again` = not the first call: the long-segment lists are reset first).
meepo_status seg_reduce(meepo_table* t, SegWork& w, const void* grads, int mode, const SegRange& r, void* reduce_out,
                        void* const* reduce_rows, cudaStream_t stream, cudaEvent_t grads_ready,
                        const char* const* names, bool again) {
  const uint32_t n32 = w.n;
  ApplyArgs a;
  a.grads = reinterpret_cast<const uint4*>(grads);
  a.sorted_slot = w.sk_out;
  a.sorted_idx = w.sv_out;
  a.seg_start = w.seg_start;
  a.seg_desc = w.seg_desc;
  a.sc = w.sc;
  a.long_seg = reinterpret_cast<LongSeg*>(w.long_seg);
  a.leaf_desc = w.leaf_desc;
  a.partial = w.partial;
  a.key_lo = r.key_lo;
  a.key_hi = r.key_hi;
  a.key_mask = r.key_mask;
  a.chunk = r.chunk;
  a.reduce_out = reinterpret_cast<uint4*>(reduce_out);
  a.reduce_rows = reinterpret_cast<uint4* const*>(reduce_rows);
  static const bool bulk_env = getenv("MEEPO_PEER_BULK") != nullptr && atoi(getenv("MEEPO_PEER_BULK")) != 0;
  a.bulk = (mode == kReduceStoreOnly && reduce_rows != nullptr && bulk_env) ? 1 : 0;
  uint32_t gl = 1;
  while (gl * 2 <= t->v.cpr && gl < 32) gl *= 2;
  a.group_lanes = gl;

  const bool bf16 = t->v.dtype == MEEPO_BF16;
  const void *k_apply = nullptr, *k_finish = nullptr;
  const bool pipelined = gl == t->v.cpr;  // one 16-byte chunk per lane: rows of <= 512 B with a power-of-two chunk count
  if (bf16)
    pick_apply<true>(mode, pipelined, k_apply, k_finish);
  else
    pick_apply<false>(mode, pipelined, k_apply, k_finish);
  const int leaf_v = t->row_bytes > 256 ? 2 : 1;  // 4-byte words per thread in the leaf kernel
  const void* k_leaf = bf16 ? (leaf_v == 2 ? (const void*)leaf_kernel<true, 2> : (const void*)leaf_kernel<true, 1>)
                            : (leaf_v == 2 ? (const void*)leaf_kernel<false, 2> : (const void*)leaf_kernel<false, 1>);
  void* args[] = {&t->v, &a};
  const uint64_t groups_per_block = 256 / gl;
  if (grads_ready) MEEPO_CUDA_TRY(cudaStreamWaitEvent(stream, grads_ready, 0));
  if (again) MEEPO_CUDA_TRY(cudaMemsetAsync(&w.sc->num_long, 0, 8, stream));  // num_long, num_leaves
  {
    ProfScope ps(t, names[2], stream);
    const uint64_t per_block = groups_per_block * (pipelined ? 2 : 1);
    const int grid = grid_for(t, k_apply, 256, 0, (n32 + per_block - 1) / per_block);
    MEEPO_CUDA_TRY(cudaLaunchKernel(k_apply, dim3(grid), dim3(256), args, 0, stream));
  }
  {
    ProfScope ps(t, names[3], stream);
    const uint32_t pieces = t->v.cpr * 4 / leaf_v;  // pieces per row, one thread each
    uint32_t tpl = std::min<uint32_t>(256, (pieces + 31) / 32 * 32);  // whole warps per leaf
    while (256 % tpl) tpl += 32;
    const uint64_t lpc = 256 / tpl;
    const int grid = grid_for(t, k_leaf, 256, 0, (w.max_leaves + lpc - 1) / lpc);
    void* largs[] = {&t->v, &a, &tpl};
    MEEPO_CUDA_TRY(cudaLaunchKernel(k_leaf, dim3(grid), dim3(256), largs, 0, stream));
  }
  {
    ProfScope ps(t, names[4], stream);
    const size_t smem = (size_t)t->v.dim * 4;
    const int grid2 = grid_for(t, k_finish, 256, smem, w.max_long);
    MEEPO_CUDA_TRY(cudaLaunchKernel(k_finish, dim3(grid2), dim3(256), args, smem, stream));
  }
  return MEEPO_OK;
}

// both phases over every key below `limit`
meepo_status run_segmented(meepo_table* t, SegWork& w, uint32_t limit, const void* grads, int mode,
                           void* reduce_out, cudaStream_t stream, cudaEvent_t grads_ready,
                           const char* const* names, void* const* reduce_rows) {
  MEEPO_TRY(seg_sort_heads(t, w, limit, mode != kStoreOnly, nullptr, stream, names));
  return seg_reduce(t, w, grads, mode, SegRange{0u, limit, 0xFFFFFFFFu, -1}, reduce_out, reduce_rows, stream, grads_ready,
                    names, false);
}

meepo_status launch_apply_gradients(meepo_table* t, const uint64_t* keys, const void* grads, uint64_t n,
                                    cudaStream_t stream, cudaEvent_t grads_ready, const uint32_t* bag_offsets,
                                    uint32_t n_bags) {
  if (n == 0) return MEEPO_OK;
  const int end_bit = bits_for(t->v.slots);
  MEEPO_TRY(t->ws.reserve(SegWork::bytes(n, t->v.dim, end_bit), stream));
  SegWork w;
  w.take(t->ws, n, t->v.dim, end_bit);
  {
    ProfScope ps(t, "apply.grad_slots", stream);
    const int grid = grid_for(t, (const void*)grad_slots_kernel, 256, 0, (n + 255) / 256);
    const uint32_t cache_n = t->cache_valid ? (uint32_t)std::min<uint64_t>(t->cache_n, n) : 0u;
    grad_slots_kernel<<<grid, 256, 0, stream>>>(t->v, keys, (uint32_t)n, w.sk_in, w.sv_in, t->cache, cache_n, bag_offsets,
                                                n_bags);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  static const char* const names[5] = {"apply.radix_sort", "apply.segments",
                                       "apply.reduce_optimizer", "apply.long_leaves", "apply.long_finish"};
  return run_segmented(t, w, t->v.slots, grads, t->v.opt, nullptr, stream, grads_ready, names);
}

}  // namespace meepo

using namespace meepo;

extern "C" MEEPO_API meepo_status meepo_apply_gradients(meepo_table* t, const uint64_t* keys,
                                                        const void* grads, uint64_t n, void* stream) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!keys || !grads)) return fail(MEEPO_EINVAL, "null buffer");
  DeviceGuard guard(t->device);
  VerbScope vs(t, (cudaStream_t)stream);
  MEEPO_TRY(vs.rc);
  return launch_apply_gradients(t, keys, grads, n, (cudaStream_t)stream);
}
