// pool.cu — the pooled (bag) verbs (include/meepo.h "Pooling"): find_or_insert / lookup fused with
// the sum / mean pooling that follows them in CTR models, and the matching backward verb.
//
// Forward = two kernels. probe_slots resolves every key (the tile body of probe_gather.cuh in its NOOUT mode:
// probe, CAS insert, promotion from the host tier, scores, counters, per-key status — but no per-key row
// leaves the kernel; a CAS winner writes the new / restored row into the arena). pooled_gather then has one
// group of lanes per bag walk the bag's slots in order, with 8 row loads in flight per lane, accumulate in fp32
// registers and store ONE row per bag. Against the unfused sequence this removes the [n][dim] row write and its
// read-back by the pooling op: per key R bytes are read instead of R read + R written + R read.
// Backward = apply_gradients with the bag index as the sort value (update.cu grad_slots_kernel): the reduce
// kernels read the bag's gradient row for every occurrence; no [n][dim] gradient is ever expanded.
#include "optimizer.cuh"
#include "probe_gather.cuh"

namespace meepo {

template <bool INSERT, bool TIER>
__global__ void __launch_bounds__(256, 4) probe_slots_kernel(TableView t, const uint64_t* __restrict__ keys, uint32_t n,
                                                             uint8_t* __restrict__ status, NewList nl, SlotCache sc,
                                                             uint32_t* __restrict__ tslab) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t ntiles = (n + 31u) >> 5;
  TileCounts cnt;
  __shared__ uint32_t sc_slot[kScoreCells], sc_freq[kScoreCells];
  const ScoreCache scache{sc_slot, sc_freq};
  score_cache_init(t, scache);
  for (uint32_t tile = warp; tile < ntiles; tile += nwarps) {
    const uint32_t i = tile * 32u + lane;
    const uint32_t tile_keys = min(32u, n - tile * 32u);
    const uint64_t key = i < n ? __ldg(keys + i) : MEEPO_KEY_EMPTY;
    probe_gather_tile<0, INSERT, false, TIER, true>(t, key, i < n, tile_keys, nullptr, status ? status + i : nullptr,
                                                    sc.slots + i, sc.keys ? sc.keys + i : nullptr, 1u,
                                                    nl.slots ? nl.slots + i : nullptr, cnt, scache, lane, 0ull,
                                                    tslab ? tslab + i : nullptr);
  }
  score_cache_flush(t, scache);
  flush_tile_counts(t, cnt, lane);
}

// One group of GL lanes per bag (GL = the power of two >= cpr, at most 32; lane gl owns chunks gl, gl + GL, ...).
template <bool BF16>
__global__ void __launch_bounds__(256) pooled_gather_kernel(TableView t, const uint32_t* __restrict__ slot,
                                                            const uint32_t* __restrict__ tslab,
                                                            const uint32_t* __restrict__ offsets, uint32_t n,
                                                            uint32_t n_bags, int pool, uint32_t GL,
                                                            uint4* __restrict__ out) {
  constexpr int E = Chunk<BF16>::E;
  constexpr int U = 8;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gl = lane & (GL - 1);
  const uint32_t groups_per_block = blockDim.x / GL;
  const uint32_t ngroups = gridDim.x * groups_per_block;
  const uint32_t cpr = t.cpr;
  for (uint32_t b = blockIdx.x * groups_per_block + threadIdx.x / GL; b < n_bags; b += ngroups) {
    const uint32_t lo = min(__ldg(offsets + b), n);
    const uint32_t hi = min(max(__ldg(offsets + b + 1), lo), n);
    for (uint32_t q = gl; q < cpr; q += GL) {
      float acc[E];
#pragma unroll
      for (int e = 0; e < E; e++) acc[e] = 0.0f;
      for (uint32_t i = lo; i < hi; i += U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          v[u] = make_uint4(0, 0, 0, 0);
          if (i + u < hi) {
            const uint32_t s = __ldg(slot + i + u);
            if (s != kNil) {
              v[u] = ld_stream(t.rows + (size_t)s * cpr + q);
            } else if (tslab) {  // lookup read-through: the row sits in the host tier
              const uint32_t d = __ldg(tslab + i + u);
              if (d != kNil) v[u] = tier_tuple(t, d).rows[q];
            }
          }
        }
        // in bag order; a missing row is +0.0 and leaves the sum as it is (the sum is never -0.0)
#pragma unroll
        for (int u = 0; u < U; u++) {
          float w[E];
          widen<BF16>(v[u], w);
#pragma unroll
          for (int e = 0; e < E; e++) acc[e] = __fadd_rn(acc[e], w[e]);
        }
      }
      if (pool == MEEPO_POOL_MEAN && hi > lo) {
        const float len = (float)(hi - lo);
#pragma unroll
        for (int e = 0; e < E; e++) acc[e] = __fdiv_rn(acc[e], len);
      }
      st_stream(out + (size_t)b * cpr + q, narrow<BF16>(acc));
    }
  }
}

// MEAN backward: scaled[b] = round_to_dtype(widen(bag_grads[b]) / len_b)
template <bool BF16>
__global__ void __launch_bounds__(256) mean_scale_kernel(const uint4* __restrict__ bag_grads,
                                                         const uint32_t* __restrict__ offsets, uint32_t n_bags,
                                                         uint32_t cpr, uint4* __restrict__ scaled) {
  constexpr int E = Chunk<BF16>::E;
  const uint64_t total = (uint64_t)n_bags * cpr;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = (uint32_t)(c / cpr);
    const uint32_t o0 = __ldg(offsets + b), o1 = __ldg(offsets + b + 1);
    float g[E];
    widen<BF16>(ld_nc(bag_grads + c), g);
    if (o1 > o0) {
      const float len = (float)(o1 - o0);
#pragma unroll
      for (int e = 0; e < E; e++) g[e] = __fdiv_rn(g[e], len);
    }
    scaled[c] = narrow<BF16>(g);
  }
}

static meepo_status pooled_forward(meepo_table* t, const uint64_t* keys, uint64_t n, const uint32_t* offsets,
                                   uint64_t n_bags, int32_t pool, void* pooled_out, uint8_t* status_out, bool insert,
                                   cudaStream_t stream) {
  const bool tier = t->v.tier.slabs != 0;
  const size_t extra = 2 * Workspace::pad(n * 4) + 512;
  MEEPO_TRY(probe_gather_begin(t, n, insert, stream, extra));
  uint32_t* slots = nullptr;
  uint32_t* tslab = nullptr;
  if (n) {
    SlotCache sc{nullptr, nullptr};
    if (t->cache_n) {  // the slots double as the cache of the following apply_gradients(_pooled)
      sc = t->cache;
      t->cache_off = n;
      t->cache_valid = true;
    } else {
      sc.slots = t->ws.take<uint32_t>(n);
    }
    slots = sc.slots;
    if (tier && !insert) tslab = t->ws.take<uint32_t>(n);
    NewList nl{insert ? t->cur_new.slots : nullptr};
    t->cur_new_off = n;
    const void* kern = insert ? (tier ? (const void*)probe_slots_kernel<true, true> : (const void*)probe_slots_kernel<true, false>)
                              : (tier ? (const void*)probe_slots_kernel<false, true> : (const void*)probe_slots_kernel<false, false>);
    uint32_t n32 = (uint32_t)n;
    void* args[] = {&t->v, &keys, &n32, &status_out, &nl, &sc, &tslab};
    ProfScope ps(t, insert ? "find_or_insert_pooled.probe" : "lookup_pooled.probe", stream);
    const int grid = grid_for(t, kern, 256, 0, ((n + 31) / 32 + 7) / 8);
    MEEPO_CUDA_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(256), args, 0, stream));
  }
  MEEPO_TRY(probe_gather_end(t, n, insert, stream));
  if (n_bags) {
    uint32_t gl = 1;
    while (gl < t->v.cpr && gl < 32) gl *= 2;
    const uint64_t groups_per_block = 256 / gl;
    const void* kern = t->v.dtype == MEEPO_BF16 ? (const void*)pooled_gather_kernel<true> : (const void*)pooled_gather_kernel<false>;
    uint32_t n32 = (uint32_t)n, nb32 = (uint32_t)n_bags;
    int pool_i = pool;
    uint4* out = reinterpret_cast<uint4*>(pooled_out);
    void* args[] = {&t->v, &slots, &tslab, &offsets, &n32, &nb32, &pool_i, &gl, &out};
    ProfScope ps(t, insert ? "find_or_insert_pooled.gather" : "lookup_pooled.gather", stream);
    const int grid = grid_for(t, kern, 256, 0, (n_bags + groups_per_block - 1) / groups_per_block);
    MEEPO_CUDA_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(256), args, 0, stream));
  }
  return MEEPO_OK;
}

}  // namespace meepo

using namespace meepo;

static meepo_status check_pooled(meepo_table* t, const void* keys, uint64_t n, const void* offsets, uint64_t n_bags,
                                 int32_t pool, const void* rows) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull || n_bags > 0xFFFFFFFEull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (pool != MEEPO_POOL_SUM && pool != MEEPO_POOL_MEAN) return fail(MEEPO_EINVAL, "bad pooling mode");
  if ((n && !keys) || (n_bags && (!offsets || !rows))) return fail(MEEPO_EINVAL, "null buffer");
  if (n && !n_bags) return fail(MEEPO_EINVAL, "keys without bags");
  return MEEPO_OK;
}

extern "C" {

MEEPO_API meepo_status meepo_find_or_insert_pooled(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                   const uint32_t* offsets, uint64_t n_bags, int32_t pool,
                                                   void* pooled_out, uint8_t* status_out, void* stream) {
  MEEPO_TRY(check_pooled(t, keys, n, offsets, n_bags, pool, pooled_out));
  DeviceGuard guard(t->device);
  VerbScope vs(t, (cudaStream_t)stream);
  MEEPO_TRY(vs.rc);
  return pooled_forward(t, keys, n, offsets, n_bags, pool, pooled_out, status_out, true, (cudaStream_t)stream);
}

MEEPO_API meepo_status meepo_lookup_pooled(meepo_table* t, const uint64_t* keys, uint64_t n, const uint32_t* offsets,
                                           uint64_t n_bags, int32_t pool, void* pooled_out, uint8_t* found_out,
                                           void* stream) {
  MEEPO_TRY(check_pooled(t, keys, n, offsets, n_bags, pool, pooled_out));
  DeviceGuard guard(t->device);
  VerbScope vs(t, (cudaStream_t)stream);
  MEEPO_TRY(vs.rc);
  return pooled_forward(t, keys, n, offsets, n_bags, pool, pooled_out, found_out, false, (cudaStream_t)stream);
}

MEEPO_API meepo_status meepo_apply_gradients_pooled(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                    const uint32_t* offsets, uint64_t n_bags, int32_t pool,
                                                    const void* bag_grads, void* stream_) {
  MEEPO_TRY(check_pooled(t, keys, n, offsets, n_bags, pool, bag_grads));
  if (n == 0) return MEEPO_OK;
  DeviceGuard guard(t->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  const void* grads = bag_grads;
  if (pool == MEEPO_POOL_MEAN) {
    // the scaled rows live in the head of the workspace; launch_apply_gradients reserves behind them
    // (Workspace::reserve resets the bump pointer), so they get their own allocation inside the table
    const size_t bytes = (size_t)n_bags * t->row_bytes;
    if (bytes > t->pool_scaled_bytes) {
      MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
      cudaFree(t->pool_scaled);
      t->pool_scaled = nullptr;
      t->pool_scaled_bytes = 0;
      MEEPO_CUDA_TRY(cudaMalloc(&t->pool_scaled, bytes + bytes / 8));
      t->pool_scaled_bytes = bytes + bytes / 8;
    }
    ProfScope ps(t, "apply_pooled.mean_scale", stream);
    const uint64_t chunks = n_bags * t->v.cpr;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((chunks + 255) / 256, (uint64_t)t->num_sms * 8));
    if (t->v.dtype == MEEPO_BF16)
      mean_scale_kernel<true><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(bag_grads), offsets,
                                                        (uint32_t)n_bags, t->v.cpr, reinterpret_cast<uint4*>(t->pool_scaled));
    else
      mean_scale_kernel<false><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(bag_grads), offsets,
                                                         (uint32_t)n_bags, t->v.cpr, reinterpret_cast<uint4*>(t->pool_scaled));
    MEEPO_CUDA_TRY(cudaGetLastError());
    grads = t->pool_scaled;
  }
  return launch_apply_gradients(t, keys, grads, n, stream, nullptr, offsets, (uint32_t)n_bags);
}

}  // extern "C"
