// shard.cu — device helpers of the key-hash-sharded table (include/meepo.h "sharding helpers",
// SURVEY K11 + the dedup-before-exchange rule of section 5):
//   meepo_shard_partition    stable counting sort of a batch by owner(key, G)
//   meepo_reduce_duplicates  batch-level dedup (CAS into an L2-resident scratch table) and, for the
//                            backward path, the fixed-shape pre-reduction of duplicate gradients
//   meepo_gather_rows        rows_out[i] = rows_in[index[i]], 16-byte vectorised (un-permute/expand)
#include "compact.cuh"
#include "table.h"

namespace meepo {

constexpr int kPartTile = 2048;  // keys per CTA in the partition passes
constexpr int kMaxShards = 32;

// ---------------------------------------------------------------------------------------------
// single-CTA exclusive scan over the (tile, shard) histogram of the owner partition (compact.cuh)
__global__ void __launch_bounds__(1024) excl_scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                         uint32_t n) {
  block_excl_scan_1024(in, out, n, nullptr);
}

// ---------------------------------------------------------------------------------------------
// partition by owner
__global__ void __launch_bounds__(256) part_hist_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t G,
                                                        uint32_t ntiles, uint32_t* __restrict__ tile_hist) {
  __shared__ uint32_t hist[kMaxShards];
  if (threadIdx.x < kMaxShards) hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kPartTile;
#pragma unroll
  for (int k = 0; k < kPartTile / 256; k++) {
    const uint32_t i = base + k * 256 + threadIdx.x;
    const bool ok = i < n;
    const uint32_t g = ok ? owner_of(__ldg(keys + i), G) : G;  // tail lanes form their own group
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, g);
    if (ok && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&hist[g], (uint32_t)__popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < G) tile_hist[threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(256) part_scatter_kernel(const uint64_t* __restrict__ keys, uint32_t n, uint32_t G,
                                                           uint32_t ntiles, const uint32_t* __restrict__ tile_off,
                                                           uint64_t* __restrict__ counts_out,
                                                           uint32_t* __restrict__ perm_out,
                                                           uint64_t* __restrict__ keys_sorted_out) {
  __shared__ uint32_t base_s[kMaxShards];
  __shared__ uint32_t warp_cnt[8][kMaxShards];
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < G) base_s[threadIdx.x] = tile_off[threadIdx.x * ntiles + blockIdx.x];
  if (blockIdx.x == 0 && threadIdx.x < G)
    counts_out[threadIdx.x] = (uint64_t)(tile_off[(threadIdx.x + 1) * ntiles] - tile_off[threadIdx.x * ntiles]);
  const uint32_t base = blockIdx.x * kPartTile;
#pragma unroll 1
  for (int k = 0; k < kPartTile / 256; k++) {
    for (uint32_t x = threadIdx.x; x < 8 * kMaxShards; x += 256) (&warp_cnt[0][0])[x] = 0;
    __syncthreads();
    const uint32_t i = base + k * 256 + threadIdx.x;
    const bool ok = i < n;
    const uint64_t key = ok ? __ldg(keys + i) : 0;
    const uint32_t g = ok ? owner_of(key, G) : G;
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, g);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (ok && rank_in_warp == 0) warp_cnt[w][g] = __popc(peers);
    __syncthreads();
    if (ok) {
      uint32_t before = 0;
      for (uint32_t j = 0; j < w; j++) before += warp_cnt[j][g];
      const uint32_t pos = base_s[g] + before + rank_in_warp;
      if (perm_out) perm_out[pos] = i;
      if (keys_sorted_out) keys_sorted_out[pos] = key;
    }
    __syncthreads();
    if (threadIdx.x < G) {
      uint32_t tot = 0;
      for (int j = 0; j < 8; j++) tot += warp_cnt[j][threadIdx.x];
      base_s[threadIdx.x] += tot;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// batch-level dedup
__global__ void __launch_bounds__(256) dedup_insert_kernel(const uint64_t* __restrict__ keys, uint32_t n,
                                                           uint64_t* __restrict__ scratch, uint32_t mask,
                                                           uint32_t* __restrict__ pos, uint32_t* __restrict__ canon_cell,
                                                           const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t key = __ldg(keys + i);
    uint32_t p = kNil;
    if (key_valid(key)) {
      uint32_t h = (uint32_t)(mix64(key ^ 0x5851F42D4C957F2Dull) >> 32) & mask;
      while (true) {  // read first: duplicates of a hot key must not serialise on one atomic
        unsigned long long old = __ldcg(reinterpret_cast<const unsigned long long*>(scratch + h));
        if (old == MEEPO_KEY_EMPTY)
          old = atomicCAS(reinterpret_cast<unsigned long long*>(scratch + h), (unsigned long long)MEEPO_KEY_EMPTY,
                          (unsigned long long)key);
        if (old == MEEPO_KEY_EMPTY || old == key) {
          // the occurrence whose CAS created the cell is the key's canonical one (where a sharded forward pass
          // has the owner deliver the row; the other occurrences copy it locally)
          if (old == MEEPO_KEY_EMPTY) canon_cell[h] = i;
          p = h;
          break;
        }
        h = (h + 1) & mask;
      }
    }
    pos[i] = p;
  }
}

// Unique ids in BATCH order: element i is the canonical occurrence of its key if its CAS created the key's
// scratch cell; the canonical occurrences, compacted in batch order (compact.cuh), number the unique keys. So
// unique id order == order of first... of canonical appearance in the batch, which is also the order in which the
// sharded push hands out positions in the owners' lanes: everything downstream that walks the unique keys in id
// order (the sender-side pre-reduction storing into the owners' windows) then walks remote memory forwards.
__global__ void __launch_bounds__(kCompactThreads) uid_compact_kernel(const uint64_t* __restrict__ keys, uint32_t n,
                                                                      const uint32_t* __restrict__ pos,
                                                                      const uint32_t* __restrict__ canon_cell,
                                                                      uint32_t* __restrict__ uid_of_cell,
                                                                      uint64_t* __restrict__ unique_out,
                                                                      uint32_t* __restrict__ canon_out,
                                                                      unsigned long long* __restrict__ n_unique,
                                                                      CompactState cs,
                                                                      const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  CompactTile ct = compact_begin(cs, n);
  unsigned flags = 0;
  uint32_t cell[kCompactItems];
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) {
    const uint64_t i = ct.pos(k);
    cell[k] = i < n ? pos[i] : kNil;
    if (cell[k] != kNil && canon_cell[cell[k]] == (uint32_t)i) flags |= 1u << k;
  }
  compact_rank(ct, flags, cs);
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) {
    if ((flags >> k) & 1u) {
      const uint32_t u = (uint32_t)ct.rank(k), i = (uint32_t)ct.pos(k);
      uid_of_cell[cell[k]] = u;
      unique_out[u] = keys[i];
      if (canon_out) canon_out[u] = i;
    }
  }
  if (ct.last && threadIdx.x == 0) *n_unique = ct.base + ct.tile_total;
}

__global__ void __launch_bounds__(256) dedup_inverse_kernel(const uint32_t* __restrict__ pos, uint32_t n,
                                                            const uint32_t* __restrict__ uid_of_slot,
                                                            uint32_t* __restrict__ inverse_out,
                                                            uint32_t* __restrict__ sort_key,
                                                            uint32_t* __restrict__ sort_val,
                                                            uint32_t* __restrict__ occurrences,
                                                            const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  const uint32_t lane = threadIdx.x & 31u;
  // whole warps iterate together (the occurrence count below is folded per warp)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i - lane < n; i += gridDim.x * blockDim.x) {
    const bool in = i < n;
    const uint32_t p = in ? pos[i] : kNil;
    const uint32_t u = p == kNil ? kNil : uid_of_slot[p];
    if (in && inverse_out) inverse_out[i] = u;
    if (occurrences) {
      // one atomic per distinct key and warp: the hottest Zipf key is 8% of a batch, and 84K atomics on ONE address
      // serialise in L2 (0.1 ms of a 1M-key batch's 0.17 ms dedup)
      const unsigned peers = __match_any_sync(0xFFFFFFFFu, u);
      if (u != kNil && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(occurrences + u, (uint32_t)__popc(peers));
    }
    if (in && sort_key) {
      sort_key[i] = u == kNil ? n : u;  // invalid keys sort last and are skipped
      sort_val[i] = i;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dense row gather: one warp moves 32 rows as a flat array of 16-byte chunks
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ in,
                                                          const uint32_t* __restrict__ index, uint32_t n,
                                                          uint32_t cpr, uint4* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t ntiles = (n + 31u) >> 5;
  for (uint32_t tile = warp; tile < ntiles; tile += nwarps) {
    const uint32_t i = tile * 32u + lane;
    const uint32_t src = i < n ? __ldg(index + i) : kNil;
    const uint32_t tile_rows = min(32u, n - tile * 32u);
    const uint32_t chunks = tile_rows * cpr;
    uint4* out_tile = out + (size_t)tile * 32u * cpr;
    for (uint32_t c0 = 0; c0 < chunks; c0 += 128) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t c = c0 + u * 32 + lane;
        const uint32_t j = min(c / cpr, 31u);
        const uint32_t s = __shfl_sync(0xFFFFFFFFu, src, j);
        v[u] = make_uint4(0, 0, 0, 0);
        if (c < chunks && s != kNil) v[u] = ld_nc(in + (size_t)s * cpr + (c - j * cpr));
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t c = c0 + u * 32 + lane;
        if (c < chunks) st_stream(out_tile + c, v[u]);
      }
    }
  }
}

// batch-level dedup (+ optional pre-reduction of the duplicate gradients); scratch comes out of the
// table workspace, which the caller has reserved for at least dedup_bytes().
static uint64_t dedup_cells(uint64_t n) {  // scratch cells: a power of two >= 2n
  uint64_t m = 1024;
  while (m < 2 * n) m <<= 1;
  return m;
}

size_t dedup_bytes(const meepo_table* t, uint64_t n, bool with_grads) {
  if (n == 0) return 256;
  const uint64_t m = dedup_cells(n);
  size_t need = Workspace::pad((size_t)m * 8) + Workspace::pad(n * 4) + 2 * Workspace::pad((size_t)m * 4) +
                Workspace::pad(compact_state_bytes(n)) + 4096;
  if (with_grads) need += SegWork::bytes(n, t->v.dim, bits_for((uint32_t)n));
  return need;
}

// phase 1: unique keys, inverse, occurrences; with_grads also prepares the (uid, batch index) sort input
meepo_status dedup_hash(meepo_table* t, const uint64_t* keys, uint64_t n, const DedupOut& o, bool with_grads,
                        SegWork& w, cudaStream_t stream, const uint32_t* skip) {
  if (n == 0) {
    MEEPO_CUDA_TRY(cudaMemsetAsync(o.n_unique, 0, 8, stream));
    return MEEPO_OK;
  }
  if (n > (1ull << 30)) return fail(MEEPO_EINVAL, "dedup: batch larger than 2^30 keys");
  const uint32_t m = (uint32_t)dedup_cells(n);  // <= 2^31
  const int end_bit = bits_for((uint32_t)n);
  uint64_t* scratch = t->ws.take<uint64_t>(m);
  uint32_t* pos = t->ws.take<uint32_t>(n);
  uint32_t* uid_of_slot = t->ws.take<uint32_t>(m);
  uint32_t* canon_cell = t->ws.take<uint32_t>(m);
  const size_t cbytes = compact_state_bytes(n);
  char* cstate = t->ws.take<char>(cbytes);
  if (with_grads) w.take(t->ws, n, t->v.dim, end_bit);
  ProfScope ps(t, "dedup.hash(3 kernels)", stream);
  MEEPO_CUDA_TRY(cudaMemsetAsync(scratch, 0xFF, (size_t)m * 8, stream));
  MEEPO_CUDA_TRY(cudaMemsetAsync(cstate, 0, cbytes, stream));
  if (o.occurrences) MEEPO_CUDA_TRY(cudaMemsetAsync(o.occurrences, 0, n * 4, stream));
  const int grid = grid_for(t, (const void*)dedup_insert_kernel, 256, 0, (n + 255) / 256);
  dedup_insert_kernel<<<grid, 256, 0, stream>>>(keys, (uint32_t)n, scratch, m - 1, pos, canon_cell, skip);
  uid_compact_kernel<<<compact_tiles(n), kCompactThreads, 0, stream>>>(keys, (uint32_t)n, pos, canon_cell, uid_of_slot,
                                                                       o.unique_keys, o.canon,
                                                                       (unsigned long long*)o.n_unique,
                                                                       compact_carve(cstate, t->err_word + kErrLookback),
                                                                       skip);
  dedup_inverse_kernel<<<grid, 256, 0, stream>>>(pos, (uint32_t)n, uid_of_slot, o.inverse,
                                                 with_grads ? w.sk_in : nullptr, with_grads ? w.sv_in : nullptr,
                                                 o.occurrences, skip);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

// phase 2: fixed-shape sum of each unique key's gradient rows, rounded to the table dtype, stored to
// o.grads_out[uid] or through o.grad_rows[uid] (a pointer per unique key, e.g. into a peer's window)
meepo_status dedup_reduce(meepo_table* t, SegWork& w, const void* grads, uint64_t n, const DedupOut& o,
                          cudaStream_t stream) {
  if (n == 0) return MEEPO_OK;
  static const char* const names[5] = {"dedup.radix_sort", "dedup.segments", "dedup.reduce_store",
                                       "dedup.long_leaves", "dedup.long_finish"};
  return run_segmented(t, w, (uint32_t)n, grads, kReduceStoreOnly, o.grads_out, stream, nullptr, names, o.grad_rows);
}

meepo_status dedup_run(meepo_table* t, const uint64_t* keys, const void* grads, uint64_t n, const DedupOut& o,
                       cudaStream_t stream) {
  SegWork w;
  MEEPO_TRY(dedup_hash(t, keys, n, o, grads != nullptr, w, stream));
  if (grads) MEEPO_TRY(dedup_reduce(t, w, grads, n, o, stream));
  return MEEPO_OK;
}

}  // namespace meepo

using namespace meepo;

extern "C" {

MEEPO_API meepo_status meepo_shard_partition(meepo_table* t, const uint64_t* keys, uint64_t n,
                                             uint32_t num_shards, uint64_t* counts_out, uint32_t* perm_out,
                                             uint64_t* keys_sorted_out, void* stream_) {
  if (!t || !counts_out) return fail(MEEPO_EINVAL, "null argument");
  if (num_shards == 0 || num_shards > (uint32_t)kMaxShards) return fail(MEEPO_EINVAL, "num_shards must be 1..32");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && !keys) return fail(MEEPO_EINVAL, "null buffer");
  DeviceGuard guard(t->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  if (n == 0) {
    MEEPO_CUDA_TRY(cudaMemsetAsync(counts_out, 0, 8 * num_shards, stream));
    return MEEPO_OK;
  }
  const uint32_t ntiles = (uint32_t)((n + kPartTile - 1) / kPartTile);
  const size_t cells = (size_t)ntiles * num_shards;
  MEEPO_TRY(t->ws.reserve(2 * Workspace::pad((cells + 1) * 4) + 1024, stream));
  uint32_t* hist = t->ws.take<uint32_t>(cells + 1);
  uint32_t* off = t->ws.take<uint32_t>(cells + 1);
  ProfScope ps(t, "shard.partition(3 kernels)", stream);
  part_hist_kernel<<<ntiles, 256, 0, stream>>>(keys, (uint32_t)n, num_shards, ntiles, hist);
  excl_scan_kernel<<<1, 1024, 0, stream>>>(hist, off, (uint32_t)cells);
  part_scatter_kernel<<<ntiles, 256, 0, stream>>>(keys, (uint32_t)n, num_shards, ntiles, off, counts_out, perm_out,
                                                  keys_sorted_out);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_reduce_duplicates(meepo_table* t, const uint64_t* keys, const void* grads, uint64_t n,
                                               uint64_t* unique_keys_out, void* grads_out, uint32_t* inverse_out,
                                               uint64_t* n_unique_out, void* stream_) {
  if (!t || !n_unique_out) return fail(MEEPO_EINVAL, "null argument");
  if (n > (1ull << 30)) return fail(MEEPO_EINVAL, "batch too large (at most 2^30 keys)");
  if (n && (!keys || !unique_keys_out)) return fail(MEEPO_EINVAL, "null buffer");
  if ((grads == nullptr) != (grads_out == nullptr)) return fail(MEEPO_EINVAL, "grads and grads_out go together");
  DeviceGuard guard(t->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  MEEPO_TRY(t->ws.reserve(dedup_bytes(t, n, grads != nullptr), stream));
  return dedup_run(t, keys, grads, n, DedupOut{unique_keys_out, grads_out, inverse_out, n_unique_out, nullptr, nullptr, nullptr}, stream);
}

MEEPO_API meepo_status meepo_gather_rows(meepo_table* t, const void* rows_in, const uint32_t* index, uint64_t n,
                                         void* rows_out, void* stream_) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!rows_in || !index || !rows_out)) return fail(MEEPO_EINVAL, "null buffer");
  if (n == 0) return MEEPO_OK;
  DeviceGuard guard(t->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  ProfScope ps(t, "shard.gather_rows", stream);
  const uint64_t tiles = (n + 31) / 32;
  const int grid = grid_for(t, (const void*)gather_rows_kernel, 256, 0, (tiles + 7) / 8);
  gather_rows_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(rows_in), index, (uint32_t)n, t->v.cpr,
                                               reinterpret_cast<uint4*>(rows_out));
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

}  // extern "C"
