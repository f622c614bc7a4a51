// io.cu — bulk dump / load (include/meepo.h "bulk dump / load"): export_buffers,
// import_buffers and the "MEEPOTB1" file format. The live keys are sorted with the library's own radix sort
// (two stable passes over the 32-bit halves of the key); the row movement reuses the 16-byte gather/scatter kernels.

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>

#include "compact.cuh"
#include "table.h"

namespace meepo {

// live (key, slot) pairs in slot order: one ordered compaction pass (compact.cuh)
// (delta: only the slots whose dirty bit is set — meepo.h "Incremental export")
__global__ void __launch_bounds__(kCompactThreads) live_compact_kernel(TableView t, uint32_t m,
                                                                       uint64_t* __restrict__ live_key,
                                                                       uint32_t* __restrict__ live_slot,
                                                                       CompactState cs, int delta) {
  CompactTile ct = compact_begin(cs, m);
  unsigned flags = 0;
  uint64_t key[kCompactItems];
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) {
    const uint64_t p = ct.pos(k);
    key[k] = p < m ? *key_ptr(t, (uint32_t)p) : MEEPO_KEY_EMPTY;
    if (delta && p < m && !((t.dirty[p >> 5] >> (p & 31u)) & 1u)) key[k] = MEEPO_KEY_EMPTY;
    if (key[k] != MEEPO_KEY_EMPTY) flags |= 1u << k;
  }
  compact_rank(ct, flags, cs);
#pragma unroll
  for (int k = 0; k < kCompactItems; k++) {
    if ((flags >> k) & 1u) {
      const uint64_t u = ct.rank(k);
      live_key[u] = key[k];
      live_slot[u] = (uint32_t)ct.pos(k);
    }
  }
}

// out[j] = arena[slot[j]] / arena[slot[i]] = in[i]: `cpr` 16-byte chunks per row, one warp per 32 rows
__global__ void __launch_bounds__(256) arena_gather_kernel(const uint4* __restrict__ arena,
                                                           const uint32_t* __restrict__ slot, uint32_t n, uint32_t cpr,
                                                           uint4* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t j = warp; j < n; j += nwarps) {
    const uint32_t s = slot[j];
    for (uint32_t q = lane; q < cpr; q += 32)
      out[(size_t)j * cpr + q] = s == kNil ? make_uint4(0, 0, 0, 0) : arena[(size_t)s * cpr + q];
  }
}
__global__ void __launch_bounds__(256) arena_scatter_kernel(uint4* __restrict__ arena, const uint32_t* __restrict__ slot,
                                                            uint32_t n, uint32_t cpr, const uint4* __restrict__ in,
                                                            uint4 fill, int use_fill) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < n; i += nwarps) {
    const uint32_t s = slot[i];
    if (s == kNil) continue;
    for (uint32_t q = lane; q < cpr; q += 32) arena[(size_t)s * cpr + q] = use_fill ? fill : in[(size_t)i * cpr + q];
  }
}

__global__ void export_meta_kernel(TableView t, const uint32_t* __restrict__ slot, uint32_t n,
                                   uint64_t* __restrict__ scores, uint32_t* __restrict__ steps) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t s = slot[j];
    if (scores) {
      const uint2 sc = t.scores ? t.scores[s] : make_uint2(0, 0);
      scores[j] = ((uint64_t)sc.y << 32) | sc.x;
    }
    if (steps) steps[j] = t.steps ? t.steps[s] : 0u;
  }
}
// number of dirty slots (a released slot never keeps its bit: evict.cu release_kernel)
__global__ void __launch_bounds__(256) dirty_count_kernel(const uint32_t* __restrict__ dirty, uint32_t words,
                                                          unsigned long long* __restrict__ out) {
  uint32_t c = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) c += __popc(dirty[i]);
  c = __reduce_add_sync(0xFFFFFFFFu, c);
  if ((threadIdx.x & 31u) == 0 && c) atomicAdd(out, (unsigned long long)c);
}
__global__ void import_meta_kernel(TableView t, const uint32_t* __restrict__ slot, uint32_t n,
                                   const uint64_t* __restrict__ scores, const uint32_t* __restrict__ steps) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t s = slot[i];
    if (s == kNil) continue;
    if (t.scores) {
      const uint64_t sc = scores ? scores[i] : 0ull;
      t.scores[s] = make_uint2((uint32_t)sc, (uint32_t)(sc >> 32));
    }
    if (t.steps) t.steps[s] = steps ? steps[i] : 0u;
    mark_dirty(t, s);
  }
}

// insert-or-find for import / readmit: one thread per key, claimed slots listed for publish_kernel
__global__ void __launch_bounds__(256) import_probe_kernel(TableView t, const uint64_t* __restrict__ keys, uint32_t n,
                                                           uint32_t* __restrict__ slot_out,
                                                           uint8_t* __restrict__ status_out, NewList nl) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Probe pr = probe_find_or_insert(t, keys[i]);
    slot_out[i] = pr.slot;
    if (status_out) status_out[i] = (uint8_t)pr.status;
    nl.slots[i] = pr.winner ? pr.slot : kNil;
    if (pr.status == MEEPO_KEY_FULL) atomicAdd(t.counters + C_FULL, 1ull);
  }
}

meepo_status import_probe_launch(meepo_table* t, const uint64_t* keys, uint64_t n, uint32_t* slot_out,
                                 uint8_t* status_out, NewList nl, cudaStream_t stream) {
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)t->num_sms * 8));
  import_probe_kernel<<<grid, 256, 0, stream>>>(t->v, keys, (uint32_t)n, slot_out, status_out, nl);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

static uint32_t float_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

meepo_status live_size(meepo_table* t, uint64_t* out) {
  unsigned long long v = 0;
  MEEPO_CUDA_TRY(cudaDeviceSynchronize());
  MEEPO_CUDA_TRY(cudaMemcpy(&v, t->dstate->counters + C_SIZE, 8, cudaMemcpyDeviceToHost));
  *out = v;
  return MEEPO_OK;
}

static meepo_status dirty_size(meepo_table* t, uint64_t* out) {
  if (!t->v.dirty) return fail(MEEPO_EINVAL, "incremental export needs MEEPO_FLAG_TRACK_DIRTY");
  unsigned long long v = 0;
  unsigned long long* cell = t->dstate->scratch64;
  const uint32_t words = (t->v.slots + 31) / 32;
  MEEPO_CUDA_TRY(cudaDeviceSynchronize());
  MEEPO_CUDA_TRY(cudaMemsetAsync(cell, 0, 8, nullptr));
  dirty_count_kernel<<<std::min<uint32_t>((words + 255) / 256, (uint32_t)t->num_sms * 8), 256>>>(t->v.dirty, words, cell);
  MEEPO_CUDA_TRY(cudaGetLastError());
  MEEPO_CUDA_TRY(cudaMemcpy(&v, cell, 8, cudaMemcpyDeviceToHost));
  *out = v;
  return MEEPO_OK;
}
static meepo_status dirty_clear(meepo_table* t, cudaStream_t stream) {
  MEEPO_CUDA_TRY(cudaMemsetAsync(t->v.dirty, 0, (size_t)((t->v.slots + 31) / 32) * 4, stream));
  return MEEPO_OK;
}

// sort keys of the two passes that order 64-bit keys: the low word of key[ord[i]] (ord == null: i), then the high word
__global__ void key_word_kernel(int high, const uint64_t* __restrict__ key, const uint32_t* __restrict__ ord, uint32_t n,
                                uint32_t* __restrict__ out, uint32_t* __restrict__ iota) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t k = key[ord ? ord[i] : i];
    out[i] = high ? (uint32_t)(k >> 32) : (uint32_t)k;
    if (iota) iota[i] = i;
  }
}
__global__ void permute_kernel(const uint32_t* __restrict__ ord, const uint64_t* __restrict__ k_in,
                               const uint32_t* __restrict__ s_in, uint32_t n, uint64_t* __restrict__ k_out,
                               uint32_t* __restrict__ s_out) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    k_out[j] = k_in[ord[j]];
    s_out[j] = s_in[ord[j]];
  }
}

// Live (key, slot) pairs sorted by key, left in the workspace (delta: the dirty ones only). Synchronous.
meepo_status sorted_live(meepo_table* t, uint64_t n, uint64_t** keys_sorted, uint32_t** slots_sorted,
                         cudaStream_t stream, bool delta = false) {
  const uint32_t m = t->v.slots;
  const size_t cbytes = compact_state_bytes(m);
  if (n && !radix_sort_supported(n, 32)) return fail(MEEPO_EINVAL, "export: more than 2^30 tuples in one call");
  const size_t tmp_bytes = n ? radix_sort_temp_bytes(n, 32) : 0;
  const size_t need = 2 * Workspace::pad(n * 8) + 6 * Workspace::pad(n * 4) + Workspace::pad(tmp_bytes) +
                      Workspace::pad(cbytes) + 4096;
  MEEPO_TRY(t->ws.reserve(need, stream));
  uint64_t* k_in = t->ws.take<uint64_t>(n);
  uint64_t* k_out = t->ws.take<uint64_t>(n);
  uint32_t* s_in = t->ws.take<uint32_t>(n);
  uint32_t* s_out = t->ws.take<uint32_t>(n);
  uint32_t* w_a = t->ws.take<uint32_t>(n);
  uint32_t* w_b = t->ws.take<uint32_t>(n);
  uint32_t* ord_a = t->ws.take<uint32_t>(n);
  uint32_t* ord_b = t->ws.take<uint32_t>(n);
  char* tmp = t->ws.take<char>(tmp_bytes);
  char* cstate = t->ws.take<char>(cbytes);
  MEEPO_CUDA_TRY(cudaMemsetAsync(cstate, 0, cbytes, stream));
  live_compact_kernel<<<compact_tiles(m), kCompactThreads, 0, stream>>>(t->v, m, k_in, s_in,
                                                                        compact_carve(cstate, t->err_word + kErrLookback),
                                                                        delta ? 1 : 0);
  MEEPO_CUDA_TRY(cudaGetLastError());
  if (n) {
    const int g = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)t->num_sms * 8));
    key_word_kernel<<<g, 256, 0, stream>>>(0, k_in, nullptr, (uint32_t)n, w_a, ord_a);
    MEEPO_TRY(radix_sort_pairs(t, tmp, w_a, w_b, ord_a, ord_b, (uint32_t)n, 32, stream));
    key_word_kernel<<<g, 256, 0, stream>>>(1, k_in, ord_b, (uint32_t)n, w_a, nullptr);
    MEEPO_TRY(radix_sort_pairs(t, tmp, w_a, w_b, ord_b, ord_a, (uint32_t)n, 32, stream));
    permute_kernel<<<g, 256, 0, stream>>>(ord_a, k_in, s_in, (uint32_t)n, k_out, s_out);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  *keys_sorted = k_out;
  *slots_sorted = s_out;
  return MEEPO_OK;
}

static int warp_grid(const meepo_table* t, uint64_t rows) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((rows + 7) / 8, (uint64_t)t->num_sms * 8));
}

// Shared by meepo_import_buffers and meepo_spill_readmit (device pointers).
meepo_status import_device(meepo_table* t, const uint64_t* keys, const void* rows, const void* state,
                           const uint64_t* scores, const uint32_t* steps, uint64_t n, uint8_t* status_out,
                           cudaStream_t stream, uint32_t* slot_buf, uint32_t* new_slots) {
  if (n == 0) return MEEPO_OK;
  t->cache_valid = false;
  t->slot_gen++;
  NewList nl{new_slots};
  const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)t->num_sms * 8);
  MEEPO_TRY(import_probe_launch(t, keys, n, slot_buf, status_out, nl, stream));
  MEEPO_TRY(publish_slots(t, nl.slots, n, stream));
  arena_scatter_kernel<<<warp_grid(t, n), 256, 0, stream>>>(t->v.rows, slot_buf, (uint32_t)n, t->v.cpr,
                                                            reinterpret_cast<const uint4*>(rows),
                                                            make_uint4(0, 0, 0, 0), 0);
  if (t->v.scpr) {
    const uint32_t a = float_bits(t->v.opt == MEEPO_ADAGRAD || t->v.opt == MEEPO_ADAGRAD_ROWWISE ? t->v.init_accum : 0.0f);
    const uint4 fill = t->v.opt == MEEPO_ADAGRAD_ROWWISE ? make_uint4(a, 0, 0, 0) : make_uint4(a, a, a, a);
    arena_scatter_kernel<<<warp_grid(t, n), 256, 0, stream>>>(t->v.state, slot_buf, (uint32_t)n, t->v.scpr,
                                                              reinterpret_cast<const uint4*>(state), fill, state == nullptr);
  }
  import_meta_kernel<<<grid, 256, 0, stream>>>(t->v, slot_buf, (uint32_t)n, scores, steps);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

}  // namespace meepo

using namespace meepo;

namespace {
struct FileHeader {  // 56 bytes
  char magic[8];
  uint32_t version, dim, dtype, opt;
  uint64_t n, row_bytes, state_bytes, epoch;
};
static_assert(sizeof(FileHeader) == 56, "header layout");
}  // namespace

extern "C" {

static meepo_status export_buffers_impl(meepo_table* t, uint64_t* keys, void* rows, void* state, uint64_t* scores,
                                        uint32_t* steps, uint64_t max_n, uint64_t* n_out, bool delta) {
  if (!t || !n_out) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  VerbScope vs(t, nullptr);
  MEEPO_TRY(vs.rc);
  uint64_t n = 0;
  MEEPO_TRY(delta ? dirty_size(t, &n) : live_size(t, &n));
  *n_out = n;
  if (!keys) return MEEPO_OK;
  if (max_n < n) return fail(MEEPO_EINVAL, "export buffers too small");
  if (n == 0) return MEEPO_OK;
  cudaStream_t stream = nullptr;
  uint64_t* ks;
  uint32_t* ss;
  MEEPO_TRY(sorted_live(t, n, &ks, &ss, stream, delta));
  MEEPO_CUDA_TRY(cudaMemcpyAsync(keys, ks, n * 8, cudaMemcpyDeviceToDevice, stream));
  if (rows)
    arena_gather_kernel<<<warp_grid(t, n), 256, 0, stream>>>(t->v.rows, ss, (uint32_t)n, t->v.cpr,
                                                             reinterpret_cast<uint4*>(rows));
  if (state && t->v.scpr)
    arena_gather_kernel<<<warp_grid(t, n), 256, 0, stream>>>(t->v.state, ss, (uint32_t)n, t->v.scpr,
                                                             reinterpret_cast<uint4*>(state));
  if (scores || steps) export_meta_kernel<<<warp_grid(t, n), 256, 0, stream>>>(t->v, ss, (uint32_t)n, scores, steps);
  MEEPO_CUDA_TRY(cudaGetLastError());
  if (delta) MEEPO_TRY(dirty_clear(t, stream));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_export_buffers(meepo_table* t, uint64_t* keys, void* rows, void* state,
                                            uint64_t* scores, uint32_t* steps, uint64_t max_n, uint64_t* n_out) {
  return export_buffers_impl(t, keys, rows, state, scores, steps, max_n, n_out, false);
}
MEEPO_API meepo_status meepo_export_delta_buffers(meepo_table* t, uint64_t* keys, void* rows, void* state,
                                                  uint64_t* scores, uint32_t* steps, uint64_t max_n,
                                                  uint64_t* n_out) {
  return export_buffers_impl(t, keys, rows, state, scores, steps, max_n, n_out, true);
}

MEEPO_API meepo_status meepo_import_buffers(meepo_table* t, const uint64_t* keys, const void* rows,
                                            const void* state, const uint64_t* scores, const uint32_t* steps,
                                            uint64_t n, uint8_t* status_out) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!keys || !rows)) return fail(MEEPO_EINVAL, "null buffer");
  if (n == 0) return MEEPO_OK;
  DeviceGuard guard(t->device);
  cudaStream_t stream = nullptr;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  MEEPO_TRY(t->ws.reserve(2 * Workspace::pad(n * 4) + 1024, stream));
  uint32_t* slot_buf = t->ws.take<uint32_t>(n);
  uint32_t* new_slots = t->ws.take<uint32_t>(n);
  MEEPO_TRY(import_device(t, keys, rows, state, scores, steps, n, status_out, stream, slot_buf, new_slots));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  return MEEPO_OK;
}

static meepo_status export_file_impl(meepo_table* t, const char* path, bool delta) {
  if (!t || !path) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  VerbScope vs(t, nullptr);
  MEEPO_TRY(vs.rc);
  uint64_t n = 0;
  MEEPO_TRY(delta ? dirty_size(t, &n) : live_size(t, &n));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(MEEPO_EIO, std::string("cannot open ") + path);
  std::unique_ptr<FILE, int (*)(FILE*)> closer(f, fclose);
  FileHeader h{};
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
  cudaStream_t stream = nullptr;
  uint64_t* ks;
  uint32_t* ss;
  MEEPO_TRY(sorted_live(t, n, &ks, &ss, stream, delta));  // stays valid: nothing below touches the workspace
  // one device staging buffer + one pinned bounce buffer, sections written in file order
  const uint64_t chunk = 1u << 16;
  const size_t widest = std::max<size_t>({(size_t)t->row_bytes, (size_t)t->state_bytes, 8});
  char *d_stage = nullptr, *h_stage = nullptr;
  MEEPO_CUDA_TRY(cudaMalloc(&d_stage, chunk * widest));
  if (cudaHostAlloc(&h_stage, chunk * widest, cudaHostAllocDefault) != cudaSuccess) {
    cudaFree(d_stage);
    return fail(MEEPO_ENOMEM, "cudaHostAlloc(export bounce)");
  }
  meepo_status rc = MEEPO_OK;
  auto section = [&](int what, size_t width) {
    for (uint64_t lo = 0; lo < n && rc == MEEPO_OK; lo += chunk) {
      const uint64_t m = std::min(chunk, n - lo);
      const void* src = d_stage;
      switch (what) {
        case 0: src = ks + lo; break;
        case 1:
          arena_gather_kernel<<<warp_grid(t, m), 256, 0, stream>>>(t->v.rows, ss + lo, (uint32_t)m, t->v.cpr,
                                                                   reinterpret_cast<uint4*>(d_stage));
          break;
        case 2:
          arena_gather_kernel<<<warp_grid(t, m), 256, 0, stream>>>(t->v.state, ss + lo, (uint32_t)m, t->v.scpr,
                                                                   reinterpret_cast<uint4*>(d_stage));
          break;
        case 3:
          export_meta_kernel<<<warp_grid(t, m), 256, 0, stream>>>(t->v, ss + lo, (uint32_t)m,
                                                                  reinterpret_cast<uint64_t*>(d_stage), nullptr);
          break;
        default:
          export_meta_kernel<<<warp_grid(t, m), 256, 0, stream>>>(t->v, ss + lo, (uint32_t)m, nullptr,
                                                                  reinterpret_cast<uint32_t*>(d_stage));
      }
      if (cudaMemcpyAsync(h_stage, src, m * width, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
          cudaStreamSynchronize(stream) != cudaSuccess)
        rc = fail(MEEPO_ECUDA, "export copy failed");
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
  if (rc == MEEPO_OK && delta) {  // the tuples are on disk: forget their marks
    rc = dirty_clear(t, stream);
    if (rc == MEEPO_OK && cudaStreamSynchronize(stream) != cudaSuccess) rc = fail(MEEPO_ECUDA, "dirty clear failed");
  }
  return rc;
}

MEEPO_API meepo_status meepo_export(meepo_table* t, const char* path) { return export_file_impl(t, path, false); }
MEEPO_API meepo_status meepo_export_delta(meepo_table* t, const char* path) { return export_file_impl(t, path, true); }

MEEPO_API meepo_status meepo_import(meepo_table* t, const char* path) {
  if (!t || !path) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  VerbScope vs(t, nullptr);
  MEEPO_TRY(vs.rc);
  FILE* f = fopen(path, "rb");
  if (!f) return fail(MEEPO_EIO, std::string("cannot open ") + path);
  std::unique_ptr<FILE, int (*)(FILE*)> closer(f, fclose);
  FileHeader h{};
  if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "MEEPOTB1", 8) != 0 || h.version != 1)
    return fail(MEEPO_EIO, "bad header");
  if (h.dim != t->cfg.dim || (int32_t)h.dtype != t->cfg.dtype || h.row_bytes != t->row_bytes ||
      h.state_bytes != t->state_bytes || (int32_t)h.opt != t->cfg.opt)
    return fail(MEEPO_EINVAL, "file does not match table configuration");
  const uint64_t n = h.n, R = t->row_bytes, S = t->state_bytes;
  {  // the header's n drives every offset below: it must agree with the file size, and the tuples must fit
    const uint64_t tuple_file = 8 + R + S + 8 + 4;
    if (fseeko(f, 0, SEEK_END) != 0) return fail(MEEPO_EIO, "seek failed");
    const off_t fsz = ftello(f);
    if (fsz < 0 || n > ((uint64_t)fsz - sizeof h) / tuple_file || (uint64_t)fsz != sizeof h + n * tuple_file)
      return fail(MEEPO_EIO, "file size does not match the tuple count in the header");
    if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "file holds more tuples than a table can");
  }
  if (h.epoch > t->epoch) t->epoch = h.epoch;
  if (n == 0) return MEEPO_OK;
  const uint64_t off_keys = sizeof h, off_rows = off_keys + n * 8, off_state = off_rows + n * R,
                 off_scores = off_state + n * S, off_steps = off_scores + n * 8;
  const uint64_t chunk = 1u << 16;
  const size_t tuple = 8 + R + S + 8 + 4;
  char *d_stage = nullptr, *h_stage = nullptr;
  MEEPO_CUDA_TRY(cudaMalloc(&d_stage, chunk * (tuple + 1) + 1024));
  if (cudaHostAlloc(&h_stage, chunk * (tuple + 1) + 1024, cudaHostAllocDefault) != cudaSuccess) {
    cudaFree(d_stage);
    return fail(MEEPO_ENOMEM, "cudaHostAlloc(import bounce)");
  }
  meepo_status rc = MEEPO_OK;
  uint64_t rejected = 0;  // tuples that found no slot (or carried a reserved key): reported, never dropped silently
  for (uint64_t lo = 0; lo < n && rc == MEEPO_OK; lo += chunk) {
    const uint64_t m = std::min(chunk, n - lo);
    // staging layout (each section 16-byte aligned because chunk is a multiple of 16)
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
    rc = meepo_import_buffers(t, reinterpret_cast<uint64_t*>(d_stage + o_k), d_stage + o_r, S ? d_stage + o_s : nullptr,
                              reinterpret_cast<uint64_t*>(d_stage + o_c), reinterpret_cast<uint32_t*>(d_stage + o_t),
                              m, reinterpret_cast<uint8_t*>(d_stage + o_t + chunk * 4));
    if (rc != MEEPO_OK) break;
    uint8_t* hst = reinterpret_cast<uint8_t*>(h_stage + o_t + chunk * 4);
    if (cudaMemcpy(hst, d_stage + o_t + chunk * 4, m, cudaMemcpyDeviceToHost) != cudaSuccess) {
      rc = fail(MEEPO_ECUDA, "import copy failed");
      break;
    }
    for (uint64_t i = 0; i < m; i++) rejected += hst[i] == MEEPO_KEY_FULL || hst[i] == MEEPO_KEY_INVALID;
  }
  cudaFree(d_stage);
  cudaFreeHost(h_stage);
  if (rc == MEEPO_OK && rejected)
    rc = fail(MEEPO_ENOMEM, std::to_string(rejected) + " of " + std::to_string(n) +
                                " tuples did not fit into the table (full) or carried a reserved key; the rest were imported");
  return rc;
}

}  // extern "C"
