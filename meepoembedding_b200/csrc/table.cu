// table.cu — life cycle of the CUDA table: create / destroy / stats, workspace, error state.
// Implements the life-cycle block of include/meepo.h (DERIVED API; the upstream repository has
// no code to mirror — /root/reference/README.md:1-2).
#include "compact.cuh"
#include "table.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace meepo {

static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
meepo_status fail(meepo_status s, const std::string& m) {
  g_err = m;
  return s;
}
const char* last_error_cstr() { return g_err.c_str(); }

meepo_status Workspace::reserve(size_t need, cudaStream_t stream) {
  used = 0;
  if (need <= bytes) return MEEPO_OK;
  // growing is rare (first call at a new batch size): drain the stream, then swap buffers
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  if (base) MEEPO_CUDA_TRY(cudaFree(base));
  base = nullptr;
  bytes = 0;
  size_t want = need + need / 4;
  MEEPO_CUDA_TRY(cudaMalloc(&base, want));
  bytes = want;
  return MEEPO_OK;
}

CompactState compact_carve(char* p, uint32_t* error) {
  CompactState cs;
  cs.ticket = reinterpret_cast<uint32_t*>(p);
  cs.state = reinterpret_cast<unsigned long long*>(p + 8);
  cs.error = error;
  return cs;
}

meepo_status sticky_error(meepo_table* t) {
  if (!t->err_host) return MEEPO_OK;
  if (t->err_host[kErrLookback])
    return fail(MEEPO_ECUDA, "a look-back (radix sort / compaction) gave up waiting for a tile; the table may be corrupt");
  if (t->err_host[kErrPeerTimeout])
    return fail(MEEPO_ENCCL, "sharded verb: a peer did not reach the barrier (timeout)");
  if (t->err_host[kErrPeerOverflow])
    return fail(MEEPO_ENCCL, "sharded verb: more keys for one owner than region_keys; results are incomplete");
  return MEEPO_OK;
}

meepo_status verb_begin(meepo_table* t, cudaStream_t stream) {
  MEEPO_TRY(sticky_error(t));
  if (t->order_valid && stream != t->last_stream) MEEPO_CUDA_TRY(cudaStreamWaitEvent(stream, t->order_ev, 0));
  return MEEPO_OK;
}
void verb_end(meepo_table* t, cudaStream_t stream) {
  if (cudaEventRecord(t->order_ev, stream) == cudaSuccess) {
    t->last_stream = stream;
    t->order_valid = true;
  }
}

// meepo_stats: how far from home the live keys sit (one pass over the bucket array)
__global__ void __launch_bounds__(256) probe_hist_kernel(TableView t, unsigned long long* __restrict__ hist) {
  uint32_t c[4] = {0, 0, 0, 0};
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.slots; s += gridDim.x * blockDim.x) {
    const uint64_t key = *key_ptr(t, s);
    if (key == MEEPO_KEY_EMPTY) continue;
    const uint32_t b = s / kBucket, home = bucket_of(mix64(key), t.num_buckets);
    const uint32_t d = b >= home ? b - home : b + t.num_buckets - home;
    c[d < 3 ? d : 3]++;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    c[i] = __reduce_add_sync(0xFFFFFFFFu, c[i]);
    if ((threadIdx.x & 31u) == 0 && c[i]) atomicAdd(hist + i, (unsigned long long)c[i]);
  }
}
meepo_status probe_histogram(meepo_table* t, uint64_t* out4) {
  unsigned long long* d = t->dstate->hist;
  MEEPO_CUDA_TRY(cudaMemset(d, 0, 4 * 8));
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)t->v.slots + 255) / 256, (uint64_t)t->num_sms * 8));
  probe_hist_kernel<<<grid, 256>>>(t->v, d);
  MEEPO_CUDA_TRY(cudaGetLastError());
  MEEPO_CUDA_TRY(cudaMemcpy(out4, d, 4 * 8, cudaMemcpyDeviceToHost));
  return MEEPO_OK;
}

int grid_for(const meepo_table* t, const void* kernel, int block, size_t smem, uint64_t blocks_needed) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  uint64_t cap = (uint64_t)((double)t->num_sms * per_sm * t->grid_scale);
  if (cap < (uint64_t)t->num_sms) cap = t->num_sms;
  uint64_t g = blocks_needed < cap ? blocks_needed : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace meepo

using namespace meepo;

extern "C" {

MEEPO_API uint32_t meepo_abi_version(void) { return MEEPO_ABI_VERSION; }
MEEPO_API const char* meepo_backend(void) { return "cuda-sm_100a"; }
MEEPO_API const char* meepo_last_error(void) { return meepo::last_error_cstr(); }
MEEPO_API uint32_t meepo_owner(uint64_t key, uint32_t num_shards) {
  return (uint32_t)(((unsigned __int128)mix64(key ^ MEEPO_OWNER_SALT) * num_shards) >> 64);
}

MEEPO_API meepo_status meepo_destroy(meepo_table* t);

MEEPO_API meepo_status meepo_create(const meepo_config* cfg, meepo_table** out) {
  if (!cfg || !out) return fail(MEEPO_EINVAL, "null argument");
  if (cfg->dtype != MEEPO_F32 && cfg->dtype != MEEPO_BF16) return fail(MEEPO_EINVAL, "bad dtype");
  if (cfg->opt < MEEPO_SGD || cfg->opt > MEEPO_ADAGRAD_ROWWISE) return fail(MEEPO_EINVAL, "bad optimizer");
  const uint32_t esz = cfg->dtype == MEEPO_F32 ? 4 : 2;
  if (cfg->dim == 0 || ((uint64_t)cfg->dim * esz) % 16 != 0)
    return fail(MEEPO_EINVAL, "row bytes must be a positive multiple of 16");
  if (cfg->capacity == 0 || cfg->capacity > 0xFFFFFFC0ull) return fail(MEEPO_EINVAL, "bad capacity");
  int ndev = 0;
  MEEPO_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(MEEPO_EINVAL, "bad device ordinal");
  DeviceGuard guard(cfg->device);
  if (!guard.ok) return fail(MEEPO_ECUDA, "cudaSetDevice failed");

  meepo_table* t = new meepo_table();
  t->cfg = *cfg;
  t->cache_enabled = getenv("MEEPO_NO_SLOT_CACHE") == nullptr;
  t->device = cfg->device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) {
    delete t;
    return fail(MEEPO_ECUDA, "cudaGetDeviceProperties failed");
  }
  t->num_sms = prop.multiProcessorCount;
  const uint64_t slots = (cfg->capacity + kBucket - 1) / kBucket * kBucket;
  t->row_bytes = cfg->dim * esz;
  t->state_bytes = cfg->opt == MEEPO_SGD              ? 0
                   : cfg->opt == MEEPO_ADAGRAD        ? cfg->dim * 4
                   : cfg->opt == MEEPO_ADAGRAD_ROWWISE ? 16  // one accumulator + padding to a 16-byte chunk
                                                       : cfg->dim * 8;
  TableView& v = t->v;
  v.slots = (uint32_t)slots;
  v.num_buckets = (uint32_t)(slots / kBucket);
  v.cpr = t->row_bytes / 16;
  v.scpr = t->state_bytes / 16;
  v.dim = cfg->dim;
  v.dtype = cfg->dtype;
  v.opt = cfg->opt;
  v.lr = cfg->lr;
  v.eps = cfg->eps;
  v.beta1 = cfg->beta1;
  v.beta2 = cfg->beta2;
  v.init_accum = cfg->init_accum;
  v.init_scale = cfg->init_scale;
  v.init_seed = cfg->init_seed;
  v.epoch = 0;

  auto bail = [&](cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    meepo_destroy(t);
    return fail(e == cudaErrorMemoryAllocation ? MEEPO_ENOMEM : MEEPO_ECUDA, m);
  };
  cudaError_t e;
  const size_t bucket_bytes = (size_t)v.num_buckets * sizeof(BucketLine);
  if ((e = cudaMalloc(&v.buckets, bucket_bytes)) != cudaSuccess) return bail(e, "cudaMalloc(buckets)");
  if ((e = cudaMalloc(&v.rows, slots * (size_t)t->row_bytes)) != cudaSuccess) return bail(e, "cudaMalloc(rows)");
  if (t->state_bytes &&
      (e = cudaMalloc(&v.state, slots * (size_t)t->state_bytes)) != cudaSuccess)
    return bail(e, "cudaMalloc(state)");
  if ((cfg->flags & MEEPO_FLAG_TRACK_SCORES) && (e = cudaMalloc(&v.scores, slots * 8)) != cudaSuccess)
    return bail(e, "cudaMalloc(scores)");
  if (cfg->opt == MEEPO_ADAM && (e = cudaMalloc(&v.steps, slots * 4)) != cudaSuccess)
    return bail(e, "cudaMalloc(steps)");
  if ((cfg->flags & MEEPO_FLAG_TRACK_DIRTY) && ((e = cudaMalloc(&v.dirty, (slots + 31) / 32 * 4)) != cudaSuccess ||
                                                (e = cudaMemset(v.dirty, 0, (slots + 31) / 32 * 4)) != cudaSuccess))
    return bail(e, "cudaMalloc(dirty bits)");
  if ((e = cudaMalloc(&t->dstate, sizeof(DeviceState))) != cudaSuccess) return bail(e, "cudaMalloc(dstate)");
  v.counters = t->dstate->counters;
  // keys = EMPTY (all ones), then the 16-byte headers (tags + metadata) = 0
  if ((e = cudaMemset(v.buckets, 0xFF, bucket_bytes)) != cudaSuccess) return bail(e, "memset");
  if ((e = cudaMemset2D(v.buckets, sizeof(BucketLine), 0, 16, v.num_buckets)) != cudaSuccess) return bail(e, "memset2D");
  if (v.scores && (e = cudaMemset(v.scores, 0, slots * 8)) != cudaSuccess) return bail(e, "memset");
  if (v.steps && (e = cudaMemset(v.steps, 0, slots * 4)) != cudaSuccess) return bail(e, "memset");
  if ((e = cudaMemset(t->dstate, 0, sizeof(DeviceState))) != cudaSuccess) return bail(e, "memset");
  {
    void* eh = nullptr;
    if ((e = cudaHostAlloc(&eh, kErrWords * 4, cudaHostAllocMapped | cudaHostAllocPortable)) != cudaSuccess)
      return bail(e, "cudaHostAlloc(error words)");
    memset(eh, 0, kErrWords * 4);
    t->err_host = reinterpret_cast<volatile uint32_t*>(eh);
    void* ed = nullptr;
    if ((e = cudaHostGetDevicePointer(&ed, eh, 0)) != cudaSuccess) return bail(e, "cudaHostGetDevicePointer");
    t->err_word = reinterpret_cast<uint32_t*>(ed);
    if ((e = cudaEventCreateWithFlags(&t->order_ev, cudaEventDisableTiming)) != cudaSuccess)
      return bail(e, "cudaEventCreate");
  }
  if (cfg->host_spill_bytes) {
    const meepo_status rc = tier_create(t);
    if (rc != MEEPO_OK) {
      const std::string m = meepo_last_error();
      meepo_destroy(t);
      return fail(rc, m);
    }
  }
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail(e, "cudaDeviceSynchronize");
  *out = t;
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_destroy(meepo_table* t) {
  if (!t) return MEEPO_OK;
  DeviceGuard guard(t->device);
  cudaDeviceSynchronize();
  destroy_peer(t);
  destroy_host_pipe(t);
  destroy_profiler(t);
  cudaFree(t->v.buckets);
  cudaFree(t->v.rows);
  cudaFree(t->v.state);
  cudaFree(t->v.scores);
  cudaFree(t->v.steps);
  cudaFree(t->v.dirty);
  cudaFree(t->dstate);
  cudaFree(t->ws.base);
  cudaFree(t->cache.keys);
  cudaFree(t->cache.slots);
  cudaFree(t->pool_scaled);
  tier_destroy(t);
  if (t->err_host) cudaFreeHost(const_cast<uint32_t*>(t->err_host));
  if (t->order_ev) cudaEventDestroy(t->order_ev);
  delete t;
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_stats(meepo_table* t, meepo_stats_t* out) {
  if (!t || !out) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  unsigned long long c[16];
  MEEPO_CUDA_TRY(cudaDeviceSynchronize());
  MEEPO_CUDA_TRY(cudaMemcpy(c, t->dstate->counters, sizeof c, cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof *out);
  out->capacity = t->v.slots;
  out->size = c[C_SIZE];
  out->inserts = c[C_INSERTS];
  out->hits = c[C_HITS];
  out->misses = c[C_MISSES];
  out->full = c[C_FULL];
  out->evictions = c[C_EVICTIONS];
  out->updates = c[C_UPDATES];
  out->grad_dropped = c[C_DROPPED];
  out->overflow_buckets = c[C_OVERFLOW];
  out->peer_keys_received = c[C_PEER_KEYS];
  out->peer_grads_received = c[C_PEER_GRADS];
  out->spill_keys = c[C_TIER_LIVE];
  out->spill_bytes = c[C_TIER_LIVE] * t->tuple_bytes();
  out->promotions = c[C_PROMOTIONS];
  out->tier_hits = c[C_TIER_HITS];
  out->epoch = t->epoch;
  out->row_bytes = t->row_bytes;
  out->state_bytes = t->state_bytes;
  MEEPO_TRY(probe_histogram(t, out->probe_hist));
  return sticky_error(t);
}

}  // extern "C"
