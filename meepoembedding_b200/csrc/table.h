// table.h — host-side object behind the opaque meepo_table handle (CUDA build).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace meepo {

void set_error(const std::string& m);
meepo_status fail(meepo_status s, const std::string& m);

#define MEEPO_CUDA_TRY(expr)                                                                     \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::meepo::fail(_e == cudaErrorMemoryAllocation ? MEEPO_ENOMEM : MEEPO_ECUDA,         \
                           std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
  } while (0)

#define MEEPO_TRY(expr)                \
  do {                                 \
    meepo_status _s = (expr);          \
    if (_s != MEEPO_OK) return _s;     \
  } while (0)

// Bump allocator over one growable device buffer; reset at the start of every verb.
struct Workspace {
  char* base = nullptr;
  size_t bytes = 0, used = 0;
  meepo_status reserve(size_t need, cudaStream_t stream);
  void reset() { used = 0; }
  template <typename T>
  T* take(size_t count) {
    size_t off = (used + 255) & ~size_t(255);
    used = off + count * sizeof(T);
    return reinterpret_cast<T*>(base + off);
  }
  static size_t pad(size_t b) { return (b + 255) & ~size_t(255); }
};

// Small always-resident device block.
struct DeviceState {
  unsigned long long counters[16];
  uint32_t sched[2];  // probe kernels: dynamic tile tickets of the tail + finished warps (self-resetting)
  uint32_t unused0[3];
  uint32_t evict_count;
  uint32_t pad[2];
  unsigned long long scratch64[8];
  unsigned long long hist[256];  // evict: radix-select histogram
};

struct Profiler;
struct PeerState;  // peer.cu: exchange window + peer mappings of the sharded verbs

// (key, slot) of every element of the last find_or_insert / lookup batch. apply_gradients on the
// same keys (the training loop) reuses the slots instead of probing again: one HBM line per key
// saved. Entries are validated per element (keys[i] == cache.keys[i]); any mutation other than
// find_or_insert / lookup / apply_gradients invalidates the whole cache.
struct SlotCache {
  uint64_t* keys;
  uint32_t* slots;
};

// Slots claimed by the find_or_insert call in flight, one cell per batch element: the claimed slot
// if this element's thread won the CAS, kNil otherwise. No shared counter: a batch-wide
// atomicAdd-with-return per warp serialises in L2 (measured: +0.1 ms per 1M-key batch with 5% new keys).
struct NewList {
  uint32_t* slots;
};

}  // namespace meepo

struct meepo_table {
  meepo_config cfg{};
  meepo::TableView v{};
  int device = 0;
  int num_sms = 0;
  uint32_t row_bytes = 0, state_bytes = 0;
  meepo::DeviceState* dstate = nullptr;
  meepo::Workspace ws;
  meepo::NewList cur_new{nullptr};
  uint64_t cur_new_off = 0;
  meepo::SlotCache cache{nullptr, nullptr};
  uint64_t cache_cap = 0, cache_n = 0, cache_off = 0;
  bool cache_valid = false, cache_enabled = true;
  // Sticky device-side errors, in mapped pinned host memory so that every verb can test them without a
  // synchronisation: [0] a look-back (radix sort / compaction) gave up waiting for a tile, [1] a peer missed a
  // barrier, [2] a (sender, owner) lane of the exchange window overflowed. Kernels only ever store 1.
  volatile uint32_t* err_host = nullptr;
  uint32_t* err_word = nullptr;  // the same words as the device sees them
  // Verbs of one table may be issued on different streams: each verb's stream first waits for the event the
  // previous verb recorded (workspace, slot cache and scratch counters are shared by all verbs of a table).
  float grid_scale = 1.0f;  // < 1 while two pipelines of this table share the GPU (sharded backward pass)
  cudaEvent_t order_ev = nullptr;
  cudaStream_t last_stream = nullptr;
  bool order_valid = false;
  uint64_t slot_gen = 0;  // bumped whenever slots may change owners outside the sharded verbs (evict, import, ...)
  uint64_t epoch = 0;
  struct meepo::Profiler* prof = nullptr;  // per-kernel event timing, off unless enabled
  // host-buffer front end (pinned staging + private streams), created lazily
  struct HostPipe* pipe = nullptr;
  struct meepo::PeerState* peer = nullptr;
  // host tier (meepo.h "Host tier"; evict.cu): the device-side view is v.tier
  char* spill_ring = nullptr;    // mapped pinned host memory: the ring of slabs
  uint64_t spill_cap_tuples = 0; // slabs
  uint64_t tier_head = 0;        // tuples appended so far: the next one goes to slab tier_head % slabs
  uint64_t tier_nonempty_ub = 0; // upper bound of the index cells that are not EMPTY (live + tombstones)
  char* tier_stage = nullptr;    // device staging buffer of the latest eviction's tuples
  uint64_t tier_stage_cap = 0;   // ... in tuples
  cudaStream_t tier_stream = nullptr;  // drains the staging buffer to the ring
  cudaEvent_t tier_staged = nullptr, tier_drained = nullptr;
  bool tier_draining = false;
  char* pool_scaled = nullptr;   // pooled backward, MEAN: the bag gradients divided by the bag lengths
  size_t pool_scaled_bytes = 0;

  uint64_t tuple_bytes() const { return 24 + (uint64_t)row_bytes + state_bytes; }
};

namespace meepo {
// RAII device guard: every verb runs on the table's device.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
    if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    want = dev;
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != want) cudaSetDevice(prev);
  }
  int want;
};

// Every verb: fail fast on a sticky device-side error, order after the previous verb of this table.
meepo_status verb_begin(meepo_table* t, cudaStream_t stream);
void verb_end(meepo_table* t, cudaStream_t stream);
struct VerbScope {
  meepo_table* t;
  cudaStream_t s;
  meepo_status rc;
  VerbScope(meepo_table* t_, cudaStream_t s_) : t(t_), s(s_), rc(verb_begin(t_, s_)) {}
  ~VerbScope() {
    if (rc == MEEPO_OK) verb_end(t, s);
  }
};
enum : int { kErrLookback = 0, kErrPeerTimeout = 1, kErrPeerOverflow = 2, kErrWords = 4 };
meepo_status sticky_error(meepo_table* t);
meepo_status probe_histogram(meepo_table* t, uint64_t* out4);
struct CompactState;
CompactState compact_carve(char* p, uint32_t* error);

// kernels' host launchers (defined across the .cu files)
meepo_status launch_probe_gather(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                 uint8_t* status_out, bool insert, cudaStream_t stream);
meepo_status probe_gather_begin(meepo_table* t, uint64_t n_total, bool insert, cudaStream_t stream,
                                size_t extra_bytes = 0);
meepo_status probe_gather_chunk(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                uint8_t* status_out, bool insert, cudaStream_t stream);
meepo_status probe_gather_end(meepo_table* t, uint64_t n_total, bool insert, cudaStream_t stream);
// grads_ready (optional): event the reduce kernels wait for, so a caller can overlap the copy of
// the gradients with the probe / sort / segment passes that only need the keys.
// bag_offsets (optional, pooled backward): grads holds one row per bag; occurrence i reads the row of its bag
meepo_status launch_apply_gradients(meepo_table* t, const uint64_t* keys, const void* grads, uint64_t n,
                                    cudaStream_t stream, cudaEvent_t grads_ready = nullptr,
                                    const uint32_t* bag_offsets = nullptr, uint32_t n_bags = 0);
// Device-side counters of one sort + segment + reduce pipeline: every pipeline in flight has its own (the owner
// side of a sharded backward pass runs on another stream than the sender side).
struct SegScratch {
  uint32_t num_segments, num_long, num_leaves, pad;
  uint32_t chunk_first[8];  // first segment of chunk c (kNil: the chunk is empty); only with a chunked sort key
};
struct SegRange {  // which segments a reduce pass takes, by sort key, and how a sort key maps to its output row
  uint32_t key_lo, key_hi, key_mask;
  int chunk = -1;  // >= 0: the segments of this chunk only (SegScratch::chunk_first), instead of a scan of all
};
// Scratch of one sort + segment + reduce pipeline (update.cu); carved out of the table workspace.
struct SegWork {
  uint32_t *sk_in, *sk_out, *sv_in, *sv_out, *seg_start;
  char *cub_tmp, *long_seg, *cstate;  // cstate: ticket + per-tile look-back words of the segment-head compaction
  uint2* leaf_desc;
  uint4* seg_desc;
  float* partial;
  SegScratch* sc;
  size_t cub_bytes, max_long, max_leaves, cstate_bytes;
  uint32_t n, ntiles;
  int end_bit;
  static size_t bytes(uint64_t n, uint32_t dim, int end_bit);
  void take(Workspace& ws, uint64_t n, uint32_t dim, int end_bit);
};
constexpr int kReduceStoreOnly = 15;  // == kStoreOnly (optimizer.cuh): not a meepo_opt
int bits_for(uint32_t max_value);
// radix_sort.cu: stable LSD sort of (u32 key, u32 value) pairs on key bits [0, end_bit)
bool radix_sort_supported(uint64_t n, int end_bit);
size_t radix_sort_temp_bytes(uint64_t n, int end_bit);
meepo_status radix_sort_pairs(meepo_table* t, char* temp, const uint32_t* k_in, uint32_t* k_out, const uint32_t* v_in,
                              uint32_t* v_out, uint32_t n, int end_bit, cudaStream_t stream,
                              const uint32_t* n_dev = nullptr, uint32_t expect_n = 0);
// chunk_shift >= 0: the sort key carries a chunk index above bit chunk_shift; the first segment of every chunk is
// recorded (SegScratch::chunk_first) so that seg_reduce can take one chunk's segments without visiting the others
meepo_status seg_sort_heads(meepo_table* t, SegWork& w, uint32_t miss_key, bool count_updates, const uint32_t* n_dev,
                            cudaStream_t stream, const char* const* names, uint32_t expect_n = 0,
                            int chunk_shift = -1);
meepo_status seg_reduce(meepo_table* t, SegWork& w, const void* grads, int mode, const SegRange& r, void* reduce_out,
                        void* const* reduce_rows, cudaStream_t stream, cudaEvent_t grads_ready,
                        const char* const* names, bool again);
meepo_status run_segmented(meepo_table* t, SegWork& w, uint32_t limit, const void* grads, int mode,
                           void* reduce_out, cudaStream_t stream, cudaEvent_t grads_ready,
                           const char* const* names, void* const* reduce_rows = nullptr);
// shard.cu: batch-level dedup. unique_keys[U] (order unspecified), inverse[n] (kNil for invalid
// keys), *n_unique = U (device), occurrences[U] = duplicates per unique key (optional), grads_out[U] =
// fixed-shape sum of each key's gradient rows rounded to the table dtype (with grads only).
struct DedupOut {
  uint64_t* unique_keys;
  void* grads_out;
  uint32_t* inverse;
  uint64_t* n_unique;
  uint32_t* occurrences;
  void* const* grad_rows;  // optional: destination row of unique key u instead of grads_out[u]
  uint32_t* canon;         // optional: batch index of one (arbitrary but fixed) occurrence of unique key u
};
struct SegWork;
size_t dedup_bytes(const meepo_table* t, uint64_t n, bool with_grads);
meepo_status dedup_hash(meepo_table* t, const uint64_t* keys, uint64_t n, const DedupOut& o, bool with_grads,
                        SegWork& w, cudaStream_t stream, const uint32_t* skip = nullptr);
meepo_status dedup_reduce(meepo_table* t, SegWork& w, const void* grads, uint64_t n, const DedupOut& o,
                          cudaStream_t stream);
meepo_status dedup_run(meepo_table* t, const uint64_t* keys, const void* grads, uint64_t n, const DedupOut& o,
                       cudaStream_t stream);
int grid_for(const meepo_table* t, const void* kernel, int block, size_t smem, uint64_t blocks_needed);
void destroy_host_pipe(meepo_table* t);
void destroy_profiler(meepo_table* t);
void prof_add_host(meepo_table* t, const char* name, double ms);
void destroy_peer(meepo_table* t);
// lookup.cu: write the tags of the slots listed in slots[0..n) (kNil cells skipped) and fold their
// number into the size
meepo_status publish_slots(meepo_table* t, const uint32_t* slots, uint64_t n, cudaStream_t stream);
// evict.cu: the host tier
meepo_status tier_create(meepo_table* t);
void tier_destroy(meepo_table* t);
// io.cu
meepo_status live_size(meepo_table* t, uint64_t* out);
meepo_status import_probe_launch(meepo_table* t, const uint64_t* keys, uint64_t n, uint32_t* slot_out,
                                 uint8_t* status_out, NewList nl, cudaStream_t stream);
// Times the kernels launched inside its lifetime when profiling is on (profile.cu).
struct ProfScope {
  ProfScope(meepo_table* t, const char* name, cudaStream_t s);
  ~ProfScope();
  meepo_table* t_;
  const char* name_;
  cudaStream_t s_;
  cudaEvent_t a_ = nullptr;
};
}  // namespace meepo
