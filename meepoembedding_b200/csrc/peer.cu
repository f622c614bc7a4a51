// peer.cu — the key-hash-sharded verbs fused with their exchange over NVLink peer memory
// (SURVEY rows a4 / e; include/meepo.h "sharded verbs"). One process per GPU; every rank owns the
// keys with owner(key, world) == rank and exposes ONE exchange window (cudaIpc) that its peers
// store into directly. No NCCL call, no host synchronisation and no staging copy on the data path:
//
//   find_or_insert / lookup (requester r, owner o)
//     r: dedup the batch                      -> unique keys, inverse, occurrences, one canonical occurrence per key
//     r: push           (key, occurrences, canonical batch index) -> o.recv_*[r][p]. A CTA partitions its 2048
//                       keys by (owner, chunk) in shared memory and stores each run contiguously: 16 B per
//                       unique key over NVLink in 1-2 KB pieces
//     barrier (delivers the per-pair counts and where r's output tensor lives)
//     o: owner_probe_gather  probe the LOCAL table (same tile body as the single-table kernel) and store each
//                       row over NVLink STRAIGHT INTO r's OUTPUT TENSOR at the key's canonical batch index
//                       (when the caller's rows_out lies in the window's output area, meepo_peer_output);
//                       otherwise into r's return region. Tags of the slots claimed in this call are published
//                       afterwards, locally
//     barrier
//     r: finish         status bytes, and rows_out[i] = rows_out[canonical(i)] for the duplicate occurrences
//                       only (a local copy; with uniform keys ~2% of the batch) — or, for an output tensor
//                       outside the window, the full expansion rows_out[i] = ret_rows[loc[inverse[i]]]
//   apply_gradients, the unique keys split into K chunks by a hash bit field (K = 1..3, MEEPO_PEER_CHUNKS)
//     r: dedup (or reuse the forward pass's), positions at the owners: chunk c of a (sender, owner) lane is the
//        contiguous run [cnt[c-1], cnt[c])
//     r: ONE stable sort by (chunk, unique id) + segment heads for the whole batch
//     for c in 0..K-1:
//       r (stream S): fixed-shape pre-reduction of chunk c's duplicate gradients, rounded to the table dtype,
//                     whose store IS the exchange: each summed row goes straight to o.recv_grads[r][p] — NVLink-bound
//       barrier c     (delivers the cumulative counts)
//       o (stream T): slot of every entry of chunk c, stable sort by slot (entries enumerate rank-major, so each
//                     key's partial sums stay in rank order), segment heads, fused reduce + optimizer — HBM-bound,
//                     and it runs UNDER the NVLink-bound pre-reduction of chunk c+1 (the sender kernels are launched
//                     with a fraction of the SMs' CTA slots, the owner kernels with the rest)
//     barrier
//
// A backward pass over the batch of the preceding forward pass (the training loop) reuses it: the
// sender skips the dedup and the key push (positions and owners are remembered, the keys are still in
// the owners' windows), the owner takes each entry's slot from the forward pass instead of probing.
// Whether the batch is the same is decided on the device (a compare kernel sets a flag that the
// skippable kernels read), per sender, and an owner whose table changed in between probes anyway.
//
// Only bulk stores and the barrier flags cross NVLink: every table mutation (CAS insert, tag
// publish, optimizer step) is done by the owner on its own HBM. Every count that depends on the
// data (unique keys, keys per owner, received entries) stays on the device: grids are persistent
// and read their bounds from device memory, so a verb is one uninterrupted stream of launches.
#include <unistd.h>

#include <cstring>

#include "probe_gather.cuh"

namespace meepo {

constexpr uint32_t kMaxPeers = MEEPO_MAX_PEERS;
constexpr uint32_t kMaxChunks = 3;            // 2 bits above the unique id keep the sender's sort at 3 passes for 4M keys
constexpr uint32_t kMaxBins = kMaxPeers * 4;  // (owner, chunk) bins of the push; 32 = one warp
constexpr uint32_t kBlobMagic = 0x4D50454Cu;  // "MPEL"
constexpr size_t kWindowHeader = 4096;
constexpr unsigned long long kNoOutput = ~0ull;

// A rank's exchange window as seen through a (peer or local) mapping.
struct PeerWindow {
  unsigned long long* flags;         // [kMaxPeers] barrier sequence number last signalled by each source
  uint32_t* recv_cnt;                // [kMaxPeers] entries source s has delivered so far in the current verb
  uint32_t* recv_reuse;              // [kMaxPeers] source s re-sent the entries of its last forward pass (same positions)
  unsigned long long* recv_out_off;  // [kMaxPeers] byte offset of requester s's output tensor inside s's output area,
                                     //             kNoOutput = use s's return region (written by s into the owner's window)
  uint64_t* recv_keys;               // [world][region]      keys pushed by source s
  uint32_t* recv_occ;                // [world][region]      batch occurrences behind each pushed key
  uint32_t* recv_dest;               // [world][region]      canonical batch index of the key at its requester
  uint4* recv_grads;                 // [world][region][cpr] pre-reduced gradient rows pushed by source s
  uint4* ret_rows;                   // [world][region][cpr] rows returned by owner o for my p-th key to it
  uint8_t* ret_status;               // [world][region]
  char* out;                         // output area: out_buffers x max_batch rows (meepo_peer_output)
};

struct PeerSet {
  PeerWindow w[kMaxPeers];
  uint32_t world, rank, region, cpr, chunks;
  uint32_t* err;  // this table's sticky error words (table.h kErr*): kernels only ever store 1
};
__device__ __forceinline__ void raise_error(const PeerSet& ps, int which) {
  reinterpret_cast<volatile uint32_t*>(ps.err)[which] = 1u;
}

// Device-resident bookkeeping of the owner side of one phase (filled by the barrier kernel): source s delivered
// the entries [lo[s], hi[s]) of its lane in this phase.
struct PeerWork {
  uint32_t lo[kMaxPeers], hi[kMaxPeers];
  uint32_t off[kMaxPeers + 1];  // exclusive prefix of hi - lo
  uint32_t max_hi;              // max over sources of hi
  uint32_t total;               // off[world]
  uint32_t pad[5];
};

struct PeerBlob {  // what ranks hand each other (opaque to the caller, MEEPO_PEER_BLOB_BYTES)
  uint32_t magic, abi;
  int32_t pid, device;
  uint32_t rank, world;
  uint64_t region, max_batch;
  uint32_t dim, dtype, opt, flags;
  uint64_t window_bytes;
  uint64_t raw_ptr;
  uint32_t out_buffers, chunks;
  cudaIpcMemHandle_t handle;
};
static_assert(sizeof(PeerBlob) <= MEEPO_PEER_BLOB_BYTES, "blob too large");

struct PeerState {
  bool attached = false;
  uint32_t world = 0, rank = 0, chunks = 1, out_buffers = 0;
  uint64_t max_batch = 0, region = 0;
  char* window = nullptr;  // this rank's exchange window (IPC-exported)
  size_t window_bytes = 0, out_bytes = 0;
  char* out_base = nullptr;  // output area inside the window
  char* local = nullptr;     // private scratch
  uint32_t* hist = nullptr;        // [kMaxBins]      unique keys per (owner, chunk) of the call in flight
  uint32_t* cursor = nullptr;      // [kMaxBins]      positions handed out so far inside each bin
  uint32_t* chunk_cnt = nullptr;   // [kMaxChunks][kMaxPeers] cumulative entries per owner after chunk c (kept for reuse)
  PeerWork* work = nullptr;        // [kMaxChunks + 1] owner side: one per backward chunk, the last for forward verbs
  uint32_t* loc = nullptr;         // [max_batch]     window position of unique key u: owner * region + p
  uint32_t* canon = nullptr;       // [max_batch]     canonical batch index of unique key u
  uint4* trash_row = nullptr;      // [cpr]           where rows of keys beyond a full lane go
  // the dedup of the last verb (persistent so that a backward pass can reuse the forward pass's)
  uint64_t* ukeys = nullptr;       // [max_batch]     unique keys
  uint32_t* inverse = nullptr;     // [max_batch]     unique id of every batch element
  uint64_t* n_unique = nullptr;    // [1]
  uint64_t* fwd_keys = nullptr;    // [max_batch]     copy of the last forward batch
  uint32_t* reuse_flag = nullptr;  // [1]             this backward batch == the last forward batch
  uint32_t* entry_slot = nullptr;  // [world*region]  owner side: slot of every entry of the last forward pass
  bool fwd_valid = false;          // the last sharded verb on this table was a forward pass ...
  uint64_t fwd_n = 0;              // ... of this many keys
  uint64_t entry_gen = ~0ull;      // table slot generation when entry_slot was written
  PeerSet ps{};
  void* opened[kMaxPeers] = {};
  unsigned long long seq = 0;
  unsigned long long timeout_ns = 20ull * 1000000000ull;
  cudaStream_t owner_stream = nullptr;  // T: owner side of the chunked backward pass
  cudaEvent_t ev_chunk[kMaxChunks] = {}, ev_owner = nullptr;
  float sender_share = 0.25f;  // share of the CTA slots the NVLink-bound sender kernels get while the owner side runs
};

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Offsets are a pure function of (world, region, cpr, out_bytes), hence identical on every rank.
static void carve_window(char* base, uint32_t world, uint64_t region, uint32_t cpr, size_t out_bytes, PeerWindow& w,
                         size_t* total) {
  size_t off = 0;
  w.flags = reinterpret_cast<unsigned long long*>(base + off);
  w.recv_cnt = reinterpret_cast<uint32_t*>(base + off + 128);
  w.recv_reuse = reinterpret_cast<uint32_t*>(base + off + 192);
  w.recv_out_off = reinterpret_cast<unsigned long long*>(base + off + 256);
  off += kWindowHeader;
  const size_t cells = (size_t)world * region;
  w.recv_keys = reinterpret_cast<uint64_t*>(base + off);
  off += align_up(cells * 8);
  w.recv_occ = reinterpret_cast<uint32_t*>(base + off);
  off += align_up(cells * 4);
  w.recv_dest = reinterpret_cast<uint32_t*>(base + off);
  off += align_up(cells * 4);
  w.recv_grads = reinterpret_cast<uint4*>(base + off);
  off += align_up(cells * cpr * 16);
  w.ret_rows = reinterpret_cast<uint4*>(base + off);
  off += align_up(cells * cpr * 16);
  w.ret_status = reinterpret_cast<uint8_t*>(base + off);
  off += align_up(cells);
  w.out = base + off;
  off += align_up(out_bytes);
  if (total) *total = off;
}

// --- system-scope flag accesses ----------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// data another GPU stored into this rank's window: read at L2 (the point of coherence), never L1
__device__ __forceinline__ uint64_t ld_window(const uint64_t* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long ld_window(const unsigned long long* p) { return __ldcg(p); }
__device__ __forceinline__ uint32_t ld_window(const uint32_t* p) { return __ldcg(p); }

// chunk of a unique key: a hash bit field independent of the owner (which takes the high bits of the same hash)
__device__ __forceinline__ uint32_t chunk_of_hash(uint64_t h, uint32_t chunks) {
  return (uint32_t)(((h & 0xFFFFFFFFull) * chunks) >> 32);
}

// --- barrier -----------------------------------------------------------------------------------
// All-to-all flag barrier between kernels: thread j delivers this rank's count for peer j, releases
// sequence number `seq` into j's window and then waits until j's number has arrived here. Bounded:
// a peer that never shows up raises kErrPeerTimeout instead of hanging the GPU.
//   send_cnt   (optional) [kMaxPeers] entries this rank has delivered to each owner so far in this verb
//   work/prev  (optional) owner-side bookkeeping of this phase; lo = prev->hi (0 without prev)
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerSet ps, unsigned long long seq,
                                                          const uint32_t* send_cnt, PeerWork* work,
                                                          const PeerWork* prev, int for_apply,
                                                          unsigned long long* counters, const uint32_t* reuse_flag,
                                                          int send_out, unsigned long long out_off,
                                                          unsigned long long timeout_ns) {
  const uint32_t j = threadIdx.x;
  if (j < ps.world) {
    if (send_cnt) ps.w[j].recv_cnt[ps.rank] = min(send_cnt[j], ps.region);
    if (reuse_flag) ps.w[j].recv_reuse[ps.rank] = *reuse_flag ? 1u : 0u;
    if (send_out) ps.w[j].recv_out_off[ps.rank] = out_off;
    __threadfence_system();
    st_release_sys(ps.w[j].flags + ps.rank, seq);
    const unsigned long long* mine = ps.w[ps.rank].flags + j;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(mine) < seq) {
      if (global_timer_ns() - t0 > timeout_ns) {
        raise_error(ps, kErrPeerTimeout);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (work && threadIdx.x == 0) {
    uint32_t ro = 0, mx = 0;
    for (uint32_t s = 0; s < ps.world; s++) {
      const uint32_t hi = send_cnt ? ld_window(ps.w[ps.rank].recv_cnt + s) : 0u;
      const uint32_t lo = prev ? min(prev->hi[s], hi) : 0u;
      work->lo[s] = lo;
      work->hi[s] = hi;
      work->off[s] = ro;
      ro += hi - lo;
      mx = max(mx, hi);
    }
    work->off[ps.world] = ro;
    work->total = ro;
    work->max_hi = mx;
    if (ro) atomicAdd(counters + (for_apply ? C_PEER_GRADS : C_PEER_KEYS), (unsigned long long)ro);
  }
}

// --- requester: push -----------------------------------------------------------------------------
// Both passes walk the BATCH (not the unique-key list) and take the canonical occurrence of every key, so that
// the positions one CTA claims in an owner's lane belong to 2048 consecutive batch elements: when the owner later
// stores the rows of 32 consecutive positions straight into the requester's output tensor, they land within ~1 MB
// of one another (scattering them over the whole 2 GB tensor costs an address translation per row over NVLink:
// measured 10x slower).
__device__ __forceinline__ bool canonical(const uint32_t* __restrict__ inverse, const uint32_t* __restrict__ canon,
                                          uint32_t i, uint32_t& u) {
  u = __ldg(inverse + i);
  return u != kNil && __ldg(canon + u) == i;
}
// Pass 1: unique keys per (owner, chunk) bin.
__global__ void __launch_bounds__(256) owner_hist_kernel(const __grid_constant__ PeerSet ps,
                                                         const uint64_t* __restrict__ keys,
                                                         const uint32_t* __restrict__ inverse,
                                                         const uint32_t* __restrict__ canon, uint32_t n,
                                                         uint32_t* __restrict__ hist, const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ uint32_t sh[kMaxBins];
  if (threadIdx.x < kMaxBins) sh[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t u;
    if (!canonical(inverse, canon, i, u)) continue;
    const uint64_t h = mix64(__ldg(keys + i) ^ MEEPO_OWNER_SALT);
    const uint32_t bin = (uint32_t)__umul64hi(h, (uint64_t)ps.world) * ps.chunks + chunk_of_hash(h, ps.chunks);
    atomicAdd(&sh[bin], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kMaxBins && sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh[threadIdx.x]);
}

// Pass 2: a CTA sorts the unique keys of its tile of the batch by bin in shared memory, claims one position range
// per bin and stores every run contiguously into the owner's window. Inside a (sender, owner) lane the chunks are
// consecutive runs: chunk c = [base[o][c], base[o][c+1]). loc[u] = owner * region + position (kNil: lane full).
constexpr int kPushThreads = 256, kPushItems = 8, kPushTile = kPushThreads * kPushItems;
__global__ void __launch_bounds__(kPushThreads) push_scatter_kernel(
    const __grid_constant__ PeerSet ps, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ inverse,
    const uint32_t* __restrict__ canon, const uint32_t* __restrict__ uocc, uint32_t n,
    const uint32_t* __restrict__ hist, uint32_t* __restrict__ cursor, uint32_t* __restrict__ chunk_cnt,
    uint32_t* __restrict__ loc, const uint32_t* __restrict__ skip) {
  if (skip && *skip) return;
  __shared__ uint64_t s_key[kPushTile];
  __shared__ uint32_t s_occ[kPushTile], s_dest[kPushTile], s_uid[kPushTile];
  __shared__ uint8_t s_bin[kPushTile];
  __shared__ uint32_t s_base[kMaxBins], s_cnt[kMaxBins], s_start[kMaxBins], s_gbase[kMaxBins];
  __shared__ uint32_t s_total;
  const uint32_t tid = threadIdx.x, K = ps.chunks, nbins = ps.world * K;
  if (tid < kMaxBins) {
    uint32_t base = 0;
    if (tid < nbins) {
      const uint32_t o = tid / K, c = tid % K;
      for (uint32_t cc = 0; cc < c; cc++) base += hist[o * K + cc];
      if (blockIdx.x == 0) chunk_cnt[c * kMaxPeers + o] = base + hist[tid];  // cumulative after chunk c
    }
    s_base[tid] = base;
  }
  for (uint32_t tile0 = blockIdx.x * kPushTile; tile0 < n; tile0 += gridDim.x * kPushTile) {
    if (tid < kMaxBins) s_cnt[tid] = 0;
    __syncthreads();
    uint64_t key[kPushItems];
    uint32_t bin[kPushItems], rank[kPushItems], uid[kPushItems];
#pragma unroll
    for (int k = 0; k < kPushItems; k++) {
      const uint32_t i = tile0 + k * kPushThreads + tid;
      bin[k] = kNil;
      key[k] = 0;
      rank[k] = 0;
      uid[k] = kNil;
      if (i < n && canonical(inverse, canon, i, uid[k])) {
        key[k] = __ldg(keys + i);
        const uint64_t h = mix64(key[k] ^ MEEPO_OWNER_SALT);
        bin[k] = (uint32_t)__umul64hi(h, (uint64_t)ps.world) * K + chunk_of_hash(h, K);
        rank[k] = atomicAdd(&s_cnt[bin[k]], 1u);
      }
    }
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the bin counts; one global claim per non-empty bin
      const uint32_t x = tid < nbins ? s_cnt[tid] : 0u;
      uint32_t incl = x;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (tid >= (uint32_t)d) incl += y;
      }
      s_start[tid] = incl - x;
      if (tid == 31) s_total = incl;
      if (x) s_gbase[tid] = s_base[tid] + atomicAdd(cursor + tid, x);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPushItems; k++) {
      if (bin[k] == kNil) continue;
      const uint32_t j = s_start[bin[k]] + rank[k];
      s_key[j] = key[k];
      s_occ[j] = uocc ? __ldg(uocc + uid[k]) : 1u;
      s_dest[j] = tile0 + k * kPushThreads + tid;
      s_uid[j] = uid[k];
      s_bin[j] = (uint8_t)bin[k];
    }
    __syncthreads();
    const uint32_t tile_n = s_total;
    for (uint32_t j = tid; j < tile_n; j += kPushThreads) {
      const uint32_t b = s_bin[j], o = b / K;
      const uint32_t p = s_gbase[b] + (j - s_start[b]);
      const uint32_t u = s_uid[j];
      if (p >= ps.region) {
        raise_error(ps, kErrPeerOverflow);
        loc[u] = kNil;
        continue;
      }
      const size_t e = (size_t)ps.rank * ps.region + p;
      ps.w[o].recv_keys[e] = s_key[j];
      ps.w[o].recv_occ[e] = s_occ[j];
      ps.w[o].recv_dest[e] = s_dest[j];
      loc[u] = o * ps.region + p;
    }
    __syncthreads();
  }
}

// Backward: is this the batch of the last forward pass? (flag preset to 1, cleared on any mismatch)
__global__ void __launch_bounds__(256) same_batch_kernel(const uint64_t* __restrict__ keys,
                                                         const uint64_t* __restrict__ fwd_keys, uint32_t n,
                                                         uint32_t* __restrict__ flag) {
  bool same = true;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    same &= __ldg(keys + i) == __ldg(fwd_keys + i);
  if (!__all_sync(0xFFFFFFFFu, same) && (threadIdx.x & 31u) == 0) *flag = 0u;
}

// Backward, sender side: where the summed gradient row of every unique key goes (straight into its owner's
// window), and the sort input of the pre-reduction: key = (chunk << ubits) | unique id, value = batch index;
// invalid keys get chunk K, sort last and are skipped.
__global__ void __launch_bounds__(256) grad_prep_kernel(const __grid_constant__ PeerSet ps,
                                                        const uint64_t* __restrict__ ukeys,
                                                        const unsigned long long* __restrict__ n_unique,
                                                        const uint32_t* __restrict__ loc,
                                                        const uint32_t* __restrict__ inverse, uint32_t n, int ubits,
                                                        uint4** __restrict__ row_ptrs, uint4* trash_row,
                                                        uint32_t* __restrict__ sort_key,
                                                        uint32_t* __restrict__ sort_val) {
  const uint32_t nu = (uint32_t)*n_unique;
  const uint32_t stride = gridDim.x * blockDim.x, first = blockIdx.x * blockDim.x + threadIdx.x;
  for (uint32_t u = first; u < nu; u += stride) {
    const uint32_t l = __ldg(loc + u);
    if (l == kNil) {
      row_ptrs[u] = trash_row;
    } else {
      const uint32_t o = l / ps.region, pos = l - o * ps.region;
      row_ptrs[u] = ps.w[o].recv_grads + ((size_t)ps.rank * ps.region + pos) * ps.cpr;
    }
  }
  for (uint32_t i = first; i < n; i += stride) {
    const uint32_t u = __ldg(inverse + i);
    uint32_t k = ps.chunks << ubits;
    if (u != kNil) k = (chunk_of_hash(mix64(__ldg(ukeys + u) ^ MEEPO_OWNER_SALT), ps.chunks) << ubits) | u;
    sort_key[i] = k;
    sort_val[i] = i;
  }
}

// --- owner: probe + gather + peer store ------------------------------------------------------------
// Tile = 32 consecutive positions of ONE source's lane. Every row goes to its own destination (SCATTER tile
// body): the requester's output tensor at the key's canonical batch index, or the requester's return region.
template <int CPR, bool INSERT, bool TIER = false>
__global__ void __launch_bounds__(256, 3) owner_probe_gather_kernel(TableView t, const __grid_constant__ PeerSet ps,
                                                                 const PeerWork* __restrict__ work, NewList nl,
                                                                 uint32_t* __restrict__ entry_slot) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t cpr = CPR > 0 ? (uint32_t)CPR : t.cpr;
  const PeerWindow& me = ps.w[ps.rank];
  const uint32_t max_tiles = (work->max_hi + 31u) >> 5;
  TileCounts cnt;
  __shared__ uint32_t sc_slot[kScoreCells], sc_freq[kScoreCells];
  const ScoreCache scache{sc_slot, sc_freq};
  score_cache_init(t, scache);
  // Row k of every source's tile list, sources rotated by (rank, k): at any moment the owners are
  // spread over all requesters (no incast on one NVLink port) and local tiles overlap remote ones.
  for (uint32_t k = warp; k < max_tiles; k += nwarps) {
    for (uint32_t j = 0; j < ps.world; j++) {
      const uint32_t s = (ps.rank + 1u + j + k) % ps.world;
      const uint32_t p0 = k * 32u;
      const uint32_t cnt_s = work->hi[s];
      if (p0 >= cnt_s) continue;
      const uint32_t tile_keys = min(32u, cnt_s - p0);
      const bool valid = lane < tile_keys;
      const size_t e = (size_t)s * ps.region + p0 + lane;  // in my window: [source][position]
      const uint64_t key = valid ? ld_window(me.recv_keys + e) : MEEPO_KEY_EMPTY;
      const uint32_t occ = valid ? ld_window(me.recv_occ + e) : 0u;
      const size_t r = (size_t)ps.rank * ps.region + p0 + lane;  // in the requester's window: [owner][position]
      const unsigned long long out_off = ld_window(me.recv_out_off + s);
      unsigned long long dst = 0;
      if (valid) {
        if (out_off != kNoOutput)
          dst = (unsigned long long)(ps.w[s].out + out_off + (size_t)ld_window(me.recv_dest + e) * cpr * 16u);
        else
          dst = (unsigned long long)(ps.w[s].ret_rows + r * cpr);
      }
      probe_gather_tile<CPR, INSERT, true, TIER>(t, key, valid, tile_keys, nullptr, valid ? ps.w[s].ret_status + r : nullptr,
                                           entry_slot + e, nullptr, occ, nl.slots + e, cnt, scache, lane, dst);
    }
  }
  score_cache_flush(t, scache);
  flush_tile_counts(t, cnt, lane);
}

// A shard with a host tier runs the generic-width kernel (the tier variants are not specialised per row width).
template <bool INSERT>
static const void* pick_owner_kernel(uint32_t cpr, bool tier = false) {
  if (tier) return (const void*)owner_probe_gather_kernel<0, INSERT, true>;
  switch (cpr) {
    case 1: return (const void*)owner_probe_gather_kernel<1, INSERT>;
    case 2: return (const void*)owner_probe_gather_kernel<2, INSERT>;
    case 4: return (const void*)owner_probe_gather_kernel<4, INSERT>;
    case 8: return (const void*)owner_probe_gather_kernel<8, INSERT>;
    case 16: return (const void*)owner_probe_gather_kernel<16, INSERT>;
    case 32: return (const void*)owner_probe_gather_kernel<32, INSERT>;
    case 64: return (const void*)owner_probe_gather_kernel<64, INSERT>;
    default: return (const void*)owner_probe_gather_kernel<0, INSERT>;
  }
}

// --- requester: finish -----------------------------------------------------------------------------
// One warp per 32 batch elements. Status of every element; then the rows that are not in place yet, moved as a
// flat array of 16-byte chunks with up to 4 loads in flight per lane:
//   direct (the owners stored each unique key's row at its canonical occurrence): the other occurrences copy it
//          from there, invalid keys / keys beyond a full lane get zeros, canonical occurrences are left alone;
//   else   every element copies its row out of the return region.
__global__ void __launch_bounds__(256) finish_kernel(const uint4* __restrict__ ret_rows,
                                                     const uint8_t* __restrict__ ret_status,
                                                     const uint32_t* __restrict__ loc,
                                                     const uint32_t* __restrict__ inverse,
                                                     const uint32_t* __restrict__ canon, int direct, uint32_t n,
                                                     uint32_t cpr, uint4* out, uint8_t* __restrict__ status_out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t ntiles = (n + 31u) >> 5;
  constexpr uint32_t kKeep = 0xFFFFFFFEu;  // the row is already in place
  for (uint32_t tile = warp; tile < ntiles; tile += nwarps) {
    const uint32_t i = tile * 32u + lane;
    uint32_t src = kKeep;  // row index into `from` (kNil: zeros, kKeep: nothing to do)
    if (i < n) {
      src = kNil;
      const uint32_t u = __ldg(inverse + i);
      uint8_t st = MEEPO_KEY_INVALID;
      if (u != kNil) {
        const uint32_t l = __ldg(loc + u);
        st = l != kNil ? __ldcg(ret_status + l) : (uint8_t)MEEPO_KEY_FULL;
        if (l != kNil) {
          if (direct) {
            const uint32_t c = __ldg(canon + u);
            src = c == i ? kKeep : c;
          } else {
            src = l;
          }
        }
      }
      if (status_out) status_out[i] = st;
    }
    if (__all_sync(0xFFFFFFFFu, src == kKeep)) continue;
    const uint4* from = direct ? out : ret_rows;
    const uint32_t chunks = min(32u, n - tile * 32u) * cpr;
    uint4* out_tile = out + (size_t)tile * 32u * cpr;
    for (uint32_t c0 = 0; c0 < chunks; c0 += 128) {
      uint4 v[4];
      bool keep[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t c = c0 + k * 32 + lane;
        const uint32_t j = min(c / cpr, 31u);
        const uint32_t s = __shfl_sync(0xFFFFFFFFu, src, j);
        v[k] = make_uint4(0, 0, 0, 0);
        keep[k] = s == kKeep;
        if (c < chunks && s != kNil && s != kKeep) v[k] = __ldcg(from + (size_t)s * cpr + (c - j * cpr));
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t c = c0 + k * 32 + lane;
        if (c < chunks && !keep[k]) st_stream(out_tile + c, v[k]);
      }
    }
  }
}

// --- owner: slot of every gradient entry of this chunk (sort key) + its position in the window (value) ---
// Entries enumerate rank-major, so the stable sort keeps each key's partial sums in rank order.
__global__ void __launch_bounds__(256) recv_slots_kernel(TableView t, const __grid_constant__ PeerSet ps,
                                                         const PeerWork* __restrict__ work,
                                                         uint32_t* __restrict__ sort_key,
                                                         uint32_t* __restrict__ sort_val,
                                                         const uint32_t* __restrict__ entry_slot, int entry_valid) {
  const uint32_t n = work->total;
  const PeerWindow& me = ps.w[ps.rank];
  uint32_t dropped = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t s = 0;
    for (uint32_t k = 1; k < ps.world; k++) s += work->off[k] <= i ? 1u : 0u;
    const uint32_t e = s * ps.region + work->lo[s] + (i - work->off[s]);
    uint32_t f;
    if (entry_valid && ld_window(me.recv_reuse + s)) {  // the entries of the last forward pass, re-sent
      f = __ldg(entry_slot + e);
    } else {
      const uint64_t key = ld_window(me.recv_keys + e);
      f = key_valid(key) ? probe_find<kReadOnly>(t, key) : kNil;
    }
    if (f == kNil) {
      dropped++;
      f = t.slots;  // sorts after every real slot
    }
    sort_key[i] = f;
    sort_val[i] = e;
  }
  dropped = __reduce_add_sync(0xFFFFFFFFu, dropped);
  if ((threadIdx.x & 31u) == 0 && dropped) atomicAdd(t.counters + C_DROPPED, (unsigned long long)dropped);
}

// --- host side ---------------------------------------------------------------------------------
static size_t forward_ws_bytes(const meepo_table* t, const PeerState* p) {
  const uint64_t n = p->max_batch;
  return dedup_bytes(t, n, false) + Workspace::pad(n * 4) + Workspace::pad((size_t)p->world * p->region * 4) + 4096;
}
static size_t backward_ws_bytes(const meepo_table* t, const PeerState* p) {
  const uint64_t n = p->max_batch;
  return dedup_bytes(t, n, false) + Workspace::pad(n * 8) + 256 + SegWork::bytes(n, t->v.dim, 32) +
         SegWork::bytes((uint64_t)p->world * p->region, t->v.dim, bits_for(t->v.slots)) + 8192;
}

void destroy_peer(meepo_table* t) {
  PeerState* p = t->peer;
  if (!p) return;
  for (uint32_t j = 0; j < kMaxPeers; j++)
    if (p->opened[j]) cudaIpcCloseMemHandle(p->opened[j]);
  if (p->owner_stream) cudaStreamDestroy(p->owner_stream);
  for (auto& e : p->ev_chunk)
    if (e) cudaEventDestroy(e);
  if (p->ev_owner) cudaEventDestroy(p->ev_owner);
  cudaFree(p->window);
  cudaFree(p->local);
  delete p;
  t->peer = nullptr;
}

// counts/work/prev: see peer_barrier_kernel. out_off: delivered to the owners with a forward verb's first barrier.
static meepo_status barrier(meepo_table* t, const uint32_t* send_cnt, PeerWork* work, const PeerWork* prev, bool for_apply,
                            const uint32_t* reuse_flag, bool send_out, unsigned long long out_off, cudaStream_t stream) {
  PeerState* p = t->peer;
  ProfScope ps(t, "sharded.barrier", stream);
  p->seq++;
  peer_barrier_kernel<<<1, 32, 0, stream>>>(p->ps, p->seq, send_cnt, work, prev, for_apply ? 1 : 0, t->v.counters,
                                            reuse_flag, send_out ? 1 : 0, out_off, p->timeout_ns);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

static meepo_status check_sharded(meepo_table* t, const void* keys, uint64_t n, const void* buf) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (!t->peer || !t->peer->attached) return fail(MEEPO_EINVAL, "meepo_peer_attach has not been called");
  if (n > t->peer->max_batch) return fail(MEEPO_EINVAL, "batch larger than the max_batch given to meepo_peer_prepare");
  if (n && (!keys || !buf)) return fail(MEEPO_EINVAL, "null buffer");
  return MEEPO_OK;
}

// keys -> their owners' windows (two kernels); leaves loc[] and chunk_cnt[][]. `skip`: device flag, no-op when set.
static meepo_status push_keys(meepo_table* t, const uint64_t* keys, const uint32_t* uocc, uint64_t n,
                              const uint32_t* skip, cudaStream_t stream) {
  PeerState* p = t->peer;
  ProfScope ps(t, "sharded.push_keys(2 kernels)", stream);
  MEEPO_CUDA_TRY(cudaMemsetAsync(p->hist, 0, 2 * kMaxBins * 4, stream));  // hist + cursor
  const uint64_t n1 = std::max<uint64_t>(n, 1);
  const int g1 = grid_for(t, (const void*)owner_hist_kernel, 256, 0, (n1 + 1023) / 1024);
  owner_hist_kernel<<<g1, 256, 0, stream>>>(p->ps, keys, p->inverse, p->canon, (uint32_t)n, p->hist, skip);
  const int g2 = grid_for(t, (const void*)push_scatter_kernel, kPushThreads, 0, (n1 + kPushTile - 1) / kPushTile);
  push_scatter_kernel<<<g2, kPushThreads, 0, stream>>>(p->ps, keys, p->inverse, p->canon, uocc, (uint32_t)n, p->hist,
                                                       p->cursor, p->chunk_cnt, p->loc, skip);
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

static meepo_status sharded_forward(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                    uint8_t* status_out, bool insert, cudaStream_t stream) {
  MEEPO_TRY(check_sharded(t, keys, n, rows_out));
  DeviceGuard guard(t->device);
  PeerState* p = t->peer;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  t->cache_valid = false;
  t->epoch++;
  t->v.epoch = (uint32_t)t->epoch;
  p->fwd_valid = false;
  MEEPO_TRY(t->ws.reserve(forward_ws_bytes(t, p), stream));
  uint32_t* uocc = t->ws.take<uint32_t>(std::max<uint64_t>(n, 1));  // hit/miss stats and LFU scores count occurrences
  if (n) MEEPO_CUDA_TRY(cudaMemcpyAsync(p->fwd_keys, keys, n * 8, cudaMemcpyDeviceToDevice, stream));
  const uint64_t n_window = (uint64_t)p->world * p->region;
  NewList nl{t->ws.take<uint32_t>(n_window)};  // one cell per window entry
  if (insert) MEEPO_CUDA_TRY(cudaMemsetAsync(nl.slots, 0xFF, n_window * 4, stream));
  // does the output tensor live in this rank's output area? then the owners store into it directly
  const char* ro = reinterpret_cast<const char*>(rows_out);
  // Direct stores pay on 2 GPUs (no return region, no expansion pass). From 4 GPUs on every owner scatters 512-byte
  // rows over the WHOLE output buffer of every requester instead of its own slice of a return region: the set of
  // remote 2 MB pages in flight grows with the world size and the owner kernel falls off a cliff (measured:
  // 2.61 vs 2.35 ms at N=4, 6.45 vs 2.77 ms at N=8). MEEPO_PEER_DIRECT=1 / MEEPO_PEER_NO_DIRECT force the choice.
  static const bool force_direct = getenv("MEEPO_PEER_DIRECT") != nullptr && atoi(getenv("MEEPO_PEER_DIRECT")) != 0;
  const bool direct = n && p->out_bytes && ro >= p->out_base &&
                      ro + n * (size_t)t->row_bytes <= p->out_base + p->out_bytes && !getenv("MEEPO_PEER_NO_DIRECT") &&
                      (p->world <= 2 || force_direct);
  const unsigned long long out_off = direct ? (unsigned long long)(ro - p->out_base) : kNoOutput;
  MEEPO_TRY(dedup_run(t, keys, nullptr, n, DedupOut{p->ukeys, nullptr, p->inverse, p->n_unique, uocc, nullptr, p->canon},
                      stream));
  MEEPO_TRY(push_keys(t, keys, uocc, n, nullptr, stream));
  PeerWork* fwork = p->work + kMaxChunks;
  MEEPO_TRY(barrier(t, p->chunk_cnt + (p->chunks - 1) * kMaxPeers, fwork, nullptr, false, nullptr, true, out_off, stream));
  {
    ProfScope ps(t, insert ? "sharded.owner_find_or_insert" : "sharded.owner_lookup", stream);
    const void* kern = insert ? pick_owner_kernel<true>(t->v.cpr, t->v.tier.slabs != 0)
                              : pick_owner_kernel<false>(t->v.cpr, t->v.tier.slabs != 0);
    const uint64_t tiles = ((uint64_t)p->world * p->region + 31) / 32 + p->world;
    const int grid = grid_for(t, kern, 256, 0, (tiles + 7) / 8);
    const PeerWork* work = fwork;
    void* args[] = {&t->v, &p->ps, &work, &nl, &p->entry_slot};
    MEEPO_CUDA_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(256), args, 0, stream));
  }
  if (insert) {
    ProfScope ps(t, "find_or_insert.publish", stream);
    MEEPO_TRY(publish_slots(t, nl.slots, n_window, stream));
  }
  MEEPO_TRY(barrier(t, nullptr, nullptr, nullptr, false, nullptr, false, 0, stream));
  if (n) {
    ProfScope ps(t, direct ? "sharded.finish(direct)" : "sharded.finish(expand)", stream);
    const PeerWindow& me = p->ps.w[p->rank];
    const uint64_t tiles = (n + 31) / 32;
    const int grid = grid_for(t, (const void*)finish_kernel, 256, 0, (tiles + 7) / 8);
    finish_kernel<<<grid, 256, 0, stream>>>(me.ret_rows, me.ret_status, p->loc, p->inverse, p->canon, direct ? 1 : 0,
                                            (uint32_t)n, t->v.cpr, reinterpret_cast<uint4*>(rows_out), status_out);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  p->fwd_valid = true;  // a backward pass over the same batch may reuse the dedup, the positions and the slots
  p->fwd_n = n;
  p->entry_gen = t->slot_gen;
  return MEEPO_OK;
}

static meepo_status sharded_apply(meepo_table* t, const uint64_t* keys, const void* grads, uint64_t n,
                                  cudaStream_t stream) {
  MEEPO_TRY(check_sharded(t, keys, n, grads));
  DeviceGuard guard(t->device);
  PeerState* p = t->peer;
  VerbScope vs(t, stream);
  MEEPO_TRY(vs.rc);
  MEEPO_TRY(t->ws.reserve(backward_ws_bytes(t, p), stream));
  const uint32_t K = p->chunks;
  const int ubits = bits_for((uint32_t)std::max<uint64_t>(n, 2) - 1);  // unique ids are < n
  const int sbits = ubits + bits_for(K);                                // chunk 0..K (K = invalid) above them
  uint4** row_ptrs = t->ws.take<uint4*>(std::max<uint64_t>(n, 1));
  const uint64_t n_pad = (uint64_t)p->world * p->region;
  SegWork ow;  // owner side
  ow.take(t->ws, n_pad, t->v.dim, bits_for(t->v.slots));
  SegWork sw;  // sender side
  if (n) sw.take(t->ws, n, t->v.dim, sbits);
  // the batch of the last forward pass? decided on the device; only the first backward pass after it qualifies
  const bool may_reuse = p->fwd_valid && n == p->fwd_n && n > 0 && !getenv("MEEPO_PEER_NO_REUSE");
  p->fwd_valid = false;
  {
    ProfScope ps(t, "sharded.same_batch", stream);
    MEEPO_CUDA_TRY(cudaMemsetAsync(p->reuse_flag, may_reuse ? 1 : 0, 4, stream));  // any non-zero value = "same"
    if (may_reuse) {
      const int grid = grid_for(t, (const void*)same_batch_kernel, 256, 0, (n + 255) / 256);
      same_batch_kernel<<<grid, 256, 0, stream>>>(keys, p->fwd_keys, (uint32_t)n, p->reuse_flag);
    }
  }
  DedupOut dd{p->ukeys, nullptr, p->inverse, p->n_unique, nullptr, reinterpret_cast<void* const*>(row_ptrs), p->canon};
  SegWork unused;
  MEEPO_TRY(dedup_hash(t, keys, n, dd, false, unused, stream, p->reuse_flag));  // no-op when the flag is set
  MEEPO_TRY(push_keys(t, keys, nullptr, n, p->reuse_flag, stream));             // likewise
  static const char* const snames[5] = {"dedup.radix_sort", "dedup.segments", "dedup.reduce_store", "dedup.long_leaves",
                                        "dedup.long_finish"};
  static const char* const onames[5] = {"sharded.owner_sort", "sharded.owner_segments", "sharded.owner_apply",
                                        "sharded.owner_long_leaves", "sharded.owner_long_finish"};
  if (n) {
    {
      ProfScope ps(t, "sharded.grad_prep", stream);
      const int grid = grid_for(t, (const void*)grad_prep_kernel, 256, 0, (n + 255) / 256);
      grad_prep_kernel<<<grid, 256, 0, stream>>>(p->ps, p->ukeys, (const unsigned long long*)p->n_unique, p->loc,
                                                 p->inverse, (uint32_t)n, ubits, row_ptrs, p->trash_row, sw.sk_in,
                                                 sw.sv_in);
      MEEPO_CUDA_TRY(cudaGetLastError());
    }
    MEEPO_TRY(seg_sort_heads(t, sw, K << ubits, false, nullptr, stream, snames, 0, K > 1 ? ubits : -1));
  }
  // The owner side of chunk c runs on its own stream underneath the sender side of chunk c + 1.
  const bool overlap = K > 1;
  cudaStream_t T = overlap ? p->owner_stream : stream;
  const int entry_valid = p->entry_gen == t->slot_gen ? 1 : 0;  // did anything move slots since the forward pass?
  for (uint32_t c = 0; c < K; c++) {
    if (n) {
      t->grid_scale = overlap ? p->sender_share : 1.0f;
      const SegRange r{c << ubits, (c + 1) << ubits, (1u << ubits) - 1u, K > 1 ? (int)c : -1};
      meepo_status rc = seg_reduce(t, sw, grads, kReduceStoreOnly, r, nullptr, reinterpret_cast<void* const*>(row_ptrs),
                                   stream, nullptr, snames, c > 0);  // summed rows land in the owners' windows
      t->grid_scale = 1.0f;
      MEEPO_TRY(rc);
    }
    MEEPO_TRY(barrier(t, p->chunk_cnt + c * kMaxPeers, p->work + c, c ? p->work + c - 1 : nullptr, true, p->reuse_flag,
                      false, 0, stream));
    if (overlap) {
      MEEPO_CUDA_TRY(cudaEventRecord(p->ev_chunk[c], stream));
      MEEPO_CUDA_TRY(cudaStreamWaitEvent(T, p->ev_chunk[c], 0));
    }
    t->grid_scale = overlap && c + 1 < K ? 1.0f - p->sender_share : 1.0f;
    meepo_status rc = MEEPO_OK;
    {
      ProfScope ps(t, "sharded.owner_slots", T);
      const int grid = grid_for(t, (const void*)recv_slots_kernel, 256, 0, (n_pad + 255) / 256);
      recv_slots_kernel<<<grid, 256, 0, T>>>(t->v, p->ps, p->work + c, ow.sk_in, ow.sv_in, p->entry_slot, entry_valid);
      if (cudaGetLastError() != cudaSuccess) rc = fail(MEEPO_ECUDA, "recv_slots launch failed");
    }
    if (rc == MEEPO_OK)
      rc = seg_sort_heads(t, ow, t->v.slots, true, &p->work[c].total, T, onames, (uint32_t)(n_pad / K), -1);
    if (rc == MEEPO_OK)
      rc = seg_reduce(t, ow, p->ps.w[p->rank].recv_grads, t->v.opt, SegRange{0u, t->v.slots, 0xFFFFFFFFu, -1}, nullptr,
                      nullptr, T, nullptr, onames, false);
    t->grid_scale = 1.0f;
    MEEPO_TRY(rc);
  }
  if (overlap) {
    MEEPO_CUDA_TRY(cudaEventRecord(p->ev_owner, T));
    MEEPO_CUDA_TRY(cudaStreamWaitEvent(stream, p->ev_owner, 0));
  }
  return barrier(t, nullptr, nullptr, nullptr, false, nullptr, false, 0, stream);
}

// CUDA loads kernels lazily, and the first launch of a kernel may synchronise the whole context: a
// rank whose barrier kernel is already spinning would then wait for a launch that cannot start (a
// deadlock when one process drives several shards, a stall otherwise). So every kernel of the three
// verbs is launched once here, on this rank alone (world 1), over a batch of invalid keys: no key is
// touched, no counter moves, and the window header and sequence numbers are reset afterwards.
static meepo_status warm_up(meepo_table* t) {
  PeerState* p = t->peer;
  const PeerSet saved = p->ps;
  PeerSet self{};
  self.world = 1;
  self.rank = 0;
  self.region = (uint32_t)p->region;
  self.cpr = t->v.cpr;
  self.chunks = p->chunks;
  self.err = t->err_word;
  carve_window(p->window, p->world, p->region, t->v.cpr, p->out_bytes, self.w[0], nullptr);
  p->ps = self;
  p->attached = true;
  const uint64_t epoch = t->epoch;
  meepo_status rc = MEEPO_OK;
  const uint64_t sizes[2] = {std::min<uint64_t>(p->max_batch, 1024), std::min<uint64_t>(p->max_batch, 65536)};
  void* keys = nullptr;
  void* rows = nullptr;
  if (cudaMalloc(&keys, sizes[1] * 8) != cudaSuccess || cudaMalloc(&rows, sizes[1] * (size_t)t->row_bytes) != cudaSuccess) {
    cudaFree(keys);
    rc = fail(MEEPO_ENOMEM, "cudaMalloc(warm-up)");
  } else {
    cudaMemset(keys, 0xFF, sizes[1] * 8);  // MEEPO_KEY_EMPTY: invalid, dropped by the dedup
    cudaMemset(rows, 0, sizes[1] * (size_t)t->row_bytes);
    for (int k = 0; k < 2 && rc == MEEPO_OK; k++) {
      rc = sharded_forward(t, (const uint64_t*)keys, sizes[k], rows, nullptr, false, nullptr);
      if (rc == MEEPO_OK) rc = sharded_forward(t, (const uint64_t*)keys, sizes[k], rows, nullptr, true, nullptr);
      if (rc == MEEPO_OK) rc = sharded_apply(t, (const uint64_t*)keys, rows, sizes[k], nullptr);
      if (rc == MEEPO_OK && p->out_bytes)  // the direct finish
        rc = sharded_forward(t, (const uint64_t*)keys, sizes[k], p->out_base, nullptr, false, nullptr);
    }
    if (rc == MEEPO_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(MEEPO_ECUDA, "warm-up failed");
    cudaFree(keys);
    cudaFree(rows);
  }
  cudaMemset(p->window, 0, kWindowHeader);
  cudaDeviceSynchronize();
  p->seq = 0;
  p->ps = saved;
  p->attached = false;
  p->fwd_valid = false;
  t->epoch = epoch;
  t->v.epoch = (uint32_t)epoch;
  return rc;
}

}  // namespace meepo

using namespace meepo;

extern "C" {

MEEPO_API meepo_status meepo_peer_prepare(meepo_table* t, uint32_t rank, uint32_t world, uint64_t max_batch,
                                          uint64_t region_keys, uint32_t out_buffers, void* blob_out) {
  if (!t || !blob_out) return fail(MEEPO_EINVAL, "null argument");
  if (world == 0 || world > kMaxPeers || rank >= world) return fail(MEEPO_EINVAL, "need rank < world <= MEEPO_MAX_PEERS");
  if (max_batch == 0 || max_batch > (1ull << 28)) return fail(MEEPO_EINVAL, "bad max_batch (1 .. 2^28)");
  if (region_keys == 0 || region_keys > max_batch) region_keys = max_batch;
  if ((uint64_t)world * region_keys > 0x7FFFFFFFull) return fail(MEEPO_EINVAL, "world * region_keys must fit in 31 bits");
  if (out_buffers > 16) return fail(MEEPO_EINVAL, "at most 16 output buffers");
  if (t->peer) return fail(MEEPO_EINVAL, "meepo_peer_prepare was already called on this table");
  DeviceGuard guard(t->device);
  PeerState* p = new PeerState();
  t->peer = p;
  p->world = world;
  p->rank = rank;
  p->max_batch = max_batch;
  p->region = region_keys;
  p->out_buffers = out_buffers;
  p->out_bytes = (size_t)out_buffers * max_batch * t->row_bytes;
  // Chunks of the backward pass (the owner side of chunk c on its own stream under the sender side of chunk c + 1):
  // opt-in. Measured on 2 and 4 B200 the per-chunk sorts and the SM sharing of the two pipelines cost what the
  // overlap gains (N=4 cfg3: 8.47 ms with 1 chunk, 8.52 with 3), and with 3 chunks the forward pass that follows
  // loses its store locality into the requesters' output buffers (12.96 ms).
  p->chunks = 1;
  if (const char* e = getenv("MEEPO_PEER_CHUNKS"))
    p->chunks = (uint32_t)std::min<unsigned long>(kMaxChunks, std::max<unsigned long>(1, strtoul(e, nullptr, 10)));
  if (const char* e = getenv("MEEPO_PEER_SENDER_SHARE"))
    p->sender_share = std::min(0.9f, std::max(0.05f, strtof(e, nullptr)));
  if (const char* e = getenv("MEEPO_PEER_TIMEOUT_MS")) p->timeout_ns = strtoull(e, nullptr, 10) * 1000000ull;
  PeerWindow w;
  carve_window(nullptr, world, region_keys, t->v.cpr, p->out_bytes, w, &p->window_bytes);
  const size_t out_offset = (size_t)(w.out - (char*)nullptr);
  auto bail = [&](cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    destroy_peer(t);
    return fail(e == cudaErrorMemoryAllocation ? MEEPO_ENOMEM : MEEPO_ECUDA, m);
  };
  cudaError_t e;
  if ((e = cudaMalloc(&p->window, p->window_bytes)) != cudaSuccess) return bail(e, "cudaMalloc(exchange window)");
  if ((e = cudaMemset(p->window, 0, kWindowHeader)) != cudaSuccess) return bail(e, "cudaMemset");
  p->out_base = p->window + out_offset;
  if ((e = cudaStreamCreateWithFlags(&p->owner_stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail(e, "cudaStreamCreate");
  for (auto& ev : p->ev_chunk)
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreateWithFlags(&p->ev_owner, cudaEventDisableTiming)) != cudaSuccess)
    return bail(e, "cudaEventCreate");
  // private scratch
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += align_up(bytes);
    return o;
  };
  const size_t o_hist = take(2 * kMaxBins * 4);  // hist + cursor, zeroed together
  const size_t o_ccnt = take(kMaxChunks * kMaxPeers * 4);
  const size_t o_work = take((kMaxChunks + 1) * sizeof(PeerWork));
  const size_t o_trash = take((size_t)t->v.cpr * 16);
  const size_t o_flag = take(4);
  const size_t o_nu = take(8);
  const size_t o_zero_end = off;
  const size_t o_loc = take(max_batch * 4);
  const size_t o_canon = take(max_batch * 4);
  const size_t o_inv = take(max_batch * 4);
  const size_t o_ukeys = take(max_batch * 8);
  const size_t o_fkeys = take(max_batch * 8);
  const size_t o_eslot = take((size_t)world * region_keys * 4);
  if ((e = cudaMalloc(&p->local, off)) != cudaSuccess) return bail(e, "cudaMalloc(peer scratch)");
  if ((e = cudaMemset(p->local, 0, o_zero_end)) != cudaSuccess) return bail(e, "cudaMemset");
  p->hist = reinterpret_cast<uint32_t*>(p->local + o_hist);
  p->cursor = p->hist + kMaxBins;
  p->chunk_cnt = reinterpret_cast<uint32_t*>(p->local + o_ccnt);
  p->work = reinterpret_cast<PeerWork*>(p->local + o_work);
  p->trash_row = reinterpret_cast<uint4*>(p->local + o_trash);
  p->reuse_flag = reinterpret_cast<uint32_t*>(p->local + o_flag);
  p->n_unique = reinterpret_cast<uint64_t*>(p->local + o_nu);
  p->loc = reinterpret_cast<uint32_t*>(p->local + o_loc);
  p->canon = reinterpret_cast<uint32_t*>(p->local + o_canon);
  p->inverse = reinterpret_cast<uint32_t*>(p->local + o_inv);
  p->ukeys = reinterpret_cast<uint64_t*>(p->local + o_ukeys);
  p->fwd_keys = reinterpret_cast<uint64_t*>(p->local + o_fkeys);
  p->entry_slot = reinterpret_cast<uint32_t*>(p->local + o_eslot);
  // the workspace never grows (cudaFree = device-wide sync) once the verbs are in flight
  if (t->ws.reserve(std::max(forward_ws_bytes(t, p), backward_ws_bytes(t, p)), nullptr) != MEEPO_OK) {
    destroy_peer(t);
    return MEEPO_ENOMEM;
  }
  if (meepo_status rc = warm_up(t); rc != MEEPO_OK) {
    destroy_peer(t);
    return rc;
  }
  PeerBlob b;
  memset(&b, 0, sizeof b);
  b.magic = kBlobMagic;
  b.abi = MEEPO_ABI_VERSION;
  b.pid = (int32_t)getpid();
  b.device = t->device;
  b.rank = rank;
  b.world = world;
  b.region = region_keys;
  b.max_batch = max_batch;
  b.dim = t->v.dim;
  b.dtype = (uint32_t)t->v.dtype;
  b.opt = (uint32_t)t->v.opt;
  b.flags = t->cfg.flags;
  b.window_bytes = p->window_bytes;
  b.raw_ptr = (uint64_t)(uintptr_t)p->window;
  b.out_buffers = out_buffers;
  b.chunks = p->chunks;
  if ((e = cudaIpcGetMemHandle(&b.handle, p->window)) != cudaSuccess) return bail(e, "cudaIpcGetMemHandle");
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail(e, "cudaDeviceSynchronize");
  memset(blob_out, 0, MEEPO_PEER_BLOB_BYTES);
  memcpy(blob_out, &b, sizeof b);
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_peer_output(meepo_table* t, uint32_t index, void** rows_out, uint64_t* max_rows) {
  if (!t || !rows_out) return fail(MEEPO_EINVAL, "null argument");
  PeerState* p = t->peer;
  if (!p) return fail(MEEPO_EINVAL, "call meepo_peer_prepare first");
  if (index >= p->out_buffers) return fail(MEEPO_EINVAL, "no such output buffer (out_buffers of meepo_peer_prepare)");
  *rows_out = p->out_base + (size_t)index * p->max_batch * t->row_bytes;
  if (max_rows) *max_rows = p->max_batch;
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_peer_attach(meepo_table* t, const void* blobs) {
  if (!t || !blobs) return fail(MEEPO_EINVAL, "null argument");
  PeerState* p = t->peer;
  if (!p) return fail(MEEPO_EINVAL, "call meepo_peer_prepare first");
  if (p->attached) return fail(MEEPO_EINVAL, "already attached");
  DeviceGuard guard(t->device);
  p->ps.world = p->world;
  p->ps.rank = p->rank;
  p->ps.region = (uint32_t)p->region;
  p->ps.cpr = t->v.cpr;
  p->ps.chunks = p->chunks;
  p->ps.err = t->err_word;
  for (uint32_t j = 0; j < p->world; j++) {
    PeerBlob b;
    memcpy(&b, (const char*)blobs + (size_t)j * MEEPO_PEER_BLOB_BYTES, sizeof b);
    if (b.magic != kBlobMagic || b.abi != MEEPO_ABI_VERSION) return fail(MEEPO_EINVAL, "peer blob: bad magic / ABI");
    if (b.rank != j || b.world != p->world) return fail(MEEPO_EINVAL, "peer blobs must be ordered by rank");
    if (b.region != p->region || b.dim != t->v.dim || b.dtype != (uint32_t)t->v.dtype || b.opt != (uint32_t)t->v.opt ||
        b.flags != t->cfg.flags || b.window_bytes != p->window_bytes || b.max_batch != p->max_batch ||
        b.out_buffers != p->out_buffers || b.chunks != p->chunks)
      return fail(MEEPO_EINVAL, "peer blob: table geometry differs between ranks");
    char* base = nullptr;
    if (j == p->rank) {
      base = p->window;
    } else if (b.pid == (int32_t)getpid()) {  // several tables driven by one process
      base = reinterpret_cast<char*>((uintptr_t)b.raw_ptr);
      if (b.device != t->device) {
        int can = 0;
        MEEPO_CUDA_TRY(cudaDeviceCanAccessPeer(&can, t->device, b.device));
        if (!can) return fail(MEEPO_ECUDA, "no peer access between the devices of two shards");
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(MEEPO_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
    } else {
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return fail(MEEPO_ECUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(j) + "): " + cudaGetErrorString(e));
      p->opened[j] = ptr;
      base = reinterpret_cast<char*>(ptr);
    }
    carve_window(base, p->world, p->region, t->v.cpr, p->out_bytes, p->ps.w[j], nullptr);
  }
  p->attached = true;
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_peer_detach(meepo_table* t) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  DeviceGuard guard(t->device);
  cudaDeviceSynchronize();
  destroy_peer(t);
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_sharded_find_or_insert(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                                    uint8_t* status_out, void* stream) {
  return sharded_forward(t, keys, n, rows_out, status_out, true, (cudaStream_t)stream);
}
MEEPO_API meepo_status meepo_sharded_lookup(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                            uint8_t* found_out, void* stream) {
  return sharded_forward(t, keys, n, rows_out, found_out, false, (cudaStream_t)stream);
}
MEEPO_API meepo_status meepo_sharded_apply_gradients(meepo_table* t, const uint64_t* keys, const void* grads,
                                                     uint64_t n, void* stream) {
  return sharded_apply(t, keys, grads, n, (cudaStream_t)stream);
}

}  // extern "C"
