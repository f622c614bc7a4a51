// peer.cu — the key-hash-sharded verbs fused with their exchange over NVLink peer memory
// (SURVEY rows a4 / e; include/meepo.h "sharded verbs"). One process per GPU; every rank owns the
// keys with owner(key, world) == rank and exposes ONE exchange window (cudaIpc) that its peers
// store into directly. No NCCL call, no host synchronisation and no staging copy on the data path:
//
//   find_or_insert / lookup (requester r, owner o)
//     r: dedup the batch                              -> unique keys, inverse, occurrences
//     r: push_keys      key -> o.recv_keys[r][p]      8 B per unique key over NVLink (p = arrival order)
//     barrier (delivers the per-pair counts)
//     o: owner_probe_gather  probe the LOCAL table (same tile body as the single-table kernel) and
//                            store each row straight into r.ret_rows[o][p] over NVLink; tags of the
//                            slots claimed in this call are published afterwards, locally
//     barrier
//     r: expand         rows_out[i] = ret_rows[loc[inverse[i]]]
//   apply_gradients
//     r: dedup the batch; assign every unique key its position p at its owner (key -> o.recv_keys[r][p])
//     r: fixed-shape pre-reduction of duplicate gradients (sort by unique id + segmented sum, rounded
//        to the table dtype) whose store IS the exchange: each summed row goes straight to
//        o.recv_grads[r][p] over NVLink — no local copy of the unique gradients, no separate push
//     barrier
//     o: slot of every received entry, stable sort by slot (entries enumerate rank-major, so each
//        key's partial sums stay in rank order) and the same fused reduce + optimizer kernel as the
//        single table: ONE optimizer step per key, deterministic
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
constexpr uint32_t kBlobMagic = 0x4D50454Bu;  // "MPEK"
constexpr size_t kWindowHeader = 4096;


// A rank's exchange window as seen through a (peer or local) mapping.
struct PeerWindow {
  unsigned long long* flags;  // [kMaxPeers] barrier sequence number last signalled by each source
  uint32_t* recv_cnt;         // [kMaxPeers] entries source s pushed in the current phase
  uint32_t* recv_reuse;       // [kMaxPeers] source s re-sent the entries of its last forward pass (same positions)
  uint64_t* recv_keys;        // [world][region]      keys pushed by source s
  uint32_t* recv_occ;         // [world][region]      batch occurrences behind each pushed key
  uint4* recv_grads;          // [world][region][cpr] pre-reduced gradient rows pushed by source s
  uint4* ret_rows;            // [world][region][cpr] rows returned by owner o for my p-th key to it
  uint8_t* ret_status;        // [world][region]
};

struct PeerSet {
  PeerWindow w[kMaxPeers];
  uint32_t world, rank, region, cpr;
  uint32_t* err;  // this table's sticky error words (table.h kErr*): kernels only ever store 1
};
__device__ __forceinline__ void raise_error(const PeerSet& ps, int which) {
  reinterpret_cast<volatile uint32_t*>(ps.err)[which] = 1u;
}

// Device-resident bookkeeping of the owner side of one phase (filled by the barrier kernel).
struct PeerWork {
  uint32_t cnt[kMaxPeers];           // entries received from source s
  uint32_t recv_off[kMaxPeers + 1];  // exclusive prefix of cnt
  uint32_t tile_off[kMaxPeers + 1];  // exclusive prefix of ceil(cnt / 32)
  uint32_t max_cnt;                  // max over sources of cnt
  uint32_t pad[3];
};

struct PeerBlob {  // what ranks hand each other (opaque to the caller, MEEPO_PEER_BLOB_BYTES)
  uint32_t magic, abi;
  int32_t pid, device;
  uint32_t rank, world;
  uint64_t region, max_batch;
  uint32_t dim, dtype, opt, flags;
  uint64_t window_bytes;
  uint64_t raw_ptr;
  cudaIpcMemHandle_t handle;
};
static_assert(sizeof(PeerBlob) <= MEEPO_PEER_BLOB_BYTES, "blob too large");

struct PeerState {
  bool attached = false;
  uint32_t world = 0, rank = 0;
  uint64_t max_batch = 0, region = 0;
  char* window = nullptr;  // this rank's exchange window (IPC-exported)
  size_t window_bytes = 0;
  char* local = nullptr;   // private scratch
  uint32_t* send_cnt = nullptr;
  PeerWork* work = nullptr;
  uint32_t* loc = nullptr;         // [max_batch]     window position of unique key u: owner * region + p
  uint4* trash_row = nullptr;      // [cpr]           where rows of keys beyond a full lane go
  // the dedup of the last verb (persistent so that a backward pass can reuse the forward pass's)
  uint64_t* ukeys = nullptr;       // [max_batch]     unique keys
  uint32_t* inverse = nullptr;     // [max_batch]     unique id of every batch element
  uint64_t* n_unique = nullptr;    // [1]
  uint64_t* fwd_keys = nullptr;    // [max_batch]     copy of the last forward batch
  uint32_t* fwd_send_cnt = nullptr;  // [kMaxPeers]   entries per owner of the last forward push
  uint32_t* reuse_flag = nullptr;  // [1]             this backward batch == the last forward batch
  uint32_t* entry_slot = nullptr;  // [world*region]  owner side: slot of every entry of the last forward pass
  bool fwd_valid = false;          // the last sharded verb on this table was a forward pass ...
  uint64_t fwd_n = 0;              // ... of this many keys
  uint64_t entry_gen = ~0ull;      // table slot generation when entry_slot was written
  PeerSet ps{};
  void* opened[kMaxPeers] = {};
  unsigned long long seq = 0;
  unsigned long long timeout_ns = 20ull * 1000000000ull;
};

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Offsets are a pure function of (world, region, cpr), hence identical on every rank.
static void carve_window(char* base, uint32_t world, uint64_t region, uint32_t cpr, PeerWindow& w, size_t* total) {
  size_t off = 0;
  w.flags = reinterpret_cast<unsigned long long*>(base + off);
  w.recv_cnt = reinterpret_cast<uint32_t*>(base + off + 128);
  w.recv_reuse = reinterpret_cast<uint32_t*>(base + off + 192);
  off += kWindowHeader;
  const size_t cells = (size_t)world * region;
  w.recv_keys = reinterpret_cast<uint64_t*>(base + off);
  off += align_up(cells * 8);
  w.recv_occ = reinterpret_cast<uint32_t*>(base + off);
  off += align_up(cells * 4);
  w.recv_grads = reinterpret_cast<uint4*>(base + off);
  off += align_up(cells * cpr * 16);
  w.ret_rows = reinterpret_cast<uint4*>(base + off);
  off += align_up(cells * cpr * 16);
  w.ret_status = reinterpret_cast<uint8_t*>(base + off);
  off += align_up(cells);
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
__device__ __forceinline__ uint32_t ld_window(const uint32_t* p) { return __ldcg(p); }

// --- barrier -----------------------------------------------------------------------------------
// All-to-all flag barrier between kernels: thread j delivers this rank's count for peer j, releases
// sequence number `seq` into j's window and then waits until j's number has arrived here. Bounded:
// a peer that never shows up sets PE_TIMEOUT instead of hanging the GPU.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ PeerSet ps, unsigned long long seq,
                                                          const uint32_t* send_cnt, PeerWork* work, int for_apply,
                                                          unsigned long long* counters, uint32_t* save_cnt,
                                                          const uint32_t* reuse_flag,
                                                          unsigned long long timeout_ns) {
  const uint32_t j = threadIdx.x;
  if (j < ps.world) {
    if (send_cnt) ps.w[j].recv_cnt[ps.rank] = min(send_cnt[j], ps.region);
    if (send_cnt && save_cnt) save_cnt[j] = send_cnt[j];
    if (reuse_flag) ps.w[j].recv_reuse[ps.rank] = *reuse_flag ? 1u : 0u;
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
    uint32_t ro = 0, to = 0;
    for (uint32_t s = 0; s < ps.world; s++) {
      const uint32_t c = send_cnt ? ld_window(ps.w[ps.rank].recv_cnt + s) : 0u;
      work->cnt[s] = c;
      work->recv_off[s] = ro;
      work->tile_off[s] = to;
      ro += c;
      to += (c + 31u) >> 5;
    }
    work->recv_off[ps.world] = ro;
    work->tile_off[ps.world] = to;
    uint32_t mx = 0;
    for (uint32_t s = 0; s < ps.world; s++) mx = max(mx, work->cnt[s]);
    work->max_cnt = mx;
    if (ro) atomicAdd(counters + (for_apply ? C_PEER_GRADS : C_PEER_KEYS), (unsigned long long)ro);
  }
}

// --- requester: push -----------------------------------------------------------------------------
// p = arrival position of this lane's key at owner o (one atomic per distinct owner per warp)
__device__ __forceinline__ uint32_t claim_position(uint32_t* send_cnt, uint32_t o, unsigned active, uint32_t lane) {
  const unsigned peers = __match_any_sync(active, o);
  const int leader = __ffs(peers) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(send_cnt + o, (uint32_t)__popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(256) push_keys_kernel(const __grid_constant__ PeerSet ps,
                                                        const uint64_t* __restrict__ ukeys,
                                                        const uint32_t* __restrict__ uocc,
                                                        const unsigned long long* __restrict__ n_unique,
                                                        uint32_t* __restrict__ send_cnt, uint32_t* __restrict__ loc) {
  const uint32_t n = (uint32_t)*n_unique;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t u = base + lane;
    const bool act = u < n;
    const unsigned am = __ballot_sync(0xFFFFFFFFu, act);
    if (!act) continue;
    const uint64_t key = __ldg(ukeys + u);
    const uint32_t o = owner_of(key, ps.world);
    const uint32_t p = claim_position(send_cnt, o, am, lane);
    if (p >= ps.region) {
      raise_error(ps, kErrPeerOverflow);
      loc[u] = kNil;
      continue;
    }
    const size_t e = (size_t)ps.rank * ps.region + p;
    ps.w[o].recv_keys[e] = key;
    ps.w[o].recv_occ[e] = uocc ? __ldg(uocc + u) : 1u;
    loc[u] = o * ps.region + p;
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
// sort input of the pre-reduction: (unique id, batch index); invalid keys sort last and are skipped
__global__ void __launch_bounds__(256) fill_sort_kernel(const uint32_t* __restrict__ inverse, uint32_t n,
                                                        uint32_t* __restrict__ sort_key,
                                                        uint32_t* __restrict__ sort_val) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t u = __ldg(inverse + i);
    sort_key[i] = u == kNil ? n : u;
    sort_val[i] = i;
  }
}

// Backward: position of every unique key at its owner. The key goes there now; row_ptrs[u] is where
// the reduce kernel will store the key's summed gradient row (straight into the owner's window).
__global__ void __launch_bounds__(256) assign_grad_rows_kernel(const __grid_constant__ PeerSet ps,
                                                               const uint64_t* __restrict__ ukeys,
                                                               const unsigned long long* __restrict__ n_unique,
                                                               uint32_t* __restrict__ send_cnt,
                                                               uint4** __restrict__ row_ptrs, uint4* trash_row,
                                                               const uint32_t* __restrict__ reuse_flag,
                                                               uint32_t* __restrict__ loc,
                                                               const uint32_t* __restrict__ fwd_send_cnt) {
  const uint32_t n = (uint32_t)*n_unique;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  if (*reuse_flag) {  // same batch as the forward pass: same owners, same positions, keys already there
    if (blockIdx.x == 0 && threadIdx.x < ps.world) send_cnt[threadIdx.x] = fwd_send_cnt[threadIdx.x];
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += stride) {
      const uint32_t l = loc[u];
      if (l == kNil) {
        row_ptrs[u] = trash_row;
      } else {
        const uint32_t o = l / ps.region, pos = l - o * ps.region;
        row_ptrs[u] = ps.w[o].recv_grads + ((size_t)ps.rank * ps.region + pos) * ps.cpr;
      }
    }
    return;
  }
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
    const uint32_t u = base + lane;
    const bool act = u < n;
    const unsigned am = __ballot_sync(0xFFFFFFFFu, act);
    if (!act) continue;
    const uint64_t key = __ldg(ukeys + u);
    const uint32_t o = owner_of(key, ps.world);
    const uint32_t p = claim_position(send_cnt, o, am, lane);
    if (p >= ps.region) {
      raise_error(ps, kErrPeerOverflow);
      row_ptrs[u] = trash_row;
      loc[u] = kNil;
      continue;
    }
    const size_t e = (size_t)ps.rank * ps.region + p;
    ps.w[o].recv_keys[e] = key;
    row_ptrs[u] = ps.w[o].recv_grads + e * ps.cpr;
    loc[u] = o * ps.region + p;
  }
}

// --- owner: probe + gather + peer store ------------------------------------------------------------
// Tile t of the received keys = 32 consecutive positions of ONE source's region, so the rows of a
// tile go to one contiguous block of that requester's ret_rows and the tile body is exactly the
// single-table one with a peer pointer as its output.
template <int CPR, bool INSERT>
__global__ void __launch_bounds__(256, 4) owner_probe_gather_kernel(TableView t, const __grid_constant__ PeerSet ps,
                                                                 const PeerWork* __restrict__ work, NewList nl,
                                                                 uint32_t* __restrict__ entry_slot) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t cpr = CPR > 0 ? (uint32_t)CPR : t.cpr;
  const PeerWindow& me = ps.w[ps.rank];
  const uint32_t max_tiles = (work->max_cnt + 31u) >> 5;
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
      const uint32_t cnt_s = work->cnt[s];
      if (p0 >= cnt_s) continue;
      const uint32_t tile_keys = min(32u, cnt_s - p0);
      const bool valid = lane < tile_keys;
      const size_t e = (size_t)s * ps.region + p0 + lane;  // in my window: [source][position]
      const uint64_t key = valid ? ld_window(me.recv_keys + e) : MEEPO_KEY_EMPTY;
      const uint32_t occ = valid ? ld_window(me.recv_occ + e) : 0u;
      const size_t r = (size_t)ps.rank * ps.region + p0;    // in the requester's window: [owner][position]
      probe_gather_tile<CPR, INSERT>(t, key, valid, tile_keys, ps.w[s].ret_rows + r * cpr,
                                     valid ? ps.w[s].ret_status + r + lane : nullptr, entry_slot + e, nullptr, occ,
                                     nl.slots + e, cnt, scache, lane);
    }
  }
  score_cache_flush(t, scache);
  flush_tile_counts(t, cnt, lane);
}

template <bool INSERT>
static const void* pick_owner_kernel(uint32_t cpr) {
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

// --- requester: expand ---------------------------------------------------------------------------
// rows_out[i] = ret_rows[loc[inverse[i]]]: one warp moves 32 output rows as a flat chunk array.
__global__ void __launch_bounds__(256) expand_kernel(const uint4* __restrict__ ret_rows,
                                                     const uint8_t* __restrict__ ret_status,
                                                     const uint32_t* __restrict__ loc,
                                                     const uint32_t* __restrict__ inverse, uint32_t n, uint32_t cpr,
                                                     uint4* __restrict__ out, uint8_t* __restrict__ status_out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t ntiles = (n + 31u) >> 5;
  for (uint32_t tile = warp; tile < ntiles; tile += nwarps) {
    const uint32_t i = tile * 32u + lane;
    uint32_t src = kNil;
    if (i < n) {
      const uint32_t u = __ldg(inverse + i);
      uint8_t st = MEEPO_KEY_INVALID;
      if (u != kNil) {
        src = __ldg(loc + u);
        st = src != kNil ? __ldcg(ret_status + src) : (uint8_t)MEEPO_KEY_FULL;
      }
      if (status_out) status_out[i] = st;
    }
    const uint32_t chunks = min(32u, n - tile * 32u) * cpr;
    uint4* out_tile = out + (size_t)tile * 32u * cpr;
    for (uint32_t c0 = 0; c0 < chunks; c0 += 128) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t c = c0 + k * 32 + lane;
        const uint32_t j = min(c / cpr, 31u);
        const uint32_t s = __shfl_sync(0xFFFFFFFFu, src, j);
        v[k] = make_uint4(0, 0, 0, 0);
        if (c < chunks && s != kNil) v[k] = ld_stream(ret_rows + (size_t)s * cpr + (c - j * cpr));
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t c = c0 + k * 32 + lane;
        if (c < chunks) st_stream(out_tile + c, v[k]);
      }
    }
  }
}

// --- owner: slot of every received gradient entry (sort key) + its position in the window (value) ---
// Entries enumerate rank-major, so the stable sort keeps each key's partial sums in rank order.
// Positions past the received count pad the sort to its host-known size with the "absent" key.
__global__ void __launch_bounds__(256) recv_slots_kernel(TableView t, const __grid_constant__ PeerSet ps,
                                                         const PeerWork* __restrict__ work, uint32_t n_pad,
                                                         uint32_t* __restrict__ sort_key,
                                                         uint32_t* __restrict__ sort_val,
                                                         const uint32_t* __restrict__ entry_slot, int entry_valid) {
  const uint32_t n = work->recv_off[ps.world];
  const PeerWindow& me = ps.w[ps.rank];
  uint32_t dropped = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
    uint32_t slot = t.slots, e = 0;
    if (i < n) {
      uint32_t s = 0;
      for (uint32_t k = 1; k < ps.world; k++) s += work->recv_off[k] <= i ? 1u : 0u;
      e = s * ps.region + (i - work->recv_off[s]);
      uint32_t f;
      if (entry_valid && ld_window(me.recv_reuse + s)) {  // the entries of the last forward pass, re-sent
        f = __ldg(entry_slot + e);
      } else {
        const uint64_t key = ld_window(me.recv_keys + e);
        f = key_valid(key) ? probe_find<kReadOnly>(t, key) : kNil;
      }
      if (f == kNil)
        dropped++;
      else
        slot = f;
    }
    sort_key[i] = slot;
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
  return dedup_bytes(t, n, true) + Workspace::pad(n * 8) + 256 +
         SegWork::bytes((uint64_t)p->world * p->region, t->v.dim, bits_for(t->v.slots)) + 4096;
}

void destroy_peer(meepo_table* t) {
  PeerState* p = t->peer;
  if (!p) return;
  for (uint32_t j = 0; j < kMaxPeers; j++)
    if (p->opened[j]) cudaIpcCloseMemHandle(p->opened[j]);
  cudaFree(p->window);
  cudaFree(p->local);
  delete p;
  t->peer = nullptr;
}

static meepo_status barrier(meepo_table* t, const uint32_t* send_cnt, bool for_apply, cudaStream_t stream) {
  PeerState* p = t->peer;
  ProfScope ps(t, "sharded.barrier", stream);
  p->seq++;
  peer_barrier_kernel<<<1, 32, 0, stream>>>(p->ps, p->seq, send_cnt, send_cnt ? p->work : nullptr, for_apply ? 1 : 0,
                                            t->v.counters, send_cnt && !for_apply ? p->fwd_send_cnt : nullptr,
                                            send_cnt && for_apply ? p->reuse_flag : nullptr, p->timeout_ns);
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
  uint64_t* ukeys = p->ukeys;
  uint32_t* inverse = p->inverse;
  uint32_t* uocc = t->ws.take<uint32_t>(std::max<uint64_t>(n, 1));  // hit/miss stats and LFU scores count occurrences
  uint64_t* n_unique = p->n_unique;
  if (n) MEEPO_CUDA_TRY(cudaMemcpyAsync(p->fwd_keys, keys, n * 8, cudaMemcpyDeviceToDevice, stream));
  const uint64_t n_window = (uint64_t)p->world * p->region;
  NewList nl{t->ws.take<uint32_t>(n_window)};  // one cell per window entry
  if (insert) MEEPO_CUDA_TRY(cudaMemsetAsync(nl.slots, 0xFF, n_window * 4, stream));
  MEEPO_TRY(dedup_run(t, keys, nullptr, n, DedupOut{ukeys, nullptr, inverse, n_unique, uocc}, stream));
  {
    ProfScope ps(t, "sharded.push_keys", stream);
    MEEPO_CUDA_TRY(cudaMemsetAsync(p->send_cnt, 0, kMaxPeers * 4, stream));
    const int grid = grid_for(t, (const void*)push_keys_kernel, 256, 0, (std::max<uint64_t>(n, 1) + 255) / 256);
    push_keys_kernel<<<grid, 256, 0, stream>>>(p->ps, ukeys, uocc, (const unsigned long long*)n_unique, p->send_cnt,
                                               p->loc);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  MEEPO_TRY(barrier(t, p->send_cnt, false, stream));
  {
    ProfScope ps(t, insert ? "sharded.owner_find_or_insert" : "sharded.owner_lookup", stream);
    const void* kern = insert ? pick_owner_kernel<true>(t->v.cpr) : pick_owner_kernel<false>(t->v.cpr);
    const uint64_t tiles = ((uint64_t)p->world * p->region + 31) / 32 + p->world;
    const int grid = grid_for(t, kern, 256, 0, (tiles + 7) / 8);
    const PeerWork* work = p->work;
    void* args[] = {&t->v, &p->ps, &work, &nl, &p->entry_slot};
    MEEPO_CUDA_TRY(cudaLaunchKernel(kern, dim3(grid), dim3(256), args, 0, stream));
  }
  if (insert) {
    ProfScope ps(t, "find_or_insert.publish", stream);
    MEEPO_TRY(publish_slots(t, nl.slots, n_window, stream));
  }
  MEEPO_TRY(barrier(t, nullptr, false, stream));
  if (n) {
    ProfScope ps(t, "sharded.expand", stream);
    const PeerWindow& me = p->ps.w[p->rank];
    const uint64_t tiles = (n + 31) / 32;
    const int grid = grid_for(t, (const void*)expand_kernel, 256, 0, (tiles + 7) / 8);
    expand_kernel<<<grid, 256, 0, stream>>>(me.ret_rows, me.ret_status, p->loc, inverse, (uint32_t)n, t->v.cpr,
                                            reinterpret_cast<uint4*>(rows_out), status_out);
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
  uint64_t* ukeys = p->ukeys;
  uint4** row_ptrs = t->ws.take<uint4*>(std::max<uint64_t>(n, 1));
  uint64_t* n_unique = p->n_unique;
  const uint64_t n_pad = (uint64_t)p->world * p->region;
  SegWork ow;  // owner side
  ow.take(t->ws, n_pad, t->v.dim, bits_for(t->v.slots));
  SegWork sw;  // sender side
  if (n) sw.take(t->ws, n, t->v.dim, bits_for((uint32_t)n));
  // the batch of the last forward pass? decided on the device; only the first backward pass after it qualifies
  const bool may_reuse = p->fwd_valid && n == p->fwd_n && n > 0 && !getenv("MEEPO_PEER_NO_REUSE");
  p->fwd_valid = false;
  {
    ProfScope ps(t, "sharded.same_batch", stream);
    MEEPO_CUDA_TRY(cudaMemsetAsync(p->reuse_flag, may_reuse ? 1 : 0, 4, stream));
    if (may_reuse) {
      const int grid = grid_for(t, (const void*)same_batch_kernel, 256, 0, (n + 255) / 256);
      same_batch_kernel<<<grid, 256, 0, stream>>>(keys, p->fwd_keys, (uint32_t)n, p->reuse_flag);
    }
  }
  DedupOut dd{ukeys, nullptr, p->inverse, n_unique, nullptr, reinterpret_cast<void* const*>(row_ptrs)};
  SegWork unused;
  MEEPO_TRY(dedup_hash(t, keys, n, dd, false, unused, stream, p->reuse_flag));  // no-op when the flag is set
  {
    ProfScope ps(t, "sharded.assign_grad_rows", stream);
    MEEPO_CUDA_TRY(cudaMemsetAsync(p->send_cnt, 0, kMaxPeers * 4, stream));
    const int grid = grid_for(t, (const void*)assign_grad_rows_kernel, 256, 0, (std::max<uint64_t>(n, 1) + 255) / 256);
    if (n) fill_sort_kernel<<<grid, 256, 0, stream>>>(p->inverse, (uint32_t)n, sw.sk_in, sw.sv_in);
    assign_grad_rows_kernel<<<grid, 256, 0, stream>>>(p->ps, ukeys, (const unsigned long long*)n_unique, p->send_cnt,
                                                      row_ptrs, p->trash_row, p->reuse_flag, p->loc, p->fwd_send_cnt);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  MEEPO_TRY(dedup_reduce(t, sw, grads, n, dd, stream));  // summed rows land in the owners' windows
  MEEPO_TRY(barrier(t, p->send_cnt, true, stream));
  {
    ProfScope ps(t, "sharded.owner_slots", stream);
    const int grid = grid_for(t, (const void*)recv_slots_kernel, 256, 0, (n_pad + 255) / 256);
    const int entry_valid = p->entry_gen == t->slot_gen ? 1 : 0;  // did anything move slots since the forward pass?
    recv_slots_kernel<<<grid, 256, 0, stream>>>(t->v, p->ps, p->work, (uint32_t)n_pad, ow.sk_in, ow.sv_in, p->entry_slot,
                                                entry_valid);
    MEEPO_CUDA_TRY(cudaGetLastError());
  }
  static const char* const names[5] = {"sharded.owner_sort", "sharded.owner_segments",
                                       "sharded.owner_apply", "sharded.owner_long_leaves", "sharded.owner_long_finish"};
  MEEPO_TRY(run_segmented(t, ow, t->v.slots, p->ps.w[p->rank].recv_grads, t->v.opt, nullptr, stream, nullptr, names));
  return barrier(t, nullptr, false, stream);
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
  self.err = t->err_word;
  carve_window(p->window, p->world, p->region, t->v.cpr, self.w[0], nullptr);
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
                                          uint64_t region_keys, void* blob_out) {
  if (!t || !blob_out) return fail(MEEPO_EINVAL, "null argument");
  if (world == 0 || world > kMaxPeers || rank >= world) return fail(MEEPO_EINVAL, "need rank < world <= MEEPO_MAX_PEERS");
  if (max_batch == 0 || max_batch > (1ull << 30)) return fail(MEEPO_EINVAL, "bad max_batch (1 .. 2^30)");
  if (region_keys == 0 || region_keys > max_batch) region_keys = max_batch;
  if ((uint64_t)world * region_keys > 0x7FFFFFFFull) return fail(MEEPO_EINVAL, "world * region_keys must fit in 31 bits");
  if (t->peer) return fail(MEEPO_EINVAL, "meepo_peer_prepare was already called on this table");
  DeviceGuard guard(t->device);
  PeerState* p = new PeerState();
  t->peer = p;
  p->world = world;
  p->rank = rank;
  p->max_batch = max_batch;
  p->region = region_keys;
  if (const char* e = getenv("MEEPO_PEER_TIMEOUT_MS")) p->timeout_ns = strtoull(e, nullptr, 10) * 1000000ull;
  PeerWindow w;
  carve_window(nullptr, world, region_keys, t->v.cpr, w, &p->window_bytes);
  auto bail = [&](cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorString(e);
    destroy_peer(t);
    return fail(e == cudaErrorMemoryAllocation ? MEEPO_ENOMEM : MEEPO_ECUDA, m);
  };
  cudaError_t e;
  if ((e = cudaMalloc(&p->window, p->window_bytes)) != cudaSuccess) return bail(e, "cudaMalloc(exchange window)");
  if ((e = cudaMemset(p->window, 0, kWindowHeader)) != cudaSuccess) return bail(e, "cudaMemset");
  // private scratch
  size_t off = 0;
  const size_t o_cnt = off;
  off += align_up(kMaxPeers * 4);
  const size_t o_work = off;
  off += align_up(sizeof(PeerWork));
  const size_t o_trash = off;
  off += align_up((size_t)t->v.cpr * 16);
  const size_t o_fcnt = off;
  off += align_up(kMaxPeers * 4);
  const size_t o_flag = off;
  off += align_up(4);
  const size_t o_nu = off;
  off += align_up(8);
  const size_t o_loc = off;
  off += align_up(max_batch * 4);
  const size_t o_inv = off;
  off += align_up(max_batch * 4);
  const size_t o_ukeys = off;
  off += align_up(max_batch * 8);
  const size_t o_fkeys = off;
  off += align_up(max_batch * 8);
  const size_t o_eslot = off;
  off += align_up((size_t)world * region_keys * 4);
  if ((e = cudaMalloc(&p->local, off)) != cudaSuccess) return bail(e, "cudaMalloc(peer scratch)");
  if ((e = cudaMemset(p->local, 0, o_loc)) != cudaSuccess) return bail(e, "cudaMemset");
  p->fwd_send_cnt = reinterpret_cast<uint32_t*>(p->local + o_fcnt);
  p->reuse_flag = reinterpret_cast<uint32_t*>(p->local + o_flag);
  p->n_unique = reinterpret_cast<uint64_t*>(p->local + o_nu);
  p->inverse = reinterpret_cast<uint32_t*>(p->local + o_inv);
  p->ukeys = reinterpret_cast<uint64_t*>(p->local + o_ukeys);
  p->fwd_keys = reinterpret_cast<uint64_t*>(p->local + o_fkeys);
  p->entry_slot = reinterpret_cast<uint32_t*>(p->local + o_eslot);
  p->send_cnt = reinterpret_cast<uint32_t*>(p->local + o_cnt);
  p->work = reinterpret_cast<PeerWork*>(p->local + o_work);
  p->trash_row = reinterpret_cast<uint4*>(p->local + o_trash);
  p->loc = reinterpret_cast<uint32_t*>(p->local + o_loc);
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
  if ((e = cudaIpcGetMemHandle(&b.handle, p->window)) != cudaSuccess) return bail(e, "cudaIpcGetMemHandle");
  if ((e = cudaDeviceSynchronize()) != cudaSuccess) return bail(e, "cudaDeviceSynchronize");
  memset(blob_out, 0, MEEPO_PEER_BLOB_BYTES);
  memcpy(blob_out, &b, sizeof b);
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
  p->ps.err = t->err_word;
  for (uint32_t j = 0; j < p->world; j++) {
    PeerBlob b;
    memcpy(&b, (const char*)blobs + (size_t)j * MEEPO_PEER_BLOB_BYTES, sizeof b);
    if (b.magic != kBlobMagic || b.abi != MEEPO_ABI_VERSION) return fail(MEEPO_EINVAL, "peer blob: bad magic / ABI");
    if (b.rank != j || b.world != p->world) return fail(MEEPO_EINVAL, "peer blobs must be ordered by rank");
    if (b.region != p->region || b.dim != t->v.dim || b.dtype != (uint32_t)t->v.dtype || b.opt != (uint32_t)t->v.opt ||
        b.flags != t->cfg.flags || b.window_bytes != p->window_bytes)
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
    carve_window(base, p->world, p->region, t->v.cpr, p->ps.w[j], nullptr);
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
