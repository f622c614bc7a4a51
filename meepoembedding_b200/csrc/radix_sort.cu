// radix_sort.cu — stable LSD radix sort of (u32 key, u32 value) pairs on the low `end_bit` bits of the
// key: the grouping step of apply_gradients (key = slot, value = batch index), of the sender-side
// pre-reduction (key = unique id) and of the owner side of the sharded backward pass.
//
// 8-bit digits, one scatter pass per digit, "onesweep" shape: ONE histogram kernel counts every
// digit of every pass up front, a tiny kernel turns the counts into per-pass bin bases, and each
// pass is a single kernel whose CTAs (ranges of 8 x 4096 pairs, taken in order from a device counter)
// learn their offset inside every bin from the ranges before them by decoupled look-back over a
// (range, bin) array of flagged counts — no per-pass histogram / scan launches.
//   per tile: warp-striped 16 keys per thread; stable rank of every key among the tile's keys with
//   the same digit by warp match (one ballot per digit bit) + per-warp digit counters in shared memory;
//   the tile is reordered by digit through shared memory so that every digit's run is stored to
//   global memory as one contiguous, coalesced piece.
// Arrays of a few million pairs stay in the 126 MB L2 between passes. The spin of the look-back is
// bounded: a tile that never shows up raises a flag (and the sort result is wrong) instead of hanging.
#include "table.h"

namespace meepo {

namespace {
constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 pairs
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsMaxSub = 8;  // sub-tiles per CTA (chosen per sort so that the CTAs of a pass form one wave)
constexpr uint32_t kFlagPartial = 1u << 30, kFlagInclusive = 2u << 30, kFlagMask = 3u << 30, kCountMask = ~kFlagMask;

struct RsTemp {
  uint32_t* hist;      // [passes][256] digit counts, then (in place) exclusive bin bases
  uint32_t* counters;  // [passes] next tile to hand out; [4] blocks of the histogram kernel that are done
  uint32_t* lookback;  // [passes][tiles][256]
  uint32_t* k_tmp;     // [n]
  uint32_t* v_tmp;     // [n]
};

// Digit counts of every pass in one read of the keys (16-byte loads). The last block to finish turns
// the counts into exclusive bin bases in place, which saves a launch.
__global__ void __launch_bounds__(256) rs_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n,
                                                      const uint32_t* __restrict__ n_dev, int passes, int end_bit,
                                                      uint32_t* __restrict__ hist, uint32_t* __restrict__ blocks_done) {
  __shared__ uint32_t sh[4][256];
  if (n_dev) n = min(n, __ldg(n_dev));  // the host only knows an upper bound
  __shared__ uint32_t warp_tot[8];
  __shared__ bool last;
  for (int p = 0; p < passes; p++) sh[p][threadIdx.x] = 0;
  __syncthreads();
  uint32_t mask[4];
#pragma unroll
  for (int p = 0; p < 4; p++) mask[p] = p < passes ? (1u << min(8, end_bit - 8 * p)) - 1u : 0u;
  auto count = [&](uint32_t k) {
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < passes) atomicAdd(&sh[p][(k >> (8 * p)) & mask[p]], 1u);
  };
  const uint32_t n4 = ((reinterpret_cast<uintptr_t>(keys) & 15u) == 0) ? n / 4 : 0;
  const uint4* k4 = reinterpret_cast<const uint4*>(keys);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const uint4 v = __ldg(k4 + i);
    count(v.x);
    count(v.y);
    count(v.z);
    count(v.w);
  }
  for (uint32_t i = n4 * 4 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) count(__ldg(keys + i));
  __syncthreads();
  for (int p = 0; p < passes; p++)
    if (sh[p][threadIdx.x]) atomicAdd(hist + p * 256 + threadIdx.x, sh[p][threadIdx.x]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  // counts -> exclusive bin bases, in place; one warp-shuffle scan per pass
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  for (int p = 0; p < passes; p++) {
    const uint32_t v = __ldcg(hist + p * 256 + threadIdx.x);
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
      if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t j = 0; j < w; j++) before += warp_tot[j];
    hist[p * 256 + threadIdx.x] = before + x - v;
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One pass. A CTA owns nsub consecutive sub-tiles of 4096 pairs (ranges handed out in order); nsub is
// chosen so that all CTAs of the pass are resident at once. A CTA first counts the digits of its whole
// range (one cheap read of the keys), publishes that count at once and looks back for the ranges
// before it; only then does it rank and scatter its sub-tiles one after the other, advancing the
// per-digit cursor. (Publishing after the ranking, one entry per 4096 pairs and a second wave of
// CTAs, made waiting for and walking over not-yet-inclusive predecessors 45% of the kernel — ncu.)
__global__ void __launch_bounds__(kRsThreads, 4) rs_onesweep_kernel(const uint32_t* __restrict__ k_in,
                                                                    const uint32_t* __restrict__ v_in,
                                                                    uint32_t* __restrict__ k_out,
                                                                    uint32_t* __restrict__ v_out, uint32_t n,
                                                                    const uint32_t* __restrict__ n_dev, int shift,
                                                                    int bits, uint32_t nsub,
                                                                    const uint32_t* __restrict__ bin_base,
                                                                    uint32_t* __restrict__ tile_counter,
                                                                    uint32_t* __restrict__ lookback,
                                                                    uint32_t* __restrict__ error) {
  __shared__ uint32_t s_keys[kRsTile];
  __shared__ uint32_t s_vals[kRsTile];
  __shared__ uint32_t s_cnt[kRsWarps][256];  // per-warp digit counts, then exclusive offsets over the warps
  __shared__ uint32_t s_bin_start[256];      // start of every digit's run in the reordered sub-tile
  __shared__ uint32_t s_global[256];         // global index where the next key of every digit goes
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_warp_tot[kRsWarps];
  __shared__ uint32_t s_tile;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
  const uint32_t dmask = (1u << bits) - 1u;
  if (n_dev) n = min(n, __ldg(n_dev));
  // ranges are handed out in order, so the CTAs past the end (grid sized for the host's upper bound) are the
  // ones nobody waits for: they leave at once
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  s_hist[tid] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t range0 = tile * (nsub * kRsTile);
  if (range0 >= n) return;
  const uint32_t range_n = min(nsub * kRsTile, n - range0);

  // A. digit counts of the whole range -> publish -> look back -> global cursor of every digit
  for (uint32_t e0 = 0; e0 < range_n; e0 += kRsThreads * 8) {
    uint32_t k8[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint32_t e = e0 + u * kRsThreads + tid;
      k8[u] = e < range_n ? __ldg(k_in + range0 + e) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; u++)
      if (e0 + u * kRsThreads + tid < range_n) atomicAdd(&s_hist[(k8[u] >> shift) & dmask], 1u);
  }
  __syncthreads();
  {
    const uint32_t tile_cnt = s_hist[tid];
    uint32_t* lb = lookback + (size_t)tile * 256 + tid;
    st_volatile_u32(lb, tile_cnt | (tile == 0 ? kFlagInclusive : kFlagPartial));
    uint32_t excl = 0;
    if (tile > 0) {
      // Walk back over the earlier ranges until one that already knows its inclusive prefix; the loads of a
      // window of 8 predecessors are issued together.
      uint32_t prev = tile;  // ranges [0, prev) are still to be accounted for
      bool done = false;
      while (!done && prev > 0) {
        constexpr uint32_t W = 8;
        uint32_t v[W];
#pragma unroll
        for (uint32_t u = 0; u < W; u++)
          v[u] = u < prev ? ld_volatile_u32(lookback + (size_t)(prev - 1 - u) * 256 + tid) : kFlagInclusive;
#pragma unroll
        for (uint32_t u = 0; u < W; u++) {
          if (done || u >= prev) continue;
          uint32_t x = v[u];
          const uint32_t* q = lookback + (size_t)(prev - 1 - u) * 256 + tid;
          for (uint32_t spin = 0; (x & kFlagMask) == 0; spin++) {
            if (spin > (1u << 24)) {  // a range that never shows up: give up rather than hang
              *reinterpret_cast<volatile uint32_t*>(error) = 1u;
              x = kFlagInclusive;
              break;
            }
            __nanosleep(32);
            x = ld_volatile_u32(q);
          }
          excl += x & kCountMask;
          done = (x & kFlagInclusive) != 0;
        }
        prev -= min(prev, W);
      }
      st_volatile_u32(lb, (excl + tile_cnt) | kFlagInclusive);
    }
    s_global[tid] = __ldg(bin_base + tid) + excl;
  }

  // B. sub-tile by sub-tile: rank, reorder through shared memory, store every digit's run contiguously
  for (uint32_t sub = 0; sub * kRsTile < range_n; sub++) {
    const uint32_t base = range0 + sub * kRsTile;
    const uint32_t valid = min((uint32_t)kRsTile, range_n - sub * kRsTile);
#pragma unroll
    for (int j = 0; j < kRsWarps; j++) s_cnt[j][tid] = 0;
    __syncthreads();  // also orders s_global / the previous sub-tile's stores

    // 1. load (warp-striped: element e = w*512 + k*32 + lane) and rank within the warp
    uint32_t key[kRsItems], val[kRsItems], rank[kRsItems];
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
      const uint32_t e = w * (kRsItems * 32) + k * 32 + lane;
      key[k] = e < valid ? __ldg(k_in + base + e) : 0xFFFFFFFFu;
      val[k] = e < valid ? __ldg(v_in + base + e) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
      const uint32_t e = w * (kRsItems * 32) + k * 32 + lane;
      const bool ok = e < valid;
      const uint32_t d = ok ? (key[k] >> shift) & dmask : 0u;
      // lanes holding the same digit, one ballot per digit bit: the hardware match (__match_any_sync) walks
      // the distinct values of the warp one by one (~30 per instruction here)
      unsigned peers = __ballot_sync(0xFFFFFFFFu, ok);  // padding matches nothing
#pragma unroll
      for (int b = 0; b < 8; b++) {
        if (b < bits) {
          const unsigned m = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
          peers &= ((d >> b) & 1u) ? m : ~m;
        }
      }
      const int leader = ok ? __ffs(peers) - 1 : (int)lane;
      uint32_t before = 0;
      if (ok && (int)lane == leader) {
        before = s_cnt[w][d];
        s_cnt[w][d] = before + __popc(peers);
      }
      before = __shfl_sync(0xFFFFFFFFu, before, leader);
      rank[k] = before + __popc(peers & ((1u << lane) - 1u));  // among this warp's earlier keys with digit d
      __syncwarp();
    }
    __syncthreads();

    // 2. thread d: exclusive offsets of digit d over the warps, sub-tile count of digit d
    uint32_t sub_cnt = 0;
#pragma unroll
    for (int j = 0; j < kRsWarps; j++) {
      const uint32_t c = s_cnt[j][tid];
      s_cnt[j][tid] = sub_cnt;
      sub_cnt += c;
    }
    // 3. exclusive scan of the counts over the digits -> start of every run in the reordered sub-tile
    {
      uint32_t x = sub_cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= (uint32_t)d) x += y;
      }
      if (lane == 31) s_warp_tot[w] = x;
      __syncthreads();
      uint32_t before = 0;
      for (uint32_t j = 0; j < w; j++) before += s_warp_tot[j];
      s_bin_start[tid] = before + x - sub_cnt;
    }
    __syncthreads();

    // 4. reorder by digit through shared memory (stable: digit run, then warp, then rank within the warp)
#pragma unroll
    for (int k = 0; k < kRsItems; k++) {
      const uint32_t e = w * (kRsItems * 32) + k * 32 + lane;
      if (e < valid) {
        const uint32_t d = (key[k] >> shift) & dmask;
        const uint32_t pos = s_bin_start[d] + s_cnt[w][d] + rank[k];
        s_keys[pos] = key[k];
        s_vals[pos] = val[k];
      }
    }
    __syncthreads();
    // 5. every digit's run goes out as one contiguous piece
    for (uint32_t j = tid; j < valid; j += kRsThreads) {
      const uint32_t k = s_keys[j];
      const uint32_t d = (k >> shift) & dmask;
      const uint32_t dst = s_global[d] + (j - s_bin_start[d]);
      k_out[dst] = k;
      v_out[dst] = s_vals[j];
    }
    __syncthreads();
    s_global[tid] += sub_cnt;  // the next sub-tile continues every run
  }
}

// sub-tiles per CTA: as few as keep the pass within one wave of resident CTAs (4 per SM)
uint32_t sub_tiles(const meepo_table* t, uint64_t n) {
  const uint64_t fine = (n + kRsTile - 1) / kRsTile, wave = (uint64_t)t->num_sms * 4;
  return (uint32_t)std::min<uint64_t>(kRsMaxSub, std::max<uint64_t>(1, (fine + wave - 1) / wave));
}

RsTemp carve(char* temp, uint64_t n, int passes) {
  const uint64_t tiles = (n + kRsTile - 1) / kRsTile;  // sized for one entry per sub-tile (the upper bound)
  RsTemp r;
  size_t off = 0;
  r.hist = reinterpret_cast<uint32_t*>(temp + off);
  off += Workspace::pad(4 * 256 * 4);
  r.counters = reinterpret_cast<uint32_t*>(temp + off);
  off += Workspace::pad(8 * 4);
  r.lookback = reinterpret_cast<uint32_t*>(temp + off);
  off += Workspace::pad((size_t)passes * tiles * 256 * 4);
  r.k_tmp = reinterpret_cast<uint32_t*>(temp + off);
  off += Workspace::pad(n * 4);
  r.v_tmp = reinterpret_cast<uint32_t*>(temp + off);
  return r;
}
}  // namespace

bool radix_sort_supported(uint64_t n, int end_bit) {
  static const bool use_cub = getenv("MEEPO_SORT_CUB") != nullptr;
  return !use_cub && n > 0 && n < (1ull << 30) && end_bit >= 1 && end_bit <= 32;
}

size_t radix_sort_temp_bytes(uint64_t n, int end_bit) {
  const int passes = (end_bit + 7) / 8;
  const uint64_t tiles = (n + kRsTile - 1) / kRsTile;
  return Workspace::pad(4 * 256 * 4) + Workspace::pad(8 * 4) + Workspace::pad((size_t)passes * tiles * 256 * 4) +
         2 * Workspace::pad(n * 4) + 256;
}

// (k_in, v_in) -> (k_out, v_out), stable, on key bits [0, end_bit). temp: radix_sort_temp_bytes(), 256-byte aligned.
// n_dev (optional, device): the true number of pairs, <= n; grids and scratch are sized for n.
meepo_status radix_sort_pairs(meepo_table* t, char* temp, const uint32_t* k_in, uint32_t* k_out,
                              const uint32_t* v_in, uint32_t* v_out, uint32_t n, int end_bit, cudaStream_t stream,
                              const uint32_t* n_dev, uint32_t expect_n) {
  const int passes = (end_bit + 7) / 8;
  // sub-tiles per CTA from the number of pairs the caller expects (a device-side count may be far below n:
  // the CTAs past the end leave at once, the others should still fill the GPU)
  const uint32_t nsub = sub_tiles(t, expect_n ? std::min(expect_n, n) : n);
  const uint32_t tiles = (n + nsub * kRsTile - 1) / (nsub * kRsTile);
  const uint32_t lb_stride = (n + kRsTile - 1) / kRsTile;  // look-back entries reserved per pass
  RsTemp r = carve(temp, n, passes);
  // histogram + counters + look-back state are one contiguous zero-filled block
  MEEPO_CUDA_TRY(cudaMemsetAsync(r.hist, 0, (char*)r.k_tmp - (char*)r.hist, stream));
  const int hgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n + 1023) / 1024, (uint64_t)t->num_sms * 4));
  rs_hist_kernel<<<hgrid, 256, 0, stream>>>(k_in, n, n_dev, passes, end_bit, r.hist, r.counters + 4);
  // ping-pong so that the last pass lands in the caller's output arrays
  const uint32_t* src_k = k_in;
  const uint32_t* src_v = v_in;
  for (int p = 0; p < passes; p++) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    uint32_t* dst_k = to_out ? k_out : r.k_tmp;
    uint32_t* dst_v = to_out ? v_out : r.v_tmp;
    const int bits = std::min(8, end_bit - 8 * p);
    rs_onesweep_kernel<<<tiles, kRsThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, n_dev, 8 * p, bits, nsub,
                                                         r.hist + p * 256, r.counters + p,
                                                         r.lookback + (size_t)p * lb_stride * 256,
                                                         t->err_word + kErrLookback);
    src_k = dst_k;
    src_v = dst_v;
  }
  MEEPO_CUDA_TRY(cudaGetLastError());
  return MEEPO_OK;
}

}  // namespace meepo

using namespace meepo;

// Test hook (not part of include/meepo.h): sort device arrays, report the look-back error flag.
extern "C" MEEPO_API meepo_status meepo_internal_sort_pairs(meepo_table* t, const uint32_t* k_in, uint32_t* k_out,
                                                            const uint32_t* v_in, uint32_t* v_out, uint64_t n,
                                                            int32_t end_bit, void* stream_) {
  if (!t || !radix_sort_supported(n, end_bit)) return fail(MEEPO_EINVAL, "unsupported sort size / bit count");
  DeviceGuard guard(t->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  MEEPO_TRY(t->ws.reserve(radix_sort_temp_bytes(n, end_bit), stream));
  char* temp = t->ws.take<char>(radix_sort_temp_bytes(n, end_bit));
  MEEPO_TRY(radix_sort_pairs(t, temp, k_in, k_out, v_in, v_out, (uint32_t)n, end_bit, stream, nullptr, 0));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(stream));
  return sticky_error(t);
}
