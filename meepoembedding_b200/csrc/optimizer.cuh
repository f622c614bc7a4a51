// optimizer.cuh — 16-byte-chunk views of rows / gradients and the fused sparse-optimizer step
// (include/meepo.h "Update": every operation individually rounded, no fused multiply-add). Shared by
// the single-table update path (update.cu) and the owner side of the sharded path (peer.cu).
#pragma once
#include "common.cuh"

namespace meepo {

// ---------------------------------------------------------------------------------------------
// gradient chunk helpers. E = fp32 elements per 16-byte chunk (4 for fp32 tables, 8 for bf16).
template <bool BF16>
struct Chunk {
  static constexpr int E = BF16 ? 8 : 4;
};

template <bool BF16>
__device__ __forceinline__ void widen(const uint4& raw, float (&g)[Chunk<BF16>::E]) {
  if constexpr (BF16) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      g[2 * i] = __uint_as_float(w[i] << 16);
      g[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  } else {
    g[0] = __uint_as_float(raw.x);
    g[1] = __uint_as_float(raw.y);
    g[2] = __uint_as_float(raw.z);
    g[3] = __uint_as_float(raw.w);
  }
}

constexpr int kStoreOnly = 15;  // "optimizer" of meepo_reduce_duplicates: round + store the sum (not a meepo_opt)

template <bool BF16>
__device__ __forceinline__ uint4 narrow(const float (&w)[Chunk<BF16>::E]) {
  if constexpr (BF16) {
    return make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]), pack_bf16x2(w[4], w[5]),
                      pack_bf16x2(w[6], w[7]));
  } else {
    return make_uint4(__float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]),
                      __float_as_uint(w[3]));
  }
}

// One optimizer step on chunk q of the row in `slot` (meepo.h "Update"; every op rounded once),
// split into the loads (opt_issue) and the math + stores (opt_finish) so that a caller can put the
// loads of several segments in flight before it consumes the first.
template <bool BF16, int OPT>
struct OptIn {
  static constexpr int SQ = Chunk<BF16>::E / 4;  // state uint4s per chunk
  static constexpr int NS = OPT == MEEPO_ADAGRAD ? SQ : (OPT == MEEPO_ADAM ? 2 * SQ : 1);  // row-wise Adagrad: 1
  uint4 row;
  uint4 st[NS];
};

template <bool BF16, int OPT>
__device__ __forceinline__ void opt_issue(const TableView& t, uint32_t slot, uint32_t q, OptIn<BF16, OPT>& in) {
  constexpr int SQ = OptIn<BF16, OPT>::SQ;
  if constexpr (OPT == kStoreOnly) return;
  in.row = ld_stream(t.rows + (size_t)slot * t.cpr + q);
  if constexpr (OPT == MEEPO_ADAGRAD) {
    const uint4* sp = t.state + (size_t)slot * t.scpr + (size_t)q * SQ;
#pragma unroll
    for (int k = 0; k < SQ; k++) in.st[k] = ld_stream(sp + k);
  } else if constexpr (OPT == MEEPO_ADAM) {
    const uint4* mp = t.state + (size_t)slot * t.scpr + (size_t)q * SQ;
    const uint4* vp = mp + (size_t)t.cpr * SQ;
#pragma unroll
    for (int k = 0; k < SQ; k++) {
      in.st[k] = ld_stream(mp + k);
      in.st[SQ + k] = ld_stream(vp + k);
    }
  } else if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) {
    in.st[0] = ld_stream(t.state + (size_t)slot);  // the row's one accumulator: every lane of the group reads it
  }
}

// meepo.h "ADAGRAD_ROWWISE": c_q of this lane's chunk, and the reduction of the chunk sums of a group of GL lanes
// that hold chunks 0..GL-1 (GL a power of two <= 32 = the whole row): the halving tree, as a butterfly.
template <bool BF16>
__device__ __forceinline__ float chunk_sumsq(const float (&g)[Chunk<BF16>::E]) {
  float c = __fmul_rn(g[0], g[0]);
#pragma unroll
  for (int e = 1; e < Chunk<BF16>::E; e++) c = __fadd_rn(c, __fmul_rn(g[e], g[e]));
  return c;
}
__device__ __forceinline__ float group_tree_sum(float v, uint32_t GL, unsigned gmask) {
  for (uint32_t d = GL >> 1; d >= 1; d >>= 1) v = __fadd_rn(v, __shfl_xor_sync(gmask, v, d));
  return v;
}

__device__ __forceinline__ void unpack4(const uint4& raw, float* f) {
  f[0] = __uint_as_float(raw.x);
  f[1] = __uint_as_float(raw.y);
  f[2] = __uint_as_float(raw.z);
  f[3] = __uint_as_float(raw.w);
}
__device__ __forceinline__ uint4 pack4(const float* f) {
  return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}

// alpha: Adam's scalar step size; row-wise Adagrad: the mean square of the row's gradient (s_0 / dim).
template <bool BF16, int OPT>
__device__ __forceinline__ void opt_finish(const TableView& t, uint32_t slot, uint32_t q,
                                           const OptIn<BF16, OPT>& in, const float (&g)[Chunk<BF16>::E],
                                           float alpha, uint4* reduce_row) {
  constexpr int E = Chunk<BF16>::E;
  constexpr int SQ = OptIn<BF16, OPT>::SQ;
  if constexpr (OPT == kStoreOnly) {
    st_stream(reduce_row + q, narrow<BF16>(g));  // kStoreOnly: reduce_row = output row of this sort key
    return;
  }
  uint4* rowp = t.rows + (size_t)slot * t.cpr + q;
  float w[E];
  widen<BF16>(in.row, w);
  if constexpr (OPT == MEEPO_SGD) {
#pragma unroll
    for (int e = 0; e < E; e++) w[e] = __fsub_rn(w[e], __fmul_rn(t.lr, g[e]));
  } else if constexpr (OPT == MEEPO_ADAGRAD) {
    uint4* sp = t.state + (size_t)slot * t.scpr + (size_t)q * SQ;
    float a[E];
#pragma unroll
    for (int k = 0; k < SQ; k++) unpack4(in.st[k], a + 4 * k);
#pragma unroll
    for (int e = 0; e < E; e++) {
      a[e] = __fadd_rn(a[e], __fmul_rn(g[e], g[e]));
      const float den = __fadd_rn(__fsqrt_rn(a[e]), t.eps);
      w[e] = __fsub_rn(w[e], __fdiv_rn(__fmul_rn(t.lr, g[e]), den));
    }
#pragma unroll
    for (int k = 0; k < SQ; k++) st_stream(sp + k, pack4(a + 4 * k));
  } else if constexpr (OPT == MEEPO_ADAGRAD_ROWWISE) {
    const float a = __fadd_rn(__uint_as_float(in.st[0].x), alpha);
    const float den = __fadd_rn(__fsqrt_rn(a), t.eps);
#pragma unroll
    for (int e = 0; e < E; e++) w[e] = __fsub_rn(w[e], __fdiv_rn(__fmul_rn(t.lr, g[e]), den));
    if (q == 0) st_stream(t.state + (size_t)slot, make_uint4(__float_as_uint(a), 0u, 0u, 0u));
  } else {
    uint4* mp = t.state + (size_t)slot * t.scpr + (size_t)q * SQ;
    uint4* vp = mp + (size_t)t.cpr * SQ;
    float m[E], v[E];
#pragma unroll
    for (int k = 0; k < SQ; k++) {
      unpack4(in.st[k], m + 4 * k);
      unpack4(in.st[SQ + k], v + 4 * k);
    }
    const float omb1 = __fsub_rn(1.0f, t.beta1), omb2 = __fsub_rn(1.0f, t.beta2);
#pragma unroll
    for (int e = 0; e < E; e++) {
      m[e] = __fadd_rn(__fmul_rn(t.beta1, m[e]), __fmul_rn(omb1, g[e]));
      v[e] = __fadd_rn(__fmul_rn(t.beta2, v[e]), __fmul_rn(omb2, __fmul_rn(g[e], g[e])));
      const float den = __fadd_rn(__fsqrt_rn(v[e]), t.eps);
      w[e] = __fsub_rn(w[e], __fdiv_rn(__fmul_rn(alpha, m[e]), den));
    }
#pragma unroll
    for (int k = 0; k < SQ; k++) {
      st_stream(mp + k, pack4(m + 4 * k));
      st_stream(vp + k, pack4(v + 4 * k));
    }
  }
  st_stream(rowp, narrow<BF16>(w));
}

template <bool BF16, int OPT>
__device__ __forceinline__ void optimizer_chunk(const TableView& t, uint32_t slot, uint32_t q,
                                                const float (&g)[Chunk<BF16>::E], float alpha,
                                                uint4* reduce_row) {
  OptIn<BF16, OPT> in;
  opt_issue<BF16, OPT>(t, slot, q, in);
  opt_finish<BF16, OPT>(t, slot, q, in, g, alpha, reduce_row);
}

// Adam: per-row step count -> scalar step size (double math, rounded once). Every lane of the
// group reads the old count, the group syncs, lane 0 writes the new one.
template <int OPT>
__device__ __forceinline__ float adam_alpha(const TableView& t, uint32_t slot, unsigned gmask, bool leader) {
  if constexpr (OPT != MEEPO_ADAM) return 0.0f;
  const uint32_t tt = t.steps[slot] + 1;
  __syncwarp(gmask);
  if (leader) t.steps[slot] = tt;
  const double bc1 = 1.0 - pow((double)t.beta1, (double)tt);
  const double bc2 = 1.0 - pow((double)t.beta2, (double)tt);
  return (float)((double)t.lr * sqrt(bc2) / bc1);
}

}  // namespace meepo
