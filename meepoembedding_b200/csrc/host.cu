// host.cu — host-buffer front ends (meepo_*_host): the C-ABI call a CPU-side caller makes.
//
// The caller's buffers may be pinned or pageable (cudaMemcpyAsync handles both; only pinned memory
// overlaps). Keys (8 B/key) go up in one copy; rows come back / gradients go up in chunks on their
// own copy streams so PCIe runs in both directions while the kernels work:
//   find_or_insert_host / lookup_host: kernel(chunk c) overlaps D2H(rows of chunk c-1)
//   apply_gradients_host:              probe + sort + segment passes overlap H2D(gradients)
// Chunks of one find_or_insert share one status epoch (probe_gather_begin/end), so chunking is
// invisible in the results.
#include "table.h"

struct HostPipe {
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t keys_up = nullptr, grads_up = nullptr, k_done[2] = {nullptr, nullptr},
              out_done[2] = {nullptr, nullptr};
  uint64_t* d_keys = nullptr;
  uint8_t* d_status = nullptr;
  size_t keys_cap = 0;
  char* d_rows[2] = {nullptr, nullptr};
  size_t rows_chunk_bytes = 0;
  char* d_grads = nullptr;
  size_t grads_cap = 0;
};

namespace meepo {

constexpr size_t kRowsChunkBytes = 64ull << 20;

static meepo_status get_pipe(meepo_table* t, HostPipe** out) {
  if (!t->pipe) {
    HostPipe* p = new HostPipe();
    t->pipe = p;
    MEEPO_CUDA_TRY(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    MEEPO_CUDA_TRY(cudaStreamCreateWithFlags(&p->s_k, cudaStreamNonBlocking));
    MEEPO_CUDA_TRY(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    MEEPO_CUDA_TRY(cudaEventCreateWithFlags(&p->keys_up, cudaEventDisableTiming));
    MEEPO_CUDA_TRY(cudaEventCreateWithFlags(&p->grads_up, cudaEventDisableTiming));
    for (int i = 0; i < 2; i++) {
      MEEPO_CUDA_TRY(cudaEventCreateWithFlags(&p->k_done[i], cudaEventDisableTiming));
      MEEPO_CUDA_TRY(cudaEventCreateWithFlags(&p->out_done[i], cudaEventDisableTiming));
      MEEPO_CUDA_TRY(cudaMalloc(&p->d_rows[i], kRowsChunkBytes));
    }
    p->rows_chunk_bytes = kRowsChunkBytes;
  }
  *out = t->pipe;
  return MEEPO_OK;
}

static meepo_status ensure_keys(HostPipe* p, uint64_t n) {
  if (n <= p->keys_cap) return MEEPO_OK;
  MEEPO_CUDA_TRY(cudaDeviceSynchronize());
  if (p->d_keys) cudaFree(p->d_keys);
  if (p->d_status) cudaFree(p->d_status);
  p->d_keys = nullptr;
  p->d_status = nullptr;
  p->keys_cap = 0;
  const size_t cap = n + n / 4 + 1024;
  MEEPO_CUDA_TRY(cudaMalloc(&p->d_keys, cap * 8));
  MEEPO_CUDA_TRY(cudaMalloc(&p->d_status, cap));
  p->keys_cap = cap;
  return MEEPO_OK;
}

void destroy_host_pipe(meepo_table* t) {
  HostPipe* p = t->pipe;
  if (!p) return;
  cudaFree(p->d_keys);
  cudaFree(p->d_status);
  cudaFree(p->d_grads);
  for (int i = 0; i < 2; i++) {
    cudaFree(p->d_rows[i]);
    if (p->k_done[i]) cudaEventDestroy(p->k_done[i]);
    if (p->out_done[i]) cudaEventDestroy(p->out_done[i]);
  }
  if (p->keys_up) cudaEventDestroy(p->keys_up);
  if (p->grads_up) cudaEventDestroy(p->grads_up);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_k) cudaStreamDestroy(p->s_k);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
  t->pipe = nullptr;
}

static meepo_status probe_host(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                               uint8_t* status_out, bool insert) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!keys || !rows_out)) return fail(MEEPO_EINVAL, "null buffer");
  DeviceGuard guard(t->device);
  MEEPO_CUDA_TRY(cudaDeviceSynchronize());  // order after whatever the caller queued before
  HostPipe* p;
  MEEPO_TRY(get_pipe(t, &p));
  MEEPO_TRY(ensure_keys(p, n));
  MEEPO_TRY(probe_gather_begin(t, n, insert, p->s_k));
  if (n == 0) return MEEPO_OK;
  MEEPO_CUDA_TRY(cudaMemcpyAsync(p->d_keys, keys, n * 8, cudaMemcpyHostToDevice, p->s_in));
  MEEPO_CUDA_TRY(cudaEventRecord(p->keys_up, p->s_in));
  MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_k, p->keys_up, 0));
  const uint64_t R = t->row_bytes;
  uint64_t chunk = p->rows_chunk_bytes / R;
  chunk = chunk / 32 * 32;
  if (chunk == 0) chunk = 32;
  uint64_t c = 0;
  for (uint64_t off = 0; off < n; off += chunk, c++) {
    const uint64_t m = std::min(chunk, n - off);
    const int b = (int)(c & 1);
    if (c >= 2) MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_k, p->out_done[b], 0));
    MEEPO_TRY(probe_gather_chunk(t, p->d_keys + off, m, p->d_rows[b], p->d_status + off, insert, p->s_k));
    MEEPO_CUDA_TRY(cudaEventRecord(p->k_done[b], p->s_k));
    MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_out, p->k_done[b], 0));
    MEEPO_CUDA_TRY(cudaMemcpyAsync((char*)rows_out + off * R, p->d_rows[b], m * R, cudaMemcpyDeviceToHost, p->s_out));
    MEEPO_CUDA_TRY(cudaEventRecord(p->out_done[b], p->s_out));
  }
  MEEPO_TRY(probe_gather_end(t, n, insert, p->s_k));
  if (status_out) MEEPO_CUDA_TRY(cudaMemcpyAsync(status_out, p->d_status, n, cudaMemcpyDeviceToHost, p->s_k));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(p->s_k));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(p->s_out));
  return MEEPO_OK;
}

}  // namespace meepo

using namespace meepo;

extern "C" {

MEEPO_API meepo_status meepo_find_or_insert_host(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                 void* rows_out, uint8_t* status_out) {
  return probe_host(t, keys, n, rows_out, status_out, true);
}
MEEPO_API meepo_status meepo_lookup_host(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                         uint8_t* found_out) {
  return probe_host(t, keys, n, rows_out, found_out, false);
}

MEEPO_API meepo_status meepo_apply_gradients_host(meepo_table* t, const uint64_t* keys, const void* grads,
                                                  uint64_t n) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!keys || !grads)) return fail(MEEPO_EINVAL, "null buffer");
  if (n == 0) return MEEPO_OK;
  DeviceGuard guard(t->device);
  MEEPO_CUDA_TRY(cudaDeviceSynchronize());
  HostPipe* p;
  MEEPO_TRY(get_pipe(t, &p));
  MEEPO_TRY(ensure_keys(p, n));
  const size_t gbytes = (size_t)n * t->row_bytes;
  if (gbytes > p->grads_cap) {
    if (p->d_grads) cudaFree(p->d_grads);
    p->d_grads = nullptr;
    p->grads_cap = 0;
    MEEPO_CUDA_TRY(cudaMalloc(&p->d_grads, gbytes + gbytes / 8));
    p->grads_cap = gbytes + gbytes / 8;
  }
  MEEPO_CUDA_TRY(cudaMemcpyAsync(p->d_keys, keys, n * 8, cudaMemcpyHostToDevice, p->s_in));
  MEEPO_CUDA_TRY(cudaEventRecord(p->keys_up, p->s_in));
  MEEPO_CUDA_TRY(cudaMemcpyAsync(p->d_grads, grads, gbytes, cudaMemcpyHostToDevice, p->s_in));
  MEEPO_CUDA_TRY(cudaEventRecord(p->grads_up, p->s_in));
  MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_k, p->keys_up, 0));
  MEEPO_TRY(launch_apply_gradients(t, p->d_keys, p->d_grads, n, p->s_k, p->grads_up));
  MEEPO_CUDA_TRY(cudaStreamSynchronize(p->s_k));
  return MEEPO_OK;
}

}  // extern "C"
