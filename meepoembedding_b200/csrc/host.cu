// host.cu — host-buffer front ends (meepo_*_host and meepo_*_host_async): the C-ABI calls a CPU-side
// caller makes (include/meepo.h "host-buffer front ends").
//
// Every host verb is asynchronous underneath: it enqueues its copies and kernels on three private streams
// and hands back a ticket; meepo_wait(ticket) blocks until the verb's results are in the caller's buffers
// (find_or_insert / lookup) or its inputs have been consumed (apply_gradients). The synchronous verbs are
// "issue, then wait". Table work happens strictly in ISSUE ORDER (one kernel stream), whatever the copies
// overlap with; PCIe runs in both directions at once:
//   s_in   H2D copies: keys (8 B/key, one copy) and gradients (one copy per call)
//   s_k    every kernel of the table, in issue order
//   s_out  D2H copies: rows in chunks (kernel of chunk c overlaps the copy of chunk c-1), then the status bytes
// A training caller keeps two batches in flight — find_or_insert(i+1) is issued before apply_gradients(i) — so
// that the rows of batch i+1 come down while the gradients of batch i go up (bench.py's e2e loop does that).
// Staging buffers are double-buffered per verb kind; a third call of the same kind first waits (on the
// device, not the host) for the buffers of the call two before it. The caller's buffers must stay valid until
// the ticket has been waited for; pageable buffers work but do not overlap (cudaMemcpyAsync stages them).
// Chunks of one find_or_insert share one status epoch (probe_gather_begin/end), so chunking is invisible in
// the results.
#include "table.h"

namespace {
constexpr int kTicketRing = 32;
constexpr int kDepth = 2;  // staging slots per verb kind
}  // namespace

struct HostPipe {
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  // forward verbs: device copies of the keys / status bytes of a call, rows in two rotating chunk buffers
  struct Fwd {
    uint64_t* d_keys = nullptr;
    uint8_t* d_status = nullptr;
    size_t cap = 0;
    cudaEvent_t keys_up = nullptr, kernels_done = nullptr, free_ev = nullptr;
    bool used = false;
  } fwd[kDepth];
  char* d_rows[2] = {nullptr, nullptr};
  cudaEvent_t k_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
  bool rows_used[2] = {false, false};
  uint64_t chunk_seq = 0;
  size_t rows_chunk_bytes = 0;
  // backward verb
  struct Bwd {
    uint64_t* d_keys = nullptr;
    char* d_grads = nullptr;
    size_t keys_cap = 0, grads_cap = 0;
    cudaEvent_t keys_up = nullptr, grads_up = nullptr, free_ev = nullptr;
    bool used = false;
  } bwd[kDepth];
  uint64_t fwd_seq = 0, bwd_seq = 0;
  // tickets
  cudaEvent_t ticket_ev[kTicketRing] = {};
  uint64_t next_ticket = 1;  // 0 is "everything issued so far"
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
    auto ev = [](cudaEvent_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
    for (int i = 0; i < kDepth; i++) {
      MEEPO_CUDA_TRY(ev(&p->fwd[i].keys_up));
      MEEPO_CUDA_TRY(ev(&p->fwd[i].kernels_done));
      MEEPO_CUDA_TRY(ev(&p->fwd[i].free_ev));
      MEEPO_CUDA_TRY(ev(&p->bwd[i].keys_up));
      MEEPO_CUDA_TRY(ev(&p->bwd[i].grads_up));
      MEEPO_CUDA_TRY(ev(&p->bwd[i].free_ev));
    }
    for (int i = 0; i < 2; i++) {
      MEEPO_CUDA_TRY(ev(&p->k_done[i]));
      MEEPO_CUDA_TRY(ev(&p->out_done[i]));
      MEEPO_CUDA_TRY(cudaMalloc(&p->d_rows[i], kRowsChunkBytes));
    }
    for (int i = 0; i < kTicketRing; i++) MEEPO_CUDA_TRY(ev(&p->ticket_ev[i]));
    p->rows_chunk_bytes = kRowsChunkBytes;
  }
  *out = t->pipe;
  return MEEPO_OK;
}

void destroy_host_pipe(meepo_table* t) {
  HostPipe* p = t->pipe;
  if (!p) return;
  for (int i = 0; i < kDepth; i++) {
    cudaFree(p->fwd[i].d_keys);
    cudaFree(p->fwd[i].d_status);
    cudaFree(p->bwd[i].d_keys);
    cudaFree(p->bwd[i].d_grads);
    for (cudaEvent_t e : {p->fwd[i].keys_up, p->fwd[i].kernels_done, p->fwd[i].free_ev, p->bwd[i].keys_up,
                          p->bwd[i].grads_up, p->bwd[i].free_ev})
      if (e) cudaEventDestroy(e);
  }
  for (int i = 0; i < 2; i++) {
    cudaFree(p->d_rows[i]);
    if (p->k_done[i]) cudaEventDestroy(p->k_done[i]);
    if (p->out_done[i]) cudaEventDestroy(p->out_done[i]);
  }
  for (int i = 0; i < kTicketRing; i++)
    if (p->ticket_ev[i]) cudaEventDestroy(p->ticket_ev[i]);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_k) cudaStreamDestroy(p->s_k);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
  t->pipe = nullptr;
}

// A fresh ticket whose event will be recorded on `s` by the caller (after everything the verb enqueued there).
static meepo_status new_ticket(HostPipe* p, uint64_t* ticket, cudaEvent_t* ev) {
  const uint64_t id = p->next_ticket++;
  *ev = p->ticket_ev[id % kTicketRing];
  if (id > kTicketRing) MEEPO_CUDA_TRY(cudaEventSynchronize(*ev));  // ticket id - kTicketRing: long done, normally
  *ticket = id;
  return MEEPO_OK;
}

// Staging slot of a forward call, grown on demand (rare: drains the slot first).
static meepo_status fwd_slot(HostPipe* p, uint64_t n, HostPipe::Fwd** out) {
  HostPipe::Fwd& f = p->fwd[p->fwd_seq++ % kDepth];
  if (n > f.cap) {
    if (f.used) MEEPO_CUDA_TRY(cudaEventSynchronize(f.free_ev));
    cudaFree(f.d_keys);
    cudaFree(f.d_status);
    f.d_keys = nullptr;
    f.d_status = nullptr;
    f.cap = 0;
    const size_t cap = n + n / 4 + 1024;
    MEEPO_CUDA_TRY(cudaMalloc(&f.d_keys, cap * 8));
    MEEPO_CUDA_TRY(cudaMalloc(&f.d_status, cap));
    f.cap = cap;
  }
  *out = &f;
  return MEEPO_OK;
}

static meepo_status bwd_slot(HostPipe* p, uint64_t n, size_t gbytes, HostPipe::Bwd** out) {
  HostPipe::Bwd& b = p->bwd[p->bwd_seq++ % kDepth];
  if (n > b.keys_cap || gbytes > b.grads_cap) {
    if (b.used) MEEPO_CUDA_TRY(cudaEventSynchronize(b.free_ev));
    if (n > b.keys_cap) {
      cudaFree(b.d_keys);
      b.d_keys = nullptr;
      b.keys_cap = 0;
      const size_t cap = n + n / 4 + 1024;
      MEEPO_CUDA_TRY(cudaMalloc(&b.d_keys, cap * 8));
      b.keys_cap = cap;
    }
    if (gbytes > b.grads_cap) {
      cudaFree(b.d_grads);
      b.d_grads = nullptr;
      b.grads_cap = 0;
      MEEPO_CUDA_TRY(cudaMalloc(&b.d_grads, gbytes + gbytes / 8));
      b.grads_cap = gbytes + gbytes / 8;
    }
  }
  *out = &b;
  return MEEPO_OK;
}

static meepo_status probe_host_async(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                     uint8_t* status_out, bool insert, uint64_t* ticket) {
  if (!t || !ticket) return fail(MEEPO_EINVAL, "null argument");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!keys || !rows_out)) return fail(MEEPO_EINVAL, "null buffer");
  DeviceGuard guard(t->device);
  HostPipe* p;
  MEEPO_TRY(get_pipe(t, &p));
  VerbScope vs(t, p->s_k);  // orders s_k after the previous verb of this table, whatever stream it ran on
  MEEPO_TRY(vs.rc);
  cudaEvent_t done;
  MEEPO_TRY(new_ticket(p, ticket, &done));
  MEEPO_TRY(probe_gather_begin(t, n, insert, p->s_k));
  if (n == 0) {
    MEEPO_CUDA_TRY(cudaEventRecord(done, p->s_k));
    return MEEPO_OK;
  }
  HostPipe::Fwd* f;
  MEEPO_TRY(fwd_slot(p, n, &f));
  if (f->used) MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_in, f->free_ev, 0));  // the call two before this one
  f->used = true;
  MEEPO_CUDA_TRY(cudaMemcpyAsync(f->d_keys, keys, n * 8, cudaMemcpyHostToDevice, p->s_in));
  MEEPO_CUDA_TRY(cudaEventRecord(f->keys_up, p->s_in));
  MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_k, f->keys_up, 0));
  const uint64_t R = t->row_bytes;
  uint64_t chunk = p->rows_chunk_bytes / R;
  chunk = chunk / 32 * 32;
  if (chunk == 0) chunk = 32;
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = std::min(chunk, n - off);
    const int b = (int)(p->chunk_seq++ & 1);
    if (p->rows_used[b]) MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_k, p->out_done[b], 0));
    p->rows_used[b] = true;
    MEEPO_TRY(probe_gather_chunk(t, f->d_keys + off, m, p->d_rows[b], f->d_status + off, insert, p->s_k));
    MEEPO_CUDA_TRY(cudaEventRecord(p->k_done[b], p->s_k));
    MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_out, p->k_done[b], 0));
    MEEPO_CUDA_TRY(cudaMemcpyAsync((char*)rows_out + off * R, p->d_rows[b], m * R, cudaMemcpyDeviceToHost, p->s_out));
    MEEPO_CUDA_TRY(cudaEventRecord(p->out_done[b], p->s_out));
  }
  MEEPO_TRY(probe_gather_end(t, n, insert, p->s_k));
  MEEPO_CUDA_TRY(cudaEventRecord(f->kernels_done, p->s_k));
  MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_out, f->kernels_done, 0));
  if (status_out) MEEPO_CUDA_TRY(cudaMemcpyAsync(status_out, f->d_status, n, cudaMemcpyDeviceToHost, p->s_out));
  MEEPO_CUDA_TRY(cudaEventRecord(f->free_ev, p->s_out));
  MEEPO_CUDA_TRY(cudaEventRecord(done, p->s_out));
  return MEEPO_OK;
}

static meepo_status apply_host_async(meepo_table* t, const uint64_t* keys, const void* grads, uint64_t n,
                                     uint64_t* ticket) {
  if (!t || !ticket) return fail(MEEPO_EINVAL, "null argument");
  if (n > 0xFFFFFFFFull) return fail(MEEPO_EINVAL, "batch too large (n must fit in 32 bits)");
  if (n && (!keys || !grads)) return fail(MEEPO_EINVAL, "null buffer");
  DeviceGuard guard(t->device);
  HostPipe* p;
  MEEPO_TRY(get_pipe(t, &p));
  VerbScope vs(t, p->s_k);
  MEEPO_TRY(vs.rc);
  cudaEvent_t done;
  MEEPO_TRY(new_ticket(p, ticket, &done));
  if (n == 0) {
    MEEPO_CUDA_TRY(cudaEventRecord(done, p->s_k));
    return MEEPO_OK;
  }
  const size_t gbytes = (size_t)n * t->row_bytes;
  HostPipe::Bwd* b;
  MEEPO_TRY(bwd_slot(p, n, gbytes, &b));
  if (b->used) MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_in, b->free_ev, 0));
  b->used = true;
  MEEPO_CUDA_TRY(cudaMemcpyAsync(b->d_keys, keys, n * 8, cudaMemcpyHostToDevice, p->s_in));
  MEEPO_CUDA_TRY(cudaEventRecord(b->keys_up, p->s_in));
  MEEPO_CUDA_TRY(cudaMemcpyAsync(b->d_grads, grads, gbytes, cudaMemcpyHostToDevice, p->s_in));
  MEEPO_CUDA_TRY(cudaEventRecord(b->grads_up, p->s_in));
  MEEPO_CUDA_TRY(cudaStreamWaitEvent(p->s_k, b->keys_up, 0));
  // probe / sort / segment passes only need the keys: they run underneath the copy of the gradients
  MEEPO_TRY(launch_apply_gradients(t, b->d_keys, b->d_grads, n, p->s_k, b->grads_up));
  MEEPO_CUDA_TRY(cudaEventRecord(b->free_ev, p->s_k));
  MEEPO_CUDA_TRY(cudaEventRecord(done, p->s_k));
  return MEEPO_OK;
}

static meepo_status wait_ticket(meepo_table* t, uint64_t ticket) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  HostPipe* p = t->pipe;
  if (!p) return ticket == 0 ? MEEPO_OK : fail(MEEPO_EINVAL, "unknown ticket");
  DeviceGuard guard(t->device);
  if (ticket == 0) {  // everything issued so far
    MEEPO_CUDA_TRY(cudaStreamSynchronize(p->s_in));
    MEEPO_CUDA_TRY(cudaStreamSynchronize(p->s_k));
    MEEPO_CUDA_TRY(cudaStreamSynchronize(p->s_out));
    return sticky_error(t);
  }
  if (ticket >= p->next_ticket) return fail(MEEPO_EINVAL, "unknown ticket");
  if (ticket + kTicketRing >= p->next_ticket)  // its event has not been recycled yet
    MEEPO_CUDA_TRY(cudaEventSynchronize(p->ticket_ev[ticket % kTicketRing]));
  return sticky_error(t);
}

}  // namespace meepo

using namespace meepo;

extern "C" {

MEEPO_API meepo_status meepo_find_or_insert_host_async(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                       void* rows_out, uint8_t* status_out, uint64_t* ticket) {
  return probe_host_async(t, keys, n, rows_out, status_out, true, ticket);
}
MEEPO_API meepo_status meepo_lookup_host_async(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                               uint8_t* found_out, uint64_t* ticket) {
  return probe_host_async(t, keys, n, rows_out, found_out, false, ticket);
}
MEEPO_API meepo_status meepo_apply_gradients_host_async(meepo_table* t, const uint64_t* keys, const void* grads,
                                                        uint64_t n, uint64_t* ticket) {
  return apply_host_async(t, keys, grads, n, ticket);
}
MEEPO_API meepo_status meepo_wait(meepo_table* t, uint64_t ticket) { return wait_ticket(t, ticket); }

MEEPO_API meepo_status meepo_find_or_insert_host(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                 void* rows_out, uint8_t* status_out) {
  uint64_t ticket = 0;
  MEEPO_TRY(probe_host_async(t, keys, n, rows_out, status_out, true, &ticket));
  return wait_ticket(t, ticket);
}
MEEPO_API meepo_status meepo_lookup_host(meepo_table* t, const uint64_t* keys, uint64_t n, void* rows_out,
                                         uint8_t* found_out) {
  uint64_t ticket = 0;
  MEEPO_TRY(probe_host_async(t, keys, n, rows_out, found_out, false, &ticket));
  return wait_ticket(t, ticket);
}
MEEPO_API meepo_status meepo_apply_gradients_host(meepo_table* t, const uint64_t* keys, const void* grads,
                                                  uint64_t n) {
  uint64_t ticket = 0;
  MEEPO_TRY(apply_host_async(t, keys, grads, n, &ticket));
  return wait_ticket(t, ticket);
}

}  // extern "C"
