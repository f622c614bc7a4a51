/* minimal.c — a plain C caller of libmeepo.so (include/meepo.h): host buffers in, host buffers out.
 *
 *   gcc -std=c99 -Iinclude examples/minimal.c -Lmeepoembedding_b200 -lmeepo -Wl,-rpath,$PWD/meepoembedding_b200 -o minimal
 *   ./minimal            (needs a B200; there is no CPU fallback)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "meepo.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    meepo_status s_ = (call);                                              \
    if (s_ != MEEPO_OK) {                                                  \
      fprintf(stderr, "%s -> %d: %s\n", #call, (int)s_, meepo_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void) {
  enum { DIM = 64, N = 1 << 16 };
  meepo_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.dim = DIM;
  cfg.capacity = 1u << 20;
  cfg.dtype = MEEPO_F32;
  cfg.opt = MEEPO_ADAGRAD;
  cfg.lr = 0.01f;
  cfg.eps = 1e-8f;
  cfg.init_accum = 0.1f;
  cfg.init_scale = 0.01f;
  cfg.init_seed = 1;
  cfg.flags = MEEPO_FLAG_TRACK_SCORES;
  cfg.host_spill_bytes = 64u << 20;
  meepo_table* t = NULL;
  CHECK(meepo_create(&cfg, &t));
  printf("backend %s, ABI %u\n", meepo_backend(), meepo_abi_version());

  uint64_t* keys = malloc(N * sizeof *keys);
  float* rows = malloc((size_t)N * DIM * sizeof *rows);
  float* grads = malloc((size_t)N * DIM * sizeof *grads);
  uint8_t* status = malloc(N);
  for (int i = 0; i < N; i++) keys[i] = 1000003ull * (uint64_t)(i % 50000) + 17;  /* with duplicates */
  for (size_t i = 0; i < (size_t)N * DIM; i++) grads[i] = 0.001f;

  CHECK(meepo_find_or_insert_host(t, keys, N, rows, status));   /* new keys: status == MEEPO_KEY_INSERTED */
  CHECK(meepo_apply_gradients_host(t, keys, grads, N));         /* duplicates summed, one Adagrad step per key */
  CHECK(meepo_lookup_host(t, keys, N, rows, status));           /* status == MEEPO_KEY_FOUND */

  /* the asynchronous forms: issue, overlap, wait. Two calls in flight execute in issue order. */
  uint64_t t_rows = 0, t_upd = 0;
  CHECK(meepo_find_or_insert_host_async(t, keys, N, rows, status, &t_rows));
  CHECK(meepo_apply_gradients_host_async(t, keys, grads, N, &t_upd));
  CHECK(meepo_wait(t, t_rows));                                 /* rows/status are valid from here on */
  CHECK(meepo_wait(t, 0));                                      /* everything issued so far */

  uint64_t evicted = 0;
  CHECK(meepo_evict(t, MEEPO_LFU, 0.02, &evicted, NULL));       /* lowest-frequency keys go to the host spill tier */
  CHECK(meepo_spill_readmit(t, keys, 1000, status));            /* ... and come back with row, state and score */
  CHECK(meepo_export(t, "/tmp/minimal.meepo"));

  meepo_stats_t st;
  CHECK(meepo_stats(t, &st));
  printf("size %llu of %llu slots, %llu evicted, %llu in the spill tier, row[0][0] = %g\n", (unsigned long long)st.size,
         (unsigned long long)st.capacity, (unsigned long long)evicted, (unsigned long long)st.spill_keys, rows[0]);
  CHECK(meepo_destroy(t));
  free(keys), free(rows), free(grads), free(status);
  return 0;
}
