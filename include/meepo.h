/* meepo.h — C ABI of the B200-native dynamic embedding table ("meepo-b200").
 *
 * Provenance. The upstream repository MoFHeka/MeepoEmbedding ships no code: its
 * whole product content is /root/reference/README.md:1-2 ("A distributed
 * high-performance dynamic lookuptable-style Embedding ... Supports GPU, CPU,
 * remote distributed KV (such as Redis), SSD, and other backends."). There is
 * therefore no reference FFI to bind to; every entry point below is DERIVED
 * from that sentence (R1 "dynamic lookuptable-style" -> find_or_insert / evict,
 * R2 "distributed" -> the sharded verbs, R4 "recommendation ... CTR" ->
 * apply_gradients with sparse optimizers) and from BASELINE.json:north_star.
 * The derivation table lives in BASELINE.md section 5; this header is the
 * normative statement of the semantics.
 *
 * Two shared libraries export exactly these symbols:
 *   meepoembedding_b200/libmeepo.so   CUDA sm_100a product; data pointers are
 *                                     DEVICE pointers unless the verb ends in
 *                                     _host.
 *   oracle/libmeepo_oracle.so         authored C++/OpenMP CPU restatement (test
 *                                     infrastructure only); all pointers are
 *                                     HOST pointers, `stream` is ignored.
 *
 * ---------------------------------------------------------------------------
 * Normative semantics (both libraries implement these bit for bit)
 * ---------------------------------------------------------------------------
 * Keys      uint64. 0xFFFFFFFFFFFFFFFF (EMPTY) and 0xFFFFFFFFFFFFFFFE (RESERVED)
 *           are not storable: they get per-key status MEEPO_KEY_INVALID and an
 *           all-zero row.
 * Rows      `dim` elements of fp32 or bf16, row-major, contiguous; row bytes
 *           must be a multiple of 16 (dim % 4 == 0 for fp32, dim % 8 == 0 for
 *           bf16).
 * Init      A row created by find_or_insert is a pure function of
 *           (key, init_seed, column) — never of slot or arrival order:
 *             p   = column >> 1
 *             x   = mix64(key + (init_seed ^ ((p + 1) * 0x9E3779B97F4A7C15)))
 *             u32 = (column & 1) ? (x >> 32) : (x & 0xFFFFFFFF)
 *             v   = ((float)(u32 >> 8) * 2^-23 - 1.0f) * init_scale   (fp32 ops)
 *           mix64 is the splitmix64 finaliser:
 *             z ^= z >> 30; z *= 0xBF58476D1CE4E5B9; z ^= z >> 27;
 *             z *= 0x94D049BB133111EB; z ^= z >> 31.
 *           bf16 tables store round-to-nearest-even(v).
 *           Optimizer state of a new row: Adagrad accumulator(s) = init_accum,
 *           Adam m = v = 0 and step = 0.
 * Status    find_or_insert reports per key: MEEPO_KEY_FOUND if the key was in
 *           the table when the call started; MEEPO_KEY_INSERTED if it was not
 *           (EVERY duplicate of such a key inside the batch reports INSERTED and
 *           receives the same freshly initialised row; one slot is used);
 *           MEEPO_KEY_FULL if no free slot exists anywhere (row = zeros);
 *           MEEPO_KEY_INVALID for the two reserved keys. lookup reports
 *           MEEPO_KEY_FOUND / MEEPO_KEY_MISS (row = zeros) / MEEPO_KEY_INVALID.
 *           The physical slot index is NOT part of the contract.
 * Update    apply_gradients sums the gradients of duplicate keys, then performs
 *           exactly one optimizer step per unique key. Keys that are not in the
 *           table (or invalid) are skipped. The sum is taken in fp32 in this
 *           fixed order (g_0..g_{n-1} = the key's gradients by increasing batch
 *           index; L = MEEPO_REDUCE_LEAF = 256):
 *             leaf_c = ((g_{cL} + g_{cL+1}) + g_{cL+2}) + ...   (up to L terms)
 *             G      = ((leaf_0 + leaf_1) + leaf_2) + ...
 *           Optimizers, element-wise, fp32 arithmetic, every operation rounded
 *           individually (no fused multiply-add), w = row, g = G:
 *             SGD      w = w - lr*g
 *             ADAGRAD  a = a + g*g ;  w = w - (lr*g) / (sqrt(a) + eps)
 *             ADAGRAD_ROWWISE  ONE fp32 accumulator per row (state = 16 bytes per
 *                      row: {a, 0, 0, 0}) fed the mean square of the row's gradient,
 *                      whose summation order is fixed as follows (a "chunk" is 16
 *                      bytes of the row: 4 fp32 or 8 bf16 elements):
 *                        c_q = ((g_0*g_0 + g_1*g_1) + g_2*g_2) + ...  over the
 *                              elements of chunk q, in order
 *                        s_r = sum of c_q over the chunks with q mod 32 == r, in
 *                              increasing q (r = 0..31; 0 if there is none)
 *                        for d = 16, 8, 4, 2, 1:  s_r = s_r + s_{r+d}  (r < d)
 *                      a = a + s_0 / dim ;  w = w - (lr*g) / (sqrt(a) + eps)
 *             ADAM     t = t+1 (per row); m = beta1*m + (1-beta1)*g ;
 *                      v = beta2*v + (1-beta2)*(g*g) ;
 *                      w = w - (lr * sqrt(1-beta2^t)/(1-beta1^t)) * m
 *                              / (sqrt(v) + eps)
 *                      (the scalar step size is computed in double and rounded
 *                      to float once per row)
 *           bf16 rows: w is widened to fp32, updated, stored RNE. Gradients have
 *           the table's dtype and are widened to fp32 before summation.
 * Evict     Removes the lowest-score keys until size <= floor(target_load *
 *           capacity). Two uint32 scores are kept per key when
 *           MEEPO_FLAG_TRACK_SCORES is set: freq = number of occurrences of the
 *           key in find_or_insert / lookup batches while resident (the inserting
 *           batch included, wrapping at 2^32), and last_epoch = the table's batch
 *           epoch at the last such occurrence (the epoch increments once per
 *           find_or_insert / lookup call, first call = 1). LFU orders by freq,
 *           LRU by last_epoch; ties are broken by key, smaller key evicted
 *           first. apply_gradients does not touch the scores.
 * Host tier When host_spill_bytes > 0 the table has a SECOND LEVEL in pinned host
 *           memory: a ring of T = floor(host_spill_bytes / (24 + row_bytes +
 *           state_bytes)) tuple slabs. meepo_evict appends its victims (key, row,
 *           state, scores, step) at the head of the ring in eviction order
 *           (ascending (score, key)): the a-th tuple ever appended goes to slab
 *           a mod T, so once the ring has wrapped every append overwrites (drops)
 *           the oldest slab. A key has at most one tuple in the tier: a newer copy
 *           replaces the older one, whose slab then stays empty until the head
 *           passes it. The first level (HBM) always wins: a key that sits in both
 *           (possible after meepo_import*) is served from HBM.
 *           find_or_insert of a key that is not in HBM but in the tier PROMOTES
 *           it instead of re-initialising it: the key gets a slot, its row,
 *           optimizer state and step count are restored, its scores are restored
 *           and then this batch's occurrences are counted (freq = tier freq +
 *           occurrences, last_epoch = epoch), and its tier slab becomes empty.
 *           Every occurrence reports MEEPO_KEY_FOUND and receives the restored
 *           row: "was in the table when the call started" covers both levels.
 *           When no slot is free its occurrences report MEEPO_KEY_FULL (row =
 *           zeros) and the tuple stays in the tier. lookup serves such a key from
 *           the tier without promoting it (MEEPO_KEY_FOUND, the row; the tuple's
 *           scores do not change). apply_gradients does not look into the tier
 *           (a training step calls find_or_insert on its keys first): gradients
 *           of keys that are only there are dropped and counted in grad_dropped.
 *           meepo_spill_readmit promotes keys ahead of their next use.
 * Pooling   The *_pooled verbs take the batch as n_bags BAGS: bag b is keys[offsets[b]
 *           .. offsets[b+1]) with offsets[0] = 0 <= ... <= offsets[n_bags] = n
 *           (uint32). Towards the table they are exactly find_or_insert / lookup /
 *           apply_gradients on `keys` (insertion, per-key status, scores, host
 *           tier); what changes is the row traffic: the forward verbs return one
 *           pooled row per bag instead of one row per key,
 *             acc = +0.0f; for the keys of the bag in order: acc = acc + w(key)
 *           per column in fp32 (w = the key's row widened to fp32; a key without a
 *           row — MISS, FULL, INVALID — contributes nothing), MEEPO_POOL_MEAN then
 *           divides by the fp32 bag length offsets[b+1] - offsets[b] (an empty bag
 *           yields zeros), and the result is stored in the table dtype (bf16: RNE).
 *           The backward verb takes one gradient row per bag: the gradient of
 *           every key occurrence of bag b is bag_grads[b] (SUM) or bag_grads[b]
 *           widened, divided by the fp32 bag length and rounded to the table dtype
 *           (MEAN); "Update" then applies with occurrences in batch order.
 * Sharding  owner(key, G) = umulhi64(mix64(key ^ 0xD6E8FEB86659FD93), G).
 */
#ifndef MEEPO_H_
#define MEEPO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MEEPO_API __attribute__((visibility("default")))
#else
#define MEEPO_API
#endif

#define MEEPO_ABI_VERSION 3u
#define MEEPO_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define MEEPO_KEY_RESERVED 0xFFFFFFFFFFFFFFFEull
#define MEEPO_REDUCE_LEAF 256u
#define MEEPO_BUCKET_SLOTS 14u /* slots per 128-byte bucket line; capacity is rounded up to it */
#define MEEPO_OWNER_SALT 0xD6E8FEB86659FD93ull
#define MEEPO_MAX_PEERS 8u        /* shards of one sharded table (the GPUs of one NVSwitch box) */
#define MEEPO_PEER_BLOB_BYTES 256u /* size of the opaque per-rank blob of meepo_peer_prepare */

typedef struct meepo_table meepo_table; /* opaque */

typedef enum {
  MEEPO_OK = 0,
  MEEPO_EINVAL = 1,
  MEEPO_ENOMEM = 2,
  MEEPO_ECUDA = 3,
  MEEPO_ENCCL = 4,
  MEEPO_EIO = 5
} meepo_status;

typedef enum { MEEPO_F32 = 0, MEEPO_BF16 = 1 } meepo_dtype;
typedef enum { MEEPO_SGD = 0, MEEPO_ADAGRAD = 1, MEEPO_ADAM = 2, MEEPO_ADAGRAD_ROWWISE = 3 } meepo_opt;
typedef enum { MEEPO_LRU = 0, MEEPO_LFU = 1 } meepo_policy;
typedef enum { MEEPO_POOL_SUM = 0, MEEPO_POOL_MEAN = 1 } meepo_pool;

/* per-key status bytes */
enum {
  MEEPO_KEY_MISS = 0,
  MEEPO_KEY_FOUND = 1,
  MEEPO_KEY_INSERTED = 2,
  MEEPO_KEY_FULL = 3,
  MEEPO_KEY_INVALID = 4
};

enum { MEEPO_FLAG_TRACK_SCORES = 1u, MEEPO_FLAG_TRACK_DIRTY = 2u };

typedef struct {
  uint32_t dim;          /* elements per row */
  uint32_t flags;        /* MEEPO_FLAG_* */
  uint64_t capacity;     /* number of key slots (rounded up to a multiple of MEEPO_BUCKET_SLOTS) */
  int32_t dtype;         /* meepo_dtype */
  int32_t opt;           /* meepo_opt */
  float lr, eps, beta1, beta2, init_accum;
  float init_scale;
  uint64_t init_seed;
  int32_t device;            /* CUDA ordinal (ignored by the oracle) */
  int32_t reserved0;
  uint64_t host_spill_bytes; /* size of the pinned host spill tier, 0 = none */
} meepo_config;

typedef struct {
  uint64_t capacity;     /* slots */
  uint64_t size;         /* live keys */
  uint64_t inserts;      /* keys given a slot since create (new keys and imports; not promotions) */
  uint64_t hits;         /* key occurrences that were found (foi + lookup) */
  uint64_t misses;       /* lookup occurrences not found */
  uint64_t full;         /* occurrences rejected because the table was full */
  uint64_t evictions;    /* keys removed by meepo_evict */
  uint64_t updates;      /* unique-key optimizer steps applied */
  uint64_t grad_dropped; /* gradient occurrences whose key was absent/invalid */
  uint64_t spill_keys;   /* tuples currently held in the host tier */
  uint64_t spill_bytes;  /* bytes they occupy (spill_keys * (24 + row_bytes + state_bytes)) */
  uint64_t epoch;        /* batch epoch */
  uint64_t overflow_buckets; /* home buckets with at least one key displaced into a later bucket */
  uint64_t row_bytes, state_bytes; /* per slot */
  uint64_t peer_keys_received;  /* sharded forward verbs: (sender, key) entries this owner served */
  uint64_t peer_grads_received; /* sharded apply_gradients: (sender, key) gradient rows received */
  uint64_t promotions;          /* keys find_or_insert / spill_readmit brought back from the host tier */
  uint64_t tier_hits;           /* lookup occurrences served from the host tier */
  uint64_t probe_hist[4];       /* live keys sitting 0, 1, 2, >= 3 buckets past their home bucket (probe length - 1);
                                   the oracle, which has no buckets, reports {size, 0, 0, 0} */
} meepo_stats_t;

/* --- life cycle (synchronous) ------------------------------------------- */
MEEPO_API uint32_t meepo_abi_version(void);
/* "cuda-sm_100a" for the product, "oracle-cpu" for the authored CPU library. */
MEEPO_API const char* meepo_backend(void);
MEEPO_API meepo_status meepo_create(const meepo_config* cfg, meepo_table** out);
MEEPO_API meepo_status meepo_destroy(meepo_table* t);
MEEPO_API meepo_status meepo_stats(meepo_table* t, meepo_stats_t* out);
/* thread-local, never NULL */
MEEPO_API const char* meepo_last_error(void);

/* --- per-kernel timing (bench.py's roofline leg) --------------------------- *
 * When enabled the library brackets every kernel it launches with CUDA events
 * on the launch stream. meepo_profile_read drains them and writes one line per
 * kernel group, "name launches total_ms\n", NUL-terminated, into buf. Enabling
 * resets the accumulators. The oracle accepts both calls and reports nothing. */
MEEPO_API meepo_status meepo_profile_enable(meepo_table* t, int32_t on);
MEEPO_API meepo_status meepo_profile_read(meepo_table* t, char* buf, uint64_t buf_bytes);

/* --- hot path (stream-ordered, asynchronous; device pointers) ------------ *
 * Verbs of ONE table share its scratch memory and execute one after another:
 * a verb issued on another stream than its predecessor first waits (on the
 * device) for the predecessor to finish, so a caller may mix streams freely —
 * but verbs of one table never overlap each other. Different tables are
 * independent. A sticky device-side failure (a sort / compaction look-back or a
 * peer barrier that timed out, an exchange-window lane that overflowed) makes
 * every following verb of the table fail with MEEPO_ECUDA / MEEPO_ENCCL rather
 * than compute on corrupt state. */
/* rows_out: n*dim elements; status_out: n bytes (may be NULL). */
MEEPO_API meepo_status meepo_find_or_insert(meepo_table* t, const uint64_t* keys, uint64_t n,
                                            void* rows_out, uint8_t* status_out, void* stream);
MEEPO_API meepo_status meepo_lookup(meepo_table* t, const uint64_t* keys, uint64_t n,
                                    void* rows_out, uint8_t* found_out, void* stream);
/* grads: n*dim elements of the table dtype. */
MEEPO_API meepo_status meepo_apply_gradients(meepo_table* t, const uint64_t* keys,
                                             const void* grads, uint64_t n, void* stream);

/* --- pooled (bag) forms: lookup fused with the sum / mean pooling that follows it
 *     in CTR models; see "Pooling" above. Stream-ordered, device pointers.
 *     keys: n; offsets: n_bags + 1 uint32; pooled_out / bag_grads: n_bags * dim
 *     elements of the table dtype; status_out: n bytes (may be NULL). The per-key
 *     row ([n][dim], written and read back once each by the unfused sequence) never
 *     exists; the backward verb reads bag_grads through the bag index of every
 *     occurrence instead of an expanded [n][dim] gradient. */
MEEPO_API meepo_status meepo_find_or_insert_pooled(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                   const uint32_t* offsets, uint64_t n_bags, int32_t pool,
                                                   void* pooled_out, uint8_t* status_out, void* stream);
MEEPO_API meepo_status meepo_lookup_pooled(meepo_table* t, const uint64_t* keys, uint64_t n,
                                           const uint32_t* offsets, uint64_t n_bags, int32_t pool,
                                           void* pooled_out, uint8_t* found_out, void* stream);
MEEPO_API meepo_status meepo_apply_gradients_pooled(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                    const uint32_t* offsets, uint64_t n_bags, int32_t pool,
                                                    const void* bag_grads, void* stream);

/* --- host-buffer front ends (pageable or pinned host pointers) ----------- *
 * Same semantics; the library stages through its own device buffers and
 * overlaps H2D, kernels and D2H in chunks on private streams. The plain verbs
 * return after the results are in the caller's buffers. */
MEEPO_API meepo_status meepo_find_or_insert_host(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                 void* rows_out, uint8_t* status_out);
MEEPO_API meepo_status meepo_lookup_host(meepo_table* t, const uint64_t* keys, uint64_t n,
                                         void* rows_out, uint8_t* found_out);
MEEPO_API meepo_status meepo_apply_gradients_host(meepo_table* t, const uint64_t* keys,
                                                  const void* grads, uint64_t n);

/* Asynchronous forms: enqueue and return a ticket. meepo_wait(t, ticket) blocks
 * until that verb's results are in the caller's buffers (find_or_insert /
 * lookup) or its inputs have been consumed (apply_gradients); ticket 0 waits
 * for everything issued so far. The table executes the verbs strictly in ISSUE
 * ORDER, whatever their copies overlap with: a caller that issues
 *   find_or_insert(batch i+1); apply_gradients(batch i); ...
 * (two batches in flight) gets the rows of batch i+1 coming down over PCIe
 * while the gradients of batch i go up, and the rows it reads are those of the
 * table before the update of batch i — exactly what the same sequence of
 * synchronous calls returns. Buffers must stay valid (and, for the outputs,
 * untouched) until the ticket has been waited for; only pinned buffers
 * overlap. At most 32 tickets may be outstanding. The oracle library executes
 * the verb at once and hands back a ticket that is already complete. */
MEEPO_API meepo_status meepo_find_or_insert_host_async(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                       void* rows_out, uint8_t* status_out, uint64_t* ticket);
MEEPO_API meepo_status meepo_lookup_host_async(meepo_table* t, const uint64_t* keys, uint64_t n,
                                               void* rows_out, uint8_t* found_out, uint64_t* ticket);
MEEPO_API meepo_status meepo_apply_gradients_host_async(meepo_table* t, const uint64_t* keys,
                                                        const void* grads, uint64_t n, uint64_t* ticket);
MEEPO_API meepo_status meepo_wait(meepo_table* t, uint64_t ticket);

/* --- capacity management -------------------------------------------------- */
/* Stream-ordered. The call reads the table size (one host synchronisation with
 * `stream`), then enqueues the selection, the hand-over to the host tier and
 * the slot release and returns; *n_evicted (may be NULL) is exact on return.
 * The victims' tuples are first gathered into a device staging buffer and
 * drained to the pinned ring by plain DMA copies on a private stream underneath
 * whatever the caller enqueues next; until the next meepo_evict the staged
 * tuples are served (promoted, looked up) straight from HBM. */
MEEPO_API meepo_status meepo_evict(meepo_table* t, int32_t policy, double target_load,
                                   uint64_t* n_evicted, void* stream);
/* Promote keys from the host tier ahead of their next use (row, state, step and
 * scores restored as they were; the call counts no occurrence). keys: HOST
 * pointer. status_out (host, may be NULL): FOUND = already in HBM (a tier copy,
 * if there is one, is dropped), INSERTED = restored (every duplicate of such a
 * key reports INSERTED), MISS = in neither level, FULL / INVALID as above. */
MEEPO_API meepo_status meepo_spill_readmit(meepo_table* t, const uint64_t* keys, uint64_t n,
                                           uint8_t* status_out);

/* --- bulk dump / load (synchronous) --------------------------------------- *
 * meepo_export_buffers compacts the live tuples into caller buffers (device
 * pointers for the CUDA library), sorted by key ascending. rows/state/scores/
 * steps may each be NULL. score = (last_epoch << 32) | freq; steps = Adam's
 * per-row step count (0 for the other optimizers). *n_out receives the number
 * of live tuples; keys == NULL makes the call a pure size query; max_n is the
 * size of the buffers in tuples (MEEPO_EINVAL if too small). */
MEEPO_API meepo_status meepo_export_buffers(meepo_table* t, uint64_t* keys, void* rows, void* state,
                                            uint64_t* scores, uint32_t* steps, uint64_t max_n,
                                            uint64_t* n_out);
/* meepo_import_buffers: bulk insert-or-overwrite of tuples with DISTINCT keys.
 * state == NULL -> state initialised as for a new row; scores/steps == NULL ->
 * 0. status_out (may be NULL): FOUND = overwritten, INSERTED, FULL, INVALID. */
MEEPO_API meepo_status meepo_import_buffers(meepo_table* t, const uint64_t* keys, const void* rows,
                                            const void* state, const uint64_t* scores,
                                            const uint32_t* steps, uint64_t n, uint8_t* status_out);
/* File format "MEEPOTB1": 56-byte header {magic[8], u32 version, dim, dtype,
 * opt, u64 n, row_bytes, state_bytes, epoch} then keys[n], rows[n], state[n],
 * scores[n], steps[n], tuples sorted by key. Both libraries write identical
 * files for identical tables. */
MEEPO_API meepo_status meepo_export(meepo_table* t, const char* path);
MEEPO_API meepo_status meepo_import(meepo_table* t, const char* path);
/* Incremental export (needs MEEPO_FLAG_TRACK_DIRTY). The table remembers which
 * tuples were inserted (find_or_insert), updated (apply_gradients), imported or
 * re-admitted since they were last written by a delta export (or since create).
 * meepo_export_delta_buffers returns exactly those tuples, sorted by key, with
 * the calling convention of meepo_export_buffers, and marks them clean; a pure
 * size query (keys == NULL) marks nothing. meepo_export_delta writes them as a
 * file of the same "MEEPOTB1" format, so meepo_import applies a delta as an
 * upsert: a full export followed by the deltas taken after it, imported in
 * order, reproduces the table's tuples (keys evicted in between stay in the
 * importing table: a delta carries no deletions). Full exports neither read nor
 * change the marks. Eviction forgets the mark of the evicted tuple. */
MEEPO_API meepo_status meepo_export_delta_buffers(meepo_table* t, uint64_t* keys, void* rows, void* state,
                                                  uint64_t* scores, uint32_t* steps, uint64_t max_n,
                                                  uint64_t* n_out);
MEEPO_API meepo_status meepo_export_delta(meepo_table* t, const char* path);

/* --- tier dump / load (synchronous): the host tier's side of a checkpoint --- *
 * meepo_export* cover the HBM level. A table whose host tier is in use is
 * checkpointed as TWO files: meepo_export + meepo_tier_export, and restored
 * with meepo_import + meepo_tier_import (a delta only ever holds HBM tuples: a
 * tuple has to be promoted before it can change). meepo_tier_export(_buffers)
 * writes the live tuples of the tier sorted by key (same conventions and file
 * format as meepo_export(_buffers); keys == NULL is a size query);
 * meepo_tier_import(_buffers) APPENDS tuples with distinct keys to the ring in
 * the given order, exactly as an eviction would (the a-th append goes to slab
 * a mod T, a newer copy of a key replaces the older one); state == NULL ->
 * initial state, scores / steps == NULL -> 0. After a restore the ring holds
 * the tuples in key order, i.e. the smallest keys are overwritten first. */
MEEPO_API meepo_status meepo_tier_export_buffers(meepo_table* t, uint64_t* keys, void* rows, void* state,
                                                 uint64_t* scores, uint32_t* steps, uint64_t max_n,
                                                 uint64_t* n_out);
MEEPO_API meepo_status meepo_tier_import_buffers(meepo_table* t, const uint64_t* keys, const void* rows,
                                                 const void* state, const uint64_t* scores,
                                                 const uint32_t* steps, uint64_t n);
MEEPO_API meepo_status meepo_tier_export(meepo_table* t, const char* path);
MEEPO_API meepo_status meepo_tier_import(meepo_table* t, const char* path);

/* --- sharding helpers (one process per GPU; the exchange itself is the
 *     caller's collective: torch.distributed / NCCL all-to-all, or the fused
 *     peer-memory kernels below) ------------------------------------------- */
/* owner(key, num_shards) on the host, for tests and callers. */
MEEPO_API uint32_t meepo_owner(uint64_t key, uint32_t num_shards);
/* Stable counting sort of a batch by owner(key, num_shards):
 *   counts_out[num_shards]  keys per destination
 *   perm_out[n]             perm_out[j] = batch index of the j-th key in
 *                           destination-major order (may be NULL)
 *   keys_sorted_out[n]      keys in that order (may be NULL)
 * Device pointers, stream-ordered. */
MEEPO_API meepo_status meepo_shard_partition(meepo_table* t, const uint64_t* keys, uint64_t n,
                                             uint32_t num_shards, uint64_t* counts_out,
                                             uint32_t* perm_out, uint64_t* keys_sorted_out,
                                             void* stream);
/* Sender-side pre-reduction before every exchange (NVLink is ~3x tighter than
 * HBM without it): the distinct valid keys of the batch, each with the
 * fixed-shape sum (meepo.h "Update") of its gradient rows rounded to the table
 * dtype. grads/grads_out may both be NULL (keys only — the forward path).
 * inverse_out (may be NULL): inverse_out[i] = position of keys[i] in
 * unique_keys_out, 0xFFFFFFFF for invalid keys. The ORDER of the unique keys is
 * unspecified; results are compared as a key -> row map. *n_unique_out is a
 * device pointer in the CUDA library. */
MEEPO_API meepo_status meepo_reduce_duplicates(meepo_table* t, const uint64_t* keys, const void* grads,
                                               uint64_t n, uint64_t* unique_keys_out, void* grads_out,
                                               uint32_t* inverse_out, uint64_t* n_unique_out,
                                               void* stream);
/* rows_out[i] = rows_in[index[i]] (16-byte vectorised gather of dense rows; the
 * un-permute / duplicate-expand step after the row exchange). index ==
 * 0xFFFFFFFF writes a zero row. */
MEEPO_API meepo_status meepo_gather_rows(meepo_table* t, const void* rows_in, const uint32_t* index,
                                         uint64_t n, void* rows_out, void* stream);

/* --- sharded verbs fused with their exchange over NVLink peer memory ------- *
 * One table per GPU (one process per GPU, or several tables driven by one
 * process), rank r owns the keys with meepo_owner(key, world) == r. Set-up:
 *   1. every rank: meepo_peer_prepare(t, rank, world, max_batch, region_keys,
 *      out_buffers, blob) allocates this rank's exchange window and fills
 *      `blob` (MEEPO_PEER_BLOB_BYTES, opaque; carries a CUDA IPC handle);
 *   2. the caller all-gathers the blobs (any transport; they are plain bytes);
 *   3. every rank: meepo_peer_attach(t, blobs) with the world blobs in rank
 *      order maps the peers' windows.
 * The three verbs are COLLECTIVE: every rank calls the same verb in the same
 * order (n may differ per rank and may be 0). They are stream-ordered and
 * asynchronous like the single-table verbs and do not synchronise with the
 * host; ranks meet in device-side flag barriers. Semantics are those of ONE
 * table fed the concatenation of all ranks' batches: a key that is new in the
 * collective call reports MEEPO_KEY_INSERTED on every rank; apply_gradients
 * first sums each rank's duplicate gradients (fixed-shape tree, rounded to the
 * table dtype — meepo_reduce_duplicates), then adds the ranks' partial sums in
 * rank order and performs one optimizer step per key.
 *   max_batch    largest n any call will pass on this rank
 *   region_keys  capacity of one (sender, owner) lane of the window in unique
 *                keys; 0 = max_batch (always sufficient). A smaller value
 *                (e.g. 1.25 * max_batch / world) saves window memory
 *                (2 * world * region_keys * row_bytes); exceeding it is
 *                reported by meepo_stats (and by every following verb) as
 *                MEEPO_ENCCL, never a memory error.
 *   out_buffers  number of OUTPUT BUFFERS (max_batch rows each) to place inside
 *                the window; meepo_peer_output(t, i, &ptr, &rows) returns the
 *                i-th. A sharded find_or_insert / lookup whose rows_out lies in
 *                one of them has the owners store every row straight into it
 *                over NVLink (one row per unique key, at one of the key's
 *                occurrences; the other occurrences are filled by a local copy)
 *                instead of going through a return region and a full expansion
 *                pass. Any other rows_out works too, just slower. Results are
 *                identical either way.
 * A rank that does not reach a barrier within MEEPO_PEER_TIMEOUT_MS (env,
 * default 20000) makes its peers give up and report MEEPO_ENCCL from
 * meepo_stats instead of hanging the GPU. All ranks must have finished (e.g. a
 * host barrier) before any of them calls meepo_peer_detach / meepo_destroy.
 * The oracle library exports these symbols and returns MEEPO_EINVAL: the
 * checker for the sharded verbs is ONE oracle table fed the concatenated
 * batches (tests/test_gpu_peer.py). */
MEEPO_API meepo_status meepo_peer_prepare(meepo_table* t, uint32_t rank, uint32_t world, uint64_t max_batch,
                                          uint64_t region_keys, uint32_t out_buffers, void* blob_out);
MEEPO_API meepo_status meepo_peer_output(meepo_table* t, uint32_t index, void** rows_out, uint64_t* max_rows);
MEEPO_API meepo_status meepo_peer_attach(meepo_table* t, const void* blobs);
MEEPO_API meepo_status meepo_peer_detach(meepo_table* t);
MEEPO_API meepo_status meepo_sharded_find_or_insert(meepo_table* t, const uint64_t* keys, uint64_t n,
                                                    void* rows_out, uint8_t* status_out, void* stream);
MEEPO_API meepo_status meepo_sharded_lookup(meepo_table* t, const uint64_t* keys, uint64_t n,
                                            void* rows_out, uint8_t* found_out, void* stream);
MEEPO_API meepo_status meepo_sharded_apply_gradients(meepo_table* t, const uint64_t* keys,
                                                     const void* grads, uint64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEEPO_H_ */
