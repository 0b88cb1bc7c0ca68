/* yrb200.h — C ABI of the B200-native dense-retrieval backend for Youtu-RAG.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own: its
 * vector-store interface is the Python ABC `BaseVectorStore` (utu/rag/base.py:187-232) and the
 * arithmetic sits one call below it, in chromadb's `collection.query`
 * (utu/rag/storage/implementations/chroma_store.py:118-120).  Each entry point here names the
 * reference call it stands in for; the Python subclass that binds them with ctypes is
 * youtu-rag_b200/store.py (INTEGRATION.md shows the stub a maintainer adds to the reference).
 *
 * Conventions: plain C types only; every function returns YRB_OK (0) or a negative error code
 * and never throws; `yrb_last_error()` gives the message for the calling thread; the library
 * owns all device memory; callers own every host buffer; `stream` arguments are `cudaStream_t`
 * passed as `void*` (NULL = the index's own stream).  Functions taking host buffers are
 * synchronous; `_device` variants are asynchronous on `stream`.
 */
#ifndef YRB200_H
#define YRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YRB_ABI_VERSION 1

#if defined(__GNUC__)
#define YRB_API __attribute__((visibility("default")))
#else
#define YRB_API
#endif

/* status codes */
#define YRB_OK 0
#define YRB_ERR_INVALID (-1)   /* bad argument (message says which) */
#define YRB_ERR_CUDA (-2)      /* a CUDA runtime/driver call failed */
#define YRB_ERR_NOMEM (-3)     /* device or host allocation failed */
#define YRB_ERR_UNSUPPORTED (-4)
#define YRB_ERR_NODEVICE (-5)  /* no sm_100 device visible: the product has no CPU fallback */

/* distance_metric of VectorStoreConfig (utu/rag/config.py:62) → chroma hnsw:space
 * (chroma_store.py:46-59).  score returned = 1 - distance for every metric (chroma_store.py:132-135):
 * cosine → cos-sim, dot → inner product, euclidean → 1 - ||a-b||^2. */
#define YRB_METRIC_COSINE 0
#define YRB_METRIC_DOT 1
#define YRB_METRIC_L2 2

#define YRB_DTYPE_BF16 0
#define YRB_DTYPE_F32 1

/* largest k the fused (in-register / in-epilogue) selection handles; larger k takes the
 * score-matrix + radix-select path (same results). */
#define YRB_FUSED_K_MAX 128

typedef struct yrb_index yrb_index; /* one collection's rows resident on one GPU */

YRB_API int yrb_abi_version(void);
YRB_API const char* yrb_last_error(void);
/* number of CUDA devices with compute capability 10.x; YRB_ERR_NODEVICE if none. */
YRB_API int yrb_device_count(int* out_count);

/* ------------------------------------------------------------------ residency
 * stands in for chromadb.PersistentClient.get_or_create_collection (chroma_store.py:41-59). */
YRB_API int yrb_index_create(yrb_index** out, int device, int dim, int metric, int storage_dtype,
                     int64_t reserve_rows);
YRB_API int yrb_index_destroy(yrb_index* ix);
YRB_API int yrb_index_reserve(yrb_index* ix, int64_t rows);
/* rows = appended so far (incl. tombstoned); live = rows not tombstoned (collection.count(),
 * chroma_store.py:249-255). */
YRB_API int yrb_index_count(const yrb_index* ix, int64_t* out_rows, int64_t* out_live);
YRB_API int yrb_index_info(const yrb_index* ix, int* out_dim, int* out_ld, int* out_metric,
                   int* out_dtype, int* out_device, int64_t* out_capacity);

/* collection.add(embeddings=…) (chroma_store.py:86): append n fp32 rows [n, dim]; cosine rows are
 * L2-normalised (fp64 norm) and all rows rounded to the storage dtype on the device (kernel K5).
 * New rows get ids [rows, rows+n). */
YRB_API int yrb_index_append_host_f32(yrb_index* ix, const float* rows, int64_t n);
YRB_API int yrb_index_append_device_f32(yrb_index* ix, const float* dev_rows, int64_t n, void* stream);
/* stored rows decoded back to fp32 (Chunk.embedding of get_by_id, chroma_store.py:236-245). */
YRB_API int yrb_index_read_rows(yrb_index* ix, const int64_t* row_ids, int64_t n, float* out_rows);
/* Persistence (Chroma's on-disk segments, chroma_store.py:41-44): rows exactly as stored
 * (storage dtype, ld elements per row — see yrb_index_info) plus their squared norms, so that a reload
 * is bit-identical without re-normalising.  read_raw copies rows [row_begin, row_begin+n) to the host;
 * append_raw appends n such rows. */
YRB_API int yrb_index_read_raw(yrb_index* ix, int64_t row_begin, int64_t n, void* out_rows, float* out_sqnorm);
YRB_API int yrb_index_append_raw(yrb_index* ix, const void* rows, const float* sqnorm, int64_t n);
/* collection.delete (chroma_store.py:150-160): tombstone / revive rows; tombstoned rows are
 * invisible to search. */
YRB_API int yrb_index_set_live(yrb_index* ix, const int64_t* row_ids, int64_t n, int live);
/* Drop the rows appended last so that `rows` remain (rollback of a collection.add whose on-disk segment could not
 * be written, chroma_store.py:86: memory and disk must not diverge). */
YRB_API int yrb_index_truncate(yrb_index* ix, int64_t rows);
/* client.delete_collection + recreate (chroma_store.py:257-272). */
YRB_API int yrb_index_clear(yrb_index* ix);

/* ------------------------------------------------------------------ metadata filter
 * Columnar metadata resident on the device, written by the host as chunks are added.  A column
 * holds ONE (field, type) pair; `present` marks rows that have the field with that type. */
#define YRB_COL_I64 0   /* int metadata (chunk_index, *_min_stamp, …) */
#define YRB_COL_F64 1   /* float metadata (importance_score, …) */
#define YRB_COL_CODE 2  /* dictionary-coded str, int32 codes assigned by the host */
#define YRB_COL_BOOL 3  /* uint8 0/1 */
YRB_API int yrb_index_column_write(yrb_index* ix, int col, int col_type, int64_t row_begin, int64_t n,
                           const void* values, const uint8_t* present /* n bytes 0/1 */);

#define YRB_OP_EQ 0
#define YRB_OP_NE 1
#define YRB_OP_GT 2
#define YRB_OP_GTE 3
#define YRB_OP_LT 4
#define YRB_OP_LTE 5
#define YRB_OP_IN 6
#define YRB_OP_NIN 7
/* a leaf with col < 0 refers to a (field,type) no row has: EQ/IN/GT… → false, NE/NIN → true */
typedef struct yrb_where_leaf {
    int32_t col;            /* column id given to yrb_index_column_write, or -1 */
    int32_t op;             /* YRB_OP_* */
    int32_t operand_begin;  /* first operand in yrb_where.operands */
    int32_t operand_count;  /* 1, or the list length for IN / NIN */
} yrb_where_leaf;
#define YRB_TOK_AND (-1)
#define YRB_TOK_OR (-2)
#define YRB_TOK_NOT (-3)
#define YRB_WHERE_MAX_LEAVES 64
#define YRB_WHERE_MAX_OPERANDS 256
#define YRB_WHERE_MAX_TOKENS 160
/* a compiled Chroma `where` tree (chroma_store.py:104-120): postfix over leaves. */
typedef struct yrb_where {
    const yrb_where_leaf* leaves;
    int32_t n_leaves;
    const int64_t* operands; /* raw 8-byte patterns: int64, double, or int32 code / bool widened */
    int32_t n_operands;
    const int32_t* postfix;  /* token ≥ 0 = push leaf; YRB_TOK_* = operator */
    int32_t n_postfix;
} yrb_where;
/* evaluate `w` over all rows (kernel K4) AND the live bits.  out_mask (host, may be NULL) gets
 * ceil(rows/32) uint32 words, bit i of word j = row 32j+i; out_pass = number of passing rows. */
YRB_API int yrb_index_where(yrb_index* ix, const yrb_where* w, uint32_t* out_mask, int64_t* out_pass);

/* ------------------------------------------------------------------ search
 * stands in for collection.query(query_embeddings, n_results=k, where=…) (chroma_store.py:118-120),
 * exact instead of HNSW.  queries: fp32 [nq, dim] (normalised + rounded on the device for cosine).
 * Filter: `w` (compiled where, evaluated on the device) and/or `mask` (host bitmask, ceil(rows/32)
 * words, shared by all queries); both NULL = all live rows.  Pre-filter semantics.
 * Outputs per query q: out_ids[q*k + j] (row id, -1 padding), out_scores[q*k + j], ordered
 * (score desc, id asc); out_counts[q] = number of valid results (≤ k). */
YRB_API int yrb_index_search(yrb_index* ix, const float* queries, int nq, int k, const yrb_where* w,
                     const uint32_t* mask, int64_t* out_ids, float* out_scores,
                     int32_t* out_counts);

/* Search with options.  `min_score`: the retriever's similarity threshold (base_retriever.py:71 keeps a hit when
 * `score >= threshold`) applied INSIDE the scan — it is the initial bound of every running top-k list (K1) and of the
 * epilogue's survivor test (K2), so rows below it are never candidates and out_counts counts qualifying hits only.
 * The result equals filtering the plain top-k on the host (the list is sorted; the threshold only cuts its tail).
 * -INFINITY (or opts == NULL) = no threshold.  `w` (shared) or `wheres` (one per query) or neither; `mask` as in
 * yrb_index_search (not with `wheres`). */
typedef struct yrb_search_opts {
    float min_score;
    int32_t reserved[7]; /* zero */
} yrb_search_opts;
YRB_API int yrb_index_search_ex(yrb_index* ix, const float* queries, int nq, int k, const yrb_where* w,
                                const yrb_where* const* wheres, const uint32_t* mask, const yrb_search_opts* opts,
                                int64_t* out_ids, float* out_scores, int32_t* out_counts);

/* One filter per query (wheres[q] may be NULL = no filter): the batched form of the reference's
 * per-column searches (utu/tools/text2sql/unified_schemalink_valuelink.py:289-303 loops
 * CourseSearcher.search, chroma_retrical_text2sql.py:148-194, one filtered search per column with
 * the same query vector).  Identical programs (same pointer) are evaluated once. */
YRB_API int yrb_index_search_multi(yrb_index* ix, const float* queries, int nq, int k,
                                   const yrb_where* const* wheres, int64_t* out_ids, float* out_scores,
                                   int32_t* out_counts);

/* All-device variant for resident inputs (bench `value`, sharded search): queries fp32 [nq, dim]
 * on the device, mask device words or NULL, outputs = nq*k packed 64-bit selection keys
 * (see yrb_key_*), best first, 0 = empty slot.  Asynchronous on `stream` — with one exception: a batch (nq >= 2)
 * under a shared mask over >= 65536 rows first counts the passing rows (K8 decides on the host whether to gather them),
 * which synchronises `stream` once.
 * All searches of one index share its scratch (per-CTA lists, ticket counters, candidate buffers):
 * enqueue them on ONE stream, or order streams with events; concurrent searches need separate indexes. */
YRB_API int yrb_index_search_device(yrb_index* ix, const float* dev_queries, int nq, int k,
                            const uint32_t* dev_mask, uint64_t* dev_out_keys, void* stream);

/* Same, but decoded on the device: out_ids [nq*k] (-1 padding), out_scores [nq*k], out_counts [nq]
 * (all device pointers).  Single-GPU serving path: a single-query search is ONE kernel launch. */
YRB_API int yrb_index_search_device_ids(yrb_index* ix, const float* dev_queries, int nq, int k,
                                        const uint32_t* dev_mask, int64_t* dev_out_ids, float* dev_out_scores,
                                        int32_t* dev_out_counts, void* stream);

/* The top-k merge collective's local step (kernel K3): `parts` sorted lists of k keys
 * per query, laid out [parts][nq][k] (the NCCL all-gather buffer, part p = rank p's shard) →
 * out_ids (global id = row_base[p] + local row), out_scores, out_counts.  All device pointers. */
YRB_API int yrb_merge_topk_device(int device, const uint64_t* dev_keys, int parts, int nq, int k,
                          const int64_t* dev_row_base, int64_t* dev_out_ids, float* dev_out_scores,
                          int32_t* dev_out_counts, void* stream);

/* The same collective as ONE kernel over NVLink peer memory (kernel K7): every rank stores its keys
 * directly into all peers' gather buffers (CUDA-IPC mapped), raises per-query flags, waits for the
 * peers' flags and merges — no NCCL call, no host-side collective enqueue.  One process per GPU:
 * create() allocates this rank's buffers and returns yrb_exchange_handle_bytes() bytes of IPC handles;
 * the caller all-gathers the handles of all ranks (any transport) and passes them to connect().
 * merge() must be called by every rank, in the same order, with the same nq and k. */
typedef struct yrb_exchange yrb_exchange;
YRB_API int yrb_exchange_handle_bytes(void);
YRB_API const char* yrb_exchange_last_error(void);
YRB_API int yrb_exchange_create(yrb_exchange** out, int device, int world, int rank, int nq_cap, int k_cap,
                                unsigned char* out_handles);
YRB_API int yrb_exchange_connect(yrb_exchange* ex, const unsigned char* all_handles /* world * handle_bytes */);
YRB_API int yrb_exchange_merge(yrb_exchange* ex, const uint64_t* dev_local_keys, int nq, int k,
                               const int64_t* dev_row_base, int64_t* dev_out_ids, float* dev_out_scores,
                               int32_t* dev_out_counts, void* stream);
/* One rank's whole sharded search through HOST buffers, enqueued from C on the index's stream: query upload, local
 * scan + top-k, merge(), result in the caller's arrays (global ids).  Every rank calls it with the same queries;
 * errors are reported through yrb_last_error(). */
YRB_API int yrb_exchange_search(yrb_exchange* ex, yrb_index* ix, const float* queries, int nq, int k,
                                const uint32_t* dev_mask, const int64_t* dev_row_base, int64_t* out_ids, float* out_scores,
                                int32_t* out_counts);
YRB_API int yrb_exchange_destroy(yrb_exchange* ex);

/* ------------------------------------------------------------------ one collection over several GPUs, one process
 * `VectorStoreFactory.create` hands out one store per collection inside the agents' single serving process
 * (utu/rag/storage/base_storage.py:28-42, utu/rag/rag_tools/base_toolkit.py:79-91), so the multi-GPU form of the
 * store is an object of that process too: n `yrb_index` shards, rows dealt block-cyclically (global row g lives in
 * block g / block_rows on shard block % n), one worker thread per device, and the cross-GPU top-k merge folded
 * into the kernel that finishes each query (NVLink peer stores + a ticket; the last shard to arrive merges and
 * writes the result into pinned host memory).  Every entry mirrors its yrb_index_* namesake and works in GLOBAL row
 * ids.  Peer access between the first device and the others is required (YRB_ERR_UNSUPPORTED otherwise). */
typedef struct yrb_sharded yrb_sharded;
/* the row map itself (stateless; needs no GPU): where a global row lives, the global row of a shard's local row — the
 * function the merge kernels apply —, and how many rows a shard holds when the collection has total_rows */
YRB_API int yrb_shard_locate(int n_shards, int block_rows, int64_t global_row, int* out_shard, int64_t* out_local);
YRB_API int yrb_shard_global(int n_shards, int block_rows, int shard, int64_t local_row, int64_t* out_global);
YRB_API int yrb_shard_rows(int n_shards, int block_rows, int64_t total_rows, int shard, int64_t* out_rows);
YRB_API int yrb_sharded_create(yrb_sharded** out, const int* devices, int n_devices /* 1..8 */, int dim, int metric,
                               int storage_dtype, int64_t reserve_rows, int block_rows /* 0 = 16384; power of two >= 64 */);
YRB_API int yrb_sharded_destroy(yrb_sharded* sh);
YRB_API int yrb_sharded_count(const yrb_sharded* sh, int64_t* out_rows, int64_t* out_live);
YRB_API int yrb_sharded_info(const yrb_sharded* sh, int* out_dim, int* out_ld, int* out_metric, int* out_dtype,
                             int* out_n_devices, int* out_block_rows, int64_t* out_capacity);
/* shard s as a plain index (profiling, path forcing) and the device it lives on; do not append to it directly */
YRB_API int yrb_sharded_shard(const yrb_sharded* sh, int s, yrb_index** out_index, int* out_device);
YRB_API int yrb_sharded_append_host_f32(yrb_sharded* sh, const float* rows, int64_t n);
/* rows resident on `src_device` (any device of the box; pieces for other shards travel over NVLink); the producer
 * of dev_rows must have finished */
YRB_API int yrb_sharded_append_device_f32(yrb_sharded* sh, const float* dev_rows, int64_t n, int src_device);
YRB_API int yrb_sharded_read_rows(yrb_sharded* sh, const int64_t* row_ids, int64_t n, float* out_rows);
YRB_API int yrb_sharded_read_raw(yrb_sharded* sh, int64_t row_begin, int64_t n, void* out_rows, float* out_sqnorm);
YRB_API int yrb_sharded_append_raw(yrb_sharded* sh, const void* rows, const float* sqnorm, int64_t n);
YRB_API int yrb_sharded_set_live(yrb_sharded* sh, const int64_t* row_ids, int64_t n, int live);
YRB_API int yrb_sharded_truncate(yrb_sharded* sh, int64_t rows);
YRB_API int yrb_sharded_clear(yrb_sharded* sh);
YRB_API int yrb_sharded_column_write(yrb_sharded* sh, int col, int col_type, int64_t row_begin, int64_t n,
                                     const void* values, const uint8_t* present);
YRB_API int yrb_sharded_where(yrb_sharded* sh, const yrb_where* w, uint32_t* out_mask, int64_t* out_pass);
YRB_API int yrb_sharded_search(yrb_sharded* sh, const float* queries, int nq, int k, const yrb_where* w,
                               const uint32_t* mask, int64_t* out_ids, float* out_scores, int32_t* out_counts);
YRB_API int yrb_sharded_search_multi(yrb_sharded* sh, const float* queries, int nq, int k, const yrb_where* const* wheres,
                                     int64_t* out_ids, float* out_scores, int32_t* out_counts);
YRB_API int yrb_sharded_search_ex(yrb_sharded* sh, const float* queries, int nq, int k, const yrb_where* w,
                                  const yrb_where* const* wheres, const uint32_t* mask, const yrb_search_opts* opts,
                                  int64_t* out_ids, float* out_scores, int32_t* out_counts);
YRB_API int yrb_sharded_stats(const yrb_sharded* sh, int64_t* out_kernel_launches, int64_t* out_searches);

/* Force a kernel family for tests/bench: 0 auto, 1 K1 (GEMV + in-register top-k),
 * 2 K2 (tcgen05 GEMM + fused top-k epilogue), 3 K6 (key vector + radix select),
 * 4 K2 with 129..256-query chunks on the CTA-pair (cta_group::2) kernel. */
YRB_API int yrb_index_set_path(yrb_index* ix, int path);
/* Leave `n` SMs out of the persistent scan / GEMM grids (default 0).  A sharded searcher sets 2 so that the
 * exchange kernel of the previous search finds a free SM while the next scan runs (measured: 2 GPUs,
 * 5.1 k → 5.9 k QPS on C2). */
YRB_API int yrb_index_set_reserved_sms(yrb_index* ix, int n);
/* launches issued by this index since creation (bench `gpu_launches`), and the duration in ms of
 * the dominant kernel of the last search measured with CUDA events when enabled. */
YRB_API int yrb_index_stats(const yrb_index* ix, int64_t* out_kernel_launches);
/* Filter caches: a `where` program evaluated over unchanged rows / tombstones / columns reuses its bitmask (K4 skipped)
 * and, for batches, the gathered rows of a selective filter (K8 skipped: agents repeat the same knowledge-base filter,
 * kb_search_toolkit.py:63-96).  Any append, delete or metadata write invalidates both.  Counters since creation. */
YRB_API int yrb_index_cache_stats(const yrb_index* ix, int64_t* out_filter_hits, int64_t* out_compaction_hits);
/* CUDA-event timing of the dominant kernel (K1 scan / K2 GEMM) of every search issued while
 * enabled: events are recorded on the launching stream around that kernel only.  read() waits for
 * the recorded events, returns the summed duration and launch count since the last read, resets. */
YRB_API int yrb_index_profile(yrb_index* ix, int enable);
YRB_API int yrb_index_profile_read(yrb_index* ix, double* out_total_ms, int64_t* out_launches);

/* selection key = (monotone(score) << 32) | ~row : larger key = better (score desc, row asc) */
static inline uint32_t yrb_key_row(uint64_t key) { return ~(uint32_t)(key & 0xffffffffu); }
static inline float yrb_key_score(uint64_t key) {
    uint32_t m = (uint32_t)(key >> 32);
    uint32_t u = (m & 0x80000000u) ? (m & 0x7fffffffu) : ~m;
    union { uint32_t u; float f; } c;
    c.u = u;
    return c.f;
}

#ifdef __cplusplus
}
#endif
#endif /* YRB200_H */
