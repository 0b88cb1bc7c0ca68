// kernels.h — host-side launchers of the hot-path kernels (internal to libyrb200.so).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "xshard.cuh"

namespace yrb {

struct DeviceInfo {
    int device = 0;
    int sm_count = 148;
};

// leading dimension (elements) of stored rows: rows are padded to a multiple of 128 bytes
inline int row_ld(int dim, int dtype /*0 bf16, 1 f32*/) {
    int q = dtype == 0 ? 64 : 32;
    return (dim + q - 1) / q * q;
}
inline int elem_size(int dtype) { return dtype == 0 ? 2 : 4; }

// ---- K5: ingest.  src fp32 [n, dim] (device) → dst rows [n, ld] in storage dtype; cosine rows are
// L2-normalised with an fp64 norm.  sqnorm[n] = ||stored row||^2 in fp32 (used by euclidean).
cudaError_t launch_ingest(const float* src, int64_t n, int dim, int ld, int metric, int dtype,
                          void* dst, float* sqnorm, cudaStream_t st);

// ---- K1: single-query scan + in-register top-k.  q = ONE prepared query [ld] in storage dtype.
// part_keys must hold k1_parts(sm) * k keys.  Writes per-CTA sorted top-k lists.
int k1_parts(int sm_count);
// q_raw: ONE fp32 query [dim] on the device, prepared (normalise + round to the storage dtype, same
// arithmetic as K5) in the kernel prologue.  When parts*k fits the in-kernel selection, the last CTA to
// finish also merges the per-CTA lists: final_keys[k] (and ids/scores/count when given) are complete when
// the launch is; *fused tells the caller.  Otherwise the caller runs launch_select_segments on part_keys.
struct K1Out {
    uint64_t* final_keys;  // [k]
    int64_t* ids;          // optional [k]
    float* scores;         // optional [k]
    int32_t* count;        // optional [1]
    unsigned long long* trace = nullptr;  // optional [grid][8] phase stamps (globaltimer ns), YRB_K1_TRACE
    int use_xs = 0;                       // sharded collection: the last CTA hands the keys to the cross-shard merge
    XShard xs{};
    uint64_t floor_key = 0;               // score threshold as a key: only keys above it are hits (0 = none)
};
// key every qualifying hit exceeds: score >= min_score  <=>  key > score_floor_key(min_score); -inf / NaN → 0 (no threshold)
uint64_t score_floor_key(float min_score);
cudaError_t launch_k1(const void* rows, int dtype, int64_t n_rows, int dim, int ld, const float* q_raw,
                      const float* row_sqnorm, int metric, const uint32_t* mask, int k, uint64_t* part_keys,
                      unsigned int* ticket, K1Out out, bool* fused, int sm_count, cudaStream_t st,
                      int* parts_out = nullptr);  // CTAs launched = per-CTA lists written (<= k1_parts(sm_count))

// K1Q: up to 4 prepared fp32 queries ([nq][ld] fp32 + squared norms, from launch_ingest) against fp32 rows in one
// pass; one shared mask; k <= 32.  Writes per-CTA sorted lists part_keys[q][cta][k]; finish with
// launch_select_segments(part_keys, k, parts*k, nullptr, 0, 0, parts, k, k, nullptr, nq, k, …).
constexpr int K1Q_MAX_Q = 4, K1Q_MAX_K = 32;
cudaError_t launch_k1q_f32(const void* rows, int64_t n_rows, int ld, const float* q_prep, int nq, const float* q_sqn,
                           const float* row_sqnorm, int metric, const uint32_t* mask, int k, uint64_t* part_keys,
                           int sm_count, cudaStream_t st, uint64_t floor_key = 0);

// ---- K3: selection / merge
// sorted top-k of unsorted candidates gathered from n_seg segments per query (see k3_select.cu):
// key pointer of (seg, q) = base + seg*seg_stride + q*q_stride; its length = counts[seg*cnt_seg_stride +
// q*cnt_q_stride] (or fixed_cnt), clamped to seg_cap; zero keys are skipped; thr (optional) keeps
// only keys whose score > thr[q].  k <= 256.
cudaError_t launch_select_segments(const uint64_t* base, int64_t seg_stride, int64_t q_stride, const int* counts,
                                   int64_t cnt_seg_stride, int64_t cnt_q_stride, int n_seg, int fixed_cnt, int seg_cap,
                                   const float* thr, int nq, int k, uint64_t* out, cudaStream_t st,
                                   int64_t* ids = nullptr, float* scores = nullptr, int32_t* counts_out = nullptr,
                                   const XShard* xs = nullptr);
// cross-shard finish as its own launch (paths whose keys are complete only after another kernel: K6, K8 remap):
// local_keys [nq][k_local] sorted, 0-padded
cudaError_t launch_xshard_finish(const XShard& xs, const uint64_t* local_keys, int nq, int k_local, cudaStream_t st);
// keys [nq][k] → ids / scores / counts
cudaError_t launch_decode(const uint64_t* keys, int nq, int k, int64_t* ids, float* scores,
                          int32_t* counts, cudaStream_t st);
// cross-rank merge: [parts][nq][k] local keys + row_base[parts] → global ids (score desc, id asc)
cudaError_t launch_merge_global(const uint64_t* in, int parts, int nq, int k,
                                const int64_t* row_base, int64_t* ids, float* scores,
                                int32_t* counts, cudaStream_t st);

// ---- K4: where-program evaluation → bitmask
struct WhereLeafDev {
    const void* values;
    const uint32_t* present;  // bitmask, never NULL for col >= 0
    int32_t col_type;         // YRB_COL_*, -1 = absent column
    int32_t op;
    int32_t operand_begin;
    int32_t operand_count;
};
struct WhereProgDev {
    int32_t n_leaves;
    int32_t n_postfix;
    WhereLeafDev leaves[64];
    int64_t operands[256];
    int32_t postfix[160];
};
// prog: device copy of WhereProgDev.  live / extra may be NULL.  out_mask: ceil(n/32) words,
// padded words up to an even count are zeroed.  pass_count: device int64, accumulated (+=).
cudaError_t launch_where(const WhereProgDev* prog, int64_t n_rows, const uint32_t* live,
                         const uint32_t* extra, uint32_t* out_mask, unsigned long long* pass_count,
                         cudaStream_t st);
cudaError_t launch_mask_and(const uint32_t* a, const uint32_t* b, int64_t n_words, uint32_t* out,
                            cudaStream_t st);

// ---- K8: row compaction for low-selectivity batched search (k8_compact.cu)
size_t compact_scratch_words(int64_t n_rows);
// block_sums gets the exclusive prefix of passing rows per 32768-row block, *total their number
cudaError_t launch_compact_count(const uint32_t* mask, int64_t n_rows, uint32_t* block_sums, unsigned long long* total,
                                 cudaStream_t st);
cudaError_t launch_compact_gather(const uint32_t* mask, int64_t n_rows, const uint32_t* block_off, uint32_t* rowmap,
                                  const void* rows, const float* sqnorm, int ld16, int64_t n_out, void* out_rows,
                                  float* out_sqnorm, int sm_count, cudaStream_t st);
cudaError_t launch_compact_remap(const uint32_t* rowmap, int64_t n, uint64_t* keys, int64_t* ids, cudaStream_t st);

// ---- K6: any-k path (k <= 4096).  One 64-bit key per row for one query + radix select of the top k.
cudaError_t launch_scores(const void* rows, int dtype, int64_t n_rows, int dim, int ld, const float* q_raw,
                          const float* row_sqnorm, int metric,
                          const uint32_t* mask, uint64_t* keys_out, int sm_count, cudaStream_t st, uint64_t floor_key = 0);
size_t select_scratch_bytes(int64_t n_rows, int k);
cudaError_t launch_select(const uint64_t* keys, int64_t n_rows, int k, uint64_t* out_keys,
                          void* scratch, int sm_count, cudaStream_t st);

// ---- K2: batched tcgen05 GEMM + fused top-k epilogue (bf16 storage only).



}  // namespace yrb
