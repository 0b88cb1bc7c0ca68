// k2_batched.h — host interface of K2 (batched tcgen05 GEMM + fused top-k epilogue).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <string>

namespace yrb {

struct K2State;  // driver entry point for tensor-map encoding + candidate buffers of one index
K2State* k2_create();
void k2_destroy(K2State* s);
int k2_parts(int sm_count);      // sorted k-lists K2 leaves per query before K3
bool k2_supported(int dtype, int dim, int k);
// q: prepared bf16 queries [nq, ld].  Writes nq*k keys (descending) to out_keys.  Returns YRB_* code.
int k2_search(K2State* s, const void* rows, int64_t n_rows, int64_t capacity, int dim, int ld, const void* q, int nq,
              int k, const uint32_t* mask, int64_t mask_q_stride /*words between queries' masks, 0 = shared*/, int metric,
              const float* q_sqnorm, const float* row_sqnorm,
              uint64_t* out_keys, int64_t* out_ids, float* out_scores, int32_t* out_counts, int sm_count, cudaStream_t st, int* launches, std::string& err,
              cudaEvent_t ev_start = nullptr, cudaEvent_t ev_stop = nullptr,  // recorded around the main GEMM launch
              bool force_pair = false,  // 129..256-query chunks on the cta_group::2 kernel
              // row-subset mode: `rows` is the whole matrix of matrix_rows rows, n_rows counts the entries of
              // rowmap (padded to a multiple of 128 with valid rows); returned keys carry compact indices
              const uint32_t* rowmap = nullptr, int64_t matrix_rows = 0,
              const struct XShard* xs = nullptr,   // sharded collection: K3 hands each query's keys to the cross-shard merge
              float min_score = -INFINITY);        // keep hits with score >= min_score (base_retriever.py:71)

// pair kernel (k2_pair.cu)
}  // namespace yrb
#include <cuda.h>
namespace yrb {
cudaError_t launch_gemm_pair(int grid, const CUtensorMap& mq, const CUtensorMap& mr, int64_t n_rows, int kblocks, int iters,
                             int nq, int k, const uint32_t* mask, int64_t mask_q_stride, const float* thr, uint64_t* ck,
                             int* cc, float* tops, int m_tops, const float* q_sqnorm, const float* row_sqnorm,
                             float* thr_out, unsigned int* sync_ctr, float score_floor, int nqc, int q_stride, cudaStream_t st);
}  // namespace yrb
