// K3 (selection form) — sorted top-k of each query's UNSORTED candidate keys, gathered from many
// segments (K1: one k-list per CTA; K2: one variable-length buffer per (CTA, query)).
//
// One CTA per query.  Candidates are staged in shared memory (up to SEL_STAGE keys), then a radix-style
// narrowing on the 64-bit keys finds the k best without sorting everything: histogram of
// (key - lo) >> shift over 2048 bins, keep every bin above the one that crosses k, recurse into
// that bin.  Keys are unique (the row id is part of the key) so the recursion ends; at most
// ceil(64/11) rounds.  The k survivors are bitonic-sorted.  More than SEL_STAGE candidates (adversarial
// inputs only) fall back to chunked bitonic merging in the same CTA.
// No reference counterpart (chromadb's HNSW keeps its own heap).
#include "common.cuh"
#include "kernels.h"
#include "select.cuh"

namespace yrb {

// staged keys (192 KiB): the sampled bound leaves k * N / (rows sampled) candidates per query — 5-7k for the 1M-row
// configurations, 10-20k for 10M-row shards and small k; beyond the stage the exact but slow chunked merge runs
// (measured: 168 us instead of ~30 us per 256-query launch on a 1.25M-row shard with an 8192-key stage)
constexpr int SEL_STAGE = 24576;

__global__ void __launch_bounds__(SEL_THREADS) select_segments_kernel(SelectArgs a) {
    extern __shared__ __align__(16) unsigned char sraw[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(sraw);
    SelectScratch& S = *reinterpret_cast<SelectScratch*>(sraw + (size_t)SEL_STAGE * 8);
    select_topk_block(a, blockIdx.x, sk, SEL_STAGE, S);
}

cudaError_t launch_select_segments(const uint64_t* base, int64_t seg_stride, int64_t q_stride, const int* counts,
                                   int64_t cnt_seg_stride, int64_t cnt_q_stride, int n_seg, int fixed_cnt, int seg_cap,
                                   const float* thr, int nq, int k, uint64_t* out, cudaStream_t st, int64_t* ids,
                                   float* scores, int32_t* counts_out, const XShard* xs) {
    if (k < 1 || k > SEL_KMAX || n_seg > SEL_MAX_SEG) return cudaErrorInvalidValue;
    const size_t smem = select_smem_bytes(SEL_STAGE);
    cudaError_t e = cudaFuncSetAttribute(select_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    SelectArgs a{base, seg_stride, q_stride, counts, cnt_seg_stride, cnt_q_stride, n_seg, fixed_cnt, seg_cap, thr, k, out,
                 ids, scores, counts_out};
    if (xs) {
        a.use_xs = 1;
        a.xs = *xs;
    }
    select_segments_kernel<<<nq, SEL_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

// the cross-shard finish on its own (xshard.cuh): one CTA per query
__global__ void __launch_bounds__(256) xshard_finish_kernel(XShard x, const uint64_t* __restrict__ local_keys, int k_local) {
    extern __shared__ __align__(16) unsigned char sraw[];
    uint64_t* stage = reinterpret_cast<uint64_t*>(sraw);
    const uint64_t* mine = local_keys + (size_t)blockIdx.x * k_local;
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int n = 0;
    for (int i = threadIdx.x; i < k_local; i += blockDim.x) n += mine[i] != 0ull;
    if (n) atomicAdd(&s_n, n);
    __syncthreads();
    xshard_finish(x, blockIdx.x, mine, s_n, stage);
}

cudaError_t launch_xshard_finish(const XShard& xs, const uint64_t* local_keys, int nq, int k_local, cudaStream_t st) {
    const size_t smem = (size_t)xs.n_shards * xs.k * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(xshard_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    xshard_finish_kernel<<<nq, 256, smem, st>>>(xs, local_keys, k_local);
    return cudaGetLastError();
}

}  // namespace yrb
