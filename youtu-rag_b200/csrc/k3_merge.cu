// K3 (merge form) — the local step of the cross-GPU merge after the NCCL all-gather (the baseline
// exchange; K7 in k7_exchange.cu does transfer + merge in one kernel), and the key decoder of the K6 path.
// SURVEY.md §8e; no reference counterpart — the reference is one Chroma collection in one process,
// chroma_store.py:41-59.  Per-GPU selection over unsorted candidates lives in k3_select.cu / select.cuh.
//
// One CTA per query: world*k candidates are loaded into shared memory, sorted with a block bitonic
// network on (score desc, GLOBAL id asc) and the first k written back.  KiB-scale, latency-bound.
#include "common.cuh"
#include "kernels.h"

namespace yrb {

constexpr int MERGE_THREADS = 512;

__global__ void decode_kernel(const uint64_t* __restrict__ keys, int nq, int k, int64_t* __restrict__ ids,
                              float* __restrict__ scores, int32_t* __restrict__ counts) {
    const int q = blockIdx.x;
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = keys[(int64_t)q * k + i];
        const bool ok = key != 0;
        ids[(int64_t)q * k + i] = ok ? (int64_t)key_row(key) : -1;
        scores[(int64_t)q * k + i] = ok ? key_score(key) : -INFINITY;
        local += ok;
    }
    if (local) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0 && counts) counts[q] = cnt;
}

cudaError_t launch_decode(const uint64_t* keys, int nq, int k, int64_t* ids, float* scores, int32_t* counts,
                          cudaStream_t st) {
    decode_kernel<<<nq, 128, 0, st>>>(keys, nq, k, ids, scores, counts);
    return cudaGetLastError();
}

// ---- cross-rank merge on (score, global id): lists come from row shards, part p holds global
// rows [row_base[p], row_base[p+1]) so the tie-break must use the GLOBAL id.
struct GKey {
    uint32_t sbits;  // monotone score bits
    uint32_t valid;
    int64_t gid;
};
struct BetterG {
    __device__ __forceinline__ bool operator()(const GKey& a, const GKey& b) const {
        if (a.valid != b.valid) return a.valid > b.valid;
        if (a.sbits != b.sbits) return a.sbits > b.sbits;
        return a.gid < b.gid;
    }
};
constexpr int GMERGE_CAP = 2048;

__global__ void __launch_bounds__(MERGE_THREADS)
    merge_global_kernel(const uint64_t* __restrict__ in, int parts, int nq, int k, const int64_t* __restrict__ row_base,
                        int64_t* __restrict__ ids, float* __restrict__ scores, int32_t* __restrict__ counts) {
    __shared__ GKey sk[GMERGE_CAP];
    __shared__ int cnt;
    const int q = blockIdx.x;
    const int n = parts * k;
    const int npow = next_pow2(n);
    if (threadIdx.x == 0) cnt = 0;
    for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        GKey g;
        g.sbits = 0;
        g.valid = 0;
        g.gid = 0;
        if (i < n) {
            const int p = i / k, j = i - p * k;
            const uint64_t key = in[((int64_t)p * nq + q) * k + j];
            if (key != 0) {
                g.sbits = (uint32_t)(key >> 32);
                g.valid = 1;
                g.gid = row_base[p] + (int64_t)key_row(key);
            }
        }
        sk[i] = g;
    }
    block_bitonic_desc(sk, npow, BetterG());
    int local = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = (i < n) && sk[i].valid;
        ids[(int64_t)q * k + i] = ok ? sk[i].gid : -1;
        scores[(int64_t)q * k + i] = ok ? bits_score(sk[i].sbits) : -INFINITY;
        local += ok;
    }
    if (local) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0 && counts) counts[q] = cnt;
}

cudaError_t launch_merge_global(const uint64_t* in, int parts, int nq, int k, const int64_t* row_base, int64_t* ids,
                                float* scores, int32_t* counts, cudaStream_t st) {
    if ((int64_t)parts * k > GMERGE_CAP) return cudaErrorInvalidValue;
    merge_global_kernel<<<nq, MERGE_THREADS, 0, st>>>(in, parts, nq, k, row_base, ids, scores, counts);
    return cudaGetLastError();
}

}  // namespace yrb
