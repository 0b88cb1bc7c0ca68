// K3 — top-k merge: P sorted k-lists per query → one sorted k-list.  Used (a) to merge the
// per-CTA lists of K1/K2 on one GPU and (b) as the local step of the cross-GPU merge collective
// after the NCCL all-gather (no reference counterpart — the reference is one
// Chroma collection in one process, chroma_store.py:41-59).
//
// One CTA per (query, group of lists): candidates are loaded into shared memory, sorted with a
// block bitonic network on 64-bit keys (score desc, row asc) and the first k written back.
// Work is KiB-scale and latency-bound; bytes are negligible next to the scan.
#include "common.cuh"
#include "kernels.h"

namespace yrb {

constexpr int MERGE_THREADS = 512;
constexpr int MERGE_CAP = 4096;  // keys per CTA pass (32 KiB shared)

// in: [parts][nq][k]; CTA (q, grp) merges lists [grp*G, min(parts,(grp+1)*G)) → out[grp][q][k]
__global__ void __launch_bounds__(MERGE_THREADS)
    merge_keys_kernel(const uint64_t* __restrict__ in, int parts, int nq, int k, int G, uint64_t* __restrict__ out) {
    __shared__ uint64_t sk[MERGE_CAP];
    const int q = blockIdx.x, grp = blockIdx.y;
    const int p0 = grp * G;
    const int np = min(G, parts - p0);
    const int n = np * k;
    const int npow = next_pow2(n);
    for (int i = threadIdx.x; i < npow; i += blockDim.x) {
        uint64_t v = 0;
        if (i < n) {
            const int p = p0 + i / k, j = i - (i / k) * k;
            v = in[((int64_t)p * nq + q) * k + j];
        }
        sk[i] = v;
    }
    block_bitonic_desc(sk, npow, BetterU64());
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[((int64_t)grp * nq + q) * k + i] = (i < n) ? sk[i] : 0ull;
}

cudaError_t launch_merge_keys(const uint64_t* in, int parts, int nq, int k, uint64_t* out, uint64_t* scratch,
                              cudaStream_t st) {
    if (k > MERGE_CAP / 2) return cudaErrorInvalidValue;
    const int G = MERGE_CAP / k;
    const uint64_t* cur = in;
    int p = parts;
    // ping-pong inside scratch: pass outputs are at most ceil(p/G) lists
    uint64_t* bufs[2] = {scratch, scratch ? scratch + (int64_t)((parts + G - 1) / G) * nq * k : nullptr};
    int flip = 0;
    while (true) {
        const int groups = (p + G - 1) / G;
        uint64_t* dst = (groups == 1) ? out : bufs[flip];
        if (groups > 1 && !scratch) return cudaErrorInvalidValue;
        merge_keys_kernel<<<dim3(nq, groups), MERGE_THREADS, 0, st>>>(cur, p, nq, k, G, dst);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (groups == 1) break;
        cur = dst;
        p = groups;
        flip ^= 1;
    }
    return cudaSuccess;
}

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
