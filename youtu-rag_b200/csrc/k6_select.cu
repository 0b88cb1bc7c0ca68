// K6 — any-k selection (k > 128, up to 4096): the scan writes one 64-bit key per row (0 for rows the
// filter hides) and a multi-CTA MSB-first radix select on the keys finds the exact k-th best key in
// six 11-bit rounds; rows at or above it are collected and bitonic-sorted.
//
// This is the large-n_results form of collection.query (chroma_store.py:118-120).  Keys are unique
// (the row id is in the key) so the k-th key is exact and ties need no special handling.
// HBM-bound and launch-bound (15 small launches); bytes: N·ld·s (scan) + 8N (keys) × 8 passes.
#include "common.cuh"
#include "kernels.h"

namespace yrb {

constexpr int K6_BINS = 2048;
constexpr int K6_KMAX = 4096;

struct K6State {
    unsigned long long prefix;  // high bits of the k-th key found so far
    int need;                   // how many keys are still wanted inside the current prefix
    int total;                  // number of visible rows (non-zero keys)
    int out_count;              // keys collected
    int hist[K6_BINS];
};

size_t select_scratch_bytes(int64_t n_rows, int k) {
    (void)n_rows;
    (void)k;
    return sizeof(K6State) + (size_t)K6_KMAX * 8;
}

// round r handles key bits [shift, shift+bits); keys must match `prefix` above shift+bits
__global__ void __launch_bounds__(256) k6_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int bits,
                                                      K6State* st) {
    __shared__ int sh[K6_BINS];
    for (int i = threadIdx.x; i < K6_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix;
    const int hi = shift + bits;
    const unsigned long long bmask = (1ull << bits) - 1ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key == 0ull) continue;
        if (hi < 64 && (key >> hi) != (prefix >> hi)) continue;
        atomicAdd(&sh[(int)((key >> shift) & bmask)], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K6_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&st->hist[i], sh[i]);
}

// single CTA: find the bin that crosses `need` from the top, fold it into the prefix, clear the histogram
__global__ void __launch_bounds__(1024) k6_pick_kernel(K6State* st, int shift, int first_round, int k) {
    __shared__ int suf[K6_BINS + 1];
    const int tid = threadIdx.x;
    // inclusive suffix sums: suf[b] = sum_{j >= b} hist[j] (Hillis-Steele over 2048 bins, 2 per thread)
    for (int b = tid; b < K6_BINS; b += blockDim.x) suf[b] = st->hist[b];
    if (tid == 0) suf[K6_BINS] = 0;
    __syncthreads();
    for (int off = 1; off < K6_BINS; off <<= 1) {
        int add[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int b = tid + j * 1024;
            add[j] = (b + off < K6_BINS) ? suf[b + off] : 0;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) suf[tid + j * 1024] += add[j];
        __syncthreads();
    }
    if (tid == 0 && first_round) {
        st->total = suf[0];
        st->need = suf[0] < k ? suf[0] : k;
    }
    __syncthreads();
    const int need = st->need;
    __syncthreads();  // everyone holds `need` before one thread rewrites it
    if (need > 0) {
        for (int b = tid; b < K6_BINS; b += blockDim.x) {
            if (suf[b] >= need && suf[b + 1] < need) {
                st->prefix |= ((unsigned long long)b << shift);
                st->need = need - suf[b + 1];
            }
        }
    }
    __syncthreads();
    for (int b = tid; b < K6_BINS; b += blockDim.x) st->hist[b] = 0;
}

__global__ void __launch_bounds__(256) k6_collect_kernel(const uint64_t* __restrict__ keys, int64_t n, K6State* st,
                                                         uint64_t* __restrict__ out, int cap) {
    if (st->total == 0) return;
    const unsigned long long T = st->prefix;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        if (key != 0ull && key >= T) {
            const int pos = atomicAdd(&st->out_count, 1);
            if (pos < cap) out[pos] = key;
        }
    }
}

__global__ void __launch_bounds__(1024) k6_sort_kernel(const uint64_t* __restrict__ in, const K6State* st, int k,
                                                       uint64_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sraw[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(sraw);
    int n = st->out_count;
    n = n > K6_KMAX ? K6_KMAX : n;
    const int npow = next_pow2(n > 1 ? n : 2);
    for (int i = threadIdx.x; i < npow; i += blockDim.x) sk[i] = i < n ? in[i] : 0ull;
    block_bitonic_desc(sk, npow, BetterU64());
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = i < n ? sk[i] : 0ull;
}

cudaError_t launch_select(const uint64_t* keys, int64_t n_rows, int k, uint64_t* out_keys, void* scratch, int sm_count,
                          cudaStream_t st) {
    if (k < 1 || k > K6_KMAX) return cudaErrorInvalidValue;
    K6State* state = reinterpret_cast<K6State*>(scratch);
    uint64_t* collected = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(scratch) + sizeof(K6State));
    cudaError_t e = cudaMemsetAsync(state, 0, sizeof(K6State), st);
    if (e != cudaSuccess) return e;
    const int grid = sm_count * 4;
    static const int shifts[6] = {53, 42, 31, 20, 9, 0};
    static const int widths[6] = {11, 11, 11, 11, 11, 9};
    for (int r = 0; r < 6; ++r) {
        k6_hist_kernel<<<grid, 256, 0, st>>>(keys, n_rows, shifts[r], widths[r], state);
        k6_pick_kernel<<<1, 1024, 0, st>>>(state, shifts[r], r == 0, k);
    }
    k6_collect_kernel<<<grid, 256, 0, st>>>(keys, n_rows, state, collected, K6_KMAX);
    e = cudaFuncSetAttribute(k6_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K6_KMAX * 8);
    if (e != cudaSuccess) return e;
    k6_sort_kernel<<<1, 1024, K6_KMAX * 8, st>>>(collected, state, k, out_keys);
    return cudaGetLastError();
}

}  // namespace yrb
