// common.cuh — shared device helpers: selection keys, warp utilities, block bitonic sort.
// sm_100a only.  No reference counterpart (the reference has no native code).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define YRB_WARP 32
#define YRB_FULL 0xffffffffu

namespace yrb {

// ---------------------------------------------------------------- selection keys
// key = (monotone(score) << 32) | ~row.  A larger key is a better hit: higher score first,
// lower row id on equal scores (the reference-side ordering pinned in DESIGN.md §3).
// key 0 is reserved for "empty slot": it is below every key a real row can produce
// (rows < 2^32-1, and monotone(score)==0 only for one NaN pattern).
__device__ __forceinline__ uint32_t score_bits(float s) {
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float bits_score(uint32_t m) {
    uint32_t u = (m & 0x80000000u) ? (m & 0x7fffffffu) : ~m;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float s, uint32_t row) {
    return ((uint64_t)score_bits(s) << 32) | (uint64_t)(~row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~(uint32_t)k; }
__device__ __forceinline__ float key_score(uint64_t k) { return bits_score((uint32_t)(k >> 32)); }

// ---------------------------------------------------------------- loads
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(YRB_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(YRB_FULL, v, o);
    return v;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(YRB_FULL, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(YRB_FULL, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int d) {
    uint32_t lo = __shfl_up_sync(YRB_FULL, (uint32_t)v, d);
    uint32_t hi = __shfl_up_sync(YRB_FULL, (uint32_t)(v >> 32), d);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(YRB_FULL, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(YRB_FULL, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

// ---------------------------------------------------------------- warp-distributed sorted list
// A descending list of 32*KPL keys spread over a warp: entry e lives in lane e%32, slot e/32.
// insert() is executed by all lanes with the same `key` (warp-uniform), shifting worse entries
// down by one and dropping the last.
template <int KPL>
struct WarpList {
    uint64_t v[KPL];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int s = 0; s < KPL; ++s) v[s] = 0;
    }
    __device__ __forceinline__ void insert(uint64_t key, int lane) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < KPL; ++s) pos += __popc(__ballot_sync(YRB_FULL, v[s] > key));
#pragma unroll
        for (int s = KPL - 1; s >= 0; --s) {
            uint64_t up = shfl_up_u64(v[s], 1);
            if (s > 0) {
                uint64_t wrap = shfl_u64(v[s - 1], 31);
                if (lane == 0) up = wrap;
            }
            int e = s * 32 + lane;
            v[s] = (e > pos) ? up : ((e == pos) ? key : v[s]);
        }
    }
    // key of entry e (warp-uniform e)
    __device__ __forceinline__ uint64_t at(int e) const {
        uint64_t r = 0;
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            uint64_t t = shfl_u64(v[s], e & 31);
            if ((e >> 5) == s) r = t;
        }
        return r;
    }
};

// ---------------------------------------------------------------- block bitonic sort (descending)
// Sorts n (power of two) elements of shared memory so that better() elements come first.
// All threads of the block must call it; it begins and ends with __syncthreads().
template <typename T, typename Better>
__device__ __forceinline__ void block_bitonic_desc(T* s, int n, Better better) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                T a = s[lo], b = s[hi];
                bool swap = desc ? better(b, a) : better(a, b);
                if (swap) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

struct BetterU64 {
    __device__ __forceinline__ bool operator()(uint64_t a, uint64_t b) const { return a > b; }
};

__host__ __device__ __forceinline__ int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace yrb
