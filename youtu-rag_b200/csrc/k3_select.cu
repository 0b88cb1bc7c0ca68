// K3 (selection form) — sorted top-k of each query's UNSORTED candidate keys, gathered from many
// segments (K1: one k-list per CTA; K2: one variable-length buffer per (CTA, query)).
//
// One CTA per query.  Candidates are staged in shared memory (up to 8192 keys), then a radix-style
// narrowing on the 64-bit keys finds the k best without sorting everything: histogram of
// (key - lo) >> shift over 2048 bins, keep every bin above the one that crosses k, recurse into
// that bin.  Keys are unique (the row id is part of the key) so the recursion ends; at most
// ceil(64/11) rounds.  The k survivors are bitonic-sorted.  More than 8192 candidates (adversarial
// inputs only) fall back to chunked bitonic merging in the same CTA.
// No reference counterpart (chromadb's HNSW keeps its own heap).
#include "common.cuh"
#include "kernels.h"

namespace yrb {

constexpr int SEL_THREADS = 512;
constexpr int SEL_STAGE = 8192;   // staged keys (64 KiB)
constexpr int SEL_BINS = 2048;
constexpr int SEL_KMAX = 256;

struct SelectArgs {
    const uint64_t* base;
    int64_t seg_stride, q_stride;     // key offset of (seg, q) = seg*seg_stride + q*q_stride
    const int* counts;                // NULL → fixed_cnt
    int64_t cnt_seg_stride, cnt_q_stride;
    int n_seg, fixed_cnt, seg_cap;
    const float* thr;                 // optional per-query filter: keep keys with score > thr[q]
    int k;
    uint64_t* out;                    // [nq][k]
};

__device__ __forceinline__ int seg_count(const SelectArgs& a, int seg, int q) {
    int c = a.counts ? a.counts[seg * a.cnt_seg_stride + q * a.cnt_q_stride] : a.fixed_cnt;
    return c > a.seg_cap ? a.seg_cap : c;
}

__global__ void __launch_bounds__(SEL_THREADS) select_segments_kernel(SelectArgs a) {
    extern __shared__ __align__(16) unsigned char sraw[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(sraw);      // SEL_STAGE keys
    __shared__ int hist[SEL_BINS];
    __shared__ uint64_t sel[SEL_KMAX];
    __shared__ uint64_t red[SEL_THREADS / 32 * 2];
    __shared__ int s_n, s_nsel, s_bstar, s_above;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k;
    const float thr = a.thr ? a.thr[q] : -INFINITY;
    if (tid == 0) {
        s_n = 0;
        s_nsel = 0;
    }
    __syncthreads();

    // ---- stage (filtered) candidates; count them even when they do not fit
    for (int seg = warp; seg < a.n_seg; seg += SEL_THREADS / 32) {
        const int c = seg_count(a, seg, q);
        const uint64_t* src = a.base + seg * a.seg_stride + q * a.q_stride;
        for (int i0 = 0; i0 < c; i0 += 32) {
            const int i = i0 + lane;
            uint64_t key = (i < c) ? src[i] : 0ull;
            const bool keep = key != 0ull && (a.thr == nullptr || key_score(key) > thr);
            const unsigned m = __ballot_sync(YRB_FULL, keep);
            if (m) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_n, __popc(m));
                pos = __shfl_sync(YRB_FULL, pos, 0) + __popc(m & ((1u << lane) - 1u));
                if (keep && pos < SEL_STAGE) sk[pos] = key;
            }
        }
    }
    __syncthreads();
    const int n_all = s_n;

    if (n_all > SEL_STAGE) {
        // ---- fallback: chunked bitonic merging straight from the segments (slow, exact)
        int have = 0;
        int seg = 0;
        while (seg < a.n_seg) {
            __syncthreads();
            if (tid == 0) s_n = have;
            __syncthreads();
            int seg_end = seg;
            int budget = have;
            while (seg_end < a.n_seg && budget + seg_count(a, seg_end, q) <= SEL_STAGE) {
                budget += seg_count(a, seg_end, q);
                ++seg_end;
            }
            for (int s2 = seg + warp; s2 < seg_end; s2 += SEL_THREADS / 32) {
                const int c = seg_count(a, s2, q);
                const uint64_t* src = a.base + s2 * a.seg_stride + q * a.q_stride;
                for (int i0 = 0; i0 < c; i0 += 32) {
                    const int i = i0 + lane;
                    uint64_t key = (i < c) ? src[i] : 0ull;
                    const bool keep = key != 0ull && (a.thr == nullptr || key_score(key) > thr);
                    const unsigned m = __ballot_sync(YRB_FULL, keep);
                    if (m) {
                        int pos = 0;
                        if (lane == 0) pos = atomicAdd(&s_n, __popc(m));
                        pos = __shfl_sync(YRB_FULL, pos, 0) + __popc(m & ((1u << lane) - 1u));
                        if (keep) sk[pos] = key;
                    }
                }
            }
            __syncthreads();
            const int total = s_n;
            const int npow = next_pow2(total > 1 ? total : 2);
            for (int i = total + tid; i < npow; i += SEL_THREADS) sk[i] = 0ull;
            block_bitonic_desc(sk, npow, BetterU64());
            have = total < k ? total : k;
            seg = seg_end;
        }
        for (int i = tid; i < k; i += SEL_THREADS) a.out[(int64_t)q * k + i] = (i < have) ? sk[i] : 0ull;
        return;
    }

    const int n = n_all;
    int need = n < k ? n : k;
    if (n > k) {
        // ---- key range
        uint64_t lo = ~0ull, hi = 0ull;
        for (int i = tid; i < n; i += SEL_THREADS) {
            const uint64_t x = sk[i];
            lo = x < lo ? x : lo;
            hi = x > hi ? x : hi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t l2 = shfl_xor_u64(lo, o), h2 = shfl_xor_u64(hi, o);
            lo = l2 < lo ? l2 : lo;
            hi = h2 > hi ? h2 : hi;
        }
        if (lane == 0) {
            red[warp * 2] = lo;
            red[warp * 2 + 1] = hi;
        }
        __syncthreads();
        for (int w = 0; w < SEL_THREADS / 32; ++w) {
            lo = red[w * 2] < lo ? red[w * 2] : lo;
            hi = red[w * 2 + 1] > hi ? red[w * 2 + 1] : hi;
        }
        // ---- narrowing rounds
        while (true) {
            const uint64_t span = hi - lo;
            const int bl = 64 - __clzll((long long)(span | 1ull));
            const int shift = bl > 11 ? bl - 11 : 0;
            for (int i = tid; i < SEL_BINS; i += SEL_THREADS) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += SEL_THREADS) {
                const uint64_t x = sk[i];
                if (x >= lo && x <= hi) atomicAdd(&hist[(int)((x - lo) >> shift)], 1);
            }
            __syncthreads();
            if (warp == 0) {
                // suffix scan from the top bin: lane L owns bins [64L, 64L+64)
                int mine = 0;
                for (int b = 0; b < 64; ++b) mine += hist[lane * 64 + ((b + lane) & 63)];  // rotated: no bank conflicts
                int above = 0;  // candidates in bins owned by higher lanes
                for (int L = 31; L >= 0; --L) {
                    const int v = __shfl_sync(YRB_FULL, mine, L);
                    if (L > lane) above += v;
                }
                const bool has = (above < need) && (above + mine >= need);
                if (has) {
                    int acc = above;
                    for (int b = 63; b >= 0; --b) {
                        const int h = hist[lane * 64 + b];
                        if (acc + h >= need) {
                            s_bstar = lane * 64 + b;
                            s_above = acc;
                            break;
                        }
                        acc += h;
                    }
                }
            }
            __syncthreads();
            const int bstar = s_bstar, above = s_above;
            const int in_b = hist[bstar];
            const uint64_t blo = lo + ((uint64_t)bstar << shift);
            const uint64_t bhi = (shift == 0) ? blo : (blo + ((1ull << shift) - 1ull));
            const bool take_all = (above + in_b == need);
            for (int i = tid; i < n; i += SEL_THREADS) {
                const uint64_t x = sk[i];
                if (x >= lo && x <= hi && (x > bhi || (take_all && x >= blo))) sel[atomicAdd(&s_nsel, 1)] = x;
            }
            __syncthreads();
            if (take_all) break;
            need -= above;
            lo = blo;
            hi = bhi < hi ? bhi : hi;
        }
    } else {
        for (int i = tid; i < n; i += SEL_THREADS) sel[i] = sk[i];
        if (tid == 0) s_nsel = n;
        __syncthreads();
    }
    const int nsel = s_nsel;  // == min(n, k)
    const int npow = next_pow2(nsel > 1 ? nsel : 2);
    for (int i = nsel + tid; i < npow; i += SEL_THREADS) sel[i] = 0ull;
    block_bitonic_desc(sel, npow, BetterU64());
    for (int i = tid; i < k; i += SEL_THREADS) a.out[(int64_t)q * k + i] = (i < nsel) ? sel[i] : 0ull;
}

cudaError_t launch_select_segments(const uint64_t* base, int64_t seg_stride, int64_t q_stride, const int* counts,
                                   int64_t cnt_seg_stride, int64_t cnt_q_stride, int n_seg, int fixed_cnt, int seg_cap,
                                   const float* thr, int nq, int k, uint64_t* out, cudaStream_t st) {
    if (k < 1 || k > SEL_KMAX) return cudaErrorInvalidValue;
    const size_t smem = (size_t)SEL_STAGE * 8;
    cudaError_t e = cudaFuncSetAttribute(select_segments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    SelectArgs a{base, seg_stride, q_stride, counts, cnt_seg_stride, cnt_q_stride, n_seg, fixed_cnt, seg_cap, thr, k, out};
    select_segments_kernel<<<nq, SEL_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace yrb
