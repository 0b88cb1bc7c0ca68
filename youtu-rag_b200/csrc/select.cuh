// select.cuh — block-level top-k selection over unsorted candidate keys (see k3_select.cu).
#pragma once

#include "common.cuh"
#include "xshard.cuh"

namespace yrb {

constexpr int SEL_THREADS = 512;
constexpr int SEL_BINS = 2048;
constexpr int SEL_KMAX = 256;
constexpr int SEL_MAX_SEG = 256;  // segments per query (one per CTA of the producing kernel)

struct SelectArgs {
    const uint64_t* base;
    int64_t seg_stride, q_stride;     // key offset of (seg, q) = seg*seg_stride + q*q_stride
    const int* counts;                // NULL → fixed_cnt
    int64_t cnt_seg_stride, cnt_q_stride;
    int n_seg, fixed_cnt, seg_cap;
    const float* thr;                 // optional per-query filter: keep keys with score > thr[q]
    int k;
    uint64_t* out;                    // [nq][k] sorted keys (may be NULL when ids/scores are given)
    int64_t* ids;                     // optional decoded outputs [nq][k] / [nq]
    float* scores;
    int32_t* out_counts;
    int use_xs = 0;                   // sharded collection: hand the sorted keys to the cross-shard merge instead
    XShard xs{};
};

__device__ __forceinline__ void select_emit(const SelectArgs& a, int q, int i, uint64_t key) {
    if (a.out) a.out[(int64_t)q * a.k + i] = key;
    if (a.ids) a.ids[(int64_t)q * a.k + i] = key ? (int64_t)key_row(key) : -1;
    if (a.scores) a.scores[(int64_t)q * a.k + i] = key ? key_score(key) : -INFINITY;
}

__device__ __forceinline__ int seg_count(const SelectArgs& a, int seg, int q) {
    int c = a.counts ? a.counts[seg * a.cnt_seg_stride + q * a.cnt_q_stride] : a.fixed_cnt;
    return c > a.seg_cap ? a.seg_cap : c;
}

// shared-memory scratch of one selecting CTA (placed after the staged keys in dynamic shared memory)
struct SelectScratch {
    int hist[SEL_BINS];
    uint64_t sel[SEL_KMAX];
    uint64_t red[SEL_THREADS / 32 * 2];
    int seg_cnt[SEL_MAX_SEG];
    int seg_off[SEL_MAX_SEG + 1];  // exclusive prefix of seg_cnt (dense staging)
    int s_n, s_nsel, s_bstar, s_above, s_zero;
};
__host__ __device__ constexpr size_t select_smem_bytes(int stage_keys) {
    return (size_t)stage_keys * 8 + sizeof(SelectScratch);
}

// Executed by all SEL_THREADS threads of a CTA.  sk: `stage` staged-key slots; scratch: SelectScratch.
// Writes a.out[q*k .. q*k+k) (descending keys, 0-padded).  Ends with every thread past its last barrier.
__device__ __forceinline__ void select_topk_block(const SelectArgs& a, const int q, uint64_t* sk, const int stage,
                                                  SelectScratch& S) {
    int (&hist)[SEL_BINS] = S.hist;
    uint64_t (&sel)[SEL_KMAX] = S.sel;
    uint64_t (&red)[SEL_THREADS / 32 * 2] = S.red;
    int &s_n = S.s_n, &s_nsel = S.s_nsel, &s_bstar = S.s_bstar, &s_above = S.s_above;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k;
    const float thr = a.thr ? a.thr[q] : -INFINITY;
    if (tid == 0) {
        s_n = 0;
        s_nsel = 0;
    }
    __syncthreads();

    // ---- stage (filtered) candidates; count them even when they do not fit.
    for (int seg = tid; seg < a.n_seg; seg += SEL_THREADS) S.seg_cnt[seg] = seg_count(a, seg, q);
    if (tid == 0) S.s_zero = 0;
    __syncthreads();
    bool staged = false;
    if (a.thr == nullptr) {
        // Dense staging: a prefix sum over the segment lengths numbers the candidates, thread t takes candidates
        // t, t+512, … (segment by binary search in the prefix), four loads in flight per thread.  The loop below
        // this one walks each segment with a ballot/atomic per step — one L2 round trip per step, ~28 of them
        // for C3's ~5600 candidates in 148 segments — and is kept for filtered selections.
        if (warp == 0) {
            int loc[SEL_MAX_SEG / 32], sum = 0;
#pragma unroll
            for (int j = 0; j < SEL_MAX_SEG / 32; ++j) {
                const int sg = lane * (SEL_MAX_SEG / 32) + j;
                loc[j] = sg < a.n_seg ? S.seg_cnt[sg] : 0;
                sum += loc[j];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(YRB_FULL, incl, o);
                if (lane >= o) incl += v;
            }
            int run = incl - sum;
#pragma unroll
            for (int j = 0; j < SEL_MAX_SEG / 32; ++j) {
                S.seg_off[lane * (SEL_MAX_SEG / 32) + j] = run;
                run += loc[j];
            }
            if (lane == 31) S.seg_off[SEL_MAX_SEG] = incl;
        }
        __syncthreads();
        const int total = S.seg_off[SEL_MAX_SEG];
        if (total <= stage) {
            const uint64_t* qbase = a.base + (int64_t)q * a.q_stride;
            bool zero = false;
            for (int i0 = tid; i0 < total; i0 += 4 * SEL_THREADS) {
                uint64_t key[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * SEL_THREADS;
                    key[u] = 1ull;
                    if (i < total) {
                        int lo = 0, hi = SEL_MAX_SEG - 1;  // largest sg with seg_off[sg] <= i (empty segments share offsets)
#pragma unroll
                        for (int step = 0; step < 8; ++step) {
                            const int mid = (lo + hi + 1) >> 1;
                            if (S.seg_off[mid] <= i) lo = mid; else hi = mid - 1;
                        }
                        key[u] = qbase[(int64_t)lo * a.seg_stride + (i - S.seg_off[lo])];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * SEL_THREADS;
                    if (i < total) sk[i] = key[u];
                    zero |= key[u] == 0ull;
                }
            }
            if (zero) S.s_zero = 1;  // an empty key inside a segment: recount with the compacting loop
            if (tid == 0) s_n = total;
            __syncthreads();
            staged = S.s_zero == 0;
            if (!staged && tid == 0) s_n = 0;
            __syncthreads();
        }
    }
    // Compacting staging: four threads share a segment so that all segment reads are in flight together.
    for (int seg0 = 0; !staged && seg0 < a.n_seg; seg0 += SEL_THREADS / 4) {
        const int seg = seg0 + (tid >> 2), r = tid & 3;
        const int c = seg < a.n_seg ? S.seg_cnt[seg] : 0;
        const uint64_t* src = a.base + (int64_t)seg * a.seg_stride + (int64_t)q * a.q_stride;
        int cmax = c;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cmax = max(cmax, __shfl_xor_sync(YRB_FULL, cmax, o));
        for (int i = r; i < cmax + r; i += 4) {   // same trip count on every lane of the warp
            uint64_t key = (i < c) ? src[i] : 0ull;
            const bool keep = key != 0ull && (a.thr == nullptr || key_score(key) > thr);
            const unsigned m = __ballot_sync(YRB_FULL, keep);
            if (m) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_n, __popc(m));
                pos = __shfl_sync(YRB_FULL, pos, 0) + __popc(m & ((1u << lane) - 1u));
                if (keep && pos < stage) sk[pos] = key;
            }
        }
    }
    __syncthreads();
    const int n_all = s_n;

    if (n_all > stage) {
        // ---- fallback: chunked bitonic merging straight from the segments (slow, exact)
        int have = 0;
        int seg = 0;
        const int stage2 = 1 << (31 - __clz(stage));  // the bitonic passes below pad to a power of two: stay inside the stage
        while (seg < a.n_seg) {
            __syncthreads();
            if (tid == 0) s_n = have;
            __syncthreads();
            int seg_end = seg;
            int budget = have;
            while (seg_end < a.n_seg && budget + seg_count(a, seg_end, q) <= stage2) {
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
        if (a.use_xs) {
            __syncthreads();
            for (int i = tid; i < have; i += SEL_THREADS) sel[i] = sk[i];
            xshard_finish(a.xs, q, sel, have, sk);
            return;
        }
        for (int i = tid; i < k; i += SEL_THREADS) select_emit(a, q, i, (i < have) ? sk[i] : 0ull);
        if (tid == 0 && a.out_counts) a.out_counts[q] = have;
        return;
    }

    const int n = n_all;
    if (n <= SEL_THREADS) {
        // ---- few candidates (small collections, tight bounds): rank by counting — candidate t's output slot is the
        // number of keys above it (keys are unique), one pass over shared memory, no histogram rounds, no sort.
        // A 16-query batch over 10k rows spent 16 us per selection in the general path for ~30 candidates.
        const uint64_t mine = tid < n ? sk[tid] : 0ull;
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += sk[j] > mine;
        __syncthreads();                      // every thread has read sk; sel (and, for xshard, sk) may be written now
        const int cnt = n < k ? n : k;
        if (tid < n && rank < k) sel[rank] = mine;
        __syncthreads();
        if (a.use_xs) {
            xshard_finish(a.xs, q, sel, cnt, sk);
            return;
        }
        for (int i = tid; i < k; i += SEL_THREADS) select_emit(a, q, i, (i < cnt) ? sel[i] : 0ull);
        if (tid == 0 && a.out_counts) a.out_counts[q] = cnt;
        return;
    }
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
    if (a.use_xs) {
        xshard_finish(a.xs, q, sel, nsel, sk);
        return;
    }
    for (int i = tid; i < k; i += SEL_THREADS) select_emit(a, q, i, (i < nsel) ? sel[i] : 0ull);
    if (tid == 0 && a.out_counts) a.out_counts[q] = nsel;
}


}  // namespace yrb
