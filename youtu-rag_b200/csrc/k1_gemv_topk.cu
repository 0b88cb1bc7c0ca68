// K1 — single-query scan: HBM-bound GEMV + in-register running top-k + bitmask, one pass.
//
// Stands in for the inner loop of collection.query(query_embeddings=[q], n_results=k, where=…)
// (utu/rag/storage/implementations/chroma_store.py:118-120) / faiss IndexFlat.search
// (faiss_store.py:154), computed exactly.
//
// Shape of the work: every stored row is read once with 128-bit coalesced
// streaming loads (a warp covers 512 contiguous bytes per instruction, R rows × 4 chunks = 16
// loads in flight per lane); the query lives in shared memory as fp32; a row's score is a
// warp butterfly sum; each warp keeps a sorted top-k of 64-bit keys spread over its lanes and
// only touches it when a score beats the current k-th key.  Rows failing the bitmask are never
// loaded.  The 16 per-warp lists of a CTA, and then the per-CTA lists in the last CTA to finish, are merged
// by pruning + ranking (merge_sorted_lists, k <= 32) or a tournament + K3's block selection (k <= 128), so a
// whole search is one launch.  K1Q (further down) shares a pass between four fp32 queries; k6_scores writes
// the plain key vector for the any-k path.
// Algorithmic bytes per search: sel·N·ld·esize + N/8 (mask) + ld·esize (query) + parts·k·8.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "select.cuh"

namespace yrb {

constexpr int K1_THREADS = 512;
constexpr int K1_WARPS = K1_THREADS / 32;
constexpr int K1_R = 4;   // rows per batch
constexpr int K1_CU = 4;  // chunks per unrolled step

int k1_parts(int sm_count) { return sm_count; }

__device__ __forceinline__ unsigned long long k1_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define K1_STAMP(i)                                                                  \
    do {                                                                             \
        if (out.trace && threadIdx.x == 0) out.trace[blockIdx.x * 8 + (i)] = k1_now(); \
    } while (0)

template <bool F32>
__device__ __forceinline__ float dot_chunk(uint4 v, float4 qa, float4 qb, float acc) {
    if (F32) {
        acc = fmaf(__uint_as_float(v.x), qa.x, acc);
        acc = fmaf(__uint_as_float(v.y), qa.y, acc);
        acc = fmaf(__uint_as_float(v.z), qa.z, acc);
        acc = fmaf(__uint_as_float(v.w), qa.w, acc);
    } else {
        acc = fmaf(__uint_as_float(v.x << 16), qa.x, acc);
        acc = fmaf(__uint_as_float(v.x & 0xffff0000u), qa.y, acc);
        acc = fmaf(__uint_as_float(v.y << 16), qa.z, acc);
        acc = fmaf(__uint_as_float(v.y & 0xffff0000u), qa.w, acc);
        acc = fmaf(__uint_as_float(v.z << 16), qb.x, acc);
        acc = fmaf(__uint_as_float(v.z & 0xffff0000u), qb.y, acc);
        acc = fmaf(__uint_as_float(v.w << 16), qb.z, acc);
        acc = fmaf(__uint_as_float(v.w & 0xffff0000u), qb.w, acc);
    }
    return acc;
}

// Query preparation fused into the scan's prologue (same arithmetic as K5, csrc/k5_ingest.cu, so K1 and
// K2 see bit-identical queries): every warp forms the fp64 sum of squares in K5's lane-strided order,
// then the CTA writes the normalised, storage-rounded query to shared memory as fp32, laid out
// [chunk][half][lane] float4 so that a warp's LDS.128 is conflict-free.  Returns ||stored q||^2 (fp32,
// K5's order) when `want_sq`.
template <bool F32>
__device__ __forceinline__ float prepare_query(const float* __restrict__ q_raw, int dim, int ld, int nch, int normalize,
                                               bool want_sq, float4* sq) {
    constexpr int H = F32 ? 1 : 2;
    constexpr int EPL = F32 ? 4 : 8;
    const int lane = threadIdx.x & 31;
    // K5 has two forms (k5_ingest.cu): the streaming one adds a lane's squares float4 by float4, the generic one
    // element by element; follow whichever K5 would take for this query so that both kernels see the same bits
    const bool vec = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(q_raw) & 15) == 0) && ((ld + 127) / 128 <= 16);
    double ss = 0.0;
    if (normalize) {
        if (vec) {
            const float4* q4 = reinterpret_cast<const float4*>(q_raw);
            for (int j = lane; j < (dim >> 2); j += 32) {
                const float4 v = q4[j];
                ss += (double)v.x * (double)v.x;
                ss += (double)v.y * (double)v.y;
                ss += (double)v.z * (double)v.z;
                ss += (double)v.w * (double)v.w;
            }
        } else {
            for (int i = lane; i < dim; i += 32) {
                const double v = (double)q_raw[i];
                ss += v * v;
            }
        }
        ss = warp_sum(ss);
    }
    const bool scale = normalize && ss > 0.0;
    const double inv = scale ? 1.0 / sqrt(ss) : 1.0;   // K5's arithmetic: multiply by the fp64 reciprocal of the norm
    auto stored = [&](int e) -> float {
        float y = 0.f;
        if (e < dim) y = scale ? (float)((double)q_raw[e] * inv) : q_raw[e];
        return F32 ? y : __bfloat162float(__float2bfloat16_rn(y));
    };
    float sqn = 0.f;
    if (want_sq) {
        if (vec) {
            for (int j = lane; j < (ld >> 2); j += 32) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float y = stored(4 * j + c);
                    sqn = fmaf(y, y, sqn);
                }
            }
        } else {
            for (int i = lane; i < ld; i += 32) {
                const float y = stored(i);
                sqn = fmaf(y, y, sqn);
            }
        }
        sqn = warp_sum(sqn);
    }
    // one float4 (4 elements) per thread: the fp64 divides are spread over the whole CTA
    for (int i = threadIdx.x; i < nch * 32 * H; i += blockDim.x) {
        const int item = i / H, h = i - item * H;  // item = chunk * 32 + lane
        const int c = item >> 5, l = item & 31;
        const int e0 = item * EPL + h * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (e0 + j < ld) ? stored(e0 + j) : 0.f;
        sq[(c * H + h) * 32 + l] = make_float4(v[0], v[1], v[2], v[3]);
    }
    return sqn;
}

// one unrolled step of K1_R rows x K1_CU chunks: the loads …
__device__ __forceinline__ void rows_load_step(const uint4* const (&rp)[K1_R], const bool (&va)[K1_R], int ld16, int c0,
                                               int lane, uint4 (&v)[K1_R][K1_CU]) {
#pragma unroll
    for (int cc = 0; cc < K1_CU; ++cc) {
        const int off = (c0 + cc) * 32 + lane;
        const bool in = off < ld16;
#pragma unroll
        for (int r = 0; r < K1_R; ++r) {
            v[r][cc] = make_uint4(0, 0, 0, 0);
            if (in && va[r]) v[r][cc] = ldg_stream(rp[r] + off);
        }
    }
}
// … and the multiply-adds against the staged query
template <bool F32>
__device__ __forceinline__ void rows_fma_step(const uint4 (&v)[K1_R][K1_CU], int c0, int nch, const float4* sq, int lane,
                                              float (&acc)[K1_R]) {
    constexpr int H = F32 ? 1 : 2;
#pragma unroll
    for (int cc = 0; cc < K1_CU; ++cc) {
        const int c = c0 + cc;
        if (c < nch) {
            const float4 qa = sq[(c * H) * 32 + lane];
            const float4 qb = F32 ? qa : sq[(c * H + 1) * 32 + lane];
#pragma unroll
            for (int r = 0; r < K1_R; ++r) acc[r] = dot_chunk<F32>(v[r][cc], qa, qb, acc[r]);
        }
    }
}

// dot products of K1_R rows (row pointers rp[], warp-uniform validity va[]) with the staged query;
// every lane returns the full sums.
template <bool F32>
__device__ __forceinline__ void rows_dot(const uint4* const (&rp)[K1_R], const bool (&va)[K1_R], int ld16, int nch,
                                         const float4* sq, int lane, float (&acc)[K1_R]) {
#pragma unroll
    for (int r = 0; r < K1_R; ++r) acc[r] = 0.f;
    for (int c0 = 0; c0 < nch; c0 += K1_CU) {
        uint4 v[K1_R][K1_CU];
        rows_load_step(rp, va, ld16, c0, lane, v);
        rows_fma_step<F32>(v, c0, nch, sq, lane, acc);
    }
#pragma unroll
    for (int r = 0; r < K1_R; ++r) acc[r] = warp_sum(acc[r]);
}

// Exact top-k of `n_lists` (<= 256) descending-sorted, 0-padded key lists in shared memory (list l at
// sk[l*stride .. +k)), k <= 32, by pruning + ranking instead of a serial k-step tournament.  No atomics, every
// step spread over the whole CTA:
//   1. bound: with depth = ceil(k / n_lists) and m = ceil(k / depth), the m-th largest of the lists' depth-th
//      entries (T) has at least m*depth >= k keys at or above it.  The entries are ranked by counting, each
//      entry's comparisons split over up to 4 threads.  Keys are unique, so the ranks are a permutation.
//   2. n_lists >= k (depth 1): only the <= k lists whose head reaches T can hold a result; list of rank r is
//      copied to row r of a k x k matrix M, so step 3 scans k*k keys instead of n_lists*k.  Otherwise
//      (n_lists < k) the lists themselves are the rows.  Rows are sorted: a row's survivors (keys >= T) are a prefix.
//   3. a prefix sum over the rows' survivor counts numbers the survivors; warp w takes survivors w, w+16, …, and
//      counts the keys above each one (lanes over the matrix, redux): that count is its output slot.
// Called by every thread of the CTA; emit(slot, key) runs exactly once per slot in [0, k) (key 0 = empty).
struct MergeScratch {
    unsigned long long T;
    int hot[32];                     // list holding the r-th largest head (step 2), -1 = none
    int cnt[32];                     // survivors per row
    unsigned short partial[4][256];  // partial ranks of step 1
};
template <class Emit>
__device__ __forceinline__ int merge_sorted_lists(const uint64_t* sk, int n_lists, int stride, int k, uint64_t* M,
                                                  MergeScratch* ms, Emit emit, unsigned long long* dbg = nullptr) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int depth = (k + n_lists - 1) / n_lists;
    const int m = (k + depth - 1) / depth;
    const int d1 = depth - 1;
    const bool copy = depth == 1;
    int P = nt / n_lists;
    P = P > 4 ? 4 : (P < 1 ? 1 : P);
    const int chunk = (n_lists + P - 1) / P;
    if (tid == 0) ms->T = 0ull;
    if (tid < 32) {
        ms->hot[tid] = -1;
        ms->cnt[tid] = 0;
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[1] = k1_now();
    for (int t = tid; t < n_lists * P; t += nt) {  // 1a. thread (entry i, part p): comparisons against one slice
        const int p = t / n_lists, i = t - p * n_lists;
        const uint64_t h = sk[i * stride + d1];
        const int j1 = (p + 1) * chunk < n_lists ? (p + 1) * chunk : n_lists;
        int c = 0;
        for (int j = p * chunk; j < j1; ++j) c += sk[j * stride + d1] > h;
        ms->partial[p][i] = (unsigned short)c;
    }
    __syncthreads();
    for (int i = tid; i < n_lists; i += nt) {  // 1b. ranks; the bound; the lists that can reach it
        const uint64_t h = sk[i * stride + d1];
        if (h != 0ull) {
            int c = 0;
            for (int p = 0; p < P; ++p) c += ms->partial[p][i];
            if (c < m) {
                if (c == m - 1) ms->T = h;
                if (copy) ms->hot[c] = i;
            }
        }
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[2] = k1_now();
    const uint64_t T = ms->T;
    const uint64_t* src = copy ? M : sk;
    const int rows = copy ? k : n_lists, rstride = copy ? k : stride;  // rows <= 32 either way
    for (int r = warp; r < rows; r += nw) {  // 2. rows and their survivor counts
        uint64_t mine = 0ull;
        if (copy) {
            const int l = ms->hot[r];
            if (l >= 0 && lane < k) mine = sk[l * stride + lane];
            if (lane < k) M[r * k + lane] = mine;
        } else if (lane < k) {
            mine = sk[r * stride + lane];
        }
        const unsigned int live = __ballot_sync(YRB_FULL, mine != 0ull && mine >= T);
        if (lane == 0) ms->cnt[r] = __popc(live);
    }
    __syncthreads();
    if (dbg && tid == 0) dbg[3] = k1_now();
    const int myc = lane < rows ? ms->cnt[lane] : 0;
    int incl = myc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(YRB_FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int ns = __shfl_sync(YRB_FULL, incl, 31);
    const int span = rows * rstride;
    for (int sv = warp; sv < ns; sv += nw) {  // 3. survivor sv: row = number of rows that end at or before it
        const int r = __popc(__ballot_sync(YRB_FULL, incl <= sv));
        const int excl = __shfl_sync(YRB_FULL, incl - myc, r);
        const uint64_t key = src[r * rstride + (sv - excl)];
        int c = 0;
        for (int j = lane; j < span; j += 32) c += src[j] > key;
        c = __reduce_add_sync(YRB_FULL, c);
        if (lane == 0 && c < k) emit(c, key);
    }
    for (int r = ns + tid; r < k; r += nt) emit(r, 0ull);
    if (dbg && tid == 0) {
        dbg[4] = k1_now();
        dbg[6] = (unsigned long long)ns;
    }
    return ns < k ? ns : k;
}
constexpr int K1_RANK_K = 32;  // largest k merged by ranking

template <bool F32, int KPL, bool HAS_MASK>
__global__ void __launch_bounds__(K1_THREADS, 1)
    k1_scan_topk(const uint4* __restrict__ rows, int64_t n_rows, int dim, int ld, int ld16, int nch,
                 const float* __restrict__ q_raw, int normalize, const float* __restrict__ row_sqnorm, int l2,
                 const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ part_keys, unsigned int* ticket,
                 K1Out out, int fuse_stage) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sq = reinterpret_cast<float4*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    K1_STAMP(0);

    const int64_t gw = (int64_t)blockIdx.x * K1_WARPS + warp;
    const int64_t tw = (int64_t)gridDim.x * K1_WARPS;
    const float q_sqn = prepare_query<F32>(q_raw, dim, ld, nch, normalize, l2 != 0, sq);
    __syncthreads();
    K1_STAMP(1);
    const float l2_bias = l2 ? (1.f - q_sqn) : 0.f;

    WarpList<KPL> list;
    list.clear();
    // a score threshold (base_retriever.py:71: keep hits with score >= threshold) is the initial bound of every warp's
    // list: rows below it never enter a list, so the merges and the count see only qualifying hits
    const uint64_t floor_key = out.floor_key;
    uint64_t thr = floor_key;

    auto offer = [&](const int64_t (&rid)[K1_R], const bool (&va)[K1_R], const float (&acc)[K1_R]) {
#pragma unroll
        for (int r = 0; r < K1_R; ++r) {
            if (va[r]) {
                float s = acc[r];
                if (l2) s = fmaf(2.f, s, l2_bias - row_sqnorm[rid[r]]);
                const uint64_t key = make_key(s, (uint32_t)rid[r]);
                if (key > thr) {
                    list.insert(key, lane);
                    const uint64_t kth = list.at(k - 1);
                    thr = kth > floor_key ? kth : floor_key;
                }
            }
        }
    };
    auto consume = [&](const int64_t (&rid)[K1_R], const bool (&va)[K1_R]) {
        const uint4* rp[K1_R];
#pragma unroll
        for (int r = 0; r < K1_R; ++r) rp[r] = rows + rid[r] * (int64_t)ld16;
        float acc[K1_R];
        rows_dot<F32>(rp, va, ld16, nch, sq, lane, acc);
        offer(rid, va, acc);
    };

    // Masked scans hand out 64-row mask groups dynamically (first group = the warp's global index, further
    // ones from an atomic counter): the number of passing rows per group varies, so static dealing would
    // leave some warps with more work.
    // the atomic is issued when a chunk starts and its result is only broadcast when the chunk is done,
    // so its round trip hides behind the chunk's loads
    auto grab_issue = [&]() -> unsigned int { return lane == 0 ? atomicAdd(ticket + 1, 1u) : 0u; };
    auto grab_get = [&](unsigned int c) -> int64_t { return (int64_t)__shfl_sync(YRB_FULL, c, 0) + tw; };
    if (!HAS_MASK) {
        // dense: groups of K1_R rows dealt round-robin to all warps of the grid, so every warp sweeps the same
        // moving window of the matrix (best DRAM locality; 64-row dynamic chunks measured 8-20 % slower, their
        // quantisation costs more than it saves).  SMs do not all get the same share of the bandwidth, though
        // (CTAs of a 125k-row scan finished between 37 and 45 us), so the last eighth of the groups is handed
        // out by tickets of two groups.  A warp draws its first ticket at kernel start and the next one before it
        // consumes the current one, so the atomic's round trip is never waited for.  (Measured alternatives, all
        // slower: tickets drawn late, one-group tickets — more same-address atomics in the tail —, and requesting
        // the first group's rows before the query is prepared, which delays the query's own loads.)
        const int64_t n_groups4 = (n_rows + K1_R - 1) / K1_R;
        const int64_t n_static = (n_groups4 - n_groups4 / 8) / tw * tw;  // whole sweeps
        unsigned int nx = grab_issue();
        auto do_group = [&](int64_t g) {
            int64_t rid[K1_R];
            bool va[K1_R];
#pragma unroll
            for (int r = 0; r < K1_R; ++r) {
                rid[r] = g * K1_R + r;
                va[r] = rid[r] < n_rows;
                if (!va[r]) rid[r] = 0;
            }
            consume(rid, va);
        };
        for (int64_t g = gw; g < n_static; g += tw) {
            do_group(g);
            if (out.trace && g == gw && threadIdx.x == 0) out.trace[blockIdx.x * 8 + 7] = k1_now();  // first group done
        }
        for (;;) {
            const int64_t g = n_static + 2 * (int64_t)__shfl_sync(YRB_FULL, nx, 0);
            if (g >= n_groups4) break;
            nx = grab_issue();
            do_group(g);
            if (g + 1 < n_groups4) do_group(g + 1);
        }
    } else {
        // Passing rows are queued across 64-row mask groups so that every batch carries K1_R rows:
        // at 10 % selectivity a group holds ~6 rows and per-group batching would leave loads idle.
        const int64_t n_groups = (n_rows + 63) / 64;
        const uint2* mask2 = reinterpret_cast<const uint2*>(mask);
        int64_t rid[K1_R];
        bool va[K1_R];
#pragma unroll
        for (int r = 0; r < K1_R; ++r) {
            rid[r] = 0;
            va[r] = false;
        }
        int have = 0;  // warp-uniform
        int64_t g = gw;
        while (g < n_groups) {
            const unsigned int nx = grab_issue();
            const uint2 mw = mask2[g];
            uint64_t m = ((uint64_t)mw.y << 32) | mw.x;
            while (m) {
                const int64_t row = g * 64 + (__ffsll((long long)m) - 1);
                m &= m - 1;
                if (row >= n_rows) continue;
#pragma unroll
                for (int r = 0; r < K1_R; ++r)
                    if (r == have) {
                        rid[r] = row;
                        va[r] = true;
                    }
                if (++have == K1_R) {
                    consume(rid, va);
                    have = 0;
#pragma unroll
                    for (int r = 0; r < K1_R; ++r) va[r] = false;
                }
            }
            g = grab_get(nx);
        }
        if (have) consume(rid, va);
    }

    // ---- CTA merge: every warp's first k entries → shared → tournament → first k to global
    if (out.trace && lane == 0 && warp == 0) out.trace[blockIdx.x * 8 + 6] = k1_now();  // warp 0 left the scan
    __syncthreads();  // query no longer needed; shared memory is reused for keys
    K1_STAMP(2);
    uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
    const int kp = next_pow2(k);  // ≤ 32*KPL
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
        const int e = s * 32 + lane;
        if (e < kp) sk[warp * kp + e] = (e < k) ? list.v[s] : 0ull;
    }
    __syncthreads();
    if constexpr (KPL == 1) {
        // k <= 32: prune + rank over the 16 sorted per-warp lists, all threads (see merge_sorted_lists)
        uint64_t* M = sk + K1_WARPS * kp;  // k*k <= 16*kp keys when the lists are copied (k <= 16)
        MergeScratch* ms = reinterpret_cast<MergeScratch*>(M + K1_WARPS * kp);
        uint64_t* dst = part_keys + (int64_t)blockIdx.x * k;
        unsigned long long* dbg = (out.trace && blockIdx.x == 0) ? out.trace + (size_t)(gridDim.x + 1) * 8 : nullptr;
        if (dbg && threadIdx.x == 0) dbg[0] = k1_now();
        merge_sorted_lists(sk, K1_WARPS, kp, k, M, ms, [&](int slot, uint64_t key) { dst[slot] = key; }, dbg);
    } else if (warp == 0) {
        // k > 32: warp 0 merges the lists with a k-step tournament (lane l < 16 holds the head of list l, a
        // butterfly max picks the winner, whose lane advances)
        int pos = 0;
        uint64_t head = (lane < K1_WARPS) ? sk[lane * kp] : 0ull;
        for (int i = 0; i < k; ++i) {
            uint64_t best = head;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint64_t other = shfl_xor_u64(best, o);
                best = other > best ? other : best;
            }
            if (lane == 0) part_keys[(int64_t)blockIdx.x * k + i] = best;
            if (best != 0ull && head == best) {   // keys are unique: exactly one lane advances
                ++pos;
                head = (pos < k) ? sk[lane * kp + pos] : 0ull;
            }
        }
    }

    // ---- fused K3: the last CTA to arrive merges all per-CTA lists (no second launch)
    {
        __shared__ int s_last;
        __syncthreads();  // the CTA's part_keys stores are issued
        K1_STAMP(3);
        if (threadIdx.x == 0) {
            __threadfence();  // cumulative over the barrier: publishes every thread's stores before the ticket
            const unsigned int t = atomicAdd(ticket, 1u);
            s_last = (t == gridDim.x - 1);
            if (s_last) {
                ticket[0] = 0u;  // every CTA has left the scan: rearm both counters for the next launch
                ticket[1] = 0u;
                __threadfence();
            }
        }
        __syncthreads();
        K1_STAMP(4);
        if (s_last && fuse_stage > 0) {
            SelectArgs a{part_keys, k, 0, nullptr, 0, 0, (int)gridDim.x, k, k, nullptr, k, out.final_keys,
                         out.ids, out.scores, out.count};
            a.use_xs = out.use_xs;
            a.xs = out.xs;
            const int parts = (int)gridDim.x;
            if (KPL == 1 && parts <= 256) {
                // k <= 32: stage the sorted per-CTA lists (one parallel round of L2 reads), then prune + rank
                unsigned long long* dbg = out.trace ? out.trace + (size_t)gridDim.x * 8 : nullptr;
                if (dbg && threadIdx.x == 0) dbg[0] = k1_now();
                for (int i = threadIdx.x; i < parts * k; i += blockDim.x) sk[i] = __ldcg(part_keys + i);
                uint64_t* M = sk + parts * k;
                MergeScratch* ms = reinterpret_cast<MergeScratch*>(M + K1_RANK_K * K1_RANK_K);
                if (a.use_xs) {
                    // sharded collection: the merged list goes to shared memory and from there to the cross-shard merge
                    uint64_t* fin = reinterpret_cast<uint64_t*>(ms + 1);
                    const int n_out = merge_sorted_lists(sk, parts, k, k, M, ms, [&](int slot, uint64_t key) { fin[slot] = key; }, dbg);
                    xshard_finish(a.xs, 0, fin, n_out, sk);
                } else {
                    const int n_out =
                        merge_sorted_lists(sk, parts, k, k, M, ms, [&](int slot, uint64_t key) { select_emit(a, 0, slot, key); }, dbg);
                    if (threadIdx.x == 0 && a.out_counts) a.out_counts[0] = n_out;
                }
            } else {
                SelectScratch& S = *reinterpret_cast<SelectScratch*>(smem_raw + (size_t)fuse_stage * 8);
                select_topk_block(a, 0, sk, fuse_stage, S);
            }
            __syncthreads();
            K1_STAMP(5);
        }
    }
}

// ---- K1Q: fp32-storage batches.  K2 is a bf16 tensor-core kernel; on an fp32 index a batch used to loop K1, one
// scan of the matrix per query.  Here K1Q_NQ = 4 queries share each pass: a row is loaded once and multiplied
// into four accumulators (fp32 rows need one FMA per element and no unpacking, so four queries still leave the
// scan HBM-bound: 16 FMAs per 16 loaded bytes ≈ 20 % of the FMA pipe at the full stream rate).  Same
// per-row arithmetic and summation order as K1, so scores are bit-identical to the single-query path.
// Queries arrive prepared (K5: normalised fp32 rows [nq][ld] + their squared norms); one shared mask; k <= 32.
// Per-CTA lists are merged per query with merge_sorted_lists; the cross-CTA selection is K3's launch.
constexpr int K1Q_NQ = 4;
constexpr int K1Q_CU = 4;  // chunks per unrolled step (4 rows x 4 chunks = 16 loads in flight per lane, as K1)

template <bool HAS_MASK>
__global__ void __launch_bounds__(K1_THREADS, 1)
    k1q_scan_f32(const uint4* __restrict__ rows, int64_t n_rows, int ld16, int nch, const float4* __restrict__ q_prep,
                 int nq, const float* __restrict__ q_sqn, const float* __restrict__ row_sqnorm, int l2,
                 const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ part_keys, uint64_t floor_key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sq = reinterpret_cast<float4*>(smem_raw);  // [K1Q_NQ][nch*32] float4, zero for missing queries / padding
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < K1Q_NQ * nch * 32; i += blockDim.x) {
        const int qi = i / (nch * 32), c = i - qi * (nch * 32);
        sq[i] = (qi < nq && c < ld16) ? q_prep[(int64_t)qi * ld16 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float l2_bias[K1Q_NQ];
#pragma unroll
    for (int qi = 0; qi < K1Q_NQ; ++qi) l2_bias[qi] = (l2 && qi < nq) ? 1.f - q_sqn[qi] : 0.f;
    __syncthreads();

    WarpList<1> list[K1Q_NQ];
    uint64_t thr[K1Q_NQ];
#pragma unroll
    for (int qi = 0; qi < K1Q_NQ; ++qi) {
        list[qi].clear();
        thr[qi] = floor_key;
    }
    const int64_t gw = (int64_t)blockIdx.x * K1_WARPS + warp;
    const int64_t tw = (int64_t)gridDim.x * K1_WARPS;
    const int64_t n_groups4 = (n_rows + K1_R - 1) / K1_R;
    for (int64_t g = gw; g < n_groups4; g += tw) {
        int64_t rid[K1_R];
        bool va[K1_R];
        bool any = false;
#pragma unroll
        for (int r = 0; r < K1_R; ++r) {
            rid[r] = g * K1_R + r;
            va[r] = rid[r] < n_rows;
            if (va[r] && HAS_MASK) va[r] = (mask[rid[r] >> 5] >> (rid[r] & 31)) & 1u;
            if (!va[r]) rid[r] = 0;
            any |= va[r];
        }
        if (!any) continue;  // warp-uniform
        float acc[K1_R][K1Q_NQ];
#pragma unroll
        for (int r = 0; r < K1_R; ++r)
#pragma unroll
            for (int qi = 0; qi < K1Q_NQ; ++qi) acc[r][qi] = 0.f;
        for (int c0 = 0; c0 < nch; c0 += K1Q_CU) {
            uint4 v[K1_R][K1Q_CU];
#pragma unroll
            for (int cc = 0; cc < K1Q_CU; ++cc) {
                const int off = (c0 + cc) * 32 + lane;
                const bool in = off < ld16;
#pragma unroll
                for (int r = 0; r < K1_R; ++r) {
                    v[r][cc] = make_uint4(0, 0, 0, 0);
                    if (in && va[r]) v[r][cc] = ldg_stream(rows + rid[r] * (int64_t)ld16 + off);
                }
            }
#pragma unroll
            for (int cc = 0; cc < K1Q_CU; ++cc) {
                const int c = c0 + cc;
                if (c < nch) {
#pragma unroll
                    for (int qi = 0; qi < K1Q_NQ; ++qi) {
                        const float4 qa = sq[(qi * nch + c) * 32 + lane];
#pragma unroll
                        for (int r = 0; r < K1_R; ++r) acc[r][qi] = dot_chunk<true>(v[r][cc], qa, qa, acc[r][qi]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < K1_R; ++r) {
            const float xs = (l2 && va[r]) ? row_sqnorm[rid[r]] : 0.f;
#pragma unroll
            for (int qi = 0; qi < K1Q_NQ; ++qi) {
                float sc = warp_sum(acc[r][qi]);
                if (va[r] && qi < nq) {
                    if (l2) sc = fmaf(2.f, sc, l2_bias[qi] - xs);
                    const uint64_t key = make_key(sc, (uint32_t)rid[r]);
                    if (key > thr[qi]) {
                        list[qi].insert(key, lane);
                        const uint64_t kth = list[qi].at(k - 1);
                        thr[qi] = kth > floor_key ? kth : floor_key;
                    }
                }
            }
        }
    }

    // per-CTA merge, one query after the other (the query tiles in shared memory are no longer needed)
    __syncthreads();
    uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
    const int kp = next_pow2(k);
    uint64_t* M = sk + K1_WARPS * kp;
    MergeScratch* ms = reinterpret_cast<MergeScratch*>(M + K1_WARPS * kp);
#pragma unroll
    for (int qi = 0; qi < K1Q_NQ; ++qi) {
        if (qi < nq) {  // CTA-uniform
            if (lane < kp) sk[warp * kp + lane] = (lane < k) ? list[qi].v[0] : 0ull;
            __syncthreads();
            uint64_t* dst = part_keys + ((int64_t)qi * gridDim.x + blockIdx.x) * k;
            merge_sorted_lists(sk, K1_WARPS, kp, k, M, ms, [&](int slot, uint64_t key) { dst[slot] = key; });
            __syncthreads();
        }
    }
}

cudaError_t launch_k1q_f32(const void* rows, int64_t n_rows, int ld, const float* q_prep, int nq, const float* q_sqn,
                           const float* row_sqnorm, int metric, const uint32_t* mask, int k, uint64_t* part_keys,
                           int sm_count, cudaStream_t st, uint64_t floor_key) {
    if (nq < 1 || nq > K1Q_NQ || k < 1 || k > K1_RANK_K) return cudaErrorInvalidValue;
    const int ld16 = ld * 4 / 16;
    const int nch = (ld16 + 31) / 32;
    const size_t q_bytes = (size_t)K1Q_NQ * nch * 32 * sizeof(float4);
    const size_t m_bytes = (size_t)K1_WARPS * next_pow2(k) * sizeof(uint64_t) * 2 + sizeof(MergeScratch);
    const size_t smem = std::max(q_bytes, m_bytes);
    const int l2 = (metric == 2);
    const int grid = k1_parts(sm_count);
    auto launch = [&](auto kern) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, K1_THREADS, smem, st>>>(reinterpret_cast<const uint4*>(rows), n_rows, ld16, nch,
                                             reinterpret_cast<const float4*>(q_prep), nq, q_sqn, row_sqnorm, l2, mask, k, part_keys, floor_key);
        return cudaGetLastError();
    };
    return mask ? launch(k1q_scan_f32<true>) : launch(k1q_scan_f32<false>);
}

// ---- K6a: plain scores (any k): score of every row for one query; masked-out rows get key-0
// semantics downstream by writing -inf-like NaN pattern?  No: they are written as -INFINITY and
// the select step also receives the mask.
template <bool F32, bool HAS_MASK>
__global__ void __launch_bounds__(K1_THREADS, 1)
    k6_scores(const uint4* __restrict__ rows, int64_t n_rows, int dim, int ld, int ld16, int nch,
              const float* __restrict__ q_raw, int normalize, const float* __restrict__ row_sqnorm, int l2,
              const uint32_t* __restrict__ mask, uint64_t* __restrict__ keys_out, uint64_t floor_key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sq = reinterpret_cast<float4*>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const float q_sqn = prepare_query<F32>(q_raw, dim, ld, nch, normalize, l2 != 0, sq);
    __syncthreads();
    const float l2_bias = l2 ? (1.f - q_sqn) : 0.f;
    const int64_t gw = (int64_t)blockIdx.x * K1_WARPS + warp;
    const int64_t tw = (int64_t)gridDim.x * K1_WARPS;
    const int64_t n_groups = (n_rows + K1_R - 1) / K1_R;
    for (int64_t g = gw; g < n_groups; g += tw) {
        int64_t rid[K1_R];
        bool va[K1_R];
        const uint4* rp[K1_R];
#pragma unroll
        for (int r = 0; r < K1_R; ++r) {
            rid[r] = g * K1_R + r;
            va[r] = rid[r] < n_rows;
            if (va[r] && HAS_MASK) va[r] = (mask[rid[r] >> 5] >> (rid[r] & 31)) & 1u;
            rp[r] = rows + (va[r] ? rid[r] : 0) * (int64_t)ld16;
        }
        float acc[K1_R];
        rows_dot<F32>(rp, va, ld16, nch, sq, lane, acc);
        if (lane < K1_R) {
            uint64_t key = 0ull;  // rows the filter hides get the empty key
            const int64_t row = g * K1_R + lane;
#pragma unroll
            for (int r = 0; r < K1_R; ++r)
                if (lane == r && va[r])
                    key = make_key(l2 ? fmaf(2.f, acc[r], l2_bias - row_sqnorm[rid[r]]) : acc[r], (uint32_t)row);
            if (row < n_rows) keys_out[row] = key > floor_key ? key : 0ull;   // below the score threshold = hidden
        }
    }
}

static size_t k1_smem_bytes(int dtype, int nch, int k, int fuse_stage) {
    size_t q = (size_t)nch * 32 * (dtype == 1 ? 1 : 2) * sizeof(float4);
    size_t m = (size_t)K1_WARPS * next_pow2(k) * sizeof(uint64_t) * 2 + sizeof(MergeScratch);  // lists + survivors
    size_t f = fuse_stage > 0 ? select_smem_bytes(fuse_stage) : 0;
    if (fuse_stage > 0 && k <= K1_RANK_K)  // staged lists + k x k matrix + scratch of the ranking merge
        f = std::max(f, (size_t)fuse_stage * 8 + (size_t)K1_RANK_K * K1_RANK_K * 8 + sizeof(MergeScratch) + (size_t)K1_RANK_K * 8);
    size_t r = q > m ? q : m;
    return r > f ? r : f;
}

template <bool F32, int KPL, bool HAS_MASK>
static cudaError_t k1_launch_t(const void* rows, int64_t n_rows, int dim, int ld, int ld16, int nch, const float* q_raw,
                               int normalize, const float* row_sqnorm, int l2, const uint32_t* mask, int k,
                               uint64_t* part_keys, unsigned int* ticket, K1Out out, int fuse_stage, int grid,
                               size_t smem, cudaStream_t st) {
    auto kern = k1_scan_topk<F32, KPL, HAS_MASK>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<grid, K1_THREADS, smem, st>>>(reinterpret_cast<const uint4*>(rows), n_rows, dim, ld, ld16, nch, q_raw,
                                          normalize, row_sqnorm, l2, mask, k, part_keys, ticket, out, fuse_stage);
    return cudaGetLastError();
}

cudaError_t launch_k1(const void* rows, int dtype, int64_t n_rows, int dim, int ld, const float* q_raw,
                      const float* row_sqnorm, int metric, const uint32_t* mask, int k, uint64_t* part_keys,
                      unsigned int* ticket, K1Out out, bool* fused, int sm_count, cudaStream_t st, int* parts_out) {
    if (k < 1 || k > 128) return cudaErrorInvalidValue;
    static_assert(K1_THREADS == SEL_THREADS, "the fused selection runs on the scan's CTA");
    const int ld16 = ld * elem_size(dtype) / 16;
    const int nch = (ld16 + 31) / 32;
    // one CTA per SM, but never more CTAs than there are row groups for their warps: a 1000-row collection runs on 16
    // CTAs, whose lists the last one merges in a fraction of the time 148 would take (18.6 us -> see profiles/)
    const int64_t n_groups = mask ? (n_rows + 63) / 64 : (n_rows + K1_R - 1) / K1_R;
    const int grid = (int)std::min<int64_t>(k1_parts(sm_count), std::max<int64_t>(1, (n_groups + K1_WARPS - 1) / K1_WARPS));
    if (parts_out) *parts_out = grid;
    // fuse the final merge when every per-CTA list fits the in-kernel selection's staging area
    const int fuse_stage = (grid * k <= 8192) ? grid * k : 0;
    if (fused) *fused = fuse_stage > 0;
    const size_t smem = k1_smem_bytes(dtype, nch, k, fuse_stage);
    const int l2 = (metric == 2), normalize = (metric == 0);
    const bool f32 = (dtype == 1), big = (k > 32), hm = (mask != nullptr);
#define YRB_K1(F, K, M)                                                                                              \
    return k1_launch_t<F, K, M>(rows, n_rows, dim, ld, ld16, nch, q_raw, normalize, row_sqnorm, l2, mask, k, part_keys, \
                                ticket, out, fuse_stage, grid, smem, st)
    if (!f32 && !big && !hm) YRB_K1(false, 1, false);
    if (!f32 && !big && hm) YRB_K1(false, 1, true);
    if (!f32 && big && !hm) YRB_K1(false, 4, false);
    if (!f32 && big && hm) YRB_K1(false, 4, true);
    if (f32 && !big && !hm) YRB_K1(true, 1, false);
    if (f32 && !big && hm) YRB_K1(true, 1, true);
    if (f32 && big && !hm) YRB_K1(true, 4, false);
    YRB_K1(true, 4, true);
#undef YRB_K1
}

cudaError_t launch_scores(const void* rows, int dtype, int64_t n_rows, int dim, int ld, const float* q_raw,
                          const float* row_sqnorm, int metric, const uint32_t* mask, uint64_t* keys_out, int sm_count,
                          cudaStream_t st, uint64_t floor_key) {
    const int ld16 = ld * elem_size(dtype) / 16;
    const int nch = (ld16 + 31) / 32;
    const size_t smem = (size_t)nch * 32 * (dtype == 1 ? 1 : 2) * sizeof(float4);
    const int l2 = (metric == 2), normalize = (metric == 0);
    const uint4* r4 = reinterpret_cast<const uint4*>(rows);
#define YRB_K6(F, M)                                                                                            \
    {                                                                                                           \
        auto kern = k6_scores<F, M>;                                                                            \
        if (smem > 48 * 1024) {                                                                                 \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                     \
        }                                                                                                       \
        kern<<<sm_count, K1_THREADS, smem, st>>>(r4, n_rows, dim, ld, ld16, nch, q_raw, normalize, row_sqnorm, l2, \
                                                 mask, keys_out, floor_key);                                    \
        return cudaGetLastError();                                                                              \
    }
    if (dtype == 1) {
        if (mask) YRB_K6(true, true) else YRB_K6(true, false)
    } else {
        if (mask) YRB_K6(false, true) else YRB_K6(false, false)
    }
#undef YRB_K6
}

}  // namespace yrb
