// k2_common.cuh — constants, PTX wrappers and the warp sort shared by the batched kernels
// (k2_batched.cu: one CTA per tile; k2_pair.cu: CTA pairs, cta_group::2).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace yrb {
namespace k2 {

constexpr int BLOCK_Q = 128;   // MMA M: queries per query block
constexpr int BLOCK_R = 128;   // MMA N: corpus rows per tile (256 measured slower: no accumulator double-buffering)
constexpr int BLOCK_K = 64;    // bf16 elements per k-block = one 128-byte swizzle atom
constexpr int UMMA_K = 16;
constexpr int CAP = 256;       // candidate slots per (CTA, query)
constexpr int MAX_Q = 256;     // queries per chunk (2 query blocks): one accumulator set
constexpr int MAX_QC = 4;      // chunks the pair kernel sweeps per row tile: up to 1024 queries read the corpus ONCE
constexpr int QTILE_BYTES = BLOCK_Q * BLOCK_K * 2;  // 16 KiB: one query block x one k-block
constexpr int RTILE_BYTES = BLOCK_R * BLOCK_K * 2;  // 16 KiB: one row tile x one k-block
constexpr int MAX_TOPS = 4;   // best scores each (CTA, query) publishes in the sampling pass

__host__ __device__ constexpr int stages(int qb) { return (220 * 1024) / (qb * QTILE_BYTES + RTILE_BYTES); }
// accumulator buffers in the 512 TMEM columns: 2 x (QB x 128) columns, the epilogue of tile i overlaps tile i+1
__host__ __device__ constexpr int acc_buffers(int qb) { return 512 / (qb * BLOCK_R) >= 2 ? 2 : 1; }
__host__ __device__ constexpr int stage_bytes(int qb) { return qb * QTILE_BYTES + RTILE_BYTES; }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug becomes a trap (reported as a CUDA error) instead of a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask`, and each
// destination CTA's barrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                  uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// tile::gather4 — four rows of the 2-D tensor (row coordinates r0..r3, same column c0) land as four consecutive
// box rows in shared memory: lets the GEMM read an arbitrary row subset (the rows passing a filter) in place
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int r0, int r1, int r2,
                                            int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}
// One lane of a converged warp (the compiler keeps warp-uniform operands of the predicated instructions in
// uniform registers; an `if (lane == 0)` around a whole loop makes it re-elect a lane and move five registers to
// the uniform file with R2UR for EVERY tcgen05 / TMA instruction — measured in round 2: the MMA issuer's own
// instruction stream, not a barrier, was what K2 waited for).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the barrier at the same offset in every CTA of `mask` once the MMAs issued so far complete
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
        "%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (version 1 = Blackwell):
// start address >> 4 | LBO(1) << 16 | SBO (8 rows x 128 B = 1024 B >> 4) << 32 | version << 46 | layout(2) << 61
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_R >> 3) << 17) |
                           ((uint32_t)(BLOCK_Q >> 4) << 24);

// ---------------------------------------------------------------- warp bitonic sort of 256 keys
// element e = i*32 + lane, descending.
__device__ __forceinline__ void warp_sort256_desc(uint64_t (&v)[8], int lane) {
#pragma unroll
    for (int size = 2; size <= 256; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int ri = stride >> 5;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if ((i & ri) == 0) {
                        const int e = i * 32 + lane;
                        const bool desc = (e & size) == 0;
                        uint64_t a = v[i], b = v[i | ri];
                        const bool sw = desc ? (b > a) : (a > b);
                        v[i] = sw ? b : a;
                        v[i | ri] = sw ? a : b;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int e = i * 32 + lane;
                    const bool desc = (e & size) == 0;
                    const bool lower = (lane & stride) == 0;
                    const uint64_t o = shfl_xor_u64(v[i], stride);
                    const uint64_t mx = v[i] > o ? v[i] : o, mn = v[i] > o ? o : v[i];
                    v[i] = (lower == desc) ? mx : mn;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- epilogue pieces shared by both kernels
// An epilogue thread owns ONE query (TMEM lane); a chunk is 32 consecutive corpus rows (TMEM columns).
// With one epilogue warp per scheduler every branch latency is exposed, and a compare-and-branch per score
// made the epilogue, not the tensor pipe, the pace of the CTA-pair kernel (round-2 ncu: the MMA issuer waited on
// tmem-empty 43 % of the time).  So the common path is branch-free: 32 compares build a survivor bitmask in four
// independent chains, and only set bits — a handful per warp and chunk — take the append path, which fetches the
// score with a select tree instead of dynamic register indexing.

// survivors of one chunk: bit j set <=> score j beats the threshold
__device__ __forceinline__ uint32_t epi_hits(const uint32_t (&v)[32], float thr) {
    uint32_t h[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j & 3] |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
    return (h[0] | h[1]) | (h[2] | h[3]);
}
// v[j] for a run-time j without spilling v to local memory: five levels of selects
__device__ __forceinline__ uint32_t epi_pick(const uint32_t (&v)[32], int j) {
    uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
    return (j & 16) ? d[1] : d[0];
}
// append the chunk's survivors (rows r0 + j) to this thread's candidate buffer; the caller made room for 32
__device__ __forceinline__ void epi_append(const uint32_t (&v)[32], uint32_t mw, float thr, int64_t r0,
                                           uint64_t* __restrict__ buf_keys, int& cnt) {
    uint32_t hit = epi_hits(v, thr) & mw;
    while (hit) {
        const int j = __ffs((int)hit) - 1;
        hit &= hit - 1;
        buf_keys[cnt++] = make_key(__uint_as_float(epi_pick(v, j)), (uint32_t)(r0 + j));
    }
}
// phase A: the MAX_TOPS best scores seen so far (descending), a compare-exchange chain per score
__device__ __forceinline__ void epi_sample(const uint32_t (&v)[32], uint32_t mw, float (&t)[MAX_TOPS]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float s = ((mw >> j) & 1u) ? __uint_as_float(v[j]) : -INFINITY;
#pragma unroll
        for (int i = 0; i < MAX_TOPS; ++i) {
            const float lo = fminf(s, t[i]);
            t[i] = fmaxf(s, t[i]);
            s = lo;
        }
    }
}
// euclidean: score = 1 - ||q||^2 - ||x||^2 + 2 q.x (chroma_store.py:132-135 on the l2 space); xn = ||x||^2 of the
// chunk's 32 rows (32-aligned; the array is padded to a multiple of 256 rows)
__device__ __forceinline__ void epi_l2(uint32_t (&v)[32], float l2_bias, const float* __restrict__ xn) {
    const float4* x4 = reinterpret_cast<const float4*>(xn);
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 n4 = x4[j4];
        v[4 * j4 + 0] = __float_as_uint(fmaf(2.f, __uint_as_float(v[4 * j4 + 0]), l2_bias - n4.x));
        v[4 * j4 + 1] = __float_as_uint(fmaf(2.f, __uint_as_float(v[4 * j4 + 1]), l2_bias - n4.y));
        v[4 * j4 + 2] = __float_as_uint(fmaf(2.f, __uint_as_float(v[4 * j4 + 2]), l2_bias - n4.z));
        v[4 * j4 + 3] = __float_as_uint(fmaf(2.f, __uint_as_float(v[4 * j4 + 3]), l2_bias - n4.w));
    }
}
// ---- in-kernel sampling (round 2: replaces the separate sampling launch + threshold kernel)
// The first tile of every CTA stays in its TMEM buffer while the epilogue threads (a) read it once to publish
// each query's best scores, (b) meet grid-wide, (c) turn the published scores into per-query thresholds — one
// warp per query, spread over all CTAs —, (d) meet again, and then read the same accumulator a second time with the
// bound.  The MMA warp keeps going on the other buffer meanwhile.  Needs every CTA resident: cooperative launch.
__device__ __forceinline__ void named_bar_sync(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all n_epi epilogue threads of every CTA; `ctr` counts CTAs, `target` = arrivals that complete this rendezvous
__device__ __forceinline__ void epi_grid_barrier(unsigned int* ctr, unsigned int target, int n_epi, bool first_thread) {
    __threadfence();
    named_bar_sync(1, n_epi);
    if (first_thread) {
        atomicAdd(ctr, 1u);
        const long long t0 = clock64();
        while (ld_acquire_gpu_u32(ctr) < target)
            if (clock64() - t0 > 4000000000ll) __trap();  // a CTA never arrived: fail instead of hanging the GPU
    }
    named_bar_sync(1, n_epi);
}
// k-th largest of the n_cta * m published scores of query q, minus one ulp (rows tying with it still pass the strict
// `s > thr`): every published score belongs to a distinct real row, so at least k rows reach it and it cannot exceed
// the true k-th best.  Fewer than k published → -inf.  cta_stride 1: every CTA published for q; 2 (pair kernel):
// CTA 2i + (q >= 128).  Bisection on the monotone bit patterns (32 rounds of compare + warp popcount); whole warp.
// q_stride: query slots per CTA row of `tops` (MAX_Q, or MAX_Q * chunks for the pair kernel's multi-chunk launches).
__device__ __forceinline__ float warp_threshold(const float* tops, int n_cta, int m, int k, int q, int cta_stride, int lane,
                                                int q_stride = MAX_Q) {
    constexpr int NV = (148 + 16) * MAX_TOPS / 32 + 1;  // n <= (SMs + cluster padding) * MAX_TOPS
    const int n = n_cta * m;
    uint32_t v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i = lane + 32 * j;
        float f = -INFINITY;
        if (i < n) {
            const int cta = (i / m) * cta_stride + (cta_stride == 2 && (q % MAX_Q) >= BLOCK_Q ? 1 : 0);
            f = __ldcg(tops + ((int64_t)cta * MAX_TOPS + (i % m)) * q_stride + q);
        }
        v[j] = score_bits(f);
    }
    uint32_t t = 0;
    if (n >= k) {
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = t | (1u << bit);
            int c = 0;
#pragma unroll
            for (int j = 0; j < NV; ++j) c += (lane + 32 * j < n) && (v[j] >= cand);
            c = __reduce_add_sync(YRB_FULL, c);
            if (c >= k) t = cand;
        }
    }
    return (n >= k) ? nextafterf(bits_score(t), -INFINITY) : -INFINITY;
}
// steps (b)-(d) above.  epi_warp / n_epi_warps number the epilogue warps of the whole grid.  One call handles the
// queries [q_begin, q_end) (a chunk); `round` counts the calls of this launch (the counter only grows).
__device__ __forceinline__ float epi_exchange_thresholds(const float (&tops_l)[MAX_TOPS], float* tops, int m_tops, float* thr_out,
                                                         unsigned int* sync_ctr, int n_cta_pub, int cta_stride, int k, int qi,
                                                         bool active, int epi_warp, int n_epi_warps, int n_epi_threads,
                                                         bool first_thread, int lane, int q_begin, int q_end,
                                                         int q_stride = MAX_Q, int round = 0) {
    if (active) {
#pragma unroll
        for (int i = 0; i < MAX_TOPS; ++i)
            if (i < m_tops) tops[((int64_t)blockIdx.x * MAX_TOPS + i) * q_stride + qi] = tops_l[i];
    }
    epi_grid_barrier(sync_ctr, (2u * round + 1u) * gridDim.x, n_epi_threads, first_thread);
    for (int q = q_begin + epi_warp; q < q_end; q += n_epi_warps) {
        const float t = warp_threshold(tops, n_cta_pub, m_tops, k, q, cta_stride, lane, q_stride);
        if (lane == 0) thr_out[q] = t;
    }
    epi_grid_barrier(sync_ctr, (2u * round + 2u) * gridDim.x, n_epi_threads, first_thread);
    return active ? __ldcg(thr_out + qi) : INFINITY;
}

// make room for up to 32 appends per lane: a lane whose buffer is nearly full has it compacted in place by the
// whole warp (bitonic sort of the 256 slots, the best k stay) and its threshold raised to the k-th kept score
__device__ __forceinline__ void epi_make_room(int& cnt, float& thr, uint64_t* buf_keys, int k, int lane) {
    unsigned need = __ballot_sync(YRB_FULL, cnt > CAP - 32);
    while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        const int n = __shfl_sync(YRB_FULL, cnt, L);
        uint64_t* bp = reinterpret_cast<uint64_t*>(shfl_u64((uint64_t)buf_keys, L));
        uint64_t v[8];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (i * 32 + lane < n) ? bp[i * 32 + lane] : 0ull;
        warp_sort256_desc(v, lane);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i * 32 + lane < k) bp[i * 32 + lane] = v[i];
        uint64_t kth = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint64_t x = shfl_u64(v[i], (k - 1) & 31);
            if (((k - 1) >> 5) == i) kth = x;
        }
        __syncwarp();
        if (lane == L) {
            cnt = n < k ? n : k;
            if (n >= k) thr = key_score(kth);
        }
    }
}

}  // namespace k2
}  // namespace yrb
