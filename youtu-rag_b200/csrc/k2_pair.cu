// K2 (pair form) — the batched GEMM + fused top-k epilogue on CTA PAIRS: tcgen05.mma.cta_group::2,
// M = 256 queries (128 per CTA) x N = 256 corpus rows (128 loaded by each CTA) x K = 16.
//
// Why pairs: the one-CTA kernel (k2_batched.cu) is bound by bytes entering each SM — every k-block it
// ingests its 128 corpus rows (16 KiB, from HBM) AND the whole 256-query k-block (32 KiB, from L2), 48 KiB
// per 2.1 M MACs, at a measured ~52 B/clk/SM.  In a pair each SM ingests only its half of the queries and
// its half of the rows (32 KiB per 2.1 M MACs); the tensor cores read the other halves from the peer SM's
// shared memory.  Used for 129..256-query chunks; smaller chunks stay on the one-CTA kernel.
//
// Protocol (CTA 0 of the pair = leader):
//   * both CTAs' producers TMA their halves with the .cta_group::2 form, signalling the LEADER's full barrier
//     (64 KiB of transactions per stage); each waits on its OWN empty barrier,
//   * the leader's MMA thread issues the 2-SM MMAs; commits are multicast to both CTAs' empty / tmem-full
//     barriers,
//   * each CTA's 4 epilogue warps drain its 128 TMEM lanes (= its 128 queries) x 256 columns (= all 256 rows
//     of the pair's tile) and arrive on the LEADER's tmem-empty barrier (count 8).
// Epilogue, thresholds, candidate buffers and selection are those of k2_batched.cu.
//
// More than 256 queries (C5: 1024): the launch sweeps up to MAX_QC = 4 query chunks PER ROW TILE — tile t is multiplied
// with chunk 0, then chunk 1, … before the pair moves to tile t + n_pairs.  The tile's rows come from HBM once (chunk 0)
// and from L2 for the other chunks (a pair's tile is 0.4-0.5 MiB; all pairs together ~35 MiB of the 126 MiB L2), so a
// 1024-query batch reads the corpus ONCE instead of four times.  Each (tile, chunk) is one
// accumulator buffer in the same double-buffered pipeline; a thread's per-chunk state (candidate count, threshold)
// sits in shared memory between its visits.
#include <cstdlib>

#include "../../include/yrb200.h"
#include "k2_batched.h"
#include "k2_common.cuh"
#include "kernels.h"

namespace yrb {
namespace k2 {

constexpr int PAIR_N = 256;              // rows per pair tile (MMA N)
constexpr int PAIR_STAGES = 6;           // 32 KiB per CTA per stage
constexpr int PAIR_STAGE_BYTES = 2 * QTILE_BYTES;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-parity bit of a shared::cluster address → leader
// kind::f16, D = F32, A = B = BF16, K-major, M = 256 (pair), N = 256
constexpr uint32_t IDESC_PAIR = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(PAIR_N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
    k2_gemm_topk_pair(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_r, int64_t n_rows,
                      int kblocks, int iters, int nq, int k, const uint32_t* __restrict__ mask, int64_t mask_q_stride,
                      const float* __restrict__ thr_init, uint64_t* __restrict__ cand_keys, int* __restrict__ cand_cnt,
                      float* __restrict__ tops, int m_tops, const float* __restrict__ q_sqnorm,
                      const float* __restrict__ row_sqnorm, float* __restrict__ thr_out, unsigned int* __restrict__ sync_ctr,
                      float score_floor, int nqc /* query chunks of 256, 1..MAX_QC */, int q_stride /* query slots per CTA row */) {
    constexpr int S = PAIR_STAGES;
    constexpr int ACC_COLS = PAIR_N;  // per buffer; two buffers = all 512 columns
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2 * S + 4];
    __shared__ uint32_t tmem_base_s;
    __shared__ int s_cnt[MAX_QC][BLOCK_Q];     // per (chunk, epilogue thread): candidates appended so far
    __shared__ float s_thr[MAX_QC][BLOCK_Q];   // … and the running bound

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    const uint32_t smem0 = (smem_u32(smem) + 1023u) & ~1023u;
    auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
    auto empty_bar = [&](int s) { return smem_u32(&bars[S + s]); };
    auto tfull_bar = [&](int b) { return smem_u32(&bars[2 * S + b]); };
    auto tempty_bar = [&](int b) { return smem_u32(&bars[2 * S + 2 + b]); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);   // leader: its producer's arrive.expect_tx; peer's copy is unused
            mbar_init(empty_bar(s), 1);  // one multicast commit per round
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 8);  // 4 epilogue warps of each CTA arrive at the LEADER's barrier
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // the same warp id in both CTAs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int n_pairs = (int)gridDim.x >> 1;
    const int pair = (int)blockIdx.x >> 1;

    if (warp == 0) {
        // converged warps, one elected lane issues (see elect_one in k2_common.cuh)
        int s = 0;
        uint32_t ph = 0;
        const uint32_t lead_full0 = full_bar(0) & PEER_MASK;
        for (int it = 0; it < iters; ++it) {
            const int t = pair + it * n_pairs;
            for (int qc = 0; qc < nqc; ++qc) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(full_bar(s), 2 * PAIR_STAGE_BYTES);  // both CTAs' halves
                        const uint32_t dst = smem0 + s * PAIR_STAGE_BYTES;
                        const uint32_t lbar = lead_full0 + (uint32_t)s * 8u;
                        tma_load_2d_pair(dst, &tmap_q, lbar, kb * BLOCK_K, qc * MAX_Q + (int)crank * BLOCK_Q);
                        tma_load_2d_pair(dst + QTILE_BYTES, &tmap_r, lbar, kb * BLOCK_K, t * PAIR_N + (int)crank * 128);
                    }
                    __syncwarp();
                    if (++s == S) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            int s = 0;
            uint32_t ph = 0;
            int buf = 0;
            uint32_t bph = 0;
            for (int itq = 0; itq < iters * nqc; ++itq) {   // one accumulator buffer per (tile, query chunk)
                mbar_wait(tempty_bar(buf), bph ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t a0 = smem0 + s * PAIR_STAGE_BYTES;
                    if (elect_one()) {
                        const uint64_t adesc = smem_desc(a0);
                        const uint64_t bdesc = smem_desc(a0 + QTILE_BYTES);
                        const uint32_t d = tmem_base + buf * ACC_COLS;
#pragma unroll
                        for (int k4 = 0; k4 < BLOCK_K / UMMA_K; ++k4)
                            umma_bf16_pair(d, adesc + 2 * k4, bdesc + 2 * k4, IDESC_PAIR, (kb | k4) != 0);
                        umma_commit_pair(empty_bar(s));
                        if (kb == kblocks - 1) umma_commit_pair(tfull_bar(buf));
                    }
                    __syncwarp();
                    if (++s == S) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                buf ^= 1;
                if (buf == 0) bph ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        const int et = quarter * 32 + lane;                                // epilogue thread 0..127 = query within the half
        const int qh = (int)crank * BLOCK_Q + et;                          // … within a 256-query chunk
        // tops && thr_out: sampling fused into this launch (the first tile is read twice, epi_exchange_thresholds)
        const bool fuse = tops != nullptr && thr_out != nullptr;
        const bool sample_only = tops != nullptr && !fuse;                 // stand-alone sampling pass (single chunk only)
        const bool l2 = q_sqnorm != nullptr;
        for (int qc = 0; qc < nqc; ++qc) {
            const int qi = qc * MAX_Q + qh;
            const bool active = qi < nq;
            s_cnt[qc][et] = active ? cand_cnt[(int64_t)blockIdx.x * q_stride + qi] : 0;
            // score_floor: one ulp below the caller's score threshold, -inf = none (see k2_batched.cu)
            s_thr[qc][et] = active ? fmaxf(thr_init ? thr_init[qi] : -INFINITY, score_floor) : INFINITY;
        }
        float tops_l[MAX_TOPS];
        int buf = 0;
        uint32_t bph = 0;
        for (int it = 0; it < iters; ++it) {
            const int t = pair + it * n_pairs;
            const int64_t row0 = (int64_t)t * PAIR_N;
            for (int qc = 0; qc < nqc; ++qc) {
                const int qi = qc * MAX_Q + qh;                            // this thread's query in this chunk
                const bool active = qi < nq;
                uint64_t* buf_keys = cand_keys + ((int64_t)blockIdx.x * q_stride + qi) * CAP;
                int cnt = s_cnt[qc][et];
                float thr = s_thr[qc][et];
                const float l2_bias = (l2 && active) ? 1.f - q_sqnorm[qi] : 0.f;
                const uint32_t* qmask = mask ? mask + (active ? (int64_t)qi * mask_q_stride : 0) : nullptr;
                mbar_wait(tfull_bar(buf), bph);
                tc_fence_after();
                const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * ACC_COLS;
                for (int pass = (fuse && it == 0) ? 0 : 1; pass < 2; ++pass) {
                    const bool sampling = sample_only || pass == 0;
                    if (sampling) {
#pragma unroll
                        for (int i = 0; i < MAX_TOPS; ++i) tops_l[i] = -INFINITY;
                    }
#pragma unroll 1
                    for (int c = 0; c < PAIR_N / 32; ++c) {
                        if (!sampling) epi_make_room(cnt, thr, buf_keys, k, lane);
                        uint32_t v[32];
                        tmem_ld32(tacc + c * 32, v);
                        const int64_t r0 = row0 + c * 32;
                        uint32_t mw = 0u;
                        if (r0 < n_rows) {
                            mw = qmask ? qmask[r0 >> 5] : 0xffffffffu;
                            if (r0 + 32 > n_rows) mw &= (1u << (int)(n_rows - r0)) - 1u;
                            if (l2) epi_l2(v, l2_bias, row_sqnorm + r0);
                        }
                        if (sampling) epi_sample(v, mw, tops_l);
                        else epi_append(v, mw, thr, r0, buf_keys, cnt);
                    }
                    if (pass == 0) {  // CTA 2i + r published for the queries [128r, 128r + 128) of the chunk: n_pairs publishers per query
                        const int q_end = nq < (qc + 1) * MAX_Q ? nq : (qc + 1) * MAX_Q;
                        thr = fmaxf(score_floor, epi_exchange_thresholds(tops_l, tops, m_tops, thr_out, sync_ctr, n_pairs, 2, k, qi, active,
                                                                         (int)blockIdx.x * 4 + quarter, (int)gridDim.x * 4, 128,
                                                                         threadIdx.x == 128, lane, qc * MAX_Q, q_end, q_stride, qc));
                    }
                }
                s_cnt[qc][et] = cnt;
                s_thr[qc][et] = thr;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive(tempty_bar(buf));
                    else mbar_arrive_remote(tempty_bar(buf), 0);
                }
                buf ^= 1;
                if (buf == 0) bph ^= 1;
            }
        }
        for (int qc = 0; qc < nqc; ++qc) {
            const int qi = qc * MAX_Q + qh;
            if (qi >= nq) continue;
            if (!sample_only) cand_cnt[(int64_t)blockIdx.x * q_stride + qi] = s_cnt[qc][et];
        }
        if (sample_only && qh < nq) {   // nqc == 1 here: the tops of the single chunk's single sampled tile
#pragma unroll
            for (int i = 0; i < MAX_TOPS; ++i)
                if (i < m_tops) tops[((int64_t)blockIdx.x * MAX_TOPS + i) * q_stride + qh] = tops_l[i];
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace k2

cudaError_t launch_gemm_pair(int grid, const CUtensorMap& mq, const CUtensorMap& mr, int64_t n_rows, int kblocks, int iters,
                             int nq, int k, const uint32_t* mask, int64_t mask_q_stride, const float* thr, uint64_t* ck,
                             int* cc, float* tops, int m_tops, const float* q_sqnorm, const float* row_sqnorm,
                             float* thr_out, unsigned int* sync_ctr, float score_floor, int nqc, int q_stride, cudaStream_t st) {
    const size_t smem = (size_t)k2::PAIR_STAGES * k2::PAIR_STAGE_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(k2::k2_gemm_topk_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;   // the fused sampling meets grid-wide: every CTA must be resident
    at[0].val.cooperative = thr_out != nullptr ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k2::k2_gemm_topk_pair, mq, mr, n_rows, kblocks, iters, nq, k, mask, mask_q_stride, thr, ck, cc,
                              tops, m_tops, q_sqnorm, row_sqnorm, thr_out, sync_ctr, score_floor, nqc, q_stride);
}

}  // namespace yrb
