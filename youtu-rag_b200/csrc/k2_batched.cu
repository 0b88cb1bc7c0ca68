// K2 — batched search: bf16 tcgen05 GEMM (TMA-fed, TMEM accumulators) with a fused top-k epilogue,
// so the [queries x rows] score matrix never reaches HBM.
//
// Stands in for the loop of single searches the reference runs for a batch
// (VectorRetriever.batch_retrieve, utu/rag/knowledge_retrieval/base_retriever.py:95-99, each a
// collection.query, chroma_store.py:118-120), computed exactly.
//
// Shape of the work.  The GEMM is issued "transposed": MMA-M = 128 queries (A operand),
// MMA-N = 128 corpus rows (B operand), K = 16 per instruction, both operands K-major in shared memory
// with the 128-byte swizzle TMA produces.  TMEM lane = query, TMEM column = corpus row of the tile, so
// an epilogue thread owns ONE query: its running threshold and candidate count live in registers and
// there are no atomics or shared-memory lists in the epilogue.
//   warp 0     TMA producer (corpus tile from HBM, query k-block from L2) into a 4/6-stage ring
//   warp 1     tcgen05.mma issuer (one thread), commits release ring slots and publish accumulators
//   warp 2     TMEM allocator
//   warps 4..  epilogue: tcgen05.ld 32 columns, compare with the query's threshold, append survivors
//              (64-bit keys) to the per-(CTA,query) candidate buffer in global memory (L2 resident)
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// Thresholds: phase A scans one tile per CTA with no threshold and publishes each query's best
// scores; a tiny kernel turns them into a valid lower bound of the k-th best score (the k-th largest
// of a set of real row scores); phase B scans the rest with that bound, so only ~O(k) candidates per
// query survive chip-wide.  A buffer that fills up is compacted in place (warp bitonic sort).  The final
// kernel selects the sorted top-k per query from all candidate buffers.
// Algorithmic flops 2·nq·N·D; algorithmic bytes N·ld·2 (corpus, once) + nq·ld·2.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/yrb200.h"
#include "common.cuh"
#include "k2_batched.h"
#include "k2_common.cuh"
#include "kernels.h"

namespace yrb {

namespace k2 {

// ---------------------------------------------------------------- the kernel
template <int QB>
__device__ __forceinline__ void
    k2_gemm_topk_body(const CUtensorMap& tmap_q, const CUtensorMap& tmap_r, int64_t n_rows,
                 int kblocks, int tile_begin, int iters, int nq, int k, const uint32_t* __restrict__ mask,
                 const float* __restrict__ thr_init, uint64_t* __restrict__ cand_keys, int* __restrict__ cand_cnt,
                 float* __restrict__ tops, int m_tops, const float* __restrict__ q_sqnorm, const float* __restrict__ row_sqnorm,
                 int64_t mask_q_stride, const uint32_t* __restrict__ rowmap, float* __restrict__ thr_out,
                 unsigned int* __restrict__ sync_ctr, float score_floor) {
    constexpr int S = stages(QB);
    constexpr int SB = stage_bytes(QB);
    constexpr int ACC_COLS = QB * BLOCK_R;  // TMEM columns per accumulator buffer
    constexpr int NBUF = acc_buffers(QB);
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2 * S + 4];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem) + 1023u) & ~1023u;
    auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
    auto empty_bar = [&](int s) { return smem_u32(&bars[S + s]); };
    auto tfull_bar = [&](int b) { return smem_u32(&bars[2 * S + b]); };
    auto tempty_bar = [&](int b) { return smem_u32(&bars[2 * S + 2 + b]); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), cluster_nctarank());  // one commit-arrival from every CTA of the cluster
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4 * QB);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(NBUF * ACC_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // Thread-block cluster: the query k-block is the same for every CTA, so each CTA fetches 1/CL of it
    // and multicasts it to its peers (L2→SM bytes per k-block drop from (QB+1)·16 KiB to (QB/CL+1)·16 KiB).
    const uint32_t CL = cluster_nctarank();
    const uint32_t crank = cluster_ctarank();
    const uint16_t cmask = (uint16_t)((1u << CL) - 1u);
    if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anything is multicast at them

    // this CTA's tiles: tile_begin + blockIdx.x + i * gridDim.x for i < iters — the SAME trip count on every
    // CTA (cluster peers advance the ring in lockstep); tiles past the end are zero-filled by TMA and masked.
    const int first = tile_begin + (int)blockIdx.x;
    const int step = (int)gridDim.x;
    const int tile_end = first + iters * step;

    if (warp == 0) {
        if (rowmap) {
            // Row-subset mode (selective shared filter): tile t covers compact rows [128t, 128t+128) whose matrix
            // rows are rowmap[...].  All 32 lanes take part: lane l gathers rows 4l..4l+3 of the tile with one
            // tile::gather4 per k-block (tmap_r is then the gather map, box = one 64-element row).
            int s = 0;
            uint32_t ph = 0;
            for (int t = first; t < tile_end; t += step) {
                uint4 rr = make_uint4(0, 0, 0, 0);  // tiles past the end (equal trip counts) gather row 0 and are masked
                if ((int64_t)t * BLOCK_R < n_rows) rr = reinterpret_cast<const uint4*>(rowmap)[(int64_t)t * (BLOCK_R / 4) + lane];
                for (int kb = 0; kb < kblocks; ++kb) {
                    if (lane == 0) {
                        mbar_wait(empty_bar(s), ph ^ 1);
                        mbar_expect_tx(full_bar(s), SB);
                        const uint32_t dq = smem0 + s * SB;
                        const uint32_t piece_rows = (QB * BLOCK_Q) / CL;
                        if (CL == 1)
                            tma_load_2d(dq, &tmap_q, full_bar(s), kb * BLOCK_K, 0);
                        else
                            tma_load_2d_mcast(dq + crank * piece_rows * (BLOCK_K * 2), &tmap_q, full_bar(s), kb * BLOCK_K,
                                              (int)(crank * piece_rows), cmask);
                    }
                    __syncwarp();
                    tma_gather4(smem0 + s * SB + QB * QTILE_BYTES + lane * (4 * BLOCK_K * 2), &tmap_r, full_bar(s),
                                kb * BLOCK_K, (int)rr.x, (int)rr.y, (int)rr.z, (int)rr.w);
                    if (++s == S) {
                        s = 0;
                        ph ^= 1;
                    }
                }
            }
        } else {
            // converged warp, one elected lane issues (see elect_one)
            int s = 0;
            uint32_t ph = 0;
            const uint32_t piece_rows = (QB * BLOCK_Q) / CL;  // this CTA's slice of the query block (= the tensor map's box)
            for (int t = first; t < tile_end; t += step) {
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(full_bar(s), SB);
                        const uint32_t dst = smem0 + s * SB;
                        if (CL == 1)
                            tma_load_2d(dst, &tmap_q, full_bar(s), kb * BLOCK_K, 0);
                        else
                            tma_load_2d_mcast(dst + crank * piece_rows * (BLOCK_K * 2), &tmap_q, full_bar(s), kb * BLOCK_K,
                                              (int)(crank * piece_rows), cmask);
                        tma_load_2d(dst + QB * QTILE_BYTES, &tmap_r, full_bar(s), kb * BLOCK_K, t * BLOCK_R);
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
        // the whole warp walks the ring (every operand below is warp-uniform); one elected lane issues
        int s = 0;
        uint32_t ph = 0;
        int buf = 0;
        uint32_t bph = 0;
        for (int t = first; t < tile_end; t += step) {
            mbar_wait(tempty_bar(buf), bph ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t a0 = smem0 + s * SB;
                if (elect_one()) {
                    const uint64_t bdesc = smem_desc(a0 + QB * QTILE_BYTES);
                    // the two query blocks accumulate into different TMEM tiles: alternate them so that
                    // consecutive MMAs never depend on each other's accumulator
#pragma unroll
                    for (int k4 = 0; k4 < BLOCK_K / UMMA_K; ++k4) {
#pragma unroll
                        for (int qb = 0; qb < QB; ++qb) {
                            const uint64_t adesc = smem_desc(a0 + qb * QTILE_BYTES);
                            const uint32_t d = tmem_base + buf * ACC_COLS + qb * BLOCK_R;
                            umma_bf16(d, adesc + 2 * k4, bdesc + 2 * k4, IDESC, (kb | k4) != 0);
                        }
                    }
                    if (CL == 1) umma_commit(empty_bar(s));
                    else umma_commit_mcast(empty_bar(s), cmask);
                    if (kb == kblocks - 1) umma_commit(tfull_bar(buf));
                }
                __syncwarp();
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (++buf == NBUF) {
                buf = 0;
                bph ^= 1;
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int qb = ew >> 2, quarter = warp & 3;
        const int qi = qb * BLOCK_Q + quarter * 32 + lane;
        const bool active = qi < nq;
        const int64_t slot = (int64_t)blockIdx.x * MAX_Q + qi;
        uint64_t* buf_keys = cand_keys + slot * CAP;
        int cnt = active ? cand_cnt[slot] : 0;
        // score_floor: one ulp below the caller's score threshold (keep hits with score >= threshold), -inf = none
        float thr = fmaxf((active && thr_init) ? thr_init[qi] : -INFINITY, score_floor);
        if (!active) thr = INFINITY;
        float tops_l[MAX_TOPS];
#pragma unroll
        for (int i = 0; i < MAX_TOPS; ++i) tops_l[i] = -INFINITY;
        // tops && thr_out: sampling fused into this launch (first tile read twice, see epi_exchange_thresholds);
        // tops only: stand-alone sampling pass (publishes, appends nothing) — kept for non-cooperative launches
        const bool fuse = tops != nullptr && thr_out != nullptr;
        const bool sample_only = tops != nullptr && !fuse;
        // euclidean: score = 1 - ||q||^2 - ||x||^2 + 2 q.x (chroma_store.py:132-135 on the l2 space)
        const bool l2 = q_sqnorm != nullptr;
        const float l2_bias = (l2 && active) ? 1.f - q_sqnorm[qi] : 0.f;
        // filter bitmask: shared by all queries (stride 0) or one per query (text2sql-style batches)
        const uint32_t* qmask = mask ? mask + (active ? (int64_t)qi * mask_q_stride : 0) : nullptr;
        int buf = 0;
        uint32_t bph = 0;
        for (int t = first; t < tile_end; t += step) {
            mbar_wait(tfull_bar(buf), bph);
            tc_fence_after();
            const int64_t row0 = (int64_t)t * BLOCK_R;
            const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * ACC_COLS + qb * BLOCK_R;
            for (int pass = (fuse && t == first) ? 0 : 1; pass < 2; ++pass) {
                const bool sampling = sample_only || pass == 0;
#pragma unroll 1
                for (int c = 0; c < BLOCK_R / 32; ++c) {
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
                if (pass == 0)
                    thr = fmaxf(score_floor, epi_exchange_thresholds(tops_l, tops, m_tops, thr_out, sync_ctr, (int)gridDim.x, 1, k, qi,
                                                                     active, (int)blockIdx.x * 4 * QB + ew, (int)gridDim.x * 4 * QB,
                                                                     128 * QB, threadIdx.x == 128, lane, 0, nq));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
            if (++buf == NBUF) {
                buf = 0;
                bph ^= 1;
            }
        }
        if (active) {
            if (!sample_only) cand_cnt[slot] = cnt;
            if (sample_only) {
#pragma unroll
                for (int i = 0; i < MAX_TOPS; ++i)
                    if (i < m_tops) tops[((int64_t)blockIdx.x * MAX_TOPS + i) * MAX_Q + qi] = tops_l[i];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NBUF * ACC_COLS));
    }
}

// Two entry points over the same body.  The cluster size is a COMPILE-TIME attribute of the kernel (as in k2_pair.cu):
// with the cluster dimension passed as a launch attribute next to the cooperative attribute, the launch failed under
// Nsight Compute (the profiler recorded a (0,0,0) grid and the driver reported LaunchFailed), which would break any
// tooling that lists the kernels of a run.
#define YRB_K2_PARAMS                                                                                                       \
    const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_r, int64_t n_rows, int kblocks,   \
        int tile_begin, int iters, int nq, int k, const uint32_t *__restrict__ mask, const float *__restrict__ thr_init,    \
        uint64_t *__restrict__ cand_keys, int *__restrict__ cand_cnt, float *__restrict__ tops, int m_tops,                 \
        const float *__restrict__ q_sqnorm, const float *__restrict__ row_sqnorm, int64_t mask_q_stride,                    \
        const uint32_t *__restrict__ rowmap, float *__restrict__ thr_out, unsigned int *__restrict__ sync_ctr, float score_floor
#define YRB_K2_ARGS                                                                                                          \
    tmap_q, tmap_r, n_rows, kblocks, tile_begin, iters, nq, k, mask, thr_init, cand_keys, cand_cnt, tops, m_tops, q_sqnorm,  \
        row_sqnorm, mask_q_stride, rowmap, thr_out, sync_ctr, score_floor
template <int QB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 128 * QB, 1) k2_gemm_topk(YRB_K2_PARAMS) {
    k2_gemm_topk_body<QB>(YRB_K2_ARGS);
}
template <int QB>
__global__ void __launch_bounds__(128 + 128 * QB, 1) k2_gemm_topk_c1(YRB_K2_PARAMS) {
    k2_gemm_topk_body<QB>(YRB_K2_ARGS);
}
#undef YRB_K2_PARAMS
#undef YRB_K2_ARGS

// ---------------------------------------------------------------- thresholds from phase A
// thr0[q] = k-th largest of the (n_cta * m) published best scores: every one is the score of a distinct
// real row, so at least k rows score >= thr0[q] and thr0[q] <= the true k-th best.  Fewer than k → -inf.
// cta_stride 1: every CTA published for every query; 2 (pair kernel): CTA 2i + (q >= 128) published for q.
// One warp per query: the k-th largest of the n = n_cta*m published scores by bisection on their monotone bit
// patterns (32 rounds of compare + warp popcount) — no shared memory, no block barriers.
__global__ void __launch_bounds__(256) k2_threshold_kernel(const float* __restrict__ tops, int n_cta, int m, int k,
                                                           float* __restrict__ thr0, int cta_stride, int nq) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float t = warp_threshold(tops, n_cta, m, k, q, cta_stride, lane);
    if (lane == 0) thr0[q] = t;
}

}  // namespace k2

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct K2State {
    EncodeTiledFn encode = nullptr;
    int slots = 0;
    int qcap = 0;   // query slots per CTA row of the candidate buffers: MAX_Q, or MAX_Q * MAX_QC once a batch of more than 256 queries was seen
    uint64_t* cand_keys = nullptr;
    int* cand_cnt = nullptr;
    float* tops = nullptr;
    float* thr0 = nullptr;
    bool attrs_set = false;
};

K2State* k2_create() { return new K2State(); }
void k2_destroy(K2State* s) {
    if (!s) return;
    if (s->cand_keys) cudaFree(s->cand_keys);
    if (s->cand_cnt) cudaFree(s->cand_cnt);
    if (s->tops) cudaFree(s->tops);
    if (s->thr0) cudaFree(s->thr0);
    delete s;
}
int k2_parts(int sm_count) { return sm_count; }
bool k2_supported(int dtype, int dim, int k) { return dtype == 0 && dim >= 1 && k >= 1 && k <= YRB_FUSED_K_MAX; }

static bool make_map(K2State* s, CUtensorMap* m, const void* base, uint64_t rows, int ld, int box_rows, std::string& err) {
    cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)k2::BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = s->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return false;
    }
    return true;
}

#define K2CK(call)                                                                     \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                  \
            return YRB_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)

template <int QB>
static cudaError_t launch_gemm(int grid, int cluster, const CUtensorMap& mq, const CUtensorMap& mr, int64_t n_rows,
                               int kblocks, int tile_begin, int iters, int nq, int k, const uint32_t* mask,
                               const float* thr, uint64_t* ck, int* cc, float* tops, int m_tops, const float* q_sqnorm,
                               const float* row_sqnorm, int64_t mask_q_stride, const uint32_t* rowmap, float* thr_out,
                               unsigned int* sync_ctr, float score_floor, cudaStream_t st) {
    const size_t smem = (size_t)k2::stages(QB) * k2::stage_bytes(QB) + 1024;
    auto kern = cluster == 2 ? k2::k2_gemm_topk<QB> : k2::k2_gemm_topk_c1<QB>;   // cluster of 2 (compile-time) or none
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128 + 128 * QB);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;   // the fused sampling meets grid-wide: every CTA must be resident
    at[0].val.cooperative = thr_out != nullptr ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, mq, mr, n_rows, kblocks, tile_begin, iters, nq, k, mask, thr, ck, cc, tops, m_tops,
                              q_sqnorm, row_sqnorm, mask_q_stride, rowmap, thr_out, sync_ctr, score_floor);
}

// CTA-pair kernel for 129..256-query chunks — the default since round 2 (C3: 0.383 ms against 0.449 ms for the
// one-CTA kernel, whose 128x128 SS-mode MMAs are bound by shared-memory operand reads); YRB_K2_PAIR=0 keeps the
// one-CTA kernel for A/B runs
static bool k2_use_pair(bool forced) {
    static int v = -1;
    if (v < 0) v = getenv("YRB_K2_PAIR") ? atoi(getenv("YRB_K2_PAIR")) : 1;
    return forced || v != 0;
}

// cluster size for the query multicast: 2 (always packs the 148 SMs: 74 TPCs) or, with YRB_K2_CLUSTER=1, none.
// (4 was measured in round 1 — 0.87 ms against 0.50: it strands SMs — and is no longer selectable: the cluster size
// is a compile-time attribute of the kernel now.)
static int k2_cluster(int grid) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("YRB_K2_CLUSTER");
        forced = e ? atoi(e) : 0;
    }
    return (forced == 1 || grid % 2) ? 1 : 2;
}

int k2_search(K2State* s, const void* rows, int64_t n_rows, int64_t capacity, int dim, int ld, const void* q, int nq,
              int k, const uint32_t* mask_all, int64_t mask_q_stride, int metric, const float* q_sqnorm,
              const float* row_sqnorm, uint64_t* out_keys, int64_t* out_ids, float* out_scores, int32_t* out_counts, int sm_count, cudaStream_t st,
              int* launches, std::string& err,
              cudaEvent_t ev_start, cudaEvent_t ev_stop, bool force_pair, const uint32_t* rowmap, int64_t matrix_rows,
              const XShard* xs_in, float min_score) {
    (void)capacity; (void)dim;
    if (rowmap) force_pair = false;  // the row-subset producer lives in the one-CTA kernel
    const float* xn = metric == YRB_METRIC_L2 ? row_sqnorm : nullptr;
    // the epilogue keeps scores strictly above its bound: one ulp below the threshold keeps score == threshold
    const float score_floor = (min_score > -INFINITY) ? nextafterf(min_score, -INFINITY) : -INFINITY;
    if (!s->encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        K2CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            err = "cuTensorMapEncodeTiled not available from the driver";
            return YRB_ERR_UNSUPPORTED;
        }
        s->encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const int want_slots = sm_count + 16;  // grids are rounded up to the cluster size
    // batches of more than 256 queries run the pair kernel over up to MAX_QC chunks per launch (one pass over the corpus):
    // the candidate buffers then hold MAX_Q * MAX_QC query slots per CTA (344 MB per device instead of 86 MB)
    static int fuse_env = -1;
    if (fuse_env < 0) fuse_env = getenv("YRB_K2_FUSE") ? atoi(getenv("YRB_K2_FUSE")) : 1;
    static int multi_env = -1;
    if (multi_env < 0) multi_env = getenv("YRB_K2_MULTICHUNK") ? atoi(getenv("YRB_K2_MULTICHUNK")) : 1;
    static bool fuse_ok[2] = {true, true};  // [0] one-CTA kernel, [1] pair kernel: cleared if a cooperative launch is refused
    const bool multichunk = nq > k2::MAX_Q && !rowmap && multi_env != 0 && fuse_env != 0 && fuse_ok[1] && k2_use_pair(force_pair);
    const int want_qcap = multichunk ? k2::MAX_Q * k2::MAX_QC : k2::MAX_Q;
    if (s->slots < want_slots || s->qcap < want_qcap) {
        const int slots = std::max(want_slots, s->slots), qcap = std::max(want_qcap, s->qcap);
        K2CK(cudaStreamSynchronize(st));
        if (s->cand_keys) cudaFree(s->cand_keys);
        if (s->cand_cnt) cudaFree(s->cand_cnt);
        if (s->tops) cudaFree(s->tops);
        if (s->thr0) cudaFree(s->thr0);
        s->cand_keys = nullptr; s->cand_cnt = nullptr; s->tops = nullptr; s->thr0 = nullptr;
        s->slots = s->qcap = 0;
        K2CK(cudaMalloc(&s->cand_keys, (size_t)slots * qcap * k2::CAP * 8));
        K2CK(cudaMalloc(&s->cand_cnt, ((size_t)slots * qcap + 1) * 4));  // + the grid-barrier counter
        K2CK(cudaMalloc(&s->tops, (size_t)slots * k2::MAX_TOPS * qcap * 4));
        K2CK(cudaMalloc(&s->thr0, (size_t)qcap * 4));
        s->slots = slots;
        s->qcap = qcap;
    }
    const int kblocks = ld / k2::BLOCK_K;
    const int tiles = (int)((n_rows + k2::BLOCK_R - 1) / k2::BLOCK_R);
    CUtensorMap mr;
    if (rowmap) {
        // gather map over the WHOLE matrix: box = one row x 64 elements, rows chosen per instruction
        static int gbox = -1;
        if (gbox < 0) gbox = getenv("YRB_K2_GATHER_BOX") ? atoi(getenv("YRB_K2_GATHER_BOX")) : 1;
        if (!make_map(s, &mr, rows, (uint64_t)matrix_rows, ld, gbox, err)) return YRB_ERR_CUDA;
    } else if (!make_map(s, &mr, rows, (uint64_t)n_rows, ld, k2::BLOCK_R, err)) return YRB_ERR_CUDA;

    // sampling inside the main launch (cooperative) unless YRB_K2_FUSE=0 (the round-1 sequence: sampling launch,
    // threshold kernel, main launch — kept for A/B runs)
    unsigned int* sync_ctr = reinterpret_cast<unsigned int*>(s->cand_cnt + (size_t)s->slots * s->qcap);  // zeroed with the counts
    int step_q = k2::MAX_Q;
    for (int c0 = 0; c0 < nq; c0 += step_q) {
        const bool use_pair_now = (nq - c0 > k2::BLOCK_Q) && !rowmap && k2_use_pair(force_pair);
        const bool fuse = fuse_env != 0 && fuse_ok[use_pair_now ? 1 : 0];
        // queries of this launch: up to MAX_QC chunks on the pair kernel with fused sampling, one chunk otherwise
        const int chunks = (use_pair_now && fuse && multichunk) ? std::min(k2::MAX_QC, (nq - c0 + k2::MAX_Q - 1) / k2::MAX_Q) : 1;
        step_q = chunks * k2::MAX_Q;
        const int nqc = nq - c0 < step_q ? nq - c0 : step_q;
        const int QB = nqc > k2::BLOCK_Q ? 2 : 1;
        const float* qn = metric == YRB_METRIC_L2 ? q_sqnorm + c0 : nullptr;
        const uint32_t* mask = mask_all ? mask_all + (size_t)c0 * mask_q_stride : nullptr;
        uint64_t* o_keys = out_keys + (size_t)c0 * k;
        int64_t* o_ids = out_ids ? out_ids + (size_t)c0 * k : nullptr;
        float* o_scores = out_scores ? out_scores + (size_t)c0 * k : nullptr;
        int32_t* o_counts = out_counts ? out_counts + c0 : nullptr;
        XShard xs_c{};
        const XShard* xs = nullptr;
        if (xs_in) {
            xs_c = *xs_in;
            xs_c.q0 += c0;
            xs = &xs_c;
        }
        K2CK(cudaMemsetAsync(s->cand_cnt, 0, ((size_t)s->slots * s->qcap + 1) * 4, st));
        if (nqc > k2::BLOCK_Q && !rowmap && k2_use_pair(force_pair)) {
            // CTA pairs (cta_group::2): 256-row tiles, CTA r of a pair owns queries [128r, 128r+128)
            const int tiles2 = (int)((n_rows + 255) / 256);
            int n_pairs = sm_count / 2;
            if (tiles2 < n_pairs) n_pairs = tiles2;
            const int grid2 = 2 * n_pairs;
            const int iters2 = (tiles2 + n_pairs - 1) / n_pairs;
            const int q_stride2 = chunks > 1 ? s->qcap : k2::MAX_Q;   // single chunks keep the 256-slot rows (round-1 layout)
            CUtensorMap mq2;
            if (!make_map(s, &mq2, reinterpret_cast<const char*>(q) + (size_t)c0 * ld * 2, (uint64_t)((nqc + 127) / 128 * 128), ld,
                          k2::BLOCK_Q, err))
                return YRB_ERR_CUDA;
            // thresholds: k-th largest of the best scores every pair publishes for its first tile — a valid lower bound
            // of the k-th best overall as soon as n_pairs * m_tops >= k scores are published
            const int m_tops2 = std::min(k2::MAX_TOPS, std::max(1, (3 * k + n_pairs - 1) / n_pairs));
            const bool sampled2 = (int64_t)n_pairs * m_tops2 >= k && (fuse || tiles2 > 2 * n_pairs);
            const float* thr2 = nullptr;
            if (sampled2 && !fuse) {
                K2CK(launch_gemm_pair(grid2, mq2, mr, n_rows, kblocks, 1, nqc, k, mask, mask_q_stride, nullptr, s->cand_keys,
                                      s->cand_cnt, s->tops, m_tops2, qn, xn, nullptr, nullptr, score_floor, 1, k2::MAX_Q, st));
                k2::k2_threshold_kernel<<<(nqc + 7) / 8, 256, 0, st>>>(s->tops, n_pairs, m_tops2, k, s->thr0, 2, nqc);
                K2CK(cudaGetLastError());
                *launches += 2;
                thr2 = s->thr0;
            }
            const bool fz = sampled2 && fuse;
            if (ev_start && c0 == 0) K2CK(cudaEventRecord(ev_start, st));
            {
                cudaError_t e = launch_gemm_pair(grid2, mq2, mr, n_rows, kblocks, iters2, nqc, k, mask, mask_q_stride, thr2, s->cand_keys,
                                                 s->cand_cnt, fz ? s->tops : nullptr, fz ? m_tops2 : 0, qn, xn, fz ? s->thr0 : nullptr,
                                                 fz ? sync_ctr : nullptr, score_floor, chunks, q_stride2, st);
                if (e != cudaSuccess && fz) {  // cooperative launch refused: redo this chunk with the separate sampling pass
                    fprintf(stderr, "yrb200: cooperative launch of k2_gemm_topk_pair refused (%s); sampling runs as separate launches\n",
                            cudaGetErrorString(e));
                    cudaGetLastError();
                    fuse_ok[1] = false;
                    step_q = 0;   // redo from the same c0 (as single chunks with the separate sampling pass)
                    continue;
                }
                K2CK(e);
            }
            if (ev_start && c0 == 0) K2CK(cudaEventRecord(ev_stop, st));
            K2CK(launch_select_segments(s->cand_keys, (int64_t)q_stride2 * k2::CAP, k2::CAP, s->cand_cnt, q_stride2, 1, grid2, 0,
                                        k2::CAP, nullptr, nqc, k, o_keys, st, o_ids, o_scores, o_counts, xs));
            *launches += 2;
            continue;
        }
        // grids are multiples of the cluster size; every CTA runs the same number of tiles
        int grid = tiles < sm_count ? tiles : sm_count;
        const int cluster = k2_cluster(sm_count);
        grid = (grid + cluster - 1) / cluster * cluster;
        const int iters = (tiles + grid - 1) / grid;
        CUtensorMap mq;
        if (!make_map(s, &mq, reinterpret_cast<const char*>(q) + (size_t)c0 * ld * 2, (uint64_t)((nqc + 127) / 128 * 128), ld,
                      QB * k2::BLOCK_Q / cluster, err))
            return YRB_ERR_CUDA;
        auto gemm = [&](int it, const float* thr, float* tops, int m, float* thr_out, unsigned int* ctr) -> cudaError_t {
            return QB == 2 ? launch_gemm<2>(grid, cluster, mq, mr, n_rows, kblocks, 0, it, nqc, k, mask, thr, s->cand_keys, s->cand_cnt,
                                            tops, m, qn, xn, mask_q_stride, rowmap, thr_out, ctr, score_floor, st)
                           : launch_gemm<1>(grid, cluster, mq, mr, n_rows, kblocks, 0, it, nqc, k, mask, thr, s->cand_keys, s->cand_cnt,
                                            tops, m, qn, xn, mask_q_stride, rowmap, thr_out, ctr, score_floor, st);
        };
        // sampling: each CTA's first tile publishes its best scores per query (CTAs past the last tile publish -inf)
        const int gridA = tiles < grid ? tiles : grid;   // CTAs that see a real tile
        const int m_tops = std::min(k2::MAX_TOPS, std::max(1, (3 * k + gridA - 1) / gridA));
        const bool sampled = (int64_t)gridA * m_tops >= k && (fuse || tiles > 2 * grid);
        const float* thr = nullptr;
        if (sampled && !fuse) {
            K2CK(gemm(1, nullptr, s->tops, m_tops, nullptr, nullptr));
            k2::k2_threshold_kernel<<<(nqc + 7) / 8, 256, 0, st>>>(s->tops, gridA, m_tops, k, s->thr0, 1, nqc);
            K2CK(cudaGetLastError());
            *launches += 2;
            thr = s->thr0;
        }
        const bool fz = sampled && fuse;
        if (ev_start && c0 == 0) K2CK(cudaEventRecord(ev_start, st));
        {
            cudaError_t e = gemm(iters, thr, fz ? s->tops : nullptr, fz ? m_tops : 0, fz ? s->thr0 : nullptr, fz ? sync_ctr : nullptr);
            if (e != cudaSuccess && fz) {  // cooperative launch refused: redo this chunk with the separate sampling pass
                fprintf(stderr, "yrb200: cooperative launch of k2_gemm_topk refused (%s); sampling runs as separate launches\n",
                        cudaGetErrorString(e));
                cudaGetLastError();
                fuse_ok[0] = false;
                step_q = 0;
                continue;
            }
            K2CK(e);
        }
        if (ev_start && c0 == 0) K2CK(cudaEventRecord(ev_stop, st));
        K2CK(launch_select_segments(s->cand_keys, (int64_t)k2::MAX_Q * k2::CAP, k2::CAP, s->cand_cnt, k2::MAX_Q, 1, grid, 0,
                                    k2::CAP, nullptr, nqc, k, o_keys, st, o_ids, o_scores, o_counts, xs));
        *launches += 2;
    }
    return YRB_OK;
}

}  // namespace yrb
