// K5 — ingest: fp32 rows → (L2-normalise for cosine) → storage dtype, one warp per row.
// Replaces what chromadb/hnswlib do inside collection.add (chroma_store.py:86) and
// faiss.normalize_L2 + index.add (faiss_store.py:107-110).  Also used to prepare queries
// (faiss_store.py:146-149 normalises the query the same way).
//
// Reproducibility pin: sum of squares and the divide are fp64, the quotient is
// rounded fp64→fp32→bf16(RNE), so the oracle (oracle/exact_search.py:prepare) lands on the same
// stored bits.  HBM-bound: 4·dim B read + 2·ld B written per row.
#include <cuda_bf16.h>

#include <cstdint>

#include "common.cuh"
#include "kernels.h"

namespace yrb {

// Generic form (any dim): one warp per row, two passes over the row (the second one hits L1/L2).
template <bool F32>
__global__ void __launch_bounds__(256) ingest_kernel(const float* __restrict__ src, int64_t n, int dim,
                                                     int ld, int normalize, void* __restrict__ dst,
                                                     float* __restrict__ sqnorm) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
        const float* x = src + row * (int64_t)dim;
        double ss = 0.0;
        for (int i = lane; i < dim; i += 32) {
            double v = (double)x[i];
            ss += v * v;
        }
        ss = warp_sum(ss);
        const bool scale = normalize && ss > 0.0;
        const double inv = scale ? 1.0 / sqrt(ss) : 1.0;   // x * (1/||x||): see the note on the streaming kernel
        float sq = 0.f;
        for (int i = lane; i < ld; i += 32) {
            float y = 0.f;
            if (i < dim) y = scale ? (float)((double)x[i] * inv) : x[i];
            if (F32) {
                reinterpret_cast<float*>(dst)[row * (int64_t)ld + i] = y;
                sq = fmaf(y, y, sq);
            } else {
                __nv_bfloat16 b = __float2bfloat16_rn(y);
                reinterpret_cast<__nv_bfloat16*>(dst)[row * (int64_t)ld + i] = b;
                float yr = __bfloat162float(b);
                sq = fmaf(yr, yr, sq);
            }
        }
        sq = warp_sum(sq);
        if (lane == 0 && sqnorm) sqnorm[row] = sq;
    }
}

// Streaming form (dim % 4 == 0, ld <= 128 * NV): one warp per row, the row is read ONCE with 128-bit loads — all NV
// of a lane in flight together — and held in registers; the fp64 sum of squares, the fp64 divides and the rounding
// are the same operations as above (so the oracle's stored bits are reproduced), only the order in which a lane adds
// its squares differs (float4 j = lane + 32 i, components x..w).  Stores are 64-bit (bf16) / 128-bit (fp32), coalesced.
// K1's fused query preparation (k1_gemv_topk.cu: prepare_query) follows the same order, so K1 and K2 see
// bit-identical queries.  Round 1's two-pass scalar kernel ran at 43 % of the HBM rate.
template <bool F32, int NV>
__global__ void __launch_bounds__(256) ingest_stream_kernel(const float* __restrict__ src, int64_t n, int dim, int ld,
                                                            int normalize, void* __restrict__ dst, float* __restrict__ sqnorm) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int dim4 = dim >> 2, ld4 = ld >> 2;
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
        const float4* x4 = reinterpret_cast<const float4*>(src + row * (int64_t)dim);
        float4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = lane + 32 * i;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < dim4) {
                const uint4 u = ldg_stream(reinterpret_cast<const uint4*>(x4 + j));
                v[i] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
            }
        }
        double ss = 0.0;
        if (normalize) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                ss += (double)v[i].x * (double)v[i].x;
                ss += (double)v[i].y * (double)v[i].y;
                ss += (double)v[i].z * (double)v[i].z;
                ss += (double)v[i].w * (double)v[i].w;
            }
            ss = warp_sum(ss);
        }
        const bool scale = normalize && ss > 0.0;
        // one fp64 reciprocal per row, one fp64 multiply per element: the fp64 divide per element of round 1 kept the
        // kernel on the fp64 pipe instead of HBM.  x * (1/n) and x / n differ by at most one fp64 ulp, which changes the
        // fp32 rounding of about one element in 2^29 (the oracle divides; test_ingest_matches_oracle_bits allows 1e-6).
        const double inv = scale ? 1.0 / sqrt(ss) : 1.0;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = lane + 32 * i;
            if (j >= ld4) continue;
            float y[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
            if (scale) {
#pragma unroll
                for (int c = 0; c < 4; ++c) y[c] = (float)((double)y[c] * inv);
            }
            if (F32) {
                reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + row * (int64_t)ld)[j] = make_float4(y[0], y[1], y[2], y[3]);
#pragma unroll
                for (int c = 0; c < 4; ++c) sq = fmaf(y[c], y[c], sq);
            } else {
                uint32_t b[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const __nv_bfloat16 h = __float2bfloat16_rn(y[c]);
                    b[c] = (uint32_t)__bfloat16_as_ushort(h);
                    const float yr = __bfloat162float(h);
                    sq = fmaf(yr, yr, sq);
                }
                reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + row * (int64_t)ld)[j] =
                    make_uint2(b[0] | (b[1] << 16), b[2] | (b[3] << 16));
            }
        }
        sq = warp_sum(sq);
        if (lane == 0 && sqnorm) sqnorm[row] = sq;
    }
}

template <bool F32>
static cudaError_t launch_ingest_t(const float* src, int64_t n, int dim, int ld, int normalize, void* dst, float* sqnorm,
                                   cudaStream_t st) {
    const int threads = 256;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const bool aligned = (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    const int nv = (ld + 127) / 128;
    if (aligned && nv <= 16) {
#define YRB_INGEST(NV) ingest_stream_kernel<F32, NV><<<(unsigned)blocks, threads, 0, st>>>(src, n, dim, ld, normalize, dst, sqnorm)
        if (nv <= 1) YRB_INGEST(1);
        else if (nv <= 2) YRB_INGEST(2);
        else if (nv <= 4) YRB_INGEST(4);
        else if (nv <= 8) YRB_INGEST(8);
        else YRB_INGEST(16);
#undef YRB_INGEST
    } else {
        ingest_kernel<F32><<<(unsigned)blocks, threads, 0, st>>>(src, n, dim, ld, normalize, dst, sqnorm);
    }
    return cudaGetLastError();
}

cudaError_t launch_ingest(const float* src, int64_t n, int dim, int ld, int metric, int dtype, void* dst,
                          float* sqnorm, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int normalize = (metric == 0);
    return dtype == 1 ? launch_ingest_t<true>(src, n, dim, ld, normalize, dst, sqnorm, st)
                      : launch_ingest_t<false>(src, n, dim, ld, normalize, dst, sqnorm, st);
}

}  // namespace yrb
