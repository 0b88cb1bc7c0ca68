// K5 — ingest: fp32 rows → (L2-normalise for cosine) → storage dtype, one warp per row.
// Replaces what chromadb/hnswlib do inside collection.add (chroma_store.py:86) and
// faiss.normalize_L2 + index.add (faiss_store.py:107-110).  Also used to prepare queries
// (faiss_store.py:146-149 normalises the query the same way).
//
// Reproducibility pin: sum of squares and the divide are fp64, the quotient is
// rounded fp64→fp32→bf16(RNE), so the oracle (oracle/exact_search.py:prepare) lands on the same
// stored bits.  HBM-bound: 4·dim B read + 2·ld B written per row.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace yrb {

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
        const double nrm = scale ? sqrt(ss) : 1.0;
        float sq = 0.f;
        for (int i = lane; i < ld; i += 32) {
            float y = 0.f;
            if (i < dim) y = scale ? (float)((double)x[i] / nrm) : x[i];
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

cudaError_t launch_ingest(const float* src, int64_t n, int dim, int ld, int metric, int dtype, void* dst,
                          float* sqnorm, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const int normalize = (metric == 0);
    if (dtype == 1)
        ingest_kernel<true><<<(unsigned)blocks, threads, 0, st>>>(src, n, dim, ld, normalize, dst, sqnorm);
    else
        ingest_kernel<false><<<(unsigned)blocks, threads, 0, st>>>(src, n, dim, ld, normalize, dst, sqnorm);
    return cudaGetLastError();
}

}  // namespace yrb
