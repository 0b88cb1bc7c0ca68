// K8 — row compaction for low-selectivity batched search (BASELINE config C4 with a query batch).
//
// K2 is a dense GEMM: with a 10 % metadata mask it would still stream and multiply every row and drop
// 90 % of the scores in the epilogue.  When a filter shared by the whole batch passes at most a quarter of
// the rows, the passing rows are first gathered into a contiguous scratch matrix (whole 2 KiB rows, warp per
// row, coalesced both ways) together with their row numbers; K2 then runs unmasked on the compact matrix
// and the selected keys are mapped back.  Compaction keeps row order, so the (score desc, id asc) rule is
// unaffected.  HBM-bound byte work: N/8 mask bytes + 2·sel·N·ld·s row bytes (read + write).
// Pre-filter semantics of collection.query(where=…) (chroma_store.py:118-120) are unchanged.
#include "common.cuh"
#include "kernels.h"

namespace yrb {

constexpr int CP_WORDS = 1024;  // mask words per block of the scan (32768 rows)

// per-block popcount
__global__ void __launch_bounds__(256) compact_count_kernel(const uint32_t* __restrict__ mask, int64_t n_words,
                                                            uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t ws[8];
    const int64_t w0 = (int64_t)blockIdx.x * CP_WORDS;
    uint32_t c = 0;
    for (int i = threadIdx.x; i < CP_WORDS; i += 256)
        if (w0 + i < n_words) c += __popc(mask[w0 + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(YRB_FULL, c, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int i = 0; i < 8; ++i) t += ws[i];
        block_sums[blockIdx.x] = t;
    }
}

// exclusive scan of the block sums (single CTA; nb <= 1 << 20) and the total
__global__ void __launch_bounds__(1024) compact_scan_kernel(uint32_t* __restrict__ block_sums, int nb,
                                                            unsigned long long* __restrict__ total) {
    __shared__ uint32_t carry_s;
    __shared__ uint32_t tile[1024];
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nb ? block_sums[i] : 0u;
        tile[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const uint32_t add = threadIdx.x >= off ? tile[threadIdx.x - off] : 0u;
            __syncthreads();
            tile[threadIdx.x] += add;
            __syncthreads();
        }
        const uint32_t incl = tile[threadIdx.x], carry = carry_s;
        if (i < nb) block_sums[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

// rowmap[compact index] = original row, in row order
__global__ void __launch_bounds__(256) compact_map_kernel(const uint32_t* __restrict__ mask, int64_t n_words,
                                                          const uint32_t* __restrict__ block_off,
                                                          uint32_t* __restrict__ rowmap) {
    const int64_t w0 = (int64_t)blockIdx.x * CP_WORDS;
    // exclusive prefix of word popcounts inside the block: 4 words per thread + block scan over 256 threads
    uint32_t wv[4], wc[4], mine = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t w = w0 + threadIdx.x * 4 + j;
        wv[j] = w < n_words ? mask[w] : 0u;
        wc[j] = __popc(wv[j]);
        mine += wc[j];
    }
    __shared__ uint32_t tsum[256];
    tsum[threadIdx.x] = mine;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const uint32_t add = threadIdx.x >= off ? tsum[threadIdx.x - off] : 0u;
        __syncthreads();
        tsum[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = block_off[blockIdx.x] + tsum[threadIdx.x] - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t m = wv[j];
        const uint32_t row0 = (uint32_t)((w0 + threadIdx.x * 4 + j) * 32);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            rowmap[run++] = row0 + b;
        }
    }
}

// warp per compact row: copy the row (ld16 uint4) and its squared norm
__global__ void __launch_bounds__(256) compact_gather_kernel(const uint4* __restrict__ rows, const float* __restrict__ sqnorm,
                                                             const uint32_t* __restrict__ rowmap, int64_t n_out, int ld16,
                                                             uint4* __restrict__ out_rows, float* __restrict__ out_sqnorm) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n_out; i += warps) {
        const uint32_t r = rowmap[i];
        const uint4* src = rows + (int64_t)r * ld16;
        uint4* dst = out_rows + i * ld16;
        for (int c = lane; c < ld16; c += 32) dst[c] = ldg_stream(src + c);
        if (lane == 0) out_sqnorm[i] = sqnorm[r];
    }
}

// row-subset mode of K2 (TMA gather4): only the squared norms are gathered, the rows are read in place
__global__ void __launch_bounds__(256) compact_sqnorm_kernel(const float* __restrict__ sqnorm, const uint32_t* __restrict__ rowmap,
                                                             int64_t n_out, float* __restrict__ out_sqnorm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) out_sqnorm[i] = sqnorm[rowmap[i]];
}

// keys / ids produced on the compact matrix → original row numbers
__global__ void __launch_bounds__(256) compact_remap_kernel(const uint32_t* __restrict__ rowmap, int64_t n, uint64_t* __restrict__ keys,
                                                            int64_t* __restrict__ ids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keys) {
        const uint64_t key = keys[i];
        if (key) keys[i] = (key & 0xffffffff00000000ull) | (uint64_t)(~rowmap[key_row(key)]);
    }
    if (ids) {
        const int64_t r = ids[i];
        if (r >= 0) ids[i] = (int64_t)rowmap[r];
    }
}

size_t compact_scratch_words(int64_t n_rows) { return (size_t)((n_rows + 32 * CP_WORDS - 1) / (32 * CP_WORDS)); }

cudaError_t launch_compact_count(const uint32_t* mask, int64_t n_rows, uint32_t* block_sums, unsigned long long* total,
                                 cudaStream_t st) {
    const int64_t n_words = (n_rows + 31) / 32;
    const int nb = (int)compact_scratch_words(n_rows);
    compact_count_kernel<<<nb, 256, 0, st>>>(mask, n_words, block_sums);
    compact_scan_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
    return cudaGetLastError();
}

cudaError_t launch_compact_gather(const uint32_t* mask, int64_t n_rows, const uint32_t* block_off, uint32_t* rowmap,
                                  const void* rows, const float* sqnorm, int ld16, int64_t n_out, void* out_rows,
                                  float* out_sqnorm, int sm_count, cudaStream_t st) {
    const int64_t n_words = (n_rows + 31) / 32;
    const int nb = (int)compact_scratch_words(n_rows);
    compact_map_kernel<<<nb, 256, 0, st>>>(mask, n_words, block_off, rowmap);
    if (n_out > 0 && !out_rows) {  // map + norms only
        compact_sqnorm_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(sqnorm, rowmap, n_out, out_sqnorm);
        return cudaGetLastError();
    }
    if (n_out > 0)
        compact_gather_kernel<<<sm_count * 8, 256, 0, st>>>(reinterpret_cast<const uint4*>(rows), sqnorm, rowmap, n_out, ld16,
                                                            reinterpret_cast<uint4*>(out_rows), out_sqnorm);
    return cudaGetLastError();
}

cudaError_t launch_compact_remap(const uint32_t* rowmap, int64_t n, uint64_t* keys, int64_t* ids, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    compact_remap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowmap, n, keys, ids);
    return cudaGetLastError();
}

}  // namespace yrb
