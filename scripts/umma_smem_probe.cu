// umma_smem_probe.cu — stand-alone microbenchmark for the one untested explanation of K2's 0.50 ms:
// is a CTA that issues SS-mode 128x128x16 bf16 MMAs (both operands read from shared memory: 8 KiB per MMA)
// while bulk copies fill the same shared memory at K2's rate (48 KiB per 8 MMAs) limited by shared-memory bandwidth?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I youtu-rag_b200/csrc -o /tmp/umma_probe scripts/umma_smem_probe.cu
//   /tmp/umma_probe            # prints clk per k-block (8 MMAs; 512 clk = tensor pipe saturated) for each mode
//
// Modes (one CTA per SM, operands are whatever the buffers hold — the result is discarded):
//   0  MMAs only, operands resident                           → the MMA issue/operand-read ceiling
//   1  MMAs + bulk copies global(L2-resident)→shared, 48 KiB per k-block, NOT waited for by the MMAs
//   2  bulk copies only
//   3  as 1 but 16 KiB per k-block (rows only: what a query-resident design would fill)
// If mode 1 is ≈ mode 0 the shared-memory hypothesis is dead; if it degrades towards K2's 56-60 % it is the limiter.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "k2_common.cuh"

using namespace yrb::k2;

constexpr int P_STAGES = 4;
constexpr int P_STAGE_BYTES = 2 * QTILE_BYTES + RTILE_BYTES;  // 48 KiB, K2's stage for 256 queries

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const unsigned char* __restrict__ src, size_t src_bytes, int iters, int mode,
                                                int fill_bytes, long long* __restrict__ out_clk) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[2 * P_STAGES + 2];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem) + 1023u) & ~1023u;
    auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
    auto mma_bar = [&](int s) { return smem_u32(&bars[P_STAGES + s]); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < P_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(mma_bar(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < P_STAGES * P_STAGE_BYTES / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem + (smem0 - smem_u32(smem)))[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes above → visible to the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const long long t0 = clock64();
    if (warp == 0 && lane == 0 && mode != 0 && mode != 4) {
        // filler: fill_bytes per k-block into the ring, paced only by its own completion two stages back
        const unsigned char* base = src + ((size_t)blockIdx.x * 65536) % (src_bytes - (size_t)P_STAGE_BYTES);
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            if (it >= P_STAGES) mbar_wait(full_bar(s), ph ^ 1);  // the previous fill of this stage has landed
            mbar_expect_tx(full_bar(s), fill_bytes);
            for (int o = 0; o < fill_bytes; o += 16384) bulk_g2s(smem0 + s * P_STAGE_BYTES + o, base + o, 16384, full_bar(s));
            if (++s == P_STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
        for (int k = 0; k < P_STAGES && k < iters; ++k) {  // drain
            mbar_wait(full_bar(s), ph ^ 1);
            if (++s == P_STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1 && lane == 0 && mode != 2) {
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < iters; ++it) {
            if (it >= P_STAGES) mbar_wait(mma_bar(s), ph ^ 1);  // bound the MMA queue: wait for the k-block issued 4 ago
            const uint32_t a0 = smem0 + s * P_STAGE_BYTES;
            if (mode >= 4) {
                // 128 queries x 256 rows per instruction: A = 16 KiB query block, B = 32 KiB (two row tiles back to back)
                constexpr uint32_t IDESC256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint64_t bdesc = smem_desc(a0 + QTILE_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < BLOCK_K / UMMA_K; ++k4)
                    umma_bf16(tmem_base, smem_desc(a0) + 2 * k4, bdesc + 2 * k4, IDESC256, 1u);
            } else {
                const uint64_t bdesc = smem_desc(a0 + 2 * QTILE_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < BLOCK_K / UMMA_K; ++k4) {
#pragma unroll
                    for (int qb = 0; qb < 2; ++qb)
                        umma_bf16(tmem_base + qb * BLOCK_R, smem_desc(a0 + qb * QTILE_BYTES) + 2 * k4, bdesc + 2 * k4, IDESC, 1u);
                }
            }
            umma_commit(mma_bar(s));
            if (++s == P_STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
        for (int k = 0; k < P_STAGES && k < iters; ++k) {
            mbar_wait(mma_bar(s), ph ^ 1);
            if (++s == P_STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out_clk[blockIdx.x] = t1 - t0;
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t src_bytes = 64ull << 20;  // L2-resident source for the fills
    unsigned char* src = nullptr;
    long long* clk = nullptr;
    CK(cudaMalloc(&src, src_bytes));
    CK(cudaMemset(src, 0x3c, src_bytes));
    CK(cudaMalloc(&clk, sizeof(long long) * sms));
    const size_t smem = (size_t)P_STAGES * P_STAGE_BYTES + 1024;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 20000;
    const struct { int mode, fill; const char* what; } runs[] = {
        {0, 0, "MMAs only (8 x 128x128x16 per k-block)"},
        {1, P_STAGE_BYTES, "MMAs + 48 KiB fills per k-block (K2, 256 queries)"},
        {3, RTILE_BYTES, "MMAs + 16 KiB fills per k-block (rows only)"},
        {2, P_STAGE_BYTES, "48 KiB fills only"},
        {4, 0, "N=256: MMAs only (4 x 128x256x16 per k-block)"},
        {5, P_STAGE_BYTES, "N=256: MMAs + 48 KiB fills per k-block"},
        {5, 2 * RTILE_BYTES, "N=256: MMAs + 32 KiB fills per k-block"},
    };
    for (const auto& r : runs) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        float ms = 0.f;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            probe<<<sms, 128, smem>>>(src, src_bytes, iters, r.mode, r.fill, clk);
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        long long* h = (long long*)malloc(sizeof(long long) * sms);
        CK(cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
        double avg = 0, mx = 0;
        for (int i = 0; i < sms; ++i) {
            avg += (double)h[i] / sms;
            if ((double)h[i] > mx) mx = (double)h[i];
        }
        printf("%-52s  clk per k-block: avg %.1f  max %.1f   (512 = tensor pipe saturated)  %.2f ms, SM clock ~%.0f MHz\n", r.what, avg / iters, mx / iters, ms, mx / ms * 1e-3);
        free(h);
    }
    return 0;
}
