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
constexpr int MAX_Q = 256;     // queries per launch (2 query blocks)
constexpr int QTILE_BYTES = BLOCK_Q * BLOCK_K * 2;  // 16 KiB: one query block x one k-block
constexpr int RTILE_BYTES = BLOCK_R * BLOCK_K * 2;  // 16 KiB: one row tile x one k-block
constexpr int MAX_TOPS = 2;

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


}  // namespace k2
}  // namespace yrb
