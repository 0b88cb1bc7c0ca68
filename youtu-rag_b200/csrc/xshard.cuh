// xshard.cuh — the cross-GPU top-k merge folded into the kernel that finishes a query.
//
// One process, one collection row-sharded over the GPUs of a box (sharded.cu).  Every shard runs its own scan +
// top-k; the CTA that completes a query on a shard (K1's last CTA, K3's selecting CTA) then
//   1. stores the shard's k keys into the gather buffer on the ROOT device (plain stores — NVLink peer stores when
//      the shard is not the root),
//   2. takes a ticket from the query's counter on the root (system-scope atomic over NVLink),
//   3. and, if it holds the last ticket, stages all shards' lists (one round of peer loads), merges them by
//      (score desc, GLOBAL row id asc) — ranks by binary search, no sort — and writes ids / scores / count to the
//      caller's result buffer (pinned host memory, mapped: no copy-engine round trip), then bumps `done`.
// So a sharded search is ONE launch per shard with no separate exchange kernel, no NCCL call and no host-side
// merge; the host only waits for `done == nq`.  (The torchrun form — one process per GPU, every rank needs the
// result in stream order — keeps K7, k7_exchange.cu.)  The reference has no counterpart: one Chroma collection in
// one process (utu/rag/storage/implementations/chroma_store.py:41-59).
//
// Rows are dealt to shards block-cyclically: global row g sits in block b = g >> block_shift on shard b % n_shards at
// local row ((b / n_shards) << block_shift) | (g & mask).  Local order is monotone in global order, so each shard's
// list — sorted by (score desc, local row asc) — is also sorted by (score desc, global row asc).
#pragma once

#include <stdint.h>

#include "common.cuh"

namespace yrb {

constexpr int XS_MAX_SHARDS = 8;

struct XShard {
    uint64_t* gather;       // root device: [nq][n_shards][k] keys of this search
    unsigned int* tickets;  // root device: [nq], zero between searches
    int64_t* out_ids;       // [nq][k_out]  (-1 padding)
    float* out_scores;      // [nq][k_out]  (-inf padding)
    int32_t* out_counts;    // [nq]
    unsigned int* done;     // queries finished (system scope)
    int n_shards, shard;
    int n_active;           // shards taking part in this search (those that hold rows)
    uint32_t active_mask;
    int block_shift;
    int k;                  // list length in `gather` and in the outputs
    int q0;                 // query index of this launch's first query (K1 runs one launch per query)
};

__host__ __device__ __forceinline__ int64_t xs_global_row(int n_shards, int block_shift, int shard, uint32_t local) {
    const int64_t lb = (int64_t)(local >> block_shift);
    return ((lb * n_shards + shard) << block_shift) | (int64_t)(local & ((1u << block_shift) - 1u));
}

__device__ __forceinline__ uint64_t xs_ld_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// is key a (from shard pa) a better hit than key b (from shard pb)?  0 = empty slot, worse than everything
__device__ __forceinline__ bool xs_better(uint64_t a, int pa, uint64_t b, int pb, int n_shards, int shift) {
    if (b == 0ull) return a != 0ull;
    if (a == 0ull) return false;
    const uint32_t sa = (uint32_t)(a >> 32), sb = (uint32_t)(b >> 32);
    if (sa != sb) return sa > sb;
    return xs_global_row(n_shards, shift, pa, key_row(a)) < xs_global_row(n_shards, shift, pb, key_row(b));
}

// Called by every thread of the CTA that finished query q (launch-relative index) on this shard.  `mine`: the shard's
// sorted keys (n valid, n <= x.k), in shared or global memory written before the call.  `stage`: n_shards * x.k
// uint64 slots of shared memory the caller no longer needs.  Ends with a __syncthreads().
__device__ __forceinline__ void xshard_finish(const XShard& x, int q, const uint64_t* mine, int n, uint64_t* stage) {
    __shared__ int xs_last;
    const int tid = threadIdx.x, nt = blockDim.x, k = x.k;
    q += x.q0;
    uint64_t* g = x.gather + (size_t)q * x.n_shards * k;
    __syncthreads();  // `mine` is complete; `stage` is free
    for (int i = tid; i < k; i += nt) g[(size_t)x.shard * k + i] = (i < n) ? mine[i] : 0ull;
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd_system(x.tickets + q, 1u);
        xs_last = (t == (unsigned int)x.n_active - 1u);
    }
    __syncthreads();
    if (!xs_last) return;
    __threadfence_system();  // the other shards' stores precede their tickets
    const int total = x.n_shards * k;
    for (int i = tid; i < total; i += nt) stage[i] = ((x.active_mask >> (i / k)) & 1u) ? xs_ld_sys(g + i) : 0ull;
    __syncthreads();
    // rank of every key = keys better than it: its index in its own list + a binary search in every other list
    int valid = 0;
    for (int e = tid; e < total; e += nt) {
        const uint64_t key = stage[e];
        if (key == 0ull) continue;
        ++valid;
        const int p = e / k;
        int rank = e - p * k;
        for (int o = 0; o < x.n_shards; ++o) {
            if (o == p) continue;
            const uint64_t* lst = stage + o * k;
            int lo = 0, hi = k;  // first index whose entry is NOT better than key
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (xs_better(lst[mid], o, key, p, x.n_shards, x.block_shift)) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            x.out_ids[(size_t)q * k + rank] = xs_global_row(x.n_shards, x.block_shift, p, key_row(key));
            x.out_scores[(size_t)q * k + rank] = key_score(key);
        }
    }
    // total number of valid keys (block-wide sum of `valid`)
    __shared__ int xs_total;
    if (tid == 0) xs_total = 0;
    __syncthreads();
    if (valid) atomicAdd(&xs_total, valid);
    __syncthreads();
    const int count = xs_total < k ? xs_total : k;
    for (int i = count + tid; i < k; i += nt) {
        x.out_ids[(size_t)q * k + i] = -1;
        x.out_scores[(size_t)q * k + i] = -INFINITY;
    }
    if (tid == 0) x.out_counts[q] = count;
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        x.tickets[q] = 0u;  // every shard has arrived; the next search is enqueued after the host saw `done`
        __threadfence_system();
        atomicAdd_system(x.done, 1u);
    }
}

}  // namespace yrb
