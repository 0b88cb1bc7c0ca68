// K7 — the top-k merge collective as ONE kernel over NVLink peer memory.
//
// Replaces the NCCL all-gather + merge launch pair: every rank's CTA q stores its query's k selection
// keys straight into the gather buffer of every peer (plain st.global on CUDA-IPC-mapped peer memory,
// i.e. NVLink / NVSwitch stores), publishes a per-(source, query) flag with release semantics, waits for
// the same flag from all peers, and merges the world·k candidates by (score, GLOBAL id) — transfer,
// synchronisation and merge in a single launch, no host-side collective enqueue.
// The reference has no counterpart (one process, one collection: chroma_store.py:41-59).
//
// Buffers live in the library (cudaMalloc) and are exported / opened with CUDA IPC handles that the
// Python side passes around with torch.distributed (plumbing only).  Slots and flags are double-buffered by
// the parity of a monotonically increasing epoch: a rank can be at most one exchange ahead of its peers
// (it needs their flags of epoch e+1, which they raise only after finishing epoch e).
// Payload is KiB-scale (C5: 80 KiB per rank): latency-bound, a few microseconds.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "index_internal.h"
#include "common.cuh"

namespace yrb {

constexpr int EX_THREADS = 256;
constexpr int EX_CAP = 2048;  // world * k candidates per query

struct GKeyX {
    uint32_t sbits, valid;
    int64_t gid;
};
struct BetterGX {
    __device__ __forceinline__ bool operator()(const GKeyX& a, const GKeyX& b) const {
        if (a.valid != b.valid) return a.valid > b.valid;
        if (a.sbits != b.sbits) return a.sbits > b.sbits;
        return a.gid < b.gid;
    }
};

struct ExArgs {
    uint64_t* slots[8];     // slots[p] = rank p's gather buffer  [2][world][nq_cap][k_cap]  (p == rank: local)
    uint32_t* flags[8];     // flags[p] = rank p's flag array      [2][world][nq_cap]
    int world, rank, nq, k, nq_cap, k_cap;
    uint32_t epoch;
    const uint64_t* local_keys;  // [nq][k]
    const int64_t* row_base;     // [world]
    int64_t* ids;
    float* scores;
    int32_t* counts;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// A bounded grid walks the queries (CTA c takes q = c, c + grid, …): every rank visits the queries in the same
// order, so the lowest unfinished query of every rank is always held by a resident CTA and the flag waits cannot
// depend on how many CTAs the scheduler keeps resident beside a running scan.
constexpr int EX_MAX_GRID = 64;
constexpr long long EX_TIMEOUT_CLK = 120000000000ll;  // ~60 s: a peer that is this late is gone; trap instead of hanging

// CAP = candidates per query this instance can sort.  The small instance (1024 candidates = 16 KiB of shared memory)
// fits on an SM beside a K2 CTA (197.6 KiB), the large one does not; with world <= 8 and k <= 128 the small one is always
// enough.  (Measured: no effect on 8-GPU batch steps — the co-residency was not what limited them)
template <int CAP>
__global__ void __launch_bounds__(EX_THREADS) exchange_merge_kernel(ExArgs a) {
    __shared__ GKeyX sk[CAP];
    __shared__ int cnt;
    const int tid = threadIdx.x;
    const int par = a.epoch & 1u;
    for (int q = blockIdx.x; q < a.nq; q += gridDim.x) {
        const size_t slot_q = ((size_t)par * a.world + a.rank) * a.nq_cap + q;  // where MY keys land in every buffer
        // 1. push this query's keys into every rank's buffer (own copy included)
        for (int i = tid; i < a.world * a.k; i += EX_THREADS) {
            const int p = i / a.k, j = i - p * a.k;
            a.slots[p][slot_q * a.k_cap + j] = a.local_keys[(size_t)q * a.k + j];
        }
        __threadfence_system();
        __syncthreads();
        // 2. raise my flag on every rank, then wait for every rank's flag here
        if (tid < a.world) st_release_sys(a.flags[tid] + slot_q, a.epoch);
        if (tid < a.world) {
            const uint32_t* f = a.flags[a.rank] + ((size_t)par * a.world + tid) * a.nq_cap + q;
            const long long t0 = clock64();
            while ((int32_t)(ld_acquire_sys(f) - a.epoch) < 0) {
                if (clock64() - t0 > EX_TIMEOUT_CLK) __trap();
            }
        }
        __syncthreads();
        // 3. merge world * k candidates by (score desc, global id asc)
        const int n = a.world * a.k;
        const int npow = next_pow2(n);
        if (tid == 0) cnt = 0;
        const uint64_t* mine = a.slots[a.rank];
        for (int i = tid; i < npow; i += EX_THREADS) {
            GKeyX g{0u, 0u, 0};
            if (i < n) {
                const int p = i / a.k, j = i - p * a.k;
                const uint64_t key = __ldcg(mine + (((size_t)par * a.world + p) * a.nq_cap + q) * a.k_cap + j);
                if (key != 0ull) {
                    g.sbits = (uint32_t)(key >> 32);
                    g.valid = 1u;
                    g.gid = a.row_base[p] + (int64_t)key_row(key);
                }
            }
            sk[i] = g;
        }
        block_bitonic_desc(sk, npow, BetterGX());
        int local = 0;
        for (int i = tid; i < a.k; i += EX_THREADS) {
            const bool ok = (i < n) && sk[i].valid;
            a.ids[(size_t)q * a.k + i] = ok ? sk[i].gid : -1;
            a.scores[(size_t)q * a.k + i] = ok ? bits_score(sk[i].sbits) : -INFINITY;
            local += ok;
        }
        if (local) atomicAdd(&cnt, local);
        __syncthreads();
        if (tid == 0 && a.counts) a.counts[q] = cnt;
        __syncthreads();  // sk / cnt are reused by the next query
    }
}

}  // namespace yrb

struct yrb_exchange {
    int device = 0, world = 1, rank = 0, nq_cap = 0, k_cap = 0;
    uint64_t* slots = nullptr;  // local
    uint32_t* flags = nullptr;  // local
    uint64_t* peer_slots[8] = {};
    uint32_t* peer_flags[8] = {};
    bool opened[8] = {};
    uint32_t epoch = 0;
};

namespace {
thread_local std::string x_err;
int xfail(int code, const std::string& m) {
    x_err = m;
    return code;
}
}  // namespace

extern "C" {

const char* yrb_exchange_last_error(void) { return x_err.c_str(); }

int yrb_exchange_create(yrb_exchange** out, int device, int world, int rank, int nq_cap, int k_cap, unsigned char* out_handles) {
    if (!out || !out_handles) return xfail(YRB_ERR_INVALID, "NULL argument");
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return xfail(YRB_ERR_INVALID, "world must be 1..8 and rank in range");
    if ((int64_t)world * k_cap > yrb::EX_CAP) return xfail(YRB_ERR_UNSUPPORTED, "world * k exceeds 2048");
    yrbi::DevGuard dev_guard_;
    if (cudaSetDevice(device) != cudaSuccess) return xfail(YRB_ERR_CUDA, "cudaSetDevice failed");
    yrb_exchange* ex = new yrb_exchange();
    ex->device = device;
    ex->world = world;
    ex->rank = rank;
    ex->nq_cap = nq_cap;
    ex->k_cap = k_cap;
    const size_t sb = (size_t)2 * world * nq_cap * k_cap * 8, fb = (size_t)2 * world * nq_cap * 4;
    cudaError_t e = cudaMalloc(&ex->slots, sb);
    if (e == cudaSuccess) e = cudaMalloc(&ex->flags, fb);
    if (e == cudaSuccess) e = cudaMemset(ex->slots, 0, sb);
    if (e == cudaSuccess) e = cudaMemset(ex->flags, 0, fb);
    cudaIpcMemHandle_t hs, hf;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hs, ex->slots);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hf, ex->flags);
    if (e != cudaSuccess) {
        std::string m = std::string("exchange buffers: ") + cudaGetErrorString(e);
        if (ex->slots) cudaFree(ex->slots);
        if (ex->flags) cudaFree(ex->flags);
        delete ex;
        return xfail(YRB_ERR_CUDA, m);
    }
    memcpy(out_handles, &hs, sizeof hs);
    memcpy(out_handles + sizeof hs, &hf, sizeof hf);
    *out = ex;
    return YRB_OK;
}

int yrb_exchange_handle_bytes(void) { return 2 * (int)sizeof(cudaIpcMemHandle_t); }

int yrb_exchange_connect(yrb_exchange* ex, const unsigned char* all_handles) {
    if (!ex || !all_handles) return xfail(YRB_ERR_INVALID, "NULL argument");
    yrbi::DevGuard dev_guard_;
    if (cudaSetDevice(ex->device) != cudaSuccess) return xfail(YRB_ERR_CUDA, "cudaSetDevice failed");
    const int hb = yrb_exchange_handle_bytes();
    for (int p = 0; p < ex->world; ++p) {
        if (p == ex->rank) {
            ex->peer_slots[p] = ex->slots;
            ex->peer_flags[p] = ex->flags;
            continue;
        }
        cudaIpcMemHandle_t hs, hf;
        memcpy(&hs, all_handles + (size_t)p * hb, sizeof hs);
        memcpy(&hf, all_handles + (size_t)p * hb + sizeof hs, sizeof hf);
        void *ps = nullptr, *pf = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ps, hs, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return xfail(YRB_ERR_CUDA, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(p) + "): " + cudaGetErrorString(e));
        ex->peer_slots[p] = static_cast<uint64_t*>(ps);
        ex->peer_flags[p] = static_cast<uint32_t*>(pf);
        ex->opened[p] = true;
    }
    return YRB_OK;
}

int yrb_exchange_merge(yrb_exchange* ex, const uint64_t* dev_local_keys, int nq, int k, const int64_t* dev_row_base,
                       int64_t* dev_out_ids, float* dev_out_scores, int32_t* dev_out_counts, void* stream) {
    if (!ex || !dev_local_keys || !dev_row_base || !dev_out_ids || !dev_out_scores) return xfail(YRB_ERR_INVALID, "NULL argument");
    if (nq < 1 || nq > ex->nq_cap || k < 1 || k > ex->k_cap) return xfail(YRB_ERR_INVALID, "nq / k exceed the exchange's capacity");
    yrbi::DevGuard dev_guard_;
    if (cudaSetDevice(ex->device) != cudaSuccess) return xfail(YRB_ERR_CUDA, "cudaSetDevice failed");
    yrb::ExArgs a{};
    for (int p = 0; p < ex->world; ++p) {
        a.slots[p] = ex->peer_slots[p];
        a.flags[p] = ex->peer_flags[p];
    }
    a.world = ex->world;
    a.rank = ex->rank;
    a.nq = nq;
    a.k = k;
    a.nq_cap = ex->nq_cap;
    a.k_cap = ex->k_cap;
    a.epoch = ++ex->epoch;
    a.local_keys = dev_local_keys;
    a.row_base = dev_row_base;
    a.ids = dev_out_ids;
    a.scores = dev_out_scores;
    a.counts = dev_out_counts;
    if (ex->world * k <= 1024)
        yrb::exchange_merge_kernel<1024><<<std::min(nq, yrb::EX_MAX_GRID), yrb::EX_THREADS, 0, (cudaStream_t)stream>>>(a);
    else
        yrb::exchange_merge_kernel<yrb::EX_CAP><<<std::min(nq, yrb::EX_MAX_GRID), yrb::EX_THREADS, 0, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return xfail(YRB_ERR_CUDA, std::string("exchange_merge_kernel: ") + cudaGetErrorString(e));
    return YRB_OK;
}

// A whole sharded search of one rank through host buffers, enqueued from C on ONE stream: query upload, local scan +
// top-k, the exchange/merge kernel writing the merged result into this rank's pinned result buffer (straight into host
// memory when it is small), one synchronisation.  What `ShardedSearcher.search` used to assemble from torch calls
// (60-90 us of Python per search at 8 GPUs).
int yrb_exchange_search(yrb_exchange* ex, yrb_index* ix, const float* queries, int nq, int k, const uint32_t* dev_mask,
                        const int64_t* dev_row_base, int64_t* out_ids, float* out_scores, int32_t* out_counts) {
    using namespace yrbi;
    Nvtx nvtx_("yrb_exchange_search");
    if (!ex || !ix || !queries || !dev_row_base || !out_ids || !out_scores) return fail(YRB_ERR_INVALID, "NULL argument");
    if (nq < 1 || nq > ex->nq_cap || k < 1 || k > ex->k_cap) return fail(YRB_ERR_INVALID, "nq / k exceed the exchange's capacity");
    if (ix->device != ex->device) return fail(YRB_ERR_INVALID, "index and exchange live on different devices");
    if (k > ix->rows) return fail(YRB_ERR_INVALID, "k=%d exceeds this shard's rows=%lld", k, (long long)ix->rows);
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    if ((rc = ensure_scratch(ix, nq, k))) return rc;
    cudaStream_t st = ix->stream;
    memcpy(ix->h_q, queries, (size_t)nq * ix->dim * 4);
    CK(cudaMemcpyAsync(ix->d_qf32, ix->h_q, (size_t)nq * ix->dim * 4, cudaMemcpyHostToDevice, st));
    const uint32_t* m = nullptr;
    if ((rc = resolve_mask(ix, nullptr, dev_mask, &m, st, false))) return rc;
    if ((rc = scan_select(ix, ix->d_qf32, nq, k, m, 0, ix->d_keys, nullptr, nullptr, nullptr, st))) return rc;
    const size_t res_bytes = (size_t)nq * k * 12 + (size_t)nq * 4;
    const bool zero_copy = ix->d_result_host && res_bytes <= 4096;
    unsigned char* base = zero_copy ? ix->d_result_host : ix->d_result;
    int64_t* d_ids = reinterpret_cast<int64_t*>(base);
    float* d_scores = reinterpret_cast<float*>(base + (size_t)nq * k * 8);
    int32_t* d_counts = reinterpret_cast<int32_t*>(base + (size_t)nq * k * 12);
    if ((rc = yrb_exchange_merge(ex, ix->d_keys, nq, k, dev_row_base, d_ids, d_scores, d_counts, st))) return fail(rc, "%s", x_err.c_str());
    ix->launches++;
    if (!zero_copy) CK(cudaMemcpyAsync(ix->h_result, ix->d_result, res_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out_ids, ix->h_result, (size_t)nq * k * 8);
    memcpy(out_scores, ix->h_result + (size_t)nq * k * 8, (size_t)nq * k * 4);
    if (out_counts) memcpy(out_counts, ix->h_result + (size_t)nq * k * 12, (size_t)nq * 4);
    return YRB_OK;
}

int yrb_exchange_destroy(yrb_exchange* ex) {
    if (!ex) return YRB_OK;
    yrbi::DevGuard dev_guard_;
    cudaSetDevice(ex->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < ex->world; ++p)
        if (ex->opened[p]) {
            cudaIpcCloseMemHandle(ex->peer_slots[p]);
            cudaIpcCloseMemHandle(ex->peer_flags[p]);
        }
    if (ex->slots) cudaFree(ex->slots);
    if (ex->flags) cudaFree(ex->flags);
    delete ex;
    return YRB_OK;
}

}  // extern "C"
