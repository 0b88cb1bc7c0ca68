// sharded.cu — one collection row-sharded over the GPUs of ONE box, inside ONE process.
//
// The agents serve every collection from a single FastAPI process and cache one store per collection
// (utu/rag/rag_tools/base_toolkit.py:79-91; factory utu/rag/storage/base_storage.py:28-42), so the drop-in store
// cannot be a torchrun job: `B200VectorStore(index_params={"devices": [0..7]})` creates a `yrb_sharded` instead of a
// `yrb_index` and everything else stays the same.  A `yrb_sharded` is n `yrb_index` shards plus
//   * a block-cyclic row map (global row g -> block g >> shift -> shard block % n), so a growing collection stays
//     balanced and every mapping is arithmetic (no table);
//   * one worker thread per device that enqueues its shard's part of a search (query upload, K4 filter, scan, top-k)
//     the moment the job is posted — the shards start within microseconds of each other instead of one launch
//     latency apart;
//   * the cross-shard merge folded into the kernel that finishes a query (xshard.cuh): keys travel to the root GPU as
//     NVLink peer stores, the CTA holding the last ticket merges and writes the result straight into pinned host
//     memory, the calling thread just waits for `done == nq`.  No NCCL, no IPC, no exchange launch.
// Peer access between the root device and every other device is required (NVSwitch boxes have it); without it
// create() fails — there is no host-side merge to fall back to.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <thread>

#include "index_internal.h"

using namespace yrbi;

namespace {

enum JobKind { JOB_NONE = 0, JOB_SEARCH = 1, JOB_QUIT = 2 };

struct SearchJob {
    int nq = 0, ke = 0;
    const yrb_where* w = nullptr;                 // shared filter (or NULL)
    const yrb_where* const* per_query = nullptr;  // one per query (or NULL)
    const uint32_t* mask = nullptr;               // caller's host bitmask in GLOBAL row order (or NULL)
    uint32_t active_mask = 0;
    int n_active = 0;
    float min_score = -INFINITY;
};

}  // namespace

struct yrb_sharded {
    int n = 0, root = 0;
    int devices[yrb::XS_MAX_SHARDS] = {};
    yrb_index* shard[yrb::XS_MAX_SHARDS] = {};
    int dim = 0, metric = 0, dtype = 0, ld = 0, block_shift = 14;
    int64_t rows = 0;
    // root-device merge buffers
    uint64_t* d_gather = nullptr;
    size_t gather_keys = 0;
    unsigned int* d_tickets = nullptr;
    int tickets_cap = 0;
    // pinned + mapped + portable host memory: queries in, results out
    float* h_q = nullptr;
    size_t hq_floats = 0;
    unsigned char* h_result = nullptr;
    size_t result_bytes = 0;
    unsigned int* h_done = nullptr;
    // per-shard peer staging for device-side appends
    float* d_peer_stage[yrb::XS_MAX_SHARDS] = {};
    size_t peer_stage_floats[yrb::XS_MAX_SHARDS] = {};
    // workers
    std::thread workers[yrb::XS_MAX_SHARDS];
    std::atomic<uint64_t> job_seq{0};
    std::atomic<int> acks{0};
    std::atomic<int> sleepers{0};
    std::mutex wake_mu;
    std::condition_variable wake_cv;
    int job_kind = JOB_NONE;
    SearchJob job;
    int job_rc[yrb::XS_MAX_SHARDS] = {};
    std::string job_err[yrb::XS_MAX_SHARDS];
    std::mutex mu;  // serialises the public entry points
    int64_t searches = 0;
};

namespace {

inline int64_t block_rows(const yrb_sharded* sh) { return int64_t(1) << sh->block_shift; }

// rows shard s holds when the collection has `total` rows
int64_t rows_on_shard(const yrb_sharded* sh, int64_t total, int s) {
    int64_t r = 0;
    yrb_shard_rows(sh->n, (int)block_rows(sh), total, s, &r);
    return r;
}
inline void locate(const yrb_sharded* sh, int64_t g, int* s, int64_t* local) {
    yrb_shard_locate(sh->n, (int)block_rows(sh), g, s, local);
}
// [g0, g0 + n) cut at block boundaries: fn(shard, local_begin, global_begin, count) → rc
template <class Fn>
int for_pieces(const yrb_sharded* sh, int64_t g0, int64_t n, Fn fn) {
    const int64_t B = block_rows(sh);
    while (n > 0) {
        const int64_t cnt = std::min<int64_t>(n, B - (g0 & (B - 1)));
        int s;
        int64_t local;
        locate(sh, g0, &s, &local);
        const int rc = fn(s, local, g0, cnt);
        if (rc) return rc;
        g0 += cnt;
        n -= cnt;
    }
    return YRB_OK;
}

// ------------------------------------------------------------------ one shard's part of a search (worker thread)
int shard_search(yrb_sharded* sh, int s) {
    Nvtx nvtx_("yrb_sharded_search/shard");
    const SearchJob& j = sh->job;
    yrb_index* ix = sh->shard[s];
    if (!((j.active_mask >> s) & 1u)) return YRB_OK;
    std::lock_guard<std::mutex> g(ix->mu);
    int rc = set_dev(ix);
    if (rc) return rc;
    const int k_s = (int)std::min<int64_t>(j.ke, ix->rows);
    if ((rc = ensure_scratch(ix, j.nq, k_s))) return rc;
    cudaStream_t st = ix->stream;
    CK(cudaMemcpyAsync(ix->d_qf32, sh->h_q, (size_t)j.nq * ix->dim * 4, cudaMemcpyHostToDevice, st));
    const uint32_t* dev_extra = nullptr;
    if (j.mask) {
        // the caller's bitmask is in global row order; blocks are whole mask words (block_rows >= 64)
        const int64_t wpb = block_rows(sh) / 32, lw = (ix->rows + 31) / 32;
        std::vector<uint32_t> local((size_t)lw);
        const int64_t gw_total = (sh->rows + 31) / 32;
        for (int64_t w = 0; w < lw; ++w) {
            const int64_t lb = w / wpb;
            const int64_t gw = (lb * sh->n + s) * wpb + (w % wpb);
            local[w] = gw < gw_total ? j.mask[gw] : 0u;
        }
        if ((rc = upload_user_mask(ix, local.data(), &dev_extra, st))) return rc;
    }
    const uint32_t* m = nullptr;
    int64_t m_stride = 0;
    if (j.per_query) rc = resolve_masks_multi(ix, j.per_query, j.nq, dev_extra, &m, &m_stride, st);
    else rc = resolve_mask(ix, j.w, dev_extra, &m, st, false);
    if (rc) return rc;
    yrb::XShard xs{};
    xs.gather = sh->d_gather;
    xs.tickets = sh->d_tickets;
    xs.out_ids = reinterpret_cast<int64_t*>(sh->h_result);
    xs.out_scores = reinterpret_cast<float*>(sh->h_result + (size_t)j.nq * j.ke * 8);
    xs.out_counts = reinterpret_cast<int32_t*>(sh->h_result + (size_t)j.nq * j.ke * 12);
    xs.done = sh->h_done;
    xs.n_shards = sh->n;
    xs.shard = s;
    xs.n_active = j.n_active;
    xs.active_mask = j.active_mask;
    xs.block_shift = sh->block_shift;
    xs.k = j.ke;
    xs.q0 = 0;
    return scan_select(ix, ix->d_qf32, j.nq, k_s, m, m_stride, ix->d_keys, nullptr, nullptr, nullptr, st, &xs, j.min_score);
}

void worker_main(yrb_sharded* sh, int s) {
    cudaSetDevice(sh->devices[s]);
    uint64_t seen = 0;
    for (;;) {
        // spin for a while (a serving loop posts searches back to back), then sleep
        int spins = 0;
        while (sh->job_seq.load(std::memory_order_acquire) == seen) {
            if (++spins < 40000) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                continue;
            }
            std::unique_lock<std::mutex> lk(sh->wake_mu);
            sh->sleepers.fetch_add(1);
            sh->wake_cv.wait(lk, [&] { return sh->job_seq.load(std::memory_order_acquire) != seen; });
            sh->sleepers.fetch_sub(1);
        }
        seen = sh->job_seq.load(std::memory_order_acquire);
        if (sh->job_kind == JOB_QUIT) return;
        int rc = YRB_OK;
        if (sh->job_kind == JOB_SEARCH) rc = shard_search(sh, s);
        sh->job_rc[s] = rc;
        if (rc) sh->job_err[s] = last_error();
        sh->acks.fetch_add(1, std::memory_order_release);
    }
}

// posts the current job to every worker and waits until all have enqueued their part
void run_job(yrb_sharded* sh, int kind) {
    sh->job_kind = kind;
    sh->acks.store(0, std::memory_order_relaxed);
    sh->job_seq.fetch_add(1, std::memory_order_release);
    if (sh->sleepers.load() > 0) {
        std::lock_guard<std::mutex> lk(sh->wake_mu);
        sh->wake_cv.notify_all();
    }
    if (kind == JOB_QUIT) return;
    while (sh->acks.load(std::memory_order_acquire) < sh->n) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

int ensure_buffers(yrb_sharded* sh, int nq, int ke) {
    const size_t need_keys = (size_t)nq * sh->n * ke;
    CK(cudaSetDevice(sh->devices[sh->root]));
    if (need_keys > sh->gather_keys || nq > sh->tickets_cap) {
        CK(cudaDeviceSynchronize());
        if (sh->d_gather) cudaFree(sh->d_gather);
        if (sh->d_tickets) cudaFree(sh->d_tickets);
        sh->d_gather = nullptr;
        sh->d_tickets = nullptr;
        const size_t keys = std::max(need_keys, sh->gather_keys), tk = (size_t)std::max(nq, sh->tickets_cap);
        CK(cudaMalloc(&sh->d_gather, keys * 8));
        CK(cudaMalloc(&sh->d_tickets, tk * 4));
        CK(cudaMemset(sh->d_tickets, 0, tk * 4));
        sh->gather_keys = keys;
        sh->tickets_cap = (int)tk;
    }
    const size_t qf = (size_t)nq * sh->dim;
    if (qf > sh->hq_floats) {
        if (sh->h_q) cudaFreeHost(sh->h_q);
        sh->h_q = nullptr;
        CK(cudaHostAlloc(&sh->h_q, qf * 4, cudaHostAllocPortable | cudaHostAllocMapped));
        sh->hq_floats = qf;
    }
    const size_t rb = (size_t)nq * ke * 12 + (size_t)nq * 4;
    if (rb > sh->result_bytes) {
        if (sh->h_result) cudaFreeHost(sh->h_result);
        sh->h_result = nullptr;
        CK(cudaHostAlloc(&sh->h_result, rb, cudaHostAllocPortable | cudaHostAllocMapped));
        sh->result_bytes = rb;
    }
    return YRB_OK;
}

// any shard's stream in an error state?  (called while waiting for `done`)
int poll_errors(yrb_sharded* sh) {
    for (int s = 0; s < sh->n; ++s) {
        if (cudaSetDevice(sh->devices[s]) != cudaSuccess) continue;
        const cudaError_t e = cudaStreamQuery(sh->shard[s]->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady)
            return fail(YRB_ERR_CUDA, "shard %d (device %d): %s", s, sh->devices[s], cudaGetErrorString(e));
    }
    return YRB_OK;
}

int sharded_search(yrb_sharded* sh, const float* queries, int nq, int k, const yrb_where* w, const yrb_where* const* per_query,
                   const uint32_t* mask, int64_t* out_ids, float* out_scores, int32_t* out_counts, float min_score = -INFINITY) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (nq < 1 || !queries) return fail(YRB_ERR_INVALID, "need at least one query");
    if (k < 1) return fail(YRB_ERR_INVALID, "k must be >= 1 (got %d)", k);
    if (!out_ids || !out_scores) return fail(YRB_ERR_INVALID, "output buffers are NULL");
    Nvtx nvtx_("yrb_sharded_search");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    int64_t live = 0;
    for (int s = 0; s < sh->n; ++s) live += sh->shard[s]->rows - sh->shard[s]->n_dead;
    auto fill_empty = [&] {
        for (int64_t i = 0; i < (int64_t)nq * k; ++i) {
            out_ids[i] = -1;
            out_scores[i] = -INFINITY;
        }
        if (out_counts)
            for (int q = 0; q < nq; ++q) out_counts[q] = 0;
    };
    if (live == 0) {
        fill_empty();
        return YRB_OK;
    }
    const int ke = (int)std::min<int64_t>(k, sh->rows);
    if (ke > YRB_FUSED_K_MAX && (size_t)sh->n * ke * 8 > 200 * 1024)
        return fail(YRB_ERR_UNSUPPORTED, "k=%d over %d shards exceeds the cross-shard merge's staging (n_shards * k <= 25600)", ke, sh->n);
    int rc = ensure_buffers(sh, nq, ke);
    if (rc) return rc;
    memcpy(sh->h_q, queries, (size_t)nq * sh->dim * 4);
    SearchJob& j = sh->job;
    j = SearchJob{};
    j.nq = nq;
    j.ke = ke;
    j.w = w;
    j.per_query = per_query;
    j.mask = mask;
    j.min_score = min_score;
    for (int s = 0; s < sh->n; ++s)
        if (sh->shard[s]->rows > 0) {
            j.active_mask |= 1u << s;
            j.n_active++;
        }
    *reinterpret_cast<volatile unsigned int*>(sh->h_done) = 0u;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    run_job(sh, JOB_SEARCH);
    sh->searches++;
    for (int s = 0; s < sh->n; ++s)
        if (sh->job_rc[s]) rc = sh->job_rc[s], set_error("shard " + std::to_string(s) + ": " + sh->job_err[s]);
    if (!rc) {
        // the kernels write the merged result into h_result and count finished queries in h_done
        const auto t0 = std::chrono::steady_clock::now();
        volatile unsigned int* done = sh->h_done;
        uint64_t spins = 0;
        while (*done < (unsigned int)nq) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((++spins & 0xffff) == 0) {
                const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (el > 0.05 && (rc = poll_errors(sh))) break;
                if (el > 60.0) {
                    rc = fail(YRB_ERR_CUDA, "sharded search timed out after 60 s (%u of %d queries finished)", *done, nq);
                    break;
                }
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (rc) {
        // drain every shard and rearm the tickets so that the next search starts clean
        for (int s = 0; s < sh->n; ++s) {
            cudaSetDevice(sh->devices[s]);
            cudaStreamSynchronize(sh->shard[s]->stream);
        }
        cudaSetDevice(sh->devices[sh->root]);
        cudaMemset(sh->d_tickets, 0, (size_t)sh->tickets_cap * 4);
        cudaGetLastError();
        return rc;
    }
    const int64_t* h_ids = reinterpret_cast<const int64_t*>(sh->h_result);
    const float* h_scores = reinterpret_cast<const float*>(sh->h_result + (size_t)nq * ke * 8);
    const int32_t* h_counts = reinterpret_cast<const int32_t*>(sh->h_result + (size_t)nq * ke * 12);
    for (int q = 0; q < nq; ++q) {
        for (int i = 0; i < k; ++i) {
            const bool ok = i < ke;
            out_ids[(int64_t)q * k + i] = ok ? h_ids[(int64_t)q * ke + i] : -1;
            out_scores[(int64_t)q * k + i] = ok ? h_scores[(int64_t)q * ke + i] : -INFINITY;
        }
        if (out_counts) out_counts[q] = h_counts[q];
    }
    return YRB_OK;
}

}  // namespace

extern "C" {

// ---- the block-cyclic row map, stateless (also what tests/test_shard_map.py checks on the CPU)
static bool shard_args_ok(int n_shards, int block_rows) {
    return n_shards >= 1 && n_shards <= yrb::XS_MAX_SHARDS && block_rows >= 64 && (block_rows & (block_rows - 1)) == 0;
}
int yrb_shard_locate(int n_shards, int block_rows, int64_t global_row, int* out_shard, int64_t* out_local) {
    if (!shard_args_ok(n_shards, block_rows) || global_row < 0 || !out_shard || !out_local) return fail(YRB_ERR_INVALID, "bad shard map arguments");
    const int64_t b = global_row / block_rows;
    *out_shard = (int)(b % n_shards);
    *out_local = (b / n_shards) * block_rows + global_row % block_rows;
    return YRB_OK;
}
int yrb_shard_global(int n_shards, int block_rows, int shard, int64_t local_row, int64_t* out_global) {
    if (!shard_args_ok(n_shards, block_rows) || shard < 0 || shard >= n_shards || local_row < 0 || !out_global)
        return fail(YRB_ERR_INVALID, "bad shard map arguments");
    int shift = 0;
    while ((1 << shift) < block_rows) ++shift;
    *out_global = yrb::xs_global_row(n_shards, shift, shard, (uint32_t)local_row);   // the function the merge kernels use
    return YRB_OK;
}
int yrb_shard_rows(int n_shards, int block_rows, int64_t total_rows, int shard, int64_t* out_rows) {
    if (!shard_args_ok(n_shards, block_rows) || shard < 0 || shard >= n_shards || total_rows < 0 || !out_rows)
        return fail(YRB_ERR_INVALID, "bad shard map arguments");
    const int64_t nb = total_rows / block_rows, rem = total_rows % block_rows;
    int64_t r = (nb / n_shards + (shard < nb % n_shards ? 1 : 0)) * block_rows;
    if (shard == nb % n_shards) r += rem;
    *out_rows = r;
    return YRB_OK;
}

int yrb_sharded_create(yrb_sharded** out, const int* devices, int n_devices, int dim, int metric, int storage_dtype,
                       int64_t reserve_rows, int block_rows_arg) {
    if (!out) return fail(YRB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > yrb::XS_MAX_SHARDS)
        return fail(YRB_ERR_INVALID, "need 1..%d devices (got %d)", yrb::XS_MAX_SHARDS, n_devices);
    // a device may be listed more than once (several shards on one GPU): that is how the cross-shard merge is
    // exercised on a one-GPU box (tests); it buys nothing in production
    int64_t B = block_rows_arg > 0 ? block_rows_arg : 16384;
    if (B < 64 || (B & (B - 1))) return fail(YRB_ERR_INVALID, "block_rows must be a power of two >= 64 (got %lld)", (long long)B);
    int shift = 0;
    while ((int64_t(1) << shift) < B) ++shift;
    DevGuard dev_guard_;
    yrb_sharded* sh = new (std::nothrow) yrb_sharded();
    if (!sh) return fail(YRB_ERR_NOMEM, "host allocation failed");
    sh->n = n_devices;
    sh->dim = dim;
    sh->metric = metric;
    sh->dtype = storage_dtype;
    sh->block_shift = shift;
    auto bail = [&](int code) {
        const std::string keep = last_error();
        yrb_sharded_destroy(sh);
        set_error(keep);
        return code;
    };
    for (int s = 0; s < n_devices; ++s) {
        sh->devices[s] = devices[s];
        const int rc = yrb_index_create(&sh->shard[s], devices[s], dim, metric, storage_dtype,
                                        (reserve_rows + n_devices - 1) / n_devices + (reserve_rows ? B : 0));
        if (rc) return bail(rc);
        // the finishing kernel of every search runs beside nothing else; keep the full grid
    }
    sh->ld = sh->shard[0]->ld;
    // peer access: every shard stores into / reads from the root's gather buffer and tickets
    for (int s = 1; s < n_devices; ++s) {
        if (devices[s] == devices[sh->root]) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[s], devices[sh->root]) != cudaSuccess || !can)
            return bail(fail(YRB_ERR_UNSUPPORTED, "device %d cannot access device %d directly (peer access is required for the "
                                                   "cross-shard merge)", devices[s], devices[sh->root]));
        cudaSetDevice(devices[s]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[sh->root], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return bail(fail(YRB_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[s], devices[sh->root], cudaGetErrorString(e)));
        cudaGetLastError();
        cudaSetDevice(devices[sh->root]);  // and back, for device-side appends that arrive through the root
        e = cudaDeviceEnablePeerAccess(devices[s], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        cudaGetLastError();
    }
    if (cudaHostAlloc(reinterpret_cast<void**>(&sh->h_done), 64, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
        return bail(fail(YRB_ERR_NOMEM, "pinned allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    *sh->h_done = 0u;
    for (int s = 0; s < n_devices; ++s) sh->workers[s] = std::thread(worker_main, sh, s);
    *out = sh;
    return YRB_OK;
}

int yrb_sharded_destroy(yrb_sharded* sh) {
    if (!sh) return YRB_OK;
    DevGuard dev_guard_;
    bool have_workers = false;
    for (int s = 0; s < sh->n; ++s) have_workers |= sh->workers[s].joinable();
    if (have_workers) {
        run_job(sh, JOB_QUIT);
        for (int s = 0; s < sh->n; ++s)
            if (sh->workers[s].joinable()) sh->workers[s].join();
    }
    for (int s = 0; s < sh->n; ++s) {
        if (sh->d_peer_stage[s]) {
            cudaSetDevice(sh->devices[s]);
            cudaFree(sh->d_peer_stage[s]);
        }
        if (sh->shard[s]) yrb_index_destroy(sh->shard[s]);
    }
    cudaSetDevice(sh->devices[sh->root]);
    if (sh->d_gather) cudaFree(sh->d_gather);
    if (sh->d_tickets) cudaFree(sh->d_tickets);
    if (sh->h_q) cudaFreeHost(sh->h_q);
    if (sh->h_result) cudaFreeHost(sh->h_result);
    if (sh->h_done) cudaFreeHost(sh->h_done);
    cudaGetLastError();
    delete sh;
    return YRB_OK;
}

int yrb_sharded_count(const yrb_sharded* sh, int64_t* out_rows, int64_t* out_live) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    int64_t dead = 0;
    for (int s = 0; s < sh->n; ++s) dead += sh->shard[s]->n_dead;
    if (out_rows) *out_rows = sh->rows;
    if (out_live) *out_live = sh->rows - dead;
    return YRB_OK;
}

int yrb_sharded_info(const yrb_sharded* sh, int* out_dim, int* out_ld, int* out_metric, int* out_dtype, int* out_n_devices,
                     int* out_block_rows, int64_t* out_capacity) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (out_dim) *out_dim = sh->dim;
    if (out_ld) *out_ld = sh->ld;
    if (out_metric) *out_metric = sh->metric;
    if (out_dtype) *out_dtype = sh->dtype;
    if (out_n_devices) *out_n_devices = sh->n;
    if (out_block_rows) *out_block_rows = (int)block_rows(sh);
    if (out_capacity) {
        *out_capacity = 0;
        for (int s = 0; s < sh->n; ++s) *out_capacity += sh->shard[s]->capacity;
    }
    return YRB_OK;
}

int yrb_sharded_shard(const yrb_sharded* sh, int s, yrb_index** out_index, int* out_device) {
    if (!sh || s < 0 || s >= sh->n) return fail(YRB_ERR_INVALID, "shard index out of range");
    if (out_index) *out_index = sh->shard[s];
    if (out_device) *out_device = sh->devices[s];
    return YRB_OK;
}

int yrb_sharded_append_host_f32(yrb_sharded* sh, const float* rows, int64_t n) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && !rows)) return fail(YRB_ERR_INVALID, "bad rows/n");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    const int64_t g0 = sh->rows;
    int rc = for_pieces(sh, g0, n, [&](int s, int64_t local, int64_t gb, int64_t cnt) {
        if (sh->shard[s]->rows != local) return fail(YRB_ERR_INVALID, "shard %d holds %lld rows, expected %lld", s, (long long)sh->shard[s]->rows, (long long)local);
        return yrb_index_append_host_f32(sh->shard[s], rows + (gb - g0) * sh->dim, cnt);
    });
    if (rc) {  // all or nothing: drop the pieces that did land
        for (int s = 0; s < sh->n; ++s) yrb_index_truncate(sh->shard[s], std::min(sh->shard[s]->rows, rows_on_shard(sh, g0, s)));
        return rc;
    }
    sh->rows += n;
    return YRB_OK;
}

int yrb_sharded_append_device_f32(yrb_sharded* sh, const float* dev_rows, int64_t n, int src_device) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && !dev_rows)) return fail(YRB_ERR_INVALID, "bad rows/n");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    const int64_t g0 = sh->rows;
    int rc = for_pieces(sh, g0, n, [&](int s, int64_t local, int64_t gb, int64_t cnt) {
        yrb_index* ix = sh->shard[s];
        if (ix->rows != local) return fail(YRB_ERR_INVALID, "shard %d holds %lld rows, expected %lld", s, (long long)ix->rows, (long long)local);
        const float* src = dev_rows + (gb - g0) * sh->dim;
        if (sh->devices[s] == src_device) return yrb_index_append_device_f32(ix, src, cnt, nullptr);
        CK(cudaSetDevice(sh->devices[s]));
        const size_t need = (size_t)cnt * sh->dim;
        if (need > sh->peer_stage_floats[s]) {
            if (sh->d_peer_stage[s]) cudaFree(sh->d_peer_stage[s]);
            sh->d_peer_stage[s] = nullptr;
            sh->peer_stage_floats[s] = 0;
            CK(cudaMalloc(&sh->d_peer_stage[s], need * 4));
            sh->peer_stage_floats[s] = need;
        }
        // on the shard's own stream: the ingest kernel that follows is ordered behind the copy (a plain
        // cudaMemcpyPeer runs on the legacy stream, which a non-blocking stream does not wait for)
        CK(cudaMemcpyPeerAsync(sh->d_peer_stage[s], sh->devices[s], src, src_device, need * 4, ix->stream));
        return yrb_index_append_device_f32(ix, sh->d_peer_stage[s], cnt, nullptr);
    });
    if (rc) {
        for (int s = 0; s < sh->n; ++s) yrb_index_truncate(sh->shard[s], std::min(sh->shard[s]->rows, rows_on_shard(sh, g0, s)));
        return rc;
    }
    sh->rows += n;
    return YRB_OK;
}

int yrb_sharded_read_rows(yrb_sharded* sh, const int64_t* row_ids, int64_t n, float* out_rows) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && (!row_ids || !out_rows))) return fail(YRB_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    // runs of consecutive ids inside one block become one read on their shard
    int64_t i = 0;
    while (i < n) {
        const int64_t r0 = row_ids[i];
        if (r0 < 0 || r0 >= sh->rows) return fail(YRB_ERR_INVALID, "row id %lld out of range", (long long)r0);
        int64_t j = i;
        while (j + 1 < n && row_ids[j + 1] == row_ids[j] + 1 && ((row_ids[j + 1] >> sh->block_shift) == (r0 >> sh->block_shift))) ++j;
        int s;
        int64_t local;
        locate(sh, r0, &s, &local);
        std::vector<int64_t> ids((size_t)(j - i + 1));
        for (size_t t = 0; t < ids.size(); ++t) ids[t] = local + (int64_t)t;
        const int rc = yrb_index_read_rows(sh->shard[s], ids.data(), (int64_t)ids.size(), out_rows + i * sh->dim);
        if (rc) return rc;
        i = j + 1;
    }
    return YRB_OK;
}

int yrb_sharded_read_raw(yrb_sharded* sh, int64_t row_begin, int64_t n, void* out_rows, float* out_sqnorm) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || row_begin < 0 || row_begin + n > sh->rows) return fail(YRB_ERR_INVALID, "row range out of bounds");
    if (n == 0) return YRB_OK;
    if (!out_rows || !out_sqnorm) return fail(YRB_ERR_INVALID, "output buffers are NULL");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    const size_t rb = (size_t)sh->ld * yrb::elem_size(sh->dtype);
    return for_pieces(sh, row_begin, n, [&](int s, int64_t local, int64_t gb, int64_t cnt) {
        return yrb_index_read_raw(sh->shard[s], local, cnt, static_cast<char*>(out_rows) + (size_t)(gb - row_begin) * rb,
                                  out_sqnorm + (gb - row_begin));
    });
}

int yrb_sharded_append_raw(yrb_sharded* sh, const void* rows, const float* sqnorm, int64_t n) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && (!rows || !sqnorm))) return fail(YRB_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    const int64_t g0 = sh->rows;
    const size_t rb = (size_t)sh->ld * yrb::elem_size(sh->dtype);
    int rc = for_pieces(sh, g0, n, [&](int s, int64_t local, int64_t gb, int64_t cnt) {
        if (sh->shard[s]->rows != local) return fail(YRB_ERR_INVALID, "shard %d holds %lld rows, expected %lld", s, (long long)sh->shard[s]->rows, (long long)local);
        return yrb_index_append_raw(sh->shard[s], static_cast<const char*>(rows) + (size_t)(gb - g0) * rb, sqnorm + (gb - g0), cnt);
    });
    if (rc) {
        for (int s = 0; s < sh->n; ++s) yrb_index_truncate(sh->shard[s], std::min(sh->shard[s]->rows, rows_on_shard(sh, g0, s)));
        return rc;
    }
    sh->rows += n;
    return YRB_OK;
}

int yrb_sharded_set_live(yrb_sharded* sh, const int64_t* row_ids, int64_t n, int live) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && !row_ids)) return fail(YRB_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    std::vector<int64_t> per[yrb::XS_MAX_SHARDS];
    for (int64_t i = 0; i < n; ++i) {
        if (row_ids[i] < 0 || row_ids[i] >= sh->rows) return fail(YRB_ERR_INVALID, "row id %lld out of range", (long long)row_ids[i]);
        int s;
        int64_t local;
        locate(sh, row_ids[i], &s, &local);
        per[s].push_back(local);
    }
    for (int s = 0; s < sh->n; ++s) {
        if (per[s].empty()) continue;
        const int rc = yrb_index_set_live(sh->shard[s], per[s].data(), (int64_t)per[s].size(), live);
        if (rc) return rc;
    }
    return YRB_OK;
}

int yrb_sharded_truncate(yrb_sharded* sh, int64_t rows) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (rows < 0 || rows > sh->rows) return fail(YRB_ERR_INVALID, "truncate to %lld rows: index holds %lld", (long long)rows, (long long)sh->rows);
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    for (int s = 0; s < sh->n; ++s) {
        const int rc = yrb_index_truncate(sh->shard[s], rows_on_shard(sh, rows, s));
        if (rc) return rc;
    }
    sh->rows = rows;
    return YRB_OK;
}

int yrb_sharded_clear(yrb_sharded* sh) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    for (int s = 0; s < sh->n; ++s) {
        const int rc = yrb_index_clear(sh->shard[s]);
        if (rc) return rc;
    }
    sh->rows = 0;
    return YRB_OK;
}

int yrb_sharded_column_write(yrb_sharded* sh, int col, int col_type, int64_t row_begin, int64_t n, const void* values,
                             const uint8_t* present) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (col < 0) return fail(YRB_ERR_INVALID, "column id must be >= 0");
    if (col_type < YRB_COL_I64 || col_type > YRB_COL_BOOL) return fail(YRB_ERR_INVALID, "bad column type %d", col_type);
    if (n < 0 || row_begin < 0 || (n > 0 && (!values || !present))) return fail(YRB_ERR_INVALID, "bad arguments");
    if (row_begin + n > sh->rows) return fail(YRB_ERR_INVALID, "column rows beyond appended rows");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    // the column must exist on EVERY shard: the same compiled where program runs on all of them
    for (int s = 0; s < sh->n; ++s) {
        yrb_index* ix = sh->shard[s];
        std::lock_guard<std::mutex> gi(ix->mu);
        int rc = set_dev(ix);
        if (!rc) rc = ensure_column(ix, col, col_type);
        if (rc) return rc;
    }
    const size_t w = (col_type == YRB_COL_I64 || col_type == YRB_COL_F64) ? 8 : (col_type == YRB_COL_CODE ? 4 : 1);
    return for_pieces(sh, row_begin, n, [&](int s, int64_t local, int64_t gb, int64_t cnt) {
        return yrb_index_column_write(sh->shard[s], col, col_type, local, cnt, static_cast<const char*>(values) + (size_t)(gb - row_begin) * w,
                                      present + (gb - row_begin));
    });
}

int yrb_sharded_where(yrb_sharded* sh, const yrb_where* w, uint32_t* out_mask, int64_t* out_pass) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    std::lock_guard<std::mutex> g(sh->mu);
    DevGuard dev_guard_;
    int64_t pass = 0;
    const int64_t wpb = block_rows(sh) / 32, gw_total = (sh->rows + 31) / 32;
    if (out_mask) memset(out_mask, 0, (size_t)gw_total * 4);
    for (int s = 0; s < sh->n; ++s) {
        yrb_index* ix = sh->shard[s];
        if (ix->rows == 0) continue;
        const int64_t lw = (ix->rows + 31) / 32;
        std::vector<uint32_t> local(out_mask ? (size_t)lw : 0);
        int64_t p = 0;
        const int rc = yrb_index_where(ix, w, out_mask ? local.data() : nullptr, &p);
        if (rc) return rc;
        pass += p;
        if (out_mask)
            for (int64_t lwi = 0; lwi < lw; ++lwi) {
                const int64_t gw = ((lwi / wpb) * sh->n + s) * wpb + (lwi % wpb);
                if (gw < gw_total) out_mask[gw] = local[lwi];
            }
    }
    if (out_pass) *out_pass = pass;
    return YRB_OK;
}

int yrb_sharded_search(yrb_sharded* sh, const float* queries, int nq, int k, const yrb_where* w, const uint32_t* mask,
                       int64_t* out_ids, float* out_scores, int32_t* out_counts) {
    return sharded_search(sh, queries, nq, k, w, nullptr, mask, out_ids, out_scores, out_counts);
}

int yrb_sharded_search_multi(yrb_sharded* sh, const float* queries, int nq, int k, const yrb_where* const* wheres,
                             int64_t* out_ids, float* out_scores, int32_t* out_counts) {
    if (!wheres) return fail(YRB_ERR_INVALID, "wheres is NULL (use yrb_sharded_search for a shared filter)");
    return sharded_search(sh, queries, nq, k, nullptr, wheres, nullptr, out_ids, out_scores, out_counts);
}

int yrb_sharded_search_ex(yrb_sharded* sh, const float* queries, int nq, int k, const yrb_where* w, const yrb_where* const* wheres,
                          const uint32_t* mask, const yrb_search_opts* opts, int64_t* out_ids, float* out_scores, int32_t* out_counts) {
    if (w && wheres) return fail(YRB_ERR_INVALID, "pass a shared filter or per-query filters, not both");
    if (wheres && mask) return fail(YRB_ERR_INVALID, "a host bitmask cannot be combined with per-query filters");
    const float ms = opts ? opts->min_score : -INFINITY;
    if (ms != ms) return fail(YRB_ERR_INVALID, "min_score is NaN");
    return sharded_search(sh, queries, nq, k, w, wheres, mask, out_ids, out_scores, out_counts, ms);
}

int yrb_sharded_stats(const yrb_sharded* sh, int64_t* out_kernel_launches, int64_t* out_searches) {
    if (!sh) return fail(YRB_ERR_INVALID, "index is NULL");
    if (out_kernel_launches) {
        *out_kernel_launches = 0;
        for (int s = 0; s < sh->n; ++s) *out_kernel_launches += sh->shard[s]->launches;
    }
    if (out_searches) *out_searches = sh->searches;
    return YRB_OK;
}

}  // extern "C"
