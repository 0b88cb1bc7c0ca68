// index_internal.h — the index object and the internal entry points shared by capi.cu (one index on one GPU)
// and sharded.cu (one collection spread over the GPUs of a box).  Internal to libyrb200.so.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/yrb200.h"
#include "k2_batched.h"
#include "kernels.h"
#include "xshard.cuh"

namespace yrbi {

// NVTX range around an ABI entry or a phase of a search (nsys timelines of the sharded path read as
// entry → filter → scan → merge).  Header-only NVTX v3: no cost unless a profiler is attached.
struct Nvtx {
    explicit Nvtx(const char* name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
    Nvtx(const Nvtx&) = delete;
    Nvtx& operator=(const Nvtx&) = delete;
};

// Restores the calling thread's current CUDA device when an entry point returns: the library switches to the index's
// device (and, for a sharded collection, walks over several), and a caller that shares the thread with another CUDA
// user (torch in the serving process) must find its device unchanged.
struct DevGuard {
    int prev = -1;
    DevGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            prev = -1;
            cudaGetLastError();
        }
    }
    ~DevGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
    DevGuard(const DevGuard&) = delete;
    DevGuard& operator=(const DevGuard&) = delete;
};

// sets the calling thread's error message (yrb_last_error) and returns `code`
int fail(int code, const char* fmt, ...);
const std::string& last_error();
void set_error(const std::string& m);

struct Column {
    int type = -1;
    void* values = nullptr;       // device, capacity rows
    uint32_t* present = nullptr;  // device bitmask, capacity words
    std::vector<uint32_t> present_host;
};

inline int64_t mask_words(int64_t rows) { return (((rows + 31) / 32) + 1) & ~int64_t(1); }

// What every index on one GPU shares (an agent process holds one index per collection / per user
// memory, and round 1 gave each its own stream and 86 MB of K2 candidate buffers): ONE stream — so the searches of a
// device are ordered and may share scratch — and ONE set of K2 buffers.  Enqueue sequences that use the shared K2
// state hold `mu`.  Lives as long as the process.  YRB_PRIVATE_STREAMS=1 restores one stream + K2 state per index.
struct DevicePool {
    cudaStream_t stream = nullptr;
    yrb::K2State* k2 = nullptr;
    std::mutex mu;
};
DevicePool* device_pool(int device);   // NULL if the stream cannot be created

}  // namespace yrbi

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return yrbi::fail(e_ == cudaErrorMemoryAllocation ? YRB_ERR_NOMEM : YRB_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                              cudaGetErrorString(e_), __FILE__, __LINE__);                            \
    } while (0)

struct yrb_index {
    int device = 0, dim = 0, ld = 0, metric = 0, dtype = 0, sm_count = 148;
    int64_t rows = 0, capacity = 0, n_dead = 0;
    void* d_rows = nullptr;
    float* d_sqnorm = nullptr;
    uint32_t* d_live = nullptr;  // mask_words(capacity)
    uint32_t* d_mask = nullptr;  // filter scratch, same size
    uint32_t* d_usermask = nullptr;  // device copy of a caller-supplied host bitmask, same size
    std::vector<uint32_t> h_live;
    std::map<int, yrbi::Column> cols;
    cudaStream_t stream = nullptr;
    // search scratch
    int nq_cap = 0, k_cap = 0;
    float* d_qf32 = nullptr;
    void* d_q = nullptr;
    float* d_qsq = nullptr;
    uint64_t* d_parts = nullptr;
    uint64_t* d_keys = nullptr;
    unsigned char* d_result = nullptr;  // [ids nq*k i64 | scores nq*k f32 | counts nq i32], one D2H
    unsigned char* d_result_host = nullptr;  // device alias of h_result (pinned, mapped): small results are written there
    size_t result_bytes = 0;
    int64_t* d_ids = nullptr;           // views into d_result for the current (nq, k)
    float* d_scores = nullptr;
    int32_t* d_counts = nullptr;
    unsigned int* d_ticket = nullptr;   // K1's last-CTA-done counter
    unsigned long long* d_k1trace = nullptr;  // YRB_K1_TRACE=1: per-CTA phase stamps of the last K1 launch
    // K8 compaction scratch (grow-only)
    uint32_t* d_cp_blocks = nullptr;
    size_t cp_blocks_cap = 0;
    void* d_cp_rows = nullptr;
    float* d_cp_sqnorm = nullptr;
    uint32_t* d_cp_map = nullptr;
    int64_t cp_rows_cap = 0;
    // filter caches (agents repeat the same knowledge-base / source filter).  A key is the hash of
    // the compiled where program plus `epoch`, which every mutation of rows, tombstones or columns bumps; 0 = no key.
    uint64_t epoch = 1;
    uint64_t cur_mask_key = 0;   // key of the mask the current search scans with (set by resolve_mask)
    uint64_t dmask_key = 0;      // d_mask holds the evaluated mask of this key (K4 skipped on a hit)
    uint64_t cp_key = 0;         // d_cp_rows / d_cp_map / d_cp_sqnorm hold the compaction of this key (K8 skipped)
    int64_t cp_pass = 0;
    bool cp_rowmap = false;
    int64_t cache_hits_k4 = 0, cache_hits_k8 = 0;
    unsigned long long* h_pass = nullptr;  // pinned
    uint64_t* d_rowkeys = nullptr;  // K6: one key per row, allocated on first use
    int64_t rowkeys_cap = 0;
    void* d_select = nullptr;
    size_t select_bytes = 0;
    yrb::WhereProgDev* d_prog = nullptr;
    yrb::WhereProgDev* h_prog = nullptr;  // pinned
    yrb::WhereProgDev* d_progs = nullptr;  // per-query filters of a batch
    yrb::WhereProgDev* h_progs = nullptr;
    size_t progs_cap = 0;
    uint32_t* d_qmasks = nullptr;          // [nq][mask_words]
    size_t qmasks_bytes = 0;
    unsigned long long* d_pass = nullptr;
    // pinned staging
    float* h_q = nullptr;
    unsigned char* h_result = nullptr;
    void* h_stage = nullptr;
    size_t stage_bytes = 0;
    float* d_append = nullptr;  // device staging of append_host_f32 (grow-only; round 1 allocated it per call)
    size_t append_bytes = 0;
    yrb::K2State* k2 = nullptr;
    yrbi::DevicePool* pool = nullptr;   // non-NULL: stream and k2 belong to the device pool
    int path = 0;
    int reserved_sms = 0;
    int64_t launches = 0;
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
    size_t prof_used = 0;
    double prof_ms = 0.0;
    int64_t prof_n = 0;
    std::mutex mu;
};


namespace yrbi {

int set_dev(const yrb_index* ix);
int ensure_scratch(yrb_index* ix, int nq, int k);
// evaluates w (and/or ANDs a device mask) into ix->d_mask; *out = the mask to scan with (NULL = all rows)
int resolve_mask(yrb_index* ix, const yrb_where* w, const uint32_t* dev_extra, const uint32_t** out, cudaStream_t st, bool count);
int resolve_masks_multi(yrb_index* ix, const yrb_where* const* wheres, int nq, const uint32_t* dev_extra, const uint32_t** out,
                        int64_t* out_stride, cudaStream_t st);
// uploads a caller-supplied host bitmask (ceil(rows/32) words) into ix->d_usermask; *out = its device copy
int upload_user_mask(yrb_index* ix, const uint32_t* mask, const uint32_t** out, cudaStream_t st);
// raw fp32 queries [nq, dim] on the device → nq*k keys (and ids/scores/counts when given).  With `xs` the kernel that
// finishes a query hands its k keys to the cross-shard merge instead (xshard.cuh) and ids/scores/counts are ignored.
int scan_select(yrb_index* ix, const float* dev_q, int nq, int k, const uint32_t* mask, int64_t mask_q_stride, uint64_t* out_keys,
                int64_t* ids, float* scores, int32_t* counts, cudaStream_t st, const yrb::XShard* xs = nullptr,
                float min_score = -INFINITY);   // keep hits with score >= min_score
// makes column `col` exist (all rows absent) so that a where program naming it compiles on this shard too
int ensure_column(yrb_index* ix, int col, int col_type);

}  // namespace yrbi
