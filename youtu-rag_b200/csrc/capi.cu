// capi.cu — the C ABI declared in include/yrb200.h: one `yrb_index` = one collection's rows,
// tombstones and metadata columns resident on one B200, plus the search entry points that chain
// the kernels (K5 query prep → K4 filter → K1/K2/K6 scan+select → K3 merge → decode).
// The reference interface each entry stands in for is cited in the header.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "index_internal.h"

namespace yrbi {

namespace {
thread_local std::string g_err;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
const std::string& last_error() { return g_err; }
void set_error(const std::string& m) { g_err = m; }

}  // namespace yrbi
namespace yrb {
uint64_t score_floor_key(float min_score) {
    if (!(min_score > -INFINITY)) return 0ull;   // -inf or NaN: no threshold
    // monotone bit pattern of the largest float below min_score, all row bits set: key > floor  <=>  score >= min_score
    const float lo = nextafterf(min_score, -INFINITY);
    uint32_t u;
    memcpy(&u, &lo, 4);
    const uint32_t m = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((uint64_t)m << 32) | 0xffffffffull;
}
}  // namespace yrb
namespace yrbi {
int col_width(int t) { return t == YRB_COL_I64 || t == YRB_COL_F64 ? 8 : (t == YRB_COL_CODE ? 4 : 1); }

DevicePool* device_pool(int device) {
    static std::mutex reg_mu;
    static std::map<int, DevicePool*> reg;
    std::lock_guard<std::mutex> g(reg_mu);
    auto it = reg.find(device);
    if (it != reg.end()) return it->second;
    DevicePool* p = new DevicePool();
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        delete p;
        return nullptr;
    }
    p->k2 = yrb::k2_create();
    reg[device] = p;
    return p;
}

}  // namespace yrbi

using namespace yrbi;

namespace yrbi {

int set_dev(const yrb_index* ix) {
    CK(cudaSetDevice(ix->device));
    return YRB_OK;
}

template <typename T>
int regrow(T** p, size_t old_bytes, size_t new_bytes, bool zero_tail, cudaStream_t st) {
    T* np_ = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&np_), new_bytes ? new_bytes : 1));
    if (*p && old_bytes) CK(cudaMemcpyAsync(np_, *p, std::min(old_bytes, new_bytes), cudaMemcpyDeviceToDevice, st));
    if (zero_tail && new_bytes > old_bytes)
        CK(cudaMemsetAsync(reinterpret_cast<char*>(np_) + old_bytes, 0, new_bytes - old_bytes, st));
    CK(cudaStreamSynchronize(st));
    if (*p) CK(cudaFree(*p));
    *p = np_;
    return YRB_OK;
}

int ensure_capacity(yrb_index* ix, int64_t want) {
    if (want <= ix->capacity) return YRB_OK;
    int64_t cap = std::max<int64_t>(want, ix->capacity + ix->capacity / 2);
    cap = std::max<int64_t>(cap, 256);   // tiny collections (per-user memories) stay tiny: 512 KiB of rows at 1024-d
    cap = (cap + 255) / 256 * 256;
    const size_t es = yrb::elem_size(ix->dtype);
    int rc;
    if ((rc = regrow(reinterpret_cast<char**>(&ix->d_rows), (size_t)ix->rows * ix->ld * es, (size_t)cap * ix->ld * es,
                     false, ix->stream)))
        return rc;
    if ((rc = regrow(&ix->d_sqnorm, (size_t)ix->rows * 4, (size_t)cap * 4, false, ix->stream))) return rc;
    const size_t ow = (size_t)mask_words(ix->capacity) * 4, nw = (size_t)mask_words(cap) * 4;
    if ((rc = regrow(&ix->d_live, ix->capacity ? ow : 0, nw, true, ix->stream))) return rc;
    if ((rc = regrow(&ix->d_mask, 0, nw, true, ix->stream))) return rc;
    ix->dmask_key = 0;   // the cached filter mask lived in the old buffer
    if ((rc = regrow(&ix->d_usermask, 0, nw, true, ix->stream))) return rc;
    ix->h_live.resize(mask_words(cap), 0u);
    for (auto& kv : ix->cols) {
        Column& c = kv.second;
        const size_t w = col_width(c.type);
        if ((rc = regrow(reinterpret_cast<char**>(&c.values), (size_t)ix->capacity * w, (size_t)cap * w, true,
                         ix->stream)))
            return rc;
        if ((rc = regrow(&c.present, ix->capacity ? ow : 0, nw, true, ix->stream))) return rc;
        c.present_host.resize(mask_words(cap), 0u);
    }
    ix->capacity = cap;
    return YRB_OK;
}

int ensure_stage(yrb_index* ix, size_t bytes) {
    if (bytes <= ix->stage_bytes) return YRB_OK;
    if (ix->h_stage) CK(cudaFreeHost(ix->h_stage));
    ix->h_stage = nullptr;
    ix->stage_bytes = 0;
    CK(cudaMallocHost(&ix->h_stage, bytes));
    ix->stage_bytes = bytes;
    return YRB_OK;
}

#define FREE_DEV(p)            \
    do {                       \
        if (p) cudaFree(p);    \
        p = nullptr;           \
    } while (0)
#define FREE_HOST(p)            \
    do {                        \
        if (p) cudaFreeHost(p); \
        p = nullptr;            \
    } while (0)

void free_scratch(yrb_index* ix) {
    FREE_DEV(ix->d_qf32);
    FREE_DEV(ix->d_q);
    FREE_DEV(ix->d_qsq);
    FREE_DEV(ix->d_parts);
    FREE_DEV(ix->d_keys);
    FREE_DEV(ix->d_result);
    ix->d_ids = nullptr;
    ix->d_scores = nullptr;
    ix->d_counts = nullptr;
    FREE_HOST(ix->h_q);
    FREE_HOST(ix->h_result);
    ix->d_result_host = nullptr;
    ix->nq_cap = ix->k_cap = 0;
}

int ensure_scratch(yrb_index* ix, int nq, int k) {
    if (nq <= ix->nq_cap && k <= ix->k_cap) return YRB_OK;
    const int nqc = std::max(nq, ix->nq_cap), kc = std::max(k, ix->k_cap);
    CK(cudaStreamSynchronize(ix->stream));
    free_scratch(ix);
    const int parts = std::max(yrb::k1_parts(ix->sm_count), yrb::k2_parts(ix->sm_count));
    const size_t es = yrb::elem_size(ix->dtype);
    CK(cudaMalloc(&ix->d_qf32, (size_t)nqc * ix->dim * 4));
    // prepared queries, padded to whole 128-query blocks (zero rows) so K2's TMA boxes never leave the tensor
    CK(cudaMalloc(&ix->d_q, (size_t)((nqc + 127) / 128 * 128) * ix->ld * es));
    CK(cudaMalloc(&ix->d_qsq, (size_t)nqc * 4));
    CK(cudaMalloc(&ix->d_parts, (size_t)parts * nqc * kc * 8));
    CK(cudaMalloc(&ix->d_keys, (size_t)nqc * kc * 8));
    ix->result_bytes = (size_t)nqc * kc * 12 + (size_t)nqc * 4;
    CK(cudaMalloc(&ix->d_result, ix->result_bytes));
    CK(cudaMallocHost(&ix->h_q, (size_t)nqc * ix->dim * 4));
    CK(cudaMallocHost(&ix->h_result, ix->result_bytes));
    {
        void* alias = nullptr;  // pinned memory is device-accessible under unified addressing; no alias → D2H copies only
        if (cudaHostGetDevicePointer(&alias, ix->h_result, 0) != cudaSuccess) {
            cudaGetLastError();
            alias = nullptr;
        }
        ix->d_result_host = static_cast<unsigned char*>(alias);
    }
    ix->nq_cap = nqc;
    ix->k_cap = kc;
    return YRB_OK;
}

// carve d_result for this call's (nq, k): ids | scores | counts, contiguous → one D2H copy
// `base`: d_result, or the device alias of the pinned host buffer (results of a few hundred bytes are stored
// straight to host memory by the selecting kernel: no copy-engine round trip after the scan)
size_t result_views(yrb_index* ix, int nq, int k, unsigned char* base) {
    ix->d_ids = reinterpret_cast<int64_t*>(base);
    ix->d_scores = reinterpret_cast<float*>(base + (size_t)nq * k * 8);
    ix->d_counts = reinterpret_cast<int32_t*>(base + (size_t)nq * k * 12);
    return (size_t)nq * k * 12 + (size_t)nq * 4;
}
constexpr size_t ZERO_COPY_RESULT_MAX = 4096;

// upload the host mirror of a bitmask range [w0, w1)
int upload_words(uint32_t* dev, const std::vector<uint32_t>& host, int64_t w0, int64_t w1, cudaStream_t st) {
    if (w1 <= w0) return YRB_OK;
    CK(cudaMemcpyAsync(dev + w0, host.data() + w0, (size_t)(w1 - w0) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return YRB_OK;
}

int mark_appended(yrb_index* ix, int64_t r0, int64_t n) {
    for (int64_t r = r0; r < r0 + n; ++r) ix->h_live[r >> 5] |= (1u << (r & 31));
    return upload_words(ix->d_live, ix->h_live, r0 >> 5, ((r0 + n - 1) >> 5) + 1, ix->stream);
}

int build_prog(yrb_index* ix, const yrb_where* w) {
    if (!w) return fail(YRB_ERR_INVALID, "where is NULL");
    if (w->n_leaves < 0 || w->n_leaves > YRB_WHERE_MAX_LEAVES) return fail(YRB_ERR_INVALID, "where: too many leaves (%d > %d)", w->n_leaves, YRB_WHERE_MAX_LEAVES);
    if (w->n_operands < 0 || w->n_operands > YRB_WHERE_MAX_OPERANDS) return fail(YRB_ERR_INVALID, "where: too many operands (%d > %d)", w->n_operands, YRB_WHERE_MAX_OPERANDS);
    if (w->n_postfix < 0 || w->n_postfix > YRB_WHERE_MAX_TOKENS) return fail(YRB_ERR_INVALID, "where: too many tokens (%d > %d)", w->n_postfix, YRB_WHERE_MAX_TOKENS);
    yrb::WhereProgDev* p = ix->h_prog;
    memset(p, 0, sizeof *p);
    p->n_leaves = w->n_leaves;
    p->n_postfix = w->n_postfix;
    for (int i = 0; i < w->n_leaves; ++i) {
        const yrb_where_leaf& l = w->leaves[i];
        if (l.op < YRB_OP_EQ || l.op > YRB_OP_NIN) return fail(YRB_ERR_INVALID, "where leaf %d: bad op %d", i, l.op);
        if (l.operand_count < 1 || l.operand_begin < 0 || l.operand_begin + l.operand_count > w->n_operands)
            return fail(YRB_ERR_INVALID, "where leaf %d: operand range out of bounds", i);
        yrb::WhereLeafDev& d = p->leaves[i];
        d.op = l.op;
        d.operand_begin = l.operand_begin;
        d.operand_count = l.operand_count;
        d.col_type = -1;
        if (l.col >= 0) {
            auto it = ix->cols.find(l.col);
            if (it == ix->cols.end()) return fail(YRB_ERR_INVALID, "where leaf %d: unknown column %d", i, l.col);
            d.col_type = it->second.type;
            d.values = it->second.values;
            d.present = it->second.present;
        }
    }
    for (int i = 0; i < w->n_operands; ++i) p->operands[i] = w->operands[i];
    int depth = 0;
    for (int i = 0; i < w->n_postfix; ++i) {
        const int t = w->postfix[i];
        if (t >= 0) {
            if (t >= w->n_leaves) return fail(YRB_ERR_INVALID, "where token %d: leaf %d out of range", i, t);
            ++depth;
        } else if (t == YRB_TOK_NOT) {
            if (depth < 1) return fail(YRB_ERR_INVALID, "where: malformed postfix");
        } else if (t == YRB_TOK_AND || t == YRB_TOK_OR) {
            if (depth < 2) return fail(YRB_ERR_INVALID, "where: malformed postfix");
            --depth;
        } else {
            return fail(YRB_ERR_INVALID, "where token %d: bad token %d", i, t);
        }
        if (depth > 64) return fail(YRB_ERR_INVALID, "where: expression too deep");
        p->postfix[i] = t;
    }
    if (w->n_postfix > 0 && depth != 1) return fail(YRB_ERR_INVALID, "where: malformed postfix");
    return YRB_OK;
}

static uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

// evaluates w (and/or ANDs a device mask) into ix->d_mask; returns the mask pointer to scan with
int resolve_mask(yrb_index* ix, const yrb_where* w, const uint32_t* dev_extra, const uint32_t** out, cudaStream_t st,
                 bool count) {
    Nvtx nvtx_("k4_filter");
    const uint32_t* live = ix->n_dead > 0 ? ix->d_live : nullptr;
    ix->cur_mask_key = 0;
    if (w) {
        int rc = build_prog(ix, w);
        if (rc) return rc;
        // the same program over unchanged rows, tombstones and columns selects the same rows: reuse the mask
        uint64_t key = 0;
        if (!dev_extra && !count) {
            const yrb::WhereProgDev* p = ix->h_prog;
            key = fnv1a(&p->n_leaves, sizeof p->n_leaves);
            key = fnv1a(&p->n_postfix, sizeof p->n_postfix, key);
            key = fnv1a(p->leaves, sizeof(yrb::WhereLeafDev) * (size_t)p->n_leaves, key);
            key = fnv1a(p->operands, sizeof(int64_t) * (size_t)w->n_operands, key);
            key = fnv1a(p->postfix, sizeof(int32_t) * (size_t)p->n_postfix, key);
            key = fnv1a(&ix->epoch, sizeof ix->epoch, key);
            key |= 1ull;
        }
        if (key && key == ix->dmask_key) {
            ix->cache_hits_k4++;
        } else {
            CK(cudaMemcpyAsync(ix->d_prog, ix->h_prog, sizeof(yrb::WhereProgDev), cudaMemcpyHostToDevice, st));
            if (count) CK(cudaMemsetAsync(ix->d_pass, 0, 8, st));
            CK(yrb::launch_where(ix->d_prog, ix->rows, live, dev_extra, ix->d_mask, count ? ix->d_pass : nullptr, st));
            ix->launches++;
            ix->dmask_key = key;
        }
        ix->cur_mask_key = key;
        *out = ix->d_mask;
    } else if (dev_extra) {
        if (live) {
            CK(yrb::launch_mask_and(dev_extra, live, mask_words(ix->rows), ix->d_mask, st));
            ix->launches++;
            ix->dmask_key = 0;
            *out = ix->d_mask;
        } else {
            *out = dev_extra;
        }
    } else {
        *out = live;
    }
    return YRB_OK;
}

// per-query filters → ix->d_qmasks [nq][mask_words]; programs are deduplicated by pointer
int resolve_masks_multi(yrb_index* ix, const yrb_where* const* wheres, int nq, const uint32_t* dev_extra,
                        const uint32_t** out, int64_t* out_stride, cudaStream_t st) {
    const int64_t words = mask_words(ix->rows);
    const size_t need = (size_t)nq * words * 4;
    if (need > ix->qmasks_bytes) {
        CK(cudaStreamSynchronize(st));
        FREE_DEV(ix->d_qmasks);
        CK(cudaMalloc(&ix->d_qmasks, need));
        ix->qmasks_bytes = need;
    }
    if ((int)ix->progs_cap < nq) {
        CK(cudaStreamSynchronize(st));
        FREE_DEV(ix->d_progs);
        FREE_HOST(ix->h_progs);
        CK(cudaMalloc(&ix->d_progs, (size_t)nq * sizeof(yrb::WhereProgDev)));
        CK(cudaMallocHost(&ix->h_progs, (size_t)nq * sizeof(yrb::WhereProgDev)));
        ix->progs_cap = nq;
    }
    const uint32_t* live = ix->n_dead > 0 ? ix->d_live : nullptr;
    yrb::WhereProgDev* saved = ix->h_prog;
    yrb_where all = {nullptr, 0, nullptr, 0, nullptr, 0};
    for (int j = 0; j < nq; ++j) {
        uint32_t* dst = ix->d_qmasks + (size_t)j * words;
        int dup = -1;
        for (int i = 0; i < j && dup < 0; ++i)
            if (wheres[i] == wheres[j]) dup = i;
        if (dup >= 0) {
            CK(cudaMemcpyAsync(dst, ix->d_qmasks + (size_t)dup * words, (size_t)words * 4, cudaMemcpyDeviceToDevice, st));
            continue;
        }
        ix->h_prog = ix->h_progs + j;  // build_prog writes into ix->h_prog
        int rc = build_prog(ix, wheres[j] ? wheres[j] : &all);
        ix->h_prog = saved;
        if (rc) return rc;
        CK(cudaMemcpyAsync(ix->d_progs + j, ix->h_progs + j, sizeof(yrb::WhereProgDev), cudaMemcpyHostToDevice, st));
        CK(yrb::launch_where(ix->d_progs + j, ix->rows, live, dev_extra, dst, nullptr, st));
        ix->launches++;
    }
    *out = ix->d_qmasks;
    *out_stride = words;
    return YRB_OK;
}

int prof_flush(yrb_index* ix) {
    for (size_t i = 0; i + 1 < ix->prof_used; i += 2) {
        CK(cudaEventSynchronize(ix->prof_ev[i + 1]));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ix->prof_ev[i], ix->prof_ev[i + 1]));
        ix->prof_ms += ms;
        ix->prof_n++;
    }
    ix->prof_used = 0;
    return YRB_OK;
}

// reserves the next (start, stop) event pair for code that records them itself
int prof_pair(yrb_index* ix, cudaEvent_t* a, cudaEvent_t* b) {
    *a = *b = nullptr;
    if (!ix->prof) return YRB_OK;
    if (ix->prof_used + 2 > ix->prof_ev.size()) {
        if (ix->prof_ev.size() >= 8192) {
            int rc = prof_flush(ix);
            if (rc) return rc;
        } else {
            cudaEvent_t x, y;
            CK(cudaEventCreate(&x));
            CK(cudaEventCreate(&y));
            ix->prof_ev.push_back(x);
            ix->prof_ev.push_back(y);
        }
    }
    *a = ix->prof_ev[ix->prof_used];
    *b = ix->prof_ev[ix->prof_used + 1];
    ix->prof_used += 2;
    return YRB_OK;
}

int prof_mark(yrb_index* ix, cudaStream_t st) {
    if (!ix->prof) return YRB_OK;
    if (ix->prof_used == ix->prof_ev.size()) {
        if (ix->prof_ev.size() >= 8192) {
            int rc = prof_flush(ix);
            if (rc) return rc;
        } else {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            ix->prof_ev.push_back(a);
            ix->prof_ev.push_back(b);
        }
    }
    CK(cudaEventRecord(ix->prof_ev[ix->prof_used++], st));
    return YRB_OK;
}

// raw fp32 queries [nq, dim] on the device → nq*k keys (and, when `decode`, ix->d_ids/d_scores/d_counts).
// K1 prepares the query in its own prologue; K2 needs the prepared bf16 matrix (K5 launch).
int scan_select(yrb_index* ix, const float* dev_q, int nq, int k, const uint32_t* mask, int64_t mask_q_stride,
                uint64_t* out_keys, int64_t* ids, float* scores, int32_t* counts, cudaStream_t st, const yrb::XShard* xs,
                float min_score) {
    Nvtx nvtx_("scan_select");
    const uint64_t floor_key = yrb::score_floor_key(min_score);
    if (xs) ids = nullptr, scores = nullptr, counts = nullptr;  // the cross-shard merge writes the results
    const bool decode = ids != nullptr;
    const int sms = std::max(2, ix->sm_count - ix->reserved_sms) & ~1;  // even: K2 runs CTA clusters of 2
    int path = ix->path;
    if (path == 0) {
        if (k > YRB_FUSED_K_MAX) path = 3;
        else if (nq >= 2 && yrb::k2_supported(ix->dtype, ix->dim, k)) path = 2;  // K2 beats a K1 loop from 2 queries (scripts/crossover.py)
        else path = 1;
    }
    const bool pair = (path == 4);
    if (pair) path = 2;
    if (path == 2 && !yrb::k2_supported(ix->dtype, ix->dim, k))
        return fail(YRB_ERR_UNSUPPORTED, "K2 (tcgen05 batched) needs bf16 storage and k <= %d", YRB_FUSED_K_MAX);
    if ((path == 1 || path == 2) && k > YRB_FUSED_K_MAX)
        return fail(YRB_ERR_UNSUPPORTED, "fused selection handles k <= %d", YRB_FUSED_K_MAX);
    if (path == 2) {
        // the K2 buffers may be shared by every index of this device: one enqueue sequence at a time
        std::unique_lock<std::mutex> pool_lock;
        if (ix->pool) pool_lock = std::unique_lock<std::mutex>(ix->pool->mu);
        // K8: a filter shared by the batch that passes at most a quarter of the rows → gather the passing rows
        // and run the GEMM on the compact matrix (needs the pass count on the host: one stream synchronisation)
        const void* k2_rows = ix->d_rows;
        const float* k2_sqnorm = ix->d_sqnorm;
        int64_t k2_n = ix->rows;
        bool compacted = false, use_rowmap = false;
        const uint64_t mkey = ix->cur_mask_key;
        if (mask && mask_q_stride == 0 && ix->rows >= 65536 && mkey && mkey == ix->cp_key && ix->cp_pass >= k) {
            // this filter's rows were gathered by an earlier search and nothing changed since: no count, no
            // synchronisation, no gather
            ix->cache_hits_k8++;
            use_rowmap = ix->cp_rowmap;
            if (!use_rowmap) k2_rows = ix->d_cp_rows;
            k2_sqnorm = ix->d_cp_sqnorm;
            k2_n = ix->cp_pass;
            mask = nullptr;
            compacted = true;
        } else if (mask && mask_q_stride == 0 && ix->rows >= 65536) {
            const size_t nb = yrb::compact_scratch_words(ix->rows);
            if (nb > ix->cp_blocks_cap) {
                CK(cudaStreamSynchronize(st));
                FREE_DEV(ix->d_cp_blocks);
                CK(cudaMalloc(&ix->d_cp_blocks, nb * 4));
                ix->cp_blocks_cap = nb;
            }
            if (!ix->h_pass) CK(cudaMallocHost(&ix->h_pass, 8));
            ix->cp_key = 0;
            CK(yrb::launch_compact_count(mask, ix->rows, ix->d_cp_blocks, ix->d_pass, st));
            CK(cudaMemcpyAsync(ix->h_pass, ix->d_pass, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            ix->launches += 2;
            const int64_t pass = (int64_t)*ix->h_pass;
            if (pass >= k && pass <= ix->rows / 4) {
                // "gather4": K2 reads the passing rows in place through TMA tile::gather4 and the row map;
                // "copy": the rows are first copied into a contiguous scratch matrix.
                static int mode_g4 = -1;
                if (mode_g4 < 0) mode_g4 = getenv("YRB_K8_MODE") ? (strcmp(getenv("YRB_K8_MODE"), "gather4") == 0) : 0;
                use_rowmap = mode_g4 != 0;
                if (pass > ix->cp_rows_cap || (!use_rowmap && !ix->d_cp_rows)) {
                    const int64_t cap = (pass + pass / 4 + 255) / 256 * 256;
                    FREE_DEV(ix->d_cp_rows);
                    FREE_DEV(ix->d_cp_sqnorm);
                    FREE_DEV(ix->d_cp_map);
                    ix->cp_rows_cap = 0;
                    if (!use_rowmap) CK(cudaMalloc(&ix->d_cp_rows, (size_t)cap * ix->ld * yrb::elem_size(ix->dtype)));
                    CK(cudaMalloc(&ix->d_cp_sqnorm, (size_t)(cap + 256) * 4));
                    CK(cudaMalloc(&ix->d_cp_map, (size_t)(cap + 256) * 4));
                    ix->cp_rows_cap = cap;
                }
                CK(yrb::launch_compact_gather(mask, ix->rows, ix->d_cp_blocks, ix->d_cp_map, ix->d_rows, ix->d_sqnorm,
                                              ix->ld * yrb::elem_size(ix->dtype) / 16, pass, use_rowmap ? nullptr : ix->d_cp_rows,
                                              ix->d_cp_sqnorm, ix->sm_count, st));
                ix->launches += 2;
                if (use_rowmap) {
                    // pad the map to whole 128-row tiles with a valid row; the epilogue ignores compact rows >= pass
                    CK(cudaMemsetAsync(ix->d_cp_map + pass, 0, (size_t)((pass + 127) / 128 * 128 - pass) * 4, st));
                } else {
                    k2_rows = ix->d_cp_rows;
                }
                k2_sqnorm = ix->d_cp_sqnorm;
                k2_n = pass;
                mask = nullptr;
                compacted = true;
                ix->cp_key = mkey;     // 0 (no key: a caller-supplied mask) is never matched
                ix->cp_pass = pass;
                ix->cp_rowmap = use_rowmap;
            }
        }
        CK(yrb::launch_ingest(dev_q, nq, ix->dim, ix->ld, ix->metric, ix->dtype, ix->d_q, ix->d_qsq, st));
        const int nq_pad = (nq + 127) / 128 * 128;
        if (nq_pad > nq)
            CK(cudaMemsetAsync(reinterpret_cast<char*>(ix->d_q) + (size_t)nq * ix->ld * yrb::elem_size(ix->dtype), 0,
                               (size_t)(nq_pad - nq) * ix->ld * yrb::elem_size(ix->dtype), st));
        int launches = 1;
        cudaEvent_t ea, eb;
        int rc = prof_pair(ix, &ea, &eb);
        if (rc) return rc;
        std::string err;
        // compacted keys carry compact row numbers until K8 maps them back: the cross-shard finish then runs on its own
        rc = yrb::k2_search(ix->k2, k2_rows, k2_n, ix->capacity, ix->dim, ix->ld, ix->d_q, nq, k, mask,
                            mask_q_stride, ix->metric, ix->d_qsq, k2_sqnorm, out_keys, ids, scores, counts,
                            ix->sm_count & ~1 /* every SM: the exchange kernel's 16 KiB CTAs share an SM with a K2 CTA */, st,
                            &launches, err, ea, eb, pair, (compacted && use_rowmap) ? ix->d_cp_map : nullptr, ix->rows,
                            compacted ? nullptr : xs, min_score);
        if (rc) set_error(err);
        ix->launches += launches;
        if (!rc && compacted) {
            CK(yrb::launch_compact_remap(ix->d_cp_map, (int64_t)nq * k, out_keys, ids, st));
            ix->launches++;
            if (xs) {
                CK(yrb::launch_xshard_finish(*xs, out_keys, nq, k, st));
                ix->launches++;
            }
        }
        return rc;
    }
    if (path == 1 && ix->path == 0 && ix->dtype == 1 && nq >= 2 && k <= yrb::K1Q_MAX_K && mask_q_stride == 0) {
        // K1Q: fp32-storage batches share each pass over the rows between four queries (K2 is bf16-only)
        const int parts = yrb::k1_parts(sms);
        CK(yrb::launch_ingest(dev_q, nq, ix->dim, ix->ld, ix->metric, ix->dtype, ix->d_q, ix->d_qsq, st));
        ix->launches++;
        for (int c0 = 0; c0 < nq; c0 += yrb::K1Q_MAX_Q) {
            const int n = std::min(yrb::K1Q_MAX_Q, nq - c0);
            CK(yrb::launch_k1q_f32(ix->d_rows, ix->rows, ix->ld, reinterpret_cast<const float*>(ix->d_q) + (size_t)c0 * ix->ld, n,
                                   ix->d_qsq + c0, ix->d_sqnorm, ix->metric, mask, k, ix->d_parts + (size_t)c0 * parts * k, sms, st,
                                   floor_key));
            ix->launches++;
        }
        // one selection launch for the whole batch: query q's per-CTA lists sit at d_parts[q][cta][k]
        CK(yrb::launch_select_segments(ix->d_parts, k, (int64_t)parts * k, nullptr, 0, 0, parts, k, k, nullptr, nq, k, out_keys, st,
                                       ids, scores, counts, xs));
        ix->launches++;
        return YRB_OK;
    }
    if (path == 1) {
        const int parts = yrb::k1_parts(sms);
        for (int j = 0; j < nq; ++j) {
            uint64_t* pk = ix->d_parts + (size_t)j * parts * k;
            yrb::K1Out o{out_keys + (size_t)j * k, ids ? ids + (size_t)j * k : nullptr,
                         scores ? scores + (size_t)j * k : nullptr, counts ? counts + j : nullptr};
            o.floor_key = floor_key;
            yrb::XShard xj{};
            if (xs) {
                xj = *xs;
                xj.q0 += j;
                o.use_xs = 1;
                o.xs = xj;
            }
            bool fused = false;
            static const bool trace = getenv("YRB_K1_TRACE") != nullptr;
            if (trace) {
                if (!ix->d_k1trace) CK(cudaMalloc(&ix->d_k1trace, (size_t)256 * 8 * 8));
                CK(cudaMemsetAsync(ix->d_k1trace, 0, (size_t)256 * 8 * 8, st));
                o.trace = ix->d_k1trace;
            }
            int rc = prof_mark(ix, st);
            if (rc) return rc;
            int lists = parts;
            CK(yrb::launch_k1(ix->d_rows, ix->dtype, ix->rows, ix->dim, ix->ld, dev_q + (size_t)j * ix->dim, ix->d_sqnorm,
                              ix->metric, mask ? mask + (size_t)j * mask_q_stride : nullptr, k, pk, ix->d_ticket, o, &fused, sms, st,
                              &lists));
            if ((rc = prof_mark(ix, st))) return rc;
            ix->launches++;
            if (trace) {  // debug aid: phase breakdown of this launch on stderr (µs relative to the first CTA's start)
                std::vector<unsigned long long> h((size_t)(lists + 2) * 8);
                CK(cudaStreamSynchronize(st));
                CK(cudaMemcpy(h.data(), ix->d_k1trace, h.size() * 8, cudaMemcpyDeviceToHost));
                unsigned long long t0 = ~0ull;
                for (int c = 0; c < lists; ++c) t0 = std::min(t0, h[(size_t)c * 8]);
                const char* names[8] = {"start", "prep_done", "scan_done", "cta_merge_done", "ticket_done", "final_done", "warp0_scan_done", "first_group_done"};
                for (int ph = 0; ph < 8; ++ph) {
                    double lo = 1e30, hi = -1e30, sum = 0;
                    int n = 0;
                    for (int c = 0; c < lists; ++c) {
                        const unsigned long long t = h[(size_t)c * 8 + ph];
                        if (!t) continue;
                        const double us = (double)(t - t0) * 1e-3;
                        lo = std::min(lo, us);
                        hi = std::max(hi, us);
                        sum += us;
                        ++n;
                    }
                    if (n) fprintf(stderr, "[k1trace] %-16s n=%3d min %8.2f avg %8.2f max %8.2f us\n", names[ph], n, lo, sum / n, hi);
                }
                const unsigned long long* f = h.data() + (size_t)lists * 8;  // the last CTA's final merge
                const unsigned long long* c0 = f + 8;  // CTA 0's own merge
                if (c0[0])
                    fprintf(stderr, "[k1trace] cta0 merge:  begin %.2f init %.2f bound %.2f ranked %.2f (survivors %llu) emitted %.2f us\n",
                            (c0[0] - t0) * 1e-3, (c0[1] - t0) * 1e-3, (c0[2] - t0) * 1e-3, (c0[3] - t0) * 1e-3, c0[6], (c0[4] - t0) * 1e-3);
                if (f[0])
                    fprintf(stderr, "[k1trace] final merge: begin %.2f staged %.2f bound %.2f ranked %.2f (survivors %llu) emitted %.2f us\n",
                            (f[0] - t0) * 1e-3, (f[1] - t0) * 1e-3, (f[2] - t0) * 1e-3, (f[3] - t0) * 1e-3, f[6], (f[4] - t0) * 1e-3);
            }
            if (!fused) {
                CK(yrb::launch_select_segments(pk, k, 0, nullptr, 0, 0, lists, k, k, nullptr, 1, k, o.final_keys, st, o.ids,
                                               o.scores, o.count, xs ? &xj : nullptr));
                ix->launches++;
            }
        }
        return YRB_OK;
    }
    // path 3: key vector + radix select, k <= 4096
    if (k > 4096) return fail(YRB_ERR_UNSUPPORTED, "k=%d exceeds the largest supported n_results (4096)", k);
    if (ix->rowkeys_cap < ix->rows) {
        CK(cudaStreamSynchronize(st));
        FREE_DEV(ix->d_rowkeys);
        CK(cudaMalloc(&ix->d_rowkeys, (size_t)ix->capacity * 8));
        ix->rowkeys_cap = ix->capacity;
    }
    const size_t need = yrb::select_scratch_bytes(ix->rows, k);
    if (need > ix->select_bytes) {
        CK(cudaStreamSynchronize(st));
        FREE_DEV(ix->d_select);
        CK(cudaMalloc(&ix->d_select, need));
        ix->select_bytes = need;
    }
    for (int j = 0; j < nq; ++j) {
        CK(yrb::launch_scores(ix->d_rows, ix->dtype, ix->rows, ix->dim, ix->ld, dev_q + (size_t)j * ix->dim, ix->d_sqnorm,
                              ix->metric, mask ? mask + (size_t)j * mask_q_stride : nullptr, ix->d_rowkeys, ix->sm_count, st,
                              floor_key));
        CK(yrb::launch_select(ix->d_rowkeys, ix->rows, k, out_keys + (size_t)j * k, ix->d_select, ix->sm_count, st));
        ix->launches += 15;
    }
    if (decode) {
        CK(yrb::launch_decode(out_keys, nq, k, ids, scores, counts, st));
        ix->launches++;
    }
    if (xs) {
        CK(yrb::launch_xshard_finish(*xs, out_keys, nq, k, st));
        ix->launches++;
    }
    return YRB_OK;
}

int upload_user_mask(yrb_index* ix, const uint32_t* mask, const uint32_t** out, cudaStream_t st) {
    uint32_t* d_user = ix->d_usermask;
    const int64_t nw = (ix->rows + 31) / 32, nwp = mask_words(ix->rows);
    cudaError_t e = cudaMemsetAsync(d_user, 0, (size_t)nwp * 4, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_user, mask, (size_t)nw * 4, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && (ix->rows & 31)) {
        // clear bits past the last row of the last word
        uint32_t last = mask[nw - 1] & ((1u << (ix->rows & 31)) - 1u);
        e = cudaMemcpyAsync(d_user + nw - 1, &last, 4, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (e != cudaSuccess) return fail(YRB_ERR_CUDA, "mask upload failed: %s", cudaGetErrorString(e));
    *out = d_user;
    return YRB_OK;
}

int ensure_column(yrb_index* ix, int col, int col_type) {
    Column& c = ix->cols[col];
    if (c.type >= 0) return c.type == col_type ? YRB_OK : fail(YRB_ERR_INVALID, "column %d already has type %d", col, c.type);
    int rc;
    const size_t w = col_width(col_type);
    c.type = col_type;
    if ((rc = regrow(reinterpret_cast<char**>(&c.values), 0, (size_t)ix->capacity * w, true, ix->stream))) return rc;
    if ((rc = regrow(&c.present, 0, (size_t)mask_words(ix->capacity) * 4, true, ix->stream))) return rc;
    c.present_host.assign(mask_words(ix->capacity), 0u);
    return YRB_OK;
}

int append_device_locked(yrb_index* ix, const float* dev_rows, int64_t n, cudaStream_t st) {
    int rc = ensure_capacity(ix, ix->rows + n);
    if (rc) return rc;
    const size_t es = yrb::elem_size(ix->dtype);
    CK(yrb::launch_ingest(dev_rows, n, ix->dim, ix->ld, ix->metric, ix->dtype,
                          reinterpret_cast<char*>(ix->d_rows) + (size_t)ix->rows * ix->ld * es,
                          ix->d_sqnorm + ix->rows, st));
    ix->launches++;
    CK(cudaStreamSynchronize(st));
    rc = mark_appended(ix, ix->rows, n);
    if (rc) return rc;
    ix->rows += n;
    ix->epoch++;
    return YRB_OK;
}

}  // namespace yrbi

extern "C" {

int yrb_abi_version(void) { return YRB_ABI_VERSION; }
const char* yrb_last_error(void) { return last_error().c_str(); }

int yrb_device_count(int* out_count) {
    if (!out_count) return fail(YRB_ERR_INVALID, "out_count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        *out_count = 0;
        cudaGetLastError();
        return fail(YRB_ERR_NODEVICE, "no CUDA device visible (%s); this backend has no CPU fallback",
                    e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    *out_count = ok;
    if (!ok) return fail(YRB_ERR_NODEVICE, "no sm_100 (B200) device among %d CUDA devices; no fallback path", n);
    return YRB_OK;
}

int yrb_index_create(yrb_index** out, int device, int dim, int metric, int storage_dtype, int64_t reserve_rows) {
    if (!out) return fail(YRB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (dim < 1 || dim > 65536) return fail(YRB_ERR_INVALID, "dim %d out of range [1, 65536]", dim);
    if (metric < 0 || metric > 2) return fail(YRB_ERR_INVALID, "unknown metric %d", metric);
    if (storage_dtype < 0 || storage_dtype > 1) return fail(YRB_ERR_INVALID, "unknown storage dtype %d", storage_dtype);
    DevGuard dev_guard_;
    int n = 0;
    int rc = yrb_device_count(&n);
    if (rc) return rc;
    int major = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(YRB_ERR_NODEVICE, "device %d is not sm_100 (compute capability %d.x)", device, major);
    yrb_index* ix = new (std::nothrow) yrb_index();
    if (!ix) return fail(YRB_ERR_NOMEM, "host allocation failed");
    ix->device = device;
    ix->dim = dim;
    ix->metric = metric;
    ix->dtype = storage_dtype;
    ix->ld = yrb::row_ld(dim, storage_dtype);
    auto bail = [&](int code) {
        yrb_index_destroy(ix);
        return code;
    };
#define CKB(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return bail(fail(YRB_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    CKB(cudaSetDevice(device));
    CKB(cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, device));
    {
        static const bool private_streams = getenv("YRB_PRIVATE_STREAMS") && getenv("YRB_PRIVATE_STREAMS")[0] == '1';
        if (!private_streams) {
            ix->pool = device_pool(device);
            if (!ix->pool) return bail(fail(YRB_ERR_CUDA, "cannot create the stream of device %d", device));
            ix->stream = ix->pool->stream;
            ix->k2 = ix->pool->k2;
        } else {
            CKB(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
            ix->k2 = yrb::k2_create();
        }
    }
    CKB(cudaMalloc(&ix->d_prog, sizeof(yrb::WhereProgDev)));
    CKB(cudaMallocHost(&ix->h_prog, sizeof(yrb::WhereProgDev)));
    CKB(cudaMalloc(&ix->d_pass, 8));
    CKB(cudaMalloc(&ix->d_ticket, 8));  // [0] CTAs done, [1] next chunk
    CKB(cudaMemset(ix->d_ticket, 0, 8));
#undef CKB
    rc = ensure_capacity(ix, std::max<int64_t>(reserve_rows, 1));
    if (rc) return bail(rc);
    *out = ix;
    return YRB_OK;
}

int yrb_index_destroy(yrb_index* ix) {
    if (!ix) return YRB_OK;
    DevGuard dev_guard_;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    free_scratch(ix);
    FREE_DEV(ix->d_rows);
    FREE_DEV(ix->d_sqnorm);
    FREE_DEV(ix->d_live);
    FREE_DEV(ix->d_mask);
    FREE_DEV(ix->d_usermask);
    FREE_DEV(ix->d_rowkeys);
    FREE_DEV(ix->d_select);
    FREE_DEV(ix->d_prog);
    FREE_DEV(ix->d_pass);
    FREE_DEV(ix->d_ticket);
    FREE_DEV(ix->d_k1trace);
    FREE_DEV(ix->d_progs);
    FREE_HOST(ix->h_progs);
    FREE_DEV(ix->d_qmasks);
    FREE_DEV(ix->d_cp_blocks);
    FREE_DEV(ix->d_cp_rows);
    FREE_DEV(ix->d_cp_sqnorm);
    FREE_DEV(ix->d_cp_map);
    FREE_HOST(ix->h_pass);
    FREE_HOST(ix->h_prog);
    FREE_HOST(ix->h_stage);
    FREE_DEV(ix->d_append);
    for (auto& kv : ix->cols) {
        FREE_DEV(kv.second.values);
        FREE_DEV(kv.second.present);
    }
    if (ix->k2 && !ix->pool) yrb::k2_destroy(ix->k2);
    for (cudaEvent_t e : ix->prof_ev) cudaEventDestroy(e);
    if (ix->stream && !ix->pool) cudaStreamDestroy(ix->stream);
    delete ix;
    return YRB_OK;
}

int yrb_index_reserve(yrb_index* ix, int64_t rows) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    return ensure_capacity(ix, rows);
}

int yrb_index_count(const yrb_index* ix, int64_t* out_rows, int64_t* out_live) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (out_rows) *out_rows = ix->rows;
    if (out_live) *out_live = ix->rows - ix->n_dead;
    return YRB_OK;
}

int yrb_index_info(const yrb_index* ix, int* out_dim, int* out_ld, int* out_metric, int* out_dtype, int* out_device,
                   int64_t* out_capacity) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (out_dim) *out_dim = ix->dim;
    if (out_ld) *out_ld = ix->ld;
    if (out_metric) *out_metric = ix->metric;
    if (out_dtype) *out_dtype = ix->dtype;
    if (out_device) *out_device = ix->device;
    if (out_capacity) *out_capacity = ix->capacity;
    return YRB_OK;
}

int yrb_index_append_host_f32(yrb_index* ix, const float* rows, int64_t n) {
    Nvtx nvtx_("yrb_index_append_host_f32");
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && !rows)) return fail(YRB_ERR_INVALID, "bad rows/n");
    if (n == 0) return YRB_OK;
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    if (ix->rows + n > 0xfffffffell) return fail(YRB_ERR_UNSUPPORTED, "more than 2^32-2 rows per GPU shard");
    if ((rc = ensure_capacity(ix, ix->rows + n))) return rc;
    // staged in chunks: pageable → pinned → device fp32 scratch → K5 into place
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(32u << 20) / ((int64_t)ix->dim * 4));
    if ((rc = ensure_stage(ix, (size_t)chunk * ix->dim * 4))) return rc;
    const size_t need = (size_t)std::min<int64_t>(chunk, n) * ix->dim * 4;
    if (need > ix->append_bytes) {
        FREE_DEV(ix->d_append);
        ix->append_bytes = 0;
        CK(cudaMalloc(&ix->d_append, need));
        ix->append_bytes = need;
    }
    for (int64_t r = 0; r < n; r += chunk) {
        const int64_t m = std::min(chunk, n - r);
        memcpy(ix->h_stage, rows + r * ix->dim, (size_t)m * ix->dim * 4);
        CK(cudaMemcpyAsync(ix->d_append, ix->h_stage, (size_t)m * ix->dim * 4, cudaMemcpyHostToDevice, ix->stream));
        if ((rc = append_device_locked(ix, ix->d_append, m, ix->stream))) return rc;
    }
    return YRB_OK;
}

int yrb_index_append_device_f32(yrb_index* ix, const float* dev_rows, int64_t n, void* stream) {
    Nvtx nvtx_("yrb_index_append_device_f32");
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && !dev_rows)) return fail(YRB_ERR_INVALID, "bad rows/n");
    if (n == 0) return YRB_OK;
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    if (ix->rows + n > 0xfffffffell) return fail(YRB_ERR_UNSUPPORTED, "more than 2^32-2 rows per GPU shard");
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    if (stream) CK(cudaStreamSynchronize(st));  // producer of dev_rows may still be running on `stream`
    return append_device_locked(ix, dev_rows, n, st);
}

int yrb_index_read_rows(yrb_index* ix, const int64_t* row_ids, int64_t n, float* out_rows) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && (!row_ids || !out_rows))) return fail(YRB_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    const size_t es = yrb::elem_size(ix->dtype);
    std::vector<unsigned char> tmp((size_t)ix->ld * es);
    // coalesce runs of consecutive ids into one copy
    int64_t i = 0;
    while (i < n) {
        int64_t j = i;
        while (j + 1 < n && row_ids[j + 1] == row_ids[j] + 1) ++j;
        const int64_t r0 = row_ids[i], cnt = j - i + 1;
        if (r0 < 0 || r0 + cnt > ix->rows) return fail(YRB_ERR_INVALID, "row id %lld out of range", (long long)r0);
        tmp.resize((size_t)cnt * ix->ld * es);
        CK(cudaMemcpyAsync(tmp.data(), reinterpret_cast<char*>(ix->d_rows) + (size_t)r0 * ix->ld * es, tmp.size(),
                           cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
        for (int64_t r = 0; r < cnt; ++r) {
            float* o = out_rows + (i + r) * ix->dim;
            if (ix->dtype == YRB_DTYPE_F32) {
                memcpy(o, tmp.data() + (size_t)r * ix->ld * 4, (size_t)ix->dim * 4);
            } else {
                const uint16_t* b = reinterpret_cast<const uint16_t*>(tmp.data()) + (size_t)r * ix->ld;
                for (int d = 0; d < ix->dim; ++d) {
                    uint32_t u = (uint32_t)b[d] << 16;
                    memcpy(&o[d], &u, 4);
                }
            }
        }
        i = j + 1;
    }
    return YRB_OK;
}

int yrb_index_read_raw(yrb_index* ix, int64_t row_begin, int64_t n, void* out_rows, float* out_sqnorm) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || row_begin < 0 || row_begin + n > ix->rows) return fail(YRB_ERR_INVALID, "row range out of bounds");
    if (n == 0) return YRB_OK;
    if (!out_rows || !out_sqnorm) return fail(YRB_ERR_INVALID, "output buffers are NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    const size_t rb = (size_t)ix->ld * yrb::elem_size(ix->dtype);
    CK(cudaMemcpyAsync(out_rows, reinterpret_cast<char*>(ix->d_rows) + (size_t)row_begin * rb, (size_t)n * rb,
                       cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaMemcpyAsync(out_sqnorm, ix->d_sqnorm + row_begin, (size_t)n * 4, cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    return YRB_OK;
}

int yrb_index_append_raw(yrb_index* ix, const void* rows, const float* sqnorm, int64_t n) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && (!rows || !sqnorm))) return fail(YRB_ERR_INVALID, "bad arguments");
    if (n == 0) return YRB_OK;
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    if (ix->rows + n > 0xfffffffell) return fail(YRB_ERR_UNSUPPORTED, "more than 2^32-2 rows per GPU shard");
    if ((rc = ensure_capacity(ix, ix->rows + n))) return rc;
    const size_t rb = (size_t)ix->ld * yrb::elem_size(ix->dtype);
    CK(cudaMemcpyAsync(reinterpret_cast<char*>(ix->d_rows) + (size_t)ix->rows * rb, rows, (size_t)n * rb,
                       cudaMemcpyHostToDevice, ix->stream));
    CK(cudaMemcpyAsync(ix->d_sqnorm + ix->rows, sqnorm, (size_t)n * 4, cudaMemcpyHostToDevice, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    if ((rc = mark_appended(ix, ix->rows, n))) return rc;
    ix->rows += n;
    ix->epoch++;
    return YRB_OK;
}

int yrb_index_set_live(yrb_index* ix, const int64_t* row_ids, int64_t n, int live) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || (n > 0 && !row_ids)) return fail(YRB_ERR_INVALID, "bad arguments");
    if (n == 0) return YRB_OK;
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i)
        if (row_ids[i] < 0 || row_ids[i] >= ix->rows)
            return fail(YRB_ERR_INVALID, "row id %lld out of range", (long long)row_ids[i]);
    int64_t wmin = INT64_MAX, wmax = -1;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = row_ids[i], w = r >> 5;
        const uint32_t bit = 1u << (r & 31);
        const bool was = ix->h_live[w] & bit;
        if (live && !was) {
            ix->h_live[w] |= bit;
            ix->n_dead--;
        } else if (!live && was) {
            ix->h_live[w] &= ~bit;
            ix->n_dead++;
        }
        wmin = std::min(wmin, w);
        wmax = std::max(wmax, w);
    }
    ix->epoch++;
    return upload_words(ix->d_live, ix->h_live, wmin, wmax + 1, ix->stream);
}

int yrb_index_truncate(yrb_index* ix, int64_t rows) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (rows < 0 || rows > ix->rows) return fail(YRB_ERR_INVALID, "truncate to %lld rows: index holds %lld", (long long)rows, (long long)ix->rows);
    if (rows == ix->rows) return YRB_OK;
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ix->stream));
    // rows [rows, ix->rows) disappear: their live / presence bits are cleared so a later append starts clean
    const int64_t w0 = rows >> 5, w1 = ((ix->rows - 1) >> 5) + 1;
    auto wipe = [&](std::vector<uint32_t>& bits, bool count_dead) {
        for (int64_t r = rows; r < ix->rows; ++r) {
            const uint32_t bit = 1u << (r & 31);
            if (count_dead && !(bits[r >> 5] & bit)) ix->n_dead--;
            bits[r >> 5] &= ~bit;
        }
    };
    wipe(ix->h_live, true);
    if ((rc = upload_words(ix->d_live, ix->h_live, w0, w1, ix->stream))) return rc;
    for (auto& kv : ix->cols) {
        wipe(kv.second.present_host, false);
        if ((rc = upload_words(kv.second.present, kv.second.present_host, w0, w1, ix->stream))) return rc;
    }
    ix->rows = rows;
    ix->epoch++;
    return YRB_OK;
}

int yrb_index_clear(yrb_index* ix) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ix->stream));
    ix->rows = 0;
    ix->epoch++;
    ix->n_dead = 0;
    std::fill(ix->h_live.begin(), ix->h_live.end(), 0u);
    CK(cudaMemsetAsync(ix->d_live, 0, (size_t)mask_words(ix->capacity) * 4, ix->stream));
    for (auto& kv : ix->cols) {
        FREE_DEV(kv.second.values);
        FREE_DEV(kv.second.present);
    }
    ix->cols.clear();
    CK(cudaStreamSynchronize(ix->stream));
    return YRB_OK;
}

int yrb_index_column_write(yrb_index* ix, int col, int col_type, int64_t row_begin, int64_t n, const void* values,
                           const uint8_t* present) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (col < 0) return fail(YRB_ERR_INVALID, "column id must be >= 0");
    if (col_type < YRB_COL_I64 || col_type > YRB_COL_BOOL) return fail(YRB_ERR_INVALID, "bad column type %d", col_type);
    if (n < 0 || row_begin < 0 || (n > 0 && (!values || !present))) return fail(YRB_ERR_INVALID, "bad arguments");
    if (n == 0) return YRB_OK;
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    if (row_begin + n > ix->rows) return fail(YRB_ERR_INVALID, "column rows [%lld,%lld) beyond appended rows %lld",
                                               (long long)row_begin, (long long)(row_begin + n), (long long)ix->rows);
    if ((rc = ensure_column(ix, col, col_type))) return rc;
    Column& c = ix->cols[col];
    const size_t w = col_width(col_type);
    CK(cudaMemcpyAsync(reinterpret_cast<char*>(c.values) + (size_t)row_begin * w, values, (size_t)n * w,
                       cudaMemcpyHostToDevice, ix->stream));
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = row_begin + i;
        if (present[i]) c.present_host[r >> 5] |= (1u << (r & 31));
        else c.present_host[r >> 5] &= ~(1u << (r & 31));
    }
    ix->epoch++;
    return upload_words(c.present, c.present_host, row_begin >> 5, ((row_begin + n - 1) >> 5) + 1, ix->stream);
}

int yrb_index_where(yrb_index* ix, const yrb_where* w, uint32_t* out_mask, int64_t* out_pass) {
    Nvtx nvtx_("yrb_index_where");
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    if (ix->rows == 0) {
        if (out_pass) *out_pass = 0;
        return YRB_OK;
    }
    yrb_where all = {nullptr, 0, nullptr, 0, nullptr, 0};
    const uint32_t* m = nullptr;
    if ((rc = resolve_mask(ix, w ? w : &all, nullptr, &m, ix->stream, true))) return rc;
    unsigned long long pass = 0;
    CK(cudaMemcpyAsync(&pass, ix->d_pass, 8, cudaMemcpyDeviceToHost, ix->stream));
    if (out_mask)
        CK(cudaMemcpyAsync(out_mask, m, (size_t)((ix->rows + 31) / 32) * 4, cudaMemcpyDeviceToHost, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    if (out_pass) *out_pass = (int64_t)pass;
    return YRB_OK;
}

static int search_host(yrb_index* ix, const float* queries, int nq, int k, const yrb_where* w,
                       const yrb_where* const* per_query, const uint32_t* mask, int64_t* out_ids, float* out_scores,
                       int32_t* out_counts, float min_score = -INFINITY) {
    Nvtx nvtx_("yrb_index_search");
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (nq < 1 || !queries) return fail(YRB_ERR_INVALID, "need at least one query");
    if (k < 1) return fail(YRB_ERR_INVALID, "k must be >= 1 (got %d)", k);
    if (!out_ids || !out_scores) return fail(YRB_ERR_INVALID, "output buffers are NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    const int64_t live_rows = ix->rows - ix->n_dead;
    if (live_rows == 0) {
        for (int64_t i = 0; i < (int64_t)nq * k; ++i) {
            out_ids[i] = -1;
            out_scores[i] = -INFINITY;
        }
        if (out_counts)
            for (int q = 0; q < nq; ++q) out_counts[q] = 0;
        return YRB_OK;
    }
    // never select more than exist: keeps the fused paths usable for small collections
    const int ke = (int)std::min<int64_t>(k, ix->rows);
    if ((rc = ensure_scratch(ix, nq, ke))) return rc;
    cudaStream_t st = ix->stream;
    memcpy(ix->h_q, queries, (size_t)nq * ix->dim * 4);
    CK(cudaMemcpyAsync(ix->d_qf32, ix->h_q, (size_t)nq * ix->dim * 4, cudaMemcpyHostToDevice, st));
    static const bool zc_enabled = [] {
        const char* e = getenv("YRB_ZERO_COPY");  // "0" keeps the D2H copy (A/B measurements)
        return !(e && e[0] == '0');
    }();
    const bool zero_copy = zc_enabled && ix->d_result_host && (size_t)nq * ke * 12 + (size_t)nq * 4 <= ZERO_COPY_RESULT_MAX;
    const size_t res_bytes = result_views(ix, nq, ke, zero_copy ? ix->d_result_host : ix->d_result);
    const uint32_t* dev_extra = nullptr;
    if (mask && (rc = upload_user_mask(ix, mask, &dev_extra, st))) return rc;
    const uint32_t* m = nullptr;
    int64_t m_stride = 0;
    if (per_query) {
        // one filter per query (text2sql-style batches): every distinct program is evaluated once into
        // its row of the [nq, words] mask matrix
        rc = resolve_masks_multi(ix, per_query, nq, dev_extra, &m, &m_stride, st);
    } else {
        rc = resolve_mask(ix, w, dev_extra, &m, st, false);
    }
    if (!rc) rc = scan_select(ix, ix->d_qf32, nq, ke, m, m_stride, ix->d_keys, ix->d_ids, ix->d_scores, ix->d_counts, st, nullptr, min_score);
    if (!rc) {
        cudaError_t e = zero_copy ? cudaSuccess : cudaMemcpyAsync(ix->h_result, ix->d_result, res_bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(YRB_ERR_CUDA, "search failed: %s", cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(st);
    }
    if (rc) return rc;
    const int64_t* h_ids = reinterpret_cast<const int64_t*>(ix->h_result);
    const float* h_scores = reinterpret_cast<const float*>(ix->h_result + (size_t)nq * ke * 8);
    const int32_t* h_counts = reinterpret_cast<const int32_t*>(ix->h_result + (size_t)nq * ke * 12);
    if (ke == k) {  // the usual case: the packed result has the caller's layout
        memcpy(out_ids, h_ids, (size_t)nq * k * 8);
        memcpy(out_scores, h_scores, (size_t)nq * k * 4);
        if (out_counts) memcpy(out_counts, h_counts, (size_t)nq * 4);
        return YRB_OK;
    }
    for (int q = 0; q < nq; ++q) {
        for (int j = 0; j < k; ++j) {
            const bool ok = j < ke;
            out_ids[(int64_t)q * k + j] = ok ? h_ids[(int64_t)q * ke + j] : -1;
            out_scores[(int64_t)q * k + j] = ok ? h_scores[(int64_t)q * ke + j] : -INFINITY;
        }
        if (out_counts) out_counts[q] = h_counts[q];
    }
    return YRB_OK;
}

int yrb_index_search(yrb_index* ix, const float* queries, int nq, int k, const yrb_where* w, const uint32_t* mask,
                     int64_t* out_ids, float* out_scores, int32_t* out_counts) {
    return search_host(ix, queries, nq, k, w, nullptr, mask, out_ids, out_scores, out_counts);
}

int yrb_index_search_ex(yrb_index* ix, const float* queries, int nq, int k, const yrb_where* w, const yrb_where* const* wheres,
                        const uint32_t* mask, const yrb_search_opts* opts, int64_t* out_ids, float* out_scores, int32_t* out_counts) {
    if (w && wheres) return fail(YRB_ERR_INVALID, "pass a shared filter or per-query filters, not both");
    if (wheres && mask) return fail(YRB_ERR_INVALID, "a host bitmask cannot be combined with per-query filters");
    const float ms = opts ? opts->min_score : -INFINITY;
    if (ms != ms) return fail(YRB_ERR_INVALID, "min_score is NaN");
    return search_host(ix, queries, nq, k, w, wheres, mask, out_ids, out_scores, out_counts, ms);
}

int yrb_index_search_multi(yrb_index* ix, const float* queries, int nq, int k, const yrb_where* const* wheres,
                           int64_t* out_ids, float* out_scores, int32_t* out_counts) {
    if (!wheres) return fail(YRB_ERR_INVALID, "wheres is NULL (use yrb_index_search for a shared filter)");
    return search_host(ix, queries, nq, k, nullptr, wheres, nullptr, out_ids, out_scores, out_counts);
}

int yrb_index_search_device(yrb_index* ix, const float* dev_queries, int nq, int k, const uint32_t* dev_mask,
                            uint64_t* dev_out_keys, void* stream) {
    Nvtx nvtx_("yrb_index_search_device");
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (nq < 1 || !dev_queries || !dev_out_keys) return fail(YRB_ERR_INVALID, "bad arguments");
    if (k < 1) return fail(YRB_ERR_INVALID, "k must be >= 1 (got %d)", k);
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    if (ix->rows == 0) {
        CK(cudaMemsetAsync(dev_out_keys, 0, (size_t)nq * k * 8, st));
        return YRB_OK;
    }
    if (k > ix->rows) return fail(YRB_ERR_INVALID, "k=%d exceeds rows=%lld (device variant does not clamp)", k, (long long)ix->rows);
    if ((rc = ensure_scratch(ix, nq, k))) return rc;
    const uint32_t* m = nullptr;
    if ((rc = resolve_mask(ix, nullptr, dev_mask, &m, st, false))) return rc;
    return scan_select(ix, dev_queries, nq, k, m, 0, dev_out_keys, nullptr, nullptr, nullptr, st);
}

int yrb_index_search_device_ids(yrb_index* ix, const float* dev_queries, int nq, int k, const uint32_t* dev_mask,
                                int64_t* dev_out_ids, float* dev_out_scores, int32_t* dev_out_counts, void* stream) {
    Nvtx nvtx_("yrb_index_search_device_ids");
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (nq < 1 || !dev_queries || !dev_out_ids || !dev_out_scores || !dev_out_counts)
        return fail(YRB_ERR_INVALID, "bad arguments");
    if (k < 1) return fail(YRB_ERR_INVALID, "k must be >= 1 (got %d)", k);
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    if (k > ix->rows) return fail(YRB_ERR_INVALID, "k=%d exceeds rows=%lld (device variant does not clamp)", k, (long long)ix->rows);
    if ((rc = ensure_scratch(ix, nq, k))) return rc;
    const uint32_t* m = nullptr;
    if ((rc = resolve_mask(ix, nullptr, dev_mask, &m, st, false))) return rc;
    return scan_select(ix, dev_queries, nq, k, m, 0, ix->d_keys, dev_out_ids, dev_out_scores, dev_out_counts, st);
}

int yrb_merge_topk_device(int device, const uint64_t* dev_keys, int parts, int nq, int k, const int64_t* dev_row_base,
                          int64_t* dev_out_ids, float* dev_out_scores, int32_t* dev_out_counts, void* stream) {
    if (!dev_keys || !dev_row_base || !dev_out_ids || !dev_out_scores) return fail(YRB_ERR_INVALID, "NULL buffer");
    if (parts < 1 || nq < 1 || k < 1) return fail(YRB_ERR_INVALID, "bad parts/nq/k");
    if ((int64_t)parts * k > 2048) return fail(YRB_ERR_UNSUPPORTED, "parts*k = %lld exceeds 2048", (long long)parts * k);
    DevGuard dev_guard_;
    CK(cudaSetDevice(device));
    CK(yrb::launch_merge_global(dev_keys, parts, nq, k, dev_row_base, dev_out_ids, dev_out_scores, dev_out_counts,
                                (cudaStream_t)stream));
    return YRB_OK;
}

int yrb_index_set_path(yrb_index* ix, int path) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (path < 0 || path > 4) return fail(YRB_ERR_INVALID, "path must be 0..4");
    ix->path = path;
    return YRB_OK;
}

int yrb_index_profile(yrb_index* ix, int enable) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    ix->prof = enable != 0;
    return YRB_OK;
}

int yrb_index_profile_read(yrb_index* ix, double* out_total_ms, int64_t* out_launches) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    int rc = set_dev(ix);
    if (rc) return rc;
    if ((rc = prof_flush(ix))) return rc;
    if (out_total_ms) *out_total_ms = ix->prof_ms;
    if (out_launches) *out_launches = ix->prof_n;
    ix->prof_ms = 0.0;
    ix->prof_n = 0;
    return YRB_OK;
}

int yrb_index_set_reserved_sms(yrb_index* ix, int n) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (n < 0 || n >= ix->sm_count) return fail(YRB_ERR_INVALID, "reserved SMs must be in [0, %d)", ix->sm_count);
    std::lock_guard<std::mutex> g(ix->mu);
    DevGuard dev_guard_;
    ix->reserved_sms = n;
    return YRB_OK;
}

int yrb_index_cache_stats(const yrb_index* ix, int64_t* out_filter_hits, int64_t* out_compaction_hits) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (out_filter_hits) *out_filter_hits = ix->cache_hits_k4;
    if (out_compaction_hits) *out_compaction_hits = ix->cache_hits_k8;
    return YRB_OK;
}

int yrb_index_stats(const yrb_index* ix, int64_t* out_kernel_launches) {
    if (!ix) return fail(YRB_ERR_INVALID, "index is NULL");
    if (out_kernel_launches) *out_kernel_launches = ix->launches;
    return YRB_OK;
}

}  // extern "C"
