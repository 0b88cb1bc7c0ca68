// K4 — metadata filter: a compiled Chroma `where` tree evaluated over columnar metadata into
// the row bitmask that K1/K2 consume.  Replaces the sqlite metadata pre-filter inside
// collection.query(where=…) (chroma_store.py:104-120); the filter language is the one the
// reference's producers emit (kb_search_toolkit.py:63-96, meta_retrieval_toolkit.py:102-255,
// memory_store.py:403-417).  Semantics: DESIGN.md §5 (typed compare, $ne/$nin true on missing).
//
// One thread per row, a warp ballot forms each 32-row mask word; column reads are coalesced.
// HBM-bound: Σ(referenced column widths)·N + N/8 bytes.
#include "common.cuh"
#include "kernels.h"

namespace yrb {

enum { COL_I64 = 0, COL_F64 = 1, COL_CODE = 2, COL_BOOL = 3 };
enum { OP_EQ = 0, OP_NE, OP_GT, OP_GTE, OP_LT, OP_LTE, OP_IN, OP_NIN };

template <typename T>
__device__ __forceinline__ bool leaf_cmp(T v, int op, const int64_t* opnd, int cnt) {
    auto get = [&](int i) -> T {
        if constexpr (sizeof(T) == 8) {
            T r;
            memcpy(&r, &opnd[i], 8);
            return r;
        } else {
            return (T)opnd[i];
        }
    };
    switch (op) {
        case OP_EQ:
        case OP_NE: return v == get(0);
        case OP_GT: return v > get(0);
        case OP_GTE: return v >= get(0);
        case OP_LT: return v < get(0);
        case OP_LTE: return v <= get(0);
        default: {
            bool any = false;
            for (int i = 0; i < cnt; ++i) any |= (v == get(i));
            return any;
        }
    }
}

__global__ void __launch_bounds__(256)
    where_kernel(const WhereProgDev* __restrict__ prog, int64_t n_rows, const uint32_t* __restrict__ live,
                 const uint32_t* __restrict__ extra, uint32_t* __restrict__ out_mask, int64_t n_words_out,
                 unsigned long long* __restrict__ pass_count) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t word = row >> 5;
    bool pass = false;
    if (row < n_rows) {
        uint64_t stack = 0;
        const int np = prog->n_postfix;
        for (int t = 0; t < np; ++t) {
            const int tok = prog->postfix[t];
            if (tok >= 0) {
                const WhereLeafDev& lf = prog->leaves[tok];
                bool hit = false;
                if (lf.col_type >= 0) {
                    const bool present = (lf.present[word] >> (row & 31)) & 1u;
                    if (present) {
                        const int64_t* opnd = prog->operands + lf.operand_begin;
                        switch (lf.col_type) {
                            case COL_I64:
                                hit = leaf_cmp<int64_t>(reinterpret_cast<const int64_t*>(lf.values)[row], lf.op, opnd,
                                                        lf.operand_count);
                                break;
                            case COL_F64:
                                hit = leaf_cmp<double>(reinterpret_cast<const double*>(lf.values)[row], lf.op, opnd,
                                                       lf.operand_count);
                                break;
                            case COL_CODE:
                                hit = leaf_cmp<int32_t>(reinterpret_cast<const int32_t*>(lf.values)[row], lf.op, opnd,
                                                        lf.operand_count);
                                break;
                            default:
                                hit = leaf_cmp<int32_t>((int32_t) reinterpret_cast<const uint8_t*>(lf.values)[row],
                                                        lf.op, opnd, lf.operand_count);
                                break;
                        }
                    }
                }
                if (lf.op == OP_NE || lf.op == OP_NIN) hit = !hit;
                stack = (stack << 1) | (hit ? 1ull : 0ull);
            } else if (tok == -3) {
                stack ^= 1ull;
            } else {
                const uint64_t a = stack & 1ull, b = (stack >> 1) & 1ull;
                stack = ((stack >> 2) << 1) | (tok == -1 ? (a & b) : (a | b));
            }
        }
        pass = (np == 0) ? true : (stack & 1ull);
        if (live) pass = pass && ((live[word] >> (row & 31)) & 1u);
        if (extra) pass = pass && ((extra[word] >> (row & 31)) & 1u);
    }
    const uint32_t bits = __ballot_sync(YRB_FULL, pass);
    if ((threadIdx.x & 31) == 0) {
        if (word < n_words_out) out_mask[word] = bits;
        if (pass_count && bits) atomicAdd(pass_count, (unsigned long long)__popc(bits));
    }
}

__global__ void mask_and_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int64_t n,
                                uint32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] & (b ? b[i] : 0xffffffffu);
}

cudaError_t launch_where(const WhereProgDev* prog, int64_t n_rows, const uint32_t* live, const uint32_t* extra,
                         uint32_t* out_mask, unsigned long long* pass_count, cudaStream_t st) {
    // cover whole mask words up to an even word count so K1's 64-row groups read defined bits
    int64_t n_words = (n_rows + 31) / 32;
    n_words = (n_words + 1) & ~1ll;
    if (n_words == 0) return cudaSuccess;
    const int64_t threads_needed = n_words * 32;
    const int64_t blocks = (threads_needed + 255) / 256;
    where_kernel<<<(unsigned)blocks, 256, 0, st>>>(prog, n_rows, live, extra, out_mask, n_words, pass_count);
    return cudaGetLastError();
}

cudaError_t launch_mask_and(const uint32_t* a, const uint32_t* b, int64_t n_words, uint32_t* out, cudaStream_t st) {
    if (n_words <= 0) return cudaSuccess;
    mask_and_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, st>>>(a, b, n_words, out);
    return cudaGetLastError();
}

}  // namespace yrb
