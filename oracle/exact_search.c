/* exact_search.c — plain-C restatement of the dense-retrieval hot path (TEST INFRASTRUCTURE, not product).
 *
 * PARITY UNPINNED with respect to the reference's engines (chromadb 1.3.4 / faiss-cpu 1.12.0 are not
 * installable here; see oracle/exact_search.py and DESIGN.md §6).  Follows the same call sites:
 *   utu/rag/storage/implementations/chroma_store.py:118-135   query → score = 1 - distance, pre-filter
 *   utu/rag/storage/implementations/faiss_store.py:98-110,143-154   cosine = normalised rows + inner product
 * and the numeric contract of DESIGN.md §3: operands are the STORED values (bf16 bit patterns or fp32),
 * accumulation in double, order (score desc, row id asc), bitmask bit i of word j = row 32j+i.
 *
 * It is a third, independent statement of the path (numpy oracle, this file, the CUDA kernels); tests compare
 * them.  Built by oracle/Makefile into oracle/_build/liboracle.so and called through ctypes from tests only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float bf16_to_f32(uint16_t b) {
    uint32_t u = (uint32_t)b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* fp32 -> bf16 bits, round to nearest even (what K5 stores) */
uint16_t oracle_bf16_rne(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* NaN stays NaN */
    uint32_t lsb = (u >> 16) & 1u;
    return (uint16_t)((u + 0x7fffu + lsb) >> 16);
}

/* faiss.normalize_L2 with the fp64 pin: y = (float)(x / sqrt(sum x^2)), zero rows untouched */
void oracle_normalize(const float* x, int64_t n, int dim, float* out) {
    for (int64_t r = 0; r < n; ++r) {
        double ss = 0.0;
        for (int d = 0; d < dim; ++d) ss += (double)x[r * dim + d] * (double)x[r * dim + d];
        double nrm = ss > 0.0 ? sqrt(ss) : 1.0;
        for (int d = 0; d < dim; ++d) out[r * dim + d] = (float)((double)x[r * dim + d] / nrm);
    }
}

static inline int better(double sa, int64_t ia, double sb, int64_t ib) { return sa > sb || (sa == sb && ia < ib); }

/* rows: n x ld stored values (dtype 0: uint16 bf16 bits, 1: float); q: dim prepared values (float, already rounded
 * to the storage dtype); metric 0 cosine / 1 dot: score = q.x ; 2 euclidean: score = 1 - |q-x|^2.
 * mask: NULL or ceil(n/32) words.  Writes up to k (id, score) pairs, best first; returns their number. */
int oracle_topk(const void* rows, int dtype, int64_t n, int dim, int ld, const float* q, int metric, const uint32_t* mask,
                int k, int64_t* out_ids, double* out_scores) {
    int have = 0;
    for (int64_t r = 0; r < n; ++r) {
        if (mask && !((mask[r >> 5] >> (r & 31)) & 1u)) continue;
        double s = 0.0;
        if (dtype == 0) {
            const uint16_t* x = (const uint16_t*)rows + r * ld;
            if (metric == 2) {
                for (int d = 0; d < dim; ++d) {
                    double df = (double)q[d] - (double)bf16_to_f32(x[d]);
                    s += df * df;
                }
            } else {
                for (int d = 0; d < dim; ++d) s += (double)q[d] * (double)bf16_to_f32(x[d]);
            }
        } else {
            const float* x = (const float*)rows + r * ld;
            if (metric == 2) {
                for (int d = 0; d < dim; ++d) {
                    double df = (double)q[d] - (double)x[d];
                    s += df * df;
                }
            } else {
                for (int d = 0; d < dim; ++d) s += (double)q[d] * (double)x[d];
            }
        }
        if (metric == 2) s = 1.0 - s;
        /* insertion into the sorted result list */
        if (have == k && !better(s, r, out_scores[k - 1], out_ids[k - 1])) continue;
        int pos = have < k ? have : k - 1;
        while (pos > 0 && better(s, r, out_scores[pos - 1], out_ids[pos - 1])) {
            out_scores[pos] = out_scores[pos - 1];
            out_ids[pos] = out_ids[pos - 1];
            --pos;
        }
        out_scores[pos] = s;
        out_ids[pos] = r;
        if (have < k) ++have;
    }
    return have;
}
