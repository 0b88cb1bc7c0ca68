/* A plain-C11 consumer of include/yrb200.h: proves the header is C (no C++-isms), that every entry links from C, and —
 * on a box without a B200 — that the library refuses to work instead of falling back to the CPU.
 * Built and run by tests/test_abi.py with `gcc -std=c11 -Wall -Wextra -Werror -pedantic`. */
#include <stdio.h>
#include <string.h>

#include "yrb200.h"

int main(void) {
    if (yrb_abi_version() != YRB_ABI_VERSION) {
        printf("FAIL abi version %d\n", yrb_abi_version());
        return 1;
    }
    /* the row map of a sharded collection is stateless host arithmetic */
    int shard = -1;
    int64_t local = -1, global = -1, rows = -1;
    if (yrb_shard_locate(8, 16384, 10000000 - 1, &shard, &local) != YRB_OK || yrb_shard_global(8, 16384, shard, local, &global) != YRB_OK ||
        global != 10000000 - 1 || yrb_shard_rows(8, 16384, 10000000, shard, &rows) != YRB_OK || local >= rows) {
        printf("FAIL shard map: shard %d local %lld global %lld rows %lld\n", shard, (long long)local, (long long)global, (long long)rows);
        return 1;
    }
    if (yrb_shard_locate(0, 16384, 1, &shard, &local) != YRB_ERR_INVALID || strlen(yrb_last_error()) == 0) {
        printf("FAIL: bad arguments were accepted\n");
        return 1;
    }
    int n = -1;
    int rc = yrb_device_count(&n);
    yrb_index* ix = NULL;
    yrb_search_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.min_score = 0.5f;
    if (rc == YRB_ERR_NODEVICE) {
        /* no sm_100 device: every constructor must fail loudly, there is no CPU path */
        int devs[2] = {0, 1};
        yrb_sharded* sh = NULL;
        if (yrb_index_create(&ix, 0, 1024, YRB_METRIC_COSINE, YRB_DTYPE_BF16, 0) != YRB_ERR_NODEVICE || ix != NULL ||
            yrb_sharded_create(&sh, devs, 2, 1024, YRB_METRIC_COSINE, YRB_DTYPE_BF16, 0, 0) == YRB_OK || sh != NULL) {
            printf("FAIL: an index was created without a device\n");
            return 1;
        }
        printf("OK no-device: %s\n", yrb_last_error());
        return 0;
    }
    if (rc != YRB_OK || n < 1) {
        printf("FAIL device count rc=%d n=%d\n", rc, n);
        return 1;
    }
    /* with a B200: a 3-row collection, one search with a score threshold */
    {
        float rowsf[3 * 8] = {1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0};
        float q[8] = {1, 0.1f, 0, 0, 0, 0, 0, 0};
        int64_t ids[3];
        float scores[3];
        int32_t count = 0;
        if (yrb_index_create(&ix, 0, 8, YRB_METRIC_COSINE, YRB_DTYPE_F32, 0) != YRB_OK || yrb_index_append_host_f32(ix, rowsf, 3) != YRB_OK ||
            yrb_index_search_ex(ix, q, 1, 3, NULL, NULL, NULL, &opts, ids, scores, &count) != YRB_OK) {
            printf("FAIL search: %s\n", yrb_last_error());
            return 1;
        }
        if (count != 2 || ids[0] != 0 || ids[1] != 2 || ids[2] != -1) {
            printf("FAIL result: count %d ids %lld %lld %lld\n", count, (long long)ids[0], (long long)ids[1], (long long)ids[2]);
            return 1;
        }
        yrb_index_destroy(ix);
        printf("OK device: top hit row %lld score %.4f, %d hits above %.2f\n", (long long)ids[0], scores[0], count, opts.min_score);
    }
    return 0;
}
