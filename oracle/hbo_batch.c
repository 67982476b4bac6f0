/*
 * TEST INFRASTRUCTURE (oracle) -- NOT part of the product path.
 *
 * Multi-threaded driver of the plain-C port for bench.py's cpu_baseline / --impl reference legs:
 * per candidate  copy search + star MSA (hbo_align.c)  ->  remove_sparse_col (Util.py:10344)  ->
 * judge_boundary_v5 (Util.py:9145), fanned out over host threads the way the reference fans
 * run_find_members_v8 out over a ProcessPoolExecutor (Util.py:8141-8147).
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "hbo.h"

typedef struct {
    const hbo_genome *g; const uint64_t *idx; uint64_t n_idx; int k, w;
    const uint8_t *cand; const uint64_t *coff; uint32_t n_cand;
    int occ_cap, max_loci, flank, max_rows, plant;
    hbo_result *out;
    atomic_uint next;
} batch_t;

static void *worker(void *arg) {
    batch_t *b = (batch_t *)arg;
    hbo_copy *copies = (hbo_copy *)malloc(sizeof(hbo_copy) * (size_t)(b->max_loci + 1));
    for (;;) {
        uint32_t c = atomic_fetch_add(&b->next, 1);
        if (c >= b->n_cand) break;
        const uint8_t *q = b->cand + b->coff[c];
        int L = (int)(b->coff[c + 1] - b->coff[c]);
        uint64_t cap = (uint64_t)(b->max_rows + 1) * (uint64_t)(L + 2 * b->flank) * 4 + 1024;
        uint8_t *mat = (uint8_t *)malloc(cap);
        uint8_t *clean = (uint8_t *)malloc(cap);
        char *cons = (char *)malloc((size_t)(L + 2 * b->flank) * 4 + 1024);
        int ncp = 0, R = 0, C = 0, C2 = 0;
        hbo_result *o = &b->out[c];
        memset(o, 0, sizeof(*o));
        if (hbo_candidate_msa(b->g, b->idx, b->n_idx, b->k, b->w, q, L, b->occ_cap, b->max_loci, b->flank, b->max_rows,
                              copies, &ncp, mat, cap, &R, &C) == 0 && R > 0) {
            hbo_remove_sparse_cols((const char *)mat, R, C, (char *)clean, &C2);
            hbo_call call;
            hbo_judge_v5((const char *)q, L, (const char *)clean, R, C2, b->plant, 1, &call, cons);
            o->is_te = call.is_te; o->info = call.info; o->copy_num = call.copy_num; o->cons_len = call.cons_len;
        }
        free(mat); free(clean); free(cons);
    }
    free(copies);
    return NULL;
}

int hbo_run_batch(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w, const uint8_t *cand_ascii,
                  const uint64_t *coff, uint32_t n_cand, int occ_cap, int max_loci, int flank, int max_rows, int plant,
                  int threads, hbo_result *out) {
    batch_t b = {g, idx, n_idx, k, w, cand_ascii, coff, n_cand, occ_cap, max_loci, flank, max_rows, plant, out, 0};
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; ++t) pthread_create(&th[t], NULL, worker, &b);
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th);
    return 0;
}
