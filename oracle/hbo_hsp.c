/*
 * TEST INFRASTRUCTURE (oracle) -- NOT part of the product path.
 *
 * Plain C restatement of HB-hsp: the seed-pair self-join that replaces the
 * n^2 `blastn` runs of process_blast_alignments (reference module/Util.py:4724-4780).  blastn
 * (rmblast 2.14.0, environment.yml:23) is third-party and absent, so its raw HSPs are "parity
 * unpinned"; the CUDA kernels of hite_b200/csrc/hb_coarse.cu must match THIS list bit for bit, and the
 * FMEA stage downstream is pinned on the reference (hbo_fmea.c).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "hbo.h"

#define HSP_MAXGAP 64
#define HSP_MINSPAN 50

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
static uint64_t lb(const uint64_t *a, uint64_t lo, uint64_t hi, uint64_t key) {
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
static int seg_of(const uint64_t *goff, int n, uint64_t pos) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (goff[mid] <= pos) lo = mid + 1; else hi = mid; }
    return lo - 1;
}

/* seg_goff: global padded start of every segment (ascending).  Returns the number of HSPs (writes <= cap). */
uint64_t hbo_self_hsps(const uint64_t *mm_pos, const uint64_t *idx, uint64_t n, uint64_t tot, int k,
                       const uint64_t *seg_goff, int n_seg, int occ_cap, hbo_hsp_rec *out, uint64_t cap) {
    uint64_t np = 0, pcap = 1 << 16;
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * pcap);
    for (uint64_t qi = 0; qi < n; ++qi) {
        const uint64_t e = mm_pos[qi];
        const uint64_t h = e >> 33;
        const uint64_t l = lb(idx, 0, n, h << 33), r = lb(idx, l, n, (h + 1) << 33);
        uint64_t win = r - l; if (win > (uint64_t)occ_cap + 1) win = (uint64_t)occ_cap + 1;
        const int64_t qpos = (int64_t)((e >> 1) & 0xFFFFFFFFull);
        int taken = 0;
        for (uint64_t x = 0; x < win && taken < occ_cap; ++x) {
            const uint64_t g = idx[l + x];
            if (g == e) continue;
            const int64_t gpos = (int64_t)((g >> 1) & 0xFFFFFFFFull);
            const uint64_t rel = (e ^ g) & 1;
            const int64_t d = (rel ? ((int64_t)tot - gpos - k) : gpos) - qpos + (1ll << 32);
            if (np == pcap) { pcap *= 2; keys = (uint64_t *)realloc(keys, sizeof(uint64_t) * pcap); }
            keys[np++] = (rel << 62) | ((uint64_t)d << 29) | qi;
            ++taken;
        }
    }
    qsort(keys, np, sizeof(uint64_t), cmp_u64);
    uint64_t nh = 0;
    for (uint64_t i = 0; i < np;) {
        uint64_t j = i + 1;
        const int rel = (int)(keys[i] >> 62) & 1;
        const int64_t d = (int64_t)((keys[i] >> 29) & 0x1FFFFFFFFull) - (1ll << 32);
        while (j < np && (keys[j] >> 29) == (keys[i] >> 29)) {
            const int64_t qa = (int64_t)((mm_pos[keys[j - 1] & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
            const int64_t qb = (int64_t)((mm_pos[keys[j] & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
            if (qb - qa > HSP_MAXGAP) break;
            const int64_t sa = rel ? (int64_t)tot - (qa + d) - k : qa + d;
            const int64_t sb = rel ? (int64_t)tot - (qb + d) - k : qb + d;
            if (seg_of(seg_goff, n_seg, (uint64_t)qa) != seg_of(seg_goff, n_seg, (uint64_t)qb)) break;
            if (seg_of(seg_goff, n_seg, (uint64_t)sa) != seg_of(seg_goff, n_seg, (uint64_t)sb)) break;
            ++j;
        }
        const int64_t q0 = (int64_t)((mm_pos[keys[i] & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
        const int64_t q1 = (int64_t)((mm_pos[keys[j - 1] & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
        if (q1 + k - q0 >= HSP_MINSPAN) {
            const int64_t s0 = rel ? (int64_t)tot - (q0 + d) - k : q0 + d;
            const int64_t s1 = rel ? (int64_t)tot - (q1 + d) - k : q1 + d;
            const int qseg = seg_of(seg_goff, n_seg, (uint64_t)q0), sseg = seg_of(seg_goff, n_seg, (uint64_t)s0);
            const int64_t qg = (int64_t)seg_goff[qseg], sg = (int64_t)seg_goff[sseg];
            if (nh < cap) {
                hbo_hsp_rec *o = &out[nh];
                o->qseg = qseg; o->subj = sseg;
                o->qs = (int32_t)(q0 - qg + 1); o->qe = (int32_t)(q1 + k - qg);
                if (!rel) { o->ss = (int32_t)(s0 - sg + 1); o->se = (int32_t)(s1 + k - sg); }
                else { o->ss = (int32_t)(s0 + k - sg); o->se = (int32_t)(s1 - sg + 1); }
            }
            ++nh;
        }
        i = j;
    }
    free(keys);
    return nh;
}
