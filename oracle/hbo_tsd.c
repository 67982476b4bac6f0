/*
 * TEST INFRASTRUCTURE (oracle) -- NOT part of the product path.
 *
 * Plain-C restatement of the TSD-bounded variant generation that runs right before the
 * boundary-adjustment path in HiTE's TIR module:
 *     search_confident_tir_v4   reference module/Util.py:7734-7845
 *     get_short_tir_contigs     reference module/Util.py:7297-7334   (classification only)
 *     getReverseSequence        reference module/Util.py:1635-1648
 * Pinned against tests/golden/tsd_v4.json.gz (outputs of the imported, unmodified reference).
 *
 * One thing cannot be restated bit for bit: the reference sorts a Python *set* of
 * (tsd, start, end) tuples by distance with a stable sort, so entries of equal distance come out
 * in set-iteration order, which depends on the per-process string hash seed.  The index `i` that
 * goes into the record name ('-C_<i>-') and the choice among equal-distance entries at the
 * 100-record cut are therefore not reproducible even between two runs of the reference.  Here ties
 * are ordered by (distance, start, end, k); every record also carries the range [idx_lo, idx_hi]
 * its index can take under any tie order, which is what the golden test checks.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "hbo.h"

static const int KS[9] = {2, 3, 4, 5, 6, 8, 9, 10, 11};   /* TIR_TSDs, Util.py:7760 */

typedef struct { int k, start, end, dist; const char *kmer; } tup_t;

static int tup_cmp(const void *a, const void *b) {
    const tup_t *x = (const tup_t *)a, *y = (const tup_t *)b;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    return (x->k > y->k) - (x->k < y->k);
}

static char rc_base(char c) {                                /* getReverseSequence, Util.py:1635 */
    switch (c) { case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; default: return 'N'; }
}

/* get_short_tir_contigs (Util.py:7297-7334) on tir = seq[start .. end]:
   0 none, 1 hAT, 2 Mutator, 3 CACTA, 4 CCC */
int hbo_short_tir_class(const char *tir, int len, int tsd_len, int plant) {
    if (len < 5) return 0;
    int eq5 = 1, eq3 = 1;
    for (int i = 0; i < 5; ++i) if (tir[i] != rc_base(tir[len - 1 - i])) eq5 = 0;
    for (int i = 0; i < 3; ++i) if (tir[i] != rc_base(tir[len - 1 - i])) eq3 = 0;
    if (eq5) {
        if (tsd_len == 8 && len < 4000) return 1;
        if (tsd_len >= 9 && tsd_len <= 11) return 2;
        if (plant == 1 && tsd_len == 3 && (memcmp(tir, "CACTA", 5) == 0 || memcmp(tir, "CACTG", 5) == 0)) return 3;
        return 0;
    }
    if (eq3 && memcmp(tir, "CCC", 3) == 0) return 4;
    return 0;
}

/* search_confident_tir_v4 (Util.py:7734-7845).  raw_start1 / raw_end1 are the 1-based boundaries the
   reference receives.  Writes at most 100 records (cap >= 100) in output order; returns their number. */
int hbo_search_confident_tir_v4(const char *seq, int n, int raw_start1, int raw_end1, int d, int plant,
                                hbo_tsd_rec *out, int cap) {
    const int rs = raw_start1 - 1, re = raw_end1 - 1;                          /* :7736-7737 */
    int ls = rs - d; if (ls < 0) ls = 0;                                        /* :7743-7747 */
    int le = rs + d + 1; if (le > n) le = n; if (le < ls) le = ls;
    int rts = re - d; if (rts < 0) rts = 0;                                     /* :7750-7754 */
    int rte = re + d + 1; if (rte > n) rte = n; if (rte < rts) rte = rts;
    if (ls > n) ls = n;
    if (rts > n) rts = n;
    const char *L = seq + ls, *R = seq + rts;
    const int nl = le - ls, nr = rte - rts;
    tup_t *tups = (tup_t *)malloc(sizeof(tup_t) * (size_t)(9 * (nr > 0 ? nr : 1)));
    int nt = 0;
    for (int ki = 0; ki < 9; ++ki) {
        const int k = KS[ki];
        for (int j = 0; j + k <= nr; ++j) {
            const int pos = rts + j - 1;                                        /* :7783 */
            if (pos < 0 || pos > n - 1) continue;
            /* left_pos of this k-mer: closest to rs, first one on ties (:7763-7778) */
            int best_l = -1, best_ld = 0;
            for (int i = 0; i + k <= nl; ++i) {
                const int lp = ls + i + k;
                if (lp > n - 1) continue;
                if (memcmp(L + i, R + j, (size_t)k) != 0) continue;
                const int dd = abs(lp - rs);
                if (best_l < 0 || dd < best_ld) { best_l = lp; best_ld = dd; }
            }
            if (best_l < 0) continue;
            /* right_pos after this occurrence = running closest to re over the occurrences so far (:7788-7796):
               this occurrence contributes a NEW tuple only when it is itself the running best */
            int is_best = 1;
            const int my = abs(pos - re);
            for (int j2 = 0; j2 < j; ++j2) {
                const int p2 = rts + j2 - 1;
                if (p2 < 0 || p2 > n - 1) continue;
                if (memcmp(R + j2, R + j, (size_t)k) != 0) continue;
                if (abs(p2 - re) <= my) { is_best = 0; break; }
            }
            if (!is_best) continue;
            const int ts = best_l, te = pos;
            const char *km = R + j;
            int ok = 0;                                                         /* :7802-7805 */
            if (k != 2 && k != 4) ok = 1;
            else if (k == 4) ok = memcmp(km, "TTAA", 4) == 0;
            else {
                if (memcmp(km, "TA", 2) == 0) ok = 1;
                else if (plant == 0) {
                    /* first_3bp = seq[ts:ts+3], last_3bp = seq[te-2:te+1]; a negative te-2 wraps in Python and
                       gives an empty or unrelated slice of a > 100 bp sequence: only te >= 2 can yield 'GGG' here */
                    const int f3 = ts + 3 <= n && memcmp(seq + ts, "CCC", 3) == 0;
                    const int l3 = te >= 2 && te + 1 <= n && memcmp(seq + te - 2, "GGG", 3) == 0;
                    ok = f3 && l3;
                }
            }
            if (!ok) continue;
            tups[nt].k = k; tups[nt].start = ts; tups[nt].end = te; tups[nt].kmer = km;
            tups[nt].dist = abs(ts - rs) + abs(te - re);
            ++nt;
        }
    }
    qsort(tups, (size_t)nt, sizeof(tup_t), tup_cmp);                            /* :7808, ties: see header */
    int n_out = 0;
    for (int i = 0; i < nt && n_out < 100 && n_out < cap; ++i) {                /* :7811-7845 */
        const tup_t *t = &tups[i];
        int has_nn = 0;
        for (int z = 0; z + 1 < t->k; ++z) if (t->kmer[z] == 'N' && t->kmer[z + 1] == 'N') has_nn = 1;
        if (has_nn) continue;
        const int len = t->end >= t->start ? t->end - t->start + 1 : 0;
        if (len < 100) continue;
        const char *tir = seq + t->start;
        if (tir[0] == 'T' && tir[1] == 'G' && tir[len - 2] == 'C' && tir[len - 1] == 'A') continue;
        if (memcmp(tir, "TATATATA", 8) == 0 || memcmp(tir, "ATATATAT", 8) == 0) continue;
        int lo = i, hi = i;
        while (lo > 0 && tups[lo - 1].dist == t->dist) --lo;
        while (hi + 1 < nt && tups[hi + 1].dist == t->dist) ++hi;
        hbo_tsd_rec *o = &out[n_out++];
        o->k = t->k; o->start = t->start; o->end = t->end; o->dist = t->dist; o->idx = i; o->idx_lo = lo; o->idx_hi = hi;
        o->short_class = hbo_short_tir_class(tir, len, t->k, plant);
    }
    free(tups);
    return n_out;
}
