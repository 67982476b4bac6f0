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

/* ---------------------------------------------------------------------------------------------- */
/* search_polyA_TSD (reference module/Util.py:10915-11006) with find_nearest_polyA / find_nearest_polyT (:10865, :10903),
   find_longest_polyA / polyT (:10839, :10877) and find_nearest_tandem (:9772-9819): the per-candidate scan of the non-LTR
   candidate generator (get_candidate_non_LTR :11008, called right before flank_region_align_v5 in
   judge_Non_LTR_transposons.py:34).  Pinned against tests/golden/polya_tsd.json.gz (outputs of the imported reference;
   fuzzysearch.find_near_matches is only asked "is there a match", which is unambiguous: some substring of the k-mer
   within Levenshtein distance 1 of the TSD). */
static int py_clip(int x, int n) { return x < 0 ? 0 : (x > n ? n : x); }

/* longest run of `base` of length >= min_length in s[0..n): first longest; -> start/end or -1 */
static void longest_run(const char *s, int n, char base, int min_length, int *out_start, int *out_end) {
    int max_len = 0, cur = 0, start = 0, ms = -1, me = -1;
    for (int i = 0; i < n; ++i) {
        if (s[i] == base) { if (++cur == 1) start = i; }
        else { if (cur >= min_length && cur > max_len) { max_len = cur; ms = start; me = i; } cur = 0; }
    }
    if (cur >= min_length && cur > max_len) { ms = start; me = n; }
    *out_start = ms; *out_end = me;
}

/* find_nearest_polyA / polyT: -> (start, end, len of sequence[start:end]) or (-1, -1, 0) */
static void nearest_run(const char *seq, int n, int position, char base, int *ps, int *pe, int *plen) {
    const int lo = py_clip(position - 25 < 0 ? 0 : position - 25, n), hi = py_clip(position + 25, n);
    int ms = -1, me = -1;
    if (hi > lo) longest_run(seq + lo, hi - lo, base, 6, &ms, &me);
    if (ms == -1 || me == -1) { *ps = -1; *pe = -1; *plen = 0; return; }
    ms = position - 25 + ms; if (ms < 0) ms = 0;               /* the offset ignores the clipping of the window (:10872) */
    me = position - 25 + me; if (me < 0) me = 0;
    *ps = ms; *pe = me;
    const int a = py_clip(ms, n), b = py_clip(me, n);
    *plen = b > a ? b - a : 0;
}

/* find_nearest_tandem: 4 exact copies of a 2..6-mer starting inside the +-25 window; longest motif wins, first position */
static void nearest_tandem(const char *seq, int n, int pos, int *ps, int *pe, int *plen) {
    const int start = pos - 25 < 0 ? 0 : pos - 25, end = pos + 25 < n ? pos + 25 : n;
    *ps = -1; *pe = -1; *plen = 0;
    for (int m = 2; m <= 6; ++m)
        for (int i = start; i <= end - 4 * m; ++i) {
            int ok = 1;
            for (int z = m; z < 4 * m && ok; ++z) if (seq[i + z] != seq[i + z % m]) ok = 0;
            if (ok && 4 * m > *plen) { *plen = 4 * m; *ps = i; *pe = i + 4 * m; }
        }
}

/* is some substring of kmer[0..k) within Levenshtein distance 1 of tsd[0..k) ? */
static int near1(const char *tsd, const char *kmer, int k) {
    int ham = 0;
    for (int i = 0; i < k; ++i) ham += tsd[i] != kmer[i];
    if (ham <= 1) return 1;
    /* tsd minus one character == kmer[0..k-1) or kmer[1..k) */
    for (int off = 0; off <= 1; ++off) {
        const char *x = kmer + off;
        int lcp = 0, lcs = 0;
        while (lcp < k - 1 && tsd[lcp] == x[lcp]) ++lcp;
        while (lcs < k - 1 && tsd[k - 1 - lcs] == x[k - 2 - lcs]) ++lcs;
        if (lcp + lcs >= k - 1) return 1;
    }
    return 0;
}

int hbo_search_polyA_TSD(const char *seq, int n, int flanking_len, hbo_polya_rec *out) {
    const int raw_start = flanking_len + 1, raw_end = n - flanking_len;          /* :10917-10918 */
    int end_3 = -1, end_5 = -1, direct = 0;
    int as, ae, alen, ts, te, tlen;
    if (raw_end >= 0) {
        nearest_run(seq, n, raw_end, 'A', &as, &ae, &alen);
        nearest_tandem(seq, n, raw_end, &ts, &te, &tlen);
        if ((alen < tlen ? alen : tlen) > 0) { end_3 = alen > tlen ? ae : te; end_5 = raw_start; direct = 1; }   /* :10926-10934 */
    }
    nearest_run(seq, n, raw_start, 'T', &as, &ae, &alen);
    nearest_tandem(seq, n, raw_start, &ts, &te, &tlen);
    if ((alen < tlen ? alen : tlen) > 0) { end_3 = alen > tlen ? as : ts; end_5 = raw_end; direct = 2; }         /* :10940-10948 */
    int found = 0, tsd_k = 0;
    if (end_3 != -1 && end_5 != -1 && direct) {
        const int lo5 = end_5 - 25 < 0 ? 0 : end_5 - 25;
        const int lo = py_clip(lo5, n), hi = py_clip(end_5 + 25 < 0 ? 0 : end_5 + 25, n);
        const int sublen = hi > lo ? hi - lo : 0;
        for (int k = 20; k >= 8 && !found; --k) {
            const int t0 = direct == 2 ? end_3 - k : end_3;
            if (t0 < 0 || t0 + k > n) continue;                                   /* len(TSD) != k */
            const char *tsd = seq + t0;
            int has_n = 0;
            for (int z = 0; z < k; ++z) if (tsd[z] == 'N') has_n = 1;
            for (int i = 0; i + k <= sublen; ++i)
                if (near1(tsd, seq + lo + i, k) && !has_n) {
                    end_5 = lo5 + i + (direct == 1 ? k : 0);                      /* :10971 / :10992 */
                    found = 1; tsd_k = k;
                    break;
                }
        }
    }
    out->found = found; out->direct = direct; out->end_5 = end_5; out->end_3 = end_3; out->tsd_k = tsd_k;
    int a = end_5 < end_3 ? end_5 : end_3, b = end_5 < end_3 ? end_3 : end_5;
    if (!direct) { a = b = 0; }
    out->seq_start = a < 0 ? (a + n < 0 ? 0 : a + n) : py_clip(a, n);               /* Python slice bounds of seq[min:max] (:11002) */
    out->seq_end = b < 0 ? (b + n < 0 ? 0 : b + n) : py_clip(b, n);
    if (out->seq_end < out->seq_start) out->seq_end = out->seq_start;
    return found;
}

/* ---- HB-itr: checker of hb_itr_search (hite_b200/csrc/hb_itr.cu) -------------------------------------------------------
   The reference runs the closed binary `tools/itrsearch -i 0.7 -l 7` (module/Util.py:216-224); its behaviour was probed as
   a black box (tests/golden/make_golden_itr.py keeps what it returned): 5' end against the reverse complement of the 3' end,
   anchored at the record's ends, window min(max_w, len / 2), match +10, mismatch -16, gap of n bases -(32 + 32 n), best end
   cell; kept when score - p > 0 (p = 64 for arms of equal length, 32 for a longer 5' arm, 0 for a shorter one: the binary's
   "Quality", fitted to its outputs), m / (m + x) >= min_ident and the 5' arm >= min_len.  This is the exact optimum
   of that scoring, written the slow obvious way: three full tables of (score, m, x, g) tuples, compared as the kernel's
   packed keys compare (score, then m, then x, then g), best cell = first maximum in row-major order.  PARITY UNPINNED with
   respect to the binary (its own search is heuristic and not always optimal); agreement is measured in the golden file. */
typedef struct { int ok; int sc, m, x, g; } itr_cell;
static int itr_gt(const itr_cell *a, const itr_cell *b) {          /* a > b as packed keys; unreachable = smallest */
    if (!a->ok) return 0;
    if (!b->ok) return 1;
    if (a->sc != b->sc) return a->sc > b->sc;
    if (a->m != b->m) return a->m > b->m;
    if (a->x != b->x) return a->x > b->x;
    return a->g > b->g;
}
static itr_cell itr_max(itr_cell a, itr_cell b) { return itr_gt(&b, &a) ? b : a; }
static itr_cell itr_add(itr_cell a, int dsc, int dm, int dx, int dg) {
    if (a.ok) { a.sc += dsc; a.m += dm; a.x += dx; a.g += dg; }
    return a;
}
static int itr_code(char c) {
    switch (c & 0xDF) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}
/* out[7] = found, len5, len3, score, n_match, n_mismatch, n_gap */
void hbo_itr_search(const char *s, int len, double min_ident, int min_len, int max_w, int32_t *out) {
    const int w = len / 2 < max_w ? len / 2 : max_w;
    const size_t W = (size_t)w + 1;
    itr_cell *D = (itr_cell *)calloc(W * W, sizeof(itr_cell)), *E = (itr_cell *)calloc(W * W, sizeof(itr_cell)),
             *F = (itr_cell *)calloc(W * W, sizeof(itr_cell));
    D[0].ok = 1;
    itr_cell best = D[0];
    int bi = 0, bj = 0;
    for (int i = 0; i <= w; ++i)
        for (int j = 0; j <= w; ++j) {
            if (i == 0 && j == 0) continue;
            const size_t c = (size_t)i * W + (size_t)j;
            if (j > 0) E[c] = itr_max(itr_add(itr_max(D[c - 1], F[c - 1]), -64, 0, 0, 1), itr_add(E[c - 1], -32, 0, 0, 1));
            if (i > 0) F[c] = itr_max(itr_add(itr_max(D[c - W], E[c - W]), -64, 0, 0, 1), itr_add(F[c - W], -32, 0, 0, 1));
            if (i > 0 && j > 0) {
                const int ca = itr_code(s[i - 1]), cb = itr_code(s[len - j]);
                const int m = ca < 4 && cb < 4 && ca == 3 - cb;
                D[c] = itr_add(itr_max(itr_max(D[c - W - 1], E[c - W - 1]), F[c - W - 1]), m ? 10 : -16, m, !m, 0);
            }
            if (i > 0 && j > 0) {                                   /* row 0 / column 0 never beat the origin */
                const itr_cell cm = itr_max(itr_max(D[c], E[c]), F[c]);
                if (itr_gt(&cm, &best)) { best = cm; bi = i; bj = j; }
            }
        }
    out[1] = bi; out[2] = bj; out[3] = best.sc; out[4] = best.m; out[5] = best.x; out[6] = best.g;
    out[0] = (best.sc - (bi == bj ? 64 : (bi > bj ? 32 : 0)) > 0) && (best.m + best.x > 0) &&
             ((double)best.m / (double)(best.m + best.x) >= min_ident) && bi >= min_len;
    free(D); free(E); free(F);
}
