/*
 * TEST INFRASTRUCTURE (oracle) -- NOT part of the product path.
 *
 * Plain scalar C restatement of the reference's boundary-judging functions, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the
 * checker for the CUDA path in hite_b200/csrc.  Every function cites the reference lines it
 * follows (paths relative to /root/reference/module/).
 *
 * Pinned against golden fixtures generated from the *imported, unmodified* reference functions
 * (tests/golden/make_golden.py -> tests/golden/judge_v5.json.gz ...).  The only unpinned piece is
 * fuzzysearch.find_near_matches (third-party, absent): it follows oracle/fnm.py ("HB-fnm").
 *
 * Alphabet: A C G T N '-' ; any other byte is treated as N (documented in DESIGN.md).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "hbo.h"

/* ------------------------------------------------------------------------------------------ */
/* symbols: 0..3 = ACGT, 4 = N/other, 5 = gap                                                  */
static inline int sym_of(char c) {
    switch (c) {
        case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
        case '-': return 5; default: return 4;
    }
}
static const char SYM_CHR[6] = {'A', 'C', 'G', 'T', 'N', '-'};

/* Python slice bounds s[a:b] on a sequence of length n (step 1). */
static void py_slice(int n, long a, long b, int *lo, int *hi) {
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    if (b < a) b = a;
    *lo = (int)a; *hi = (int)b;
}

/* ------------------------------------------------------------------------------------------ */
/* Levenshtein.distance (Util.py:9403)                                                         */
int hbo_levenshtein(const char *a, int la, const char *b, int lb) {
    int *prev = (int *)malloc(sizeof(int) * (size_t)(lb + 1));
    int *cur = (int *)malloc(sizeof(int) * (size_t)(lb + 1));
    for (int j = 0; j <= lb; ++j) prev[j] = j;
    for (int i = 1; i <= la; ++i) {
        cur[0] = i;
        for (int j = 1; j <= lb; ++j) {
            int v = prev[j] + 1;
            if (cur[j - 1] + 1 < v) v = cur[j - 1] + 1;
            int d = prev[j - 1] + (a[i - 1] != b[j - 1]);
            if (d < v) v = d;
            cur[j] = v;
        }
        int *t = prev; prev = cur; cur = t;
    }
    int r = prev[lb];
    free(prev); free(cur);
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* HB-fnm: oracle/fnm.py (stand-in for fuzzysearch.find_near_matches, Util.py:9173-9174).      */
/* Returns the number of consolidated matches (sorted); fills first/last match.                */
typedef struct { int start, end, dist; } fnm_match;

static int fnm_longest_start(const char *p, int m, const char *s, int e, int dist) {
    int lo = e - m - dist; if (lo < 0) lo = 0;
    int maxk = e - lo;
    int *row = (int *)malloc(sizeof(int) * (size_t)(maxk + 1));
    int *nw = (int *)malloc(sizeof(int) * (size_t)(maxk + 1));
    for (int k = 0; k <= maxk; ++k) row[k] = k;
    for (int j = 1; j <= m; ++j) {
        char pj = p[m - j];
        nw[0] = j;
        for (int k = 1; k <= maxk; ++k) {
            int v = row[k] + 1;
            if (nw[k - 1] + 1 < v) v = nw[k - 1] + 1;
            int d = row[k - 1] + (pj != s[e - k]);
            if (d < v) v = d;
            nw[k] = v;
        }
        int *t = row; row = nw; nw = t;
    }
    int best = -1;
    for (int k = maxk; k >= 0; --k) if (row[k] == dist) { best = e - k; break; }
    free(row); free(nw);
    return best;
}

static int fnm_better(const fnm_match *a, const fnm_match *b) { /* a strictly better than b */
    if (a->dist != b->dist) return a->dist < b->dist;
    int la = a->end - a->start, lb = b->end - b->start;
    if (la != lb) return la > lb;
    return a->start < b->start;
}

static int cmp_match(const void *x, const void *y) {
    const fnm_match *a = (const fnm_match *)x, *b = (const fnm_match *)y;
    if (a->start != b->start) return a->start < b->start ? -1 : 1;
    if (a->end != b->end) return a->end < b->end ? -1 : 1;
    return (a->dist > b->dist) - (a->dist < b->dist);
}

int hbo_find_near_matches(const char *p, int m, const char *s, int n, int d,
                          int *first_start, int *first_end, int *last_start, int *last_end) {
    int *col = (int *)malloc(sizeof(int) * (size_t)(m + 1));
    fnm_match *c = (fnm_match *)malloc(sizeof(fnm_match) * (size_t)(n + 2));
    int nc = 0;
    for (int j = 0; j <= m; ++j) col[j] = j;
    if (col[m] <= d) { c[nc].start = 0; c[nc].end = 0; c[nc].dist = col[m]; ++nc; }
    for (int e = 1; e <= n; ++e) {
        char ch = s[e - 1];
        int prev_diag = col[0];
        col[0] = 0;
        for (int j = 1; j <= m; ++j) {
            int tmp = col[j];
            int v = col[j] + 1;
            if (col[j - 1] + 1 < v) v = col[j - 1] + 1;
            int dg = prev_diag + (p[j - 1] != ch);
            if (dg < v) v = dg;
            col[j] = v;
            prev_diag = tmp;
        }
        if (col[m] <= d) { c[nc].end = e; c[nc].dist = col[m]; c[nc].start = -1; ++nc; }
    }
    for (int i = 0; i < nc; ++i)
        if (c[i].start < 0) c[i].start = fnm_longest_start(p, m, s, c[i].end, c[i].dist);
    qsort(c, (size_t)nc, sizeof(fnm_match), cmp_match);
    int ngroups = 0;
    int gs = 0, ge = 0;
    fnm_match best = {0, 0, 0}, first_best = {0, 0, 0}, last_best = {0, 0, 0};
    for (int i = 0; i < nc; ++i) {
        if (ngroups > 0 && c[i].start < ge && c[i].end > gs) {
            if (c[i].start < gs) gs = c[i].start;
            if (c[i].end > ge) ge = c[i].end;
            if (fnm_better(&c[i], &best)) best = c[i];
        } else {
            if (ngroups == 1) first_best = best;
            ++ngroups;
            gs = c[i].start; ge = c[i].end; best = c[i];
        }
    }
    if (ngroups == 1) first_best = best;
    last_best = best;
    /* "sorted(best_matches)": group bests are already ordered by start because groups are
       disjoint start-ordered intervals */
    if (ngroups > 0) {
        if (first_start) *first_start = first_best.start;
        if (first_end) *first_end = first_best.end;
        if (last_start) *last_start = last_best.start;
        if (last_end) *last_end = last_best.end;
    }
    free(col); free(c);
    return ngroups;
}

/* ------------------------------------------------------------------------------------------ */
/* remove_sparse_col_in_align_file  Util.py:10344-10405                                        */
int hbo_remove_sparse_cols(const char *m, int R, int C, char *out, int *C_out) {
    if (R <= 0 || C <= 0) { *C_out = 0; return 0; }
    int *keep = (int *)malloc(sizeof(int) * (size_t)C);
    int nk = 0;
    for (int c = 0; c < C; ++c) {
        int gaps = 0;
        for (int r = 0; r < R; ++r) gaps += (m[(size_t)r * C + c] == '-');
        /* :10384-10391: first and last column are kept unconditionally (elif order);
           otherwise dropped iff gap_num > row_num / 2  (float compare == 2*gaps > R) */
        if (c == 0 || c == C - 1 || !(2 * gaps > R)) keep[nk++] = c;
    }
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < nk; ++k) out[(size_t)r * nk + k] = m[(size_t)r * C + keep[k]];
    *C_out = nk;
    free(keep);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* column composition with Python-dict insertion order (Util.py:9251-9266, 8900-8915)          */
typedef struct {
    int cnt[6];      /* per symbol */
    int8_t order[6]; /* symbols in first-seen order ('-' appended last if unseen) */
    int8_t nkeys;
} colmap;

static void build_colmap(const char *const *rows, int R, int C, colmap *cm) {
    for (int c = 0; c < C; ++c) {
        colmap *x = &cm[c];
        memset(x, 0, sizeof(*x));
        int seen = 0;
        for (int r = 0; r < R; ++r) {
            int s = sym_of(rows[r][c]);
            if (!(seen & (1 << s))) { seen |= 1 << s; x->order[x->nkeys++] = (int8_t)s; }
            x->cnt[s]++;
        }
        if (!(seen & (1 << 5))) x->order[x->nkeys++] = 5;
    }
}

/* calculate_window_homology  Util.py:8827-8885.  cols[] in iteration order. */
static int window_homology(const char *const *rows, int R, const int *cols, int ncols, double thr) {
    int *valid = (int *)malloc(sizeof(int) * (size_t)R);
    int nv = 0;
    for (int r = 0; r < R; ++r) {
        int g = 0;
        for (int k = 0; k < ncols; ++k) g += (rows[r][cols[k]] == '-');
        if (2 * g <= ncols) valid[nv++] = r;               /* gap_count <= len/2 (:8846) */
    }
    if (nv < 2) { free(valid); return -1; }                /* :8850 */
    double total = 0.0;
    int first_candidate = -1;
    for (int k = 0; k < ncols; ++k) {
        int cnt[6] = {0, 0, 0, 0, 0, 0};
        for (int v = 0; v < nv; ++v) cnt[sym_of(rows[valid[v]][cols[k]])]++;
        int mx = 0;
        for (int s = 0; s < 5; ++s) if (cnt[s] > mx) mx = cnt[s];
        double ratio = mx > 0 ? (double)mx / (double)nv : 0.0;       /* :8865-8868 */
        if (ratio >= thr - 0.1 && first_candidate == -1) first_candidate = cols[k];
        total += ratio;
    }
    free(valid);
    double avg = total / (double)ncols;
    return (avg >= thr) ? first_candidate : -1;
}

/* slide windows of `wsize` over the list of valid columns vc[0..n) (already in the order the
   reference holds them); evaluate columns range(first, last+1, step). */
static int slide_windows(const char *const *rows, int R, const int *vc, int n, int wsize, int step,
                         double thr) {
    int *cols = NULL;
    int res = -1;
    for (int i = 0; i + wsize <= n; ++i) {
        int first = vc[i], last = vc[i + wsize - 1];
        int ncols = 0;
        if (step > 0) {
            ncols = last + 1 - first; if (ncols < 0) ncols = 0;
            cols = (int *)realloc(cols, sizeof(int) * (size_t)(ncols + 1));
            for (int k = 0; k < ncols; ++k) cols[k] = first + k;
        } else {
            /* range(first, last + 1, -1): first, first-1, ..., last+2  (Util.py:9076, 9134) */
            ncols = first - (last + 1); if (ncols < 0) ncols = 0;
            cols = (int *)realloc(cols, sizeof(int) * (size_t)(ncols + 1));
            for (int k = 0; k < ncols; ++k) cols[k] = first - k;
        }
        res = window_homology(rows, R, cols, ncols, thr);
        if (res != -1) break;
    }
    free(cols);
    return res;
}

/* search_boundary_homo_v3  Util.py:8887-9143 */
int hbo_search_boundary_homo_v3(int valid_col_threshold, int pos, const char *const *rows, int R, int C,
                                int is_start, double thr, int int_win, int out_win, const void *cmv) {
    const colmap *cm = (const colmap *)cmv;
    const int search_len = 100;
    int *vc = (int *)malloc(sizeof(int) * (size_t)(C + 1));
    int n, cur_boundary, nb;
    if (is_start) {
        n = 0;
        for (int c = pos; n < search_len && 2 * c < C; ++c)            /* :8924 */
            if (cm[c].cnt[5] <= valid_col_threshold) vc[n++] = c;
        cur_boundary = pos;
        int w = n < int_win ? n : int_win;
        if (w < 10) cur_boundary = -1;                                  /* :8954-8956 */
        else cur_boundary = slide_windows(rows, R, vc, n, w, +1, thr);  /* :8958-8969 */
        n = 0;
        for (int c = cur_boundary; n < search_len && c >= 0; --c)       /* :8976 */
            if (cm[c].cnt[5] <= valid_col_threshold) vc[n++] = c;
        for (int a = 0, b = n - 1; a < b; ++a, --b) { int t = vc[a]; vc[a] = vc[b]; vc[b] = t; } /* :9003 */
        w = n < out_win ? n : out_win;
        if (w < 10) cur_boundary = -1;                                  /* :9006-9008 */
        else {
            nb = slide_windows(rows, R, vc, n, w, +1, thr);
            if (nb != -1) {
                if (nb < 10) cur_boundary = -1;                         /* :9020-9023 */
                else if (nb != cur_boundary) cur_boundary = nb;
            }
        }
    } else {
        n = 0;
        for (int c = pos; n < search_len && c < C; ++c)                 /* :9036 */
            if (cm[c].cnt[5] <= valid_col_threshold) vc[n++] = c;
        cur_boundary = pos;
        for (int a = 0, b = n - 1; a < b; ++a, --b) { int t = vc[a]; vc[a] = vc[b]; vc[b] = t; } /* :9065 */
        int w = n < out_win ? n : out_win;
        if (w < 10) cur_boundary = -1;                                  /* :9068-9070 */
        else {
            nb = slide_windows(rows, R, vc, n, w, -1, thr);
            if (nb != -1) {
                if (C - nb < 10) cur_boundary = -1;                     /* :9082-9085 */
                else if (nb != cur_boundary) cur_boundary = nb;
            }
        }
        n = 0;
        for (int c = cur_boundary; n < search_len && 2 * c >= C && c >= 0; --c)   /* :9096 */
            if (cm[c].cnt[5] <= valid_col_threshold) vc[n++] = c;
        w = n < int_win ? n : int_win;
        if (w < 10) cur_boundary = -1;                                  /* :9126-9128 */
        else cur_boundary = slide_windows(rows, R, vc, n, w, -1, thr);  /* :9130-9141 */
    }
    free(vc);
    return cur_boundary;
}

/* get_boundary_ungap_str  Util.py:2407-2429.  Returns length; out[] holds the string. */
static int ungap_str(const char *row, int C, int pos, int want, int right, char *out) {
    int n = 0;
    if (right) {
        for (int c = pos; n < want && c < C; ++c) if (row[c] != '-') out[n++] = row[c];
    } else {
        char tmp[32];
        for (int c = pos; n < want && c >= 0; --c) if (row[c] != '-') tmp[n++] = row[c];
        for (int k = 0; k < n; ++k) out[k] = tmp[n - 1 - k];
    }
    out[n] = 0;
    return n;
}

/* TSDsearch_v5  Util.py:2460-2492: returns 1 iff left_tsd_seq != '' */
int hbo_tsd_search_v5(const char *row, int C, int bstart, int bend, int plant) {
    static const int lens[9] = {11, 10, 9, 8, 6, 5, 4, 3, 2};
    char f5[8], f3[8], l5[8], l3[8], lt[16], rt[16];
    int nf5 = ungap_str(row, C, bstart, 5, 1, f5);
    int nf3 = ungap_str(row, C, bstart, 3, 1, f3);
    int nl5 = ungap_str(row, C, bend, 5, 0, l5);
    int nl3 = ungap_str(row, C, bend, 3, 0, l3);
    (void)nf5; (void)nf3; (void)nl5; (void)nl3;
    int found = 0;
    for (int i = 0; i < 9; ++i) {
        int k = lens[i];
        int nl = ungap_str(row, C, bstart - 1, k, 0, lt);
        int nr = ungap_str(row, C, bend + 1, k, 1, rt);
        if (nl != nr || nl != k) continue;
        if (memcmp(lt, rt, (size_t)k) == 0) {
            int ok;
            if (k != 2 && k != 3 && k != 4) ok = 1;
            else if (k == 4) ok = (memcmp(lt, "TTAA", 4) == 0);
            else if (k == 2)
                ok = (memcmp(lt, "TA", 2) == 0) ||
                     (plant == 0 && strcmp(f3, "CCC") == 0 && strcmp(l3, "GGG") == 0);
            else
                ok = (memcmp(lt, "TAA", 3) == 0) || (memcmp(lt, "TTA", 3) == 0) ||
                     (plant == 1 && ((strcmp(f5, "CACTA") == 0 && strcmp(l5, "TAGTG") == 0) ||
                                     (strcmp(f5, "CACTG") == 0 && strcmp(l5, "CAGTG") == 0)));
            if (ok) found = 1;
        } else if (k >= 8) {
            int mm = 0;
            for (int t = 0; t < k; ++t) mm += (lt[t] != rt[t]);
            if (mm <= 1) found = 1;
        }
    }
    return found;
}

/* consensus vote of one column  Util.py:9314-9339; returns symbol or -1 (column skipped) */
static int vote_col(const colmap *x, int half) {
    int mxc = 0, mxs = -1;
    for (int k = 0; k < x->nkeys; ++k) {
        int s = x->order[k];
        if (x->cnt[s] > mxc) { mxc = x->cnt[s]; mxs = s; }
    }
    if (mxc >= half) return (mxs != 5) ? mxs : -1;
    mxc = 0; mxs = -2;                       /* -2: max_base stays '' (nothing appended) */
    for (int k = 0; k < x->nkeys; ++k) {
        int s = x->order[k];
        if (s == 5) continue;
        if (x->cnt[s] > mxc) { mxc = x->cnt[s]; mxs = s; }
    }
    return mxs == -2 ? -1 : mxs;
}

static int starts_with(const char *s, int n, const char *p) {
    int lp = (int)strlen(p);
    return n >= lp && memcmp(s, p, (size_t)lp) == 0;
}
static int ends_with(const char *s, int n, const char *p) {
    int lp = (int)strlen(p);
    return n >= lp && memcmp(s + n - lp, p, (size_t)lp) == 0;
}

static void revcomp5(const char *s, int n, char *out) { /* getReverseSequence Util.py:1635 */
    for (int i = 0; i < n; ++i) {
        char b = s[n - 1 - i], r;
        switch (b) { case 'A': r = 'T'; break; case 'T': r = 'A'; break; case 'C': r = 'G'; break;
                     case 'G': r = 'C'; break; default: r = 'N'; }
        out[i] = r;
    }
}

/* anchors (Util.py:9151-9181): first row (top-down) where both 20-mers hit with <= 2 edits */
int hbo_find_anchor(const char *cand, int cand_len, const char *m, int R, int C,
                    int *align_start, int *align_end) {
    int alen = cand_len < 20 ? cand_len : 20;
    const char *first = cand;
    const char *last = cand + (cand_len - alen);
    char *u = (char *)malloc((size_t)C + 1);
    int *pos = (int *)malloc(sizeof(int) * (size_t)(C + 1));
    *align_start = -1; *align_end = -1;
    for (int r = 0; r < R; ++r) {
        const char *row = m + (size_t)r * C;
        int n = 0;
        for (int c = 0; c < C; ++c) if (row[c] != '-') { u[n] = row[c]; pos[n] = c; ++n; }
        int fs, fe, ls, le;
        int nf = hbo_find_near_matches(first, alen, u, n, 2, &fs, &fe, NULL, NULL);
        if (nf <= 0) continue;
        int nl = hbo_find_near_matches(last, alen, u, n, 2, NULL, NULL, &ls, &le);
        if (nl <= 0) continue;
        *align_start = pos[fs];
        *align_end = pos[le - 1];
        break;
    }
    free(u); free(pos);
    return (*align_start >= 0 && *align_end >= 0) ? 1 : 0;
}

/* judge_boundary_v5  Util.py:9145-9480 (TE_type == 'tir' branch; the only one
   is_TE_from_align_file :10427-10428 routes here) */
int hbo_judge_v5(const char *cand, int cand_len, const char *m, int R, int C, int plant,
                 int result_type_cons, hbo_call *out, char *cons) {
    memset(out, 0, sizeof(*out));
    out->homo_start = out->homo_end = -1;
    cons[0] = 0;
    int as, ae;
    if (!hbo_find_anchor(cand, cand_len, m, R, C, &as, &ae)) {
        out->info = HBO_INFO_NB; out->align_start = as; out->align_end = ae;  /* :9184-9187 */
        return 0;
    }
    out->align_start = as; out->align_end = ae;
    /* full-length member filter :9196-9217 */
    const char **rows = (const char **)malloc(sizeof(char *) * (size_t)(R + 1));
    int nr = 0;
    for (int r = 0; r < R; ++r) {
        if (nr > 100) break;
        const char *row = m + (size_t)r * C;
        int lo, hi, any1 = 0, any2 = 0;
        py_slice(C, as - 10 >= 0 ? as - 10 : 0, (long)as + 10, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { any1 = 1; break; }
        py_slice(C, (long)ae - 10, ae + 10 < C ? ae + 10 : C, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { any2 = 1; break; }
        if (any1 && any2) rows[nr++] = row;
    }
    int row_num = nr;
    if (row_num <= 1) {                                           /* :9222-9225 */
        out->info = HBO_INFO_FL1; out->copy_num = 1;
        free(rows); return 0;
    }
    int valid_col_threshold = row_num / 2;                        /* int(row_num/2) :9238 */
    double thr = row_num <= 2 ? 0.95 : (row_num <= 5 ? 0.9 : 0.7);  /* :9240-9245 */
    colmap *cm = (colmap *)malloc(sizeof(colmap) * (size_t)C);
    build_colmap(rows, row_num, C, cm);
    int vl = -1, vr = -1;                                          /* :9269-9294 */
    for (int c = 0; c < C; ++c) if (2 * cm[c].cnt[5] <= row_num) { vl = c; break; }
    if (vl != -1) for (int c = C - 1; c >= 0; --c) if (2 * cm[c].cnt[5] <= row_num) { vr = c; break; }
    if (!(vl != -1 && vr != -1 && vl < vr)) { vl = -1; vr = -1; }
    out->copy_num = row_num;
    int hs = hbo_search_boundary_homo_v3(valid_col_threshold, as, rows, row_num, C, 1, thr, 20, 10, cm);
    out->homo_start = hs;
    if (hs == -1) { free(cm); free(rows); return 0; }              /* :9300-9301 */
    int he = hbo_search_boundary_homo_v3(valid_col_threshold, ae, rows, row_num, C, 0, thr, 20, 10, cm);
    out->homo_end = he;
    if (he == -1) { free(cm); free(rows); return 0; }
    /* consensus :9312-9346 */
    char *model = (char *)malloc((size_t)C + 2);
    int ml = 0;
    for (int c = hs; c <= he; ++c) {
        if (result_type_cons) {
            int s = vote_col(&cm[c], row_num / 2);
            if (s >= 0) model[ml++] = SYM_CHR[s];
        } else if (rows[0][c] != '-') model[ml++] = rows[0][c];
    }
    model[ml] = 0;
    int final_lo = 0, final_hi = 0;
    if (!(hs <= vl || he >= vr)) {                                 /* :9353 */
        /* TIR trim options :9358-9379 */
        int fo[5], nfo = 0, eo[5], neo = 0;
        fo[nfo++] = 0; eo[neo++] = 0;
        if (starts_with(model, ml, "A")) fo[nfo++] = 1;
        if (starts_with(model, ml, "AA") || starts_with(model, ml, "TA")) fo[nfo++] = 2;
        if (starts_with(model, ml, "TAA") || starts_with(model, ml, "TTA")) fo[nfo++] = 3;
        if (starts_with(model, ml, "TTAA")) fo[nfo++] = 4;
        if (ends_with(model, ml, "T")) eo[neo++] = 1;
        if (ends_with(model, ml, "TT") || ends_with(model, ml, "TA")) eo[neo++] = 2;
        if (ends_with(model, ml, "TAA") || ends_with(model, ml, "TTA")) eo[neo++] = 3;
        if (ends_with(model, ml, "TTAA")) eo[neo++] = 4;
        int have = 0, best_ed = 0, best_tsd = 0, best_f = 0, best_e = 0;
        for (int a = 0; a < nfo; ++a) for (int b = 0; b < neo; ++b) {
            int bs = hs + fo[a], be = he - eo[b];
            int tsd = 0;
            for (int r = 0; r < R; ++r) {                          /* ALL rows :9391 */
                const char *row = m + (size_t)r * C;
                /* Python indexing raw_align_seq[i]: negative wraps, out of range raises */
                int ibs = bs < 0 ? bs + C : bs, ibe = be < 0 ? be + C : be;
                if (ibs < 0 || ibs >= C || ibe < 0 || ibe >= C) { tsd = -1; break; }
                if (row[ibs] == '-' || row[ibe] == '-') continue;
                if (hbo_tsd_search_v5(row, C, bs, be, plant)) ++tsd;
            }
            if (tsd > 0) {
                /* first_5bp = model[k:k+5]; end_5bp = model[len-5-k : len-k] (python slices) */
                int lo, hi, lo2, hi2;
                char f[8], e[8], rc[8];
                py_slice(ml, fo[a], fo[a] + 5, &lo, &hi);
                if (eo[b] == 0) py_slice(ml, -5, ml, &lo2, &hi2);
                else py_slice(ml, (long)ml - 5 - eo[b], (long)ml - eo[b], &lo2, &hi2);
                memcpy(f, model + lo, (size_t)(hi - lo));
                memcpy(e, model + lo2, (size_t)(hi2 - lo2));
                revcomp5(f, hi - lo, rc);
                int ed = hbo_levenshtein(rc, hi - lo, e, hi2 - lo2);
                /* sort key (ed, -tsd), stable, generation order :9403-9405 */
                if (!have || ed < best_ed || (ed == best_ed && tsd > best_tsd)) {
                    have = 1; best_ed = ed; best_tsd = tsd; best_f = fo[a]; best_e = eo[b];
                }
            }
        }
        if (have) {                                                /* :9406-9413 */
            if (best_e != 0) py_slice(ml, best_f, -(long)best_e, &final_lo, &final_hi);
            else py_slice(ml, best_f, ml, &final_lo, &final_hi);
            out->trim_lead = best_f; out->trim_trail = best_e;
        }
    }
    int fl = final_hi - final_lo;
    if (fl > 0) { memcpy(cons, model + final_lo, (size_t)fl); }
    cons[fl > 0 ? fl : 0] = 0;
    out->cons_len = fl > 0 ? fl : 0;
    out->is_te = fl > 0;                                           /* :9473-9476 */
    free(model); free(cm); free(rows);
    return 0;
}

/* ========================================================================================== */
/* Helitron: judge_boundary_v6 (Util.py:9821-10159) + search_boundary_homo_v4 (Util.py:8556-8824) */

static int collect_valid_cols(const colmap *cm, int vthr, int start, int dir, int lo, int hi, int *vc) {
    int n = 0;
    for (int c = start; n < 100 && c >= lo && c < hi; c += dir)
        if (cm[c].cnt[5] <= vthr) vc[n++] = c;
    return n;
}

static double col_max_ratio(const colmap *x, int row_num) {      /* stored max_homo_ratio :8607-8613 */
    int mx = 0;
    for (int s = 0; s < 5; ++s) if (x->cnt[s] > mx) mx = x->cnt[s];
    return mx > 0 ? (double)mx / (double)row_num : 0.0;
}

/* returns valid flag; *boundary gets the second tuple element */
int hbo_search_boundary_homo_v4(int vthr, int pos, const char *const *rows, int R, int C, int is_start,
                                double thr, double int_thr, double out_thr, int int_win, int out_win,
                                const void *cmv, int *boundary) {
    const colmap *cm = (const colmap *)cmv;
    int *vc = (int *)malloc(sizeof(int) * (size_t)(C + 1));
    int n, w, nb, cur;
    if (is_start) {
        n = (pos >= 0) ? collect_valid_cols(cm, vthr, pos, +1, 0, (C + 1) / 2, vc) : 0;      /* :8592 */
        cur = pos;
        w = n < int_win ? n : int_win;
        if (w < 10) cur = -1;                                                                  /* :8622-8624 */
        else {
            nb = slide_windows(rows, R, vc, n, w, +1, thr);
            if (nb != cur && nb != -1) cur = nb;                                               /* :8634-8637 */
        }
        n = (cur - 1 >= 0) ? collect_valid_cols(cm, vthr, cur - 1, -1, 0, C, vc) : 0;          /* :8639-8669 */
        for (int a = 0, b = n - 1; a < b; ++a, --b) { int t = vc[a]; vc[a] = vc[b]; vc[b] = t; }
        w = n < out_win ? n : out_win;
        if (w < 10) cur = -1;                                                                  /* :8676-8678 */
        else {
            nb = slide_windows(rows, R, vc, n, w, +1, thr);
            if (nb != -1) { if (nb < 10) cur = -1; else if (nb != cur) cur = nb; }             /* :8689-8697 */
        }
        *boundary = cur;
        free(vc);
        return 1;                                                                              /* :8699 */
    }
    n = collect_valid_cols(cm, vthr, pos + 1, +1, 0, C, vc);                                   /* :8704-8732 */
    w = n < out_win ? n : out_win;
    *boundary = -1;
    if (w < 10) { free(vc); return 0; }                                                        /* :8739-8741 */
    {
        double avg = 0.0;                                                                      /* window i == 0 only */
        for (int k = 0; k < w; ++k) avg += col_max_ratio(&cm[vc[k]], R);
        avg = avg / (double)w;
        if (avg >= out_thr) { free(vc); return 0; }       /* new_boundary_end != -1 -> both branches return False (:8758-8767) */
    }
    n = collect_valid_cols(cm, vthr, pos, -1, (C + 1) / 2, C, vc);                             /* :8769-8798 */
    w = n < int_win ? n : int_win;
    if (w < 10) { free(vc); return 0; }                                                        /* :8804-8806 */
    {
        double avg = 0.0;
        for (int k = 0; k < w; ++k) avg += col_max_ratio(&cm[vc[k]], R);
        avg = avg / (double)w;
        if (avg < int_thr) {
            int ne = vc[w - 1];
            if (ne != pos && ne != -1) { free(vc); return 0; }                                 /* :8817-8823 */
        }
    }
    *boundary = pos;
    free(vc);
    return 1;                                                                                  /* :8824 */
}

static int most_common(const int *arr, int n) {          /* most_common_element :9722-9730 */
    if (n <= 0) return -1;
    int best = arr[0], bc = 0;
    for (int i = 0; i < n; ++i) {
        int seen = 0;
        for (int j = 0; j < i; ++j) if (arr[j] == arr[i]) { seen = 1; break; }
        if (seen) continue;
        int c = 0;
        for (int j = i; j < n; ++j) c += (arr[j] == arr[i]);
        if (c > bc) { bc = c; best = arr[i]; }            /* ties: first inserted wins */
    }
    return best;
}

static int max_key(const colmap *x, int *count) {         /* first key with the largest count (:10082-10088) */
    int mxc = 0, mxs = -1;
    for (int k = 0; k < x->nkeys; ++k) { int s = x->order[k]; if (x->cnt[s] > mxc) { mxc = x->cnt[s]; mxs = s; } }
    *count = mxc;
    return mxs;
}

int hbo_judge_v6(const char *cand, int cand_len, const char *m, int R, int C, hbo_call *out, char *cons) {
    memset(out, 0, sizeof(*out));
    out->homo_start = out->homo_end = -1; out->align_start = out->align_end = -1;
    cons[0] = 0;
    int alen = cand_len < 20 ? cand_len : 20;
    const char *first = cand, *last = cand + (cand_len - alen);
    char *u = (char *)malloc((size_t)C + 1);
    int *pos = (int *)malloc(sizeof(int) * (size_t)(C + 1));
    int *as_l = (int *)malloc(sizeof(int) * (size_t)(R + 1)), *ae_l = (int *)malloc(sizeof(int) * (size_t)(R + 1));
    int na = 0;
    for (int r = 0; r < R; ++r) {                                                       /* :9837-9863, all rows */
        const char *row = m + (size_t)r * C;
        int n = 0;
        for (int c = 0; c < C; ++c) if (row[c] != '-') { u[n] = row[c]; pos[n] = c; ++n; }
        int fs, fe, ls, le;
        if (hbo_find_near_matches(first, alen, u, n, 2, &fs, &fe, NULL, NULL) <= 0) continue;
        if (hbo_find_near_matches(last, alen, u, n, 2, NULL, NULL, &ls, &le) <= 0) continue;
        as_l[na] = pos[fs]; ae_l[na] = pos[le - 1]; ++na;
    }
    int as = most_common(as_l, na), ae = most_common(ae_l, na);                         /* :9865-9866 */
    free(u); free(pos); free(as_l); free(ae_l);
    out->align_start = as; out->align_end = ae;
    if (as == -1 || ae == -1) { out->info = HBO_INFO_NB; return 0; }                    /* :9870-9873 */
    /* start / end member sets :9884-9936 */
    const char **srows = (const char **)malloc(sizeof(char *) * (size_t)(R + 1));
    const char **erows = (const char **)malloc(sizeof(char *) * (size_t)(R + 1));
    int ns = 0, ne = 0;
    for (int r = 0; r < R; ++r) {
        if (ns > 100) break;
        if (ne > 100) break;
        const char *row = m + (size_t)r * C;
        int lo, hi, any = 0;
        py_slice(C, as - 1 >= 0 ? as - 1 : 0, (long)as + 1, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { any = 1; break; }
        if (any) srows[ns++] = row;
        any = 0;
        py_slice(C, (long)ae - 1, ae + 1 < C ? ae + 1 : C, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { any = 1; break; }
        if (any) erows[ne++] = row;
    }
    if (ne <= 0) { free(srows); free(erows); return 0; }                                /* :9940-9941 */
    colmap *cm = (colmap *)malloc(sizeof(colmap) * (size_t)C);
    {
        double thr = ne <= 2 ? 0.95 : (ne <= 5 ? 0.9 : 0.7);
        double ithr = ne <= 2 ? 0.9 : (ne <= 5 ? 0.85 : 0.65);
        double othr = ne <= 2 ? 0.95 : (ne <= 5 ? 0.9 : 0.7);
        build_colmap(erows, ne, C, cm);
        int he;
        int ok = hbo_search_boundary_homo_v4(ne / 2, ae, erows, ne, C, 0, thr, ithr, othr, 20, 10, cm, &he);
        if (!ok) { free(cm); free(srows); free(erows); return 0; }                      /* :9976-9979 */
        out->homo_end = he;
    }
    int hs;
    {
        double thr = ns <= 2 ? 0.95 : (ns <= 5 ? 0.9 : 0.7);
        double ithr = ns <= 2 ? 0.9 : (ns <= 5 ? 0.85 : 0.65);
        double othr = ns <= 2 ? 0.95 : (ns <= 5 ? 0.9 : 0.7);
        build_colmap(srows, ns, C, cm);
        hbo_search_boundary_homo_v4(ns / 2, as, srows, ns, C, 1, thr, ithr, othr, 20, 10, cm, &hs);
        out->homo_start = hs;
    }
    const int he = out->homo_end;
    /* full-length members :10012-10038 (rows reuse srows storage) */
    int nr = 0;
    for (int r = 0; r < R; ++r) {
        if (nr > 100) break;
        const char *row = m + (size_t)r * C;
        int lo, hi, a1 = 0, a2 = 0;
        py_slice(C, hs - 1 >= 0 ? hs - 1 : 0, (long)hs + 1, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { a1 = 1; break; }
        py_slice(C, (long)he - 1, he + 1 < C ? he + 1 : C, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { a2 = 1; break; }
        if (a1 && a2) srows[nr++] = row;
    }
    if (nr <= 0) { free(cm); free(srows); free(erows); return 0; }
    const int row_num = nr;
    out->copy_num = row_num;
    build_colmap(srows, row_num, C, cm);
    /* consensus with 'N' for weak columns :10078-10095, +-1 bp extension :10097-10130 */
    char *model = (char *)malloc((size_t)C + 8);
    int ml = 0;
    {
        int col = hs - 1, cnt;                                                          /* left extension first in the result */
        char lext = 0;
        while (col >= 0) { int s = max_key(&cm[col], &cnt); if (cnt >= row_num / 2 && s != 5) { lext = SYM_CHR[s]; break; } --col; }
        if (lext) model[ml++] = lext;
        for (int c = hs; c <= he; ++c) {
            int s = max_key(&cm[c], &cnt);
            if (cnt >= row_num / 2) { if (s != 5) model[ml++] = SYM_CHR[s]; }
            else model[ml++] = 'N';
        }
        col = he + 1;
        while (col < C) { int s = max_key(&cm[col], &cnt); if (cnt >= row_num / 2 && s != 5) { model[ml++] = SYM_CHR[s]; break; } ++col; }
    }
    model[ml] = 0;
    /* motif search :10135-10150 */
    static const char *tails[4] = {"CTAGT", "CTAAT", "CTGGT", "CTGAT"};
    int f_lo, f_hi, l_lo, l_hi;
    py_slice(ml, 0, 10, &f_lo, &f_hi);
    py_slice(ml, -10, ml, &l_lo, &l_hi);
    int final_lo = 0, final_hi = 0;
    for (int t = 0; t < 4; ++t) {
        int end_index = -1;
        for (int k = l_hi - l_lo - 5; k >= 0; --k) if (memcmp(model + l_lo + k, tails[t], 5) == 0) { end_index = k; break; }
        if (end_index == -1) continue;
        int start_index = -1;
        for (int k = 0; k + 3 <= f_hi - f_lo; ++k) if (memcmp(model + f_lo + k, "ATC", 3) == 0) { start_index = k; break; }
        if (start_index != -1) {
            int cut = 10 - (end_index + 3) - 1;
            if (cut == 0) py_slice(ml, start_index + 1, ml, &final_lo, &final_hi);
            else py_slice(ml, start_index + 1, -(long)cut, &final_lo, &final_hi);
            break;
        }
    }
    int fl = final_hi - final_lo;
    if (fl > 0) memcpy(cons, model + final_lo, (size_t)fl);
    cons[fl > 0 ? fl : 0] = 0;
    out->cons_len = fl > 0 ? fl : 0;
    out->is_te = fl > 0;
    free(model); free(cm); free(srows); free(erows);
    return 0;
}

/* ========================================================================================== */
/* non-LTR: judge_boundary_v9 (Util.py:9483-9720), find_tail_polyA (:10832),                    */
/* find_longest_tandem_repeat_tail (:9732-9770)                                                 */

static int tail_polyA_end(const char *s, int n) {                 /* find_tail_polyA, min_length 6 */
    for (int i = n - 5 - 1 + 0; i >= 0; --i) {                    /* range(len-(6-1), -1, -1): i = n-5 .. 0 */
        (void)0;
    }
    for (int i = n - 5; i >= 0; --i) {
        if (i + 6 <= n && memcmp(s + i, "AAAAAA", 6) == 0) return i + 6;
    }
    return -1;
}

static int tandem_tail_end(const char *seq, int n) {              /* find_longest_tandem_repeat_tail */
    const int tail_length = 30, min_repeats = 4;
    const char *tail = n >= tail_length ? seq + (n - tail_length) : seq;
    int tl = n >= tail_length ? tail_length : n;
    int longest_end = -1, max_len = 0;
    for (int ul = 2; ul <= 6; ++ul) {
        for (int start = tl - ul * min_repeats; start >= 0; --start) {
            int count = 1;
            for (int i = 1; i < (tl - start) / ul; ++i) {
                if (memcmp(tail + start + i * ul, tail + start, (size_t)ul) != 0) break;
                ++count;
            }
            if (count >= min_repeats) {
                int total = ul * count;
                if (total > max_len) { max_len = total; longest_end = n - tl + start + total; }
            }
        }
    }
    return longest_end;
}

/* fuzzysearch.find_near_matches(TSD, kmer, max_l_dist=1) non-empty, |TSD| == |kmer| == k (Util.py:9650) */
static int near1(const char *tsd, const char *kmer, int k) {
    int mm = 0;
    for (int t = 0; t < k; ++t) mm += (tsd[t] != kmer[t]);
    if (mm <= 1) return 1;
    /* one deletion from TSD equals kmer[1:] or kmer[:-1] */
    for (int which = 0; which < 2; ++which) {
        const char *s = kmer + (which == 0 ? 1 : 0);          /* length k-1 */
        int a = 0;
        while (a < k - 1 && tsd[a] == s[a]) ++a;
        int ok = 1;
        for (int t = a; t < k - 1; ++t) if (tsd[t + 1] != s[t]) { ok = 0; break; }
        if (ok) return 1;
    }
    return 0;
}

int hbo_judge_v9(const char *cand, int cand_len, const char *m, int R, int C, hbo_call *out, char *cons) {
    memset(out, 0, sizeof(*out));
    out->homo_start = out->homo_end = -1;
    cons[0] = 0;
    int as, ae;
    if (!hbo_find_anchor(cand, cand_len, m, R, C, &as, &ae)) {                  /* :9522-9525 */
        out->info = HBO_INFO_NB; out->align_start = as; out->align_end = ae;
        return 0;
    }
    out->align_start = as; out->align_end = ae;
    const char **rows = (const char **)malloc(sizeof(char *) * (size_t)(R + 1));
    int nr = 0;
    for (int r = 0; r < R; ++r) {                                                /* :9535-9556 */
        if (nr > 100) break;
        const char *row = m + (size_t)r * C;
        int lo, hi, any1 = 0, any2 = 0;
        py_slice(C, as - 10 >= 0 ? as - 10 : 0, (long)as + 10, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { any1 = 1; break; }
        py_slice(C, (long)ae - 10, ae + 10 < C ? ae + 10 : C, &lo, &hi);
        for (int c = lo; c < hi; ++c) if (row[c] != '-') { any2 = 1; break; }
        if (any1 && any2) rows[nr++] = row;
    }
    int row_num = nr;
    if (row_num <= 1) { out->info = HBO_INFO_FL1; out->copy_num = row_num; free(rows); return 0; }   /* :9561-9564 */
    double thr = row_num <= 2 ? 0.95 : (row_num <= 5 ? 0.9 : 0.8);              /* :9579-9584 */
    colmap *cm = (colmap *)malloc(sizeof(colmap) * (size_t)C);
    build_colmap(rows, row_num, C, cm);
    out->copy_num = row_num;
    int hs = hbo_search_boundary_homo_v3(row_num / 2, as, rows, row_num, C, 1, thr, 20, 10, cm);
    out->homo_start = hs;
    if (hs == -1) { free(cm); free(rows); return 0; }
    int he = hbo_search_boundary_homo_v3(row_num / 2, ae, rows, row_num, C, 0, thr, 20, 10, cm);
    out->homo_end = he;
    if (he == -1) { free(cm); free(rows); return 0; }
    /* per-row polyA / tandem 3' end + TSD search over ALL rows :9604-9664 */
    int tsd_count = 0, have_first = 0;
    const int cur_boundary_start = hs;
    char *u = (char *)malloc((size_t)C + 1);
    int *n2g = (int *)malloc(sizeof(int) * (size_t)(C + 1));
    for (int r = 0; r < R; ++r) {
        const char *row = m + (size_t)r * C;
        int n = 0, end_5 = 0, h3 = 0;
        for (int c = 0; c < C; ++c) {
            if (c == cur_boundary_start) end_5 = n;
            if (c == he) h3 = n;
            if (row[c] != '-') { u[n] = row[c]; n2g[n] = c; ++n; }
        }
        int cl = h3 + 10; if (cl > n) cl = n;                                     /* cur_seq = align_seq[0:h3+10] */
        int end_3 = tail_polyA_end(u, cl);
        if (end_3 == -1) end_3 = tandem_tail_end(u, cl);
        int diff = end_3 - h3; if (diff < 0) diff = -diff;
        if (diff > 10) continue;                                                  /* :9632 */
        int found = 0;
        if (end_3 != -1 && end_5 != -1) {
            int left_pos = end_5 - 50; if (left_pos < 0) left_pos = 0;
            int sl = end_5 - left_pos;                                            /* len(subsequence) */
            for (int k = 20; k >= 8 && !found; --k) {
                int tl = n - end_3; if (tl > k) tl = k; if (tl < 0) tl = 0;      /* len(align_seq[end_3:end_3+k]) */
                if (tl != k) continue;
                for (int i = 0; i + k <= sl; ++i) {
                    if (near1(u + end_3, u + left_pos + i, k)) { end_5 = left_pos + i + k; found = 1; break; }
                }
            }
        }
        if (found) {
            ++tsd_count;
            if (!have_first) {
                int fbs = end_5 < end_3 ? end_5 : end_3, fbe = end_5 < end_3 ? end_3 : end_5;
                if (fbe > fbs) have_first = 1;                                    /* first_non_ltr_seq != '' */
                if (fbs < n) hs = n2g[fbs];                                       /* :9664 (KeyError if fbs == n) */
            }
        }
    }
    free(u); free(n2g);
    out->homo_start = hs;
    out->trim_lead = tsd_count;
    int ml = 0;
    if (tsd_count >= 5 || 2 * tsd_count > row_num) {                              /* :9668 */
        for (int c = hs; c <= he; ++c) {
            int s = vote_col(&cm[c], row_num / 2);
            if (s >= 0) cons[ml++] = SYM_CHR[s];
        }
    }
    cons[ml] = 0;
    out->cons_len = ml;
    out->is_te = !(ml == 0 || ml < 80);                                           /* :9713-9716 */
    free(cm); free(rows);
    return 0;
}
