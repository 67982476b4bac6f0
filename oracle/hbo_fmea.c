/*
 * TEST INFRASTRUCTURE (oracle) -- NOT part of the product path.
 *
 * Plain C restatement of the reference's FMEA interval chaining for the coarse-boundary stage:
 *   get_longest_repeats_v4   module/Util.py:4122-4400
 *   get_overlap_len          module/Util.py:4402-4405
 *   get_integer_pos          module/Util.py:4566-4569
 *   process_all_seqs / process_seq_group   module/Util.py:4529-4563
 * for ONE query segment (one `.final.out` file of process_blast_alignments, module/Util.py:4763-4770).
 * Pinned against tests/golden/fmea.json.gz, produced by the imported, unmodified reference function.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "hbo.h"

typedef struct { int32_t qs, qe, ss, se; int order; } frag_t;

static int cmp_fwd(const void *a, const void *b) {          /* key (s_start, s_end), stable */
    const frag_t *x = (const frag_t *)a, *y = (const frag_t *)b;
    if (x->ss != y->ss) return x->ss < y->ss ? -1 : 1;
    if (x->se != y->se) return x->se < y->se ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int cmp_rev(const void *a, const void *b) {          /* key (-s_start, -s_end), stable */
    const frag_t *x = (const frag_t *)a, *y = (const frag_t *)b;
    if (x->ss != y->ss) return x->ss > y->ss ? -1 : 1;
    if (x->se != y->se) return x->se > y->se ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int cmp_q(const void *a, const void *b) {            /* key (q_start, q_end), stable */
    const frag_t *x = (const frag_t *)a, *y = (const frag_t *)b;
    if (x->qs != y->qs) return x->qs < y->qs ? -1 : 1;
    if (x->qe != y->qe) return x->qe < y->qe ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int same_frag(const frag_t *a, const frag_t *b) {
    return a->qs == b->qs && a->qe == b->qe && a->ss == b->ss && a->se == b->se;
}

typedef struct { int32_t qs, qe, ss, se, subj; } longest_t;

/* chain one cluster (:4229-4319), append to lq */
static void chain_cluster(frag_t *c, int n, int skip_gap, int subj, longest_t **lq, int *nlq, int *caplq) {
    for (int i = 0; i < n; ++i) c[i].order = i;
    qsort(c, (size_t)n, sizeof(frag_t), cmp_q);                                  /* :4231 */
    uint8_t *visited = (uint8_t *)calloc((size_t)n + 1, 1);
    for (int i = 0; i < n; ++i) {
        if (visited[i]) continue;
        int32_t pqs = c[i].qs, pqe = c[i].qe, pss = c[i].ss, pse = c[i].se;
        for (int t = 0; t < n; ++t) if (same_frag(&c[t], &c[i])) visited[t] = 1;   /* dict keyed by the tuple */
        for (int j = i + 1; j < n; ++j) {
            if (visited[j]) continue;
            const int32_t cqs = c[j].qs, cqe = c[j].qe, css = c[j].ss, cse = c[j].se;
            if (cqe > pqe) {                                                      /* :4278 */
                if (pss < pse && css < cse) {
                    if (cse > pse) {
                        if (cqs - pqe < skip_gap && cqe > pqe && css - pse < skip_gap) {
                            pqe = cqe; pss = pss < css ? pss : css; pse = cse;
                            for (int t = 0; t < n; ++t) if (same_frag(&c[t], &c[j])) visited[t] = 1;
                        } else if (cqs - pqe >= skip_gap) break;
                    }
                } else if (pss > pse && css > cse) {
                    if (cse < pse) {
                        if (cqs - pqe < skip_gap && cqe > pqe && pse - css < skip_gap) {
                            pqe = cqe; pss = pss > css ? pss : css; pse = cse;
                            for (int t = 0; t < n; ++t) if (same_frag(&c[t], &c[j])) visited[t] = 1;
                        } else if (cqs - pqe >= skip_gap) break;
                    }
                }
            }
        }
        if (*nlq == *caplq) { *caplq = *caplq * 2 + 16; *lq = (longest_t *)realloc(*lq, sizeof(longest_t) * (size_t)*caplq); }
        longest_t *o = &(*lq)[(*nlq)++];
        o->qs = pqs; o->qe = pqe; o->ss = pss; o->se = pse; o->subj = subj;
    }
    free(visited);
}

/* tiny open-addressing set of (chr, start, end) */
typedef struct { int32_t chr; int64_t a, b; int used; } key_t;
typedef struct { key_t *t; size_t cap, n; } set_t;
static size_t key_hash(int32_t chr, int64_t a, int64_t b) {
    uint64_t h = (uint64_t)chr * 0x9E3779B97F4A7C15ull ^ (uint64_t)a * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)b * 0x165667B19E3779F9ull;
    h ^= h >> 29; return (size_t)h;
}
static int set_has(const set_t *s, int32_t chr, int64_t a, int64_t b) {
    size_t i = key_hash(chr, a, b) & (s->cap - 1);
    while (s->t[i].used) { if (s->t[i].chr == chr && s->t[i].a == a && s->t[i].b == b) return 1; i = (i + 1) & (s->cap - 1); }
    return 0;
}
static void set_add(set_t *s, int32_t chr, int64_t a, int64_t b) {
    if ((s->n + 1) * 2 > s->cap) {
        set_t ns; ns.cap = s->cap * 2; ns.n = 0; ns.t = (key_t *)calloc(ns.cap, sizeof(key_t));
        for (size_t i = 0; i < s->cap; ++i) if (s->t[i].used) set_add(&ns, s->t[i].chr, s->t[i].a, s->t[i].b);
        free(s->t); *s = ns;
    }
    size_t i = key_hash(chr, a, b) & (s->cap - 1);
    while (s->t[i].used) { if (s->t[i].chr == chr && s->t[i].a == a && s->t[i].b == b) return; i = (i + 1) & (s->cap - 1); }
    s->t[i].chr = chr; s->t[i].a = a; s->t[i].b = b; s->t[i].used = 1; s->n++;
}

typedef struct { int64_t a, b; int order; } iv_t;
static int cmp_len_desc(const void *x, const void *y) {      /* sort by length desc, stable */
    const iv_t *p = (const iv_t *)x, *q = (const iv_t *)y;
    int64_t lp = p->b - p->a, lq = q->b - q->a;
    if (lp != lq) return lp > lq ? -1 : 1;
    return (p->order > q->order) - (p->order < q->order);
}

/* chr_pos_candidates is created once per BLAST result FILE (:4155), i.e. it is shared by all the query records of one split
   file (split_and_store_sequences :4987: records are packed into a file until it holds >= 1 000 000 bases, so a contig's
   short tail segment shares a file with the records that follow it).  `shared` carries that dict from one query of a file to
   the next; NULL = a file with this single query. */
void *hbo_fmea_set_new(void) {
    set_t *s = (set_t *)malloc(sizeof(set_t));
    s->cap = 1024; s->n = 0; s->t = (key_t *)calloc(s->cap, sizeof(key_t));
    return s;
}
void hbo_fmea_set_free(void *p) { if (p) { free(((set_t *)p)->t); free(p); } }

int hbo_fmea_shared(const hbo_hsp *h, int n, int q_chr, int64_t q_seg_start, int self_subj,
                    const int32_t *subj_chr, const int64_t *subj_seg_start, int n_subj,
                    int skip_gap, int max_len, void *shared, hbo_interval *out, int cap);

int hbo_fmea(const hbo_hsp *h, int n, int q_chr, int64_t q_seg_start, int self_subj,
             const int32_t *subj_chr, const int64_t *subj_seg_start, int n_subj,
             int skip_gap, int max_len, hbo_interval *out, int cap) {
    return hbo_fmea_shared(h, n, q_chr, q_seg_start, self_subj, subj_chr, subj_seg_start, n_subj, skip_gap, max_len, NULL, out, cap);
}

int hbo_fmea_shared(const hbo_hsp *h, int n, int q_chr, int64_t q_seg_start, int self_subj,
                    const int32_t *subj_chr, const int64_t *subj_seg_start, int n_subj,
                    int skip_gap, int max_len, void *shared, hbo_interval *out, int cap) {
    /* subjects in first-appearance order (:4140-4147) */
    int *order = (int *)malloc(sizeof(int) * (size_t)(n_subj + 1));
    int *seen = (int *)calloc((size_t)n_subj + 1, sizeof(int));
    int n_order = 0;
    for (int i = 0; i < n; ++i) {
        if (h[i].subj == self_subj && h[i].qs == h[i].ss && h[i].qe == h[i].se) continue;   /* :4138 */
        if (!seen[h[i].subj]) { seen[h[i].subj] = 1; order[n_order++] = h[i].subj; }
    }
    longest_t *lq = NULL; int nlq = 0, caplq = 0;
    frag_t *fw = (frag_t *)malloc(sizeof(frag_t) * (size_t)(n + 1)), *rv = (frag_t *)malloc(sizeof(frag_t) * (size_t)(n + 1));
    frag_t *cl = (frag_t *)malloc(sizeof(frag_t) * (size_t)(n + 1));
    for (int oi = 0; oi < n_order; ++oi) {
        const int s = order[oi];
        int nf = 0, nr = 0;
        for (int i = 0; i < n; ++i) {
            if (h[i].subj != s) continue;
            if (s == self_subj && h[i].qs == h[i].ss && h[i].qe == h[i].se) continue;
            frag_t f = {h[i].qs, h[i].qe, h[i].ss, h[i].se, 0};
            if (f.ss > f.se) { f.order = nr; rv[nr++] = f; } else { f.order = nf; fw[nf++] = f; }      /* :4168-4172 */
        }
        qsort(fw, (size_t)nf, sizeof(frag_t), cmp_fwd);
        qsort(rv, (size_t)nr, sizeof(frag_t), cmp_rev);
        for (int pass = 0; pass < 2; ++pass) {                       /* forward clusters, then reverse (:4178-4227) */
            frag_t *p = pass ? rv : fw; const int np = pass ? nr : nf;
            int nc = 0;
            for (int k = 0; k < np; ++k) {
                int closed = 0;
                if (k > 0) {
                    for (int e = nc - 1; e >= 0; --e) {
                        const int gap = pass ? (cl[e].se - p[k].ss) : (p[k].ss - cl[e].se);
                        if (gap < skip_gap && p[k].qe > cl[e].qe) { closed = 1; break; }
                    }
                }
                if (k == 0 || closed) cl[nc++] = p[k];
                else { chain_cluster(cl, nc, skip_gap, s, &lq, &nlq, &caplq); nc = 0; cl[nc++] = p[k]; }
            }
            if (nc > 0) chain_cluster(cl, nc, skip_gap, s, &lq, &nlq, &caplq);
        }
    }
    /* position de-duplication (:4330-4390) */
    set_t own; own.cap = 1024; own.n = 0; own.t = shared ? NULL : (key_t *)calloc(own.cap, sizeof(key_t));
    set_t *setp = shared ? (set_t *)shared : &own;
#define set (*setp)
    iv_t *names = (iv_t *)malloc(sizeof(iv_t) * (size_t)(nlq + 1));
    int nn = 0;
    for (int r = 0; r < nlq; ++r) {
        const int32_t schr = subj_chr[lq[r].subj];
        const int64_t sa = subj_seg_start[lq[r].subj] + (lq[r].ss - 1), sb = subj_seg_start[lq[r].subj] + lq[r].se;
        const int64_t qa = q_seg_start + (lq[r].qs - 1), qb = q_seg_start + lq[r].qe;
        const int64_t sa1 = (sa / 10) * 10, sa2 = sa1 + 10, sb1 = (sb / 10) * 10, sb2 = sb1 + 10;
        const int64_t qa1 = (qa / 10) * 10, qa2 = qa1 + 10, qb1 = (qb / 10) * 10, qb2 = qb1 + 10;
        const int64_t qlen = lq[r].qe - (lq[r].qs - 1) >= 0 ? lq[r].qe - (lq[r].qs - 1) : -(lq[r].qe - (lq[r].qs - 1));
        const int hit = set_has(&set, schr, sa1, sb1) || set_has(&set, schr, sa1, sb2) || set_has(&set, schr, sa2, sb1) ||
                        set_has(&set, schr, sa2, sb2) || set_has(&set, q_chr, qa1, qb1) || set_has(&set, q_chr, qa1, qb2) ||
                        set_has(&set, q_chr, qa2, qb1) || set_has(&set, q_chr, qa2, qb2);
        if (!hit && qlen >= 80 && qlen < max_len) { names[nn].a = qa; names[nn].b = qb; names[nn].order = nn; ++nn; }
        set_add(&set, schr, sa1, sb1); set_add(&set, schr, sa1, sb2); set_add(&set, schr, sa2, sb1); set_add(&set, schr, sa2, sb2);
        set_add(&set, q_chr, qa1, qb1); set_add(&set, q_chr, qa1, qb2); set_add(&set, q_chr, qa2, qb1); set_add(&set, q_chr, qa2, qb2);
    }
    /* process_all_seqs on one chromosome (all names carry the query's chromosome) (:4529-4563) */
    qsort(names, (size_t)nn, sizeof(iv_t), cmp_len_desc);
    uint8_t *keep = (uint8_t *)malloc((size_t)nn + 1);
    memset(keep, 1, (size_t)nn + 1);
    for (int i = 0; i < nn; ++i) {
        if (!keep[i]) continue;
        for (int j = i + 1; j < nn; ++j) {
            if (!keep[j]) continue;
            int64_t lo = names[i].a > names[j].a ? names[i].a : names[j].a;
            int64_t hi = names[i].b < names[j].b ? names[i].b : names[j].b;
            int64_t ov = hi - lo > 0 ? hi - lo : 0;
            if ((double)ov / (double)(names[j].b - names[j].a) >= 0.95) keep[j] = 0;
        }
    }
    int no = 0;
    for (int i = 0; i < nn; ++i) {
        if (!keep[i]) continue;
        /* longest_repeats is a dict keyed by the name: a repeated name keeps its first position */
        int dup = 0;
        for (int t = 0; t < no; ++t) if (out[t].start == names[i].a && out[t].end == names[i].b) { dup = 1; break; }
        if (dup) continue;
        if (no < cap) { out[no].chr = q_chr; out[no].start = names[i].a; out[no].end = names[i].b; }
        ++no;
    }
#undef set
    free(keep); free(names); free(own.t); free(lq); free(fw); free(rv); free(cl); free(order); free(seen);
    return no;
}

/* ========================================================================================== */
/* get_query_copies  module/Util.py:6828-7030 -- BLAST-HSP copy chaining of ONE query           */
typedef struct { int32_t qs, qe, ss, se; double ident; int order; } cfrag_t;

static int ccmp_fwd(const void *a, const void *b) {
    const cfrag_t *x = (const cfrag_t *)a, *y = (const cfrag_t *)b;
    if (x->ss != y->ss) return x->ss < y->ss ? -1 : 1;
    if (x->se != y->se) return x->se < y->se ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int ccmp_rev(const void *a, const void *b) {
    const cfrag_t *x = (const cfrag_t *)a, *y = (const cfrag_t *)b;
    if (x->ss != y->ss) return x->ss > y->ss ? -1 : 1;
    if (x->se != y->se) return x->se > y->se ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int ccmp_q(const void *a, const void *b) {
    const cfrag_t *x = (const cfrag_t *)a, *y = (const cfrag_t *)b;
    if (x->qs != y->qs) return x->qs < y->qs ? -1 : 1;
    if (x->qe != y->qe) return x->qe < y->qe ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int csame(const cfrag_t *a, const cfrag_t *b) {
    return a->qs == b->qs && a->qe == b->qe && a->ss == b->ss && a->se == b->se && a->ident == b->ident;
}
typedef struct { int32_t qs, qe, qlen, ss, se, slen, subj; int order; } clong_t;
static int ccmp_len(const void *a, const void *b) {            /* sort(key=-len), stable */
    const clong_t *x = (const clong_t *)a, *y = (const clong_t *)b;
    if (x->qlen != y->qlen) return x->qlen > y->qlen ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}

static void copies_cluster(cfrag_t *c, int n, int qthr, int sthr, int subj, clong_t **lq, int *nlq, int *cap) {
    for (int i = 0; i < n; ++i) c[i].order = i;
    qsort(c, (size_t)n, sizeof(cfrag_t), ccmp_q);                                   /* :6899 */
    uint8_t *vis = (uint8_t *)calloc((size_t)n + 1, 1);
    int32_t bqs = -1, bqe = -1, bql = -1, bss = -1, bse = -1, bsl = -1;
    for (int i = 0; i < n; ++i) {
        if (vis[i]) continue;
        int32_t lqs = c[i].qs, lqe = c[i].qe, lss = c[i].ss, lse = c[i].se;
        int32_t cur_len = c[i].qe - c[i].qs + 1;                                     /* :6919 */
        for (int t = 0; t < n; ++t) if (csame(&c[t], &c[i])) vis[t] = 1;
        for (int j = i + 1; j < n; ++j) {
            if (vis[j]) continue;
            const cfrag_t *e = &c[j];
            if (e->qe > lqe) {
                if (lss < lse && e->ss < e->se) {
                    if (e->se > lse) {
                        if (e->qs - lqe < qthr && e->ss - lse < sthr) {
                            lqe = e->qe; lss = lss < e->ss ? lss : e->ss; lse = e->se;
                            cur_len = lqe - lqs;                                    /* :6951 (no +1) */
                            for (int t = 0; t < n; ++t) if (csame(&c[t], e)) vis[t] = 1;
                        } else if (e->qs - lqe >= qthr) break;
                    }
                } else if (lss > lse && e->ss > e->se) {
                    if (e->se < lse) {
                        if (e->qs - lqe < qthr && lse - e->ss < sthr) {
                            lqe = e->qe; lss = lss > e->ss ? lss : e->ss; lse = e->se;
                            cur_len = lqe - lqs;
                            for (int t = 0; t < n; ++t) if (csame(&c[t], e)) vis[t] = 1;
                        } else if (e->qs - lqe >= qthr) break;
                    }
                }
            }
        }
        if (cur_len > bql) {                                                        /* :6979-6989 */
            bqs = lqs; bqe = lqe; bql = cur_len; bss = lss; bse = lse;
            bsl = (lse > lss ? lse - lss : lss - lse) + 1;
        }
    }
    free(vis);
    if (bql != -1) {
        if (*nlq == *cap) { *cap = *cap * 2 + 16; *lq = (clong_t *)realloc(*lq, sizeof(clong_t) * (size_t)*cap); }
        clong_t *o = &(*lq)[*nlq];
        o->qs = bqs; o->qe = bqe; o->qlen = bql; o->ss = bss; o->se = bse; o->slen = bsl; o->subj = subj; o->order = *nlq;
        ++*nlq;
    }
}

int hbo_query_copies(const hbo_hsp *h, const double *ident, int n, int query_len, const int32_t *subj_len, int n_subj,
                     double query_coverage, double subject_coverage, int qthr, int sthr, int max_copy_num,
                     hbo_copy_out *out, int cap) {
    int *order = (int *)malloc(sizeof(int) * (size_t)(n_subj + 1));
    int *seen = (int *)calloc((size_t)n_subj + 1, sizeof(int));
    int n_order = 0;
    for (int i = 0; i < n; ++i) if (!seen[h[i].subj]) { seen[h[i].subj] = 1; order[n_order++] = h[i].subj; }
    clong_t *lq = NULL; int nlq = 0, caplq = 0;
    cfrag_t *fw = (cfrag_t *)malloc(sizeof(cfrag_t) * (size_t)(n + 1)), *rv = (cfrag_t *)malloc(sizeof(cfrag_t) * (size_t)(n + 1));
    cfrag_t *cl = (cfrag_t *)malloc(sizeof(cfrag_t) * (size_t)(n + 1));
    for (int oi = 0; oi < n_order; ++oi) {
        const int s = order[oi];
        int nf = 0, nr = 0;
        for (int i = 0; i < n; ++i) {
            if (h[i].subj != s) continue;
            cfrag_t f = {h[i].qs, h[i].qe, h[i].ss, h[i].se, ident[i], 0};
            if (f.ss > f.se) { f.order = nr; rv[nr++] = f; } else { f.order = nf; fw[nf++] = f; }
        }
        qsort(fw, (size_t)nf, sizeof(cfrag_t), ccmp_fwd);
        qsort(rv, (size_t)nr, sizeof(cfrag_t), ccmp_rev);
        for (int pass = 0; pass < 2; ++pass) {
            cfrag_t *p = pass ? rv : fw; const int np = pass ? nr : nf;
            int nc = 0;
            for (int k = 0; k < np; ++k) {
                int closed = 0;
                if (k > 0)
                    for (int e = nc - 1; e >= 0; --e) {
                        const int gap = pass ? (cl[e].se - p[k].ss) : (p[k].ss - cl[e].se);
                        if (gap < sthr && p[k].qe > cl[e].qe) { closed = 1; break; }
                    }
                if (k == 0 || closed) cl[nc++] = p[k];
                else { copies_cluster(cl, nc, qthr, sthr, s, &lq, &nlq, &caplq); nc = 0; cl[nc++] = p[k]; }
            }
            if (nc > 0) copies_cluster(cl, nc, qthr, sthr, s, &lq, &nlq, &caplq);
        }
    }
    qsort(lq, (size_t)nlq, sizeof(clong_t), ccmp_len);                              /* :6997 */
    int no = 0;
    for (int r = 0; r < nlq; ++r) {
        if (no > max_copy_num) break;                                               /* :7002 */
        int32_t a = lq[r].ss, b = lq[r].se, strand = 0;
        if (a > b) { int32_t t = a; a = b; b = t; strand = 1; }
        int ok = ((double)lq[r].qlen / (double)query_len >= query_coverage);
        if (subject_coverage > 0) ok = ok && ((double)lq[r].slen / (double)subj_len[lq[r].subj] >= subject_coverage);
        if (ok) for (int t = 0; t < no && t < cap; ++t) if (out[t].subj == lq[r].subj && out[t].start == a && out[t].end == b) { ok = 0; break; }
        if (!ok) continue;
        if (no < cap) { out[no].subj = lq[r].subj; out[no].start = a; out[no].end = b; out[no].qlen = lq[r].qlen; out[no].strand = strand; }
        ++no;
    }
    free(lq); free(fw); free(rv); free(cl); free(order); free(seen);
    return no;
}

/* ---------------------------------------------------------------------------------------------- */
/* get_full_length_copies_from_blastn_v1 (reference module/Util.py:5907-6128) for ONE query, as mask_genome_intactTE
   drives it (Util.py:6389-6431 -> generate_full_length_out_v1 :6288-6316: search_struct = False): the HSPs of the query
   are chained per subject into candidate copies; copies covering >= full_length_threshold of the query are returned as
   (subject, start, end, strand) with start 0-based and end exclusive -- the 'chr:start-end' key of :6107, i.e. the slice
   the caller fills with N (:6414-6420).  Keys are unique, in first-insertion order (Python dict).  Pinned against
   tests/golden/full_length_copies.json.gz (outputs of the imported reference). */
typedef struct { int32_t qs, qe, ss, se; int order; } mfrag_t;

static int mcmp_fwd(const void *a, const void *b) {           /* key (s_start, s_end), stable (:5965) */
    const mfrag_t *x = (const mfrag_t *)a, *y = (const mfrag_t *)b;
    if (x->ss != y->ss) return x->ss < y->ss ? -1 : 1;
    if (x->se != y->se) return x->se < y->se ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int mcmp_rev(const void *a, const void *b) {           /* key (-s_start, -s_end), stable (:5966) */
    const mfrag_t *x = (const mfrag_t *)a, *y = (const mfrag_t *)b;
    if (x->ss != y->ss) return x->ss > y->ss ? -1 : 1;
    if (x->se != y->se) return x->se > y->se ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int mcmp_q(const void *a, const void *b) {             /* key (q_start, q_end), stable (:6023) */
    const mfrag_t *x = (const mfrag_t *)a, *y = (const mfrag_t *)b;
    if (x->qs != y->qs) return x->qs < y->qs ? -1 : 1;
    if (x->qe != y->qe) return x->qe < y->qe ? -1 : 1;
    return (x->order > y->order) - (x->order < y->order);
}
static int mfrag_same(const mfrag_t *a, const mfrag_t *b) { return a->qs == b->qs && a->qe == b->qe && a->ss == b->ss && a->se == b->se; }

int hbo_full_length_copies_v1(const hbo_hsp *h, int n, int query_len, double full_length_threshold, hbo_copy_out *out, int cap) {
    const double skip_gap = (double)query_len * full_length_threshold;                    /* :5947 */
    int n_subj = 0;
    for (int i = 0; i < n; ++i) if (h[i].subj + 1 > n_subj) n_subj = h[i].subj + 1;
    int *order = (int *)malloc(sizeof(int) * (size_t)(n_subj + 1));
    int *seen = (int *)calloc((size_t)n_subj + 1, sizeof(int));
    int n_order = 0;
    for (int i = 0; i < n; ++i) if (!seen[h[i].subj]) { seen[h[i].subj] = 1; order[n_order++] = h[i].subj; }
    mfrag_t *fw = (mfrag_t *)malloc(sizeof(mfrag_t) * (size_t)(n + 1)), *rv = (mfrag_t *)malloc(sizeof(mfrag_t) * (size_t)(n + 1));
    mfrag_t *cl = (mfrag_t *)malloc(sizeof(mfrag_t) * (size_t)(n + 1));
    char *visited = (char *)malloc((size_t)n + 1);
    int n_out = 0;
    for (int oi = 0; oi < n_order; ++oi) {
        const int s = order[oi];
        int nf = 0, nr = 0;
        for (int i = 0; i < n; ++i) {
            if (h[i].subj != s) continue;
            mfrag_t f = {h[i].qs, h[i].qe, h[i].ss, h[i].se, 0};
            if (f.ss > f.se) { f.order = nr; rv[nr++] = f; } else { f.order = nf; fw[nf++] = f; }   /* :5960-5964 */
        }
        qsort(fw, (size_t)nf, sizeof(mfrag_t), mcmp_fwd);
        qsort(rv, (size_t)nr, sizeof(mfrag_t), mcmp_rev);
        for (int pass = 0; pass < 2; ++pass) {
            const mfrag_t *p = pass ? rv : fw; const int np = pass ? nr : nf;
            int k = 0;
            while (k < np) {
                /* one cluster: consecutive fragments each "closed" to some earlier member (:5970-6018) */
                int nc = 0;
                cl[nc++] = p[k++];
                for (; k < np; ++k) {
                    int closed = 0;
                    for (int e = nc - 1; e >= 0; --e) {
                        const int gap = pass ? (cl[e].se - p[k].ss) : (p[k].ss - cl[e].se);
                        if ((double)gap < skip_gap && p[k].qe > cl[e].qe) { closed = 1; break; }
                    }
                    if (!closed) break;
                    cl[nc++] = p[k];
                }
                for (int i = 0; i < nc; ++i) cl[i].order = i;
                qsort(cl, (size_t)nc, sizeof(mfrag_t), mcmp_q);
                memset(visited, 0, (size_t)nc);
                for (int i = 0; i < nc; ++i) {
                    /* visited_frag is keyed by the tuple's VALUE: an identical tuple seen earlier counts (:6029) */
                    int vis = visited[i];
                    for (int z = 0; z < nc && !vis; ++z) if (visited[z] && mfrag_same(&cl[z], &cl[i])) vis = 1;
                    if (vis) continue;
                    int32_t pqs = cl[i].qs, pqe = cl[i].qe, pss = cl[i].ss, pse = cl[i].se;
                    int32_t longest = pqe - pqs; if (longest < 0) longest = -longest;                  /* :6038, 6040 */
                    visited[i] = 1;
                    for (int j = i + 1; j < nc; ++j) {
                        int vj = visited[j];
                        for (int z = 0; z < nc && !vj; ++z) if (visited[z] && mfrag_same(&cl[z], &cl[j])) vj = 1;
                        if (vj) continue;
                        const int32_t cqs = cl[j].qs, cqe = cl[j].qe, css = cl[j].ss, cse = cl[j].se;
                        if (!(cqe > pqe)) continue;                                                      /* :6058 */
                        if (pss < pse && css < cse) {
                            if (cse > pse) {
                                if ((double)(cqs - pqe) < skip_gap && cqe > pqe && (double)(css - pse) < skip_gap) {
                                    pqe = cqe; if (css < pss) pss = css; pse = cse; longest = pqe - pqs; visited[j] = 1;
                                } else if ((double)(cqs - pqe) >= skip_gap) break;
                            }
                        } else if (pss > pse && css > cse) {
                            if (cse < pse) {
                                if ((double)(cqs - pqe) < skip_gap && cqe > pqe && (double)(pse - css) < skip_gap) {
                                    pqe = cqe; if (css > pss) pss = css; pse = cse; longest = pqe - pqs; visited[j] = 1;
                                } else if ((double)(cqs - pqe) >= skip_gap) break;
                            }
                        }
                    }
                    /* :6096-6125 */
                    if ((double)longest < full_length_threshold * (double)query_len) continue;
                    int32_t start, end, strand;
                    if (pss > pse) { strand = 1; start = pse - 1; end = pss; } else { strand = 0; start = pss - 1; end = pse; }
                    if (!((double)longest / (double)query_len >= full_length_threshold)) continue;
                    int dup = 0;
                    for (int z = 0; z < n_out; ++z) if (out[z].subj == s && out[z].start == start && out[z].end == end) { dup = 1; out[z].strand = strand; break; }
                    if (dup) continue;
                    if (n_out < cap) { out[n_out].subj = s; out[n_out].start = start; out[n_out].end = end; out[n_out].qlen = longest; out[n_out].strand = strand; }
                    ++n_out;
                }
            }
        }
    }
    free(visited); free(cl); free(fw); free(rv); free(order); free(seen);
    return n_out;
}
