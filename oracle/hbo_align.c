/*
 * TEST INFRASTRUCTURE (oracle) -- NOT part of the product path.
 *
 * Plain scalar C restatement of the HB-align specification: the
 * deterministic seed -> cluster -> banded affine Smith-Waterman -> candidate-anchored star MSA
 * procedure that this repository uses where the reference shells out to third-party aligners
 * that are absent from /root/reference and from this image:
 *     minimap2 2.28  (module/Util.py:7941, 7952-7961)  full-length copy search
 *     mafft          (module/Util.py:10416)            per-candidate MSA
 * The aligners' heuristics cannot be reproduced ("parity unpinned" for their raw output); what IS
 * taken from the reference are the selection rules around them, each cited below:
 *     >= 0.95 query / target coverage            module/Util.py:8003-8022
 *     1-based inclusive copy coords, strand      module/Util.py:8015-8025
 *     +-flanking_len extension, off-contig skip, <100 bp skip, revcomp of '-' copies
 *                                                module/Util.py:8095-8115
 *     100 longest members                        module/Util.py:10409-10412, tools/ready_for_MSA.sh:63-73
 * The CUDA kernels (hite_b200/csrc/hb_index.cu, hb_seed.cu, hb_extend.cu, hb_msa.cu) must
 * reproduce every array produced here bit for bit.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hbo.h"

#define PAD 64
#define INVALID_H 0xFFFFFFFFu

struct hbo_genome {
    uint8_t *code;     /* TOT entries: 0..3 ACGT, 4 N (pads are N) */
    uint64_t tot;
    uint32_t n_contig;
    uint64_t *goff;    /* n_contig global offsets */
    uint64_t *glen;
    int32_t *crank;    /* rank of every contig by (name + ':') bytewise, or NULL: see hbo_genome_set_tie_rule */
    int tie_rule;      /* S9: 1 = 100 longest, ties by ready_for_MSA.sh's name order (default), 0 = ties to the earlier-ranked copy,
                          2 = the FIRST max_rows copies (FiLTR's extract_copies, HB_PRM_ROWS_FIRST) */
};

static inline uint8_t code_of(uint8_t c) {
    switch (c & 0xDF) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}

hbo_genome *hbo_genome_create(const uint8_t *ascii, const uint64_t *off, uint32_t n) {
    hbo_genome *g = (hbo_genome *)calloc(1, sizeof(*g));
    g->n_contig = n;
    g->tie_rule = 1;
    g->goff = (uint64_t *)malloc(sizeof(uint64_t) * (n + 1));
    g->glen = (uint64_t *)malloc(sizeof(uint64_t) * (n + 1));
    uint64_t p = PAD;
    for (uint32_t c = 0; c < n; ++c) { g->goff[c] = p; g->glen[c] = off[c + 1] - off[c]; p += g->glen[c] + PAD; }
    g->tot = (p + 63) & ~(uint64_t)63;
    g->code = (uint8_t *)malloc(g->tot);
    memset(g->code, 4, g->tot);
    for (uint32_t c = 0; c < n; ++c)
        for (uint64_t i = 0; i < g->glen[c]; ++i) g->code[g->goff[c] + i] = code_of(ascii[off[c] + i]);
    return g;
}
void hbo_genome_free(hbo_genome *g) { if (g) { free(g->code); free(g->goff); free(g->glen); free(g->crank); free(g); } }

/* S9 tie rule of the reference's tools/ready_for_MSA.sh (the default; copy_select_kernel follows it): `sort -nk 2 -r` on the .fai lines breaks a length tie by a reversed whole-line comparison, i.e. the member
   whose name "contig:start1-end1(strand)" is LARGER bytewise is the "longer" one (C / POSIX locale; pinned by running the real
   script, tests/test_flank_loop_dropin_cpu.py).  contig_rank[c] = rank of contig c by (name + ':') bytewise ascending. */
void hbo_genome_set_tie_rule(hbo_genome *g, int tie_rule, const int32_t *contig_rank) {
    g->tie_rule = tie_rule;
    free(g->crank); g->crank = NULL;
    if (contig_rank) {
        g->crank = (int32_t *)malloc(sizeof(int32_t) * (g->n_contig + 1));
        memcpy(g->crank, contig_rank, sizeof(int32_t) * g->n_contig);
    }
}
/* decimal strings of two non-negative numbers compared bytewise, a shorter string that is a prefix being smaller (what follows
   it in the name, '-' or '(', sorts below every digit) */
static int dec_lex_cmp(int64_t a, int64_t b) {
    char sa[24], sb[24];
    snprintf(sa, sizeof(sa), "%lld", (long long)a);
    snprintf(sb, sizeof(sb), "%lld", (long long)b);
    return strcmp(sa, sb);
}
/* is copy y's member name bytewise larger than copy x's ? */
static int name_greater(const hbo_genome *g, const hbo_copy *y, const hbo_copy *x) {
    const int ry = g->crank ? g->crank[y->contig] : y->contig, rx = g->crank ? g->crank[x->contig] : x->contig;
    if (ry != rx) return ry > rx;
    int c = dec_lex_cmp(y->start1, x->start1);
    if (c) return c > 0;
    c = dec_lex_cmp(y->end1, x->end1);
    if (c) return c > 0;
    return y->rel > x->rel;                                   /* '-' (0x2D) > '+' (0x2B) */
}
uint64_t hbo_genome_total(const hbo_genome *g) { return g->tot; }
int hbo_genome_code(const hbo_genome *g, uint64_t pos) { return pos < g->tot ? g->code[pos] : 4; }
uint64_t hbo_genome_contig_off(const hbo_genome *g, uint32_t c) { return g->goff[c]; }

/* oriented accessor: rel = 0 forward, 1 reverse-complement genome */
static inline int gbase(const hbo_genome *g, int rel, int64_t j) {
    if (j < 0 || (uint64_t)j >= g->tot) return 4;
    if (!rel) return g->code[j];
    uint8_t c = g->code[g->tot - 1 - (uint64_t)j];
    return c < 4 ? 3 - c : 4;
}

/* ---- minimizers (k, w): invertible 2k-bit mix (the hash64 of minimap2's sketch.c, masked) ---- */
static inline uint32_t hash_kmer(uint64_t key, int k) {
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return (uint32_t)key;
}

/* h (and strand) of the k-mer starting at every position; INVALID_H if it holds an N or is its own
   reverse complement */
static void kmer_hashes(const uint8_t *code, uint64_t tot, int k, uint32_t *h, uint8_t *strand) {
    for (uint64_t p = 0; p < tot; ++p) {
        h[p] = INVALID_H; strand[p] = 0;
        if (p + (uint64_t)k > tot) continue;
        uint64_t f = 0, r = 0; int ok = 1;
        for (int t = 0; t < k; ++t) {
            uint8_t c = code[p + t];
            if (c > 3) { ok = 0; break; }
            f |= (uint64_t)c << (2 * (k - 1 - t));
            r |= (uint64_t)(3 - c) << (2 * t);
        }
        if (!ok || f == r) continue;
        strand[p] = r < f;
        h[p] = hash_kmer(r < f ? r : f, k);
    }
}

/* position-ordered minimizer keys  h<<33 | pos<<1 | strand ; returns count (writes at most cap) */
uint64_t hbo_minimizers_codes(const uint8_t *code, uint64_t tot, int k, int w, uint64_t *out, uint64_t cap) {
    uint32_t *h = (uint32_t *)malloc(sizeof(uint32_t) * tot);
    uint8_t *st = (uint8_t *)malloc(tot);
    uint8_t *sel = (uint8_t *)calloc(tot, 1);
    kmer_hashes(code, tot, k, h, st);
    if (tot >= (uint64_t)(k + w - 1)) {
        for (uint64_t s = 0; s + (uint64_t)w <= tot; ++s) {
            uint32_t best = INVALID_H; uint64_t bp = 0;
            for (int t = 0; t < w; ++t) if (h[s + t] < best) { best = h[s + t]; bp = s + t; }   /* leftmost smallest */
            if (best != INVALID_H) sel[bp] = 1;
        }
    }
    uint64_t n = 0;
    for (uint64_t p = 0; p < tot; ++p) if (sel[p]) {
        if (n < cap) out[n] = ((uint64_t)h[p] << 33) | (p << 1) | st[p];
        ++n;
    }
    free(h); free(st); free(sel);
    return n;
}
uint64_t hbo_minimizers(const hbo_genome *g, int k, int w, uint64_t *out, uint64_t cap) {
    return hbo_minimizers_codes(g->code, g->tot, k, w, out, cap);
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
void hbo_sort_u64(uint64_t *keys, uint64_t n) { qsort(keys, n, sizeof(uint64_t), cmp_u64); }

/* ---------------------------------------------------------------------------------------------- */
/* banded affine local alignment (HB-align S7).  Query rows i in [0,L), oriented target j = j0+j',
   band o = j' - i in [-32, 31].  Scores: match +2, mismatch/N -4, gap open 4, gap extend 2.
   Outside the band / outside the genome: H = 0, E = F = -6.                                       */
/* band half-width: 16 when the seed cluster's diagonal spread is <= HB_NARROW_SPREAD (o in [-16,15]),
   else 32 (o in [-32,31])  -- HB-align S7 */
#define HB_NARROW_SPREAD 12
#define HB_TIGHT_SPREAD 3
#define HB16_MAX_L 16000
/* half band of a locus (HB-align S7): 8 (o in [-8,7]) when the cluster's trimmed diagonal spread is <= HB_TIGHT_SPREAD and the
   candidate fits the 16-bit kernel, 16 when it is <= HB_NARROW_SPREAD, else 32 */
static int half_band_of(int64_t spread, int L) {
    if (spread > HB_NARROW_SPREAD) return 32;
    return (spread <= HB_TIGHT_SPREAD && L <= HB16_MAX_L) ? 8 : 16;
}
#define SC_MATCH 2
#define SC_MIS (-4)
#define SC_GAPO 4
#define SC_GAPE 2
#define NEG6 (-6)

typedef struct { int32_t h, e, f; uint8_t tb; } cell_t;   /* tb: bits0-1 hsrc (0 stop,1 diag,2 E,3 F), bit2 eext, bit3 fext */

int hbo_align_pair(const hbo_genome *g, const uint8_t *q /*codes*/, int L, int rel, int64_t j0, int half_band,
                   hbo_aln *out, uint32_t *qrec /* L entries or NULL */) {
    const int BAND_LO = -half_band, BAND_HI = half_band - 1;
    const int W = BAND_HI - BAND_LO + 1;
    /* per-thread scratch, grown on demand (avoids an mmap/munmap + page-fault storm per pair) */
    static _Thread_local cell_t *tl_rows = NULL;
    static _Thread_local size_t tl_cap = 0;
    if ((size_t)L * W > tl_cap) { free(tl_rows); tl_cap = (size_t)L * W * 2; tl_rows = (cell_t *)malloc(sizeof(cell_t) * tl_cap); }
    cell_t *rows = tl_rows;
    int32_t best = 0; int bi = -1, bo = 0;
    for (int i = 0; i < L; ++i) {
        for (int oi = 0; oi < W; ++oi) {
            int o = BAND_LO + oi;
            int64_t jp = (int64_t)i + o;              /* j' */
            int64_t j = j0 + jp;
            cell_t *c = &rows[(size_t)i * W + oi];
            int inside = (j >= 0 && (uint64_t)j < g->tot);
            if (!inside) { c->h = 0; c->e = NEG6; c->f = NEG6; c->tb = 0; continue; }
            /* neighbours: left (i, j'-1) -> oi-1 ; up (i-1, j') -> row i-1, oi+1 ; diag (i-1,j'-1) -> row i-1, oi */
            int32_t hl = 0, el = NEG6, hu = 0, fu = NEG6, hd = 0;
            if (oi - 1 >= 0) { cell_t *l = &rows[(size_t)i * W + oi - 1]; hl = l->h; el = l->e; }
            if (i > 0 && oi + 1 < W) { cell_t *u = &rows[(size_t)(i - 1) * W + oi + 1]; hu = u->h; fu = u->f; }
            if (i > 0) { cell_t *d = &rows[(size_t)(i - 1) * W + oi]; hd = d->h; }
            int tb = 0;
            int32_t eo = hl - SC_GAPO, e;
            if (el > eo) { e = el - SC_GAPE; tb |= 4; } else e = eo - SC_GAPE;      /* tie -> open */
            int32_t fo = hu - SC_GAPO, f;
            if (fu > fo) { f = fu - SC_GAPE; tb |= 8; } else f = fo - SC_GAPE;
            int tbb = gbase(g, rel, j), qb = q[i];
            int32_t s = (qb < 4 && tbb < 4 && qb == tbb) ? SC_MATCH : SC_MIS;
            int32_t dg = hd + s;
            int32_t h = 0; int src = 0;
            if (dg > h) { h = dg; src = 1; }                                         /* priority diag > E > F */
            if (e > h) { h = e; src = 2; }
            if (f > h) { h = f; src = 3; }
            c->h = h; c->e = e; c->f = f; c->tb = (uint8_t)(tb | src);
            /* best cell: max score, then smallest i, then smallest j */
            if (h > best) { best = h; bi = i; bo = oi; }
        }
    }
    memset(out, 0, sizeof(*out));
    out->score = best;
    if (qrec) for (int i = 0; i < L; ++i) qrec[i] = 0;
    if (best <= 0) { out->i_s = out->i_e = -1; return 0; }
    /* traceback */
    int i = bi, oi = bo, state = 0; /* 0 H, 1 E, 2 F */
    int64_t je = j0 + bi + (BAND_LO + bo);
    int pending_d = 0;
    int i_s = bi; int64_t j_s = je;
    for (;;) {
        if (i < 0 || oi < 0 || oi >= W) break;
        cell_t *c = &rows[(size_t)i * W + oi];
        int64_t j = j0 + i + (BAND_LO + oi);
        if (state == 0) {
            int src = c->tb & 3;
            if (src == 0) break;
            if (src == 1) {
                if (qrec) {
                    int64_t drift = (j - j0) - i;                         /* = o */
                    qrec[i] = (uint32_t)(drift + 32768) | ((uint32_t)(pending_d > 255 ? 255 : pending_d) << 16) | (1u << 25);
                }
                pending_d = 0; i_s = i; j_s = j;
                --i;                                                      /* (i-1, j'-1): same oi */
            } else if (src == 2) state = 1;
            else state = 2;
        } else if (state == 1) {                                          /* target base j consumed alone */
            ++pending_d;
            int ext = (c->tb >> 2) & 1;
            --oi;                                                         /* (i, j'-1) */
            state = ext ? 1 : 0;
        } else {                                                          /* query base i consumed alone */
            if (qrec) qrec[i] = (uint32_t)((BAND_LO + oi) + 32768) | ((uint32_t)(pending_d > 255 ? 255 : pending_d) << 16) | (1u << 24) | (1u << 25);
            pending_d = 0; i_s = i;
            int ext = (c->tb >> 3) & 1;
            --i; ++oi;                                                    /* (i-1, j') */
            state = ext ? 2 : 0;
        }
    }
    out->i_s = i_s; out->i_e = bi; out->j_s = j_s; out->j_e = je;
    return 1;
}

/* ---------------------------------------------------------------------------------------------- */
/* S3..S6: candidate minimizers -> hits -> clusters -> ranked loci                                 */
#define DIAG_OFF 65536
#define DGAP 16
#define MIN_SEEDS 3
#define SPLIT_SPREAD 47

static uint64_t lower_bound_h(const uint64_t *idx, uint64_t n, uint32_t h) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if ((uint32_t)(idx[mid] >> 33) < h) lo = mid + 1; else hi = mid; }
    return lo;
}

typedef struct { uint32_t cand, rel, n; int64_t dmin, dmax; int64_t tlo, thi, med; } cluster_t;   /* tlo / thi / med: S5c */

static int better_cluster(const cluster_t *a, const cluster_t *b) {   /* a ranks before b */
    if (a->n != b->n) return a->n > b->n;
    if (a->rel != b->rel) return a->rel < b->rel;
    return a->dmin < b->dmin;
}

/* S5c: the band of a cluster is centred on the MEDIAN diagonal of its hits and sized by their 5 %-trimmed spread (hits a..b-1,
   sorted by diagonal): a stray hit within DGAP of the chain no longer drags the band off the copy or widens it */
static void cluster_trim(const uint64_t *hits, uint64_t a, uint64_t b, cluster_t *c) {
    const uint64_t n = b - a, t = n / 20;
    c->tlo = (int64_t)((hits[a + t] >> 13) & 0x1FFFFFFFFull);
    c->thi = (int64_t)((hits[b - 1 - t] >> 13) & 0x1FFFFFFFFull);
    c->med = (int64_t)((hits[a + n / 2] >> 13) & 0x1FFFFFFFFull);
}

/* Returns number of loci written (<= cap).  cands: codes (0..4), cand i at coff[i]..coff[i+1]. */
uint64_t hbo_find_loci(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w,
                       const uint8_t *cand_codes, const uint64_t *coff, uint32_t n_cand,
                       int occ_cap, int max_loci, hbo_locus *out, uint64_t cap) {
    /* S3: candidates laid out like a genome (PAD Ns between) and sketched with the same routine */
    uint64_t *cgoff = (uint64_t *)malloc(sizeof(uint64_t) * (n_cand + 1));
    uint64_t p = PAD;
    for (uint32_t c = 0; c < n_cand; ++c) { cgoff[c] = p; p += (coff[c + 1] - coff[c]) + PAD; }
    uint64_t ctot = (p + 63) & ~(uint64_t)63;
    uint8_t *cc = (uint8_t *)malloc(ctot);
    memset(cc, 4, ctot);
    for (uint32_t c = 0; c < n_cand; ++c) memcpy(cc + cgoff[c], cand_codes + coff[c], coff[c + 1] - coff[c]);
    uint64_t ncm = hbo_minimizers_codes(cc, ctot, k, w, NULL, 0);
    uint64_t *cm = (uint64_t *)malloc(sizeof(uint64_t) * (ncm + 1));
    hbo_minimizers_codes(cc, ctot, k, w, cm, ncm);
    /* S4: hits */
    uint64_t nh = 0, hcap = 1 << 16;
    uint64_t *hits = (uint64_t *)malloc(sizeof(uint64_t) * hcap);
    uint32_t cur = 0, qidx = 0;
    for (uint64_t m = 0; m < ncm; ++m) {
        uint64_t pos = (cm[m] >> 1) & 0xFFFFFFFFull; uint32_t h = (uint32_t)(cm[m] >> 33); uint32_t qs = cm[m] & 1;
        uint32_t c = cur;
        while (c + 1 < n_cand && pos >= cgoff[c + 1]) ++c;
        if (c != cur) { cur = c; qidx = 0; }
        uint32_t my_qidx = qidx++;
        if (my_qidx >= 8192) continue;
        int64_t qpos = (int64_t)(pos - cgoff[c]);
        uint64_t lo = lower_bound_h(idx, n_idx, h), hi = lo;
        while (hi < n_idx && (uint32_t)(idx[hi] >> 33) == h) ++hi;
        uint64_t take = hi - lo; if (take > (uint64_t)occ_cap) take = (uint64_t)occ_cap;
        for (uint64_t t = 0; t < take; ++t) {
            uint64_t e = idx[lo + t];
            int64_t gpos = (int64_t)((e >> 1) & 0xFFFFFFFFull); uint32_t gs = e & 1;
            uint32_t rel = qs ^ gs;
            int64_t d = rel ? ((int64_t)g->tot - gpos - k) - qpos + DIAG_OFF : gpos - qpos + DIAG_OFF;
            if (nh == hcap) { hcap *= 2; hits = (uint64_t *)realloc(hits, sizeof(uint64_t) * hcap); }
            hits[nh++] = ((uint64_t)c << 47) | ((uint64_t)rel << 46) | ((uint64_t)d << 13) | my_qidx;
        }
    }
    hbo_sort_u64(hits, nh);
    /* S5: clusters, then rank per candidate */
    cluster_t *cl = (cluster_t *)malloc(sizeof(cluster_t) * (nh + 1));
    uint64_t ncl = 0;
    for (uint64_t i = 0; i < nh;) {
        uint64_t key = hits[i]; uint32_t c = (uint32_t)(key >> 47), rel = (key >> 46) & 1;
        int64_t d0 = (int64_t)((key >> 13) & 0x1FFFFFFFFull), dl = d0; uint64_t j = i + 1;
        while (j < nh) {
            uint64_t k2 = hits[j];
            if ((k2 >> 46) != (key >> 46)) break;
            int64_t d = (int64_t)((k2 >> 13) & 0x1FFFFFFFFull);
            if (d - dl > DGAP) break;
            dl = d; ++j;
        }
        if (dl - d0 <= SPLIT_SPREAD) {
            if (j - i >= MIN_SEEDS) { cl[ncl].cand = c; cl[ncl].rel = rel; cl[ncl].n = (uint32_t)(j - i); cl[ncl].dmin = d0; cl[ncl].dmax = dl; cluster_trim(hits, i, j, &cl[ncl]); ++ncl; }
        } else {
            /* S5b: a run wider than the band (a copy whose diagonal drifts: net insertions / deletions) is cut into pieces of
               SPLIT_SPREAD + 1 diagonals counted from its first diagonal; each piece is aligned in its own band and the
               pieces are chained afterwards (S8b) */
            uint64_t a = i;
            while (a < j) {
                const int64_t da = (int64_t)((hits[a] >> 13) & 0x1FFFFFFFFull);
                const int64_t bucket = (da - d0) / (SPLIT_SPREAD + 1);
                uint64_t b = a + 1;
                int64_t db = da;
                while (b < j) {
                    const int64_t d = (int64_t)((hits[b] >> 13) & 0x1FFFFFFFFull);
                    if ((d - d0) / (SPLIT_SPREAD + 1) != bucket) break;
                    db = d; ++b;
                }
                if (b - a >= MIN_SEEDS) { cl[ncl].cand = c; cl[ncl].rel = rel; cl[ncl].n = (uint32_t)(b - a); cl[ncl].dmin = da; cl[ncl].dmax = db; cluster_trim(hits, a, b, &cl[ncl]); ++ncl; }
                a = b;
            }
        }
        i = j;
    }
    uint64_t nout = 0;
    for (uint64_t a = 0; a < ncl;) {
        uint64_t b = a;
        while (b < ncl && cl[b].cand == cl[a].cand) ++b;
        for (uint64_t x = a; x < b; ++x) {
            uint64_t rank = 0;
            for (uint64_t y = a; y < b; ++y) if (y != x && better_cluster(&cl[y], &cl[x])) ++rank;
            if (rank < (uint64_t)max_loci && nout < cap) {
                /* stable slot: candidate base + rank, filled below */
            }
            cl[x].dmax = (cl[x].dmax << 12) | (int64_t)(rank > 4095 ? 4095 : rank);   /* stash rank */
        }
        uint64_t cnt = (b - a) < (uint64_t)max_loci ? (b - a) : (uint64_t)max_loci;
        for (uint64_t x = a; x < b; ++x) {
            uint64_t rank = (uint64_t)(cl[x].dmax & 4095); int64_t dmax = cl[x].dmax >> 12;
            if (rank < cnt && nout + rank < cap) {
                hbo_locus *L = &out[nout + rank];
                L->cand = cl[x].cand; L->rel = cl[x].rel; L->nseeds = cl[x].n; L->rank = (uint32_t)rank;
                (void)dmax;
                L->dmin = cl[x].tlo; L->dmax = cl[x].thi;
                L->j0 = cl[x].med - DIAG_OFF;
            }
        }
        nout += cnt;
        a = b;
    }
    free(cl); free(hits); free(cm); free(cc); free(cgoff);
    return nout;
}

/* ---------------------------------------------------------------------------------------------- */
/* S8..S10 for one candidate: align every ranked locus, apply the reference's copy rules, keep the
   <=max_rows longest, and lay the copies out as a candidate-anchored star MSA (symbol ASCII).      */
static int contig_of(const hbo_genome *g, int64_t fs, int64_t fe) {
    for (uint32_t c = 0; c < g->n_contig; ++c)
        if ((uint64_t)fs >= g->goff[c] && (uint64_t)fe < g->goff[c] + g->glen[c]) return (int)c;
    return -1;
}

/* S8b: chains of local alignments.  A copy that differs from the candidate by an indel longer than the band, or whose
   diagonal drifts out of it, comes back from S7 as several partial alignments ("pieces") that are colinear in the candidate
   and in the genome -- what minimap2 reports as ONE alignment with a long I / D operation (its chaining bandwidth is 500,
   `-r`; the copy rules of Util.py:8007-8022 then judge the whole span).  Pieces of one strand are chained left to right:
   the leftmost unused partial piece (query coverage < 0.95) starts a chain and repeatedly takes the piece B that
     starts after the chain's last piece A started, ends >= CH_MINADD query bases after A's end, ends later in the
     genome, and whose junction  cut = min(A.i_e, B.i_s - 1), lowered (by at most CH_MAXTRIM) while A's genome column there is
     not left of B.j_s,  cut >= A.i_s,  tA = genome column of A at query position cut,
     B.j_s > tA,  query gap  gq = B.i_s - 1 - cut <= CH_MAXGAP,  genome gap  gt = B.j_s - tA - 1 <= CH_MAXGAP
   minimises (|gq - gt|, rank).  The chain lives in the slot of its best-ranked piece (j0 of that piece): A's records beyond
   cut are dropped, the gq skipped query bases become "query base without target" records, the gt skipped genome bases
   become target-only bases after position cut (shown up to 255), B's records follow.  Absorbed pieces are cleared. */
#define CH_MINADD 20
#define CH_MAXGAP 1000
#define CH_MAXTRIM 64
#define CH_MAXPIECES 32
static int64_t piece_tpos(const hbo_locus *loci, const uint32_t *qrecs, int L, int x, int qpos) {
    return loci[x].j0 + qpos + (int64_t)(qrecs[(size_t)x * L + qpos] & 0xFFFFu) - 32768;
}
static void chain_pieces(int nl, const hbo_locus *loci, hbo_copy *cps, uint32_t *qrecs, int L) {
    uint8_t *used = (uint8_t *)calloc((size_t)nl + 1, 1);
    int *piece = (int *)malloc(sizeof(int) * ((size_t)nl + 1));
    int *cuts = (int *)malloc(sizeof(int) * ((size_t)nl + 1));
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)L + 1));
    for (int x = 0; x < nl; ++x)
        if (cps[x].score <= 0 || (double)(cps[x].i_e - cps[x].i_s + 1) / (double)L >= 0.95) used[x] = 1;   /* not a partial piece */
    for (;;) {
        int lead = -1;                                   /* leftmost unused partial piece, best rank on ties */
        for (int x = 0; x < nl; ++x)
            if (!used[x] && (lead < 0 || cps[x].i_s < cps[lead].i_s)) lead = x;
        if (lead < 0) break;
        used[lead] = 1;
        int np = 0;
        piece[np++] = lead;
        while (np < CH_MAXPIECES) {
            const int A = piece[np - 1];
            int best = -1, best_cut = 0;
            int64_t best_key = 0;
            for (int y = 0; y < nl; ++y) {
                if (used[y] || cps[y].rel != cps[A].rel) continue;
                if (!(cps[y].i_s > cps[A].i_s && cps[y].i_e >= cps[A].i_e + CH_MINADD && cps[y].j_e > cps[A].j_e)) continue;
                int cut = cps[A].i_e < cps[y].i_s - 1 ? cps[A].i_e : cps[y].i_s - 1;
                /* the two local alignments may share a few genome bases around the junction: A gives them up */
                for (int n = 0; n < CH_MAXTRIM && cut >= cps[A].i_s && piece_tpos(loci, qrecs, L, A, cut) >= cps[y].j_s; ++n) --cut;
                if (cut < cps[A].i_s) continue;
                const int64_t tA = piece_tpos(loci, qrecs, L, A, cut);
                if (!(cps[y].j_s > tA)) continue;
                const int64_t gq = cps[y].i_s - 1 - cut, gt = cps[y].j_s - tA - 1;
                if (gq > CH_MAXGAP || gt > CH_MAXGAP) continue;
                const int64_t key = ((gq > gt ? gq - gt : gt - gq) << 12) | (int64_t)y;
                if (best < 0 || key < best_key) { best = y; best_key = key; best_cut = cut; }
            }
            if (best < 0) break;
            cuts[np - 1] = best_cut;
            used[best] = 1;
            piece[np++] = best;
        }
        if (np == 1) continue;
        int h = piece[0];
        for (int t = 1; t < np; ++t) if (piece[t] < h) h = piece[t];     /* best rank = smallest locus index */
        const int first = piece[0], last = piece[np - 1];
        const int cs = cps[first].i_s, ce = cps[last].i_e;
        const int64_t js = cps[first].j_s, je = cps[last].j_e;
        int32_t score = 0;
        for (int t = 0; t < np; ++t) {
            const int x = piece[t];
            const int from = cps[x].i_s, to = (t + 1 < np) ? cuts[t] : cps[x].i_e;
            const int64_t delta = loci[x].j0 - loci[h].j0;
            for (int p = from; p <= to; ++p) {
                const uint32_t v = qrecs[(size_t)x * L + p];
                tmp[p] = (v & 0xFFFF0000u) | (uint32_t)((int64_t)(v & 0xFFFFu) + delta);
            }
            score += cps[x].score;
            if (t + 1 < np) {
                const int y = piece[t + 1];
                const int64_t tA = piece_tpos(loci, qrecs, L, x, to);
                const int64_t gt = cps[y].j_s - tA - 1;
                tmp[to] = (tmp[to] & ~(255u << 16)) | ((uint32_t)(gt > 255 ? 255 : gt) << 16);
                for (int p = to + 1; p < cps[y].i_s; ++p)
                    tmp[p] = (uint32_t)(tA - loci[h].j0 - p + 32768) | (1u << 24) | (1u << 25);
            }
        }
        for (int p = 0; p < L; ++p) qrecs[(size_t)h * L + p] = (p >= cs && p <= ce) ? tmp[p] : 0;
        for (int t = 0; t < np; ++t) {
            const int x = piece[t];
            if (x == h) continue;
            cps[x].score = 0; cps[x].i_s = -1; cps[x].i_e = -1; cps[x].j_s = 0; cps[x].j_e = 0;
        }
        cps[h].score = score; cps[h].i_s = cs; cps[h].i_e = ce; cps[h].j_s = js; cps[h].j_e = je;
    }
    free(used); free(piece); free(cuts); free(tmp);
}

int hbo_candidate_msa(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w,
                      const uint8_t *cand_ascii, int L, int occ_cap, int max_loci, int flank, int max_rows,
                      hbo_copy *copies, int *n_copies, uint8_t *matrix, uint64_t matrix_cap, int *R_out, int *C_out) {
    uint8_t *q = (uint8_t *)malloc((size_t)L + 1);
    for (int i = 0; i < L; ++i) q[i] = code_of(cand_ascii[i]);
    uint64_t coff[2] = {0, (uint64_t)L};
    hbo_locus *loci = (hbo_locus *)malloc(sizeof(hbo_locus) * (size_t)(max_loci + 1));
    int nl = (int)hbo_find_loci(g, idx, n_idx, k, w, q, coff, 1, occ_cap, max_loci, loci, (uint64_t)max_loci);
    uint32_t *qrecs = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nl > 0 ? nl : 1) * (size_t)L);
    int nk = 0;
    for (int x = 0; x < nl; ++x) {
        hbo_aln a;
        hbo_copy *cp = &copies[x];
        memset(cp, 0, sizeof(*cp));
        cp->rank = x; cp->rel = (int32_t)loci[x].rel; cp->contig = -1;
        hbo_align_pair(g, q, L, (int)loci[x].rel, loci[x].j0,
                       half_band_of(loci[x].dmax - loci[x].dmin, L), &a, qrecs + (size_t)x * L);
        cp->score = a.score; cp->i_s = a.i_s; cp->i_e = a.i_e; cp->j_s = a.j_s; cp->j_e = a.j_e;
    }
    if (getenv("HBO_DEBUG_SPREAD")) for (int x = 0; x < nl; ++x) fprintf(stderr, "SPREAD %lld %u\n", (long long)(loci[x].dmax - loci[x].dmin), loci[x].nseeds);
    if (!getenv("HBO_NO_CHAIN")) chain_pieces(nl, loci, copies, qrecs, L);   /* S8b (the switch is a debugging aid) */
    for (int x = 0; x < nl; ++x) {
        hbo_copy *cp = &copies[x];
        hbo_aln a;
        a.score = cp->score; a.i_s = cp->i_s; a.i_e = cp->i_e; a.j_s = cp->j_s; a.j_e = cp->j_e;
        if (a.score <= 0) continue;
        const int aligned = a.i_e - a.i_s + 1;                         /* query_alignment_length (M+I) */
        const int64_t tlen = a.j_e - a.j_s + 1;                        /* sum(M+D)                      */
        if (!((double)aligned / (double)L >= 0.95)) continue;          /* Util.py:8007, 8022            */
        if (!((double)aligned / (double)tlen >= 0.95)) continue;       /* Util.py:8013, 8022            */
        int64_t fs = cp->rel ? (int64_t)g->tot - 1 - a.j_e : a.j_s;
        int64_t fe = cp->rel ? (int64_t)g->tot - 1 - a.j_s : a.j_e;
        int c = contig_of(g, fs, fe);
        if (c < 0) continue;
        cp->contig = c;
        cp->start1 = fs - (int64_t)g->goff[c] + 1;
        cp->end1 = fe - (int64_t)g->goff[c] + 1;
        if (cp->start1 - 1 - flank < 0 || cp->end1 + flank > (int64_t)g->glen[c]) continue;   /* Util.py:8103 */
        cp->ext_len = (int32_t)(cp->end1 - cp->start1 + 1 + 2 * flank);
        if (cp->ext_len < 100) continue;                                                       /* Util.py:8108 */
        int dup = 0;
        for (int y = 0; y < x; ++y)
            if (copies[y].keep && copies[y].rel == cp->rel && copies[y].j_s == cp->j_s && copies[y].j_e == cp->j_e) { dup = 1; break; }
        if (dup) continue;
        cp->keep = 1; ++nk;
    }
    if (nk > max_rows) {                                   /* 100 longest, file order kept (Util.py:10409-10412) */
        for (int x = 0; x < nl; ++x) {
            if (!copies[x].keep) continue;
            int longer = 0;
            for (int y = 0; y < nl; ++y) {
                if (y == x || !copies[y].keep) continue;
                if (g->tie_rule == 2) { longer += y < x; continue; }      /* FiLTR's extract_copies: the first max_rows of the file */
                const int first = g->tie_rule ? name_greater(g, &copies[y], &copies[x]) : (y < x);
                if (copies[y].ext_len > copies[x].ext_len || (copies[y].ext_len == copies[x].ext_len && first)) ++longer;
            }
            if (longer >= max_rows) copies[x].keep = 2;    /* mark, drop after the pass */
        }
        for (int x = 0; x < nl; ++x) if (copies[x].keep == 2) copies[x].keep = 0;
    }
    *n_copies = nl;
    /* star MSA */
    int *ins = (int *)calloc((size_t)L + 1, sizeof(int));
    int R = 0;
    for (int x = 0; x < nl; ++x) {
        if (!copies[x].keep) continue;
        ++R;
        const uint32_t *qr = qrecs + (size_t)x * L;
        for (int p = copies[x].i_s; p <= copies[x].i_e; ++p) {
            int d = (int)((qr[p] >> 16) & 255);
            if (d > ins[p]) ins[p] = d;
        }
    }
    int64_t *colb = (int64_t *)malloc(sizeof(int64_t) * (size_t)(L + 2 * flank + 1));   /* index p + flank */
    int64_t acc = 0;
    for (int p = -flank; p < L + flank; ++p) {
        colb[p + flank] = (p + flank) + acc;
        if (p >= 0 && p < L) acc += ins[p];
    }
    const int64_t C = (int64_t)L + 2 * flank + acc;
    *R_out = R; *C_out = (int)C;
    if ((uint64_t)R * (uint64_t)C > matrix_cap) { free(ins); free(colb); free(qrecs); free(loci); free(q); return -1; }
    memset(matrix, '-', (size_t)R * (size_t)C);
    int r = 0;
    static const char ASC[5] = {'A', 'C', 'G', 'T', 'N'};
    for (int x = 0; x < nl; ++x) {
        if (!copies[x].keep) continue;
        uint8_t *row = matrix + (size_t)r * (size_t)C;
        const uint32_t *qr = qrecs + (size_t)x * L;
        const hbo_copy *cp = &copies[x];
        for (int p = cp->i_s; p <= cp->i_e; ++p) {
            uint32_t v = qr[p];
            /* j = last target column consumed at or before query position p (for an unaligned query
               base: the column at which it was inserted) */
            int64_t j = loci[x].j0 + p + ((int64_t)(v & 0xFFFF) - 32768);
            if (!((v >> 24) & 1)) row[colb[p + flank]] = (uint8_t)ASC[gbase(g, cp->rel, j)];
            int d = (int)((v >> 16) & 255);
            for (int t = 0; t < d; ++t) row[colb[p + flank] + 1 + t] = (uint8_t)ASC[gbase(g, cp->rel, j + 1 + t)];
        }
        for (int t = 1; t <= flank; ++t) {
            row[colb[cp->i_s - t + flank]] = (uint8_t)ASC[gbase(g, cp->rel, cp->j_s - t)];
            row[colb[cp->i_e + t + flank]] = (uint8_t)ASC[gbase(g, cp->rel, cp->j_e + t)];
        }
        ++r;
    }
    free(ins); free(colb); free(qrecs); free(loci); free(q);
    return 0;
}

/* ---------------------------------------------------------------------------------------------- */
/* Library-vs-genome HSPs (checker of hb_query_hsps): every positive local alignment of a query x ranked locus
   pair inside one contig, as the outfmt-6 fields get_copies_v1 reads (Util.py:7032-7056): 1-based inclusive
   spans, minus strand as ss > se, matches and alignment length from the traceback's edit record. */
int hbo_query_hsps(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w, const uint8_t *query_ascii,
                   int L, int occ_cap, int max_loci, hbo_qhsp *out, int cap) {
    uint8_t *q = (uint8_t *)malloc((size_t)L + 1);
    for (int i = 0; i < L; ++i) q[i] = code_of(query_ascii[i]);
    uint64_t coff[2] = {0, (uint64_t)L};
    hbo_locus *loci = (hbo_locus *)malloc(sizeof(hbo_locus) * (size_t)(max_loci + 1));
    int nl = (int)hbo_find_loci(g, idx, n_idx, k, w, q, coff, 1, occ_cap, max_loci, loci, (uint64_t)max_loci);
    uint32_t *qr = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)L);
    int n = 0;
    for (int x = 0; x < nl && n < cap; ++x) {
        hbo_aln a;
        const int rel = (int)loci[x].rel;
        hbo_align_pair(g, q, L, rel, loci[x].j0, half_band_of(loci[x].dmax - loci[x].dmin, L), &a, qr);
        if (a.score <= 0) continue;
        const int64_t fs = rel ? (int64_t)g->tot - 1 - a.j_e : a.j_s;
        const int64_t fe = rel ? (int64_t)g->tot - 1 - a.j_s : a.j_e;
        const int c = contig_of(g, fs, fe);
        if (c < 0) continue;
        int match = 0, extra = 0;
        for (int i = a.i_s; i <= a.i_e; ++i) {
            const uint32_t v = qr[i];
            extra += (int)((v >> 16) & 255u);
            if ((v >> 24) & 1u) continue;
            const int tb = gbase(g, rel, loci[x].j0 + i + (int64_t)(v & 0xFFFFu) - 32768);
            if (q[i] < 4 && q[i] == tb) ++match;
        }
        const int32_t s1 = (int32_t)(fs - (int64_t)g->goff[c] + 1), e1 = (int32_t)(fe - (int64_t)g->goff[c] + 1);
        hbo_qhsp *o = &out[n++];
        o->rank = x; o->contig = c; o->qs = a.i_s + 1; o->qe = a.i_e + 1;
        o->ss = rel ? e1 : s1; o->se = rel ? s1 : e1;
        o->n_match = match; o->aln_len = a.i_e - a.i_s + 1 + extra; o->score = a.score;
    }
    free(qr); free(loci); free(q);
    return n;
}
