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

/* ---- HB-hsp S11b: base-level ends.  A seed run is grown along its diagonal over 8-base blocks (aligned in query
   coordinates): a block is GOOD when at least 5 of its 8 positions match (N, or a position outside the query / subject
   segment, is a mismatch).  The segment of a seed = the maximal run of good blocks around the first block that lies
   wholly inside the seed, at most HSP_MAXEXT blocks to either side; at each end it takes, from the adjacent bad block,
   the prefix with the best +1 / -2 score (first maximum, only if positive), and otherwise shrinks to the outermost match
   of its last good block.  Everything depends on positions only, so every seed inside one segment yields the same segment:
   duplicates are dropped by comparing neighbours.  (blastn extends its word hits by x-drop; this is the deterministic,
   word-parallel stand-in -- parity unpinned like the rest of HB-hsp.) */
#define HSP_MAXEXT 4096

typedef struct { const hbo_genome *g; int rel; int64_t d; int64_t qlo, qhi, slo, shi; uint64_t tot; } diag_t;

static int diag_match(const diag_t *D, int64_t q) {
    if (q < D->qlo || q >= D->qhi) return 0;
    const int64_t j = q + D->d;                                   /* oriented subject coordinate */
    const int64_t f = D->rel ? (int64_t)D->tot - 1 - j : j;       /* forward position            */
    if (f < D->slo || f >= D->shi) return 0;
    const int a = hbo_genome_code(D->g, (uint64_t)q);
    int b = hbo_genome_code(D->g, (uint64_t)f);
    if (a > 3 || b > 3) return 0;
    if (D->rel) b = 3 - b;
    return a == b;
}
static int block_mask(const diag_t *D, int64_t b) {
    int m = 0;
    for (int i = 0; i < 8; ++i) m |= diag_match(D, 8 * b + i) << i;
    return m;
}
static int popc8(int m) { int c = 0; for (; m; m &= m - 1) ++c; return c; }
/* segment [*qa, *qb] (inclusive, global query coordinates) of the seed starting at q0 */
static void seed_segment(const diag_t *D, int64_t q0, int64_t *qa, int64_t *qb) {
    const int64_t bc = (q0 + 7) >> 3;
    int64_t bl = bc, br = bc;
    if (block_mask(D, bc) == 0) { *qa = *qb = q0; return; }     /* a segment boundary cuts the whole core block off */
    for (int n = 0; n < HSP_MAXEXT && popc8(block_mask(D, bl - 1)) >= 5; ++n) --bl;
    for (int n = 0; n < HSP_MAXEXT && popc8(block_mask(D, br + 1)) >= 5; ++n) ++br;
    {   /* left end */
        const int m = block_mask(D, bl - 1);
        int sc = 0, best = 0, ext = 0;
        for (int i = 7; i >= 0; --i) { sc += ((m >> i) & 1) ? 1 : -2; if (sc > best) { best = sc; ext = 8 - i; } }
        if (ext) *qa = 8 * bl - ext;
        else { const int ml = block_mask(D, bl); int i = 0; while (!((ml >> i) & 1)) ++i; *qa = 8 * bl + i; }
    }
    {   /* right end */
        const int m = block_mask(D, br + 1);
        int sc = 0, best = 0, ext = 0;
        for (int i = 0; i < 8; ++i) { sc += ((m >> i) & 1) ? 1 : -2; if (sc > best) { best = sc; ext = i + 1; } }
        if (ext) *qb = 8 * br + 7 + ext;
        else { const int mr = block_mask(D, br); int i = 7; while (!((mr >> i) & 1)) --i; *qb = 8 * br + i; }
    }
}

/* seg_goff: global padded start of every segment (ascending), seg_len[g] its length.  Returns the number of HSPs
   (writes <= cap). */
uint64_t hbo_self_hsps(const hbo_genome *g, const uint64_t *mm_pos, const uint64_t *idx, uint64_t n, uint64_t tot, int k,
                       const uint64_t *seg_goff, const uint64_t *seg_len, int n_seg, int occ_cap, hbo_hsp_rec *out, uint64_t cap) {
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
            const uint64_t gk = idx[l + x];
            if (gk == e) continue;
            const int64_t gpos = (int64_t)((gk >> 1) & 0xFFFFFFFFull);
            const uint64_t rel = (e ^ gk) & 1;
            const int64_t d = (rel ? ((int64_t)tot - gpos - k) : gpos) - qpos + (1ll << 32);
            if (np == pcap) { pcap *= 2; keys = (uint64_t *)realloc(keys, sizeof(uint64_t) * pcap); }
            keys[np++] = (rel << 62) | ((uint64_t)d << 29) | qi;
            ++taken;
        }
    }
    qsort(keys, np, sizeof(uint64_t), cmp_u64);
    uint64_t nh = 0;
    int have_prev = 0;
    uint64_t prev_diag = 0;
    int64_t prev_a = 0, prev_b = 0;                    /* the last segment computed on the current diagonal */
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
        const int64_t s0 = rel ? (int64_t)tot - (q0 + d) - k : q0 + d;
        const int qseg = seg_of(seg_goff, n_seg, (uint64_t)q0), sseg = seg_of(seg_goff, n_seg, (uint64_t)s0);
        const int64_t qg = (int64_t)seg_goff[qseg], sg = (int64_t)seg_goff[sseg];
        diag_t D = {g, rel, d, qg, qg + (int64_t)seg_len[qseg], sg, sg + (int64_t)seg_len[sseg], tot};
        int64_t seg[2][2];
        seed_segment(&D, q0, &seg[0][0], &seg[0][1]);
        seg[1][0] = seg[0][0]; seg[1][1] = seg[0][1];
        if (q1 + k - 1 > seg[0][1]) seed_segment(&D, q1, &seg[1][0], &seg[1][1]);
        const uint64_t diag = keys[i] >> 29;
        for (int z = 0; z < 2; ++z) {
            const int64_t a = seg[z][0], b = seg[z][1];
            const int dup = (z == 1 && a == seg[0][0] && b == seg[0][1]) ||
                            (z == 0 && have_prev && prev_diag == diag && prev_a == a && prev_b == b);
            if (!dup && b - a + 1 >= HSP_MINSPAN) {
                if (nh < cap) {
                    hbo_hsp_rec *o = &out[nh];
                    o->qseg = qseg; o->subj = sseg;
                    o->qs = (int32_t)(a - qg + 1); o->qe = (int32_t)(b - qg + 1);
                    if (!rel) { o->ss = (int32_t)(a + d - sg + 1); o->se = (int32_t)(b + d - sg + 1); }
                    else { o->ss = (int32_t)((int64_t)tot - 1 - (a + d) - sg + 1); o->se = (int32_t)((int64_t)tot - 1 - (b + d) - sg + 1); }
                }
                ++nh;
            }
        }
        have_prev = 1; prev_diag = diag; prev_a = seg[1][0]; prev_b = seg[1][1];
        i = j;
    }
    free(keys);
    return nh;
}
