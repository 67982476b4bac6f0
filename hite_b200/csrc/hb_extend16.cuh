// hite_b200 -- kernel (ii), banded affine Smith-Waterman extension with TWO alignments per lane.
//
// Same recurrence, band geometry, tie rules and traceback stream as hb_extend.cu (HB-align S7,
// DESIGN.md section 4; replaces the base-level alignment of `minimap2 -ax map-ont`, reference
// module/Util.py:7952-7961, whose CIGAR feeds get_copies_minimap2, module/Util.py:7977-8030), but
// every register holds the cells of two independent candidate x locus pairs as signed 16-bit halves
// (pair A = bits 0-15, pair B = bits 16-31) and the recurrences are the Blackwell/Hopper DPX
// instructions VIADDMNMX.S16x2 / VIMNMX3.S16x2.RELU / VIADD.16x2.  Scores are even and bounded by 2L,
// so candidates up to HB16_MAX_L bases fit; longer ones take the 32-bit kernel of hb_extend.cu.
//
// The warp body is written against a small "warp" interface (lane id, shuffles, sync) so that the
// SAME source runs on the device and in the host lock-step emulator the CPU test-suite uses
// (tests/emu/emu_extend16.cu): the emulator is test infrastructure, the product only ever
// instantiates DevWarp.
#pragma once
#include "hb_common.cuh"
#include "hb_extend.cuh"

#define HB_HD __host__ __device__ __forceinline__

namespace hb16 {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int HB16_MAX_L = HB_EXTEND16_MAX_L;       // 2 * L + 4 + 6 < 32768
constexpr int NARROW_SPREAD = 12;        // HB-align S7 (5 %-trimmed diagonal spread of the cluster, band centred on its median)
constexpr int TIGHT_SPREAD = 3;          // ... and 16 diagonals are enough when the hits sit within 3 diagonals
// half band (lanes per pair) of a locus: 8 / 16 / 32; candidates beyond the 16-bit kernel only have the 16 / 32 kernels
HB_HD int half_band(int64_t spread, int L) {
    if (spread > NARROW_SPREAD) return 32;
    return (spread <= TIGHT_SPREAD && L <= HB16_MAX_L) ? 8 : 16;
}

// ---- packed signed-16x2 primitives (device: one SASS instruction each; host: plain C) ------------
#ifndef __CUDA_ARCH__
static inline int16_t h_lo(uint32_t v) { return (int16_t)(v & 0xFFFFu); }
static inline int16_t h_hi(uint32_t v) { return (int16_t)(v >> 16); }
static inline uint32_t h_pack(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | (((uint32_t)hi & 0xFFFFu) << 16); }
#endif
HB_HD uint32_t add2(uint32_t a, uint32_t b) {                       // wrapping add per half
#ifdef __CUDA_ARCH__
    return __vadd2(a, b);
#else
    return h_pack((int)(a & 0xFFFF) + (int)(b & 0xFFFF), (int)(a >> 16) + (int)(b >> 16));
#endif
}
HB_HD uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) {         // max(a + b, c)
#ifdef __CUDA_ARCH__
    return __viaddmax_s16x2(a, b, c);
#else
    const uint32_t s = add2(a, b);
    return h_pack(h_lo(s) > h_lo(c) ? h_lo(s) : h_lo(c), h_hi(s) > h_hi(c) ? h_hi(s) : h_hi(c));
#endif
}
HB_HD uint32_t addmax_relu(uint32_t a, uint32_t b, uint32_t c) {    // max(a + b, c, 0)
#ifdef __CUDA_ARCH__
    return __viaddmax_s16x2_relu(a, b, c);
#else
    const uint32_t m = addmax(a, b, c);
    return h_pack(h_lo(m) > 0 ? h_lo(m) : 0, h_hi(m) > 0 ? h_hi(m) : 0);
#endif
}
HB_HD uint32_t max3_relu(uint32_t a, uint32_t b, uint32_t c) {      // max(a, b, c, 0)
#ifdef __CUDA_ARCH__
    return __vimax3_s16x2_relu(a, b, c);
#else
    int lo = h_lo(a), hi = h_hi(a);
    if (h_lo(b) > lo) lo = h_lo(b);
    if (h_lo(c) > lo) lo = h_lo(c);
    if (h_hi(b) > hi) hi = h_hi(b);
    if (h_hi(c) > hi) hi = h_hi(c);
    return h_pack(lo > 0 ? lo : 0, hi > 0 ? hi : 0);
#endif
}
HB_HD uint32_t max3(uint32_t a, uint32_t b, uint32_t c) {           // max(a, b, c)
#ifdef __CUDA_ARCH__
    return __vimax3_s16x2(a, b, c);
#else
    int lo = h_lo(a), hi = h_hi(a);
    if (h_lo(b) > lo) lo = h_lo(b);
    if (h_lo(c) > lo) lo = h_lo(c);
    if (h_hi(b) > hi) hi = h_hi(b);
    if (h_hi(c) > hi) hi = h_hi(c);
    return h_pack(lo, hi);
#endif
}
HB_HD uint32_t minu2(uint32_t a, uint32_t b) {                      // unsigned min per half
#ifdef __CUDA_ARCH__
    return __vminu2(a, b);
#else
    const uint32_t al = a & 0xFFFF, bl = b & 0xFFFF, ah = a >> 16, bh = b >> 16;
    return (al < bl ? al : bl) | ((ah < bh ? ah : bh) << 16);
#endif
}
HB_HD uint32_t min3u2(uint32_t a, uint32_t b, uint32_t c) {         // unsigned min(a, b, c) per half
#ifdef __CUDA_ARCH__
    return __vimin3_u16x2(a, b, c);
#else
    return minu2(minu2(a, b), c);
#endif
}
// max(a, b) per half; *keep_x = (a >= b) for that half
HB_HD uint32_t bmax(uint32_t a, uint32_t b, bool* keep_hi, bool* keep_lo) {
#ifdef __CUDA_ARCH__
    return __vibmax_s16x2(a, b, keep_hi, keep_lo);
#else
    *keep_lo = h_lo(a) >= h_lo(b);
    *keep_hi = h_hi(a) >= h_hi(b);
    return h_pack(*keep_lo ? h_lo(a) : h_lo(b), *keep_hi ? h_hi(a) : h_hi(b));
#endif
}
HB_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {         // byte permute (no sign modes used)
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

// ---- packed genome / candidate accessors (same layout as hb_index.cu::pack_kernel) ---------------
HB_HD int base_at(const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask, uint64_t p) {
    if ((nmask[p >> 5] >> (p & 31)) & 1u) return 4;
    return (int)((seq2[p >> 4] >> (2 * (p & 15))) & 3u);
}
HB_HD int oriented_base_at(const uint32_t* __restrict__ gseq2, const uint32_t* __restrict__ gnmask, uint64_t tot,
                           int rel, int64_t j) {
    if (j < 0 || (uint64_t)j >= tot) return 4;
    if (!rel) return base_at(gseq2, gnmask, (uint64_t)j);
    const int c = base_at(gseq2, gnmask, tot - 1 - (uint64_t)j);
    return c < 4 ? 3 - c : 4;
}
// one-hot base codes: A 8, C 16, G 32, T 64, N / outside 0.  (q & t) != 0 <=> both the same real base.
HB_HD uint32_t onehot(int c) { return c < 4 ? (8u << c) : 0u; }

constexpr uint32_t K_M2 = 0xFFFEFFFEu;   // (-2, -2)  gap extend
constexpr uint32_t K_M6 = 0xFFFAFFFAu;   // (-6, -6)  gap open + extend
constexpr uint32_t K_6 = 0x00060006u;    // match - mismatch
constexpr uint32_t K_4 = 0x00040004u;    // the bias: stored E, F, Z and the H inside a cell are score + 4
constexpr uint32_t K_1 = 0x00010001u;

// One DP cell for both halves.  PAR 0: even anti-diagonal (left neighbour was shuffled in, the lane at
// the band edge has none), PAR 1: odd (up neighbour shuffled in).
//
// Representation.  E and F are kept clamped at 0 like in hb_extend.cu, and every stored value carries a
// +4 bias (E', F', Z' = score + 4, so "zero" is 4); the diagonal input hd is the UNBIASED H of two steps
// ago.  With that, hd + {6, 0} is directly the biased diagonal candidate (H_d + 4 + {+2, -4}) and
// h' - 4 the unbiased H: both are plain 32-bit adds of non-negative halves that can never carry across
// the halves, so ptxas is free to issue them on the FMA pipe (IMAD.IADD) -- the kernel is bound by the
// ALU pipe -- and the RELU disappears because E', F' >= 4 already floor the maximum.
//   z = max(H' - 6, 4) is the clamped "open" value of a cell, computed ONCE by its owner (Z) and
//   shuffled instead of H, so a step costs three VIADDMNMX, not four.
//   Gap-extend flag = (E' != z): differs from the oracle's (E_l - 2 > H_l - 6) only where E == 0, and
//   such a cell is never walked in gap state.  At a band-edge lane the forced "zero" E' / F' may be below
//   the garbage z it was compared with: both halves sit on the edge there, the borrow stays inside
//   that cell's flags, and they are never read.
// NW: traceback nibbles, pair A at bits 0-3, pair B at bits 16-19:
//   bits0-1 source of H (0 stop, 1 diag, 2 E, 3 F), bit2 E extended, bit3 F extended
template <int PAR>
HB_HD void cell2(uint32_t zl, uint32_t el, uint32_t zu, uint32_t fu, uint32_t hd, uint32_t qb, uint32_t tb,
                 uint32_t edge_keep, uint32_t edge_fill, uint32_t& HU, uint32_t& E, uint32_t& F, uint32_t& Z,
                 uint32_t& NW) {
    uint32_t e = addmax(el, K_M2, zl);
    uint32_t f = addmax(fu, K_M2, zu);
    if (PAR == 0) e = (e & edge_keep) | edge_fill; else f = (f & edge_keep) | edge_fill;
    const uint32_t s6 = minu2(qb & tb, K_6);              // 6 on a match, 0 otherwise
    const uint32_t dg = hd + s6;                          // biased diagonal candidate, >= 0
    const uint32_t h = max3(dg, e, f);                    // biased H (>= 4)
    const uint32_t hu = h - K_4;                          // unbiased H, >= 0
    // flags as 0 / 1 per half.  u = H - diag, v = H - E, hu = H are >= 0, so "x > 0" is min(x, 1) and a
    // conjunction is a three-input min; v > 0 implies H > 0 (H' > E' >= 4).  Source of H:
    //   0 stop, 1 diag, 2 E, 3 F  =  [H > 0] + [H > 0 and H != diag] + [H != diag and H != E]
    const uint32_t u = h - dg, v = h - e;
    const uint32_t fa = minu2(e - zl, K_1);               // E extended
    const uint32_t fb = minu2(f - zu, K_1);               // F extended
    const uint32_t p0 = minu2(hu, K_1);
    const uint32_t p1 = min3u2(u, hu, K_1);
    const uint32_t p2 = min3u2(u, v, K_1);
    NW = p0 + p1 + p2 + 4u * fa + 8u * fb;
    HU = hu; E = e; F = f;
    Z = addmax(h, K_M6, K_4);
}

// ---- the warp body ---------------------------------------------------------------------------------
// HB = lanes per lane-group (half band): 16 -> band o in [-16,15], two groups (4 pairs) per warp;
// 32 -> o in [-32,31], one group (2 pairs) per warp.  Group g extends pairs order[2g] and order[2g+1].
// s_rq / s_rt: (32 / HB) * (HB + 32) words each, private to the warp.
template <int HB, class W>
HB_HD void extend16_warp(W& w, uint32_t group0, const uint32_t* __restrict__ order, uint32_t n_order,
                         const HbLocus* __restrict__ pairs, const uint32_t* __restrict__ gseq2,
                         const uint32_t* __restrict__ gnmask, uint64_t tot, const uint32_t* __restrict__ cseq2,
                         const uint32_t* __restrict__ cnmask, const uint64_t* __restrict__ cgoff,
                         const uint32_t* __restrict__ clen, const uint32_t* __restrict__ tb_unit_off,
                         uint32_t tb_unit_base, uint32_t* __restrict__ tbuf, HbAln* __restrict__ alns,
                         uint32_t* s_rq, uint32_t* s_rt) {
    constexpr int WIN = HB + 32;                 // sequence window of one super-step (64 anti-diagonals)
    const int lane = w.lane();
    const int sub = lane / HB, sl = lane % HB;
    const uint32_t g = group0 + (uint32_t)sub;
    uint32_t pid[2];
    int L[2], rel[2], msup[2];
    uint64_t cstart[2];
    int64_t j0[2];
    uint32_t* tbp[2];
#pragma unroll
    for (int z = 0; z < 2; ++z) {
        const uint64_t idx = 2ull * g + (uint64_t)z;
        pid[z] = idx < n_order ? order[idx] : NONE;
        L[z] = 0; rel[z] = 0; msup[z] = 0; cstart[z] = 0; j0[z] = 0; tbp[z] = tbuf;
        if (pid[z] != NONE) {
            const HbLocus pr = pairs[pid[z]];
            L[z] = (int)clen[pr.cand];
            cstart[z] = cgoff[pr.cand];
            rel[z] = (int)pr.rel;
            j0[z] = pr.j0;
            tbp[z] = tbuf + ((size_t)(tb_unit_off[pid[z]] - tb_unit_base) << 7);     // 128 words per unit
            msup[z] = (2 * L[z] + 2 * HB - 2 + 63) >> 6;
        }
    }
    const int n_super = w.max_all(msup[0] > msup[1] ? msup[0] : msup[1]);
    if (n_super == 0) return;
    uint32_t* rq = s_rq + sub * WIN;
    uint32_t* rt = s_rt + sub * WIN;

    // Branch-free one-hot fetches (the window refill runs once per 64 anti-diagonals but for both pairs,
    // both sequences and 32 / HB positions per lane): out-of-range positions read word 0 and select 0.
    auto code_at = [&](const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask, uint64_t pos, bool ok,
                       bool comp) -> uint32_t {
        const uint64_t p = ok ? pos : 0;
        const uint32_t nm = (nmask[p >> 5] >> (p & 31)) & 1u;
        const uint32_t b = (seq2[p >> 4] >> (2 * (p & 15))) & 3u;
        const uint32_t oh = comp ? (64u >> b) : (8u << b);           // complement: 3 - b
        return (ok && !nm) ? oh : 0u;
    };
    auto qcode = [&](int x) -> uint32_t {
        uint32_t v = 0;
#pragma unroll
        for (int z = 0; z < 2; ++z)
            v |= code_at(cseq2, cnmask, cstart[z] + (uint64_t)(int64_t)x, (unsigned)x < (unsigned)L[z], false) << (16 * z);
        return v;
    };
    auto tcode = [&](int x) -> uint32_t {
        uint32_t v = 0;
#pragma unroll
        for (int z = 0; z < 2; ++z) {
            const int64_t j = j0[z] + x;
            const bool ok = pid[z] != NONE && j >= 0 && (uint64_t)j < tot;
            v |= code_at(gseq2, gnmask, rel[z] ? tot - 1 - (uint64_t)j : (uint64_t)j, ok, rel[z] != 0) << (16 * z);
        }
        return v;
    };
    // window of the first super-step: positions [-HB, 32)
    for (int t = sl; t < WIN; t += HB) { rq[t] = qcode(-HB + t); rt[t] = tcode(-HB + t); }
    w.sync();

    uint32_t HU1 = 0, E1 = K_4, F1 = K_4, Z1 = K_4, Hd = 0;   // HU1 / Hd: unbiased H one / two steps ago; E1, F1, Z1 biased
    uint32_t best = 0;                                    // running maximum of H per half
    int rbest[2] = {0, 0};                                // earliest anti-diagonal that reached it
    const uint32_t m_up = sl == 0 ? 0u : 0xffffffffu;     // band edge: neighbour outside = (H 0, E/F 0)
    const uint32_t m_dn = sl == HB - 1 ? 0u : 0xffffffffu;
    const uint32_t f_up = ~m_up & K_4, f_dn = ~m_dn & K_4;  // ... i.e. the biased zero
    const uint32_t* pq = rq + (HB - sl);                  // query index of pair-step u:  u - sl + HB
    const uint32_t* pt = rt + sl;                         // target index: u + sl (even), u + sl + 1 (odd)
    for (int sup = 0; sup < n_super; ++sup) {
        if (sup > 0) {
            const uint32_t a = rq[32 + sl], b = rt[32 + sl];
            uint32_t nq[32 / HB], nt[32 / HB];
#pragma unroll
            for (int z = 0; z < 32 / HB; ++z) { nq[z] = qcode(32 * sup + sl + z * HB); nt[z] = tcode(32 * sup + sl + z * HB); }
            w.sync();
            rq[sl] = a; rt[sl] = b;
#pragma unroll
            for (int z = 0; z < 32 / HB; ++z) { rq[HB + sl + z * HB] = nq[z]; rt[HB + sl + z * HB] = nt[z]; }
            w.sync();
        }
        const int r0 = 64 * sup;
        uint32_t twA[8], twB[8];                          // the 8 traceback words of this super-step, per pair
        uint32_t fq[4];
        uint32_t tnext = pt[0];
#pragma unroll 8
        for (int u = 0; u < 32; ++u) {
            const uint32_t qb = pq[u];
            const uint32_t te = tnext;
            const uint32_t to = pt[u + 1];
            tnext = to;
            uint32_t hu, e, f, z, nw0, nw1;
            bool kh, kl;
            // even step (r = 64 sup + 2u): left from lane-1, up/diag own
            {
                const uint32_t zl = w.shfl_up(Z1, 1, HB), el = w.shfl_up(E1, 1, HB);
                cell2<0>(zl, el, Z1, F1, Hd, qb, te, m_up, f_up, hu, e, f, z, nw0);
                best = bmax(best, hu, &kh, &kl);
                if (!kl) rbest[0] = r0 + 2 * u;
                if (!kh) rbest[1] = r0 + 2 * u;
                Hd = HU1; HU1 = hu; E1 = e; F1 = f; Z1 = z;
            }
            // odd step (r + 1): up from lane+1, left/diag own
            {
                const uint32_t zu = w.shfl_down(Z1, 1, HB), fu = w.shfl_down(F1, 1, HB);
                cell2<1>(Z1, E1, zu, fu, Hd, qb, to, m_dn, f_dn, hu, e, f, z, nw1);
                best = bmax(best, hu, &kh, &kl);
                if (!kl) rbest[0] = r0 + 2 * u + 1;
                if (!kh) rbest[1] = r0 + 2 * u + 1;
                Hd = HU1; HU1 = hu; E1 = e; F1 = f; Z1 = z;
            }
            fq[u & 3] = nw0 + 16u * nw1;                  // byte 0: two nibbles of pair A, byte 2: pair B
            if ((u & 3) == 3) {
                const uint32_t g01 = prmt(fq[0], fq[1], 0x6240), g23 = prmt(fq[2], fq[3], 0x6240);
                twA[u >> 2] = prmt(g01, g23, 0x5410);
                twB[u >> 2] = prmt(g01, g23, 0x7632);
            }
        }
        // lane-major layout (as hb_extend.cu): the 8 words (64 anti-diagonals) of one band column are one
        // 32 B sector of the pair's own stream
        if (sup < msup[0]) {
            uint4* dst = reinterpret_cast<uint4*>(tbp[0] + ((size_t)sup * HB + sl) * 8);
            dst[0] = make_uint4(twA[0], twA[1], twA[2], twA[3]);
            dst[1] = make_uint4(twA[4], twA[5], twA[6], twA[7]);
        }
        if (sup < msup[1]) {
            uint4* dst = reinterpret_cast<uint4*>(tbp[1] + ((size_t)sup * HB + sl) * 8);
            dst[0] = make_uint4(twB[0], twB[1], twB[2], twB[3]);
            dst[1] = make_uint4(twB[4], twB[5], twB[6], twB[7]);
        }
    }
    // ---- best cell per pair: max score, then smallest i, then smallest j ---------------------------
#pragma unroll
    for (int z = 0; z < 2; ++z) {
        const int score = (int)((best >> (16 * z)) & 0xFFFFu);
        unsigned long long key = 0;
        if (pid[z] != NONE && score != 0) {
            const int s = rbest[z] - HB;
            const int o = 2 * sl - HB + (s & 1);
            const int i = (s - o) >> 1;
            if (i >= 0 && i < L[z])   // rows below the query are never better than the row above them
                key = ((unsigned long long)score << 32) | ((unsigned long long)(0xFFFF - i) << 16) |
                      (unsigned long long)(127 - (o + HB));
        }
#pragma unroll
        for (int off = HB / 2; off; off >>= 1) {
            const unsigned long long t = w.shfl_xor64(key, off);
            key = t > key ? t : key;
        }
        if (pid[z] == NONE || sl != 0) continue;
        HbAln out;
        out.score = (int32_t)(key >> 32);
        out.keep = 0; out.contig = -1; out.ext_len = 0; out.start1 = 0; out.end1 = 0; out.row = -1; out.trow = -1;
        if (out.score <= 0) {
            out.score = 0; out.i_s = -1; out.i_e = -1; out.j_s = 0; out.j_e = 0;
        } else {
            const int bi = 0xFFFF - (int)((key >> 16) & 0xFFFF);
            const int bo = (127 - (int)(key & 0xFFFF)) - HB;
            // the walk back is done by traceback (one thread per pair); leave the end cell here
            out.i_s = bi; out.i_e = bi; out.j_s = j0[z] + bi + bo; out.j_e = j0[z] + bi + bo;
        }
        alns[pid[z]] = out;
    }
}

// ---- traceback of one pair (one thread per pair; the walk is inherently serial) ---------------------
// A 32 B sector holds the nibbles of 64 consecutive anti-diagonals of one band column and a word the
// nibbles of 8; a diagonal move stays in the column and steps two anti-diagonals back, so a word covers
// up to 4 consecutive diagonal moves: they are recognised with one mask operation and retired together
// (alignments are mostly diagonal runs).  Edit records are staged 8 query positions (one 32 B sector)
// at a time instead of one scattered 4 B store per step.
HB_HD int hb_clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
HB_HD void traceback_pair(uint32_t p, const HbLocus* __restrict__ pairs, const uint32_t* __restrict__ clen,
                          const uint32_t* __restrict__ tb_unit_off,
                          uint32_t tb_unit_base, const uint32_t* __restrict__ tbuf,
                          const uint32_t* __restrict__ qrec_unit_off, uint32_t* __restrict__ qrec,
                          HbAln* __restrict__ alns) {
    HbAln a = alns[p];
    if (a.score <= 0) return;
    const HbLocus pr = pairs[p];
    const int64_t j0 = pr.j0;
    const int HB = half_band(pr.dmax - pr.dmin, (int)clen[pr.cand]);
    const uint32_t* tb = tbuf + ((size_t)(tb_unit_off[p] - tb_unit_base) << 7);
    uint32_t* qr = qrec + ((size_t)qrec_unit_off[p] << 5);
    int i = a.i_e, o = (int)(a.j_e - j0) - a.i_e, pending_d = 0;
    int i_s = i, o_s = o;
    int c_sec = -1, c_g8 = -1;
    uint4 s0 = make_uint4(0, 0, 0, 0), s1 = s0;
    uint32_t w = 0;                                       // cached word: anti-diagonals [8 g8, 8 g8 + 8) of the sector
    // nibble of cell (ii, oo); leaves its word in w and its position inside the word in *k
    auto fetch = [&](int ii, int oo, int* k) -> uint32_t {
        const int r = 2 * ii + oo + HB;
        const int sec = (r >> 6) * HB + ((oo + HB) >> 1);
        const int g8 = (r >> 3) & 7;
        if (sec != c_sec) {
            c_sec = sec; c_g8 = -1;
            const uint4* src = reinterpret_cast<const uint4*>(tb + (size_t)sec * 8);
            s0 = src[0]; s1 = src[1];
        }
        if (g8 != c_g8) {
            c_g8 = g8;
            w = g8 == 0 ? s0.x : g8 == 1 ? s0.y : g8 == 2 ? s0.z : g8 == 3 ? s0.w :
                g8 == 4 ? s1.x : g8 == 5 ? s1.y : g8 == 6 ? s1.z : s1.w;
        }
        *k = r & 7;
        return (w >> (4 * (r & 7))) & 15u;
    };
    uint32_t blk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto record = [&](int ii, uint32_t v) {
        blk[ii & 7] = v;
        if ((ii & 7) == 0) {
            uint4* dst = reinterpret_cast<uint4*>(qr + ii);
            dst[0] = make_uint4(blk[0], blk[1], blk[2], blk[3]);
            dst[1] = make_uint4(blk[4], blk[5], blk[6], blk[7]);
#pragma unroll
            for (int z = 0; z < 8; ++z) blk[z] = 0;
        }
    };
    while (i >= 0) {
        int k;
        uint32_t nib = fetch(i, o, &k);
        const uint32_t src = nib & 3u;
        if (src == 0) break;
        if (src == 1) {
            // diagonal run inside the cached word: nibbles k, k-2, ... whose source is "diag"
            const uint32_t x = w ^ 0x11111111u;
            const uint32_t notdiag = (x | (x >> 1)) & 0x11111111u;             // bit 4t: nibble t is not a diag cell
            const uint32_t lanes = ((k & 1) ? 0x10101010u : 0x01010101u) & ((2u << (4 * k)) - 1u);
            const uint32_t stop = notdiag & lanes;                             // same-parity nibbles <= k that end the run
            int n = stop ? (k - ((31 - hb_clz32(stop)) >> 2)) >> 1 : (k >> 1) + 1;
            if (n > i + 1) n = i + 1;
            const uint32_t base = (uint32_t)(o + 32768) | (1u << 25);
            record(i, base | ((uint32_t)(pending_d < 255 ? pending_d : 255) << 16));
            for (int t = 1; t < n; ++t) record(i - t, base);
            pending_d = 0;
            i -= n;
            i_s = i + 1; o_s = o;
        } else if (src == 2) {                            // run of target-only bases
            for (;;) {
                ++pending_d;
                const bool ext = (nib >> 2) & 1u;
                --o;
                if (!ext) break;
                nib = fetch(i, o, &k);
            }
        } else {                                          // run of query-only bases
            for (;;) {
                record(i, (uint32_t)(o + 32768) | ((uint32_t)(pending_d < 255 ? pending_d : 255) << 16) | (1u << 24) | (1u << 25));
                pending_d = 0; i_s = i; o_s = o;
                const bool ext = (nib >> 3) & 1u;
                --i; ++o;
                if (!ext || i < 0) break;
                nib = fetch(i, o, &k);
            }
        }
    }
    if (i >= 0 && (i & 7) != 7) {                         // partial sector holding i_s
        uint4* dst = reinterpret_cast<uint4*>(qr + (i & ~7));
        dst[0] = make_uint4(blk[0], blk[1], blk[2], blk[3]);
        dst[1] = make_uint4(blk[4], blk[5], blk[6], blk[7]);
    }
    a.i_s = i_s; a.j_s = j0 + i_s + o_s;
    alns[p] = a;
}

}  // namespace hb16
