// hite_b200 -- kernel (ii), second half: banded affine Smith-Waterman extension of every
// candidate x locus pair, one warp per pair, anti-diagonal wavefront held in registers and stepped
// by warp shuffle, 4-bit traceback streamed to HBM and walked back by the same warp.
//
// Replaces the base-level alignment stage of `minimap2 -ax map-ont` (reference
// module/Util.py:7952-7961) whose CIGAR feeds get_copies_minimap2 (module/Util.py:7977-8030);
// follows HB-align S7, bit-identical to oracle/hbo_align.c::hbo_align_pair.
//
// Geometry: query row i, oriented target column j = j0 + j', band offset o = j' - i in [-32, 31].
// Anti-diagonal s = i + j' = 2i + o; lane l owns o = 2l - 32 + (s & 1), so
//   even s: left neighbour comes from lane l-1, up neighbour from the same lane,
//   odd  s: left neighbour from the same lane, up neighbour from lane l+1,
//   diagonal neighbour: same lane two steps earlier.
// Integer DP (no tensor cores).  Scores: match +2, mismatch/N -4, gap open 4, gap extend 2.
#include "hb_common.cuh"
#include "hb_extend.cuh"

namespace {

constexpr int EXT_WARPS = 8;
constexpr int NEG6 = -6;

__device__ __forceinline__ int packed_base(const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask,
                                           uint64_t p) {
    if ((nmask[p >> 5] >> (p & 31)) & 1u) return 4;
    return (seq2[p >> 4] >> (2 * (p & 15))) & 3u;
}
__device__ __forceinline__ int target_base(const uint32_t* __restrict__ gseq2, const uint32_t* __restrict__ gnmask,
                                           uint64_t tot, int rel, int64_t j) {
    if (j < 0 || (uint64_t)j >= tot) return 4;
    if (!rel) return packed_base(gseq2, gnmask, (uint64_t)j);
    const int c = packed_base(gseq2, gnmask, tot - 1 - (uint64_t)j);
    return c < 4 ? 3 - c : 4;
}

__global__ void __launch_bounds__(EXT_WARPS * 32)
extend_kernel(const HbLocus* __restrict__ pairs, uint32_t pair_lo, uint32_t pair_hi,
              const uint32_t* __restrict__ gseq2, const uint32_t* __restrict__ gnmask, uint64_t tot,
              const uint32_t* __restrict__ cseq2, const uint32_t* __restrict__ cnmask,
              const uint64_t* __restrict__ cgoff, const uint32_t* __restrict__ clen,
              const uint32_t* __restrict__ tb_unit_off, uint32_t tb_unit_base, uint32_t* __restrict__ tbuf,
              const uint32_t* __restrict__ qrec_unit_off, uint32_t* __restrict__ qrec,
              HbAln* __restrict__ alns) {
    __shared__ uint8_t s_rq[EXT_WARPS][128];
    __shared__ uint8_t s_rt[EXT_WARPS][128];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t p = pair_lo + blockIdx.x * EXT_WARPS + wid;
    if (p >= pair_hi) return;
    const HbLocus pr = pairs[p];
    const int L = (int)clen[pr.cand];
    const uint64_t cstart = cgoff[pr.cand];
    const int rel = (int)pr.rel;
    const int64_t j0 = pr.j0;
    uint8_t* rq = s_rq[wid];
    uint8_t* rt = s_rt[wid];
    uint32_t* tb = tbuf + ((size_t)(tb_unit_off[p] - tb_unit_base) << 8);      // 256 words per unit
    const int n_super = (2 * L + 62 + 63) >> 6;

    auto fill = [&](int x) {
        rq[x & 127] = (uint8_t)((x >= 0 && x < L) ? packed_base(cseq2, cnmask, cstart + (uint64_t)x) : 4);
        rt[x & 127] = (uint8_t)target_base(gseq2, gnmask, tot, rel, j0 + x);
    };
    fill(-48 + lane);
    fill(-16 + lane);

    int H1 = 0, E1 = NEG6, F1 = NEG6, H2 = 0;
    int best_h = 0, best_r = -1;
    uint32_t tbword = 0;
    for (int sup = 0; sup < n_super; ++sup) {
        const int half0 = -16 + 32 * sup;          // s0 / 2
        fill(half0 + 16 + lane);
        __syncwarp();
#pragma unroll 4
        for (int u = 0; u < 32; ++u) {
            const int half = half0 + u;            // s/2 for the even step
            const int i = half - lane + 16;
            const int jp = half + lane - 16;
            const int r = 64 * sup + 2 * u;        // s + 32
            const bool row_ok = (i >= 0 && i < L);
            const int qb = rq[i & 127];
            // ---------------- even step: cell (i, jp) ----------------
            {
                const int tbse = rt[jp & 127];
                uint32_t v = ((uint32_t)H1 << 16) | (uint32_t)(E1 + 8);
                uint32_t rv = __shfl_up_sync(HB_FULL, v, 1);
                if (lane == 0) rv = 2u;                                   // H = 0, E = -6
                const int hl = (int)(rv >> 16), el = (int)(rv & 0xFFFF) - 8;
                const int hu = H1, fu = F1, hd = H2;
                const int eo = hl - 4, fo = hu - 4;
                const bool eext = el > eo, fext = fu > fo;
                int e = (eext ? el : eo) - 2;
                int f = (fext ? fu : fo) - 2;
                const int sc = (qb < 4 && qb == tbse) ? 2 : -4;
                const int dg = hd + sc;
                int h = 0, src = 0;
                if (dg > h) { h = dg; src = 1; }
                if (e > h) { h = e; src = 2; }
                if (f > h) { h = f; src = 3; }
                const int64_t j = j0 + jp;
                const bool act = row_ok && j >= 0 && (uint64_t)j < tot;
                uint32_t nib = (uint32_t)src | (eext ? 4u : 0u) | (fext ? 8u : 0u);
                if (!act) { h = 0; e = NEG6; f = NEG6; nib = 0; }
                if (h > best_h) { best_h = h; best_r = r; }
                tbword |= nib << (4 * (r & 7));
                H2 = H1; H1 = h; E1 = e; F1 = f;
            }
            // ---------------- odd step: cell (i, jp + 1) ----------------
            {
                const int tbse = rt[(jp + 1) & 127];
                uint32_t v = ((uint32_t)H1 << 16) | (uint32_t)(F1 + 8);
                uint32_t rv = __shfl_down_sync(HB_FULL, v, 1);
                if (lane == 31) rv = 2u;
                const int hu = (int)(rv >> 16), fu = (int)(rv & 0xFFFF) - 8;
                const int hl = H1, el = E1, hd = H2;
                const int eo = hl - 4, fo = hu - 4;
                const bool eext = el > eo, fext = fu > fo;
                int e = (eext ? el : eo) - 2;
                int f = (fext ? fu : fo) - 2;
                const int sc = (qb < 4 && qb == tbse) ? 2 : -4;
                const int dg = hd + sc;
                int h = 0, src = 0;
                if (dg > h) { h = dg; src = 1; }
                if (e > h) { h = e; src = 2; }
                if (f > h) { h = f; src = 3; }
                const int64_t j = j0 + jp + 1;
                const bool act = row_ok && j >= 0 && (uint64_t)j < tot;
                uint32_t nib = (uint32_t)src | (eext ? 4u : 0u) | (fext ? 8u : 0u);
                if (!act) { h = 0; e = NEG6; f = NEG6; nib = 0; }
                if (h > best_h) { best_h = h; best_r = r + 1; }
                tbword |= nib << (4 * ((r + 1) & 7));
                H2 = H1; H1 = h; E1 = e; F1 = f;
            }
            if ((u & 3) == 3) {
                tb[(size_t)(r >> 3) * 32 + lane] = tbword;               // 128 B per warp, coalesced
                tbword = 0;
            }
        }
        __syncwarp();
    }
    // ---- best cell: max score, then smallest i, then smallest j ------------------------------
    unsigned long long key = 0;
    if (best_r >= 0) {
        const int s = best_r - 32;
        const int o = 2 * lane - 32 + (s & 1);
        const int i = (s - o) >> 1;
        key = ((unsigned long long)(uint32_t)best_h << 32) | ((unsigned long long)(0xFFFF - i) << 16) |
              (unsigned long long)(63 - (o + 32));
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        unsigned long long t = __shfl_xor_sync(HB_FULL, key, off);
        key = t > key ? t : key;
    }
    HbAln out;
    out.score = (int32_t)(key >> 32);
    out.keep = 0; out.contig = -1; out.ext_len = 0; out.start1 = 0; out.end1 = 0; out.row = -1; out.pad = 0;
    if (out.score <= 0) {
        out.score = 0; out.i_s = -1; out.i_e = -1; out.j_s = 0; out.j_e = 0;
        if (lane == 0) alns[p] = out;
        return;
    }
    const int bi = 0xFFFF - (int)((key >> 16) & 0xFFFF);
    const int bo = (63 - (int)(key & 0xFFFF)) - 32;
    // ---- traceback (warp-uniform walk; 128 B line of 8 steps held across the lanes) ------------
    __syncwarp();
    uint32_t* qr = qrec + ((size_t)qrec_unit_off[p] << 5);
    int i = bi, o = bo, st = 0, pending_d = 0;
    int i_s = bi; int64_t j_s = j0 + bi + bo;
    int g_cur = -1;
    uint32_t w_cur = 0, w_next = 0;
    uint32_t qv = 0;
    for (;;) {
        if (i < 0 || o < -32 || o > 31) break;
        const int r = 2 * i + o + 32;
        const int g = r >> 3;
        if (g != g_cur) {
            if (g == g_cur - 1) w_cur = w_next; else w_cur = tb[(size_t)g * 32 + lane];
            g_cur = g;
            if (g > 0) w_next = tb[(size_t)(g - 1) * 32 + lane];
        }
        const uint32_t nib = (__shfl_sync(HB_FULL, w_cur, (o + 32) >> 1) >> (4 * (r & 7))) & 15u;
        if (st == 0) {
            const int src = (int)(nib & 3u);
            if (src == 0) break;
            if (src == 1) {
                if (lane == (i & 31)) qv = (uint32_t)(o + 32768) | ((uint32_t)min(pending_d, 255) << 16) | (1u << 25);
                pending_d = 0; i_s = i; j_s = j0 + i + o;
                if ((i & 31) == 0) { qr[(size_t)(i >> 5) * 32 + lane] = qv; qv = 0; }
                --i;
            } else if (src == 2) st = 1;
            else st = 2;
        } else if (st == 1) {
            ++pending_d;
            st = ((nib >> 2) & 1u) ? 1 : 0;
            --o;
        } else {
            if (lane == (i & 31)) qv = (uint32_t)(o + 32768) | ((uint32_t)min(pending_d, 255) << 16) | (1u << 24) | (1u << 25);
            pending_d = 0; i_s = i;
            if ((i & 31) == 0) { qr[(size_t)(i >> 5) * 32 + lane] = qv; qv = 0; }
            st = ((nib >> 3) & 1u) ? 2 : 0;
            --i; ++o;
        }
    }
    if (i >= 0 && (i & 31) != 31) qr[(size_t)(i >> 5) * 32 + lane] = qv;   // partial block holding i_s
    else if (i < 0) { /* i_s == 0 was flushed at (i & 31) == 0 */ }
    out.i_s = i_s; out.i_e = bi; out.j_s = j_s; out.j_e = j0 + bi + bo;
    if (lane == 0) alns[p] = out;
}

// sizes: tb units (1 KB = 64 steps x 32 lanes x 4 bit) and qrec units (32 words) per pair
__global__ void pair_sizes_kernel(const HbLocus* __restrict__ pairs, uint32_t n_pairs, const uint32_t* __restrict__ clen,
                                  uint32_t* __restrict__ tb_units, uint32_t* __restrict__ qrec_units) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const uint32_t L = clen[pairs[p].cand];
    tb_units[p] = (2 * L + 62 + 63) >> 6;
    qrec_units[p] = (L + 31) >> 5;
}

}  // namespace

cudaError_t hb_pair_sizes_launch(const HbLocus* pairs, uint32_t n_pairs, const uint32_t* clen, uint32_t* tb_units,
                                 uint32_t* qrec_units, cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    pair_sizes_kernel<<<(n_pairs + 255) / 256, 256, 0, s>>>(pairs, n_pairs, clen, tb_units, qrec_units);
    return cudaGetLastError();
}

cudaError_t hb_extend_launch(const HbLocus* pairs, uint32_t pair_lo, uint32_t pair_hi, const uint32_t* gseq2,
                             const uint32_t* gnmask, uint64_t tot, const uint32_t* cseq2, const uint32_t* cnmask,
                             const uint64_t* cgoff, const uint32_t* clen, const uint32_t* tb_unit_off,
                             uint32_t tb_unit_base, uint32_t* tbuf, const uint32_t* qrec_unit_off, uint32_t* qrec,
                             HbAln* alns, cudaStream_t s) {
    if (pair_hi <= pair_lo) return cudaSuccess;
    const uint32_t n = pair_hi - pair_lo;
    extend_kernel<<<(n + EXT_WARPS - 1) / EXT_WARPS, EXT_WARPS * 32, 0, s>>>(
        pairs, pair_lo, pair_hi, gseq2, gnmask, tot, cseq2, cnmask, cgoff, clen, tb_unit_off, tb_unit_base, tbuf,
        qrec_unit_off, qrec, alns);
    return cudaGetLastError();
}
