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
//
// Two implementations of the same recurrence live here:
//   * extend16_kernel (body in hb_extend16.cuh): two pairs per lane as s16x2 halves on the DPX
//     instructions -- every candidate of at most HB16_MAX_L bases, i.e. practically all of them;
//   * extend_kernel (below): one pair per lane in 32-bit registers, for longer candidates.
// Pairs are routed by class through the order lists built by hb_extend_order_launch.
#include "hb_common.cuh"
#include "hb_extend.cuh"
#include "hb_extend16.cuh"
#include "hb_prims.cuh"

namespace {

constexpr int EXT_WARPS = 8;
constexpr int HB_NARROW_SPREAD = hb16::NARROW_SPREAD;      // HB-align S7

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

// One DP cell.  H >= 0; E and F are kept clamped at 0 (only positive gap states can ever be chosen
// by H or walked by the traceback), which lets every recurrence be a single
// 3-input max (VIMNMX3) and the left/up neighbour travel as one packed 32-bit shuffle.
//   nib: bits0-1 source of H (0 stop, 1 diag, 2 E, 3 F), bit2 E extended, bit3 F extended
#define HB_CELL(hl, el, hu, fu, hd, qb, tbse, H, E, F, NIB)                         \
    {                                                                               \
        const int t_e = (el) - 2, hm_e = (hl) - 6;                                  \
        const int t_f = (fu) - 2, hm_f = (hu) - 6;                                  \
        E = __vimax3_s32(t_e, hm_e, 0);                                             \
        F = __vimax3_s32(t_f, hm_f, 0);                                             \
        const int dg = (hd) + (((qb) == (tbse)) ? 2 : -4);                          \
        H = __vimax3_s32_relu(dg, E, F);                                            \
        const uint32_t src = (dg == H) ? 1u : ((E == H) ? 2u : 3u);                 \
        const uint32_t fl = src | ((t_e > hm_e) ? 4u : 0u) | ((t_f > hm_f) ? 8u : 0u); \
        NIB = (H > 0) ? fl : 0u;                                                    \
    }

// HB = half band = lanes per pair: 32 (o in [-32,31], one pair per warp) or 16 (o in [-16,15], two
// pairs per warp) -- HB-align S7: the narrow band is used when the seed cluster's diagonal spread is
// <= HB_NARROW_SPREAD.  Traceback stream: word ((r>>6)*HB + sublane)*8 + ((r>>3)&7), nibble r&7, 128-word units.
template <int HB>
__global__ void __launch_bounds__(EXT_WARPS * 32)
extend_kernel(const HbLocus* __restrict__ pairs, uint32_t pair_lo, uint32_t pair_hi,
              const uint32_t* __restrict__ gseq2, const uint32_t* __restrict__ gnmask, uint64_t tot,
              const uint32_t* __restrict__ cseq2, const uint32_t* __restrict__ cnmask,
              const uint64_t* __restrict__ cgoff, const uint32_t* __restrict__ clen,
              const uint32_t* __restrict__ tb_unit_off, uint32_t tb_unit_base, uint32_t* __restrict__ tbuf,
              HbAln* __restrict__ alns) {
    constexpr int PPW = 32 / HB;                 // pairs per warp
    constexpr int WIN = HB + 32;                 // sequence window of one super-step (64 anti-diagonals)
    __shared__ uint8_t s_rq[EXT_WARPS * PPW][WIN];
    __shared__ uint8_t s_rt[EXT_WARPS * PPW][WIN];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / HB, sl = lane % HB;   // which pair of the warp, lane inside the pair
    const uint32_t p = pair_lo + (blockIdx.x * EXT_WARPS + wid) * PPW + sub;
    bool active = p < pair_hi;
    HbLocus pr;
    pr.cand = 0; pr.rel = 0; pr.j0 = 0; pr.dmin = 0; pr.dmax = 0; pr.nseeds = 0; pr.rank = 0;
    if (active) {
        pr = pairs[p];
        active = (int)clen[pr.cand] > hb16::HB16_MAX_L && hb16::half_band(pr.dmax - pr.dmin, (int)clen[pr.cand]) == HB;
    }
    if (!__ballot_sync(HB_FULL, active)) return;
    const int L = active ? (int)clen[pr.cand] : 0;
    const uint64_t cstart = active ? cgoff[pr.cand] : 0;
    const int rel = (int)pr.rel;
    const int64_t j0 = pr.j0;
    uint8_t* rq = s_rq[wid * PPW + sub];
    uint8_t* rt = s_rt[wid * PPW + sub];
    uint32_t* tb = tbuf + (active ? ((size_t)(tb_unit_off[p] - tb_unit_base) << 7) : 0);   // 128 words per unit
    const int my_super = active ? ((2 * L + 2 * HB - 2 + 63) >> 6) : 0;
    int n_super = my_super;
    if (PPW == 2) n_super = max(__shfl_sync(HB_FULL, my_super, 0), __shfl_sync(HB_FULL, my_super, 16));

    // query N = 4, target N = 5: never equal, so one compare scores a cell
    auto qsym = [&](int x) -> uint8_t {
        return (uint8_t)((active && x >= 0 && x < L) ? packed_base(cseq2, cnmask, cstart + (uint64_t)x) : 4);
    };
    auto tsym = [&](int x) -> uint8_t {
        if (!active) return 5;
        const int c = target_base(gseq2, gnmask, tot, rel, j0 + x);
        return (uint8_t)(c < 4 ? c : 5);
    };
    // window of the first super-step (half0 = -HB/2): positions [-HB, 32)
    for (int t = sl; t < WIN; t += HB) { rq[t] = qsym(-HB + t); rt[t] = tsym(-HB + t); }
    __syncwarp();

    int H1 = 0, E1 = 0, F1 = 0, H2 = 0;
    uint32_t best = 0;                                   // (h << 16) | (65535 - r): max = best score, earliest r
    uint32_t tbword = 0;
    const uint32_t m_up = sl == 0 ? 0u : 0xffffffffu;    // band edge: neighbour outside = (H 0, E/F 0)
    const uint32_t m_dn = sl == HB - 1 ? 0u : 0xffffffffu;
    const uint8_t* pq = rq + (HB - sl);                  // query index of pair-step u:  u - sl + HB
    const uint8_t* pt = rt + sl;                         // target index: u + sl (even), u + sl + 1 (odd)
    for (int sup = 0; sup < n_super; ++sup) {
        if (sup > 0) {
            const int half0 = -HB / 2 + 32 * sup;
            const uint8_t a = rq[32 + sl], b = rt[32 + sl];
            uint8_t nq[32 / HB], nt[32 / HB];
#pragma unroll
            for (int z = 0; z < 32 / HB; ++z) { nq[z] = qsym(half0 + HB / 2 + sl + z * HB); nt[z] = tsym(half0 + HB / 2 + sl + z * HB); }
            __syncwarp();
            rq[sl] = a; rt[sl] = b;
#pragma unroll
            for (int z = 0; z < 32 / HB; ++z) { rq[HB + sl + z * HB] = nq[z]; rt[HB + sl + z * HB] = nt[z]; }
            __syncwarp();
        }
        const uint32_t rbase = 65535u - (uint32_t)(64 * sup);
        const bool store = sup < my_super;
        uint32_t tw[8];                                   // the 8 traceback words of this super-step
        int tnext = pt[0];
#pragma unroll 4
        for (int u = 0; u < 32; ++u) {
            const int qb = pq[u];
            const int te = tnext;
            const int to = pt[u + 1];
            tnext = to;
            int h, e, f; uint32_t nib;
            // even step (r = 64 sup + 2u): left from lane-1, up/diag own
            {
                uint32_t rv = __shfl_up_sync(HB_FULL, __byte_perm((uint32_t)E1, (uint32_t)H1, 0x5410), 1, HB) & m_up;
                const int hl = (int)(rv >> 16), el = (int)(rv & 0xFFFFu);
                HB_CELL(hl, el, H1, F1, H2, qb, te, h, e, f, nib);
                best = max(best, ((uint32_t)h << 16) | (rbase - (uint32_t)(2 * u)));
                tbword |= nib << (4 * ((2 * u) & 7));
                H2 = H1; H1 = h; E1 = e; F1 = f;
            }
            // odd step (r + 1): up from lane+1, left/diag own
            {
                uint32_t rv = __shfl_down_sync(HB_FULL, __byte_perm((uint32_t)F1, (uint32_t)H1, 0x5410), 1, HB) & m_dn;
                const int hu = (int)(rv >> 16), fu = (int)(rv & 0xFFFFu);
                HB_CELL(H1, E1, hu, fu, H2, qb, to, h, e, f, nib);
                best = max(best, ((uint32_t)h << 16) | (rbase - (uint32_t)(2 * u + 1)));
                tbword |= nib << (4 * ((2 * u + 1) & 7));
                H2 = H1; H1 = h; E1 = e; F1 = f;
            }
            if ((u & 3) == 3) { tw[u >> 2] = tbword; tbword = 0; }
        }
        // lane-major layout: the 8 words (64 anti-diagonals) of one band column are one 32 B sector, so
        // the traceback thread that walks down a diagonal reads one sector per 32 steps
        if (store) {
            uint4* dst = reinterpret_cast<uint4*>(tb + ((size_t)sup * HB + sl) * 8);
            dst[0] = make_uint4(tw[0], tw[1], tw[2], tw[3]);
            dst[1] = make_uint4(tw[4], tw[5], tw[6], tw[7]);
        }
    }
    // ---- best cell: max score, then smallest i, then smallest j ------------------------------
    unsigned long long key = 0;
    if (active && (best >> 16) != 0) {
        const int r = 65535 - (int)(best & 0xFFFFu);
        const int s = r - HB;
        const int o = 2 * sl - HB + (s & 1);
        const int i = (s - o) >> 1;
        if (i >= 0 && i < L)      // rows below the query are never better than the row above them
            key = ((unsigned long long)(best >> 16) << 32) | ((unsigned long long)(0xFFFF - i) << 16) |
                  (unsigned long long)(127 - (o + HB));
    }
#pragma unroll
    for (int off = HB / 2; off; off >>= 1) {
        unsigned long long t = __shfl_xor_sync(HB_FULL, key, off);
        key = t > key ? t : key;
    }
    if (!active) return;
    HbAln out;
    out.score = (int32_t)(key >> 32);
    out.keep = 0; out.contig = -1; out.ext_len = 0; out.start1 = 0; out.end1 = 0; out.row = -1; out.trow = -1;
    if (out.score <= 0) {
        out.score = 0; out.i_s = -1; out.i_e = -1; out.j_s = 0; out.j_e = 0;
        if (sl == 0) alns[p] = out;
        return;
    }
    const int bi = 0xFFFF - (int)((key >> 16) & 0xFFFF);
    const int bo = (127 - (int)(key & 0xFFFF)) - HB;
    // the walk back is done by traceback_kernel (one thread per pair); leave the end cell here
    out.i_s = bi; out.i_e = bi; out.j_s = j0 + bi + bo; out.j_e = j0 + bi + bo;
    if (sl == 0) alns[p] = out;
}

// ---- two pairs per lane (hb_extend16.cuh) -----------------------------------------------------------
struct DevWarp {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ uint32_t shfl_up(uint32_t v, int d, int width) const { return __shfl_up_sync(HB_FULL, v, d, width); }
    __device__ __forceinline__ uint32_t shfl_down(uint32_t v, int d, int width) const { return __shfl_down_sync(HB_FULL, v, d, width); }
    __device__ __forceinline__ unsigned long long shfl_xor64(unsigned long long v, int off) const { return __shfl_xor_sync(HB_FULL, v, off); }
    __device__ __forceinline__ int max_all(int v) const { return __reduce_max_sync(HB_FULL, v); }
};

template <int HB>
__global__ void __launch_bounds__(EXT_WARPS * 32, 4)
extend16_kernel(const uint32_t* __restrict__ order, const uint64_t* __restrict__ n_order_p,
                const HbLocus* __restrict__ pairs, const uint32_t* __restrict__ gseq2,
                const uint32_t* __restrict__ gnmask, uint64_t tot, const uint32_t* __restrict__ cseq2,
                const uint32_t* __restrict__ cnmask, const uint64_t* __restrict__ cgoff,
                const uint32_t* __restrict__ clen, const uint32_t* __restrict__ tb_unit_off, uint32_t tb_unit_base,
                uint32_t* __restrict__ tbuf, HbAln* __restrict__ alns) {
    constexpr int PPW = 32 / HB, WIN = HB + 32;
    __shared__ uint32_t s_rq[EXT_WARPS * PPW * WIN];
    __shared__ uint32_t s_rt[EXT_WARPS * PPW * WIN];
    const uint32_t n_order = (uint32_t)*n_order_p;
    const int wid = threadIdx.x >> 5;
    const uint32_t group0 = (blockIdx.x * EXT_WARPS + wid) * PPW;
    if (2ull * group0 >= n_order) return;                 // the grid is sized for the worst case
    DevWarp w;
    hb16::extend16_warp<HB>(w, group0, order, n_order, pairs, gseq2, gnmask, tot, cseq2, cnmask, cgoff, clen,
                            tb_unit_off, tb_unit_base, tbuf, alns, s_rq + wid * PPW * WIN, s_rt + wid * PPW * WIN);
}

// class of every pair of [pair_lo, pair_hi): flag arrays for the three order lists (tight / narrow / wide band)
__global__ void order_flag_kernel(const HbLocus* __restrict__ pairs, uint32_t pair_lo, uint32_t pair_hi,
                                  const uint32_t* __restrict__ clen, uint32_t* __restrict__ flag_t, uint32_t* __restrict__ flag_n,
                                  uint32_t* __restrict__ flag_w) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair_lo + t >= pair_hi) return;
    const HbLocus pr = pairs[pair_lo + t];
    const int L = (int)clen[pr.cand];
    const bool fits = L <= hb16::HB16_MAX_L;
    const int hb = hb16::half_band(pr.dmax - pr.dmin, L);
    flag_t[t] = (fits && hb == 8) ? 1u : 0u;
    flag_n[t] = (fits && hb == 16) ? 1u : 0u;
    flag_w[t] = (fits && hb == 32) ? 1u : 0u;
}
__global__ void order_scatter_kernel(uint32_t pair_lo, uint32_t n, const uint32_t* __restrict__ flag_t,
                                     const uint32_t* __restrict__ flag_n, const uint32_t* __restrict__ flag_w,
                                     const uint32_t* __restrict__ off_t, const uint32_t* __restrict__ off_n,
                                     const uint32_t* __restrict__ off_w, uint32_t* __restrict__ order_t,
                                     uint32_t* __restrict__ order_n, uint32_t* __restrict__ order_w) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (flag_t[t]) order_t[off_t[t]] = pair_lo + t;
    if (flag_n[t]) order_n[off_n[t]] = pair_lo + t;
    if (flag_w[t]) order_w[off_w[t]] = pair_lo + t;
}

// Traceback: ONE THREAD per pair (the walk is inherently serial; 32 independent walks per warp keep
// the issue slots full where a warp-per-pair walk leaves 31 lanes idle).
__global__ void __launch_bounds__(128)
traceback_kernel(const HbLocus* __restrict__ pairs, uint32_t pair_lo, uint32_t pair_hi, const uint32_t* __restrict__ clen,
                 const uint32_t* __restrict__ tb_unit_off, uint32_t tb_unit_base, const uint32_t* __restrict__ tbuf,
                 const uint32_t* __restrict__ qrec_unit_off, uint32_t* __restrict__ qrec, HbAln* __restrict__ alns) {
    const uint32_t p = pair_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= pair_hi) return;
    hb16::traceback_pair(p, pairs, clen, tb_unit_off, tb_unit_base, tbuf, qrec_unit_off, qrec, alns);
}

// sizes: tb units (512 B = 64 steps x 16 lanes x 4 bit) and qrec units (32 words) per pair
__global__ void pair_sizes_kernel(const HbLocus* __restrict__ pairs, uint32_t n_pairs, const uint32_t* __restrict__ clen,
                                  uint32_t* __restrict__ tb_units, uint32_t* __restrict__ qrec_units) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const HbLocus pr = pairs[p];
    const uint32_t L = clen[pr.cand];
    const uint32_t hb = (uint32_t)hb16::half_band(pr.dmax - pr.dmin, (int)L);
    tb_units[p] = (((2 * L + 2 * hb - 2 + 63) >> 6) * hb + 15) / 16;  // 128-word units: a super-step is hb x 8 words
    qrec_units[p] = (L + 31) >> 5;
}

}  // namespace

cudaError_t hb_pair_sizes_launch(const HbLocus* pairs, uint32_t n_pairs, const uint32_t* clen, uint32_t* tb_units,
                                 uint32_t* qrec_units, cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    pair_sizes_kernel<<<(n_pairs + 255) / 256, 256, 0, s>>>(pairs, n_pairs, clen, tb_units, qrec_units);
    return cudaGetLastError();
}

size_t hb_extend_order_words(uint32_t n_pairs) {
    // flag_t/n/w, off_t/n/w, order_t/n/w (n each), three totals (8 words), scan scratch
    return 9 * ((size_t)n_pairs + 2) + 8 + hb_scan_tmp_words((size_t)n_pairs + 1);
}

cudaError_t hb_extend_launch(const HbLocus* pairs, uint32_t pair_lo, uint32_t pair_hi, const uint32_t* gseq2,
                             const uint32_t* gnmask, uint64_t tot, const uint32_t* cseq2, const uint32_t* cnmask,
                             const uint64_t* cgoff, const uint32_t* clen, const uint32_t* tb_unit_off,
                             uint32_t tb_unit_base, uint32_t* tbuf, const uint32_t* qrec_unit_off, uint32_t* qrec,
                             HbAln* alns, uint32_t* order_ws, int any_long, uint32_t* launches, cudaStream_t s) {
    if (pair_hi <= pair_lo) return cudaSuccess;
    const uint32_t n = pair_hi - pair_lo;
    const size_t st = (size_t)n + 2;
    uint32_t *flag_t = order_ws, *flag_n = flag_t + st, *flag_w = flag_n + st, *off_t = flag_w + st, *off_n = off_t + st,
             *off_w = off_n + st, *order_t = off_w + st, *order_n = order_t + st, *order_w = order_n + st;
    uint64_t* totals = reinterpret_cast<uint64_t*>(order_ws + ((9 * st + 1) & ~(size_t)1));     // 8 B aligned
    uint32_t* scan_tmp = reinterpret_cast<uint32_t*>(totals + 3);
    // ---- order lists: tight / narrow / wide pairs whose candidate fits the 16-bit kernel ----------
    order_flag_kernel<<<(n + 255) / 256, 256, 0, s>>>(pairs, pair_lo, pair_hi, clen, flag_t, flag_n, flag_w);
    cudaError_t e = hb_exclusive_scan_u32(flag_t, off_t, n, scan_tmp, totals + 2, s);
    if (e != cudaSuccess) return e;
    e = hb_exclusive_scan_u32(flag_n, off_n, n, scan_tmp, totals, s);
    if (e != cudaSuccess) return e;
    e = hb_exclusive_scan_u32(flag_w, off_w, n, scan_tmp, totals + 1, s);
    if (e != cudaSuccess) return e;
    order_scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(pair_lo, n, flag_t, flag_n, flag_w, off_t, off_n, off_w, order_t, order_n, order_w);
    // ---- 16-bit kernels: grids sized for "all pairs in this class", surplus blocks exit at once ----
    extend16_kernel<32><<<(n + 2 * EXT_WARPS - 1) / (2 * EXT_WARPS), EXT_WARPS * 32, 0, s>>>(
        order_w, totals + 1, pairs, gseq2, gnmask, tot, cseq2, cnmask, cgoff, clen, tb_unit_off, tb_unit_base, tbuf, alns);
    extend16_kernel<16><<<(n + 4 * EXT_WARPS - 1) / (4 * EXT_WARPS), EXT_WARPS * 32, 0, s>>>(
        order_n, totals, pairs, gseq2, gnmask, tot, cseq2, cnmask, cgoff, clen, tb_unit_off, tb_unit_base, tbuf, alns);
    extend16_kernel<8><<<(n + 8 * EXT_WARPS - 1) / (8 * EXT_WARPS), EXT_WARPS * 32, 0, s>>>(
        order_t, totals + 2, pairs, gseq2, gnmask, tot, cseq2, cnmask, cgoff, clen, tb_unit_off, tb_unit_base, tbuf, alns);
    *launches += 4;
    *launches += 10;
    if (any_long) {
        // 32-bit kernels: one pair per warp (wide band) / two pairs per warp (narrow); each skips the
        // pairs of the other class and those the 16-bit kernels took
        extend_kernel<32><<<(n + EXT_WARPS - 1) / EXT_WARPS, EXT_WARPS * 32, 0, s>>>(
            pairs, pair_lo, pair_hi, gseq2, gnmask, tot, cseq2, cnmask, cgoff, clen, tb_unit_off, tb_unit_base, tbuf, alns);
        extend_kernel<16><<<(n + 2 * EXT_WARPS - 1) / (2 * EXT_WARPS), EXT_WARPS * 32, 0, s>>>(
            pairs, pair_lo, pair_hi, gseq2, gnmask, tot, cseq2, cnmask, cgoff, clen, tb_unit_off, tb_unit_base, tbuf, alns);
        *launches += 2;
    }
    traceback_kernel<<<(n + 127) / 128, 128, 0, s>>>(pairs, pair_lo, pair_hi, clen, tb_unit_off, tb_unit_base, tbuf,
                                                      qrec_unit_off, qrec, alns);
    *launches += 1;
    return cudaGetLastError();
}
