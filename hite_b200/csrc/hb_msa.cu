// hite_b200 -- copy selection and candidate-anchored star-MSA assembly from the pairwise edit
// scripts of kernel (ii); feeds the column-vote kernel (iii).
//
// Replaces (reference module/Util.py):
//   get_copies_minimap2 coverage rules            :8003-8025   -> copy_filter_kernel
//   copy extension / off-contig / <100 bp skips   :8095-8115   -> copy_filter_kernel
//   "100 longest members" (ready_for_MSA.sh)      :10409-10412 -> copy_select_kernel
//   mafft --preservecase                          :10416       -> ins_max / layout / msa_write kernels
// Follows HB-align S8..S10; bit-identical to oracle/hbo_align.c::hbo_candidate_msa.
#include "hb_common.cuh"
#include "hb_msa.cuh"

namespace {

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

// S8: thread per pair
__global__ void copy_filter_kernel(const HbLocus* __restrict__ pairs, uint32_t n_pairs, const uint32_t* __restrict__ clen,
                                   const uint64_t* __restrict__ goff, const uint64_t* __restrict__ glen, uint32_t n_contig,
                                   uint64_t tot, int flank, HbAln* __restrict__ alns) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    HbAln a = alns[p];
    a.keep = 0; a.contig = -1; a.row = -1; a.trow = -1;
    if (a.score > 0) {
        const int L = (int)clen[pairs[p].cand];
        const int aligned = a.i_e - a.i_s + 1;
        const long long tlen = a.j_e - a.j_s + 1;
        const bool cov = ((double)aligned / (double)L >= 0.95) && ((double)aligned / (double)tlen >= 0.95);   // :8007-8022
        if (cov) {
            const int rel = (int)pairs[p].rel;
            const int64_t fs = rel ? (int64_t)tot - 1 - a.j_e : a.j_s;
            const int64_t fe = rel ? (int64_t)tot - 1 - a.j_s : a.j_e;
            int lo = 0, hi = (int)n_contig;
            while (lo < hi) { int mid = (lo + hi) >> 1; if ((int64_t)goff[mid] <= fs) lo = mid + 1; else hi = mid; }
            const int c = lo - 1;
            if (c >= 0 && (uint64_t)fe < goff[c] + glen[c]) {
                a.contig = c;
                a.start1 = fs - (int64_t)goff[c] + 1;
                a.end1 = fe - (int64_t)goff[c] + 1;
                const bool off = (a.start1 - 1 - flank < 0) || (a.end1 + flank > (int64_t)glen[c]);            // :8103
                a.ext_len = (int32_t)(a.end1 - a.start1 + 1 + 2 * flank);
                if (!off && a.ext_len >= 100) a.keep = 1;                                                       // :8108
            }
        }
    }
    alns[p] = a;
}

// Library-vs-genome HSPs (hb_query_hsps): every positive local alignment of a query x locus pair as the outfmt-6
// fields get_copies_v1 reads (module/Util.py:7032-7056): 1-based inclusive query / subject spans (minus strand as
// s_start > s_end), matches and alignment length (pident = 100 * n_match / aln_len).  Warp per pair over the
// traceback's per-query-base edit record: word & 0xFFFF = diagonal drift + 32768, bits 16-23 = target-only bases
// after the query base, bit 24 = query base without a target base.
__global__ void __launch_bounds__(256)
qhsp_kernel(const HbLocus* __restrict__ pairs, uint32_t n_pairs, const HbAln* __restrict__ alns,
            const uint32_t* __restrict__ qrec_unit_off, const uint32_t* __restrict__ qrec, const uint32_t* __restrict__ cseq2,
            const uint32_t* __restrict__ cnmask, const uint64_t* __restrict__ cgoff, const uint32_t* __restrict__ gseq2,
            const uint32_t* __restrict__ gnmask, const uint64_t* __restrict__ goff, const uint64_t* __restrict__ glen,
            uint32_t n_contig, uint64_t tot, hb_qhsp* __restrict__ out) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    const HbAln a = alns[p];
    const HbLocus pr = pairs[p];
    hb_qhsp r;
    r.query = pr.cand; r.contig = -1; r.qs = r.qe = r.ss = r.se = 0; r.n_match = r.aln_len = 0;
    r.score = a.score; r.rank = (int32_t)pr.rank;
    if (a.score > 0) {
        const int rel = (int)pr.rel;
        const uint32_t* qr = qrec + ((size_t)qrec_unit_off[p] << 5);
        const uint64_t cbase = cgoff[pr.cand];
        int match = 0, extra = 0;
        for (int q = a.i_s + lane; q <= a.i_e; q += 32) {
            const uint32_t v = qr[q];
            extra += (int)((v >> 16) & 255u);
            if ((v >> 24) & 1u) continue;
            const int qb = packed_base(cseq2, cnmask, cbase + (uint64_t)q);
            const int tb = target_base(gseq2, gnmask, tot, rel, pr.j0 + q + (int64_t)(v & 0xFFFFu) - 32768);
            match += (qb < 4 && qb == tb) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) { match += __shfl_xor_sync(HB_FULL, match, o); extra += __shfl_xor_sync(HB_FULL, extra, o); }
        const int64_t fs = rel ? (int64_t)tot - 1 - a.j_e : a.j_s;
        const int64_t fe = rel ? (int64_t)tot - 1 - a.j_s : a.j_e;
        int lo = 0, hi = (int)n_contig;
        while (lo < hi) { int mid = (lo + hi) >> 1; if ((int64_t)goff[mid] <= fs) lo = mid + 1; else hi = mid; }
        const int c = lo - 1;
        if (c >= 0 && (uint64_t)fe < goff[c] + glen[c]) {
            const int32_t s1 = (int32_t)(fs - (int64_t)goff[c] + 1), e1 = (int32_t)(fe - (int64_t)goff[c] + 1);
            r.contig = c;
            r.qs = a.i_s + 1; r.qe = a.i_e + 1;
            r.ss = rel ? e1 : s1; r.se = rel ? s1 : e1;
            r.n_match = match; r.aln_len = a.i_e - a.i_s + 1 + extra;
        }
    }
    if (lane == 0) out[p] = r;
}

// Bytewise order of the decimal strings of two positive numbers, the shorter string being smaller when it is a
// prefix of the longer one (in a member name the number is followed by '-' or '(', both below every digit).
__device__ __forceinline__ int dec_lex_cmp(int64_t a, int64_t b) {
    if (a == b) return 0;
    int da = 1, db = 1;
    for (int64_t t = a; t >= 10; t /= 10) ++da;
    for (int64_t t = b; t >= 10; t /= 10) ++db;
    int64_t pa = a, pb = b;                               // leading min(da, db) digits of each
    for (int t = da; t > db; --t) pa /= 10;
    for (int t = db; t > da; --t) pb /= 10;
    if (pa != pb) return pa < pb ? -1 : 1;
    return da < db ? -1 : 1;
}
// is copy y's member name `contig:start1-end1(strand)` bytewise larger than copy x's ?  (tools/ready_for_MSA.sh:63)
__device__ __forceinline__ bool name_greater(const int32_t* __restrict__ crank, const HbAln& y, uint32_t rel_y, const HbAln& x,
                                             uint32_t rel_x) {
    const int ry = crank[y.contig], rx = crank[x.contig];
    if (ry != rx) return ry > rx;
    int c = dec_lex_cmp(y.start1, x.start1);
    if (c) return c > 0;
    c = dec_lex_cmp(y.end1, x.end1);
    if (c) return c > 0;
    return rel_y > rel_x;                                 // '-' (0x2D) > '+' (0x2B)
}

// S8b: chains of local alignments (definition and rationale: oracle/hbo_align.c::chain_pieces, which this follows bit for
// bit).  A copy with an indel longer than the band, or a drifting diagonal, comes back from S7 as several colinear partial
// alignments; minimap2 reports ONE alignment with a long I / D operation there (module/Util.py:7952-7961) and the copy rules
// of :8007-8022 judge the whole span.  Warp per candidate: the leftmost unused partial piece starts a chain and repeatedly
// takes the best compatible piece to its right; the chain is merged into the edit record of its best-ranked piece.
constexpr int CH_MINADD = 20, CH_MAXGAP = 1000, CH_MAXTRIM = 64, CH_MAXN = 512, CH_MAXPIECES = 32;
__global__ void __launch_bounds__(128)
chain_kernel(const HbLocus* __restrict__ pairs, const uint32_t* __restrict__ pair_off, uint32_t n_cand,
             const uint32_t* __restrict__ clen, const uint32_t* __restrict__ qrec_unit_off, uint32_t* __restrict__ qrec,
             HbAln* __restrict__ alns) {
    __shared__ uint8_t s_used[4][CH_MAXN];
    __shared__ int s_piece[4][CH_MAXPIECES], s_cut[4][CH_MAXPIECES];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t c = blockIdx.x * 4 + wid;
    if (c >= n_cand) return;
    const uint32_t p0 = pair_off[c];
    const int n = min((int)(pair_off[c + 1] - p0), CH_MAXN);
    const int L = (int)clen[c];
    uint8_t* used = s_used[wid];
    int* piece = s_piece[wid];
    int* cuts = s_cut[wid];
    int n_partial = 0;
    for (int x = lane; x < n; x += 32) {
        const HbAln a = alns[p0 + x];
        const bool u = a.score <= 0 || (double)(a.i_e - a.i_s + 1) / (double)L >= 0.95;        // not a partial piece
        used[x] = u ? 1 : 0;
        n_partial += u ? 0 : 1;
    }
    if (__ballot_sync(HB_FULL, n_partial > 0) == 0u) return;
    __syncwarp();
    auto qr_of = [&](int x) -> uint32_t* { return qrec + ((size_t)qrec_unit_off[p0 + x] << 5); };
    auto tpos = [&](int x, int qpos) -> int64_t {
        return pairs[p0 + x].j0 + qpos + (int64_t)(qr_of(x)[qpos] & 0xFFFFu) - 32768;
    };
    for (;;) {
        // the leftmost unused partial piece, best rank on ties
        uint32_t key = 0xFFFFFFFFu;
        for (int x = lane; x < n; x += 32)
            if (!used[x]) key = min(key, ((uint32_t)alns[p0 + x].i_s << 16) | (uint32_t)x);
#pragma unroll
        for (int o = 16; o; o >>= 1) key = min(key, __shfl_xor_sync(HB_FULL, key, o));
        if (key == 0xFFFFFFFFu) break;
        const int lead = (int)(key & 0xFFFFu);
        __syncwarp();
        if (lane == 0) { used[lead] = 1; piece[0] = lead; }
        __syncwarp();
        int np = 1;
        while (np < CH_MAXPIECES) {
            const int A = piece[np - 1];
            const HbAln aA = alns[p0 + A];
            const uint32_t relA = pairs[p0 + A].rel;
            unsigned long long best = ~0ull;                  // (|gq - gt| << 12 | y) << 16 | cut
            for (int y = lane; y < n; y += 32) {
                if (used[y] || pairs[p0 + y].rel != relA) continue;
                const HbAln aY = alns[p0 + y];
                if (!(aY.i_s > aA.i_s && aY.i_e >= aA.i_e + CH_MINADD && aY.j_e > aA.j_e)) continue;
                int cut = aA.i_e < aY.i_s - 1 ? aA.i_e : aY.i_s - 1;
                for (int t = 0; t < CH_MAXTRIM && cut >= aA.i_s && tpos(A, cut) >= aY.j_s; ++t) --cut;
                if (cut < aA.i_s) continue;
                const int64_t tA = tpos(A, cut);
                if (!(aY.j_s > tA)) continue;
                const int64_t gq = aY.i_s - 1 - cut, gt = aY.j_s - tA - 1;
                if (gq > CH_MAXGAP || gt > CH_MAXGAP) continue;
                const unsigned long long k2 = ((((unsigned long long)(gq > gt ? gq - gt : gt - gq) << 12) | (unsigned long long)y) << 16) |
                                              (unsigned long long)cut;
                best = k2 < best ? k2 : best;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(HB_FULL, best, o); best = t < best ? t : best; }
            if (best == ~0ull) break;
            const int by = (int)((best >> 16) & 0xFFFu), bcut = (int)(best & 0xFFFFu);
            __syncwarp();
            if (lane == 0) { cuts[np - 1] = bcut; used[by] = 1; piece[np] = by; }
            __syncwarp();
            ++np;
        }
        if (np == 1) continue;
        int h = piece[0];
        for (int t = 1; t < np; ++t) h = min(h, piece[t]);                 // best rank = smallest locus index
        const int64_t j0h = pairs[p0 + h].j0;
        uint32_t* qh = qr_of(h);
        int score = 0;
        for (int t = 0; t < np; ++t) {
            const int x = piece[t];
            const HbAln ax = alns[p0 + x];
            const int from = ax.i_s, to = (t + 1 < np) ? cuts[t] : ax.i_e;
            score += ax.score;
            int64_t tA = 0;
            if (t + 1 < np) tA = tpos(x, to);                              // read before the junction is rewritten
            __syncwarp();
            if (x != h) {
                const uint32_t* qx = qr_of(x);
                const uint32_t delta = (uint32_t)(int32_t)(pairs[p0 + x].j0 - j0h);
                for (int p = from + lane; p <= to; p += 32) {
                    const uint32_t v = qx[p];
                    qh[p] = (v & 0xFFFF0000u) | ((v + delta) & 0xFFFFu);
                }
            }
            __syncwarp();
            if (t + 1 < np) {
                const HbAln ay = alns[p0 + piece[t + 1]];
                const int64_t gt = ay.j_s - tA - 1;
                if (lane == 0) qh[to] = (qh[to] & ~(255u << 16)) | ((uint32_t)(gt > 255 ? 255 : gt) << 16);
                for (int p = to + 1 + lane; p < ay.i_s; p += 32)
                    qh[p] = ((uint32_t)(tA - j0h - p + 32768) & 0xFFFFu) | (1u << 24) | (1u << 25);
            }
            __syncwarp();
        }
        const HbAln first = alns[p0 + piece[0]], last = alns[p0 + piece[np - 1]];
        __syncwarp();
        if (lane == 0) {
            for (int t = 0; t < np; ++t) {
                const int x = piece[t];
                if (x == h) continue;
                HbAln z = alns[p0 + x];
                z.score = 0; z.i_s = -1; z.i_e = -1; z.j_s = 0; z.j_e = 0;
                alns[p0 + x] = z;
            }
            HbAln m = alns[p0 + h];
            m.score = score; m.i_s = first.i_s; m.i_e = last.i_e; m.j_s = first.j_s; m.j_e = last.j_e;
            alns[p0 + h] = m;
        }
        __syncwarp();
    }
}

// S8 dedup + S9 "max_rows longest": CTA per candidate, the compared fields staged in shared memory (both steps compare every
// copy with every other one: as a warp per candidate reading the records from global memory, a 300-locus candidate took
// milliseconds and set the kernel's duration on high-copy genomes)
constexpr int SEL_MAX = 512, SEL_THREADS = 128;
__global__ void __launch_bounds__(SEL_THREADS)
copy_select_kernel(const HbLocus* __restrict__ pairs, const uint32_t* __restrict__ pair_off, uint32_t n_cand,
                   int max_rows, int trunc_gate, int rows_first, const int32_t* __restrict__ crank, HbAln* __restrict__ alns,
                   uint32_t* __restrict__ n_rows, uint32_t* __restrict__ n_trows) {
    __shared__ long long s_js[SEL_MAX], s_je[SEL_MAX];
    __shared__ int s_len[SEL_MAX];
    __shared__ uint8_t s_rel[SEL_MAX], s_raw[SEL_MAX], s_keep[SEL_MAX];
    __shared__ int s_nk;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t c = blockIdx.x;
    const uint32_t p0 = pair_off[c], p1 = pair_off[c + 1];
    const int n = min((int)(p1 - p0), SEL_MAX);
    uint8_t* keep = s_keep;
    for (int x = tid; x < n; x += SEL_THREADS) {
        const HbAln a = alns[p0 + x];
        s_js[x] = a.j_s; s_je[x] = a.j_e; s_len[x] = a.ext_len; s_raw[x] = a.keep ? 1 : 0; s_rel[x] = (uint8_t)pairs[p0 + x].rel;
    }
    if (tid == 0) s_nk = 0;
    __syncthreads();
    int mine = 0;
    for (int x = tid; x < n; x += SEL_THREADS) {
        int k = s_raw[x];
        if (k) {
            const uint8_t rel = s_rel[x];
            const long long js = s_js[x], je = s_je[x];
            for (int y = 0; y < x; ++y)
                if (s_raw[y] && s_rel[y] == rel && s_js[y] == js && s_je[y] == je) { k = 0; break; }
        }
        keep[x] = (uint8_t)k;
        mine += k;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(HB_FULL, mine, o);
    if (lane == 0 && mine) atomicAdd(&s_nk, mine);
    __syncthreads();
    const int nk = s_nk;
    if (nk > max_rows && rows_first) {                    // FiLTR's extract_copies: the first max_rows members of the file
        for (int x = tid; x < n; x += SEL_THREADS) {
            if (!keep[x]) continue;
            int before = 0;
            for (int y = 0; y < x; ++y) before += keep[y] ? 1 : 0;
            if (before >= max_rows) keep[x] = 2;
        }
    } else if (nk > max_rows) {
        for (int x = tid; x < n; x += SEL_THREADS) {
            if (!keep[x]) continue;
            // the 100 longest, length ties to the bytewise larger member name (sort -nk 2 -r, tools/ready_for_MSA.sh:63)
            const int lx = s_len[x];
            int longer = 0;
            for (int y = 0; y < n; ++y) {
                if (y == x || !keep[y]) continue;
                const int ly = s_len[y];
                if (ly > lx) ++longer;
                else if (ly == lx && name_greater(crank, alns[p0 + y], s_rel[y], alns[p0 + x], s_rel[x])) ++longer;
            }
            if (longer >= max_rows) keep[x] = 2;
        }
    }
    __syncthreads();
    for (int x = tid; x < n; x += SEL_THREADS) if (keep[x] == 2) keep[x] = 0;
    __syncthreads();
    // rows of the full MSA, and of the truncated-member pass (Util.py:8116-8124, :10442-10446): the kept
    // copies whose extended length exceeds 1000 bp, reduced to their first 500 + last 500 bases
    if (tid < 32) {
        int rows = 0, trows = 0;
        for (int x0 = 0; x0 < n; x0 += 32) {
            const int x = x0 + lane;
            const bool k = x < n && keep[x];
            const bool lg = trunc_gate && k && s_len[x < n ? x : 0] > 1000;
            const uint32_t bal = __ballot_sync(HB_FULL, k), balt = __ballot_sync(HB_FULL, lg);
            if (x < n) {
                alns[p0 + x].keep = k ? 1 : 0;
                alns[p0 + x].row = k ? rows + __popc(bal & ((1u << lane) - 1u)) : -1;
                alns[p0 + x].trow = lg ? trows + __popc(balt & ((1u << lane) - 1u)) : -1;
            }
            rows += __popc(bal);
            trows += __popc(balt);
        }
        if (lane == 0) { n_trows[c] = (uint32_t)trows; n_rows[c] = (uint32_t)rows; }
    }
    for (uint32_t x = p0 + n + tid; x < p1; x += SEL_THREADS) { alns[x].keep = 0; alns[x].row = -1; alns[x].trow = -1; }
}

// S10a: insertion-column widths, warp per kept pair
__global__ void __launch_bounds__(256)
ins_max_kernel(const HbLocus* __restrict__ pairs, uint32_t n_pairs, const HbAln* __restrict__ alns,
               const uint32_t* __restrict__ qrec_unit_off, const uint32_t* __restrict__ qrec,
               const uint64_t* __restrict__ coff, uint32_t* __restrict__ ins) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    const HbAln a = alns[p];
    if (!a.keep) return;
    const uint32_t* qr = qrec + ((size_t)qrec_unit_off[p] << 5);
    uint32_t* dst = ins + coff[pairs[p].cand];
    // 16 B per lane (512 B per warp load): the record stream is read once, at HBM speed
#pragma unroll 2
    for (int q4 = (a.i_s & ~3) + 4 * lane; q4 <= a.i_e; q4 += 128) {
        const uint4 v = *reinterpret_cast<const uint4*>(qr + q4);
        const uint32_t d0 = (v.x >> 16) & 255u, d1 = (v.y >> 16) & 255u, d2 = (v.z >> 16) & 255u, d3 = (v.w >> 16) & 255u;
        if ((d0 | d1 | d2 | d3) == 0) continue;
        if (d0 && q4 >= a.i_s) atomicMax(&dst[q4], d0);
        if (d1 && q4 + 1 >= a.i_s && q4 + 1 <= a.i_e) atomicMax(&dst[q4 + 1], d1);
        if (d2 && q4 + 2 >= a.i_s && q4 + 2 <= a.i_e) atomicMax(&dst[q4 + 2], d2);
        if (d3 && q4 + 3 >= a.i_s && q4 + 3 <= a.i_e) atomicMax(&dst[q4 + 3], d3);
    }
}

// S10b: per candidate exclusive prefix of the widths (in place) and total columns. Warp per candidate.
__global__ void __launch_bounds__(128)
layout_kernel(const uint64_t* __restrict__ coff, const uint32_t* __restrict__ clen, uint32_t n_cand, int flank,
              uint32_t* __restrict__ ins, uint32_t* __restrict__ n_cols) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n_cand) return;
    uint32_t* a = ins + coff[c];
    const int L = (int)clen[c];
    uint32_t carry = 0;
    for (int q0 = 0; q0 < L; q0 += 32) {
        const int q = q0 + lane;
        const uint32_t v = q < L ? a[q] : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(HB_FULL, x, o); if (lane >= o) x += t; }
        if (q < L) a[q] = carry + x - v;
        carry += __shfl_sync(HB_FULL, x, 31);
    }
    if (lane == 0) n_cols[c] = (uint32_t)(L + 2 * flank) + carry;
}

// S10c: column map of every candidate's MSA: entry of column x = ((q + flank) << 8) | slot, where q in
// [-flank, L + flank) is the candidate position the column belongs to, slot 0 its base column and
// slot t >= 1 the t-th insertion column after it.  Warp per candidate; shared by all rows of the candidate.
__global__ void __launch_bounds__(128)
colmap_kernel(const uint64_t* __restrict__ coff, const uint32_t* __restrict__ clen, uint32_t n_cand, int flank,
              const uint32_t* __restrict__ pre, const uint32_t* __restrict__ n_cols,
              const uint32_t* __restrict__ col_off, uint32_t* __restrict__ colq) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n_cand) return;
    const int L = (int)clen[c];
    const uint32_t* pc = pre + coff[c];
    const uint32_t total_ins = n_cols[c] - (uint32_t)(L + 2 * flank);
    uint32_t* dst = colq + col_off[c];
    for (int q = -flank + lane; q < L + flank; q += 32) {
        const uint32_t p0 = q < 0 ? 0u : (q < L ? pc[q] : total_ins);
        const uint32_t p1 = (q < 0 || q >= L) ? p0 : (q + 1 < L ? pc[q + 1] : total_ins);
        const uint32_t base = (uint32_t)(q + flank) + p0;
        for (uint32_t t = 0; t <= p1 - p0; ++t) dst[base + t] = ((uint32_t)(q + flank) << 8) | t;
    }
}

// S10d: one MSA row per kept pair, column-centric: a warp walks the row's columns in address-aligned
// groups of 4 per lane, derives every symbol (base, insertion base, flank base or gap) from the column
// map and the pair's edit records, and writes whole 32-bit words -- gaps included, so the matrix needs
// no pre-fill and each byte of it is written exactly once (HBM: R x C bytes out).
__global__ void __launch_bounds__(256)
msa_write_kernel(const HbLocus* __restrict__ pairs, uint32_t n_pairs, const HbAln* __restrict__ alns,
                 const uint32_t* __restrict__ qrec_unit_off, const uint32_t* __restrict__ qrec,
                 const uint32_t* __restrict__ n_cols, const uint32_t* __restrict__ col_off,
                 const uint32_t* __restrict__ colq, const uint64_t* __restrict__ mat_off,
                 const uint32_t* __restrict__ gseq2, const uint32_t* __restrict__ gnmask, uint64_t tot, int flank,
                 int trunc, uint8_t* __restrict__ mat) {
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    const HbAln a = alns[p];
    const int my_row = trunc ? a.trow : a.row;
    if (my_row < 0) return;
    const HbLocus pr = pairs[p];
    const uint32_t c = pr.cand;
    const int rel = (int)pr.rel;
    const int C = (int)n_cols[c];
    const uint32_t* cq = colq + col_off[c];
    uint8_t* row = mat + mat_off[c] + (size_t)my_row * C;
    const uint32_t* qr = qrec + ((size_t)qrec_unit_off[p] << 5);
    // Every position a kept row reads lies inside its contig (copy_filter_kernel dropped rows whose
    // +-flank extension leaves it), i.e. inside [0, tot) with tot < 2^32: positions are plain uint32,
    // no bounds checks.  js / je: oriented position of the first / last aligned base.
    const uint32_t j0 = (uint32_t)pr.j0, js = (uint32_t)a.j_s, je = (uint32_t)a.j_e;
    const uint32_t totm1 = (uint32_t)(tot - 1);
    const int i_s = a.i_s, i_e = a.i_e;
    // truncated pass keeps only the first 500 and last 500 bases of the extended copy
    const uint32_t ext_lo = js - (uint32_t)flank, ext_len = je - js + 2u * (uint32_t)flank + 1u;
    auto sym_at = [&](int col) -> uint32_t {
        const uint32_t e = cq[col];
        const int q = (int)(e >> 8) - flank;
        const uint32_t slot = e & 255u;
        uint32_t j;
        if (q >= i_s && q <= i_e) {
            const uint32_t v = qr[q];
            if (slot == 0 ? ((v >> 24) & 1u) != 0 : slot > ((v >> 16) & 255u)) return HB_SYM_GAP;
            j = j0 + (uint32_t)q + (v & 0xFFFFu) - 32768u + slot;
        } else {
            if (slot != 0) return HB_SYM_GAP;
            if (q < i_s) { if (q < i_s - flank) return HB_SYM_GAP; j = js - (uint32_t)(i_s - q); }
            else { if (q > i_e + flank) return HB_SYM_GAP; j = je + (uint32_t)(q - i_e); }
        }
        if (trunc) {
            const uint32_t d = j - ext_lo;                 // offset inside the extended copy
            if (d >= 500u && ext_len - 1u - d >= 500u) return HB_SYM_GAP;
        }
        const uint32_t pos = rel ? totm1 - j : j;
        if ((gnmask[pos >> 5] >> (pos & 31u)) & 1u) return HB_SYM_N;
        const uint32_t b2 = (gseq2[pos >> 4] >> (2u * (pos & 15u))) & 3u;
        return rel ? 3u - b2 : b2;
    };
    const int a0 = (int)(reinterpret_cast<uintptr_t>(row) & 3u);        // virtual column x = col + a0 is word aligned
    for (int x = 4 * lane; x < C + a0; x += 128) {
        const int col = x - a0;
        if (col >= 0 && col + 3 < C) {
            const uint32_t w = sym_at(col) | (sym_at(col + 1) << 8) | (sym_at(col + 2) << 16) | (sym_at(col + 3) << 24);
            *reinterpret_cast<uint32_t*>(row + col) = w;
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (col + t >= 0 && col + t < C) row[col + t] = (uint8_t)sym_at(col + t);
        }
    }
}

// S10d + sparse-column removal + bit-planes fused (msa_write_kernel, hb_vote.cu::hb_clean_kernel and build_planes without the
// raw matrix ever reaching HBM): one CTA per candidate, producing the clean matrix of its full judge task AND of its
// truncated-member task (rows = its copies longer than 1000 bp, cells outside each row's first / last 500 bases blanked)
// from ONE computation of the symbols.  The matrices are built 128 columns at a time in shared memory, COLUMN-major (a
// column's rows are consecutive bytes, 27 words per column so that a warp's columns fall into distinct banks): a warp per
// row derives the row's symbols from the staged column map and the pair's edit records exactly like msa_write_kernel and
// drops them into the full tile at its row and into the truncated tile at its truncated row.  Then a thread per column and
// matrix works on four rows per 32-bit word: gaps are counted by a zero-byte test, the column is kept or dropped by the
// reference's rule (remove_sparse_col_in_align_file, module/Util.py:10379-10391: first and last column always kept, a column
// goes when gaps > rows / 2), kept columns are appended to the clean matrix (row stride C) and their six bit-planes (what the
// judges vote on) are derived from three code-bit planes gathered with one multiply per word.
constexpr int MC_TILE = 128, MC_THREADS = 256, MC_MAXROWS = 104, MC_PITCH = 108;

// per-row constants of a candidate's kept copies, staged once per CTA (struct of arrays in shared memory)
struct McRows {
    uint32_t qoff[MC_MAXROWS], j0[MC_MAXROWS], js[MC_MAXROWS], je[MC_MAXROWS];
    int i_s[MC_MAXROWS], i_e[MC_MAXROWS];
    short trow[MC_MAXROWS];
    uint8_t rel[MC_MAXROWS];
};

// four cells of one row (columns col0 + 32 k), branch-free so that the 4 edit-record loads and then the 8 genome loads are in
// flight together; byte k = symbol, bit 7 = blank in the truncated-member matrix
__device__ __forceinline__ uint32_t mc_row_word(const McRows& rw, int r, const uint32_t (&e)[4], int col0, int C,
                                                const uint32_t* __restrict__ qrec, const uint32_t* __restrict__ gseq2,
                                                const uint32_t* __restrict__ gnmask, uint32_t totm1, int flank) {
    const uint32_t* qr = qrec + ((size_t)rw.qoff[r] << 5);
    const uint32_t j0 = rw.j0[r], js = rw.js[r], je = rw.je[r];
    const int i_s = rw.i_s[r], i_e = rw.i_e[r];
    const bool rel = rw.rel[r] != 0, in_trunc = rw.trow[r] >= 0;
    const uint32_t ext_lo = js - (uint32_t)flank, ext_len = je - js + 2u * (uint32_t)flank + 1u;
    uint32_t v[4];
    int q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q[k] = (int)(e[k] >> 8) - flank;
        v[k] = qr[min(max(q[k], i_s), i_e)];
    }
    uint32_t j[4], nm[4], sq[4];
    bool have[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t slot = e[k] & 255u;
        const bool inside = q[k] >= i_s && q[k] <= i_e, left = q[k] < i_s;
        const bool have_in = slot == 0 ? ((v[k] >> 24) & 1u) == 0 : slot <= ((v[k] >> 16) & 255u);
        const bool have_out = slot == 0 && (left ? q[k] >= i_s - flank : q[k] <= i_e + flank);
        const uint32_t j_in = j0 + (uint32_t)q[k] + (v[k] & 0xFFFFu) - 32768u + slot;
        const uint32_t j_out = left ? js - (uint32_t)(i_s - q[k]) : je + (uint32_t)(q[k] - i_e);
        have[k] = (col0 + 32 * k < C) && (inside ? have_in : have_out);
        j[k] = inside ? j_in : j_out;
        const uint32_t pos = min(rel ? totm1 - j[k] : j[k], totm1);       // a cell without a base reads a harmless word
        nm[k] = (gnmask[pos >> 5] >> (pos & 31u)) & 1u;
        sq[k] = (gseq2[pos >> 4] >> (2u * (pos & 15u))) & 3u;
    }
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t sym = nm[k] ? (uint32_t)HB_SYM_N : (rel ? 3u - sq[k] : sq[k]);
        const uint32_t d = j[k] - ext_lo;                                   // blank in the truncated-member matrix ?
        if (in_trunc && d >= 500u && ext_len - 1u - d >= 500u) sym |= 0x80u;
        w |= (have[k] ? sym : (uint32_t)HB_SYM_GAP) << (8 * k);
    }
    return w;
}

__global__ void __launch_bounds__(MC_THREADS, 4)
msa_clean_kernel(const HbVoteTask* __restrict__ tasks, const uint32_t* __restrict__ order, const uint32_t* __restrict__ task_cand,
                 const int32_t* __restrict__ trunc_of_full,
                 const HbLocus* __restrict__ pairs, const uint32_t* __restrict__ pair_off, const HbAln* __restrict__ alns,
                 const uint32_t* __restrict__ qrec_unit_off, const uint32_t* __restrict__ qrec,
                 const uint32_t* __restrict__ col_off, const uint32_t* __restrict__ colq, const uint32_t* __restrict__ gseq2,
                 const uint32_t* __restrict__ gnmask, uint64_t tot, int flank, uint8_t* __restrict__ clean,
                 uint32_t* __restrict__ planes, HbVoteCall* __restrict__ calls) {
    __shared__ __align__(16) uint8_t s_tile[2][MC_TILE][MC_PITCH];     // [full | truncated][column][row]
    __shared__ uint32_t s_cq[MC_TILE];
    __shared__ McRows rw;
    __shared__ int s_warp[2][MC_TILE / 32];
    static_assert(HB_SYM_GAP == 5 && MC_PITCH % 4 == 0 && (MC_PITCH / 4) % 2 == 1 && MC_PITCH >= MC_MAXROWS, "tile layout");
    const uint32_t ti = order[blockIdx.x];                // full task; the largest matrices are started first
    const int tt = trunc_of_full[ti];                     // its truncated-member task, or -1
    const HbVoteTask t = tasks[ti];
    const uint32_t c = task_cand[ti];
    const int R = t.R, C = t.C;
    const int RT = tt >= 0 ? tasks[tt].R : 0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t p0 = pair_off[c], p1 = pair_off[c + 1];
    for (uint32_t p = p0 + tid; p < p1; p += MC_THREADS) {
        const HbAln a = alns[p];
        if (a.row >= 0 && a.row < MC_MAXROWS) {
            const HbLocus pr = pairs[p];
            const int r = a.row;
            rw.qoff[r] = qrec_unit_off[p]; rw.j0[r] = (uint32_t)pr.j0; rw.js[r] = (uint32_t)a.j_s; rw.je[r] = (uint32_t)a.j_e;
            rw.i_s[r] = a.i_s; rw.i_e[r] = a.i_e; rw.trow[r] = (short)a.trow; rw.rel[r] = (uint8_t)pr.rel;
        }
    }
    // rows R.. (RT..) of a tile column are never written: zero = "not a gap", and they stay outside the planes' valid bits
    for (int i = tid; i < 2 * MC_TILE * MC_PITCH / 4; i += MC_THREADS) reinterpret_cast<uint32_t*>(&s_tile[0][0][0])[i] = 0u;
    __syncthreads();
    const uint32_t* cq = colq + col_off[c];
    const uint32_t totm1 = (uint32_t)(tot - 1);
    // column phase: threads 0..127 own the full matrix, 128..255 the truncated-member matrix
    const int half = tid >> 7, ct = tid & (MC_TILE - 1), hw = wid & 3;
    uint8_t* out = half == 0 ? clean + t.clean_off : (tt >= 0 ? clean + tasks[tt].clean_off : nullptr);
    const int RH = half == 0 ? R : RT;
    const int G = (RH + 3) >> 2;                          // 4-row words of my column
    const uint32_t* colw = reinterpret_cast<const uint32_t*>(&s_tile[half][ct][0]);
    // six bit-planes of every kept column over ALL rows of my matrix (row k = bit k & 31 of word k >> 5): what the judges
    // vote on (hb_vote.cu::Planes; they mask the rows they keep)
    uint32_t* pw = planes == nullptr ? nullptr : planes + (half == 0 ? t.plane_off : (tt >= 0 ? tasks[tt].plane_off : 0));
    int base = 0;                                         // kept columns of my matrix so far
    for (int c0 = 0; c0 < C; c0 += MC_TILE) {
        if (tid < MC_TILE) s_cq[tid] = c0 + tid < C ? cq[c0 + tid] : 0u;
        __syncthreads();
        // ---- the tile's symbols: warp per row (two rows in flight), columns lane + 32 k ----
        uint32_t e4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) e4[k] = s_cq[lane + 32 * k];
        for (int r = wid; r < R; r += 2 * (MC_THREADS / 32)) {
            const int r2 = r + MC_THREADS / 32;
            const bool two = r2 < R;                                     // (warp-uniform)
            const uint32_t wa = mc_row_word(rw, r, e4, c0 + lane, C, qrec, gseq2, gnmask, totm1, flank);
            const int ta = rw.trow[r];
            if (two) {
                const uint32_t wb = mc_row_word(rw, r2, e4, c0 + lane, C, qrec, gseq2, gnmask, totm1, flank);
                const int tb = rw.trow[r2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t bb = (wb >> (8 * k)) & 255u;
                    s_tile[0][lane + 32 * k][r2] = (uint8_t)(bb & 0x7Fu);
                    if (tb >= 0) s_tile[1][lane + 32 * k][tb] = (bb & 0x80u) ? (uint8_t)HB_SYM_GAP : (uint8_t)bb;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t ba = (wa >> (8 * k)) & 255u;
                s_tile[0][lane + 32 * k][r] = (uint8_t)(ba & 0x7Fu);
                if (ta >= 0) s_tile[1][lane + 32 * k][ta] = (ba & 0x80u) ? (uint8_t)HB_SYM_GAP : (uint8_t)ba;
            }
        }
        __syncthreads();
        // ---- sparse columns (remove_sparse_col_in_align_file): thread per column and matrix, four rows per word ----
        const int col = c0 + ct;
        int keep = 0;
        if (RH > 0 && col < C) {
            int gaps = 0;
            for (int g = 0; g < G; ++g) {
                const uint32_t x = colw[g] ^ 0x05050505u;                 // symbols are < 0x80: exact zero-byte test
                gaps += __popc(~((x + 0x7F7F7F7Fu) | x) & 0x80808080u);
            }
            keep = (col == 0) || (col == C - 1) || !(2 * gaps > RH);
        }
        const uint32_t bal = __ballot_sync(HB_FULL, keep);
        if (lane == 0) s_warp[half][hw] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w2 = 0; w2 < MC_TILE / 32; ++w2) { const int n = s_warp[half][w2]; before += (w2 < hw) ? n : 0; total += n; }
        if (keep) {
            const int at = base + before + __popc(bal & ((1u << lane) - 1u));
            uint8_t* op = out + at;
            uint32_t b0 = 0, b1 = 0, b2 = 0;              // code-bit planes of the current 32 rows
            for (int g = 0; g < G; ++g) {
                const uint32_t w = colw[g];
                const int r = 4 * g;
                if (r + 3 < RH) {
                    op[0] = (uint8_t)w; op[(size_t)C] = (uint8_t)(w >> 8); op[2 * (size_t)C] = (uint8_t)(w >> 16);
                    op[3 * (size_t)C] = (uint8_t)(w >> 24);
                } else {
                    op[0] = (uint8_t)w;
                    if (r + 1 < RH) op[(size_t)C] = (uint8_t)(w >> 8);
                    if (r + 2 < RH) op[2 * (size_t)C] = (uint8_t)(w >> 16);
                }
                op += 4 * (size_t)C;
                if (pw != nullptr) {
                    // bits 0 / 8 / 16 / 24 -> one nibble: x * 0x01020408 puts them at bits 24..27 without carries
                    const int sh = 4 * (g & 7);
                    b0 |= (((w & 0x01010101u) * 0x01020408u) >> 24) << sh;
                    b1 |= ((((w >> 1) & 0x01010101u) * 0x01020408u) >> 24) << sh;
                    b2 |= ((((w >> 2) & 0x01010101u) * 0x01020408u) >> 24) << sh;
                    if ((g & 7) == 7 || g == G - 1) {
                        const int wd = g >> 3, left = RH - 32 * wd;
                        const uint32_t valid = left >= 32 ? 0xFFFFFFFFu : (1u << left) - 1u;
                        const size_t o = (size_t)wd * C + at, pl = 4 * (size_t)C;
                        pw[o] = ~b2 & ~b1 & ~b0 & valid; pw[o + pl] = ~b2 & ~b1 & b0; pw[o + 2 * pl] = ~b2 & b1 & ~b0;
                        pw[o + 3 * pl] = ~b2 & b1 & b0; pw[o + 4 * pl] = b2 & ~b1 & ~b0; pw[o + 5 * pl] = b2 & ~b1 & b0;
                        b0 = b1 = b2 = 0;
                    }
                }
            }
        }
        base += total;
        // no barrier here: s_cq is only read before the tile's second barrier, the tiles are rewritten after the next tile's
        // first barrier and s_warp after its second
    }
    if (ct == 0) { if (half == 0) calls[ti].clean_cols = base; else if (tt >= 0) calls[tt].clean_cols = base; }
}

}  // namespace

cudaError_t hb_copy_filter_launch(const HbLocus* pairs, uint32_t n_pairs, const uint32_t* clen, const uint64_t* goff,
                                  const uint64_t* glen, uint32_t n_contig, uint64_t tot, int flank, HbAln* alns,
                                  cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    copy_filter_kernel<<<(n_pairs + 255) / 256, 256, 0, s>>>(pairs, n_pairs, clen, goff, glen, n_contig, tot, flank, alns);
    return cudaGetLastError();
}
cudaError_t hb_chain_launch(const HbLocus* pairs, const uint32_t* pair_off, uint32_t n_cand, const uint32_t* clen,
                            const uint32_t* qrec_unit_off, uint32_t* qrec, HbAln* alns, cudaStream_t s) {
    if (n_cand == 0) return cudaSuccess;
    chain_kernel<<<(n_cand + 3) / 4, 128, 0, s>>>(pairs, pair_off, n_cand, clen, qrec_unit_off, qrec, alns);
    return cudaGetLastError();
}
cudaError_t hb_copy_select_launch(const HbLocus* pairs, const uint32_t* pair_off, uint32_t n_cand, int max_rows,
                                  int trunc_gate, int rows_first, const int32_t* crank, HbAln* alns, uint32_t* n_rows,
                                  uint32_t* n_trows, cudaStream_t s) {
    if (n_cand == 0) return cudaSuccess;
    copy_select_kernel<<<n_cand, SEL_THREADS, 0, s>>>(pairs, pair_off, n_cand, max_rows, trunc_gate, rows_first, crank, alns,
                                                        n_rows, n_trows);
    return cudaGetLastError();
}
cudaError_t hb_ins_max_launch(const HbLocus* pairs, uint32_t n_pairs, const HbAln* alns, const uint32_t* qrec_unit_off,
                              const uint32_t* qrec, const uint64_t* coff, uint32_t* ins, cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_pairs * 32;
    ins_max_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(pairs, n_pairs, alns, qrec_unit_off, qrec, coff, ins);
    return cudaGetLastError();
}
cudaError_t hb_layout_launch(const uint64_t* coff, const uint32_t* clen, uint32_t n_cand, int flank, uint32_t* ins,
                             uint32_t* n_cols, cudaStream_t s) {
    if (n_cand == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_cand * 32;
    layout_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(coff, clen, n_cand, flank, ins, n_cols);
    return cudaGetLastError();
}
cudaError_t hb_colmap_launch(const uint64_t* coff, const uint32_t* clen, uint32_t n_cand, int flank, const uint32_t* pre,
                             const uint32_t* n_cols, const uint32_t* col_off, uint32_t* colq, cudaStream_t s) {
    if (n_cand == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_cand * 32;
    colmap_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(coff, clen, n_cand, flank, pre, n_cols, col_off, colq);
    return cudaGetLastError();
}
cudaError_t hb_msa_write_launch(const HbLocus* pairs, uint32_t n_pairs, const HbAln* alns, const uint32_t* qrec_unit_off,
                                const uint32_t* qrec, const uint32_t* n_cols, const uint32_t* col_off,
                                const uint32_t* colq, const uint64_t* mat_off, const uint32_t* gseq2,
                                const uint32_t* gnmask, uint64_t tot, int flank, int trunc, uint8_t* mat,
                                cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_pairs * 32;
    msa_write_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(pairs, n_pairs, alns, qrec_unit_off, qrec, n_cols,
                                                                       col_off, colq, mat_off, gseq2, gnmask, tot, flank,
                                                                       trunc, mat);
    return cudaGetLastError();
}
cudaError_t hb_qhsp_launch(const HbLocus* pairs, uint32_t n_pairs, const HbAln* alns, const uint32_t* qrec_unit_off,
                           const uint32_t* qrec, const uint32_t* cseq2, const uint32_t* cnmask, const uint64_t* cgoff,
                           const uint32_t* gseq2, const uint32_t* gnmask, const uint64_t* goff, const uint64_t* glen,
                           uint32_t n_contig, uint64_t tot, hb_qhsp* out, cudaStream_t s) {
    if (n_pairs == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_pairs * 32;
    qhsp_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(pairs, n_pairs, alns, qrec_unit_off, qrec, cseq2, cnmask, cgoff,
                                                                  gseq2, gnmask, goff, glen, n_contig, tot, out);
    return cudaGetLastError();
}
cudaError_t hb_msa_clean_launch(const HbVoteTask* tasks, const uint32_t* order, const uint32_t* task_cand, const int32_t* trunc_of_full,
                                uint32_t n_full,
                                const HbLocus* pairs, const uint32_t* pair_off, const HbAln* alns, const uint32_t* qrec_unit_off,
                                const uint32_t* qrec, const uint32_t* col_off, const uint32_t* colq, const uint32_t* gseq2,
                                const uint32_t* gnmask, uint64_t tot, int flank, uint8_t* clean, uint32_t* planes, HbVoteCall* calls,
                                cudaStream_t s) {
    if (n_full == 0) return cudaSuccess;
    msa_clean_kernel<<<n_full, MC_THREADS, 0, s>>>(tasks, order, task_cand, trunc_of_full, pairs, pair_off, alns, qrec_unit_off, qrec, col_off,
                                                   colq, gseq2, gnmask, tot, flank, clean, planes, calls);
    return cudaGetLastError();
}
