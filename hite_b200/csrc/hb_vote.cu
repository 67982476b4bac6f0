// hite_b200 -- kernel (iii): flank-copy MSA column voting + homology-window boundary search.
//
// Replaces, bit-exactly on the same aligned matrix, the reference's pure-Python
//   remove_sparse_col_in_align_file   (module/Util.py:10344-10405)   -> hb_clean_kernel
//   judge_boundary_v5  (TIR)          (module/Util.py:9145-9480)     -> hb_judge_kernel
//   search_boundary_homo_v3           (module/Util.py:8887-9143)
//   calculate_window_homology         (module/Util.py:8827-8885)
//   TSDsearch_v5                      (module/Util.py:2460-2492)
//
// Layout in HBM (per candidate): the R x C matrix row-major (1 B/cell) for the row-wise work
// (anchor search, TSD strings) and, for the <=101 full-length rows the judge keeps, six column
// bit-planes P[sym][word][col] (row r = bit r&31 of word r>>5) so that every column statistic
// is a popc and every "rows valid in this window" set is a ballot.  One warp per candidate;
// no tensor cores (integer/bit work), no CPU fallback.
#include "hb_common.cuh"

namespace {

constexpr int WARPS_PER_CTA = 4;
constexpr int CLEAN_THREADS = 256;
constexpr int MAX_KEPT = 101;   // "len(...) > 100: break"  Util.py:9201

// ------------------------------------------------------------------------------------------
// warp helpers
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(HB_FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(HB_FULL, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(HB_FULL, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(HB_FULL, v, o);
        v = t < v ? t : v;
    }
    return v;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(HB_FULL, x, o);
        if (lane >= o) x += t;
    }
    *total = __shfl_sync(HB_FULL, x, 31);
    return x - v;
}
__device__ __forceinline__ uint32_t lanemask_lt(int lane) { return (1u << lane) - 1u; }

// Python slice bounds s[a:b] for a sequence of length n
__device__ __forceinline__ void py_slice(int n, long long a, long long b, int* lo, int* hi) {
    if (a < 0) { a += n; if (a < 0) a = 0; } else if (a > n) a = n;
    if (b < 0) { b += n; if (b < 0) b = 0; } else if (b > n) b = n;
    if (b < a) b = a;
    *lo = (int)a; *hi = (int)b;
}

// ------------------------------------------------------------------------------------------
// kernel A: sparse-column removal (Util.py:10344-10405).  One CTA per candidate.
// keeps column 0 and C-1 unconditionally (the elif order at :10384-10387), drops a column iff
// gaps > row_num/2 (float compare == 2*gaps > R).  Output: symbol codes, row stride C.
__global__ void __launch_bounds__(CLEAN_THREADS)
hb_clean_kernel(const HbVoteTask* __restrict__ tasks, const uint8_t* __restrict__ raw,
                uint8_t* __restrict__ clean, HbVoteCall* __restrict__ calls) {
    const HbVoteTask t = tasks[blockIdx.x];
    const int R = t.R, C = t.C;
    const uint8_t* in = raw + t.raw_off;
    uint8_t* out = clean + t.clean_off;
    __shared__ int s_warp[CLEAN_THREADS / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < C; c0 += CLEAN_THREADS) {
        const int c = c0 + tid;
        int keep = 0;
        if (c < C) {
            int gaps = 0;
            for (int r = 0; r < R; ++r) {
                uint8_t ch = in[(size_t)r * C + c];
                gaps += t.raw_is_ascii ? (ch == '-') : (ch == HB_SYM_GAP);
            }
            keep = (c == 0) || (c == C - 1) || !(2 * gaps > R);
        }
        const uint32_t bal = __ballot_sync(HB_FULL, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < CLEAN_THREADS / 32; ++w) {
            int v = s_warp[w];
            if (w < wid) before += v;
            total += v;
        }
        const int base = s_base;
        const int pos = base + before + __popc(bal & lanemask_lt(lane));
        if (keep) {
            for (int r = 0; r < R; ++r) {
                uint8_t ch = in[(size_t)r * C + c];
                out[(size_t)r * C + pos] = t.raw_is_ascii ? (uint8_t)hb_sym_of_ascii(ch) : ch;
            }
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    if (tid == 0) calls[blockIdx.x].clean_cols = s_base;
}

// ------------------------------------------------------------------------------------------
// anchor search: HB-fnm (see oracle/fnm.py; stand-in for fuzzysearch.find_near_matches with
// max_l_dist = 2, Util.py:9173-9174).  Bit-parallel Sellers/Myers over 32 lane chunks.
struct Pattern {
    uint32_t peq[5];
    unsigned long long p0;   // symbols 0..15, 4 bits each
    uint32_t p1;             // symbols 16..19
    int m;
    __device__ __forceinline__ int sym(int j) const {
        return j < 16 ? (int)((p0 >> (4 * j)) & 15ull) : (int)((p1 >> (4 * (j - 16))) & 15u);
    }
};
__device__ __forceinline__ void pattern_init(Pattern& P, const uint8_t* ascii, int m) {
    P.m = m; P.p0 = 0; P.p1 = 0;
#pragma unroll
    for (int s = 0; s < 5; ++s) P.peq[s] = 0;
    for (int j = 0; j < m; ++j) {
        int s = hb_sym_of_ascii(ascii[j]);
        if (s == HB_SYM_GAP) s = HB_SYM_N;
#pragma unroll
        for (int q = 0; q < 5; ++q) if (q == s) P.peq[q] |= 1u << j;
        if (j < 16) P.p0 |= (unsigned long long)s << (4 * j); else P.p1 |= (uint32_t)s << (4 * (j - 16));
    }
}
struct Myers {
    uint32_t pv, mv; int score;
    __device__ __forceinline__ void init(int m) { pv = m >= 32 ? 0xffffffffu : ((1u << m) - 1u); mv = 0; score = m; }
    __device__ __forceinline__ void step(uint32_t eq, uint32_t top) {
        uint32_t xv = eq | mv;
        uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
        uint32_t ph = mv | ~(xh | pv);
        uint32_t mh = pv & xh;
        if (ph & top) ++score; else if (mh & top) --score;
        ph <<= 1; mh <<= 1;
        pv = mh | ~(xv | ph);
        mv = ph & xv;
    }
};
struct Hist {  // last 32 non-gap symbols, most recent first
    unsigned long long h0, h1;
    __device__ __forceinline__ void init() { h0 = 0; h1 = 0; }
    __device__ __forceinline__ void push(int s) { h1 = (h1 << 4) | (h0 >> 60); h0 = (h0 << 4) | (unsigned long long)s; }
    __device__ __forceinline__ int back(int k) const {  // k = 1: most recent
        return k <= 16 ? (int)((h0 >> (4 * (k - 1))) & 15ull) : (int)((h1 >> (4 * (k - 17))) & 15ull);
    }
};
// smallest a with lev(p, s[a:e]) == dist  (oracle/fnm.py::_longest_start).  Bit-parallel: the
// reversed pattern is aligned globally-at-the-start against the text read backwards from e, so after
// k symbols the running score is lev(p, s[e-k:e]); the largest k that reaches `dist` wins.
__device__ __noinline__ int longest_start(const Pattern& P, const Hist& H, int e, int dist) {
    const int m = P.m;
    int maxk = m + dist; if (maxk > e) maxk = e;
    uint32_t peq_r[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) peq_r[s] = __brev(P.peq[s]) >> (32 - m);
    const uint32_t top = 1u << (m - 1);
    uint32_t pv = m >= 32 ? 0xffffffffu : ((1u << m) - 1u), mv = 0;
    int score = m, best_k = 0;          // k = 0: lev(p, '') = m, only equal to dist when m <= dist
    for (int k = 1; k <= maxk; ++k) {
        const int c = H.back(k);
        uint32_t eq = 0;
#pragma unroll
        for (int s = 0; s < 5; ++s) eq = (c == s) ? peq_r[s] : eq;
        const uint32_t xv = eq | mv;
        const uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
        uint32_t ph = mv | ~(xh | pv);
        uint32_t mh = pv & xh;
        if (ph & top) ++score; else if (mh & top) --score;
        ph = (ph << 1) | 1u; mh <<= 1;
        pv = mh | ~(xv | ph);
        mv = ph & xv;
        if (score == dist) best_k = k;
    }
    return e - best_k;
}

struct Match { int a, e, d; };

// One row: run both patterns over the ungapped row, emit (start,end,dist) of every end with
// dist <= 2 into mlist[0..cap) / mlist[cap..2cap).  Returns via cnt[0], cnt[1] (smem).
__device__ __noinline__ void anchor_scan_row(const uint8_t* __restrict__ row, int Cc, const Pattern& P1, const Pattern& P2,
                                Match* __restrict__ mlist, int cap, int* cnt, int lane,
                                int* ub_out, int* n_out) {
    const int chunk = (Cc + 31) >> 5;
    const int lo = min(Cc, lane * chunk), hi = min(Cc, lo + chunk);
    int c_ng = 0;
    for (int c = lo; c < hi; ++c) c_ng += (row[c] != HB_SYM_GAP);
    int n;
    const int ub = warp_excl_scan(c_ng, lane, &n);
    *ub_out = ub; *n_out = n;
    if (lane < 2) cnt[lane] = 0;
    __syncwarp();
    Myers M1, M2; M1.init(P1.m); M2.init(P2.m);
    const uint32_t top1 = 1u << (P1.m - 1), top2 = 1u << (P2.m - 1);
    Hist H; H.init();
    // warm-up: the previous 40 (= 2m) non-gap symbols fully determine the DP column
    uint8_t warm[40];
    int nwarm = 0;
    for (int c = lo - 1; c >= 0 && nwarm < 40; --c) { uint8_t s = row[c]; if (s != HB_SYM_GAP) warm[nwarm++] = s; }
    for (int i = nwarm - 1; i >= 0; --i) {
        int s = warm[i];
        M1.step(P1.peq[s], top1); M2.step(P2.peq[s], top2); H.push(s);
    }
    int e = ub;
    for (int c = lo; c < hi; ++c) {
        int s = row[c];
        if (s == HB_SYM_GAP) continue;
        M1.step(P1.peq[s], top1); M2.step(P2.peq[s], top2); H.push(s);
        ++e;
        if (M1.score <= 2) {
            int i = atomicAdd(&cnt[0], 1);
            if (i < cap) { Match mt; mt.e = e; mt.d = M1.score; mt.a = longest_start(P1, H, e, M1.score); mlist[i] = mt; }
        }
        if (M2.score <= 2) {
            int i = atomicAdd(&cnt[1], 1);
            if (i < cap) { Match mt; mt.e = e; mt.d = M2.score; mt.a = longest_start(P2, H, e, M2.score); mlist[cap + i] = mt; }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned long long match_key(const Match& m) {
    // min over (dist, -(len), start); end carried in the low bits
    return ((unsigned long long)m.d << 58) | ((unsigned long long)(63 - (m.e - m.a)) << 52) |
           ((unsigned long long)m.a << 26) | (unsigned long long)m.e;
}

// best match of the first overlap-group (by start) -> its start
__device__ __noinline__ int first_group_best_start(const Match* l, int n, int lane) {
    int gs = 0x7fffffff;
    for (int i = lane; i < n; i += 32) gs = min(gs, l[i].a);
    gs = warp_min(gs);
    int ge = gs + 1;
    for (;;) {
        int ne = ge;
        for (int i = lane; i < n; i += 32) if (l[i].a < ge) ne = max(ne, l[i].e);
        ne = warp_max(ne);
        if (ne == ge) break;
        ge = ne;
    }
    unsigned long long k = ~0ull;
    for (int i = lane; i < n; i += 32) if (l[i].a < ge) { unsigned long long q = match_key(l[i]); k = q < k ? q : k; }
    k = warp_min_u64(k);
    return (int)((k >> 26) & 0x3ffffffull);
}
// best match of the last overlap-group -> its end
__device__ __noinline__ int last_group_best_end(const Match* l, int n, int lane) {
    int ge = -1;
    for (int i = lane; i < n; i += 32) ge = max(ge, l[i].e);
    ge = warp_max(ge);
    int gs = ge - 1;
    for (;;) {
        int ns = gs;
        for (int i = lane; i < n; i += 32) if (l[i].e > gs) ns = min(ns, l[i].a);
        ns = warp_min(ns);
        if (ns == gs) break;
        gs = ns;
    }
    unsigned long long k = ~0ull;
    for (int i = lane; i < n; i += 32) if (l[i].e > gs) { unsigned long long q = match_key(l[i]); k = q < k ? q : k; }
    k = warp_min_u64(k);
    return (int)(k & 0x3ffffffull);
}
// column of the idx-th non-gap symbol of the row (lanes own [lo,hi) with ungapped base ub)
__device__ __noinline__ int column_of_ungapped(const uint8_t* row, int Cc, int idx, int ub, int lane) {
    const int chunk = (Cc + 31) >> 5;
    const int lo = min(Cc, lane * chunk), hi = min(Cc, lo + chunk);
    int col = -1, k = ub;
    for (int c = lo; c < hi; ++c) {
        if (row[c] == HB_SYM_GAP) continue;
        if (k == idx) { col = c; break; }
        ++k;
    }
    return warp_max(col);
}

// ------------------------------------------------------------------------------------------
// bit-plane view of the kept rows
struct Planes {
    const uint32_t* p;   // P[sym][word][col]
    int stride;          // C
    int nw;              // words in use
    int row_num;
    uint32_t rowmask[4]; // the rows that take part: planes built by build_planes hold exactly the kept rows in list order (a prefix
                         // mask); planes prebuilt over ALL matrix rows (hb_msa.cu::msa_clean_kernel) are masked down to the kept
                         // rows -- counts and first-seen order (ascending row) are the same either way
    __device__ __forceinline__ uint32_t w(int sym, int word, int col) const {
        return p[(size_t)(sym * 4 + word) * stride + col] & rowmask[word];
    }
    __device__ __forceinline__ int count(int sym, int col) const {
        int c = 0;
        for (int k = 0; k < nw; ++k) c += __popc(w(sym, k, col));
        return c;
    }
};

// collect <= 100 valid columns (gaps <= valid_col_threshold) scanning from `start` in direction
// dir while lo_bound <= col < hi_bound  (Util.py:8924-8947, 8976-8999, 9036-9060, 9096-9120)
__device__ __noinline__ int collect_valid(const Planes& P, int vthr, int start, int dir, int lo_bound, int hi_bound,
                             int* vc, int lane) {
    int n = 0;
    for (int c0 = start; n < 100; c0 += 32 * dir) {
        const int c = c0 + dir * lane;
        const bool in = (c >= lo_bound && c < hi_bound);
        const bool ok = in && (P.count(HB_SYM_GAP, c) <= vthr);
        const uint32_t bal = __ballot_sync(HB_FULL, ok);
        const int pos = n + __popc(bal & lanemask_lt(lane));
        if (ok && pos < 100) vc[pos] = c;
        n = min(100, n + __popc(bal));
        // stop once the leading lane has left the range (ranges are contiguous in scan order)
        const int cfirst = c0, clast = c0 + dir * 31;
        const bool any_in = (dir > 0) ? (cfirst < hi_bound) : (cfirst >= lo_bound);
        const bool last_in = (dir > 0) ? (clast < hi_bound) : (clast >= lo_bound);
        if (!any_in || !last_in) break;
    }
    __syncwarp();
    return n;
}
__device__ void reverse_list(int* vc, int n, int lane) {
    for (int i = lane; i < n / 2; i += 32) { int t = vc[i]; vc[i] = vc[n - 1 - i]; vc[n - 1 - i] = t; }
    __syncwarp();
}

// calculate_window_homology (Util.py:8827-8885); window columns are first + step*k, k < ncols
__device__ __noinline__ int window_homology(const Planes& P, int first, int ncols, int step, double thr, int lane) {
    uint32_t valid[4] = {0, 0, 0, 0};
    int nv = 0;
    for (int wd = 0; wd < P.nw; ++wd) {
        int cnt = 0;
        for (int k = 0; k < ncols; ++k) cnt += (P.w(HB_SYM_GAP, wd, first + step * k) >> lane) & 1u;
        valid[wd] = __ballot_sync(HB_FULL, 2 * cnt <= ncols) & P.rowmask[wd];      // :8846
        nv += __popc(valid[wd]);
    }
    if (nv < 2) return -1;                                                          // :8850
    const double thr2 = thr - 0.1;
    double total = 0.0;
    int first_cand = -1;
    for (int k0 = 0; k0 < ncols; k0 += 32) {
        const int k = k0 + lane;
        double ratio = 0.0;
        if (k < ncols) {
            const int col = first + step * k;
            int mx = 0;
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                int c = 0;
                for (int wd = 0; wd < P.nw; ++wd) c += __popc(P.w(s, wd, col) & valid[wd]);
                mx = max(mx, c);
            }
            ratio = mx > 0 ? (double)mx / (double)nv : 0.0;                         // :8865-8868
        }
        const uint32_t cb = __ballot_sync(HB_FULL, k < ncols && ratio >= thr2);     // :8871
        if (first_cand == -1 && cb) first_cand = first + step * (k0 + __ffs(cb) - 1);
        const int lim = min(32, ncols - k0);
        for (int j = 0; j < lim; ++j) total += __shfl_sync(HB_FULL, ratio, j);      // ordered sum :8874
    }
    const double avg = total / (double)ncols;                                        // :8877
    return avg >= thr ? first_cand : -1;
}

// slide windows of wsize over the valid-column list (Util.py:8958-8965, 9010-9017, 9072-9079, 9130-9137)
__device__ __noinline__ int slide_windows(const Planes& P, const int* vc, int n, int wsize, int step, double thr, int lane) {
    for (int i = 0; i + wsize <= n; ++i) {
        const int first = vc[i], last = vc[i + wsize - 1];
        // step>0: range(first, last+1) ; step<0: range(first, last+1, -1) = first .. last+2
        const int ncols = step > 0 ? (last + 1 - first) : (first - last - 1);
        int res = -1;
        if (ncols > 0) res = window_homology(P, first, ncols, step, thr, lane);
        else {
            // empty window: len(valid_rows) == all rows (0 <= 0); total/0 raises in Python.
            // cannot occur for wsize >= 10 (ncols >= wsize-2); treated as non-homologous.
            res = -1;
        }
        if (res != -1) return res;
    }
    return -1;
}

// search_boundary_homo_v3 (Util.py:8887-9143)
__device__ __noinline__ int search_boundary_v3(const Planes& P, int vthr, int pos, int Cc, bool is_start, double thr,
                                  int int_win, int out_win, int* vc, int lane) {
    int cur, n, w, nb;
    if (is_start) {
        n = collect_valid(P, vthr, pos, +1, 0, (Cc + 1) / 2, vc, lane);             // col < C/2  :8924
        w = min(n, int_win);
        cur = (w < 10) ? -1 : slide_windows(P, vc, n, w, +1, thr, lane);            // :8954-8969
        n = (cur >= 0) ? collect_valid(P, vthr, cur, -1, 0, Cc, vc, lane) : 0;      // :8976
        reverse_list(vc, n, lane);                                                   // :9003
        w = min(n, out_win);
        if (w < 10) cur = -1;                                                        // :9006-9008
        else {
            nb = slide_windows(P, vc, n, w, +1, thr, lane);
            if (nb != -1) { if (nb < 10) cur = -1; else cur = nb; }                  // :9018-9027
        }
    } else {
        n = collect_valid(P, vthr, pos, +1, 0, Cc, vc, lane);                        // :9036
        reverse_list(vc, n, lane);                                                   // :9065
        cur = pos;
        w = min(n, out_win);
        if (w < 10) cur = -1;                                                        // :9068-9070
        else {
            nb = slide_windows(P, vc, n, w, -1, thr, lane);
            if (nb != -1) { if (Cc - nb < 10) cur = -1; else cur = nb; }             // :9080-9089
        }
        // col >= C/2  :9096
        n = (cur >= 0) ? collect_valid(P, vthr, cur, -1, (Cc + 1) / 2, Cc, vc, lane) : 0;
        w = min(n, int_win);
        cur = (w < 10) ? -1 : slide_windows(P, vc, n, w, -1, thr, lane);             // :9126-9141
    }
    return cur;
}

// consensus vote of one column (Util.py:9314-9339); returns symbol or -1 (column emits nothing)
__device__ __forceinline__ int vote_column(const Planes& P, int col, int half) {
    int cnt[6], fs[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        int c = 0, f = 1 << 20;
        for (int wd = P.nw - 1; wd >= 0; --wd) {
            uint32_t x = P.w(s, wd, col);
            c += __popc(x);
            if (x) f = wd * 32 + __ffs(x) - 1;
        }
        cnt[s] = c; fs[s] = f;   // first-seen row == dict insertion order (:9257-9266)
    }
    int mxc = 0, mxs = -1, mxf = 1 << 21;
#pragma unroll
    for (int s = 0; s < 6; ++s)
        if (cnt[s] > mxc || (cnt[s] == mxc && cnt[s] > 0 && fs[s] < mxf)) { mxc = cnt[s]; mxs = s; mxf = fs[s]; }
    if (mxc >= half) return mxs != HB_SYM_GAP ? mxs : -1;
    mxc = 0; mxs = -1; mxf = 1 << 21;
#pragma unroll
    for (int s = 0; s < 5; ++s)
        if (cnt[s] > mxc || (cnt[s] == mxc && cnt[s] > 0 && fs[s] < mxf)) { mxc = cnt[s]; mxs = s; mxf = fs[s]; }
    return mxs;
}

// TSDsearch_v5 (Util.py:2460-2492) on one row: returns 1 iff a TSD is found
__device__ __noinline__ int tsd_search_v5(const uint8_t* row, int Cc, int bs, int be, int plant) {
    // ungapped neighbourhoods (get_boundary_ungap_str :2407-2429), 4 bits per symbol
    unsigned long long L = 0, Rr = 0;   // L[i]: i-th non-gap going left from bs-1 ; Rr[i]: i-th going right from be+1
    int nL = 0, nR = 0;
    for (int c = bs - 1; c >= 0 && nL < 11; --c) { int s = row[c]; if (s != HB_SYM_GAP) { L |= (unsigned long long)s << (4 * nL); ++nL; } }
    for (int c = be + 1; c < Cc && nR < 11; ++c) { int s = row[c]; if (s != HB_SYM_GAP) { Rr |= (unsigned long long)s << (4 * nR); ++nR; } }
    uint32_t F = 0, E = 0;              // F[i]: i-th non-gap going right from bs ; E[i]: i-th going left from be
    int nF = 0, nE = 0;
    for (int c = bs; c < Cc && nF < 5; ++c) { int s = row[c]; if (s != HB_SYM_GAP) { F |= (uint32_t)s << (4 * nF); ++nF; } }
    for (int c = be; c >= 0 && nE < 5; --c) { int s = row[c]; if (s != HB_SYM_GAP) { E |= (uint32_t)s << (4 * nE); ++nE; } }
    // symbol codes: A0 C1 G2 T3
    const bool f3_ccc = nF >= 3 && (F & 0xfff) == 0x111;
    const bool l3_ggg = nE >= 3 && (E & 0xfff) == 0x222;
    // first_5bp == 'CACTA' : F = C,A,C,T,A -> nibbles (low first) 1,0,1,3,0
    const bool f5_cacta = nF >= 5 && F == 0x03101u;
    const bool f5_cactg = nF >= 5 && F == 0x23101u;
    // last_5bp == 'TAGTG' forward; E holds it reversed: G,T,G,A,T -> 2,3,2,0,3
    const bool l5_tagtg = nE >= 5 && E == 0x30232u;
    // 'CAGTG' reversed: G,T,G,A,C -> 2,3,2,0,1
    const bool l5_cagtg = nE >= 5 && E == 0x10232u;
    int found = 0;
    const int lens[9] = {11, 10, 9, 8, 6, 5, 4, 3, 2};
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int k = lens[i];
        if (nL < k || nR < k) continue;
        // left_tsd[t] = L[k-1-t], right_tsd[t] = Rr[t]
        int mm = 0;
        for (int t = 0; t < k; ++t) mm += (((L >> (4 * (k - 1 - t))) & 15ull) != ((Rr >> (4 * t)) & 15ull));
        if (mm == 0) {
            bool ok;
            if (k != 2 && k != 3 && k != 4) ok = true;
            else if (k == 4) ok = ((Rr & 0xffffull) == 0x0033ull);                        // TTAA
            else if (k == 2) ok = ((Rr & 0xffull) == 0x03ull) || (plant == 0 && f3_ccc && l3_ggg);   // TA
            else ok = ((Rr & 0xfffull) == 0x003ull) || ((Rr & 0xfffull) == 0x033ull) ||  // TAA, TTA
                      (plant == 1 && ((f5_cacta && l5_tagtg) || (f5_cactg && l5_cagtg)));
            if (ok) found = 1;
        } else if (k >= 8 && mm <= 1) found = 1;
    }
    return found;
}

__device__ __noinline__ int lev5(const char* a, int la, const char* b, int lb) {
    int prev[8], cur[8];
    for (int j = 0; j <= lb; ++j) prev[j] = j;
    for (int i = 1; i <= la; ++i) {
        cur[0] = i;
        for (int j = 1; j <= lb; ++j) {
            int v = min(prev[j] + 1, cur[j - 1] + 1);
            cur[j] = min(v, prev[j - 1] + (a[i - 1] != b[j - 1]));
        }
        for (int j = 0; j <= lb; ++j) prev[j] = cur[j];
    }
    return prev[lb];
}

__device__ __forceinline__ bool starts_with(const char* s, int n, const char* p, int lp) {
    if (n < lp) return false;
    for (int i = 0; i < lp; ++i) if (s[i] != p[i]) return false;
    return true;
}

// anchors of one row (Util.py:9158-9181): true iff both 20-mers hit; columns in *as, *ae
__device__ __noinline__ bool row_anchor(const uint8_t* __restrict__ row, int Cc, int cap, const Pattern& P1, const Pattern& P2,
                           Match* __restrict__ ml, int* cnt, int lane, int* as, int* ae) {
    int ub, n;
    anchor_scan_row(row, Cc, P1, P2, ml, cap, cnt, lane, &ub, &n);
    const int n1 = min(cap, cnt[0]), n2 = min(cap, cnt[1]);
    __syncwarp();
    if (n1 <= 0 || n2 <= 0) return false;
    const int fs = first_group_best_start(ml, n1, lane);
    const int le = last_group_best_end(ml + cap, n2, lane);
    *as = column_of_ungapped(row, Cc, fs, ub, lane);
    *ae = column_of_ungapped(row, Cc, le - 1, ub, lane);
    return true;
}

// Does the row hold at least one end with distance <= 2 for BOTH patterns -- i.e. would row_anchor succeed?  Same lane chunks
// and the same 40-symbol warm-up as anchor_scan_row, but no match lists, no start search and the pattern masks in registers:
// used to look for the first anchored row of a matrix with many warps at once (hb_judge_a2_kernel).
__device__ __noinline__ bool row_has_anchors(const uint8_t* __restrict__ row, int Cc, const Pattern& P1, const Pattern& P2, int lane) {
    const int chunk = (Cc + 31) >> 5;
    const int lo = min(Cc, lane * chunk), hi = min(Cc, lo + chunk);
    const uint32_t a0 = P1.peq[0], a1 = P1.peq[1], a2 = P1.peq[2], a3 = P1.peq[3], a4 = P1.peq[4];
    const uint32_t b0 = P2.peq[0], b1 = P2.peq[1], b2 = P2.peq[2], b3 = P2.peq[3], b4 = P2.peq[4];
    const uint32_t top1 = 1u << (P1.m - 1), top2 = 1u << (P2.m - 1);
    Myers M1, M2; M1.init(P1.m); M2.init(P2.m);
    int c0 = lo, need = 40;
    while (c0 > 0 && need > 0) { --c0; need -= (row[c0] != HB_SYM_GAP); }
    bool f1 = false, f2 = false;
    for (int c = c0; c < hi; ++c) {
        const int s = row[c];
        if (s == HB_SYM_GAP) continue;
        const uint32_t e1 = s == 0 ? a0 : s == 1 ? a1 : s == 2 ? a2 : s == 3 ? a3 : a4;
        const uint32_t e2 = s == 0 ? b0 : s == 1 ? b1 : s == 2 ? b2 : s == 3 ? b3 : b4;
        M1.step(e1, top1); M2.step(e2, top2);
        if (c >= lo) { f1 |= M1.score <= 2; f2 |= M2.score <= 2; }
    }
    return __any_sync(HB_FULL, f1) && __any_sync(HB_FULL, f2);
}

// six bit-planes of the listed rows (row k of the list = bit k&31 of word k>>5)
__device__ __noinline__ void build_planes(Planes& P, uint32_t* __restrict__ pw, const uint8_t* __restrict__ m, int C, int Cc,
                             const unsigned short* rows, int row_num, int lane) {
    P.p = pw; P.stride = C; P.nw = (row_num + 31) >> 5; P.row_num = row_num;
    for (int wd = 0; wd < 4; ++wd) {
        int rem = row_num - wd * 32;
        P.rowmask[wd] = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
    }
    __syncwarp();
    for (int col = lane; col < Cc; col += 32) {
        for (int wd = 0; wd < P.nw; ++wd) {
            uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
            const int kend = min(32, row_num - wd * 32);
#pragma unroll 8
            for (int kk = 0; kk < kend; ++kk) {
                const int s = m[(size_t)rows[wd * 32 + kk] * C + col];
                const uint32_t bit = 1u << kk;
                a0 |= (s == 0) ? bit : 0u; a1 |= (s == 1) ? bit : 0u; a2 |= (s == 2) ? bit : 0u;
                a3 |= (s == 3) ? bit : 0u; a4 |= (s == 4) ? bit : 0u; a5 |= (s == 5) ? bit : 0u;
            }
            pw[(size_t)(0 * 4 + wd) * C + col] = a0; pw[(size_t)(1 * 4 + wd) * C + col] = a1;
            pw[(size_t)(2 * 4 + wd) * C + col] = a2; pw[(size_t)(3 * 4 + wd) * C + col] = a3;
            pw[(size_t)(4 * 4 + wd) * C + col] = a4; pw[(size_t)(5 * 4 + wd) * C + col] = a5;
        }
    }
    __syncwarp();
}

// planes prebuilt over ALL rows of the matrix (hb_msa.cu::msa_clean_kernel): keeping a row list = masking the other rows
__device__ __noinline__ void planes_select(Planes& P, const uint32_t* __restrict__ pw, int C, int R, const unsigned short* rows,
                                           int row_num, int lane) {
    P.p = pw; P.stride = C; P.nw = (R + 31) >> 5; P.row_num = row_num;
#pragma unroll
    for (int wd = 0; wd < 4; ++wd) {
        uint32_t mk = 0;
        for (int k = lane; k < row_num; k += 32) mk |= (rows[k] >> 5) == wd ? 1u << (rows[k] & 31) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mk |= __shfl_xor_sync(HB_FULL, mk, o);
        P.rowmask[wd] = mk;
    }
    __syncwarp();
}

// rows (in order) that have a non-gap symbol in [lo1,hi1) and in [lo2,hi2); at most MAX_KEPT
// ("len(...) > 100: break", Util.py:9201 / :10017)
__device__ __noinline__ int select_rows(const uint8_t* __restrict__ m, int R, int C, int lo1, int hi1, int lo2, int hi2,
                           unsigned short* rows, int lane) {
    int row_num = 0;
    for (int r0 = 0; r0 < R && row_num < MAX_KEPT; r0 += 32) {
        const int r = r0 + lane;
        bool ok = false;
        if (r < R) {
            const uint8_t* row = m + (size_t)r * C;
            bool a1 = false, a2 = false;
            for (int c = lo1; c < hi1; ++c) a1 |= (row[c] != HB_SYM_GAP);
            for (int c = lo2; c < hi2; ++c) a2 |= (row[c] != HB_SYM_GAP);
            ok = a1 && a2;
        }
        const uint32_t bal = __ballot_sync(HB_FULL, ok);
        const int pos = row_num + __popc(bal & lanemask_lt(lane));
        if (ok && pos < MAX_KEPT) rows[pos] = (unsigned short)r;
        row_num = min(MAX_KEPT, row_num + __popc(bal));
    }
    __syncwarp();
    return row_num;
}

// ---- Helitron helpers: search_boundary_homo_v4 (Util.py:8556-8824) ---------------------------
__device__ __noinline__ double ordered_ratio_avg(const Planes& P, const int* vc, int w, int lane) {
    // mean of the stored per-column max_homo_ratio (denominator row_num, :8607) in list order
    double total = 0.0;
    for (int k0 = 0; k0 < w; k0 += 32) {
        const int k = k0 + lane;
        double ratio = 0.0;
        if (k < w) {
            const int col = vc[k];
            int mx = 0;
#pragma unroll
            for (int s = 0; s < 5; ++s) mx = max(mx, P.count(s, col));
            ratio = mx > 0 ? (double)mx / (double)P.row_num : 0.0;
        }
        const int lim = min(32, w - k0);
        for (int j = 0; j < lim; ++j) total += __shfl_sync(HB_FULL, ratio, j);
    }
    return total / (double)w;
}

__device__ __noinline__ bool search_boundary_v4(const Planes& P, int vthr, int pos, int Cc, bool is_start, double thr,
                                   double int_thr, double out_thr, int int_win, int out_win, int* vc, int lane,
                                   int* boundary) {
    int n, w, nb, cur;
    if (is_start) {
        n = collect_valid(P, vthr, pos, +1, 0, (Cc + 1) / 2, vc, lane);                  // :8592
        cur = pos;
        w = min(n, int_win);
        if (w < 10) cur = -1;                                                            // :8622-8624
        else { nb = slide_windows(P, vc, n, w, +1, thr, lane); if (nb != cur && nb != -1) cur = nb; }   // :8634-8637
        n = (cur - 1 >= 0) ? collect_valid(P, vthr, cur - 1, -1, 0, Cc, vc, lane) : 0;   // :8639-8669
        reverse_list(vc, n, lane);
        w = min(n, out_win);
        if (w < 10) cur = -1;                                                            // :8676-8678
        else {
            nb = slide_windows(P, vc, n, w, +1, thr, lane);
            if (nb != -1) { if (nb < 10) cur = -1; else cur = nb; }                      // :8689-8697
        }
        *boundary = cur;
        return true;                                                                     // :8699
    }
    *boundary = -1;
    n = collect_valid(P, vthr, pos + 1, +1, 0, Cc, vc, lane);                            // :8704-8732
    w = min(n, out_win);
    if (w < 10) return false;                                                            // :8739-8741
    if (ordered_ratio_avg(P, vc, w, lane) >= out_thr) return false;                      // :8754-8767
    __syncwarp();
    n = collect_valid(P, vthr, pos, -1, (Cc + 1) / 2, Cc, vc, lane);                     // :8769-8798
    w = min(n, int_win);
    if (w < 10) return false;                                                            // :8804-8806
    if (ordered_ratio_avg(P, vc, w, lane) < int_thr) {
        const int ne = vc[w - 1];
        if (ne != pos && ne != -1) return false;                                         // :8817-8823
    }
    *boundary = pos;
    return true;                                                                         // :8824
}

// first key (dict insertion order) with the largest count (Util.py:10082-10088): symbol + count
__device__ __forceinline__ int max_key(const Planes& P, int col, int* count) {
    int mxc = 0, mxs = -1, mxf = 1 << 21;
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        int c = 0, f = 1 << 20;
        for (int wd = P.nw - 1; wd >= 0; --wd) {
            uint32_t x = P.w(s, wd, col);
            c += __popc(x);
            if (x) f = wd * 32 + __ffs(x) - 1;
        }
        if (c > mxc || (c == mxc && c > 0 && f < mxf)) { mxc = c; mxs = s; mxf = f; }
    }
    *count = mxc;
    return mxs;
}

// ------------------------------------------------------------------------------------------
// Shared front of judge_boundary_v5 (Util.py:9151-9308) and judge_boundary_v9 (Util.py:9490-9597):
// anchors from the first row where both 20-mers hit, full-length member filter, bit-planes,
// outermost valid columns, search_boundary_homo_v3 for both ends.  Returns true when the caller
// should go on to the consensus; otherwise `call` already holds the verdict.
struct Front {
    Planes P;
    int row_num, hs, he, vl, vr, Cc;
};
// phase a: call defaults + anchors (:9151-9187).  false: `call` holds the verdict ('nb').
__device__ __noinline__ bool front_anchors(const HbVoteTask& t, HbVoteCall& call, const uint8_t* __restrict__ m,
                                           Match* __restrict__ matches, const uint8_t* __restrict__ cands, int* cnt,
                                           int lane, Front& F, int* as_out, int* ae_out, int max_rows = 1 << 30,
                                           bool* rows_left = nullptr) {
    const int R = t.R, C = t.C, Cc = call.clean_cols;
    F.Cc = Cc;
    call.is_te = 0; call.info = 0; call.copy_num = 0; call.align_start = -1; call.align_end = -1;
    call.homo_start = -1; call.homo_end = -1; call.cons_len = 0; call.trim_lead = 0; call.trim_trail = 0;
    call.flags = 0; call.cons_off = (int64_t)t.cons_off;
    // ---- 1. anchors (:9151-9181) ----------------------------------------------------------
    int as = -1, ae = -1;
    if (t.cand_len > 2 && Cc > 0) {
        const uint8_t* cand = cands + t.cand_off;
        const int alen = min(20, t.cand_len);
        Pattern P1, P2;
        pattern_init(P1, cand, alen);
        pattern_init(P2, cand + (t.cand_len - alen), alen);
        Match* ml = matches + t.match_off;
        bool found = false;
        const int rend = min(R, max_rows);
        for (int r = 0; r < rend && !found; ++r) found = row_anchor(m + (size_t)r * C, Cc, C + 2, P1, P2, ml, cnt, lane, &as, &ae);
        if (rows_left != nullptr) *rows_left = !found && rend < R;     // the caller has the other rows searched (hb_judge_a2_kernel)
    }
    call.align_start = as; call.align_end = ae;
    *as_out = as; *ae_out = ae;
    if (as < 0 || ae < 0) { call.info = 1; return false; }                     // :9184-9187
    return true;
}
// phase b: full-length member filter (:9196-9225) + bit-planes of the kept rows.  false: verdict 'fl1'.
__device__ __noinline__ bool front_members(const HbVoteTask& t, HbVoteCall& call, const uint8_t* __restrict__ m,
                                           uint32_t* __restrict__ planes, int as, int ae, int fl1_copy_num_is_rows,
                                           unsigned short* rows, int lane, Front& F, int prebuilt) {
    const int R = t.R, C = t.C, Cc = F.Cc;
    int e_lo, e_hi;
    py_slice(Cc, (long long)ae - 10, (ae + 10 < Cc) ? ae + 10 : Cc, &e_lo, &e_hi);
    const int row_num = select_rows(m, R, C, max(0, as - 10), min(Cc, as + 10), e_lo, e_hi, rows, lane);
    F.row_num = row_num;
    if (row_num <= 1) {                                                        // :9222-9225 / :9561-9564
        call.info = 2; call.copy_num = fl1_copy_num_is_rows ? row_num : 1;
        return false;
    }
    call.copy_num = row_num;
    if (!prebuilt) build_planes(F.P, planes + t.plane_off, m, C, Cc, rows, row_num, lane);
    else planes_select(F.P, planes + t.plane_off, C, R, rows, row_num, lane);
    return true;
}
// phase c: outermost valid columns (:9269-9294) + homology boundaries (:9297-9308).
__device__ __noinline__ bool front_boundaries(HbVoteCall& call, int as, int ae, double thr_many, int* vc, int lane,
                                              Front& F) {
    const Planes& P = F.P;
    const int Cc = F.Cc, row_num = F.row_num;
    const int vthr = row_num / 2;                                              // int(row_num/2) :9238
    const double thr = row_num <= 2 ? 0.95 : (row_num <= 5 ? 0.9 : thr_many);  // :9240-9245 / :9579-9584
    int vl = -1, vr = -1;
    for (int c0 = 0; c0 < Cc; c0 += 32) {
        const int c = c0 + lane;
        const uint32_t bal = __ballot_sync(HB_FULL, c < Cc && 2 * P.count(HB_SYM_GAP, c) <= row_num);
        if (bal) { vl = c0 + __ffs(bal) - 1; break; }
    }
    if (vl != -1) {
        for (int c0 = Cc - 1; c0 >= 0; c0 -= 32) {
            const int c = c0 - lane;
            const uint32_t bal = __ballot_sync(HB_FULL, c >= 0 && 2 * P.count(HB_SYM_GAP, c) <= row_num);
            if (bal) { vr = c0 - (__ffs(bal) - 1); break; }
        }
    }
    if (!(vl != -1 && vr != -1 && vl < vr)) { vl = -1; vr = -1; }
    F.vl = vl; F.vr = vr;
    F.hs = search_boundary_v3(P, vthr, as, Cc, true, thr, 20, 10, vc, lane);
    call.homo_start = F.hs;
    if (F.hs == -1) return false;
    F.he = search_boundary_v3(P, vthr, ae, Cc, false, thr, 20, 10, vc, lane);
    call.homo_end = F.he;
    if (F.he == -1) return false;
    return true;
}
__device__ bool judge_front_v3(const HbVoteTask& t, HbVoteCall& call, const uint8_t* __restrict__ m,
                               uint32_t* __restrict__ planes, Match* __restrict__ matches,
                               const uint8_t* __restrict__ cands, double thr_many, int fl1_copy_num_is_rows,
                               int* vc, unsigned short* rows, int* cnt, int lane, Front& F) {
    int as, ae;
    if (!front_anchors(t, call, m, matches, cands, cnt, lane, F, &as, &ae)) return false;
    if (!front_members(t, call, m, planes, as, ae, fl1_copy_num_is_rows, rows, lane, F, 0)) return false;
    return front_boundaries(call, as, ae, thr_many, vc, lane, F);
}

// column-vote consensus of columns [hs, he] written to cons (Util.py:9312-9346); returns its length
__device__ __noinline__ int vote_consensus(const Planes& P, const uint8_t* __restrict__ m, int C, const unsigned short* rows,
                              int hs, int he, int result_type, uint8_t* __restrict__ cons, int lane) {
    int ml = 0;
    for (int c0 = hs; c0 <= he; c0 += 32) {
        const int c = c0 + lane;
        int s = -1;
        if (c <= he) {
            if (result_type == 1) s = vote_column(P, c, P.row_num / 2);
            else { int x = m[(size_t)rows[0] * C + c]; s = (x != HB_SYM_GAP) ? x : -1; }
        }
        const uint32_t bal = __ballot_sync(HB_FULL, s >= 0);
        if (s >= 0) cons[ml + __popc(bal & lanemask_lt(lane))] = (uint8_t)hb_ascii_of_sym(s);
        ml += __popc(bal);
    }
    __syncwarp();
    return ml;
}

// ------------------------------------------------------------------------------------------
// kernel B: judge_boundary_v5 (TIR, Util.py:9145-9480), one warp per candidate, run as FOUR phase
// kernels (anchors | member filter + bit-planes | boundaries | consensus + TSD trimming).  As a single
// kernel the judge is ~40 k instructions and warps of one SM sit in different phases, so the warps
// starve on instruction fetch (ncu: stall_no_instruction dominant, profiles/); per phase the code is
// a few thousand instructions, registers drop and rejected candidates leave early.  The per-task
// state between phases lives behind the task's match scratch (hb_vote_state_entries()).
struct JState {
    int go, as, ae, row_num, hs, he, vl, vr, Cc, nw;
    uint32_t rowmask[4];
    unsigned short rows[MAX_KEPT + 3];
};
struct JState6 {        // Helitron judge (judge_boundary_v6)
    int go, as, ae, hs, he, Cc;
};
// a task's match scratch: MATCH_LISTS lists of C + 2 entries (two per warp that scans rows: hb_judge_v6_a_kernel runs four), then
// the judge's phase state
constexpr int MATCH_LISTS = 8;
__device__ __forceinline__ JState6* jstate6_of(Match* matches, const HbVoteTask& t) {
    return reinterpret_cast<JState6*>(matches + t.match_off + MATCH_LISTS * ((size_t)t.C + 2));
}
__device__ __forceinline__ JState* jstate_of(Match* matches, const HbVoteTask& t) {
    return reinterpret_cast<JState*>(matches + t.match_off + MATCH_LISTS * ((size_t)t.C + 2));
}
__device__ __forceinline__ void front_load(Front& F, const JState* st, uint32_t* planes, const HbVoteTask& t) {
    F.Cc = st->Cc; F.row_num = st->row_num; F.hs = st->hs; F.he = st->he; F.vl = st->vl; F.vr = st->vr;
    F.P.p = planes + t.plane_off; F.P.stride = t.C; F.P.nw = st->nw; F.P.row_num = st->row_num;
#pragma unroll
    for (int k = 0; k < 4; ++k) F.P.rowmask[k] = st->rowmask[k];
}

// Phase a tries the first A_FIRST_ROWS rows itself (one warp per task: nearly every element anchors in its first row); a
// matrix whose first rows hold no anchor pair is queued (work[0] = count, work[1..] = tasks) and its remaining rows are
// searched by hb_judge_a2_kernel with A2_WARPS warps at once -- as one warp, such a matrix cost R row scans and set the
// kernel's duration (ncu launch lists of a C5 shard: 7 ms for 1 300 candidates).
constexpr int A_FIRST_ROWS = 2, A2_WARPS = 8;
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
hb_judge_a_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ clean,
                  Match* __restrict__ matches, const uint8_t* __restrict__ cands, HbVoteCall* __restrict__ calls,
                  int* __restrict__ work) {
    __shared__ int s_cnt[WARPS_PER_CTA][2];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * WARPS_PER_CTA + wid;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    HbVoteCall call = calls[ti];          // clean_cols written by kernel A
    Front F;
    int as, ae;
    bool rows_left = false;
    const bool go = front_anchors(t, call, clean + t.clean_off, matches, cands, s_cnt[wid], lane, F, &as, &ae, A_FIRST_ROWS, &rows_left);
    __syncwarp();                          // the match scratch is dead from here on; the state sits behind it
    if (lane == 0) {
        JState* st = jstate_of(matches, t);
        st->go = go ? 1 : 0; st->as = as; st->ae = ae; st->Cc = F.Cc;
        calls[ti] = call;
        if (rows_left) work[1 + atomicAdd(&work[0], 1)] = ti;
    }
}

// rows A_FIRST_ROWS.. of the queued matrices: the warps of a CTA test rows round-robin (row_has_anchors), the smallest
// anchored row wins -- the row the sequential loop would have stopped at (:9151-9181) -- and warp 0 anchors on it.
__global__ void __launch_bounds__(A2_WARPS * 32)
hb_judge_a2_kernel(const HbVoteTask* __restrict__ tasks, const int* __restrict__ work, const uint8_t* __restrict__ clean,
                   Match* __restrict__ matches, const uint8_t* __restrict__ cands, HbVoteCall* __restrict__ calls) {
    __shared__ int s_first;
    __shared__ int s_cnt[2];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_work = work[0];
    for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const int ti = work[1 + wi];
        const HbVoteTask t = tasks[ti];
        const int R = t.R, C = t.C, Cc = calls[ti].clean_cols;
        const uint8_t* m = clean + t.clean_off;
        if (threadIdx.x == 0) s_first = R;
        __syncthreads();
        const int alen = min(20, t.cand_len);
        Pattern P1, P2;
        pattern_init(P1, cands + t.cand_off, alen);
        pattern_init(P2, cands + t.cand_off + (t.cand_len - alen), alen);
        for (int r = A_FIRST_ROWS + wid; r < R; r += A2_WARPS) {
            if (r > *(volatile int*)&s_first) break;
            if (row_has_anchors(m + (size_t)r * C, Cc, P1, P2, lane)) { if (lane == 0) atomicMin(&s_first, r); break; }
        }
        __syncthreads();
        const int fr = s_first;
        if (wid == 0 && fr < R) {
            int as = -1, ae = -1;
            row_anchor(m + (size_t)fr * C, Cc, C + 2, P1, P2, matches + t.match_off, s_cnt, lane, &as, &ae);
            __syncwarp();
            if (lane == 0) {
                JState* st = jstate_of(matches, t);
                st->as = as; st->ae = ae;
                calls[ti].align_start = as; calls[ti].align_end = ae;
                if (as >= 0 && ae >= 0) { calls[ti].info = 0; st->go = 1; }
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
hb_judge_b_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ clean,
                  uint32_t* __restrict__ planes, Match* __restrict__ matches, HbVoteCall* __restrict__ calls,
                  int fl1_copy_num_is_rows, int prebuilt) {
    __shared__ unsigned short s_rows[WARPS_PER_CTA][104];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * WARPS_PER_CTA + wid;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    JState* st = jstate_of(matches, t);
    if (!st->go) return;
    HbVoteCall call = calls[ti];
    unsigned short* rows = s_rows[wid];
    Front F;
    F.Cc = st->Cc;
    const bool go = front_members(t, call, clean + t.clean_off, planes, st->as, st->ae, fl1_copy_num_is_rows, rows, lane, F, prebuilt);
    for (int k = lane; k < F.row_num; k += 32) st->rows[k] = rows[k];
    if (lane == 0) {
        st->go = go ? 1 : 0; st->row_num = F.row_num;
        if (go) {
            st->nw = F.P.nw;
#pragma unroll
            for (int k = 0; k < 4; ++k) st->rowmask[k] = F.P.rowmask[k];
        }
        calls[ti] = call;
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
hb_judge_c_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, uint32_t* __restrict__ planes,
                  Match* __restrict__ matches, HbVoteCall* __restrict__ calls, double thr_many) {
    __shared__ int s_vc[WARPS_PER_CTA][128];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * WARPS_PER_CTA + wid;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    JState* st = jstate_of(matches, t);
    if (!st->go) return;
    HbVoteCall call = calls[ti];
    Front F;
    front_load(F, st, planes, t);
    const bool go = front_boundaries(call, st->as, st->ae, thr_many, s_vc[wid], lane, F);
    if (lane == 0) {
        st->go = go ? 1 : 0; st->hs = F.hs; st->he = F.he; st->vl = F.vl; st->vr = F.vr;
        calls[ti] = call;
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
hb_judge_d_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ clean,
                  uint32_t* __restrict__ planes, Match* __restrict__ matches, uint8_t* __restrict__ cons_buf,
                  HbVoteCall* __restrict__ calls) {
    __shared__ unsigned short s_rows[WARPS_PER_CTA][104];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * WARPS_PER_CTA + wid;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    const JState* st = jstate_of(matches, t);
    if (!st->go) return;
    HbVoteCall call = calls[ti];
    const int R = t.R, C = t.C;
    const uint8_t* m = clean + t.clean_off;
    unsigned short* rows = s_rows[wid];
    Front F;
    front_load(F, st, planes, t);
    for (int k = lane; k < F.row_num; k += 32) rows[k] = st->rows[k];
    __syncwarp();
    const Planes& P = F.P;
    const int Cc = F.Cc, hs = F.hs, he = F.he, vl = F.vl, vr = F.vr;
    // ---- 6. consensus (:9312-9346) -----------------------------------------------------------
    uint8_t* cons = cons_buf + t.cons_off;
    const int ml = vote_consensus(P, m, C, rows, hs, he, t.result_type, cons, lane);
    // ---- 7. reject rule + TIR terminal trimming supported by TSDs (:9353-9413) ---------------
    int final_lo = 0, final_hi = 0;
    if (!(hs <= vl || he >= vr) && t.te_type == 0) {
        char head[9], tail[9];
        const int nh = min(9, ml), nt = min(9, ml);
        for (int i = 0; i < nh; ++i) head[i] = (char)cons[i];
        for (int i = 0; i < nt; ++i) tail[i] = (char)cons[ml - nt + i];
        auto ends_with = [&](const char* p, int lp) {
            if (ml < lp) return false;
            for (int i = 0; i < lp; ++i) if (tail[nt - lp + i] != p[i]) return false;
            return true;
        };
        int fo[5], nfo = 0, eo[5], neo = 0;
        fo[nfo++] = 0; eo[neo++] = 0;
        if (starts_with(head, nh, "A", 1)) fo[nfo++] = 1;
        if (starts_with(head, nh, "AA", 2) || starts_with(head, nh, "TA", 2)) fo[nfo++] = 2;
        if (starts_with(head, nh, "TAA", 3) || starts_with(head, nh, "TTA", 3)) fo[nfo++] = 3;
        if (starts_with(head, nh, "TTAA", 4)) fo[nfo++] = 4;
        if (ends_with("T", 1)) eo[neo++] = 1;
        if (ends_with("TT", 2) || ends_with("TA", 2)) eo[neo++] = 2;
        if (ends_with("TAA", 3) || ends_with("TTA", 3)) eo[neo++] = 3;
        if (ends_with("TTAA", 4)) eo[neo++] = 4;
        bool have = false;
        int best_ed = 0, best_tsd = 0, best_f = 0, best_e = 0;
#pragma unroll 1
        for (int a = 0; a < nfo; ++a)
#pragma unroll 1
        for (int b = 0; b < neo; ++b) {
            const int bs = hs + fo[a], be = he - eo[b];
            int tsd = 0;
            if (bs >= 0 && bs < Cc && be >= 0 && be < Cc) {
                for (int r = lane; r < R; r += 32) {                           // ALL rows :9391
                    const uint8_t* row = m + (size_t)r * C;
                    if (row[bs] == HB_SYM_GAP || row[be] == HB_SYM_GAP) continue;
                    tsd += tsd_search_v5(row, Cc, bs, be, t.plant);
                }
                tsd = warp_sum(tsd);
            }
            if (tsd > 0) {
                int lo, hi, lo2, hi2;
                py_slice(ml, fo[a], fo[a] + 5, &lo, &hi);
                if (eo[b] == 0) py_slice(ml, -5, ml, &lo2, &hi2);
                else py_slice(ml, (long long)ml - 5 - eo[b], (long long)ml - eo[b], &lo2, &hi2);
                char f[8], e[8];
                const int nf = hi - lo, ne = hi2 - lo2;
#pragma unroll 1
                for (int i = 0; i < nf; ++i) {                                 // getReverseSequence :1635
                    char bch = (char)cons[lo + nf - 1 - i];
                    f[i] = bch == 'A' ? 'T' : bch == 'T' ? 'A' : bch == 'C' ? 'G' : bch == 'G' ? 'C' : 'N';
                }
#pragma unroll 1
                for (int i = 0; i < ne; ++i) e[i] = (char)cons[lo2 + i];
                const int ed = lev5(f, nf, e, ne);                             // :9403
                if (!have || ed < best_ed || (ed == best_ed && tsd > best_tsd)) {   // stable sort :9405
                    have = true; best_ed = ed; best_tsd = tsd; best_f = fo[a]; best_e = eo[b];
                }
            }
        }
        if (have) {                                                            // :9406-9413
            if (best_e != 0) py_slice(ml, best_f, -(long long)best_e, &final_lo, &final_hi);
            else py_slice(ml, best_f, ml, &final_lo, &final_hi);
            call.trim_lead = best_f; call.trim_trail = best_e;
        }
    }
    const int fl = max(0, final_hi - final_lo);
    call.cons_len = fl;
    call.cons_off = (int64_t)t.cons_off + final_lo;
    call.is_te = fl > 0;                                                       // :9473-9476
    if (lane == 0) calls[ti] = call;
}

// ------------------------------------------------------------------------------------------
// kernel B': judge_boundary_v6 (Helitron, Util.py:9821-10159).  Phase a anchors EVERY row (two 20-mers by Myers over the
// whole row), ~R times the work of the TIR judge's phase a: one CTA of ROW_WARPS warps per candidate, rows dealt round-robin
// to the warps (one warp per candidate left the GPU a third full at 3 000 candidates and serialised <= 100 row scans).
// Every warp scans into its own pair of match lists of the task's match scratch (MATCH_LISTS; the bit-plane area is already
// filled when the matrices come from hb_msa.cu::msa_clean_kernel).  Per-row anchors land at an[r] / an[R + r] (-1 =
// none) and are compacted in row order by warp 0: the mode's tie-break is "first inserted" (:9722-9730).
constexpr int ROW_WARPS = 4;
__global__ void __launch_bounds__(ROW_WARPS * 32)
hb_judge_v6_a_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ clean,
                     uint32_t* __restrict__ planes, Match* __restrict__ matches, int* __restrict__ anchors,
                     const uint8_t* __restrict__ cands, HbVoteCall* __restrict__ calls) {
    __shared__ int s_cnt[ROW_WARPS][2];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    HbVoteCall call = calls[ti];
    const int R = t.R, C = t.C, Cc = call.clean_cols;
    const uint8_t* m = clean + t.clean_off;
    call.is_te = 0; call.info = 0; call.copy_num = 0; call.align_start = -1; call.align_end = -1;
    call.homo_start = -1; call.homo_end = -1; call.cons_len = 0; call.trim_lead = 0; call.trim_trail = 0;
    call.flags = 0; call.cons_off = (int64_t)t.cons_off;
    // ---- 1. anchors of ALL rows, then the mode (:9837-9866) ------------------------------------
    int as = -1, ae = -1;
    const bool scan = t.cand_len > 2 && Cc > 0;
    int* an = anchors + t.anchor_off;              // [0..R) starts, [R..2R) ends
    if (scan) {
        const uint8_t* cand = cands + t.cand_off;
        const int alen = min(20, t.cand_len);
        Pattern P1, P2;
        pattern_init(P1, cand, alen);
        pattern_init(P2, cand + (t.cand_len - alen), alen);
        static_assert(MATCH_LISTS >= 2 * ROW_WARPS, "one pair of match lists per scanning warp");
        Match* ml = matches + t.match_off + (size_t)wid * 2 * ((size_t)C + 2);
        const int nw = ROW_WARPS;
        for (int r = wid; r < R && wid < nw; r += nw) {
            int a = -1, b = -1;
            const bool hit = row_anchor(m + (size_t)r * C, Cc, C + 2, P1, P2, ml, s_cnt[wid], lane, &a, &b);
            if (lane == 0) { an[r] = hit ? a : -1; an[R + r] = hit ? b : -1; }
            __syncwarp();
        }
    }
    __syncthreads();
    if (wid != 0) return;
    if (scan) {
        // rows with both anchors, in row order (in place: a chunk's survivors never land beyond the chunk)
        int na = 0;
        for (int r0 = 0; r0 < R; r0 += 32) {
            const int r = r0 + lane;
            const int a = r < R ? an[r] : -1, b = r < R ? an[R + r] : -1;
            const uint32_t bal = __ballot_sync(HB_FULL, a >= 0);
            __syncwarp();
            if (a >= 0) { const int pos = na + __popc(bal & lanemask_lt(lane)); an[pos] = a; an[R + pos] = b; }
            na += __popc(bal);
            __syncwarp();
        }
        // most_common_element (:9722-9730): highest count, ties -> first inserted
        for (int which = 0; which < 2; ++which) {
            const int* arr = an + which * R;
            unsigned long long best = 0;
            for (int i = lane; i < na; i += 32) {
                const int v = arr[i];
                int cnt = 0, first = i;
                for (int j = 0; j < na; ++j) { if (arr[j] == v) { ++cnt; if (j < first) first = j; } }
                const unsigned long long key = ((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)first);
                best = key > best ? key : best;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(HB_FULL, best, o); best = x > best ? x : best; }
            const int val = na > 0 ? arr[0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFull)] : -1;
            if (which == 0) as = val; else ae = val;
        }
    }
    call.align_start = as; call.align_end = ae;
    const bool go = !(as < 0 || ae < 0);
    if (!go) call.info = 1;                                                                 // :9870-9873
    __syncwarp();                          // the match scratch is dead from here on; the state sits behind it
    if (lane == 0) {
        JState6* st = jstate6_of(matches, t);
        st->go = go ? 1 : 0; st->as = as; st->ae = ae; st->Cc = Cc;
        calls[ti] = call;
    }
}

// phase B: start / end member sets, then the end and the start boundary (search_boundary_homo_v4)
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
hb_judge_v6_b_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ clean,
                     uint32_t* __restrict__ planes, Match* __restrict__ matches, HbVoteCall* __restrict__ calls, int prebuilt) {
    __shared__ int s_vc[WARPS_PER_CTA][128];
    __shared__ unsigned short s_rows[WARPS_PER_CTA][104];
    __shared__ unsigned short s_rows2[WARPS_PER_CTA][104];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * WARPS_PER_CTA + wid;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    JState6* st = jstate6_of(matches, t);
    if (!st->go) return;
    HbVoteCall call = calls[ti];
    const int R = t.R, C = t.C, Cc = st->Cc, as = st->as, ae = st->ae;
    const uint8_t* m = clean + t.clean_off;
    int* vc = s_vc[wid];
    unsigned short* rows = s_rows[wid];
    unsigned short* rows2 = s_rows2[wid];
    __syncwarp();
    if (lane == 0) st->go = 0;             // every early return below is a final verdict
    // ---- 2. start / end member sets (:9884-9936): the loop stops for both once either exceeds 100
    int ns = 0, ne = 0;
    {
        const int s_lo = max(0, as - 1), s_hi = min(Cc, as + 1);
        int e_lo, e_hi;
        py_slice(Cc, (long long)ae - 1, (ae + 1 < Cc) ? ae + 1 : Cc, &e_lo, &e_hi);
        for (int r0 = 0; r0 < R && ns <= 100 && ne <= 100; r0 += 32) {
            const int r = r0 + lane;
            bool fs = false, fe = false;
            if (r < R) {
                const uint8_t* row = m + (size_t)r * C;
                for (int c = s_lo; c < s_hi; ++c) fs |= (row[c] != HB_SYM_GAP);
                for (int c = e_lo; c < e_hi; ++c) fe |= (row[c] != HB_SYM_GAP);
            }
            const uint32_t bs = __ballot_sync(HB_FULL, fs), be = __ballot_sync(HB_FULL, fe);
            const int ps = ns + __popc(bs & lanemask_lt(lane)), pe = ne + __popc(be & lanemask_lt(lane));
            const bool processed = r < R && ps <= 100 && pe <= 100;        // loop-top checks :9885-9888
            const uint32_t bsp = __ballot_sync(HB_FULL, fs && processed), bep = __ballot_sync(HB_FULL, fe && processed);
            if (fs && processed) rows[ps] = (unsigned short)r;
            if (fe && processed) rows2[pe] = (unsigned short)r;
            ns += __popc(bsp); ne += __popc(bep);
            if (__ballot_sync(HB_FULL, r < R && !processed)) break;
        }
        __syncwarp();
    }
    if (ne <= 0) { if (lane == 0) calls[ti] = call; return; }                              // :9940-9941
    Planes P;
    uint32_t* pw = planes + t.plane_off;
    int he = -1, hs = -1;
    {   // end first (:9972-9979)
        if (prebuilt) planes_select(P, pw, C, R, rows2, ne, lane); else build_planes(P, pw, m, C, Cc, rows2, ne, lane);
        const double thr = ne <= 2 ? 0.95 : (ne <= 5 ? 0.9 : 0.7);
        const double ithr = ne <= 2 ? 0.9 : (ne <= 5 ? 0.85 : 0.65);
        const double othr = ne <= 2 ? 0.95 : (ne <= 5 ? 0.9 : 0.7);
        const bool ok = search_boundary_v4(P, ne / 2, ae, Cc, false, thr, ithr, othr, 20, 10, vc, lane, &he);
        if (!ok) { if (lane == 0) calls[ti] = call; return; }
        call.homo_end = he;
    }
    {   // then start (:9991-10008)
        if (prebuilt) planes_select(P, pw, C, R, rows, ns, lane); else build_planes(P, pw, m, C, Cc, rows, ns, lane);
        const double thr = ns <= 2 ? 0.95 : (ns <= 5 ? 0.9 : 0.7);
        const double ithr = ns <= 2 ? 0.9 : (ns <= 5 ? 0.85 : 0.65);
        const double othr = ns <= 2 ? 0.95 : (ns <= 5 ? 0.9 : 0.7);
        search_boundary_v4(P, ns / 2, as, Cc, true, thr, ithr, othr, 20, 10, vc, lane, &hs);
        call.homo_start = hs;
    }
    if (lane == 0) { st->go = 1; st->hs = hs; st->he = he; calls[ti] = call; }
}

// phase C: full-length members, consensus, ATC ... CTRRT motif
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
hb_judge_v6_c_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ clean,
                     uint32_t* __restrict__ planes, Match* __restrict__ matches, uint8_t* __restrict__ cons_buf,
                     HbVoteCall* __restrict__ calls, int prebuilt) {
    __shared__ unsigned short s_rows[WARPS_PER_CTA][104];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = blockIdx.x * WARPS_PER_CTA + wid;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    const JState6* st = jstate6_of(matches, t);
    if (!st->go) return;
    HbVoteCall call = calls[ti];
    const int R = t.R, C = t.C, Cc = st->Cc, hs = st->hs, he = st->he;
    const uint8_t* m = clean + t.clean_off;
    unsigned short* rows = s_rows[wid];
    Planes P;
    uint32_t* pw = planes + t.plane_off;
    // ---- 3. full-length members (+-1 around the homology boundaries, :10012-10038) -------------
    int row_num;
    {
        int s_lo, s_hi, e_lo, e_hi;
        py_slice(Cc, hs - 1 >= 0 ? hs - 1 : 0, (long long)hs + 1, &s_lo, &s_hi);
        py_slice(Cc, (long long)he - 1, (he + 1 < Cc) ? he + 1 : Cc, &e_lo, &e_hi);
        row_num = select_rows(m, R, C, s_lo, s_hi, e_lo, e_hi, rows, lane);
    }
    if (row_num <= 0) { if (lane == 0) calls[ti] = call; return; }
    call.copy_num = row_num;
    if (prebuilt) planes_select(P, pw, C, R, rows, row_num, lane); else build_planes(P, pw, m, C, Cc, rows, row_num, lane);
    // ---- 4. consensus with 'N' for weak columns (:10078-10095) and +-1 bp extension (:10097-10130)
    uint8_t* cons = cons_buf + t.cons_off;
    const int half = row_num / 2;
    int lext = -1, rext = -1;
    for (int c0 = hs - 1; c0 >= 0 && lext < 0; c0 -= 32) {
        const int c = c0 - lane;
        int cnt = 0, sy = -1;
        if (c >= 0) sy = max_key(P, c, &cnt);
        const uint32_t bal = __ballot_sync(HB_FULL, c >= 0 && cnt >= half && sy != HB_SYM_GAP);
        if (bal) lext = __shfl_sync(HB_FULL, sy, __ffs(bal) - 1);
    }
    for (int c0 = he + 1; c0 < Cc && rext < 0; c0 += 32) {
        const int c = c0 + lane;
        int cnt = 0, sy = -1;
        if (c < Cc) sy = max_key(P, c, &cnt);
        const uint32_t bal = __ballot_sync(HB_FULL, c < Cc && cnt >= half && sy != HB_SYM_GAP);
        if (bal) rext = __shfl_sync(HB_FULL, sy, __ffs(bal) - 1);
    }
    int ml = 0;
    if (lext >= 0) { if (lane == 0) cons[0] = (uint8_t)hb_ascii_of_sym(lext); ml = 1; }
    for (int c0 = hs; c0 <= he; c0 += 32) {
        const int c = c0 + lane;
        int out = -1;                                    // -1 nothing, else ASCII
        if (c <= he) {
            int cnt; const int sy = max_key(P, c, &cnt);
            if (cnt >= half) out = (sy != HB_SYM_GAP) ? hb_ascii_of_sym(sy) : -1;
            else out = 'N';
        }
        const uint32_t bal = __ballot_sync(HB_FULL, out >= 0);
        if (out >= 0) cons[ml + __popc(bal & lanemask_lt(lane))] = (uint8_t)out;
        ml += __popc(bal);
    }
    if (rext >= 0) { if (lane == 0) cons[ml] = (uint8_t)hb_ascii_of_sym(rext); ++ml; }
    __syncwarp();
    // ---- 5. ATC ... CTRRT motif (:10135-10150) ---------------------------------------------------
    int final_lo = 0, final_hi = 0;
    {
        int f_lo, f_hi, l_lo, l_hi;
        py_slice(ml, 0, 10, &f_lo, &f_hi);
        py_slice(ml, -10, ml, &l_lo, &l_hi);
        char head[10], tail[10];
        for (int i = 0; i < f_hi - f_lo; ++i) head[i] = (char)cons[f_lo + i];
        for (int i = 0; i < l_hi - l_lo; ++i) tail[i] = (char)cons[l_lo + i];
        const char motifs[4][5] = {{'C', 'T', 'A', 'G', 'T'}, {'C', 'T', 'A', 'A', 'T'}, {'C', 'T', 'G', 'G', 'T'}, {'C', 'T', 'G', 'A', 'T'}};
        for (int q = 0; q < 4; ++q) {
            int end_index = -1;
            for (int k = (l_hi - l_lo) - 5; k >= 0; --k) {
                bool eq = true;
                for (int z = 0; z < 5; ++z) eq &= (tail[k + z] == motifs[q][z]);
                if (eq) { end_index = k; break; }
            }
            if (end_index == -1) continue;
            int start_index = -1;
            for (int k = 0; k + 3 <= f_hi - f_lo; ++k)
                if (head[k] == 'A' && head[k + 1] == 'T' && head[k + 2] == 'C') { start_index = k; break; }
            if (start_index != -1) {
                const int cut = 10 - (end_index + 3) - 1;
                if (cut == 0) py_slice(ml, start_index + 1, ml, &final_lo, &final_hi);
                else py_slice(ml, start_index + 1, -(long long)cut, &final_lo, &final_hi);
                break;
            }
        }
    }
    const int fl = max(0, final_hi - final_lo);
    call.cons_len = fl;
    call.cons_off = (int64_t)t.cons_off + final_lo;
    call.is_te = fl > 0;
    if (lane == 0) calls[ti] = call;
}

// ------------------------------------------------------------------------------------------
// non-LTR helpers (Util.py:9604-9664, find_tail_polyA :10832, find_longest_tandem_repeat_tail :9732)

constexpr int NL_SUB = 76;     // <= 50 symbols before the 5' end (align_seq[left_pos:end_5]), oldest first; rest = fill 6
constexpr int NL_TSD = 24;     // the <= 20 symbols from the 3' end on (the TSD source); rest = fill 7
constexpr int NL_TSD_MAX = 20;

// The per-copy 3'-end / TSD search, part 1, for up to 32 rows (r0 + lane < r_end) at once: the serial pass over each row -- 5'
// end, poly-A or tandem 3' end (find_tail_polyA :10832, find_longest_tandem_repeat_tail :9732) -- and the two short symbol
// windows the TSD search needs, copied to shared memory (sub / tsd of this lane).  The rows are walked one lane each, but
// the matrix is read through a 32 x 128 tile staged in shared memory with coalesced loads (row stride 132: the lanes'
// byte reads fall in distinct banks) and the 64-symbol history ring lives in shared memory too: read straight from
// global / local memory every step of the walk touches 32 different cache lines per warp and the kernel is bound by L1
// wavefronts.  Returns true iff the TSD search has to run for this lane's row (:9632); *avail = symbols available from
// end_3 on (< 20 only where the row ends).  hs0 / he: homology boundary columns.
constexpr int NL_TILE = 128, NL_TSTRIDE = 132, NL_RSTRIDE = 68;
__device__ __noinline__ bool nonltr_rows_scan(const uint8_t* __restrict__ m, int C, int Cc, int r0, int r_end, int hs0, int he,
                                              int lane, uint8_t* tile, uint8_t* ring_all, int* end5_out, int* end3_out,
                                              int* sl_out, int* avail_out, uint8_t* sub, uint8_t* tsd) {
    uint8_t* ring = ring_all + lane * NL_RSTRIDE;      // last 64 non-gap symbols, index = ungapped position & 63
    const bool active = r0 + lane < r_end;
    int n = 0, end_5 = -1, h3 = -1, sl = 0;
    int run = 0, pa_end = -1;
    bool done = !active;
    const int n_rows = min(32, r_end - r0);
    for (int c0 = 0; c0 < Cc; c0 += NL_TILE) {
        if (__all_sync(HB_FULL, done)) break;
        const int w = min(NL_TILE, Cc - c0);
        for (int rr = 0; rr < n_rows; ++rr) {
            const uint8_t* src = m + (size_t)(r0 + rr) * C + c0;
#pragma unroll
            for (int z = 0; z < NL_TILE / 32; ++z) {
                const int cc = z * 32 + lane;
                if (cc < w) tile[rr * NL_TSTRIDE + cc] = src[cc];
            }
        }
        __syncwarp();
        const uint8_t* mine = tile + lane * NL_TSTRIDE;
        for (int cc = 0; cc < w; ++cc) {
            if (done) continue;
            const int c = c0 + cc;
            if (c == hs0) {
                end_5 = n;
                sl = min(50, n);
                for (int t = 0; t < sl; ++t) sub[t] = ring[(n - sl + t) & 63];
            }
            if (c == he) h3 = n;
            const uint8_t s = mine[cc];
            if (s == HB_SYM_GAP) continue;
            ring[n & 63] = s;
            ++n;
            // cur_seq = align_seq[0 : h3 + 10]: A-runs only count while inside it
            if (h3 < 0 || n <= h3 + 10) {
                run = (s == 0) ? run + 1 : 0;
                if (run >= 6) pa_end = n;
            }
            if (h3 >= 0 && n >= h3 + 30) done = true;
        }
        __syncwarp();
    }
    *end3_out = -1; *end5_out = -1; *sl_out = 0; *avail_out = 0;
    if (!active) return false;
    for (int t = sl; t < NL_SUB; ++t) sub[t] = 6;      // (sub[0 .. sl) was written when the walk passed column hs0)
    for (int t = 0; t < NL_TSD; ++t) tsd[t] = 7;
    if (h3 < 0) h3 = n;            // he beyond the last column never happens (he < Cc); keeps the maths defined
    if (end_5 < 0) end_5 = n;
    const int cl = min(n, h3 + 10);
    int end_3 = pa_end;                                                    // find_tail_polyA
    if (end_3 < 0) {                                                       // find_longest_tandem_repeat_tail
        const int tl = min(30, cl), t0 = cl - tl;
        int max_len = 0;
        for (int ul = 2; ul <= 6; ++ul) {
            for (int start = tl - ul * 4; start >= 0; --start) {
                int count = 1;
                for (int i = 1; i < (tl - start) / ul; ++i) {
                    bool eq = true;
                    for (int z = 0; z < ul; ++z) eq &= (ring[(t0 + start + i * ul + z) & 63] == ring[(t0 + start + z) & 63]);
                    if (!eq) break;
                    ++count;
                }
                if (count >= 4 && ul * count > max_len) { max_len = ul * count; end_3 = t0 + start + ul * count; }
            }
        }
    }
    *end3_out = end_3;
    *end5_out = end_5;
    *sl_out = sl;
    if (abs(end_3 - h3) > 10) return false;                                // :9632
    if (end_3 == -1 || end_5 == -1) return false;
    // the scan stops 30 symbols past h3, so n >= end_3 + 20 unless the row really ends here; the ring still holds them
    const int avail = min(NL_TSD_MAX, max(0, n - end_3));
    for (int t = 0; t < avail; ++t) tsd[t] = ring[(end_3 + t) & 63];
    *avail_out = avail;
    return true;
}

// part 2 (the whole warp on one row): TSDs of 20 .. 8 symbols, the window positions spread over the lanes, first hit of the
// longest length by ballot (:9640-9658).  Returns the hit's end (i + k) or -1.
// fuzzysearch.find_near_matches(TSD, kmer, max_l_dist=1) is non-empty for |TSD| == |kmer| == k iff Hamming <= 1, or TSD minus
// one symbol equals kmer[1:] / kmer[:-1].  All 13 lengths are answered from three 20-bit mismatch masks of the position,
//   M  bit t: tsd[t]     != kmer[t]        M1 bit t: tsd[t] != kmer[t + 1]        M2 bit t: tsd[t + 1] != kmer[t]
// built four symbols at a time (symbols are bytes < 8: (x ^ y) + 0x7f sets bit 7 of every differing byte, one multiply
// gathers the four flags): Hamming_k = popc(M & ones(k)); "minus one symbol == kmer[1:]" = M1 clear below a and M clear on
// (a, k) with a the length of the common prefix; "== kmer[:-1]" = M clear below a and M2 clear on [a, k - 1).
__device__ __forceinline__ uint32_t diff4(uint32_t x, uint32_t y) {            // 4 flags: byte z of x != byte z of y
    return (((((x ^ y) + 0x7f7f7f7fu) >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ void diff_masks(const uint32_t* __restrict__ tw, const uint32_t* __restrict__ subw, int i,
                                           uint32_t& M, uint32_t& M1, uint32_t& M2) {
    const int j = i >> 2, sh = 8 * (i & 3);
    uint32_t w[7];
#pragma unroll
    for (int z = 0; z < 7; ++z) w[z] = subw[j + z];
    M = 0; M1 = 0; M2 = 0;
#pragma unroll
    for (int z = 0; z < 5; ++z) {
        const uint32_t k0 = __funnelshift_r(w[z], w[z + 1], sh);               // kmer[4z .. 4z+4)
        const uint32_t k1 = __funnelshift_r(w[z], w[z + 1], sh + 8);           // kmer[4z+1 .. 4z+5)  (sh + 8 <= 32)
        const uint32_t k1b = sh == 24 ? w[z + 1] : k1;                         // shift 32 = the next word itself
        const uint32_t t0 = tw[z];
        const uint32_t t1 = __funnelshift_r(tw[z], tw[z + 1], 8);              // tsd[4z+1 .. 4z+5)
        M |= diff4(t0, k0) << (4 * z);
        M1 |= diff4(t0, k1b) << (4 * z);
        M2 |= diff4(t1, k0) << (4 * z);
    }
}
__device__ __forceinline__ bool near1_masks(uint32_t M, uint32_t M1, uint32_t M2, int k) {
    const uint32_t ones = (1u << k) - 1u;
    if (__popc(M & ones) <= 1) return true;
    const int a0 = min(__ffs((int)(M1 | (1u << 20))) - 1, k - 1);
    if (((M >> (a0 + 1)) & ((1u << (k - 1 - a0)) - 1u)) == 0) return true;
    const int a1 = min(__ffs((int)(M | (1u << 20))) - 1, k - 1);
    return ((M2 >> a1) & ((1u << (k - 1 - a1)) - 1u)) == 0;
}
__device__ __forceinline__ int nonltr_tsd_search(const uint8_t* sub, int sl, const uint8_t* tsd, int avail, int lane) {
    if (avail < 8 || sl < 8) return -1;
    const uint32_t* subw = reinterpret_cast<const uint32_t*>(sub);
    const uint32_t* tsw = reinterpret_cast<const uint32_t*>(tsd);
    uint32_t tw[6];
#pragma unroll
    for (int z = 0; z < 6; ++z) tw[z] = tsw[z];
    // positions 0 .. sl - 8 <= 42: two rounds of 32 lanes
    uint32_t Ma, M1a, M2a, Mb, M1b, M2b;
    diff_masks(tw, subw, min(lane, 43), Ma, M1a, M2a);
    diff_masks(tw, subw, min(32 + lane, 43), Mb, M1b, M2b);
    for (int k = min(20, avail); k >= 8; --k) {
        uint32_t bal = __ballot_sync(HB_FULL, lane + k <= sl && near1_masks(Ma, M1a, M2a, k));
        if (bal) return __ffs((int)bal) - 1 + k;
        if (32 + k <= sl) {
            bal = __ballot_sync(HB_FULL, 32 + lane + k <= sl && near1_masks(Mb, M1b, M2b, k));
            if (bal) return 32 + __ffs((int)bal) - 1 + k;
        }
    }
    return -1;
}

// ------------------------------------------------------------------------------------------
// kernel B'': judge_boundary_v9 (non-LTR, Util.py:9483-9720), one warp per candidate.
// NW = warps per candidate.  The per-copy 3'-end / TSD scan is one serial pass per row, one row per lane: with NW = 4 a
// candidate's <= 100 rows run as four chunks of 32 side by side (one CTA per candidate); with NW = 1 (batches whose matrices
// average <= 48 rows: extra warps would only hold warp slots) a CTA judges four candidates, one warp each.
template <int NW>
__global__ void __launch_bounds__(ROW_WARPS * 32)
hb_judge_v9_tail_kernel(const HbVoteTask* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ clean,
                        uint32_t* __restrict__ planes, Match* __restrict__ matches, uint8_t* __restrict__ cons_buf,
                        HbVoteCall* __restrict__ calls) {
    // phases a, b, c are the TIR judge's kernels (thr_many 0.8, 'fl1' copy number = rows, Util.py:9490-9597)
    constexpr int TPB = ROW_WARPS / NW;                  // candidates per CTA
    __shared__ unsigned short s_rows[TPB][104];
    __shared__ int s_part[ROW_WARPS][3];
    __shared__ int s_ends[TPB][128][2];                  // (end_5, end_3) of every TSD-bearing row
    __shared__ __align__(16) uint8_t s_sub[ROW_WARPS][32][NL_SUB];     // rows are read as 32-bit words
    __shared__ __align__(16) uint8_t s_tsd[ROW_WARPS][32][NL_TSD];
    __shared__ uint8_t s_tile[ROW_WARPS][32 * NL_TSTRIDE];
    __shared__ uint8_t s_ring[ROW_WARPS][32 * NL_RSTRIDE];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = wid / NW, sub = wid % NW;
    const int ti = blockIdx.x * TPB + slot;
    if (ti >= n_tasks) return;
    const HbVoteTask t = tasks[ti];
    const JState* st = jstate_of(matches, t);
    if (!st->go) return;
    HbVoteCall call = calls[ti];
    const int R = t.R, C = t.C;
    const uint8_t* m = clean + t.clean_off;
    unsigned short* rows = s_rows[slot];
    Front F;
    front_load(F, st, planes, t);
    if (sub == 0) for (int k = lane; k < F.row_num; k += 32) rows[k] = st->rows[k];
    __syncwarp();
    const int Cc = F.Cc, row_num = F.row_num, he = F.he;
    int hs = F.hs;
    // ---- per-copy 3' end + TSD (:9604-9664), all R rows: serial row pass one row per lane, then the TSD search of every
    // row that needs one by the whole warp ------------------------------------------------------------------------
    int tsd_count = 0;
    int first_nonempty = 0x7fffffff, last_found = -1;    // row indices
    for (int r0 = 32 * sub; r0 < R; r0 += 32 * NW) {
        const int r = r0 + lane;
        bool found = false;
        int e5 = -1, e3 = -1, sl = 0, avail = 0;
        const bool search = nonltr_rows_scan(m, C, Cc, r0, R, F.hs, he, lane, s_tile[wid], s_ring[wid], &e5, &e3, &sl, &avail,
                                             s_sub[wid][lane], s_tsd[wid][lane]);
        __syncwarp();
        uint32_t todo = __ballot_sync(HB_FULL, search);
        while (todo) {
            const int src = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const int sl_s = __shfl_sync(HB_FULL, sl, src), av_s = __shfl_sync(HB_FULL, avail, src);
            const int end = nonltr_tsd_search(s_sub[wid][src], sl_s, s_tsd[wid][src], av_s, lane);
            if (lane == src && end >= 0) { found = true; e5 = max(0, e5 - 50) + end; }   // left_pos + i + k (:9654)
        }
        const bool nonempty = found && (max(e5, e3) > min(e5, e3));
        if (found && r < 128) { s_ends[slot][r][0] = e5; s_ends[slot][r][1] = e3; }
        const uint32_t bf = __ballot_sync(HB_FULL, found), bn = __ballot_sync(HB_FULL, nonempty);
        tsd_count += __popc(bf);
        if (bf) last_found = r0 + (31 - __clz(bf));
        if (bn && first_nonempty == 0x7fffffff) first_nonempty = r0 + __ffs(bn) - 1;
    }
    if (NW > 1) {                                        // TPB == 1: the CTA is this candidate's warps
        if (lane == 0) { s_part[wid][0] = tsd_count; s_part[wid][1] = first_nonempty; s_part[wid][2] = last_found; }
        __syncthreads();
        if (sub != 0) return;
        tsd_count = 0; first_nonempty = 0x7fffffff; last_found = -1;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            tsd_count += s_part[k][0]; first_nonempty = min(first_nonempty, s_part[k][1]); last_found = max(last_found, s_part[k][2]);
        }
    } else
        __syncwarp();
    // homo_boundary_start follows every TSD-bearing row until the first with a non-empty element (:9660-9664)
    const int src_row = (first_nonempty != 0x7fffffff) ? first_nonempty : last_found;
    if (src_row >= 0) {
        const uint8_t* row = m + (size_t)src_row * C;
        int e5, e3;
        if (src_row < 128) { e5 = s_ends[slot][src_row][0]; e3 = s_ends[slot][src_row][1]; }
        else {                                           // beyond the stored rows (only via hb_vote_matrices): redo that row
            int sl = 0, avail = 0;
            int srch = nonltr_rows_scan(m, C, Cc, src_row, src_row + 1, F.hs, he, lane, s_tile[wid], s_ring[wid], &e5, &e3, &sl, &avail,
                                        s_sub[wid][lane], s_tsd[wid][lane]) ? 1 : 0;
            __syncwarp();
            srch = __shfl_sync(HB_FULL, srch, 0); sl = __shfl_sync(HB_FULL, sl, 0); avail = __shfl_sync(HB_FULL, avail, 0);
            e5 = __shfl_sync(HB_FULL, e5, 0); e3 = __shfl_sync(HB_FULL, e3, 0);
            if (srch) { const int end = nonltr_tsd_search(s_sub[wid][0], sl, s_tsd[wid][0], avail, lane); if (end >= 0) e5 = max(0, e5 - 50) + end; }
        }
        const int fbs = min(e5, e3);
        // column of ungapped index fbs
        const int chunk = (Cc + 31) >> 5;
        const int lo = min(Cc, lane * chunk), hi = min(Cc, lo + chunk);
        int c_ng = 0;
        for (int c = lo; c < hi; ++c) c_ng += (row[c] != HB_SYM_GAP);
        int ntot;
        const int ub = warp_excl_scan(c_ng, lane, &ntot);
        if (fbs >= 0 && fbs < ntot) hs = column_of_ungapped(row, Cc, fbs, ub, lane);
    }
    call.homo_start = hs;
    call.trim_lead = tsd_count;
    int ml = 0;
    uint8_t* cons = cons_buf + t.cons_off;
    if (tsd_count >= 5 || 2 * tsd_count > row_num)                               // :9668
        ml = vote_consensus(F.P, m, C, rows, hs, he, 1, cons, lane);
    call.cons_len = ml;
    call.is_te = !(ml == 0 || ml < 80);                                          // :9713-9716
    if (lane == 0) calls[ti] = call;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host launchers (called from hb_api.cu)
extern "C" size_t hb_vote_match_entry_size() { return sizeof(Match); }
// extra Match entries per task (after its MATCH_LISTS * (C + 2) match slots) that hold the judge's phase state
extern "C" size_t hb_vote_state_entries() { return (sizeof(JState) + sizeof(Match) - 1) / sizeof(Match); }
// Match entries of one task with C columns: its match lists + the state
extern "C" size_t hb_vote_match_entries(int C) { return (size_t)MATCH_LISTS * ((size_t)C + 2) + hb_vote_state_entries(); }

extern "C" cudaError_t hb_vote_launch(const HbVoteTask* d_tasks, int n, const uint8_t* d_raw, uint8_t* d_clean,
                                      uint32_t* d_planes, void* d_matches, int* d_anchors, const uint8_t* d_cands,
                                      uint8_t* d_cons, HbVoteCall* d_calls, int te_type, int avg_rows, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    static_assert(ROW_WARPS == WARPS_PER_CTA, "the NW = 1 tail kernel shares the one-warp-per-candidate grid");
    // d_raw == nullptr: the clean matrices (and calls[].clean_cols) were already produced by hb_msa.cu::msa_clean_kernel
    if (d_raw) hb_clean_kernel<<<n, CLEAN_THREADS, 0, stream>>>(d_tasks, d_raw, d_clean, d_calls);
    const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int prebuilt = d_raw == nullptr;               // ... and so were the bit-planes over all rows
    if (te_type == 0) {
        cudaMemsetAsync(d_anchors, 0, sizeof(int), stream);                    // (the anchor scratch of the Helitron judge: its work queue here)
        hb_judge_a_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_clean, (Match*)d_matches, d_cands, d_calls, d_anchors);
        hb_judge_a2_kernel<<<min(n, 148 * 4), A2_WARPS * 32, 0, stream>>>(d_tasks, d_anchors, d_clean, (Match*)d_matches, d_cands, d_calls);
        hb_judge_b_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches, d_calls, 0, prebuilt);
        hb_judge_c_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_planes, (Match*)d_matches, d_calls, 0.7);
        hb_judge_d_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches, d_cons,
                                                                   d_calls);
    } else if (te_type == 1) {
        hb_judge_v6_a_kernel<<<n, ROW_WARPS * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches, d_anchors, d_cands,
                                                               d_calls);
        hb_judge_v6_b_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches, d_calls, prebuilt);
        hb_judge_v6_c_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches, d_cons,
                                                                      d_calls, prebuilt);
    } else if (te_type == 2) {
        cudaMemsetAsync(d_anchors, 0, sizeof(int), stream);                    // (the anchor scratch of the Helitron judge: its work queue here)
        hb_judge_a_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_clean, (Match*)d_matches, d_cands, d_calls, d_anchors);
        hb_judge_a2_kernel<<<min(n, 148 * 4), A2_WARPS * 32, 0, stream>>>(d_tasks, d_anchors, d_clean, (Match*)d_matches, d_cands, d_calls);
        hb_judge_b_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches, d_calls, 1, prebuilt);
        hb_judge_c_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(d_tasks, n, d_planes, (Match*)d_matches, d_calls, 0.8);
        if (avg_rows > 48)
            hb_judge_v9_tail_kernel<ROW_WARPS><<<n, ROW_WARPS * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches,
                                                                                 d_cons, d_calls);
        else
            hb_judge_v9_tail_kernel<1><<<grid, ROW_WARPS * 32, 0, stream>>>(d_tasks, n, d_clean, d_planes, (Match*)d_matches,
                                                                            d_cons, d_calls);
    } else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}
