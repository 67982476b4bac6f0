// hite_b200 -- TSD-bounded variant generation, the step right before the boundary-adjustment path in
// HiTE's TIR module.  One warp per flanked candidate.
//
// Replaces (reference module/Util.py):
//   search_confident_tir_v4   :7734-7845   k-mer dictionary of the +-d window around the left boundary, matched
//                                          against the k-mers around the right boundary, k in {2,3,4,5,6,8,9,10,11}
//   get_short_tir_contigs     :7297-7334   (classification of each emitted variant)
// Bit-identical to oracle/hbo_tsd.c, which is pinned on outputs of the imported reference
// (tests/golden/tsd_v4.json.gz).  Like the oracle, equal-distance entries are ordered by
// (distance, start, end, k): the reference's own order among them is Python set iteration order.
#include "hb_common.cuh"
#include "hb_tsd.cuh"

namespace {

// k runs over TIR_TSDs = {2, 3, 4, 5, 6, 8, 9, 10, 11} (Util.py:7760)
struct Win {             // 11-byte look-ahead of every window position: bytes 0-7 in lo, 8-10 in hi
    unsigned long long* lo;
    uint32_t* hi;
};
__device__ __forceinline__ uint8_t rc_base(uint8_t c) {     // getReverseSequence, Util.py:1635
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'N';
}
__device__ __forceinline__ bool eq_lit(const uint8_t* s, const char* lit, int n) {
    for (int i = 0; i < n; ++i) if (s[i] != (uint8_t)lit[i]) return false;
    return true;
}
// get_short_tir_contigs (Util.py:7297-7334): 0 none, 1 hAT, 2 Mutator, 3 CACTA, 4 CCC
__device__ int short_tir_class(const uint8_t* tir, int len, int tsd_len, int plant) {
    if (len < 5) return 0;
    bool eq5 = true, eq3 = true;
    for (int i = 0; i < 5; ++i) if (tir[i] != rc_base(tir[len - 1 - i])) eq5 = false;
    for (int i = 0; i < 3; ++i) if (tir[i] != rc_base(tir[len - 1 - i])) eq3 = false;
    if (eq5) {
        if (tsd_len == 8 && len < 4000) return 1;
        if (tsd_len >= 9 && tsd_len <= 11) return 2;
        if (plant == 1 && tsd_len == 3 && (eq_lit(tir, "CACTA", 5) || eq_lit(tir, "CACTG", 5))) return 3;
        return 0;
    }
    if (eq3 && eq_lit(tir, "CCC", 3)) return 4;
    return 0;
}

// key of a (dist, start, end, k) tuple: the canonical order is the numeric order of the key
__device__ __forceinline__ unsigned long long tup_key(int dist, int start, int end, int k) {
    return ((unsigned long long)dist << 44) | ((unsigned long long)start << 24) | ((unsigned long long)end << 4) | (unsigned long long)k;
}

__global__ void __launch_bounds__(32)
tsd_kernel(const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ seq_off, const int32_t* __restrict__ raw_start1,
           const int32_t* __restrict__ raw_end1, uint32_t n_cand, int d, int plant, int wcap, hb_tsd_rec* __restrict__ out,
           uint32_t* __restrict__ n_out) {
    extern __shared__ unsigned long long smem[];
    const uint32_t c = blockIdx.x;
    const int lane = threadIdx.x;
    if (c >= n_cand) return;
    // shared layout: Llo[wcap] Rlo[wcap] keys[9 wcap] | Lhi[wcap] Rhi[wcap] | counter | order[9 wcap] (16-bit)
    Win L, R;
    L.lo = smem; R.lo = smem + wcap;
    unsigned long long* keys = smem + 2 * wcap;
    L.hi = reinterpret_cast<uint32_t*>(keys + 9 * wcap); R.hi = L.hi + wcap;
    int* counter = reinterpret_cast<int*>(R.hi + wcap);
    uint16_t* order = reinterpret_cast<uint16_t*>(counter + 4);
    const uint8_t* seq = seqs + seq_off[c];
    const int n = (int)(seq_off[c + 1] - seq_off[c]);
    const int rs = raw_start1[c] - 1, re = raw_end1[c] - 1;                     // :7736-7737
    int ls = max(0, rs - d), le = min(n, rs + d + 1);                            // :7743-7747
    int rts = max(0, re - d), rte = min(n, re + d + 1);                          // :7750-7754
    ls = min(ls, n); rts = min(rts, n); le = max(le, ls); rte = max(rte, rts);
    const int nl = le - ls, nr = rte - rts;
    for (int i = lane; i < nl; i += 32) {
        unsigned long long lo = 0; uint32_t hi = 0;
        for (int z = 0; z < 11 && ls + i + z < le; ++z) {
            const unsigned long long b = seq[ls + i + z];
            if (z < 8) lo |= b << (8 * z); else hi |= (uint32_t)b << (8 * (z - 8));
        }
        L.lo[i] = lo; L.hi[i] = hi;
    }
    for (int j = lane; j < nr; j += 32) {
        unsigned long long lo = 0; uint32_t hi = 0;
        for (int z = 0; z < 11 && rts + j + z < rte; ++z) {
            const unsigned long long b = seq[rts + j + z];
            if (z < 8) lo |= b << (8 * z); else hi |= (uint32_t)b << (8 * (z - 8));
        }
        R.lo[j] = lo; R.hi[j] = hi;
    }
    if (lane == 0) *counter = 0;
    __syncwarp();
    // ---- tuples: every right k-mer occurrence that sets a new "closest to the raw end" for its string ----
    // One pass over the (left i, right j) pairs serves all nine k: the pair's common-prefix length m says for which k
    // the two k-mers are equal.  Lanes own right positions; i is warp-uniform, so L is a shared-memory broadcast.
    for (int j = lane; j < nr; j += 32) {
        const int pos = rts + j - 1;                                            // :7783
        if (pos < 0 || pos > n - 1) continue;
        const unsigned long long rlo = R.lo[j];
        const uint32_t rhi = R.hi[j];
        const int kmax_r = nr - j;                                              // k-mer must end inside the right window
        int best[9];                                                            // left_pos per k: closest to rs, first on ties
#pragma unroll
        for (int ki = 0; ki < 9; ++ki) best[ki] = -1;
        const int i_end = min(nl, n - 1 - ls) - 2;                              // i + k <= nl and ls + i + k <= n - 1 for k = 2
        for (int i = 0; i <= i_end; ++i) {
            const unsigned long long x = L.lo[i] ^ rlo;
            if ((x & 0xFFFFull) != 0) continue;                                 // first two bytes differ: no k matches
            int m = x ? ((__ffsll((long long)x) - 1) >> 3) : 8;
            if (m == 8) { const uint32_t y = (L.hi[i] ^ rhi) & 0xFFFFFFu; m = 8 + (y ? ((__ffs((int)y) - 1) >> 3) : 3); }
            m = min(m, min(kmax_r, min(nl - i, n - 1 - ls - i)));
#pragma unroll
            for (int ki = 0; ki < 9; ++ki) {
                const int k = ki < 5 ? ki + 2 : ki + 3;                         // 2 3 4 5 6 8 9 10 11
                if (k <= m) {
                    const int lp = ls + i + k;
                    if (best[ki] < 0 || abs(lp - rs) < abs(best[ki] - rs)) best[ki] = lp;
                }
            }
        }
        // right_pos after this occurrence = running closest to re over the occurrences so far (:7788-7796): this
        // occurrence contributes a NEW tuple only when no earlier equal k-mer is at least as close
        const int my = abs(pos - re);
        int not_best = 0;                                                       // smallest m-bound: k <= not_best are beaten
        for (int j2 = max(0, 1 - rts); j2 < j; ++j2) {
            const unsigned long long x = R.lo[j2] ^ rlo;
            if ((x & 0xFFFFull) != 0) continue;
            if (abs(rts + j2 - 1 - re) > my) continue;
            int m = x ? ((__ffsll((long long)x) - 1) >> 3) : 8;
            if (m == 8) { const uint32_t y = (R.hi[j2] ^ rhi) & 0xFFFFFFu; m = 8 + (y ? ((__ffs((int)y) - 1) >> 3) : 3); }
            not_best = max(not_best, m);                                        // j2 < j: its k-mers end inside the window too
        }
        const uint8_t* km = seq + rts + j;
#pragma unroll
        for (int ki = 0; ki < 9; ++ki) {
            const int k = ki < 5 ? ki + 2 : ki + 3;
            const int best_l = best[ki];
            if (best_l < 0 || k <= not_best) continue;
            bool ok;                                                            // :7802-7805
            if (k != 2 && k != 4) ok = true;
            else if (k == 4) ok = eq_lit(km, "TTAA", 4);
            else {
                ok = eq_lit(km, "TA", 2);
                if (!ok && plant == 0)
                    ok = best_l + 3 <= n && eq_lit(seq + best_l, "CCC", 3) && pos >= 2 && pos + 1 <= n && eq_lit(seq + pos - 2, "GGG", 3);
            }
            if (!ok) continue;
            const int slot = atomicAdd(counter, 1);
            keys[slot] = tup_key(abs(best_l - rs) + my, best_l, pos, k);
        }
    }
    __syncwarp();
    const int nt = *counter;
    // ---- canonical order: rank by counting (keys are distinct) ------------------------------------------------
    for (int t = lane; t < nt; t += 32) {
        const unsigned long long kt = keys[t];
        int rank = 0;
        for (int u = 0; u < nt; ++u) rank += keys[u] < kt ? 1 : 0;
        order[rank] = (uint16_t)t;
    }
    __syncwarp();
    // ---- filters (:7811-7833) and the 100-record cut (:7836-7844), in sorted order ----------------------------
    int emitted = 0;
    for (int t0 = 0; t0 < nt && emitted < 100; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        int k = 0, start = 0, end = 0, dist = 0;
        if (t < nt) {
            const unsigned long long kt = keys[order[t]];
            k = (int)(kt & 15ull); end = (int)((kt >> 4) & 0xFFFFFull); start = (int)((kt >> 24) & 0xFFFFFull); dist = (int)(kt >> 44);
            const uint8_t* km = seq + end + 1;
            bool has_nn = false;
            for (int z = 0; z + 1 < k; ++z) if (km[z] == 'N' && km[z + 1] == 'N') has_nn = true;
            const int len = end >= start ? end - start + 1 : 0;
            keep = !has_nn && len >= 100;
            if (keep) {
                const uint8_t* tir = seq + start;
                if (tir[0] == 'T' && tir[1] == 'G' && tir[len - 2] == 'C' && tir[len - 1] == 'A') keep = false;
                if (eq_lit(tir, "TATATATA", 8) || eq_lit(tir, "ATATATAT", 8)) keep = false;
            }
        }
        const uint32_t bal = __ballot_sync(HB_FULL, keep);
        const int slot = emitted + __popc(bal & ((1u << lane) - 1u));
        if (keep && slot < 100) {
            hb_tsd_rec r;
            r.k = k; r.start = start; r.end = end; r.dist = dist; r.idx = t;
            r.short_class = short_tir_class(seq + start, end - start + 1, k, plant);
            out[(size_t)c * 100 + slot] = r;
        }
        emitted += __popc(bal);
    }
    if (lane == 0) n_out[c] = (uint32_t)min(emitted, 100);
}

// ------------------------------------------------------------------------------------------------
// search_polyA_TSD (module/Util.py:10915-11006): the per-candidate scan of the non-LTR candidate generator
// (get_candidate_non_LTR :11008, judge_Non_LTR_transposons.py:34).  One warp per flanked candidate; the four +-25 bp window
// scans are warp-uniform (every lane walks the same 50 bytes), the TSD search spreads the window positions over the lanes
// and takes the first hit of the longest k by ballot.  Bit-identical to oracle/hbo_tsd.c::hbo_search_polyA_TSD.
__device__ __forceinline__ int py_clip(int x, int n) { return x < 0 ? 0 : (x > n ? n : x); }

// find_nearest_polyA / polyT (:10865, :10903) -> start, end, len(sequence[start:end])
__device__ void nearest_run(const uint8_t* seq, int n, int position, uint8_t base, int& ps, int& pe, int& plen) {
    const int lo = py_clip(max(0, position - 25), n), hi = py_clip(position + 25, n);
    int max_len = 0, cur = 0, start = 0, ms = -1, me = -1;
    for (int i = lo; i < hi; ++i) {
        if (seq[i] == base) { if (++cur == 1) start = i - lo; }
        else { if (cur >= 6 && cur > max_len) { max_len = cur; ms = start; me = i - lo; } cur = 0; }
    }
    if (cur >= 6 && cur > max_len) { ms = start; me = hi - lo; }
    if (ms == -1 || me == -1) { ps = -1; pe = -1; plen = 0; return; }
    ms = max(0, position - 25 + ms);                           // the offset ignores the clipping of the window (:10872)
    me = max(0, position - 25 + me);
    ps = ms; pe = me;
    plen = max(0, py_clip(me, n) - py_clip(ms, n));
}
// find_nearest_tandem (:9772-9819): four exact copies of a 2..6-mer; longest motif wins, first position
__device__ void nearest_tandem(const uint8_t* seq, int n, int pos, int& ps, int& pe, int& plen) {
    const int start = max(0, pos - 25), end = min(n, pos + 25);
    ps = -1; pe = -1; plen = 0;
    for (int m = 2; m <= 6; ++m)
        for (int i = start; i <= end - 4 * m; ++i) {
            bool ok = true;
            for (int z = m; z < 4 * m && ok; ++z) ok = seq[i + z] == seq[i + z % m];
            if (ok && 4 * m > plen) { plen = 4 * m; ps = i; pe = i + 4 * m; }
        }
}
// some substring of kmer[0..k) within Levenshtein distance 1 of tsd[0..k): one substitution, or tsd minus one character
// equal to the k-mer without its last / first character
__device__ bool near1(const uint8_t* tsd, const uint8_t* kmer, int k) {
    int ham = 0;
    for (int i = 0; i < k; ++i) ham += tsd[i] != kmer[i] ? 1 : 0;
    if (ham <= 1) return true;
    for (int off = 0; off <= 1; ++off) {
        const uint8_t* x = kmer + off;
        int lcp = 0, lcs = 0;
        while (lcp < k - 1 && tsd[lcp] == x[lcp]) ++lcp;
        while (lcs < k - 1 && tsd[k - 1 - lcs] == x[k - 2 - lcs]) ++lcs;
        if (lcp + lcs >= k - 1) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128)
polya_tsd_kernel(const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ seq_off, uint32_t n_cand, int flank,
                 hb_polya_rec* __restrict__ out) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n_cand) return;
    const uint8_t* seq = seqs + seq_off[c];
    const int n = (int)(seq_off[c + 1] - seq_off[c]);
    const int raw_start = flank + 1, raw_end = n - flank;                           // :10917-10918
    int end_3 = -1, end_5 = -1, direct = 0;
    int as, ae, alen, ts, te, tlen;
    if (raw_end >= 0) {
        nearest_run(seq, n, raw_end, 'A', as, ae, alen);
        nearest_tandem(seq, n, raw_end, ts, te, tlen);
        if (min(alen, tlen) > 0) { end_3 = alen > tlen ? ae : te; end_5 = raw_start; direct = 1; }     // :10926-10934
    }
    nearest_run(seq, n, raw_start, 'T', as, ae, alen);
    nearest_tandem(seq, n, raw_start, ts, te, tlen);
    if (min(alen, tlen) > 0) { end_3 = alen > tlen ? as : ts; end_5 = raw_end; direct = 2; }           // :10940-10948
    int found = 0, tsd_k = 0;
    if (end_3 != -1 && end_5 != -1 && direct) {
        const int lo5 = max(0, end_5 - 25);
        const int lo = py_clip(lo5, n), hi = py_clip(max(0, end_5 + 25), n);
        const int sublen = max(0, hi - lo);
        for (int k = 20; k >= 8 && !found; --k) {
            const int t0 = direct == 2 ? end_3 - k : end_3;
            if (t0 < 0 || t0 + k > n) continue;                                     // len(TSD) != k
            const uint8_t* tsd = seq + t0;
            bool has_n = false;
            for (int z = 0; z < k; ++z) has_n |= tsd[z] == 'N';
            if (has_n) continue;
            for (int i0 = 0; i0 + k <= sublen && !found; i0 += 32) {
                const int i = i0 + lane;
                const bool hit = i + k <= sublen && near1(tsd, seq + lo + i, k);
                const uint32_t bal = __ballot_sync(HB_FULL, hit);
                if (bal) {
                    const int first = i0 + __ffs((int)bal) - 1;
                    end_5 = lo5 + first + (direct == 1 ? k : 0);                    // :10971 / :10992
                    found = 1; tsd_k = k;
                }
            }
        }
    }
    if (lane == 0) {
        hb_polya_rec r;
        r.found = found; r.direct = direct; r.end_5 = end_5; r.end_3 = end_3; r.tsd_k = tsd_k;
        int a = min(end_5, end_3), b = max(end_5, end_3);
        if (!direct) a = b = 0;
        r.seq_start = a < 0 ? max(0, a + n) : py_clip(a, n);                        // Python slice bounds of seq[min:max] (:11002)
        r.seq_end = b < 0 ? max(0, b + n) : py_clip(b, n);
        if (r.seq_end < r.seq_start) r.seq_end = r.seq_start;
        out[c] = r;
    }
}

}  // namespace

size_t hb_tsd_smem_bytes(int d) {
    const size_t wcap = (size_t)(2 * d + 1 + 3) & ~(size_t)3;
    return wcap * (2 + 9) * sizeof(unsigned long long) + wcap * 2 * sizeof(uint32_t) + 16 + 9 * wcap * sizeof(uint16_t);
}

cudaError_t hb_tsd_launch(const uint8_t* seqs, const uint64_t* seq_off, const int32_t* raw_start1, const int32_t* raw_end1,
                          uint32_t n_cand, int d, int plant, hb_tsd_rec* out, uint32_t* n_out, cudaStream_t s) {
    if (n_cand == 0) return cudaSuccess;
    const int wcap = (2 * d + 1 + 3) & ~3;
    tsd_kernel<<<n_cand, 32, hb_tsd_smem_bytes(d), s>>>(seqs, seq_off, raw_start1, raw_end1, n_cand, d, plant, wcap, out, n_out);
    return cudaGetLastError();
}

cudaError_t hb_polya_tsd_launch(const uint8_t* seqs, const uint64_t* seq_off, uint32_t n_cand, int flank, hb_polya_rec* out,
                                cudaStream_t s) {
    if (n_cand == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_cand * 32;
    polya_tsd_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(seqs, seq_off, n_cand, flank, out);
    return cudaGetLastError();
}
