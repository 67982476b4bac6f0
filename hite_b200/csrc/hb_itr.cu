// hite_b200 -- HB-itr: terminal-inverted-repeat matcher, the stand-in for the closed binary `tools/itrsearch -i 0.7 -l 7`
// (reference module/Util.py:216-224, run by search_confident_tir_batch_v1 :6563 / :6575 on the 40 + 40 bp ends of every
// TSD-bounded variant and by remove_no_tirs :13907 on whole low-copy elements).
//
// What the binary does was probed as a black box (it ships without source): it aligns the 5' end of a record against the
// reverse complement of its 3' end, both anchored at the record's ends (`Begins at (1,1)`), inside a window of
// min(500, len / 2) bases per end, match +10, mismatch -16, a gap of n bases -(32 + 32 n); it reports the best-scoring end
// cell and keeps the record when its "Quality" is positive -- empirically  score - p > 0  with p = 64 when both arms are
// equally long, 32 when the 5' arm is longer, 0 when it is shorter --,  m / (m + x) >= identity  and the 5' arm is >= the
// minimum length, where m / x = matches / mismatches of the alignment.  HB-itr is the exact
// optimum of that scoring (the binary's own search is a heuristic and is sometimes below the optimum): DP over the three
// affine states with (score, m, x, g) packed into one 64-bit key, so the maximum -- ties: more matches, then more
// mismatches, then more gap bases; between cells the first in (5' arm, 3' arm) order -- is a plain integer max.  N never
// matches.  Agreement with the binary on its own outputs: tests/golden/itrsearch.json.gz; bit-identical to
// oracle/hbo_tsd.c::hbo_itr_search.  One thread per record: the records are 80 bp in the hot caller (a 41 x 41 table) and
// the three rows of the table live in a per-record strip of global scratch.
#include "hb_common.cuh"
#include "hb_itr.cuh"

namespace {

constexpr long long ITR_BIAS = 1ll << 20;
constexpr long long K_ONE = 1ll << 30;                                  // one score unit in the packed key
constexpr long long K_MATCH = 10 * K_ONE + (1ll << 20), K_MIS = -16 * K_ONE + (1ll << 10);
constexpr long long K_OPEN = -64 * K_ONE + 1, K_EXT = -32 * K_ONE + 1;

__device__ __forceinline__ long long kadd(long long k, long long d) { return k ? k + d : 0; }   // key 0 = unreachable
__device__ __forceinline__ long long kmax(long long a, long long b) { return a > b ? a : b; }
__device__ __forceinline__ int code_of(uint8_t c) {
    switch (c & 0xDF) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}

__global__ void itr_kernel(const uint8_t* __restrict__ seqs, const uint64_t* __restrict__ seq_off, uint32_t n, double min_ident,
                           int min_len, int max_w, long long* __restrict__ scratch, const uint64_t* __restrict__ scratch_off,
                           hb_itr_rec* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint8_t* s = seqs + seq_off[r];
    const int len = (int)(seq_off[r + 1] - seq_off[r]);
    const int w = min(max_w, len / 2);
    long long* uD = scratch + scratch_off[r];                           // states of the row above, per column
    long long* uE = uD + (w + 1);
    long long* uF = uE + (w + 1);
    const long long z0 = ITR_BIAS * K_ONE;
    long long best = z0;                                                // the origin: an empty alignment
    int bi = 0, bj = 0;
    uD[0] = z0; uE[0] = 0; uF[0] = 0;
    for (int j = 1; j <= w; ++j) { uD[j] = 0; uF[j] = 0; uE[j] = kmax(kadd(kmax(uD[j - 1], uF[j - 1]), K_OPEN), kadd(uE[j - 1], K_EXT)); }
    for (int i = 1; i <= w; ++i) {
        const int ca = code_of(s[i - 1]);
        long long diag = kmax(kmax(uD[0], uE[0]), uF[0]);               // max(D, E, F)[i-1][j-1]
        long long lD = 0, lE = 0, lF = kmax(kadd(kmax(uD[0], uE[0]), K_OPEN), kadd(uF[0], K_EXT));     // column 0 of row i
        uD[0] = lD; uE[0] = lE; uF[0] = lF;
        for (int j = 1; j <= w; ++j) {
            const long long aD = uD[j], aE = uE[j], aF = uF[j];        // (i-1, j)
            const int cb = code_of(s[len - j]);                        // b[j-1] = complement of s[len - j]
            const bool m = ca < 4 && cb < 4 && ca == 3 - cb;
            const long long d = kadd(diag, m ? K_MATCH : K_MIS);
            const long long e = kmax(kadd(kmax(lD, lF), K_OPEN), kadd(lE, K_EXT));
            const long long f = kmax(kadd(kmax(aD, aE), K_OPEN), kadd(aF, K_EXT));
            diag = kmax(kmax(aD, aE), aF);
            const long long c = kmax(kmax(d, e), f);
            if (c > best) { best = c; bi = i; bj = j; }
            uD[j] = d; uE[j] = e; uF[j] = f;
            lD = d; lE = e; lF = f;
        }
    }
    hb_itr_rec o;
    const int sc = (int)((best >> 30) - ITR_BIAS), nm = (int)((best >> 20) & 1023), nx = (int)((best >> 10) & 1023), ng = (int)(best & 1023);
    o.len5 = bi; o.len3 = bj; o.score = sc; o.n_match = nm; o.n_mismatch = nx; o.n_gap = ng;
    const int pen = bi == bj ? 64 : (bi > bj ? 32 : 0);
    o.found = (sc - pen > 0) && (nm + nx > 0) && ((double)nm / (double)(nm + nx) >= min_ident) && bi >= min_len ? 1 : 0;
    out[r] = o;
}

}  // namespace

cudaError_t hb_itr_launch(const uint8_t* seqs, const uint64_t* seq_off, uint32_t n, double min_ident, int min_len, int max_w,
                          long long* scratch, const uint64_t* scratch_off, hb_itr_rec* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    itr_kernel<<<(n + 127) / 128, 128, 0, s>>>(seqs, seq_off, n, min_ident, min_len, max_w, scratch, scratch_off, out);
    return cudaGetLastError();
}
