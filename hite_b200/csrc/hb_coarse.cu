// hite_b200 -- coarse-boundary all-vs-all on the GPU: one self-join of the minimizer index replaces the
// n^2 `blastn -subject T -query Q` runs of process_blast_alignments (reference module/Util.py:4724-4780,
// sequence2sequenceBlastn :4068-4091).
//
// HB-hsp: every minimizer (position order, ordinal qi) is paired with the first
// occ_cap other occurrences of its hash; pair key = rel<<62 | diag<<29 | qi, diag = oriented subject
// position - query position + 2^32; keys are radix-sorted; an HSP is a maximal run with equal (rel, diag),
// consecutive query gap <= 64 bp, inside one query segment and one subject segment, spanning >= 50 bp.
// Gapped alignments appear as several exact-diagonal HSPs -- stitching them is what FMEA is for.
#include "hb_common.cuh"
#include "hb_coarse.cuh"

namespace {

constexpr int HSP_MAXGAP = 64;
constexpr int HSP_MINSPAN = 50;

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t* __restrict__ a, uint64_t lo, uint64_t hi, uint64_t key) {
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int seg_of(const uint64_t* __restrict__ seg_goff, int n_seg, uint64_t pos) {
    int lo = 0, hi = n_seg;       // last segment with goff <= pos
    while (lo < hi) { int mid = (lo + hi) >> 1; if (seg_goff[mid] <= pos) lo = mid + 1; else hi = mid; }
    return lo - 1;
}

// one thread per minimizer (position order)
// [qpos_lo, qpos_hi): the query positions this call (one part of a sharded self-join) owns; the others pair with nothing
__global__ void pair_count_kernel(const uint64_t* __restrict__ mm, uint32_t n_mm, const uint64_t* __restrict__ idx,
                                  const uint32_t* __restrict__ bucket, int occ_cap, uint64_t qpos_lo, uint64_t qpos_hi,
                                  uint32_t* __restrict__ lo_out, uint32_t* __restrict__ cnt_out) {
    const uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= n_mm) return;
    const uint64_t e = mm[qi];
    const uint64_t qpos = (e >> 1) & 0xFFFFFFFFull;
    if (qpos < qpos_lo || qpos >= qpos_hi) { lo_out[qi] = 0; cnt_out[qi] = 0; return; }
    const uint32_t h = (uint32_t)(e >> 33);
    const uint32_t bk = h >> (30 - HB_COARSE_BUCKET_BITS);
    const uint64_t l = lower_bound_u64(idx, bucket[bk], bucket[bk + 1], (uint64_t)h << 33);
    const uint64_t r = lower_bound_u64(idx, l, bucket[bk + 1], ((uint64_t)h + 1) << 33);
    const uint32_t win = (uint32_t)min((uint64_t)occ_cap + 1, r - l);      // window of the run that is looked at
    // the minimizer itself is skipped if it falls inside the window
    uint32_t self_in = 0;
    const uint64_t me = lower_bound_u64(idx, l, r, e);                      // exact entry (keys are unique)
    if (me - l < win) self_in = 1;
    uint32_t cnt = win - self_in;
    if (cnt > (uint32_t)occ_cap) cnt = (uint32_t)occ_cap;
    lo_out[qi] = (uint32_t)l;
    cnt_out[qi] = cnt;
}

// 8 lanes per minimizer
__global__ void pair_fill_kernel(const uint64_t* __restrict__ mm, uint32_t n_mm, const uint64_t* __restrict__ idx,
                                 uint64_t n_idx, uint64_t tot, int k, int occ_cap, const uint32_t* __restrict__ lo_in,
                                 const uint32_t* __restrict__ cnt_in, const uint32_t* __restrict__ off_in,
                                 uint64_t* __restrict__ keys) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t qi = (uint32_t)(t >> 3), sub = (uint32_t)(t & 7);
    if (qi >= n_mm) return;
    const uint32_t cnt = cnt_in[qi];
    if (cnt == 0) return;
    const uint64_t e = mm[qi];
    const int64_t qpos = (int64_t)((e >> 1) & 0xFFFFFFFFull);
    const uint32_t qs = (uint32_t)(e & 1);
    const uint64_t lo = lo_in[qi];
    const uint64_t base = off_in[qi];
    // window entries in index order, the own entry skipped: output slot x reads entry x or x+1
    const uint64_t me = lower_bound_u64(idx, lo, min(n_idx, lo + (uint64_t)occ_cap + 1), e) - lo;   // >= window size if outside
    for (uint32_t x = sub; x < cnt; x += 8) {
        const uint64_t g = idx[lo + x + (x >= me ? 1 : 0)];
        const int64_t gpos = (int64_t)((g >> 1) & 0xFFFFFFFFull);
        const uint32_t rel = qs ^ (uint32_t)(g & 1);
        const int64_t d = (rel ? ((int64_t)tot - gpos - k) : gpos) - qpos + (1ll << 32);
        keys[base + x] = ((uint64_t)rel << 62) | ((uint64_t)d << 29) | (uint64_t)qi;
    }
}

// run starts on the sorted keys
__global__ void hsp_flag_kernel(const uint64_t* __restrict__ keys, uint64_t n, const uint64_t* __restrict__ mm,
                                uint64_t tot, int k, const uint64_t* __restrict__ seg_goff, int n_seg,
                                uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t f = 1;
    if (i > 0) {
        const uint64_t a = keys[i - 1], b = keys[i];
        if ((a >> 29) == (b >> 29)) {                      // same strand and diagonal
            const int64_t qa = (int64_t)((mm[a & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
            const int64_t qb = (int64_t)((mm[b & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
            if (qb - qa <= HSP_MAXGAP) {
                const int rel = (int)(b >> 62) & 1;
                const int64_t d = (int64_t)((b >> 29) & 0x1FFFFFFFFull) - (1ll << 32);
                // forward subject positions of the two seeds
                const int64_t sa = rel ? (int64_t)tot - (qa + d) - k : qa + d;
                const int64_t sb = rel ? (int64_t)tot - (qb + d) - k : qb + d;
                if (seg_of(seg_goff, n_seg, (uint64_t)qa) == seg_of(seg_goff, n_seg, (uint64_t)qb) &&
                    seg_of(seg_goff, n_seg, (uint64_t)sa) == seg_of(seg_goff, n_seg, (uint64_t)sb))
                    f = 0;
            }
        }
    }
    flag[i] = f;
}
__global__ void hsp_start_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rid, uint64_t n,
                                 uint32_t* __restrict__ run_start) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) run_start[rid[i]] = (uint32_t)i;
}
// ---- S11b: base-level HSP ends (see oracle/hbo_hsp.c for the definition this follows bit for bit) ----------------------
// A seed run is grown along its diagonal over 8-base blocks aligned in query coordinates; a block is good when >= 5 of its
// 8 positions match.  Blocks are compared 8 bases at a time: 16 bits of 2-bit codes + 8 N-mask bits per side.
constexpr int HSP_MAXEXT = 4096;

struct Diag {
    const uint32_t* seq2; const uint32_t* nmask;
    int rel; int64_t d, qlo, qhi, slo, shi, tot;
};
__device__ __forceinline__ void fwd8(const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask, uint64_t p,
                                     uint32_t* codes, uint32_t* nm) {
    const uint64_t w = p >> 4, wn = p >> 5;
    const uint64_t v = (uint64_t)seq2[w] | ((uint64_t)seq2[w + 1] << 32);          // the buffers carry 64 B of slack
    *codes = (uint32_t)(v >> (2 * (p & 15))) & 0xFFFFu;
    const uint64_t m = (uint64_t)nmask[wn] | ((uint64_t)nmask[wn + 1] << 32);
    *nm = (uint32_t)(m >> (p & 31)) & 0xFFu;
}
__device__ __forceinline__ int base_code(const Diag& D, int64_t p) {
    if ((D.nmask[p >> 5] >> (p & 31)) & 1u) return 4;
    return (int)((D.seq2[p >> 4] >> (2 * (p & 15))) & 3u);
}
// match mask of block b: bit i = query position 8b + i matches its subject base
__device__ __forceinline__ uint32_t block_mask(const Diag& D, int64_t b) {
    const int64_t q = 8 * b, j = q + D.d;
    const int64_t f0 = D.rel ? D.tot - 1 - (j + 7) : j;                            // lowest forward subject position of the block
    if (q >= D.qlo && q + 8 <= D.qhi && f0 >= D.slo && f0 + 8 <= D.shi) {
        uint32_t cq, nq, cs, ns;
        fwd8(D.seq2, D.nmask, (uint64_t)q, &cq, &nq);
        fwd8(D.seq2, D.nmask, (uint64_t)f0, &cs, &ns);
        if (D.rel) {                                                               // reverse the 8 bases and complement them
            uint32_t r = __brev(cs) >> 16;
            cs = (((r & 0xAAAAu) >> 1) | ((r & 0x5555u) << 1)) ^ 0xFFFFu;
            ns = __brev(ns) >> 24;
        }
        uint32_t x = cq ^ cs;
        x = (x | (x >> 1)) & 0x5555u;
        x = (x | (x >> 1)) & 0x3333u;
        x = (x | (x >> 2)) & 0x0F0Fu;
        x = (x | (x >> 4)) & 0x00FFu;
        return ~(x | nq | ns) & 0xFFu;
    }
    uint32_t m = 0;                                                                // a block cut by a segment boundary
    for (int i = 0; i < 8; ++i) {
        const int64_t qq = q + i, jj = qq + D.d, f = D.rel ? D.tot - 1 - jj : jj;
        if (qq < D.qlo || qq >= D.qhi || f < D.slo || f >= D.shi) continue;
        const int a = base_code(D, qq);
        int c = base_code(D, f);
        if (a > 3 || c > 3) continue;
        if (D.rel) c = 3 - c;
        m |= (uint32_t)(a == c) << i;
    }
    return m;
}
__device__ __forceinline__ void seed_segment(const Diag& D, int64_t q0, uint32_t* qa, uint32_t* qb) {
    const int64_t bc = (q0 + 7) >> 3;
    int64_t bl = bc, br = bc;
    const uint32_t mc = block_mask(D, bc);                    // all ones unless a segment boundary cuts the seed
    if (mc == 0) { *qa = *qb = (uint32_t)q0; return; }
    uint32_t ml = mc, mr = mc, m;                             // masks of the last good block on either side
    for (int n = 0; n < HSP_MAXEXT; ++n) { m = block_mask(D, bl - 1); if (__popc(m) < 5) break; ml = m; --bl; }
    m = block_mask(D, bl - 1);
    {
        int sc = 0, best = 0, ext = 0;
#pragma unroll
        for (int i = 7; i >= 0; --i) { sc += ((m >> i) & 1u) ? 1 : -2; if (sc > best) { best = sc; ext = 8 - i; } }
        *qa = (uint32_t)(ext ? 8 * bl - ext : 8 * bl + (__ffs((int)ml) - 1));
    }
    for (int n = 0; n < HSP_MAXEXT; ++n) { m = block_mask(D, br + 1); if (__popc(m) < 5) break; mr = m; ++br; }
    m = block_mask(D, br + 1);
    {
        int sc = 0, best = 0, ext = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc += ((m >> i) & 1u) ? 1 : -2; if (sc > best) { best = sc; ext = i + 1; } }
        *qb = (uint32_t)(ext ? 8 * br + 7 + ext : 8 * br + (31 - __clz((int)mr)));
    }
}
__device__ __forceinline__ void run_geometry(const uint64_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ run_start,
                                             uint32_t n_runs, uint32_t r, const uint64_t* __restrict__ mm, int64_t* q0, int64_t* q1,
                                             int* rel, int64_t* d) {
    const uint64_t a = run_start[r], b = (r + 1 < n_runs) ? run_start[r + 1] : n;
    const uint64_t k0 = keys[a], k1 = keys[b - 1];
    *q0 = (int64_t)((mm[k0 & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
    *q1 = (int64_t)((mm[k1 & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
    *rel = (int)(k0 >> 62) & 1;
    *d = (int64_t)((k0 >> 29) & 0x1FFFFFFFFull) - (1ll << 32);
}
// one thread per run: the segment of its first seed (x, y) and of its last seed (z, w) in global query coordinates
__global__ void __launch_bounds__(128)
hsp_segment_kernel(const uint64_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ run_start, uint32_t n_runs,
                   const uint64_t* __restrict__ mm, const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask,
                   uint64_t tot, int k, const uint64_t* __restrict__ seg_goff, const uint64_t* __restrict__ seg_len, int n_seg,
                   uint4* __restrict__ segs) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    int64_t q0, q1, d; int rel;
    run_geometry(keys, n, run_start, n_runs, r, mm, &q0, &q1, &rel, &d);
    const int64_t s0 = rel ? (int64_t)tot - (q0 + d) - k : q0 + d;
    const int qseg = seg_of(seg_goff, n_seg, (uint64_t)q0), sseg = seg_of(seg_goff, n_seg, (uint64_t)s0);
    Diag D;
    D.seq2 = seq2; D.nmask = nmask; D.rel = rel; D.d = d; D.tot = (int64_t)tot;
    D.qlo = (int64_t)seg_goff[qseg]; D.qhi = D.qlo + (int64_t)seg_len[qseg];
    D.slo = (int64_t)seg_goff[sseg]; D.shi = D.slo + (int64_t)seg_len[sseg];
    uint4 o;
    seed_segment(D, q0, &o.x, &o.y);
    o.z = o.x; o.w = o.y;
    if (q1 + k - 1 > (int64_t)o.y) seed_segment(D, q1, &o.z, &o.w);
    segs[r] = o;
}
__device__ __forceinline__ void run_keep(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ run_start, uint32_t r,
                                         const uint4* __restrict__ segs, bool* ka, bool* kb) {
    const uint4 s = segs[r];
    bool dup_a = false;
    if (r > 0 && (keys[run_start[r - 1]] >> 29) == (keys[run_start[r]] >> 29)) {
        const uint4 p = segs[r - 1];
        dup_a = p.z == s.x && p.w == s.y;
    }
    *ka = !dup_a && (s.y - s.x + 1u >= (uint32_t)HSP_MINSPAN);
    *kb = !(s.z == s.x && s.w == s.y) && (s.w - s.z + 1u >= (uint32_t)HSP_MINSPAN);
}
// keep flags of the 2 slots of every run: a segment is dropped when it repeats its predecessor on the diagonal or is short
__global__ void hsp_keep_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ run_start, uint32_t n_runs,
                                const uint4* __restrict__ segs, uint32_t* __restrict__ keep) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    bool ka, kb;
    run_keep(keys, run_start, r, segs, &ka, &kb);
    keep[2 * (size_t)r] = ka ? 1u : 0u;
    keep[2 * (size_t)r + 1] = kb ? 1u : 0u;
}
// the kept segments as HSP records at their compacted positions
__global__ void hsp_emit_kernel(const uint64_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ run_start,
                                uint32_t n_runs, const uint64_t* __restrict__ mm, uint64_t tot, int k,
                                const uint64_t* __restrict__ seg_goff, int n_seg, const uint4* __restrict__ segs,
                                const uint32_t* __restrict__ keep, const uint32_t* __restrict__ pos, HbHspRec* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint32_t k0 = keep[2 * (size_t)r], k1 = keep[2 * (size_t)r + 1];
    if (!(k0 | k1)) return;
    int64_t q0, q1, d; int rel;
    run_geometry(keys, n, run_start, n_runs, r, mm, &q0, &q1, &rel, &d);
    const int64_t s0 = rel ? (int64_t)tot - (q0 + d) - k : q0 + d;
    const int qseg = seg_of(seg_goff, n_seg, (uint64_t)q0), sseg = seg_of(seg_goff, n_seg, (uint64_t)s0);
    const int64_t qg = (int64_t)seg_goff[qseg], sg = (int64_t)seg_goff[sseg];
    const uint4 s = segs[r];
#pragma unroll
    for (int z = 0; z < 2; ++z) {
        if (!(z ? k1 : k0)) continue;
        const int64_t a = z ? s.z : s.x, b = z ? s.w : s.y;
        HbHspRec o;
        o.qseg = qseg; o.subj = sseg; o.pad = 0;
        o.qs = (int32_t)(a - qg + 1); o.qe = (int32_t)(b - qg + 1);
        if (!rel) { o.ss = (int32_t)(a + d - sg + 1); o.se = (int32_t)(b + d - sg + 1); }
        else { o.ss = (int32_t)((int64_t)tot - 1 - (a + d) - sg + 1); o.se = (int32_t)((int64_t)tot - 1 - (b + d) - sg + 1); }
        out[pos[2 * (size_t)r + z]] = o;
    }
}

}  // namespace

cudaError_t hb_pair_count_launch(const uint64_t* mm, uint32_t n_mm, const uint64_t* idx, const uint32_t* bucket,
                                 int occ_cap, uint64_t qpos_lo, uint64_t qpos_hi, uint32_t* lo, uint32_t* cnt, cudaStream_t s) {
    if (n_mm == 0) return cudaSuccess;
    pair_count_kernel<<<(n_mm + 255) / 256, 256, 0, s>>>(mm, n_mm, idx, bucket, occ_cap, qpos_lo, qpos_hi, lo, cnt);
    return cudaGetLastError();
}
cudaError_t hb_pair_fill_launch(const uint64_t* mm, uint32_t n_mm, const uint64_t* idx, uint64_t n_idx, uint64_t tot, int k, int occ_cap,
                                const uint32_t* lo, const uint32_t* cnt, const uint32_t* off, uint64_t* keys,
                                cudaStream_t s) {
    if (n_mm == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_mm * 8;
    pair_fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(mm, n_mm, idx, n_idx, tot, k, occ_cap, lo, cnt, off, keys);
    return cudaGetLastError();
}
cudaError_t hb_hsp_flag_launch(const uint64_t* keys, uint64_t n, const uint64_t* mm, uint64_t tot, int k,
                               const uint64_t* seg_goff, int n_seg, uint32_t* flag, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    hsp_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys, n, mm, tot, k, seg_goff, n_seg, flag);
    return cudaGetLastError();
}
cudaError_t hb_hsp_start_launch(const uint32_t* flag, const uint32_t* rid, uint64_t n, uint32_t* run_start, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    hsp_start_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flag, rid, n, run_start);
    return cudaGetLastError();
}
cudaError_t hb_hsp_segment_launch(const uint64_t* keys, uint64_t n, const uint32_t* run_start, uint32_t n_runs, const uint64_t* mm,
                                  const uint32_t* seq2, const uint32_t* nmask, uint64_t tot, int k, const uint64_t* seg_goff,
                                  const uint64_t* seg_len, int n_seg, uint4* segs, uint32_t* keep, cudaStream_t s) {
    if (n_runs == 0) return cudaSuccess;
    hsp_segment_kernel<<<(n_runs + 127) / 128, 128, 0, s>>>(keys, n, run_start, n_runs, mm, seq2, nmask, tot, k, seg_goff, seg_len,
                                                             n_seg, segs);
    hsp_keep_kernel<<<(n_runs + 255) / 256, 256, 0, s>>>(keys, run_start, n_runs, segs, keep);
    return cudaGetLastError();
}
cudaError_t hb_hsp_emit_launch(const uint64_t* keys, uint64_t n, const uint32_t* run_start, uint32_t n_runs, const uint64_t* mm,
                               uint64_t tot, int k, const uint64_t* seg_goff, int n_seg, const uint4* segs, const uint32_t* keep,
                               const uint32_t* pos, HbHspRec* out, cudaStream_t s) {
    if (n_runs == 0) return cudaSuccess;
    hsp_emit_kernel<<<(n_runs + 255) / 256, 256, 0, s>>>(keys, n, run_start, n_runs, mm, tot, k, seg_goff, n_seg, segs, keep, pos, out);
    return cudaGetLastError();
}
