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
// one thread per run: valid HSPs are marked (keep flag) and described
__global__ void hsp_emit_kernel(const uint64_t* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ run_start,
                                uint32_t n_runs, const uint64_t* __restrict__ mm, uint64_t tot, int k,
                                const uint64_t* __restrict__ seg_goff, int n_seg, HbHspRec* __restrict__ recs,
                                uint32_t* __restrict__ keep) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint64_t a = run_start[r], b = (r + 1 < n_runs) ? run_start[r + 1] : n;
    const uint64_t k0 = keys[a], k1 = keys[b - 1];
    const int64_t q0 = (int64_t)((mm[k0 & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
    const int64_t q1 = (int64_t)((mm[k1 & 0x1FFFFFFFull] >> 1) & 0xFFFFFFFFull);
    const int rel = (int)(k0 >> 62) & 1;
    const int64_t d = (int64_t)((k0 >> 29) & 0x1FFFFFFFFull) - (1ll << 32);
    const bool ok = (q1 + k - q0) >= HSP_MINSPAN;
    keep[r] = ok ? 1u : 0u;
    if (!ok) return;
    HbHspRec o;
    const int qseg = seg_of(seg_goff, n_seg, (uint64_t)q0);
    const int64_t s0 = rel ? (int64_t)tot - (q0 + d) - k : q0 + d;      // forward start of the first / last seed
    const int64_t s1 = rel ? (int64_t)tot - (q1 + d) - k : q1 + d;
    const int sseg = seg_of(seg_goff, n_seg, (uint64_t)s0);
    const int64_t qg = (int64_t)seg_goff[qseg], sg = (int64_t)seg_goff[sseg];
    o.qseg = qseg; o.subj = sseg;
    o.qs = (int32_t)(q0 - qg + 1); o.qe = (int32_t)(q1 + k - qg);
    if (!rel) { o.ss = (int32_t)(s0 - sg + 1); o.se = (int32_t)(s1 + k - sg); }
    else { o.ss = (int32_t)(s0 + k - sg); o.se = (int32_t)(s1 - sg + 1); }
    o.pad = 0;
    recs[r] = o;
}
__global__ void hsp_compact_kernel(const HbHspRec* __restrict__ recs, const uint32_t* __restrict__ keep,
                                   const uint32_t* __restrict__ pos, uint32_t n_runs, HbHspRec* __restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_runs && keep[r]) out[pos[r]] = recs[r];
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
cudaError_t hb_hsp_emit_launch(const uint64_t* keys, uint64_t n, const uint32_t* run_start, uint32_t n_runs,
                               const uint64_t* mm, uint64_t tot, int k, const uint64_t* seg_goff, int n_seg,
                               HbHspRec* recs, uint32_t* keep, cudaStream_t s) {
    if (n_runs == 0) return cudaSuccess;
    hsp_emit_kernel<<<(n_runs + 255) / 256, 256, 0, s>>>(keys, n, run_start, n_runs, mm, tot, k, seg_goff, n_seg, recs, keep);
    return cudaGetLastError();
}
cudaError_t hb_hsp_compact_launch(const HbHspRec* recs, const uint32_t* keep, const uint32_t* pos, uint32_t n_runs,
                                  HbHspRec* out, cudaStream_t s) {
    if (n_runs == 0) return cudaSuccess;
    hsp_compact_kernel<<<(n_runs + 255) / 256, 256, 0, s>>>(recs, keep, pos, n_runs, out);
    return cudaGetLastError();
}
