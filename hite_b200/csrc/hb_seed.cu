// hite_b200 -- kernel (ii), first half: per-candidate seed lookup in the minimizer index, diagonal
// clustering of the hits and ranking of the candidate loci.
//
// Replaces the seeding/chaining stage of `minimap2 -ax map-ont -N 300 -p 0.2` as driven by
// get_full_length_copies_minimap2 (reference module/Util.py:7952-7961); follows HB-align S3..S6
//  and is bit-identical to oracle/hbo_align.c::hbo_find_loci.
#include "hb_common.cuh"
#include "hb_seed.cuh"

namespace {

__device__ __forceinline__ uint64_t lower_bound_u64(const uint64_t* __restrict__ a, uint64_t lo, uint64_t hi, uint64_t key) {
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// S3/S4 count: one thread per candidate minimizer.
__global__ void seed_count_kernel(const uint64_t* __restrict__ cm, uint32_t n_cm, const uint64_t* __restrict__ cgoff,
                                  uint32_t n_cand, const uint64_t* __restrict__ idx, uint64_t n_idx,
                                  const uint32_t* __restrict__ bucket, int bucket_shift, int occ_cap,
                                  uint32_t* __restrict__ m_cand, uint32_t* __restrict__ m_qidx,
                                  uint32_t* __restrict__ m_lo, uint32_t* __restrict__ m_cnt) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_cm) return;
    const uint64_t e = cm[m];
    const uint64_t pos = (e >> 1) & 0xFFFFFFFFull;
    const uint32_t h = (uint32_t)(e >> 33);
    // candidate = last c with cgoff[c] <= pos
    uint32_t lo = 0, hi = n_cand;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cgoff[mid] <= pos) lo = mid + 1; else hi = mid; }
    const uint32_t c = lo - 1;
    // first minimizer of this candidate: lower_bound over positions of cgoff[c]
    uint32_t a = 0, b = m;
    const uint64_t cstart = cgoff[c];
    while (a < b) { uint32_t mid = (a + b) >> 1; if (((cm[mid] >> 1) & 0xFFFFFFFFull) < cstart) a = mid + 1; else b = mid; }
    const uint32_t qidx = m - a;
    m_cand[m] = c; m_qidx[m] = qidx;
    uint32_t cnt = 0, first = 0;
    if (qidx < 8192 && n_idx > 0) {
        const uint32_t bk = h >> (30 - HB_SEED_BUCKET_BITS);
        (void)bucket_shift;
        const uint64_t blo = bucket[bk], bhi = bucket[bk + 1];
        const uint64_t l = lower_bound_u64(idx, blo, bhi, (uint64_t)h << 33);
        const uint64_t r = lower_bound_u64(idx, l, bhi, ((uint64_t)h + 1) << 33);
        first = (uint32_t)l;
        cnt = (uint32_t)min((uint64_t)occ_cap, r - l);
    }
    m_lo[m] = first; m_cnt[m] = cnt;
}

// S4 fill: 8 lanes per candidate minimizer
__global__ void seed_fill_kernel(const uint64_t* __restrict__ cm, uint32_t n_cm, const uint64_t* __restrict__ cgoff,
                                 const uint64_t* __restrict__ idx, uint64_t tot, int k,
                                 const uint32_t* __restrict__ m_cand, const uint32_t* __restrict__ m_qidx,
                                 const uint32_t* __restrict__ m_lo, const uint32_t* __restrict__ m_cnt,
                                 const uint32_t* __restrict__ m_off, uint64_t* __restrict__ hits) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = t >> 3, sub = t & 7;
    if (m >= n_cm) return;
    const uint32_t cnt = m_cnt[m];
    if (cnt == 0) return;
    const uint64_t e = cm[m];
    const uint32_t c = m_cand[m];
    const int64_t qpos = (int64_t)(((e >> 1) & 0xFFFFFFFFull) - cgoff[c]);
    const uint32_t qs = (uint32_t)(e & 1);
    const uint64_t base = m_off[m];
    const uint64_t lo = m_lo[m];
    for (uint32_t x = sub; x < cnt; x += 8) {
        const uint64_t g = idx[lo + x];
        const int64_t gpos = (int64_t)((g >> 1) & 0xFFFFFFFFull);
        const uint32_t rel = qs ^ (uint32_t)(g & 1);
        const int64_t d = rel ? ((int64_t)tot - gpos - k) - qpos + HB_DIAG_OFF : gpos - qpos + HB_DIAG_OFF;
        hits[base + x] = ((uint64_t)c << 47) | ((uint64_t)rel << 46) | ((uint64_t)d << 13) | (uint64_t)m_qidx[m];
    }
}

// S5: cluster starts
__global__ void cluster_flag_kernel(const uint64_t* __restrict__ hits, uint64_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t f = 1;
    if (i > 0) {
        const uint64_t a = hits[i - 1], b = hits[i];
        if ((a >> 46) == (b >> 46)) {
            const int64_t da = (int64_t)((a >> 13) & 0x1FFFFFFFFull), db = (int64_t)((b >> 13) & 0x1FFFFFFFFull);
            if (db - da <= HB_DGAP) f = 0;
        }
    }
    flag[i] = f;
}
__global__ void cluster_start_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ cid, uint64_t n,
                                     uint32_t* __restrict__ cl_start) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) cl_start[cid[i]] = (uint32_t)i;
}
// S5b: a cluster wider than the band (a copy whose diagonal drifts) is cut into pieces of HB_SPLIT_SPREAD + 1 diagonals
// counted from its first diagonal; each piece gets its own band and the pieces are chained after the alignment (S8b).
// cid = exclusive scan of flag, so the cluster of hit i is cid[i] - (flag[i] ? 0 : 1).
__global__ void cluster_split_kernel(const uint64_t* __restrict__ hits, uint64_t n, const uint32_t* __restrict__ flag,
                                     const uint32_t* __restrict__ cid, const uint32_t* __restrict__ cl_start, uint32_t n_cl,
                                     uint32_t* __restrict__ flag2) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t f = flag[i];
    if (!f) {
        const uint32_t c = cid[i] - 1u;
        const uint64_t a = cl_start[c], b = (c + 1 < n_cl) ? cl_start[c + 1] : n;
        const int64_t d0 = (int64_t)((hits[a] >> 13) & 0x1FFFFFFFFull), dl = (int64_t)((hits[b - 1] >> 13) & 0x1FFFFFFFFull);
        if (dl - d0 > HB_SPLIT_SPREAD) {
            const int64_t di = (int64_t)((hits[i] >> 13) & 0x1FFFFFFFFull), dp = (int64_t)((hits[i - 1] >> 13) & 0x1FFFFFFFFull);
            f = ((di - d0) / (HB_SPLIT_SPREAD + 1) != (dp - d0) / (HB_SPLIT_SPREAD + 1)) ? 1u : 0u;
        }
    }
    flag2[i] = f;
}
// per-candidate first cluster: clusters are ordered by candidate
__global__ void cand_first_cluster_kernel(const uint64_t* __restrict__ hits, const uint32_t* __restrict__ cl_start,
                                          uint32_t n_cl, uint32_t n_cand, uint32_t* __restrict__ cand_first) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_cl) return;
    const uint32_t c = (uint32_t)(hits[cl_start[x]] >> 47);
    const uint32_t pc = x > 0 ? (uint32_t)(hits[cl_start[x - 1]] >> 47) : 0xFFFFFFFFu;
    if (x == 0 || pc != c) {
        // candidates without clusters between pc and c get an empty range later (cand_first == END marker)
        cand_first[c] = x;
    }
    (void)n_cand;
}

// S5c: the band of a locus is centred on the MEDIAN diagonal of its cluster's hits (sorted by diagonal: hits a .. b-1) and
// sized by their 5 %-trimmed spread, so a stray hit within HB_DGAP of the chain neither drags the band off the copy nor
// widens it (dmin / dmax of the locus are the trimmed bounds; the extension kernels pick their band from dmax - dmin)
__device__ __forceinline__ void locus_band(const uint64_t* __restrict__ hits, uint64_t a, uint64_t b, HbLocus* L) {
    const uint64_t n = b - a, t = n / 20;
    L->dmin = (int64_t)((hits[a + t] >> 13) & 0x1FFFFFFFFull);
    L->dmax = (int64_t)((hits[b - 1 - t] >> 13) & 0x1FFFFFFFFull);
    L->j0 = (int64_t)((hits[a + n / 2] >> 13) & 0x1FFFFFFFFull) - HB_DIAG_OFF;
}

constexpr uint32_t RANK_SMEM_KEYS = 768;     // valid clusters per candidate ranked in shared memory (12 B each, 4 warps)

// S5/S6: rank the clusters of a candidate (n desc, rel asc, dmin asc), keep the best max_loci.
// One warp per candidate.
__global__ void __launch_bounds__(128)
rank_select_kernel(const uint64_t* __restrict__ hits, uint64_t n_hits, const uint32_t* __restrict__ cl_start,
                   uint32_t n_cl, const uint32_t* __restrict__ cand_first, uint32_t n_cand, int max_loci,
                   HbLocus* __restrict__ loci, uint32_t* __restrict__ n_loci) {
    __shared__ uint64_t s_keys[4][RANK_SMEM_KEYS];
    __shared__ uint32_t s_kcl[4][RANK_SMEM_KEYS];
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n_cand) return;
    const uint32_t first = cand_first[c];
    uint32_t nvalid = 0;
    if (first != 0xFFFFFFFFu) {
        // end of this candidate's cluster range
        uint32_t end = first;
        {   // scan forward (warp-strided) for the first cluster belonging to another candidate
            uint32_t lo = first, hi = n_cl;
            while (lo < hi) {   // binary search on candidate id (monotone)
                uint32_t mid = (lo + hi) >> 1;
                if ((uint32_t)(hits[cl_start[mid]] >> 47) <= c) lo = mid + 1; else hi = mid;
            }
            end = lo;
        }
        auto cl_n = [&](uint32_t x) -> uint32_t {
            const uint64_t e = (x + 1 < n_cl) ? cl_start[x + 1] : n_hits;
            return (uint32_t)(e - cl_start[x]);
        };
        // Pass 1: the valid clusters (>= HB_MIN_SEEDS seeds) as sortable keys in shared memory -- smaller key =
        // better cluster: (seeds desc, rel asc, first diagonal asc); keys are distinct.  A candidate with a
        // low-complexity stretch (polyA tail) has thousands of one-hit clusters but few valid ones, so the
        // ranking below is O(valid^2 / 32) on shared memory instead of O(clusters^2 / 32) on global memory.
        uint64_t* keys = s_keys[threadIdx.x >> 5];
        uint32_t* kcl = s_kcl[threadIdx.x >> 5];
        uint32_t nv = 0;
        for (uint32_t x0 = first; x0 < end; x0 += 32) {
            const uint32_t x = x0 + lane;
            bool valid = false;
            uint64_t key = 0;
            if (x < end) {
                const uint32_t nx = cl_n(x);
                valid = nx >= HB_MIN_SEEDS;
                const uint64_t k0 = hits[cl_start[x]];
                key = ((uint64_t)(0xFFFFFFFFu - nx) << 34) | (((k0 >> 46) & 1ull) << 33) | ((k0 >> 13) & 0x1FFFFFFFFull);
            }
            const uint32_t bal = __ballot_sync(HB_FULL, valid);
            const uint32_t pos = nv + __popc(bal & ((1u << lane) - 1u));
            if (valid && pos < RANK_SMEM_KEYS) { keys[pos] = key; kcl[pos] = x; }
            nv += __popc(bal);
        }
        __syncwarp();
        nvalid = nv;
        if (nv <= RANK_SMEM_KEYS) {
            for (uint32_t v0 = 0; v0 < nv; v0 += 32) {
                const uint32_t v = v0 + lane;
                if (v >= nv) continue;
                const uint64_t kx = keys[v];
                uint32_t rank = 0;
                for (uint32_t y = 0; y < nv; ++y) rank += keys[y] < kx ? 1u : 0u;
                if (rank < (uint32_t)max_loci) {
                    const uint32_t x = kcl[v];
                    const uint64_t e = (x + 1 < n_cl) ? cl_start[x + 1] : n_hits;
                    HbLocus L;
                    L.cand = c; L.rel = (uint32_t)((kx >> 33) & 1ull); L.nseeds = 0xFFFFFFFFu - (uint32_t)(kx >> 34);
                    L.rank = rank;
                    locus_band(hits, cl_start[x], e, &L);
                    loci[(size_t)c * max_loci + rank] = L;
                }
            }
        } else {
            // More valid clusters than the shared-memory buffer holds (families with thousands of copies): only the best
            // max_loci matter, so first find the seed-count threshold T with  #{n > T} < max_loci <= #{n >= T}  by bisection
            // (each probe is one strided pass over the clusters), then collect every cluster with n > T and, in array
            // order -- which is (strand, first diagonal) order, the tie-break of the ranking -- the first clusters with
            // n == T that still fit.  The <= max_loci collected keys are ranked in shared memory like the small case:
            // the result equals the full O(n^2) ranking.
            auto count_ge = [&](uint32_t T) -> uint32_t {
                uint32_t cnt = 0;
                for (uint32_t x = first + lane; x < end; x += 32) cnt += cl_n(x) >= T ? 1u : 0u;
#pragma unroll
                for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(HB_FULL, cnt, o);
                return cnt;
            };
            uint32_t lo = HB_MIN_SEEDS, hi = 1u << 20;                  // invariant: count_ge(lo) >= max_loci > count_ge(hi)
            while (hi - lo > 1) {
                const uint32_t mid = lo + (hi - lo) / 2;
                if (count_ge(mid) >= (uint32_t)max_loci) lo = mid; else hi = mid;
            }
            const uint32_t T = lo;
            const uint32_t n_gt = count_ge(T + 1);
            const uint32_t need_eq = (uint32_t)max_loci - n_gt;
            uint32_t m = 0, taken_eq = 0;
            for (uint32_t x0 = first; x0 < end; x0 += 32) {
                const uint32_t x = x0 + lane;
                uint32_t nx = 0;
                uint64_t key = 0;
                if (x < end) {
                    nx = cl_n(x);
                    const uint64_t k0 = hits[cl_start[x]];
                    key = ((uint64_t)(0xFFFFFFFFu - nx) << 34) | (((k0 >> 46) & 1ull) << 33) | ((k0 >> 13) & 0x1FFFFFFFFull);
                }
                const bool gt = x < end && nx > T, eq = x < end && nx == T;
                const uint32_t bal_eq = __ballot_sync(HB_FULL, eq);
                const bool take_eq = eq && (taken_eq + __popc(bal_eq & ((1u << lane) - 1u)) < need_eq);
                const bool take = gt || take_eq;
                const uint32_t bal = __ballot_sync(HB_FULL, take);
                const uint32_t pos = m + __popc(bal & ((1u << lane) - 1u));
                if (take) { keys[pos] = key; kcl[pos] = x; }
                m += __popc(bal);
                taken_eq += __popc(bal_eq);
            }
            __syncwarp();
            for (uint32_t v0 = 0; v0 < m; v0 += 32) {
                const uint32_t v = v0 + lane;
                if (v >= m) continue;
                const uint64_t kx = keys[v];
                uint32_t rank = 0;
                for (uint32_t y = 0; y < m; ++y) rank += keys[y] < kx ? 1u : 0u;
                const uint32_t x = kcl[v];
                const uint64_t e = (x + 1 < n_cl) ? cl_start[x + 1] : n_hits;
                HbLocus L;
                L.cand = c; L.rel = (uint32_t)((kx >> 33) & 1ull); L.nseeds = 0xFFFFFFFFu - (uint32_t)(kx >> 34);
                L.rank = rank;
                locus_band(hits, cl_start[x], e, &L);
                loci[(size_t)c * max_loci + rank] = L;
            }
        }
        }
    if (lane == 0) n_loci[c] = min(nvalid, (uint32_t)max_loci);
}

// compact pair list: pair_off = exclusive scan of n_loci
__global__ void pair_emit_kernel(const HbLocus* __restrict__ loci, const uint32_t* __restrict__ n_loci,
                                 const uint32_t* __restrict__ pair_off, uint32_t n_cand, int max_loci,
                                 HbLocus* __restrict__ pairs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = t / (uint32_t)max_loci, r = t % (uint32_t)max_loci;
    if (c >= n_cand || r >= n_loci[c]) return;
    pairs[pair_off[c] + r] = loci[(size_t)c * max_loci + r];
}

}  // namespace

cudaError_t hb_seed_count_launch(const uint64_t* cm, uint32_t n_cm, const uint64_t* cgoff, uint32_t n_cand,
                                 const uint64_t* idx, uint64_t n_idx, const uint32_t* bucket, int occ_cap,
                                 uint32_t* m_cand, uint32_t* m_qidx, uint32_t* m_lo, uint32_t* m_cnt, cudaStream_t s) {
    if (n_cm == 0) return cudaSuccess;
    seed_count_kernel<<<(n_cm + 255) / 256, 256, 0, s>>>(cm, n_cm, cgoff, n_cand, idx, n_idx, bucket, 0, occ_cap,
                                                          m_cand, m_qidx, m_lo, m_cnt);
    return cudaGetLastError();
}
cudaError_t hb_seed_fill_launch(const uint64_t* cm, uint32_t n_cm, const uint64_t* cgoff, const uint64_t* idx,
                                uint64_t tot, int k, const uint32_t* m_cand, const uint32_t* m_qidx,
                                const uint32_t* m_lo, const uint32_t* m_cnt, const uint32_t* m_off, uint64_t* hits,
                                cudaStream_t s) {
    if (n_cm == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n_cm * 8;
    seed_fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(cm, n_cm, cgoff, idx, tot, k, m_cand, m_qidx,
                                                                       m_lo, m_cnt, m_off, hits);
    return cudaGetLastError();
}
cudaError_t hb_cluster_flag_launch(const uint64_t* hits, uint64_t n, uint32_t* flag, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    cluster_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hits, n, flag);
    return cudaGetLastError();
}
cudaError_t hb_cluster_split_launch(const uint64_t* hits, uint64_t n, const uint32_t* flag, const uint32_t* cid,
                                    const uint32_t* cl_start, uint32_t n_cl, uint32_t* flag2, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    cluster_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hits, n, flag, cid, cl_start, n_cl, flag2);
    return cudaGetLastError();
}
cudaError_t hb_cluster_start_launch(const uint32_t* flag, const uint32_t* cid, uint64_t n, uint32_t* cl_start,
                                    cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    cluster_start_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flag, cid, n, cl_start);
    return cudaGetLastError();
}
cudaError_t hb_rank_select_launch(const uint64_t* hits, uint64_t n_hits, const uint32_t* cl_start, uint32_t n_cl,
                                  uint32_t* cand_first, uint32_t n_cand, int max_loci, HbLocus* loci,
                                  uint32_t* n_loci, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(cand_first, 0xFF, sizeof(uint32_t) * n_cand, s);
    if (e != cudaSuccess) return e;
    if (n_cl > 0)
        cand_first_cluster_kernel<<<(n_cl + 255) / 256, 256, 0, s>>>(hits, cl_start, n_cl, n_cand, cand_first);
    const uint64_t threads = (uint64_t)n_cand * 32;
    rank_select_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(hits, n_hits, cl_start, n_cl, cand_first,
                                                                          n_cand, max_loci, loci, n_loci);
    return cudaGetLastError();
}
cudaError_t hb_pair_emit_launch(const HbLocus* loci, const uint32_t* n_loci, const uint32_t* pair_off,
                                uint32_t n_cand, int max_loci, HbLocus* pairs, cudaStream_t s) {
    const uint64_t threads = (uint64_t)n_cand * (uint64_t)max_loci;
    if (threads == 0) return cudaSuccess;
    pair_emit_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(loci, n_loci, pair_off, n_cand, max_loci, pairs);
    return cudaGetLastError();
}
