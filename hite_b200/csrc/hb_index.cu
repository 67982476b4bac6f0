// hite_b200 -- kernel (i): 2-bit packing and (k,w)-minimizer index build over the genome chunks.
//
// Replaces the `minimap2 -d ref.mmi` index build of get_full_length_copies_minimap2 (reference
// module/Util.py:7938-7947) -- the reference shells out; the index here follows HB-align S1/S2
//  and is bit-identical to oracle/hbo_align.c::hbo_minimizers + sort.
//
// HBM layout: sequences live in one padded global coordinate space (64 N between contigs):
//   seq2[TOT/16]  uint32, base b of a word at bits 2*(b%16)         (2 bit/base)
//   nmask[TOT/32] uint32, bit b%32 set = not ACGT                   (1 bit/base)
//   index[M]      uint64  hash(30b)<<33 | pos(32b)<<1 | strand, sorted ascending
//   bucket[2^22+1] uint32 first index entry of every 22-bit hash prefix
#include "hb_common.cuh"
#include "hb_index.cuh"

namespace {

// ---- S1: pack ---------------------------------------------------------------------------------
// One thread per 16 bases = one seq2 word; a lane pair shares one N-mask word.  The 64-base pads guarantee that a word
// intersects at most one contig.  A word that lies wholly inside its contig takes the fast path: its 16 ASCII bytes come
// from TWO aligned 16-byte loads (LDG.128; the source of a contig is not 16-byte aligned with the packed space, so the
// bytes are re-aligned with funnel shifts) and are classified four at a time in 32-bit registers: upper-casing by & 0xDF,
// code = ((c >> 1) & 3) ^ (((c >> 1) & 3) >> 1) maps A C G T -> 0 1 2 3, validity by four byte-wise compares, and the
// 4 x 2 code bits / 4 validity bits of a register are gathered by one multiply each.  Words cut by a contig boundary (two
// per contig) and words in the pads take the byte path.
__device__ __forceinline__ void pack4(uint32_t w, uint32_t* code8, uint32_t* valid4) {
    const uint32_t u = w & 0xDFDFDFDFu;
    const uint32_t x = (u >> 1) & 0x03030303u;
    const uint32_t c = x ^ ((x >> 1) & 0x01010101u);
    *code8 = (c * 0x01041040u) >> 24;                     // b0 | b1 << 2 | b2 << 4 | b3 << 6
    const uint32_t v = (__vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u)) &
                       0x01010101u;
    *valid4 = ((v * 0x01020408u) >> 24) & 0xFu;           // bit i = byte i is one of A C G T (either case)
}
__device__ __forceinline__ uint4 ldg128(const uint4* p) {      // one LDG.E.128 (the compiler otherwise splits it into two 64-bit loads)
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t* __restrict__ ascii, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ goff,
            const uint64_t* __restrict__ glen, uint32_t n_contig, uint64_t tot, uint32_t* __restrict__ seq2,
            uint32_t* __restrict__ nmask) {
    const uint64_t wq = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;       // seq2 word
    const uint64_t g0 = wq * 16;
    uint32_t word = 0, nm16 = 0xFFFFu;
    if (g0 < tot) {
        int lo = 0, hi = (int)n_contig;                   // last contig with goff <= g0 + 15
        while (lo < hi) { int mid = (lo + hi) >> 1; if (goff[mid] <= g0 + 15) lo = mid + 1; else hi = mid; }
        const int c = lo - 1;
        if (c >= 0) {
            const uint64_t cs = goff[c], ce = cs + glen[c];
            const uint8_t* src = ascii + src_off[c];
            if (g0 >= cs && g0 + 16 <= ce) {
                const uint8_t* p = src + (g0 - cs);
                const uintptr_t a = reinterpret_cast<uintptr_t>(p);
                const uint4* q = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
                const uint4 v0 = ldg128(q);
                const int sh = (int)(a & 15);
                uint32_t w0 = v0.x, w1 = v0.y, w2 = v0.z, w3 = v0.w;
                if (sh) {                                 // bytes sh .. sh + 15 of the 32 loaded (the upload buffer has 16 B of slack)
                    const uint4 v1 = ldg128(q + 1);
                    const uint32_t r[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    const int qw = sh >> 2, bs = 8 * (sh & 3);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        uint32_t lo32 = 0, hi32 = 0;
#pragma unroll
                        for (int z = 0; z < 4; ++z) { if (qw == z) { lo32 = r[z + t]; hi32 = r[z + t + 1]; } }
                        const uint32_t o = __funnelshift_r(lo32, hi32, bs);
                        if (t == 0) w0 = o; else if (t == 1) w1 = o; else if (t == 2) w2 = o; else w3 = o;
                    }
                }
                uint32_t c0, c1, c2, c3, k0, k1, k2, k3;
                pack4(w0, &c0, &k0); pack4(w1, &c1, &k1); pack4(w2, &c2, &k2); pack4(w3, &c3, &k3);
                const uint32_t valid = k0 | (k1 << 4) | (k2 << 8) | (k3 << 12);
                // codes of invalid bytes are cleared like the byte path leaves them (0)
                uint32_t spread = valid;                  // 16 validity bits -> 2 bits each
                spread = (spread | (spread << 8)) & 0x00FF00FFu;
                spread = (spread | (spread << 4)) & 0x0F0F0F0Fu;
                spread = (spread | (spread << 2)) & 0x33333333u;
                spread = (spread | (spread << 1)) & 0x55555555u;
                word = (c0 | (c1 << 8) | (c2 << 16) | (c3 << 24)) & (spread | (spread << 1));
                nm16 = ~valid & 0xFFFFu;
            } else {
#pragma unroll 4
                for (int b = 0; b < 16; ++b) {
                    const uint64_t p = g0 + b;
                    if (p >= cs && p < ce) {
                        const int sy = hb_sym_of_ascii(src[p - cs]);
                        if (sy < 4) { nm16 &= ~(1u << b); word |= (uint32_t)sy << (2 * b); }
                    }
                }
            }
        }
        seq2[wq] = word;
    }
    const uint32_t other = __shfl_xor_sync(HB_FULL, nm16, 1);
    if (!(threadIdx.x & 1) && g0 < tot) nmask[wq >> 1] = nm16 | (other << 16);
}

// ---- known-TE masking: N-fill of [start, end) intervals (mask_genome_intactTE, module/Util.py:6414-6420) -------------
// One warp per interval; the interval's span of the N-mask is OR-ed word by word (atomicOr: intervals overlap).  The
// 2-bit codes stay: every reader tests the mask first.  *n_new counts bases that were not N before.
__global__ void __launch_bounds__(256)
mask_fill_kernel(const int32_t* __restrict__ contig, const int64_t* __restrict__ start, const int64_t* __restrict__ end,
                 uint32_t n, const uint64_t* __restrict__ goff, const uint64_t* __restrict__ glen, uint32_t n_contig,
                 uint32_t* __restrict__ nmask, unsigned long long* __restrict__ n_new) {
    const uint32_t iv = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (iv >= n) return;
    const int32_t c = contig[iv];
    if (c < 0 || (uint32_t)c >= n_contig) return;
    // numpy slice semantics of ref_seq_array[start:end] (non-negative bounds): clipped to the contig
    int64_t a = start[iv], b = end[iv];
    if (a < 0) a = 0;
    if (b > (int64_t)glen[c]) b = (int64_t)glen[c];
    if (b <= a) return;
    const uint64_t g0 = goff[c] + (uint64_t)a, g1 = goff[c] + (uint64_t)b;     // [g0, g1) in the padded space
    const uint64_t w0 = g0 >> 5, w1 = (g1 - 1) >> 5;
    unsigned long long fresh = 0;
    for (uint64_t w = w0 + lane; w <= w1; w += 32) {
        uint32_t m = 0xffffffffu;
        if (w == w0) m &= 0xffffffffu << (g0 & 31);
        if (w == w1) m &= 0xffffffffu >> (31 - ((g1 - 1) & 31));
        const uint32_t old = atomicOr(&nmask[w], m);
        fresh += __popc(m & ~old);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) fresh += __shfl_xor_sync(HB_FULL, fresh, o);
    if (lane == 0 && fresh) atomicAdd(n_new, fresh);
}

// packed contig -> ASCII (A C G T N), 16 bases per thread, one 16 B store
__global__ void export_kernel(const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask, uint64_t g0, uint64_t len,
                              uint8_t* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t o = t * 16;
    if (o >= len) return;
    uint8_t b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint64_t p = g0 + o + i;
        const bool isn = (nmask[p >> 5] >> (p & 31)) & 1u;
        b[i] = isn ? (uint8_t)'N' : (uint8_t)"ACGT"[(seq2[p >> 4] >> (2 * (p & 15))) & 3u];
    }
    if (o + 16 <= len && (((uintptr_t)(out + o)) & 15) == 0) *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(b);
    else for (int i = 0; i < 16 && o + i < len; ++i) out[o + i] = b[i];
}

// ---- S2: minimizers -----------------------------------------------------------------------------
constexpr int MM_THREADS = 256;
constexpr int MM_PER_THREAD = 8;
constexpr int MM_TILE = MM_THREADS * MM_PER_THREAD;     // positions owned by a CTA
constexpr int MM_HALO = 16;                             // >= w-1 on both sides
constexpr uint32_t INVALID_H = 0xFFFFFFFFu;
constexpr int MM_WORDS = (MM_TILE + 2 * MM_HALO + 16 + 15) / 16;   // seq2 words holding the tile + halo + k - 1 bases

__device__ __forceinline__ uint32_t hash_kmer(uint64_t key, int k) {   // sketch.c hash64, masked
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return (uint32_t)key;
}

// the same mix on 32-bit registers: every step keeps only the low 2k <= 30 bits, which +, << and ~ determine from the low
// bits of their inputs, and the right shifts act on values already masked to 2k bits
__device__ __forceinline__ uint32_t hash_kmer32(uint32_t key, uint32_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

__device__ __forceinline__ int base_at(const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask,
                                       uint64_t tot, int64_t p) {
    if (p < 0 || (uint64_t)p >= tot) return 4;
    if ((nmask[p >> 5] >> (p & 31)) & 1u) return 4;
    return (seq2[p >> 4] >> (2 * (p & 15))) & 3u;
}

// WRITE = false: count selected positions per CTA; WRITE = true: emit keys at cta_base[blockIdx]
template <bool WRITE>
__global__ void __launch_bounds__(MM_THREADS)
minimizer_kernel(const uint32_t* __restrict__ seq2, const uint32_t* __restrict__ nmask, uint64_t tot, int k, int w,
                 uint32_t* __restrict__ cta_count, const uint32_t* __restrict__ cta_base, uint64_t* __restrict__ out) {
    __shared__ uint32_t s_h[MM_TILE + 2 * MM_HALO];      // hash | strand in bit 31? no: hash < 2^30
    __shared__ uint8_t s_st[MM_TILE + 2 * MM_HALO];
    __shared__ uint8_t s_sel[MM_TILE + 2 * MM_HALO];
    __shared__ uint32_t s_warp[8];
    const int tid = threadIdx.x;
    const int64_t tile0 = (int64_t)blockIdx.x * MM_TILE;
    const int64_t ext0 = tile0 - MM_HALO;
    const int ext_n = MM_TILE + 2 * MM_HALO;
    // The packed words of the tile (+ halo + k - 1 bases) are staged once in shared memory; every hash is then cut out of a
    // 64-bit window of the 2-bit stream held in registers: no per-base loads.  The stream has base t of a k-mer at bits 2t,
    // which IS the reverse-complement k-mer once complemented; the forward k-mer is the same slice with its 2-bit groups
    // reversed (brev + swap inside the pairs).  tile0 and MM_HALO are multiples of 16, so ext0 is word aligned.
    __shared__ uint32_t s_seq[MM_WORDS + 2];
    __shared__ uint32_t s_nm[MM_WORDS / 2 + 2];
    {
        const int64_t w0 = ext0 >> 4;                                  // first seq2 word (ext0 may be -16: word -1)
        const int64_t n_words = (int64_t)((tot + 15) >> 4), n_nwords = (int64_t)((tot + 31) >> 5);
        for (int x = tid; x < MM_WORDS + 2; x += MM_THREADS) {
            const int64_t wi = w0 + x;
            s_seq[x] = (wi >= 0 && wi < n_words) ? seq2[wi] : 0u;
        }
        const int64_t nw0 = ext0 >> 5;                                 // floor, also for -16
        for (int x = tid; x < MM_WORDS / 2 + 2; x += MM_THREADS) {
            const int64_t wi = nw0 + x;
            uint32_t m = (wi >= 0 && wi < n_nwords) ? nmask[wi] : 0xFFFFFFFFu;
            // positions at or beyond tot inside the last word count as N
            if (wi >= 0 && wi < n_nwords && (uint64_t)(wi + 1) * 32 > tot) m |= ~0u << (tot & 31);
            s_nm[x] = m;
        }
    }
    __syncthreads();
    const uint32_t kmask = (k == 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    const uint32_t nmask_k = (1u << k) - 1u;
    const int nm_shift0 = (int)(ext0 - ((ext0 >> 5) << 5));            // 0 or 16: offset of ext0 inside its nmask word
    for (int x0 = tid * 9; x0 < ext_n; x0 += MM_THREADS * 9) {
        // 9 consecutive positions need 9 + k - 1 <= 24 bases = 48 bits from bit offset 2 * (x0 & 15) of word x0 >> 4
        const int wq = x0 >> 4, off = 2 * (x0 & 15);
        const uint64_t lo = (uint64_t)s_seq[wq] | ((uint64_t)s_seq[wq + 1] << 32);
        const uint64_t hi = (uint64_t)s_seq[wq + 2];
        const uint64_t stream = off ? ((lo >> off) | (hi << (64 - off))) : lo;
        const int nb = x0 + nm_shift0;                                 // bit index into the staged N-mask stream
        const uint64_t nlo = (uint64_t)s_nm[nb >> 5] | ((uint64_t)s_nm[(nb >> 5) + 1] << 32);
        const uint32_t nbits = (uint32_t)(nlo >> (nb & 31));           // >= 32 - 0 ... at least 32 valid bits: 24 needed
        const int64_t p0 = ext0 + x0;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (x0 + j >= ext_n) break;
            const uint32_t slice = (uint32_t)(stream >> (2 * j)) & kmask;
            const uint32_t r = ~slice & kmask;
            uint32_t y = __brev(slice);
            y = ((y & 0xAAAAAAAAu) >> 1) | ((y & 0x55555555u) << 1);
            const uint32_t f = y >> (32 - 2 * k);
            const bool has_n = ((nbits >> j) & nmask_k) != 0u;
            uint32_t h = INVALID_H; uint8_t st = 0;
            if (!has_n && f != r && p0 + j >= 0 && (uint64_t)(p0 + j) + (uint64_t)k <= tot) {
                st = r < f;
                h = hash_kmer32(r < f ? r : f, kmask);
            }
            s_h[x0 + j] = h; s_st[x0 + j] = st; s_sel[x0 + j] = 0;
        }
    }
    __syncthreads();
    // every window start s in [tile0 - (w-1), tile0 + MM_TILE) marks its leftmost smallest k-mer
    for (int x = tid; x < MM_TILE + (w - 1); x += MM_THREADS) {
        const int xs = x + (MM_HALO - (w - 1));          // index of window start in s_h
        const int64_t s = ext0 + xs;
        if (s < 0 || (uint64_t)s + (uint64_t)w > tot) continue;
        uint32_t best = INVALID_H; int bp = 0;
        for (int t = 0; t < w; ++t) { const uint32_t h = s_h[xs + t]; if (h < best) { best = h; bp = xs + t; } }
        if (best != INVALID_H) s_sel[bp] = 1;
    }
    __syncthreads();
    // compaction over the owned positions (blocked per thread -> position order)
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < MM_PER_THREAD; ++j) {
        const int x = MM_HALO + tid * MM_PER_THREAD + j;
        const int64_t p = ext0 + x;
        cnt += (p >= 0 && (uint64_t)p < tot && s_sel[x]) ? 1u : 0u;
    }
    // block exclusive scan
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t xsum = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(HB_FULL, xsum, o); if (lane >= o) xsum += t; }
    if (lane == 31) s_warp[wid] = xsum;
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { uint32_t sv = s_warp[q]; if (q < wid) before += sv; total += sv; }
    if (!WRITE) {
        if (tid == 0) cta_count[blockIdx.x] = total;
        return;
    }
    uint64_t o = (uint64_t)cta_base[blockIdx.x] + before + xsum - cnt;
#pragma unroll
    for (int j = 0; j < MM_PER_THREAD; ++j) {
        const int x = MM_HALO + tid * MM_PER_THREAD + j;
        const int64_t p = ext0 + x;
        if (p >= 0 && (uint64_t)p < tot && s_sel[x])
            out[o++] = ((uint64_t)s_h[x] << 33) | ((uint64_t)p << 1) | (uint64_t)s_st[x];
    }
}

// ---- bucket table over the sorted index ----------------------------------------------------------
// bucket[b] = first index entry whose hash prefix is >= b, for b in [0, nb].  One thread per BUCKET finds it by binary search in
// the sorted index.  Minimizer hashes are window minima, so they crowd the low end of the hash space (99.9 % of the keys of a
// genome sit below 0.55 of it): the boundary-detection version (one thread per key, each writing the buckets its key opens)
// left single threads filling gaps of 10^5 - 10^6 empty buckets one by one -- 1.9 ms of a 4.2 ms index build at 100 Mbp with
// the SMs idle.  Here every thread does log2(n) probes whose upper levels stay in L2; no atomics, no tail.
__global__ void bucket_bounds_kernel(const uint64_t* __restrict__ idx, uint64_t n, int shift, uint32_t nb, uint32_t* __restrict__ bucket) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if ((idx[mid] >> shift) < b) lo = mid + 1; else hi = mid;
    }
    bucket[b] = (uint32_t)lo;
}

}  // namespace

cudaError_t hb_pack_launch(const uint8_t* d_ascii, const uint64_t* d_src_off, const uint64_t* d_goff,
                           const uint64_t* d_glen, uint32_t n_contig, uint64_t tot, uint32_t* d_seq2,
                           uint32_t* d_nmask, cudaStream_t s) {
    const uint64_t words = tot / 16;                      // tot is a multiple of 64
    pack_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(d_ascii, d_src_off, d_goff, d_glen, n_contig, tot,
                                                                 d_seq2, d_nmask);
    return cudaGetLastError();
}

uint32_t hb_minimizer_ctas(uint64_t tot) { return (uint32_t)((tot + MM_TILE - 1) / MM_TILE); }

cudaError_t hb_minimizer_count_launch(const uint32_t* d_seq2, const uint32_t* d_nmask, uint64_t tot, int k, int w,
                                      uint32_t* d_cta_count, cudaStream_t s) {
    minimizer_kernel<false><<<hb_minimizer_ctas(tot), MM_THREADS, 0, s>>>(d_seq2, d_nmask, tot, k, w, d_cta_count,
                                                                          nullptr, nullptr);
    return cudaGetLastError();
}
cudaError_t hb_minimizer_write_launch(const uint32_t* d_seq2, const uint32_t* d_nmask, uint64_t tot, int k, int w,
                                      const uint32_t* d_cta_base, uint64_t* d_out, cudaStream_t s) {
    minimizer_kernel<true><<<hb_minimizer_ctas(tot), MM_THREADS, 0, s>>>(d_seq2, d_nmask, tot, k, w, nullptr,
                                                                         d_cta_base, d_out);
    return cudaGetLastError();
}
cudaError_t hb_bucket_bounds_launch(const uint64_t* d_idx, uint64_t n, int shift, uint32_t nb, uint32_t* d_bucket, cudaStream_t s) {
    bucket_bounds_kernel<<<(unsigned)(((uint64_t)nb + 1 + 255) / 256), 256, 0, s>>>(d_idx, n, shift, nb, d_bucket);
    return cudaGetLastError();
}
cudaError_t hb_mask_fill_launch(const int32_t* d_contig, const int64_t* d_start, const int64_t* d_end, uint32_t n,
                                const uint64_t* d_goff, const uint64_t* d_glen, uint32_t n_contig, uint32_t* d_nmask,
                                unsigned long long* d_n_new, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint64_t threads = (uint64_t)n * 32;
    mask_fill_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_contig, d_start, d_end, n, d_goff, d_glen, n_contig, d_nmask, d_n_new);
    return cudaGetLastError();
}
cudaError_t hb_export_launch(const uint32_t* d_seq2, const uint32_t* d_nmask, uint64_t g0, uint64_t len, uint8_t* d_out, cudaStream_t s) {
    if (len == 0) return cudaSuccess;
    const uint64_t threads = (len + 15) / 16;
    export_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_seq2, d_nmask, g0, len, d_out);
    return cudaGetLastError();
}
