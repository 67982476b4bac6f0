// hite_b200 -- device-wide primitives written for this library (no CUB/Thrust on the product path):
//   * exclusive scan of uint32 (three-kernel, tile sums in a single-CTA middle pass)
//   * stable LSD radix sort of uint64 keys, 8-bit digits, shared-memory ranking + staged,
//     coalesced scatter ("shared-memory radix binning")
#include "hb_common.cuh"
#include "hb_prims.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(HB_FULL, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t s = s_warp[w];
        if (w < wid) before += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return before + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t s_warp[8];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t v = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        size_t i = base + (size_t)it * SCAN_THREADS + threadIdx.x;
        if (i < n) v += in[i];
    }
    uint32_t tot;
    block_excl_scan_256(v, s_warp, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single CTA: exclusive scan of the tile sums in place, grand total to *total.  The offsets are 32-bit (callers index
// buffers of < 2^32 elements with them) and wrap silently beyond that, so the grand total is accumulated separately in
// 64 bits: a caller compares it with its limit BEFORE it sizes or fills anything from the offsets.
__global__ void __launch_bounds__(1024)
scan_middle(uint32_t* __restrict__ tile_sums, size_t n_tiles, uint64_t* __restrict__ total) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    __shared__ unsigned long long s_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry = 0; s_total = 0; }
    __syncthreads();
    unsigned long long mine = 0;
    for (size_t b = 0; b < n_tiles; b += 1024) {
        size_t i = b + threadIdx.x;
        uint32_t v = i < n_tiles ? tile_sums[i] : 0, x = v;
        mine += v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(HB_FULL, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t y = s_w[lane], z = y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(HB_FULL, z, o);
                if (lane >= o) z += t;
            }
            s_w[lane] = z - y;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_w[wid] + x - v;
        if (i < n_tiles) tile_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry = excl + v; }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(HB_FULL, mine, o);
    if (lane == 0 && mine) atomicAdd(&s_total, mine);
    __syncthreads();
    if (threadIdx.x == 0 && total) *total = s_total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ tile_base) {
    __shared__ uint32_t s_warp[8];
    // blocked arrangement per thread keeps the scan order == index order
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        size_t i = base + it;
        v[it] = i < n ? in[i] : 0;
        sum += v[it];
    }
    uint32_t tot;
    uint32_t off = block_excl_scan_256(sum, s_warp, &tot) + tile_base[blockIdx.x];
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        size_t i = base + it;
        if (i < n) out[i] = off;
        off += v[it];
    }
}

// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS)
rs_hist(const uint64_t* __restrict__ keys, size_t n, int shift, uint32_t mask, uint32_t* __restrict__ hist, uint32_t nblk) {
    __shared__ uint32_t cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        size_t i = base + (size_t)it * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&cnt[(uint32_t)(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = cnt[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, size_t n, int shift, uint32_t mask,
           const uint32_t* __restrict__ offs, uint32_t nblk) {
    __shared__ uint32_t warp_cnt[8][256];
    __shared__ uint32_t run[256];
    __shared__ uint32_t dstart[256];
    __shared__ uint32_t s_warp[8];
    __shared__ uint64_t stage[RS_TILE];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    run[tid] = 0;
    uint64_t key[RS_ITEMS];
    uint32_t lrank[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const size_t i = base + (size_t)it * RS_THREADS + tid;
        const bool valid = i < n;
        key[it] = valid ? in[i] : ~0ull;
        const uint32_t d = valid ? ((uint32_t)(key[it] >> shift) & mask) : 0x1FFu;
#pragma unroll
        for (int w = 0; w < 8; ++w) warp_cnt[w][tid] = 0;
        __syncthreads();
        const uint32_t peers = __match_any_sync(HB_FULL, d);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_cnt[wid][d] = __popc(peers);
        __syncthreads();
        {
            uint32_t b = run[tid];
#pragma unroll
            for (int w = 0; w < 8; ++w) { uint32_t c = warp_cnt[w][tid]; warp_cnt[w][tid] = b; b += c; }
            run[tid] = b;
        }
        __syncthreads();
        lrank[it] = valid ? warp_cnt[wid][d] + rank_in_warp : 0;
        __syncthreads();
    }
    // exclusive prefix over digits -> start of each digit's run inside the staged tile
    uint32_t tot;
    const uint32_t ds = block_excl_scan_256(run[tid], s_warp, &tot);
    dstart[tid] = ds;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const size_t i = base + (size_t)it * RS_THREADS + tid;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[it] >> shift) & mask;
            stage[dstart[d] + lrank[it]] = key[it];
        }
    }
    __syncthreads();
    const uint32_t tile_n = (uint32_t)min((size_t)RS_TILE, n - base);
    for (uint32_t j = tid; j < tile_n; j += RS_THREADS) {
        const uint64_t k = stage[j];
        const uint32_t d = (uint32_t)(k >> shift) & mask;
        out[(size_t)offs[(size_t)d * nblk + blockIdx.x] + (j - dstart[d])] = k;
    }
}

}  // namespace

size_t hb_scan_tmp_words(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

cudaError_t hb_exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_tmp,
                                  uint64_t* d_total, cudaStream_t s) {
    if (n == 0) {
        if (d_total) return cudaMemsetAsync(d_total, 0, sizeof(uint64_t), s);
        return cudaSuccess;
    }
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(d_in, n, d_tmp);
    scan_middle<<<1, 1024, 0, s>>>(d_tmp, tiles, d_total);
    scan_apply<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(d_in, d_out, n, d_tmp);
    return cudaGetLastError();
}

size_t hb_sort_tmp_words(size_t n) {
    const size_t nblk = (n + RS_TILE - 1) / RS_TILE;
    return 2 * (256 * nblk) + hb_scan_tmp_words(256 * nblk) + 16;
}

// Sorts keys by bits [lo_bit, hi_bit) (stable).  Result ends in d_keys when the pass count is even,
// otherwise in d_alt; the returned pointer says which.  launches += kernels launched.
cudaError_t hb_radix_sort_u64(uint64_t* d_keys, uint64_t* d_alt, size_t n, int lo_bit, int hi_bit,
                              uint32_t* d_tmp, uint64_t** d_result, uint32_t* launches, cudaStream_t s) {
    *d_result = d_keys;
    if (n == 0) return cudaSuccess;
    if (n >= (1ull << 32)) return cudaErrorInvalidValue;
    const uint32_t nblk = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint32_t* hist = d_tmp;
    uint32_t* offs = d_tmp + (size_t)256 * nblk;
    uint32_t* scan_tmp = offs + (size_t)256 * nblk;
    uint64_t* src = d_keys;
    uint64_t* dst = d_alt;
    for (int shift = lo_bit; shift < hi_bit; shift += 8) {
        const int bits = min(8, hi_bit - shift);
        const uint32_t mask = (1u << bits) - 1u;
        rs_hist<<<nblk, RS_THREADS, 0, s>>>(src, n, shift, mask, hist, nblk);
        cudaError_t e = hb_exclusive_scan_u32(hist, offs, (size_t)256 * nblk, scan_tmp, nullptr, s);
        if (e != cudaSuccess) return e;
        rs_scatter<<<nblk, RS_THREADS, 0, s>>>(src, dst, n, shift, mask, offs, nblk);
        if (launches) *launches += 5;
        uint64_t* t = src; src = dst; dst = t;
    }
    *d_result = src;
    return cudaGetLastError();
}
