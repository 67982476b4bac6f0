// hite_b200 -- host orchestration behind the C-ABI for the genome / index / boundary-adjustment
// entry points (include/hite_b200.h).  All compute is in the kernels of hb_index.cu, hb_prims.cu,
// hb_seed.cu, hb_extend.cu, hb_msa.cu and hb_vote.cu; this file only sizes buffers, launches and
// copies.  One stream per context; stage times are CUDA events on that stream.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdlib>

#include <chrono>
#include "hb_coarse.cuh"
#include "hb_ctx.cuh"
#include "hb_extend.cuh"
#include "hb_fmea.cuh"
#include "hb_index.cuh"
#include "hb_msa.cuh"
#include "hb_prims.cuh"
#include "hb_seed.cuh"

extern "C" size_t hb_vote_match_entry_size();
extern "C" size_t hb_vote_state_entries();
extern "C" size_t hb_vote_match_entries(int C);
extern "C" cudaError_t hb_vote_launch(const HbVoteTask* d_tasks, int n, const uint8_t* d_raw, uint8_t* d_clean,
                                      uint32_t* d_planes, void* d_matches, int* d_anchors, const uint8_t* d_cands,
                                      uint8_t* d_cons, HbVoteCall* d_calls, int te_type, int avg_rows, cudaStream_t stream);

namespace {

__global__ void u32_field_kernel(const HbVoteCall* __restrict__ calls, uint32_t n, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)calls[i].cons_len;
}
// warp per candidate: copy its trimmed consensus to the dense output, rewrite cons_off
__global__ void cons_compact_kernel(HbVoteCall* __restrict__ calls, uint32_t n, const uint32_t* __restrict__ off,
                                    const uint8_t* __restrict__ cons, uint8_t* __restrict__ dense) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= n) return;
    const int len = calls[c].cons_len;
    const uint8_t* src = cons + calls[c].cons_off;
    uint8_t* dst = dense + off[c];
    for (int i = lane; i < len; i += 32) dst[i] = src[i];
    __syncwarp();
    if (lane == 0) calls[c].cons_off = (int64_t)off[c];
}
// truncated-member gate (Util.py:10442-10446), resolved on the device before the consensus bytes are
// compacted: a failed truncated pass overrides the full pass's verdict; truncated consensi are dropped
__global__ void gate_resolve_kernel(HbVoteCall* __restrict__ calls, uint32_t n_full, uint32_t n_tasks,
                                    const uint32_t* __restrict__ trunc_to_full) {
    const uint32_t t = n_full + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    HbVoteCall c = calls[t];
    if (!c.is_te) {
        c.flags |= HB_FLAG_TRUNC_REJECT;
        c.cons_len = 0;
        calls[trunc_to_full[t - n_full]] = c;
    }
    calls[t].cons_len = 0;
}
__global__ void clen_kernel(const uint64_t* __restrict__ coff, uint32_t n, uint32_t* __restrict__ clen) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) clen[i] = (uint32_t)(coff[i + 1] - coff[i]);
}

struct Layout {            // padded global coordinate space (HB-align S1)
    std::vector<uint64_t> goff, glen;
    uint64_t tot = 0;
};
Layout make_layout(const uint64_t* off, uint32_t n) {
    Layout L;
    L.goff.resize(n); L.glen.resize(n);
    uint64_t p = HB_PAD;
    for (uint32_t c = 0; c < n; ++c) { L.goff[c] = p; L.glen[c] = off[c + 1] - off[c]; p += L.glen[c] + HB_PAD; }
    L.tot = (p + 63) & ~(uint64_t)63;
    return L;
}

#define ENSURE(name, bytes) HB_CUDA_OK(ctx->buf(name).ensure((size_t)(bytes)))
#define ENSURE_L(name, bytes) HB_CUDA_OK(LB(name).ensure((size_t)(bytes)))

// pack + sketch a sequence set into (seq2, nmask, minimizer keys in position order). Returns count.
int sketch(hb_ctx* ctx, const char* prefix, const uint8_t* h_ascii, const uint64_t* h_off, uint32_t n,
           const Layout& L, uint64_t* n_mm_out, cudaStream_t s) {
    std::string P(prefix);
    auto B = [&](const char* suffix) -> HbDevBuf& { return ctx->buf((P + suffix).c_str()); };
    const uint64_t ascii_bytes = h_off[n];
    HB_CUDA_OK(B("ascii").ensure(ascii_bytes + 1));
    HB_CUDA_OK(B("src_off").ensure(sizeof(uint64_t) * (n + 1)));
    HB_CUDA_OK(B("goff").ensure(sizeof(uint64_t) * (n + 1)));
    HB_CUDA_OK(B("glen").ensure(sizeof(uint64_t) * (n + 1)));
    HB_CUDA_OK(B("seq2").ensure(L.tot / 4 + 64));
    HB_CUDA_OK(B("nmask").ensure(L.tot / 8 + 64));
    HB_CUDA_OK(cudaMemcpyAsync(B("ascii").p, h_ascii, ascii_bytes, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(B("src_off").p, h_off, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(B("goff").p, L.goff.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(B("glen").p, L.glen.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(hb_pack_launch(B("ascii").as<uint8_t>(), B("src_off").as<uint64_t>(), B("goff").as<uint64_t>(),
                              B("glen").as<uint64_t>(), n, L.tot, B("seq2").as<uint32_t>(), B("nmask").as<uint32_t>(), s));
    const uint32_t ctas = hb_minimizer_ctas(L.tot);
    HB_CUDA_OK(B("mm_cnt").ensure(sizeof(uint32_t) * (ctas + 1)));
    HB_CUDA_OK(B("mm_base").ensure(sizeof(uint32_t) * (ctas + 1)));
    HB_CUDA_OK(B("scan_tmp").ensure(sizeof(uint32_t) * hb_scan_tmp_words(ctas)));
    HB_CUDA_OK(B("total").ensure(sizeof(uint64_t)));
    HB_CUDA_OK(hb_minimizer_count_launch(B("seq2").as<uint32_t>(), B("nmask").as<uint32_t>(), L.tot, ctx->k, ctx->w,
                                         B("mm_cnt").as<uint32_t>(), s));
    HB_CUDA_OK(hb_exclusive_scan_u32(B("mm_cnt").as<uint32_t>(), B("mm_base").as<uint32_t>(), ctas,
                                     B("scan_tmp").as<uint32_t>(), B("total").as<uint64_t>(), s));
    uint64_t n_mm = 0;
    HB_CUDA_OK(cudaMemcpyAsync(&n_mm, B("total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    HB_CUDA_OK(B("mm").ensure(sizeof(uint64_t) * (n_mm + 1)));
    HB_CUDA_OK(hb_minimizer_write_launch(B("seq2").as<uint32_t>(), B("nmask").as<uint32_t>(), L.tot, ctx->k, ctx->w,
                                         B("mm_base").as<uint32_t>(), B("mm").as<uint64_t>(), s));
    ctx->timing.launches += 6;
    *n_mm_out = n_mm;
    return HB_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" int hb_genome_load(hb_ctx* ctx, const uint8_t* ascii, const uint64_t* contig_off, uint32_t n_contig) {
    if (!ctx || !ascii || !contig_off || n_contig == 0) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_load: bad argument");
    cudaSetDevice(ctx->device);
    Layout L = make_layout(contig_off, n_contig);
    if (L.tot >= (1ull << 32)) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_load: genome + pads must be < 2^32 bp");
    ctx->g_goff = L.goff; ctx->g_glen = L.glen; ctx->g_tot = L.tot; ctx->g_ncontig = n_contig;
    ctx->have_genome = false; ctx->have_index = false;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    cudaStream_t s = ctx->stream;
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    // pack only (the sketch is done by hb_index_build so k/w can be chosen afterwards)
    const uint64_t ascii_bytes = contig_off[n_contig];
    ENSURE("g.ascii", ascii_bytes + 1);
    ENSURE("g.src_off", sizeof(uint64_t) * (n_contig + 1));
    ENSURE("g.goff", sizeof(uint64_t) * (n_contig + 1));
    ENSURE("g.glen", sizeof(uint64_t) * (n_contig + 1));
    ENSURE("g.seq2", L.tot / 4 + 64);
    ENSURE("g.nmask", L.tot / 8 + 64);
    ENSURE("g.crank", sizeof(int32_t) * (n_contig + 1));
    {   // name order of the contigs for the S9 tie rule: the contig index until hb_genome_contig_rank says otherwise
        std::vector<int32_t> ident(n_contig);
        for (uint32_t c = 0; c < n_contig; ++c) ident[c] = (int32_t)c;
        HB_CUDA_OK(cudaMemcpy(ctx->buf("g.crank").p, ident.data(), sizeof(int32_t) * n_contig, cudaMemcpyHostToDevice));
    }
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("g.ascii").p, ascii, ascii_bytes, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("g.src_off").p, contig_off, sizeof(uint64_t) * (n_contig + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("g.goff").p, L.goff.data(), sizeof(uint64_t) * n_contig, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("g.glen").p, L.glen.data(), sizeof(uint64_t) * n_contig, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    HB_CUDA_OK(hb_pack_launch(ctx->buf("g.ascii").as<uint8_t>(), ctx->buf("g.src_off").as<uint64_t>(),
                              ctx->buf("g.goff").as<uint64_t>(), ctx->buf("g.glen").as<uint64_t>(), n_contig, L.tot,
                              ctx->buf("g.seq2").as<uint32_t>(), ctx->buf("g.nmask").as<uint32_t>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    ctx->buf("g.ascii").release();     // ASCII copy is no longer needed on the device
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    ctx->timing.k_ms[0] = ctx->timing.kernel_ms;
    ctx->timing.h2d_bytes = ascii_bytes;
    ctx->timing.launches = 1;
    ctx->have_genome = true;
    return HB_OK;
}

extern "C" int hb_genome_contig_rank(hb_ctx* ctx, const int32_t* rank, uint32_t n_contig) {
    if (!ctx || !rank) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_contig_rank: null argument");
    if (!ctx->have_genome) return hb_fail_msg(ctx, HB_ESTATE, "hb_genome_contig_rank: call hb_genome_load first");
    if (n_contig != ctx->g_ncontig) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_contig_rank: n_contig differs from the loaded genome");
    std::vector<uint8_t> seen(n_contig, 0);
    for (uint32_t c = 0; c < n_contig; ++c) {
        if (rank[c] < 0 || (uint32_t)rank[c] >= n_contig || seen[(size_t)rank[c]])
            return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_contig_rank: rank must be a permutation of 0..n_contig-1");
        seen[(size_t)rank[c]] = 1;
    }
    cudaSetDevice(ctx->device);
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("g.crank").p, rank, sizeof(int32_t) * n_contig, cudaMemcpyHostToDevice, ctx->stream));
    HB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

extern "C" int hb_index_build(hb_ctx* ctx, int k, int w) {
    if (!ctx) return HB_EINVAL;
    if (!ctx->have_genome) return hb_fail_msg(ctx, HB_ESTATE, "hb_index_build: call hb_genome_load first");
    if (k < 8 || k > 15 || w < 1 || w > 16) return hb_fail_msg(ctx, HB_EINVAL, "hb_index_build: need 8<=k<=15, 1<=w<=16");
    cudaSetDevice(ctx->device);
    ctx->k = k; ctx->w = w; ctx->have_index = false;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    cudaStream_t s = ctx->stream;
    const uint64_t tot = ctx->g_tot;
    const uint32_t ctas = hb_minimizer_ctas(tot);
    ENSURE("g.mm_cnt", sizeof(uint32_t) * (ctas + 1));
    ENSURE("g.mm_base", sizeof(uint32_t) * (ctas + 1));
    ENSURE("g.scan_tmp", sizeof(uint32_t) * hb_scan_tmp_words(std::max<uint64_t>(ctas, (1u << HB_SEED_BUCKET_BITS) + 1)));
    ENSURE("g.total", sizeof(uint64_t));
    const uint32_t nb = 1u << HB_SEED_BUCKET_BITS;
    ENSURE("g.bucket", sizeof(uint32_t) * (nb + 1));
    // events bracket kernels only: buffer (re)allocation sits between the intervals, outside every stage time
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    HB_CUDA_OK(hb_minimizer_count_launch(ctx->buf("g.seq2").as<uint32_t>(), ctx->buf("g.nmask").as<uint32_t>(), tot, k, w,
                                         ctx->buf("g.mm_cnt").as<uint32_t>(), s));
    HB_CUDA_OK(hb_exclusive_scan_u32(ctx->buf("g.mm_cnt").as<uint32_t>(), ctx->buf("g.mm_base").as<uint32_t>(), ctas,
                                     ctx->buf("g.scan_tmp").as<uint32_t>(), ctx->buf("g.total").as<uint64_t>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    uint64_t n_mm = 0;
    HB_CUDA_OK(cudaMemcpyAsync(&n_mm, ctx->buf("g.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    ENSURE("g.idx_a", sizeof(uint64_t) * (n_mm + 1));
    ENSURE("g.idx_b", sizeof(uint64_t) * (n_mm + 1));
    ENSURE("g.sort_tmp", sizeof(uint32_t) * hb_sort_tmp_words(n_mm));
    ENSURE("g.mm_pos", sizeof(uint64_t) * (n_mm + 1));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    HB_CUDA_OK(hb_minimizer_write_launch(ctx->buf("g.seq2").as<uint32_t>(), ctx->buf("g.nmask").as<uint32_t>(), tot, k, w,
                                         ctx->buf("g.mm_base").as<uint32_t>(), ctx->buf("g.idx_a").as<uint64_t>(), s));
    // the position-ordered list is kept for the coarse-boundary self-join (hb_coarse.cu)
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("g.mm_pos").p, ctx->buf("g.idx_a").p, sizeof(uint64_t) * n_mm, cudaMemcpyDeviceToDevice, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[3], s));
    // stable sort on the 2k-bit hash (keys are emitted in position order, so (hash,pos) ends sorted)
    uint64_t* sorted = nullptr;
    uint32_t launches = 6;
    HB_CUDA_OK(hb_radix_sort_u64(ctx->buf("g.idx_a").as<uint64_t>(), ctx->buf("g.idx_b").as<uint64_t>(), n_mm, 33, 33 + 2 * k,
                                 ctx->buf("g.sort_tmp").as<uint32_t>(), &sorted, &launches, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[4], s));
    // bucket table on the top HB_SEED_BUCKET_BITS bits of a 30-bit hash (hash < 4^k <= 2^30)
    HB_CUDA_OK(hb_bucket_bounds_launch(sorted, n_mm, 33 + (30 - HB_SEED_BUCKET_BITS), nb, ctx->buf("g.bucket").as<uint32_t>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[5], s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    if (sorted != ctx->buf("g.idx_a").as<uint64_t>()) std::swap(ctx->bufs["g.idx_a"], ctx->bufs["g.idx_b"]);
    // idx_b is kept: a rebuild (after hb_genome_mask, or the next chunk) allocates nothing
    ctx->g_nmm = n_mm;
    float t_count = 0.f, t_write = 0.f;
    cudaEventElapsedTime(&t_count, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&t_write, ctx->ev[2], ctx->ev[3]);
    ctx->timing.k_ms[0] = t_count + t_write;                              // sketch (count pass + write pass)
    cudaEventElapsedTime(&ctx->timing.k_ms[1], ctx->ev[3], ctx->ev[4]);   // radix sort
    cudaEventElapsedTime(&ctx->timing.k_ms[2], ctx->ev[4], ctx->ev[5]);   // bucket table
    ctx->timing.kernel_ms = ctx->timing.k_ms[0] + ctx->timing.k_ms[1] + ctx->timing.k_ms[2];
    ctx->timing.launches = launches + 4;
    ctx->have_index = true;
    return HB_OK;
}

extern "C" int hb_genome_info(hb_ctx* ctx, uint64_t* total_padded, uint64_t* n_minimizers, uint64_t* contig_goff,
                              uint32_t cap) {
    if (!ctx || !ctx->have_genome) return hb_fail_msg(ctx, HB_ESTATE, "hb_genome_info: no genome loaded");
    if (total_padded) *total_padded = ctx->g_tot;
    if (n_minimizers) *n_minimizers = ctx->have_index ? ctx->g_nmm : 0;
    if (contig_goff) for (uint32_t c = 0; c < ctx->g_ncontig && c < cap; ++c) contig_goff[c] = ctx->g_goff[c];
    return HB_OK;
}

namespace {
__global__ void fill_u32_kernel(uint32_t* __restrict__ a, uint64_t n, uint32_t v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
}  // namespace

// Test hook: exclusive scan of n copies of `value`; *total must be the exact 64-bit n * value even where the 32-bit
// offsets have wrapped (every size guard of hb_adjust / hb_coarse_boundary relies on that), *last = the last offset.
extern "C" int hb_debug_scan_total(hb_ctx* ctx, uint64_t n, uint32_t value, uint64_t* total, uint32_t* last) {
    if (!ctx || !total || n == 0 || n > (1ull << 28)) return hb_fail_msg(ctx, HB_EINVAL, "hb_debug_scan_total: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ENSURE("d.in", sizeof(uint32_t) * n); ENSURE("d.out", sizeof(uint32_t) * n);
    ENSURE("d.tmp", sizeof(uint32_t) * hb_scan_tmp_words(n)); ENSURE("d.total", sizeof(uint64_t));
    fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ctx->buf("d.in").as<uint32_t>(), n, value);
    HB_CUDA_OK(hb_exclusive_scan_u32(ctx->buf("d.in").as<uint32_t>(), ctx->buf("d.out").as<uint32_t>(), n,
                                     ctx->buf("d.tmp").as<uint32_t>(), ctx->buf("d.total").as<uint64_t>(), s));
    HB_CUDA_OK(cudaMemcpyAsync(total, ctx->buf("d.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    if (last) HB_CUDA_OK(cudaMemcpyAsync(last, ctx->buf("d.out").as<uint32_t>() + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    ctx->buf("d.in").release(); ctx->buf("d.out").release();
    return HB_OK;
}

extern "C" int hb_debug_index(hb_ctx* ctx, uint64_t* keys_out, uint64_t cap) {
    if (!ctx || !ctx->have_index) return hb_fail_msg(ctx, HB_ESTATE, "hb_debug_index: no index built");
    if (cap < ctx->g_nmm) return hb_fail_msg(ctx, HB_EINVAL, "hb_debug_index: buffer too small");
    cudaSetDevice(ctx->device);
    HB_CUDA_OK(cudaMemcpyAsync(keys_out, ctx->buf("g.idx_a").p, sizeof(uint64_t) * ctx->g_nmm, cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// One batch of candidates through the whole path on one lane (stream + buffer namespace).  The
// consensus bytes land in `h_cons` (pinned), calls[i].cons_off indexes them.
static int adjust_batch(hb_ctx* ctx, HbLane& ln, HbHostBuf& h_cons, uint32_t n_cand, const uint8_t* cand_ascii,
                        const uint64_t* cand_off, const hb_params* prm, int any_long, hb_call* calls,
                        uint64_t* cons_bytes, std::vector<hb_qhsp>* hsp_out = nullptr) {
    cudaSetDevice(ctx->device);
    cudaStream_t s = ln.stream;
    memset(&ln.timing, 0, sizeof(ln.timing));
    hb_timing& tm = ln.timing;
    auto LB = [&](const char* name) -> HbDevBuf& { return ln.prefix.empty() ? ctx->buf(name) : ctx->buf(ln.prefix + name); };
    const int max_loci = prm->max_loci, flank = prm->flank;
    const uint64_t tot = ctx->g_tot;
    uint32_t launches = 0;

    const bool dbg = getenv("HB_DEBUG_TIMING") != nullptr;
    auto t_host0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!dbg) return;
        cudaStreamSynchronize(s);
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[hb_adjust] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_host0).count());
    };
    // ---- S3: pack + sketch the candidates ----------------------------------------------------
    HB_CUDA_OK(cudaEventRecord(ln.ev[0], s));
    Layout CL = make_layout(cand_off, n_cand);
    uint64_t n_cm = 0;
    int rc = sketch(ctx, (ln.prefix + "c.").c_str(), cand_ascii, cand_off, n_cand, CL, &n_cm, s);
    if (rc != HB_OK) return rc;
    launches += 6;
    ENSURE_L("c.clen", sizeof(uint32_t) * (n_cand + 1));
    clen_kernel<<<(n_cand + 255) / 256, 256, 0, s>>>(LB("c.src_off").as<uint64_t>(), n_cand, LB("c.clen").as<uint32_t>());
    ++launches;
    tm.h2d_bytes += cand_off[n_cand] + sizeof(uint64_t) * (3 * (uint64_t)n_cand + 1);
    HB_CUDA_OK(cudaEventRecord(ln.ev[1], s));
    mark("sketch");
    const uint64_t* d_cgoff = LB("c.goff").as<uint64_t>();
    const uint64_t* d_coff = LB("c.src_off").as<uint64_t>();
    const uint32_t* d_clen = LB("c.clen").as<uint32_t>();

    // ---- S4: hits ------------------------------------------------------------------------------
    uint64_t n_hits = 0;
    if (n_cm > 0) {
        ENSURE_L("s.m_cand", sizeof(uint32_t) * n_cm); ENSURE_L("s.m_qidx", sizeof(uint32_t) * n_cm);
        ENSURE_L("s.m_lo", sizeof(uint32_t) * n_cm); ENSURE_L("s.m_cnt", sizeof(uint32_t) * n_cm);
        ENSURE_L("s.m_off", sizeof(uint32_t) * n_cm);
        ENSURE_L("s.scan_tmp", sizeof(uint32_t) * hb_scan_tmp_words(std::max<uint64_t>(n_cm, n_cand + 1)));
        ENSURE_L("s.total", sizeof(uint64_t));
        HB_CUDA_OK(hb_seed_count_launch(LB("c.mm").as<uint64_t>(), (uint32_t)n_cm, d_cgoff, n_cand,
                                        ctx->buf("g.idx_a").as<uint64_t>(), ctx->g_nmm, ctx->buf("g.bucket").as<uint32_t>(),
                                        prm->occ_cap, LB("s.m_cand").as<uint32_t>(), LB("s.m_qidx").as<uint32_t>(),
                                        LB("s.m_lo").as<uint32_t>(), LB("s.m_cnt").as<uint32_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(LB("s.m_cnt").as<uint32_t>(), LB("s.m_off").as<uint32_t>(), n_cm,
                                         LB("s.scan_tmp").as<uint32_t>(), LB("s.total").as<uint64_t>(), s));
        HB_CUDA_OK(cudaMemcpyAsync(&n_hits, LB("s.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        launches += 4;
    }
    if (n_hits >= (1ull << 31)) return hb_fail_msg(ctx, HB_ENOMEM, "hb_adjust: too many seed hits in one call; pass fewer candidates");
    uint32_t n_cl = 0, n_pairs = 0;
    ENSURE_L("s.n_loci", sizeof(uint32_t) * (n_cand + 1));
    ENSURE_L("s.pair_off", sizeof(uint32_t) * (n_cand + 2));
    HB_CUDA_OK(cudaMemsetAsync(LB("s.n_loci").p, 0, sizeof(uint32_t) * (n_cand + 1), s));
    uint64_t* d_hits = nullptr;
    if (n_hits > 0) {
        ENSURE_L("s.hits_a", sizeof(uint64_t) * n_hits); ENSURE_L("s.hits_b", sizeof(uint64_t) * n_hits);
        ENSURE_L("s.sort_tmp", sizeof(uint32_t) * hb_sort_tmp_words(n_hits));
        HB_CUDA_OK(hb_seed_fill_launch(LB("c.mm").as<uint64_t>(), (uint32_t)n_cm, d_cgoff, ctx->buf("g.idx_a").as<uint64_t>(),
                                       tot, ctx->k, LB("s.m_cand").as<uint32_t>(), LB("s.m_qidx").as<uint32_t>(),
                                       LB("s.m_lo").as<uint32_t>(), LB("s.m_cnt").as<uint32_t>(),
                                       LB("s.m_off").as<uint32_t>(), LB("s.hits_a").as<uint64_t>(), s));
        ++launches;
        int cand_bits = 1;
        while ((1u << cand_bits) < n_cand) ++cand_bits;
        // total order on (cand, strand, diagonal); the low 13 bits (minimizer ordinal) do not affect
        // any cluster statistic, so they are left unsorted
        HB_CUDA_OK(hb_radix_sort_u64(LB("s.hits_a").as<uint64_t>(), LB("s.hits_b").as<uint64_t>(), n_hits, 13,
                                     47 + cand_bits, LB("s.sort_tmp").as<uint32_t>(), &d_hits, &launches, s));
        // ---- S5: clusters -----------------------------------------------------------------------
        ENSURE_L("s.flag", sizeof(uint32_t) * n_hits); ENSURE_L("s.cid", sizeof(uint32_t) * n_hits);
        ENSURE_L("s.flag2", sizeof(uint32_t) * n_hits);
        ENSURE_L("s.scan_tmp2", sizeof(uint32_t) * hb_scan_tmp_words(n_hits));
        ENSURE_L("s.cl_start", sizeof(uint32_t) * (n_hits + 1));
        HB_CUDA_OK(hb_cluster_flag_launch(d_hits, n_hits, LB("s.flag").as<uint32_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(LB("s.flag").as<uint32_t>(), LB("s.cid").as<uint32_t>(), n_hits,
                                         LB("s.scan_tmp2").as<uint32_t>(), LB("s.total").as<uint64_t>(), s));
        uint64_t ncl64 = 0;
        HB_CUDA_OK(cudaMemcpyAsync(&ncl64, LB("s.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(hb_cluster_start_launch(LB("s.flag").as<uint32_t>(), LB("s.cid").as<uint32_t>(), n_hits,
                                           LB("s.cl_start").as<uint32_t>(), s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        // S5b: clusters wider than the band are cut into pieces (second flag pass + rescan)
        HB_CUDA_OK(hb_cluster_split_launch(d_hits, n_hits, LB("s.flag").as<uint32_t>(), LB("s.cid").as<uint32_t>(),
                                           LB("s.cl_start").as<uint32_t>(), (uint32_t)ncl64, LB("s.flag2").as<uint32_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(LB("s.flag2").as<uint32_t>(), LB("s.cid").as<uint32_t>(), n_hits,
                                         LB("s.scan_tmp2").as<uint32_t>(), LB("s.total").as<uint64_t>(), s));
        HB_CUDA_OK(cudaMemcpyAsync(&ncl64, LB("s.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(hb_cluster_start_launch(LB("s.flag2").as<uint32_t>(), LB("s.cid").as<uint32_t>(), n_hits,
                                           LB("s.cl_start").as<uint32_t>(), s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        n_cl = (uint32_t)ncl64;
        launches += 6;
        ENSURE_L("s.cand_first", sizeof(uint32_t) * (n_cand + 1));
        ENSURE_L("s.loci", sizeof(HbLocus) * (size_t)n_cand * max_loci);
        HB_CUDA_OK(hb_rank_select_launch(d_hits, n_hits, LB("s.cl_start").as<uint32_t>(), n_cl,
                                         LB("s.cand_first").as<uint32_t>(), n_cand, max_loci,
                                         LB("s.loci").as<HbLocus>(), LB("s.n_loci").as<uint32_t>(), s));
        launches += 8;
    }
    // ---- S6: pair list ---------------------------------------------------------------------------
    ENSURE_L("s.scan_tmp", sizeof(uint32_t) * hb_scan_tmp_words(std::max<uint64_t>(n_cm, n_cand + 1)));
    ENSURE_L("s.total", sizeof(uint64_t));
    HB_CUDA_OK(hb_exclusive_scan_u32(LB("s.n_loci").as<uint32_t>(), LB("s.pair_off").as<uint32_t>(), n_cand + 1,
                                     LB("s.scan_tmp").as<uint32_t>(), LB("s.total").as<uint64_t>(), s));
    {
        uint64_t np64 = 0;
        HB_CUDA_OK(cudaMemcpyAsync(&np64, LB("s.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        n_pairs = (uint32_t)np64;
    }
    launches += 3;
    ENSURE_L("p.pairs", sizeof(HbLocus) * ((size_t)n_pairs + 1));
    ENSURE_L("p.alns", sizeof(HbAln) * ((size_t)n_pairs + 1));
    ENSURE_L("p.tb_units", sizeof(uint32_t) * ((size_t)n_pairs + 1)); ENSURE_L("p.tb_off", sizeof(uint32_t) * ((size_t)n_pairs + 1));
    ENSURE_L("p.qr_units", sizeof(uint32_t) * ((size_t)n_pairs + 1)); ENSURE_L("p.qr_off", sizeof(uint32_t) * ((size_t)n_pairs + 1));
    ENSURE_L("p.scan_tmp", sizeof(uint32_t) * hb_scan_tmp_words((size_t)n_pairs + 1));
    ENSURE_L("p.total", sizeof(uint64_t) * 2);
    uint64_t tb_units_total = 0, qr_units_total = 0;
    std::vector<uint32_t> h_tb_off;
    if (n_pairs > 0) {
        HB_CUDA_OK(hb_pair_emit_launch(LB("s.loci").as<HbLocus>(), LB("s.n_loci").as<uint32_t>(),
                                       LB("s.pair_off").as<uint32_t>(), n_cand, max_loci, LB("p.pairs").as<HbLocus>(), s));
        HB_CUDA_OK(hb_pair_sizes_launch(LB("p.pairs").as<HbLocus>(), n_pairs, d_clen, LB("p.tb_units").as<uint32_t>(),
                                        LB("p.qr_units").as<uint32_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(LB("p.tb_units").as<uint32_t>(), LB("p.tb_off").as<uint32_t>(), n_pairs,
                                         LB("p.scan_tmp").as<uint32_t>(), LB("p.total").as<uint64_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(LB("p.qr_units").as<uint32_t>(), LB("p.qr_off").as<uint32_t>(), n_pairs,
                                         LB("p.scan_tmp").as<uint32_t>(), LB("p.total").as<uint64_t>() + 1, s));
        uint64_t totals[2] = {0, 0};
        h_tb_off.resize(n_pairs);
        HB_CUDA_OK(cudaMemcpyAsync(totals, LB("p.total").p, sizeof(uint64_t) * 2, cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaMemcpyAsync(h_tb_off.data(), LB("p.tb_off").p, sizeof(uint32_t) * n_pairs, cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        tb_units_total = totals[0]; qr_units_total = totals[1];
        launches += 8;
    }
    HB_CUDA_OK(cudaEventRecord(ln.ev[2], s));
    mark("seed/sort/cluster/pairs");

    // ---- S7: banded extension, chunked by the traceback budget -----------------------------------
    uint64_t budget_units = (prm->tb_budget_bytes ? prm->tb_budget_bytes : (16ull << 30)) >> 9;      // 512 B units
    if (budget_units < 64) budget_units = 64;
    uint64_t ext_steps = 0;
    if (n_pairs > 0) {
        const uint64_t alloc_units = std::min<uint64_t>(tb_units_total, budget_units);
        ENSURE_L("p.tbuf", (alloc_units + 1) << 9);
        ENSURE_L("p.qrec", ((qr_units_total + 1) << 5) * sizeof(uint32_t));
        ENSURE_L("p.order", hb_extend_order_words(n_pairs) * sizeof(uint32_t));
        uint32_t lo = 0;
        while (lo < n_pairs) {
            uint32_t hi = lo + 1;
            const uint64_t base = h_tb_off[lo];
            // largest hi with tb_off[hi] - base <= alloc_units (tb_off is monotone)
            {
                uint32_t a = lo + 1, b = n_pairs;
                while (a < b) {
                    uint32_t mid = a + (b - a + 1) / 2;
                    const uint64_t end = (mid < n_pairs) ? h_tb_off[mid] : tb_units_total;
                    if (end - base <= alloc_units) a = mid; else b = mid - 1;
                }
                hi = a;
                const uint64_t end = (hi < n_pairs) ? h_tb_off[hi] : tb_units_total;
                if (end - base > alloc_units) return hb_fail_msg(ctx, HB_ENOMEM, "hb_adjust: tb_budget_bytes is smaller than one candidate's traceback");
            }
            HB_CUDA_OK(hb_extend_launch(LB("p.pairs").as<HbLocus>(), lo, hi, ctx->buf("g.seq2").as<uint32_t>(),
                                        ctx->buf("g.nmask").as<uint32_t>(), tot, LB("c.seq2").as<uint32_t>(),
                                        LB("c.nmask").as<uint32_t>(), d_cgoff, d_clen, LB("p.tb_off").as<uint32_t>(),
                                        (uint32_t)base, LB("p.tbuf").as<uint32_t>(), LB("p.qr_off").as<uint32_t>(),
                                        LB("p.qrec").as<uint32_t>(), LB("p.alns").as<HbAln>(),
                                        LB("p.order").as<uint32_t>(), any_long, &launches, s));
            lo = hi;
        }
        ext_steps = tb_units_total * 32;                 // one 512 B unit = 64 anti-diagonals of 16 lanes
    }
    HB_CUDA_OK(cudaEventRecord(ln.ev[3], s));
    mark("extend");

    if (hsp_out) {
        // ---- hb_query_hsps: the local alignments themselves are the result (no copy rules, no MSA, no judge) ----
        ENSURE_L("h.recs", sizeof(hb_qhsp) * ((size_t)n_pairs + 1));
        HB_CUDA_OK(hb_qhsp_launch(LB("p.pairs").as<HbLocus>(), n_pairs, LB("p.alns").as<HbAln>(), LB("p.qr_off").as<uint32_t>(),
                                  LB("p.qrec").as<uint32_t>(), LB("c.seq2").as<uint32_t>(), LB("c.nmask").as<uint32_t>(), d_cgoff,
                                  ctx->buf("g.seq2").as<uint32_t>(), ctx->buf("g.nmask").as<uint32_t>(),
                                  ctx->buf("g.goff").as<uint64_t>(), ctx->buf("g.glen").as<uint64_t>(), ctx->g_ncontig, tot,
                                  LB("h.recs").as<hb_qhsp>(), s));
        ++launches;
        HB_CUDA_OK(cudaEventRecord(ln.ev[4], s));
        std::vector<hb_qhsp> all(n_pairs);
        if (n_pairs)
            HB_CUDA_OK(cudaMemcpyAsync(all.data(), LB("h.recs").p, sizeof(hb_qhsp) * n_pairs, cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaEventRecord(ln.ev[5], s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        hsp_out->clear();
        for (const hb_qhsp& r : all) if (r.score > 0 && r.contig >= 0) hsp_out->push_back(r);
        cudaEventElapsedTime(&tm.k_ms[0], ln.ev[0], ln.ev[1]);
        cudaEventElapsedTime(&tm.k_ms[1], ln.ev[1], ln.ev[2]);
        cudaEventElapsedTime(&tm.k_ms[2], ln.ev[2], ln.ev[3]);
        cudaEventElapsedTime(&tm.k_ms[3], ln.ev[3], ln.ev[4]);
        cudaEventElapsedTime(&tm.d2h_ms, ln.ev[4], ln.ev[5]);
        cudaEventElapsedTime(&tm.kernel_ms, ln.ev[0], ln.ev[4]);
        tm.d2h_bytes += sizeof(hb_qhsp) * (uint64_t)n_pairs;
        tm.launches = launches;
        tm.k_ms[6] = (float)n_pairs;
        tm.k_ms[7] = (float)((double)ext_steps * 1e-6);
        ln.n_pairs = n_pairs; ln.ext_steps = ext_steps;
        return HB_OK;
    }

    // ---- S8..S10: copy rules, star MSA -------------------------------------------------------------
    ENSURE_L("m.n_rows", sizeof(uint32_t) * (n_cand + 1)); ENSURE_L("m.n_cols", sizeof(uint32_t) * (n_cand + 1));
    ENSURE_L("m.n_trows", sizeof(uint32_t) * (n_cand + 1));
    HB_CUDA_OK(cudaMemsetAsync(LB("m.n_trows").p, 0, sizeof(uint32_t) * (n_cand + 1), s));
    const int trunc_gate = prm->trunc_gate ? 1 : 0;
    ENSURE_L("m.ins", sizeof(uint32_t) * (cand_off[n_cand] + 1));
    HB_CUDA_OK(cudaMemsetAsync(LB("m.ins").p, 0, sizeof(uint32_t) * (cand_off[n_cand] + 1), s));
    HB_CUDA_OK(cudaMemsetAsync(LB("m.n_rows").p, 0, sizeof(uint32_t) * (n_cand + 1), s));
    if (n_pairs > 0)
        HB_CUDA_OK(hb_chain_launch(LB("p.pairs").as<HbLocus>(), LB("s.pair_off").as<uint32_t>(), n_cand, d_clen,
                                   LB("p.qr_off").as<uint32_t>(), LB("p.qrec").as<uint32_t>(), LB("p.alns").as<HbAln>(), s));
    HB_CUDA_OK(hb_copy_filter_launch(LB("p.pairs").as<HbLocus>(), n_pairs, d_clen, ctx->buf("g.goff").as<uint64_t>(),
                                     ctx->buf("g.glen").as<uint64_t>(), ctx->g_ncontig, tot, flank, LB("p.alns").as<HbAln>(), s));
    HB_CUDA_OK(hb_copy_select_launch(LB("p.pairs").as<HbLocus>(), LB("s.pair_off").as<uint32_t>(), n_cand,
                                     prm->max_rows, trunc_gate, (prm->flags & HB_PRM_ROWS_FIRST) ? 1 : 0,
                                     ctx->buf("g.crank").as<int32_t>(), LB("p.alns").as<HbAln>(),
                                     LB("m.n_rows").as<uint32_t>(), LB("m.n_trows").as<uint32_t>(), s));
    HB_CUDA_OK(hb_ins_max_launch(LB("p.pairs").as<HbLocus>(), n_pairs, LB("p.alns").as<HbAln>(),
                                 LB("p.qr_off").as<uint32_t>(), LB("p.qrec").as<uint32_t>(), d_coff,
                                 LB("m.ins").as<uint32_t>(), s));
    HB_CUDA_OK(hb_layout_launch(d_coff, d_clen, n_cand, flank, LB("m.ins").as<uint32_t>(), LB("m.n_cols").as<uint32_t>(), s));
    // column offsets of the per-candidate column maps (device scan: no host round trip)
    ENSURE_L("m.col_off", sizeof(uint32_t) * (n_cand + 1));
    ENSURE_L("m.col_total", sizeof(uint64_t));
    HB_CUDA_OK(hb_exclusive_scan_u32(LB("m.n_cols").as<uint32_t>(), LB("m.col_off").as<uint32_t>(), n_cand,
                                     LB("s.scan_tmp").as<uint32_t>(), LB("m.col_total").as<uint64_t>(), s));
    launches += 7;
    std::vector<uint32_t> h_rows(n_cand), h_cols(n_cand), h_trows(n_cand);
    HB_CUDA_OK(cudaMemcpyAsync(h_trows.data(), LB("m.n_trows").p, sizeof(uint32_t) * n_cand, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaMemcpyAsync(h_rows.data(), LB("m.n_rows").p, sizeof(uint32_t) * n_cand, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaMemcpyAsync(h_cols.data(), LB("m.n_cols").p, sizeof(uint32_t) * n_cand, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    // candidates with >= 1 row become judge tasks (0 rows: no member file in the reference either,
    // flank_region_align_v5 only submits queries present in extend_all_copies, Util.py:8125-8137)
    std::vector<uint64_t> h_mat_off(n_cand + 1, 0), h_tmat_off(n_cand + 1, 0);
    std::vector<HbVoteTask> tasks;
    std::vector<uint32_t> task_cand;          // candidate of each task; trunc tasks follow the full ones
    std::vector<uint32_t> trunc_to_full, full_of_cand(n_cand, 0);
    tasks.reserve(2 * (size_t)n_cand); task_cand.reserve(2 * (size_t)n_cand);
    uint64_t mat_total = 0, col_total = 0, match_total = 0, row_total = 0;
    auto add_task = [&](uint32_t c, uint32_t rows, uint64_t off) {
        HbVoteTask t;
        memset(&t, 0, sizeof(t));
        t.raw_off = off; t.clean_off = off;
        t.plane_off = col_total * 24; t.cons_off = col_total + 2ull * tasks.size(); t.match_off = match_total;
        t.anchor_off = 2 * row_total;
        t.cand_off = cand_off[c]; t.cand_len = (int32_t)(cand_off[c + 1] - cand_off[c]);
        t.R = (int32_t)rows; t.C = (int32_t)h_cols[c];
        t.te_type = prm->te_type; t.plant = prm->plant; t.result_type = prm->result_type; t.raw_is_ascii = 0;
        tasks.push_back(t); task_cand.push_back(c);
        col_total += h_cols[c];
        match_total += hb_vote_match_entries((int)h_cols[c]);
        row_total += rows;
    };
    for (uint32_t c = 0; c < n_cand; ++c) {
        h_mat_off[c] = mat_total;
        if (h_rows[c] == 0) continue;
        full_of_cand[c] = (uint32_t)tasks.size();
        add_task(c, h_rows[c], mat_total);
        mat_total += (uint64_t)h_rows[c] * h_cols[c];
    }
    h_mat_off[n_cand] = mat_total;
    const uint32_t n_full = (uint32_t)tasks.size();
    const uint64_t tmat_base = mat_total;      // truncated matrices live after the full ones in the same buffer
    for (uint32_t c = 0; c < n_cand; ++c) {
        h_tmat_off[c] = mat_total;
        if (h_rows[c] == 0 || h_trows[c] == 0) continue;
        trunc_to_full.push_back(full_of_cand[c]);
        add_task(c, h_trows[c], mat_total);
        mat_total += (uint64_t)h_trows[c] * h_cols[c];
    }
    h_tmat_off[n_cand] = mat_total;
    (void)tmat_base;
    const uint32_t n_tasks = (uint32_t)tasks.size();
    ENSURE_L("m.mat_off", sizeof(uint64_t) * (n_cand + 1));
    ENSURE_L("m.tmat_off", sizeof(uint64_t) * (n_cand + 1));
    HB_CUDA_OK(cudaMemcpyAsync(LB("m.mat_off").p, h_mat_off.data(), sizeof(uint64_t) * (n_cand + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(LB("m.tmat_off").p, h_tmat_off.data(), sizeof(uint64_t) * (n_cand + 1), cudaMemcpyHostToDevice, s));
    {
        uint64_t all_cols = 0;
        for (uint32_t c = 0; c < n_cand; ++c) all_cols += h_cols[c];
        if (all_cols >= (1ull << 32)) return hb_fail_msg(ctx, HB_ENOMEM, "hb_adjust: too many MSA columns in one call; pass fewer candidates");
        ENSURE_L("m.colq", sizeof(uint32_t) * (all_cols + 1));
        HB_CUDA_OK(hb_colmap_launch(d_coff, d_clen, n_cand, flank, LB("m.ins").as<uint32_t>(), LB("m.n_cols").as<uint32_t>(),
                                    LB("m.col_off").as<uint32_t>(), LB("m.colq").as<uint32_t>(), s));
        ++launches;
    }
    // the raw matrices are not written on the path: msa_clean_kernel below goes from the edit records straight to the clean
    // matrices.  hb_last_msa / hb_last_msa_trunc (parity getters, the FiLTR mode) materialise them on demand.
    ln.raw_valid = false; ln.raw_bytes = mat_total; ln.raw_flank = flank; ln.raw_has_trunc = n_tasks > n_full;
    HB_CUDA_OK(cudaEventRecord(ln.ev[4], s));
    mark("copy rules + msa");
    ln.tmat_off = h_tmat_off; ln.trows = h_trows;
    if (prm->flags & HB_PRM_NO_JUDGE) {
        // the FiLTR fork of the loop stops at the alignment (bin/FiLTR-main/src/Util.py:1315-1339): the matrices stay on the
        // device for hb_last_msa / hb_last_msa_trunc, the copy records for hb_last_copies
        HB_CUDA_OK(cudaStreamSynchronize(s));
        for (uint32_t c = 0; c < n_cand; ++c) {
            hb_call& o = calls[c];
            memset(&o, 0, sizeof(o));
            o.align_start = o.align_end = o.homo_start = o.homo_end = -1;
            o.copy_num = (int32_t)h_rows[c];
            o.flags = h_rows[c] ? 0 : HB_FLAG_NO_COPIES;
        }
        if (cons_bytes) *cons_bytes = 0;
        ln.cons_bytes = 0;
        ln.n_pairs = n_pairs; ln.ext_steps = ext_steps;
        ln.mat_off = h_mat_off; ln.rows = h_rows; ln.cols = h_cols;
        cudaEventElapsedTime(&tm.k_ms[0], ln.ev[0], ln.ev[1]);
        cudaEventElapsedTime(&tm.k_ms[1], ln.ev[1], ln.ev[2]);
        cudaEventElapsedTime(&tm.k_ms[2], ln.ev[2], ln.ev[3]);
        cudaEventElapsedTime(&tm.k_ms[3], ln.ev[3], ln.ev[4]);
        cudaEventElapsedTime(&tm.kernel_ms, ln.ev[0], ln.ev[4]);
        tm.launches = launches;
        tm.k_ms[6] = (float)n_pairs;
        tm.k_ms[7] = (float)((double)ext_steps * 1e-6);
        return HB_OK;
    }

    // ---- kernel (iii): clean + judge ---------------------------------------------------------------
    std::vector<HbVoteCall> h_calls(n_tasks);
    uint64_t dense_total = 0;
    ln.cons_bytes = 0;
    if (n_tasks > 0) {
        ENSURE_L("v.tasks", sizeof(HbVoteTask) * n_tasks);
        ENSURE_L("v.clean", mat_total + 1);
        ENSURE_L("v.planes", col_total * 24 * sizeof(uint32_t));
        ENSURE_L("v.matches", match_total * hb_vote_match_entry_size());
        ENSURE_L("v.cons", col_total + 2ull * n_tasks + 1);
        ENSURE_L("v.anchors", sizeof(int) * (2 * row_total + 2));
        ENSURE_L("v.calls", sizeof(HbVoteCall) * n_tasks);
        ENSURE_L("v.len", sizeof(uint32_t) * (n_tasks + 1)); ENSURE_L("v.off", sizeof(uint32_t) * (n_tasks + 1));
        ENSURE_L("v.dense", col_total + 2ull * n_tasks + 1);
        ENSURE_L("v.scan_tmp", sizeof(uint32_t) * hb_scan_tmp_words(n_tasks + 1));
        ENSURE_L("v.total", sizeof(uint64_t));
        ENSURE_L("v.task_cand", sizeof(uint32_t) * n_tasks);
        ENSURE_L("v.f2t", sizeof(int32_t) * (n_full + 1));
        std::vector<int32_t> trunc_of_full(n_full, -1);
        for (uint32_t k = 0; k < trunc_to_full.size(); ++k) trunc_of_full[trunc_to_full[k]] = (int32_t)(n_full + k);
        HB_CUDA_OK(cudaMemcpyAsync(LB("v.f2t").p, trunc_of_full.data(), sizeof(int32_t) * n_full, cudaMemcpyHostToDevice, s));
        // CTA per candidate: start the largest matrices first so that a long one is never the last to be scheduled
        ENSURE_L("v.order", sizeof(uint32_t) * (n_full + 1));
        std::vector<uint32_t> order(n_full);
        for (uint32_t k = 0; k < n_full; ++k) order[k] = k;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return (uint64_t)tasks[a].R * (uint64_t)tasks[a].C > (uint64_t)tasks[b].R * (uint64_t)tasks[b].C; });
        HB_CUDA_OK(cudaMemcpyAsync(LB("v.order").p, order.data(), sizeof(uint32_t) * n_full, cudaMemcpyHostToDevice, s));
        HB_CUDA_OK(cudaMemcpyAsync(LB("v.tasks").p, tasks.data(), sizeof(HbVoteTask) * n_tasks, cudaMemcpyHostToDevice, s));
        HB_CUDA_OK(cudaMemcpyAsync(LB("v.task_cand").p, task_cand.data(), sizeof(uint32_t) * n_tasks, cudaMemcpyHostToDevice, s));
        HB_CUDA_OK(cudaMemsetAsync(LB("v.calls").p, 0, sizeof(HbVoteCall) * n_tasks, s));
        HB_CUDA_OK(hb_msa_clean_launch(LB("v.tasks").as<HbVoteTask>(), LB("v.order").as<uint32_t>(), LB("v.task_cand").as<uint32_t>(),
                                       LB("v.f2t").as<int32_t>(), n_full,
                                       LB("p.pairs").as<HbLocus>(), LB("s.pair_off").as<uint32_t>(), LB("p.alns").as<HbAln>(),
                                       LB("p.qr_off").as<uint32_t>(), LB("p.qrec").as<uint32_t>(), LB("m.col_off").as<uint32_t>(),
                                       LB("m.colq").as<uint32_t>(), ctx->buf("g.seq2").as<uint32_t>(), ctx->buf("g.nmask").as<uint32_t>(),
                                       tot, flank, LB("v.clean").as<uint8_t>(),
                                       LB("v.planes").as<uint32_t>(),
                                       LB("v.calls").as<HbVoteCall>(), s));
        ++launches;
        HB_CUDA_OK(hb_vote_launch(LB("v.tasks").as<HbVoteTask>(), (int)n_tasks, nullptr,
                                  LB("v.clean").as<uint8_t>(), LB("v.planes").as<uint32_t>(), LB("v.matches").p,
                                  LB("v.anchors").as<int>(), LB("c.ascii").as<uint8_t>(),
                                  LB("v.cons").as<uint8_t>(), LB("v.calls").as<HbVoteCall>(), prm->te_type,
                                  (int)(row_total / n_tasks), s));
        HB_CUDA_OK(cudaEventRecord(ln.ev[5], s));
        if (n_tasks > n_full) {
            ENSURE_L("v.t2f", sizeof(uint32_t) * trunc_to_full.size());
            HB_CUDA_OK(cudaMemcpyAsync(LB("v.t2f").p, trunc_to_full.data(), sizeof(uint32_t) * trunc_to_full.size(), cudaMemcpyHostToDevice, s));
            gate_resolve_kernel<<<(n_tasks - n_full + 255) / 256, 256, 0, s>>>(LB("v.calls").as<HbVoteCall>(), n_full, n_tasks,
                                                                              LB("v.t2f").as<uint32_t>());
            ++launches;
        }
        u32_field_kernel<<<(n_tasks + 255) / 256, 256, 0, s>>>(LB("v.calls").as<HbVoteCall>(), n_tasks, LB("v.len").as<uint32_t>());
        HB_CUDA_OK(hb_exclusive_scan_u32(LB("v.len").as<uint32_t>(), LB("v.off").as<uint32_t>(), n_tasks,
                                         LB("v.scan_tmp").as<uint32_t>(), LB("v.total").as<uint64_t>(), s));
        cons_compact_kernel<<<(n_tasks * 32 + 255) / 256, 256, 0, s>>>(LB("v.calls").as<HbVoteCall>(), n_tasks,
                                                                       LB("v.off").as<uint32_t>(), LB("v.cons").as<uint8_t>(),
                                                                       LB("v.dense").as<uint8_t>());
        launches += 7;
        HB_CUDA_OK(cudaEventRecord(ln.ev[6], s));
        HB_CUDA_OK(cudaMemcpyAsync(&dense_total, LB("v.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaMemcpyAsync(h_calls.data(), LB("v.calls").p, sizeof(HbVoteCall) * n_full, cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        HB_CUDA_OK(h_cons.ensure(dense_total + 1));
        ln.cons_bytes = dense_total;
        if (dense_total)
            HB_CUDA_OK(cudaMemcpyAsync(h_cons.p, LB("v.dense").p, dense_total, cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaEventRecord(ln.ev[7], s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        tm.d2h_bytes += dense_total + sizeof(HbVoteCall) * n_full;
    } else {
        HB_CUDA_OK(cudaEventRecord(ln.ev[5], s));
        HB_CUDA_OK(cudaEventRecord(ln.ev[6], s));
        HB_CUDA_OK(cudaEventRecord(ln.ev[7], s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
    }
    mark("vote + d2h");
    // scatter results back to candidate order; candidates without copies: nothing found
    for (uint32_t c = 0; c < n_cand; ++c) {
        hb_call& o = calls[c];
        memset(&o, 0, sizeof(o));
        o.align_start = o.align_end = o.homo_start = o.homo_end = -1;
        o.flags = HB_FLAG_NO_COPIES;
    }
    for (uint32_t t = 0; t < n_full; ++t) memcpy(&calls[task_cand[t]], &h_calls[t], sizeof(hb_call));
    if (cons_bytes) *cons_bytes = dense_total;
    ln.n_pairs = n_pairs; ln.ext_steps = ext_steps;
    ln.mat_off = h_mat_off; ln.rows = h_rows; ln.cols = h_cols;
    cudaEventElapsedTime(&tm.k_ms[0], ln.ev[0], ln.ev[1]);   // pack + sketch candidates (incl. H2D)
    cudaEventElapsedTime(&tm.k_ms[1], ln.ev[1], ln.ev[2]);   // seed lookup, sort, cluster, rank
    cudaEventElapsedTime(&tm.k_ms[2], ln.ev[2], ln.ev[3]);   // banded extension + traceback
    cudaEventElapsedTime(&tm.k_ms[3], ln.ev[3], ln.ev[4]);   // copy rules + star MSA
    cudaEventElapsedTime(&tm.k_ms[4], ln.ev[4], ln.ev[5]);   // clean + judge
    cudaEventElapsedTime(&tm.k_ms[5], ln.ev[5], ln.ev[6]);   // consensus compaction
    cudaEventElapsedTime(&tm.d2h_ms, ln.ev[6], ln.ev[7]);
    cudaEventElapsedTime(&tm.kernel_ms, ln.ev[0], ln.ev[6]);
    tm.h2d_ms = 0.f;
    tm.launches = launches;
    tm.k_ms[6] = (float)n_pairs;                                  // diagnostics: pairs extended
    tm.k_ms[7] = (float)((double)ext_steps * 1e-6);               // diagnostics: million DP steps
    mark("done");
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int hb_adjust(hb_ctx* ctx, uint32_t n_cand, const uint8_t* cand_ascii, const uint64_t* cand_off,
                         const hb_params* prm, hb_call* calls, uint64_t* cons_bytes) {
    if (!ctx || !cand_ascii || !cand_off || !prm || !calls) return hb_fail_msg(ctx, HB_EINVAL, "hb_adjust: null argument");
    if (!ctx->have_index) return hb_fail_msg(ctx, HB_ESTATE, "hb_adjust: call hb_genome_load + hb_index_build first");
    if (n_cand == 0) { if (cons_bytes) *cons_bytes = 0; return HB_OK; }
    if (n_cand > (1u << 17)) return hb_fail_msg(ctx, HB_EINVAL, "hb_adjust: at most 131072 candidates per call");
    if (prm->te_type < HB_TE_TIR || prm->te_type > HB_TE_NON_LTR)
        return hb_fail_msg(ctx, HB_EINVAL, "hb_adjust: te_type must be HB_TE_TIR, HB_TE_HELITRON or HB_TE_NON_LTR");
    if (prm->max_loci < 1 || prm->max_loci > 512 || prm->max_rows < 2 || prm->max_rows > 100 || prm->occ_cap < 1 ||
        prm->flank < 0 || prm->flank > 200)
        return hb_fail_msg(ctx, HB_EINVAL, "hb_adjust: parameter out of range");
    int any_long = 0;                 // candidates beyond the 16-bit extension kernel (hb_extend16.cuh)
    for (uint32_t c = 0; c < n_cand; ++c) {
        const uint64_t L = cand_off[c + 1] - cand_off[c];
        if (L < 3 || L > 32000) return hb_fail_msg(ctx, HB_EINVAL, "hb_adjust: candidate " + std::to_string(c) + " length must be in [3, 32000]");
        if (L > HB_EXTEND16_MAX_L) any_long = 1;
    }
    cudaSetDevice(ctx->device);
    // ---- serial: one batch on the context's stream (the hb_last_copies / hb_last_msa getters read it) ----
    uint32_t n_parts = prm->overlap >= 2 ? (uint32_t)prm->overlap : 1u;
    if (n_parts > 16) n_parts = 16;
    if (n_cand < 64 * n_parts) n_parts = 1;
    ctx->cons_parts.clear();
    if (ctx->part_cons.size() < n_parts) ctx->part_cons.resize(n_parts);
    if (n_parts == 1) {
        HbLane& ln = ctx->lane[0];
        uint64_t nb = 0;
        const int rc = adjust_batch(ctx, ln, ctx->part_cons[0], n_cand, cand_ascii, cand_off, prm, any_long, calls, &nb);
        if (rc != HB_OK) return rc;
        ctx->timing = ln.timing;
        ctx->last_overlapped = false;
        ctx->last_cons_bytes = nb;
        ctx->cons_parts.push_back({(const uint8_t*)ctx->part_cons[0].p, nb});
        ctx->last_ncand = n_cand; ctx->last_npairs = ln.n_pairs;
        ctx->last_mat_off = ln.mat_off; ctx->last_rows = ln.rows; ctx->last_cols = ln.cols;
        ctx->last_tmat_off = ln.tmat_off; ctx->last_trows = ln.trows;
        if (cons_bytes) *cons_bytes = nb;
        return HB_OK;
    }
    // ---- overlapped: n_parts sub-batches (contiguous candidate ranges of about equal bases) dealt to two
    // lanes, each driven by its own host thread on its own stream.  The extension kernel saturates the
    // integer ALU pipe while sort / traceback / MSA / judge are latency- or memory-bound, so sub-batches in
    // different stages fill each other's idle issue slots; the host-side gaps (counts read back, task
    // tables) of one lane hide behind the other lane's kernels.  Results do not depend on the split.
    std::vector<uint32_t> cut(n_parts + 1, 0);
    {
        const uint64_t total = cand_off[n_cand] - cand_off[0];
        uint32_t c = 0;
        for (uint32_t k = 1; k < n_parts; ++k) {
            const uint64_t want = cand_off[0] + total * k / n_parts;
            while (c < n_cand && cand_off[c] < want) ++c;
            cut[k] = std::max(c, cut[k - 1]);
        }
        cut[n_parts] = n_cand;
    }
    cudaStream_t s0 = ctx->stream;
    HB_CUDA_OK(cudaEventRecord(ctx->ev[8], s0));
    for (int l = 1; l <= 2; ++l) HB_CUDA_OK(cudaStreamWaitEvent(ctx->lane[l].stream, ctx->ev[8], 0));
    std::vector<uint64_t> part_bytes(n_parts, 0);
    std::vector<hb_timing> part_tm(n_parts);
    std::vector<uint32_t> part_pairs(n_parts, 0);
    int lane_rc[2] = {HB_OK, HB_OK};
    auto worker = [&](int t) {
        cudaSetDevice(ctx->device);
        HbLane& ln = ctx->lane[1 + t];
        for (uint32_t k = (uint32_t)t; k < n_parts; k += 2) {
            const uint32_t c0 = cut[k], c1 = cut[k + 1];
            if (c1 <= c0) continue;
            std::vector<uint64_t> off(c1 - c0 + 1);
            for (uint32_t c = c0; c <= c1; ++c) off[c - c0] = cand_off[c] - cand_off[c0];
            int sub_long = 0;
            for (uint32_t c = c0; c < c1; ++c) if (cand_off[c + 1] - cand_off[c] > HB_EXTEND16_MAX_L) sub_long = 1;
            const int rc = adjust_batch(ctx, ln, ctx->part_cons[k], c1 - c0, cand_ascii + cand_off[c0], off.data(), prm,
                                        sub_long, calls + c0, &part_bytes[k]);
            if (rc != HB_OK) { lane_rc[t] = rc; return; }
            part_tm[k] = ln.timing; part_pairs[k] = ln.n_pairs;
        }
        cudaEventRecord(ln.ev[7], ln.stream);
    };
    {
        std::thread th(worker, 1);
        worker(0);
        th.join();
    }
    if (lane_rc[0] != HB_OK) return lane_rc[0];
    if (lane_rc[1] != HB_OK) return lane_rc[1];
    for (int l = 1; l <= 2; ++l) HB_CUDA_OK(cudaStreamWaitEvent(s0, ctx->lane[l].ev[7], 0));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[9], s0));
    HB_CUDA_OK(cudaStreamSynchronize(s0));
    // stitch: consensus parts in candidate order, cons_off rebased; stage times are sums over the
    // sub-batches (they overlap, so their sum exceeds kernel_ms = first launch .. last kernel end)
    hb_timing tm;
    memset(&tm, 0, sizeof(tm));
    uint64_t base = 0;
    uint32_t pairs = 0;
    for (uint32_t k = 0; k < n_parts; ++k) {
        if (base)
            for (uint32_t c = cut[k]; c < cut[k + 1]; ++c) calls[c].cons_off += (int64_t)base;
        ctx->cons_parts.push_back({(const uint8_t*)ctx->part_cons[k].p, part_bytes[k]});
        base += part_bytes[k];
        for (int i = 0; i < 6; ++i) tm.k_ms[i] += part_tm[k].k_ms[i];
        tm.k_ms[6] += part_tm[k].k_ms[6]; tm.k_ms[7] += part_tm[k].k_ms[7];
        tm.d2h_ms += part_tm[k].d2h_ms;
        tm.h2d_bytes += part_tm[k].h2d_bytes; tm.d2h_bytes += part_tm[k].d2h_bytes;
        tm.launches += part_tm[k].launches;
        pairs += part_pairs[k];
    }
    cudaEventElapsedTime(&tm.kernel_ms, ctx->ev[8], ctx->ev[9]);
    ctx->timing = tm;
    ctx->last_overlapped = true;
    ctx->last_cons_bytes = base;
    ctx->last_ncand = n_cand; ctx->last_npairs = pairs;
    ctx->last_mat_off.clear(); ctx->last_rows.clear(); ctx->last_cols.clear();
    if (cons_bytes) *cons_bytes = base;
    return HB_OK;
}

extern "C" int hb_query_hsps(hb_ctx* ctx, uint32_t n_query, const uint8_t* query_ascii, const uint64_t* query_off,
                             const hb_params* prm, uint64_t* n_out) {
    if (!ctx || !query_ascii || !query_off || !prm || !n_out) return hb_fail_msg(ctx, HB_EINVAL, "hb_query_hsps: null argument");
    if (!ctx->have_index) return hb_fail_msg(ctx, HB_ESTATE, "hb_query_hsps: call hb_genome_load + hb_index_build first");
    ctx->last_qhsps.clear();
    *n_out = 0;
    if (n_query == 0) return HB_OK;
    if (n_query > (1u << 17)) return hb_fail_msg(ctx, HB_EINVAL, "hb_query_hsps: at most 131072 queries per call");
    if (prm->max_loci < 1 || prm->max_loci > 512 || prm->occ_cap < 1)
        return hb_fail_msg(ctx, HB_EINVAL, "hb_query_hsps: parameter out of range");
    int any_long = 0;
    for (uint32_t c = 0; c < n_query; ++c) {
        const uint64_t L = query_off[c + 1] - query_off[c];
        if (L < 3 || L > 32000) return hb_fail_msg(ctx, HB_EINVAL, "hb_query_hsps: query " + std::to_string(c) + " length must be in [3, 32000]");
        if (L > HB_EXTEND16_MAX_L) any_long = 1;
    }
    cudaSetDevice(ctx->device);
    HbLane& ln = ctx->lane[0];
    const int rc = adjust_batch(ctx, ln, ctx->part_cons.empty() ? ctx->h_cons : ctx->part_cons[0], n_query, query_ascii, query_off,
                                prm, any_long, nullptr, nullptr, &ctx->last_qhsps);
    if (rc != HB_OK) return rc;
    ctx->timing = ln.timing;
    // the hb_last_copies / hb_last_msa / hb_last_cons getters describe an hb_adjust: none is current now
    ctx->last_ncand = 0; ctx->last_npairs = 0; ctx->last_cons_bytes = 0; ctx->cons_parts.clear();
    ctx->last_mat_off.clear(); ctx->last_rows.clear(); ctx->last_cols.clear();
    *n_out = ctx->last_qhsps.size();
    return HB_OK;
}

extern "C" int hb_last_query_hsps(hb_ctx* ctx, hb_qhsp* out, uint64_t cap, uint64_t* n) {
    if (!ctx || !n) return HB_EINVAL;
    *n = ctx->last_qhsps.size();
    if (!out) return HB_OK;
    if (cap < ctx->last_qhsps.size()) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_query_hsps: buffer too small");
    if (!ctx->last_qhsps.empty()) memcpy(out, ctx->last_qhsps.data(), sizeof(hb_qhsp) * ctx->last_qhsps.size());
    return HB_OK;
}

extern "C" int hb_last_cons(hb_ctx* ctx, uint8_t* out, uint64_t cap) {
    if (!ctx || !out) return HB_EINVAL;
    if (cap < ctx->last_cons_bytes) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_cons: buffer too small");
    uint64_t at = 0;
    for (const auto& part : ctx->cons_parts) { if (part.second) memcpy(out + at, part.first, part.second); at += part.second; }
    return HB_OK;
}

extern "C" int hb_last_cons_view(hb_ctx* ctx, const uint8_t** ptr, uint64_t* nbytes) {
    if (!ctx || !ptr || !nbytes) return HB_EINVAL;
    *nbytes = ctx->last_cons_bytes;
    *ptr = nullptr;
    if (!ctx->last_cons_bytes) return HB_OK;
    const std::pair<const uint8_t*, uint64_t>* only = nullptr;
    int n_parts = 0;
    for (const auto& part : ctx->cons_parts) if (part.second) { only = &part; ++n_parts; }
    if (n_parts == 1 && only->second == ctx->last_cons_bytes) { *ptr = only->first; return HB_OK; }
    ctx->cons_stitched.resize(ctx->last_cons_bytes);              // an overlapped run leaves one piece per sub-batch
    uint64_t at = 0;
    for (const auto& part : ctx->cons_parts) { if (part.second) memcpy(ctx->cons_stitched.data() + at, part.first, part.second); at += part.second; }
    *ptr = ctx->cons_stitched.data();
    return HB_OK;
}

extern "C" int hb_host_register(void* ptr, uint64_t nbytes) {
    if (!ptr || !nbytes) return HB_EINVAL;
    const cudaError_t e = cudaHostRegister(ptr, (size_t)nbytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return HB_ECUDA; }
    return HB_OK;
}

extern "C" int hb_host_unregister(void* ptr) {
    if (!ptr) return HB_EINVAL;
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return HB_ECUDA; }
    return HB_OK;
}

extern "C" int hb_last_cons_device(hb_ctx* ctx, const void** dptr, uint64_t* nbytes) {
    if (!ctx || !dptr || !nbytes) return HB_EINVAL;
    *nbytes = ctx->last_cons_bytes;
    *dptr = nullptr;
    if (!ctx->last_cons_bytes) return HB_OK;
    if (!ctx->last_overlapped) { *dptr = ctx->buf("v.dense").p; return HB_OK; }
    // overlapped run: the sub-batches' dense buffers were reused; stage the stitched bytes once more
    cudaSetDevice(ctx->device);
    HB_CUDA_OK(ctx->buf("v.dense_all").ensure(ctx->last_cons_bytes + 1));
    uint64_t at = 0;
    for (const auto& part : ctx->cons_parts) {
        if (part.second)
            HB_CUDA_OK(cudaMemcpyAsync((uint8_t*)ctx->buf("v.dense_all").p + at, part.first, part.second, cudaMemcpyHostToDevice, ctx->stream));
        at += part.second;
    }
    HB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    *dptr = ctx->buf("v.dense_all").p;
    return HB_OK;
}

extern "C" int hb_last_pair_counts(hb_ctx* ctx, uint32_t* out, uint32_t cap) {
    if (!ctx || !out) return HB_EINVAL;
    if (ctx->last_overlapped) return hb_fail_msg(ctx, HB_ESTATE, "hb_last_pair_counts: needs a serial hb_adjust (hb_params.overlap < 2)");
    if (cap < ctx->last_ncand) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_pair_counts: buffer too small");
    if (ctx->last_ncand == 0) return HB_OK;
    cudaSetDevice(ctx->device);
    HB_CUDA_OK(cudaMemcpyAsync(out, ctx->buf("s.n_loci").p, sizeof(uint32_t) * ctx->last_ncand, cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}

extern "C" int hb_last_copies(hb_ctx* ctx, hb_copy* out, uint64_t cap, uint64_t* n) {
    if (!ctx || !n) return HB_EINVAL;
    *n = ctx->last_npairs;
    if (!out) return HB_OK;
    if (ctx->last_overlapped) return hb_fail_msg(ctx, HB_ESTATE, "hb_last_copies: needs a serial hb_adjust (hb_params.overlap < 2)");
    if (cap < ctx->last_npairs) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_copies: buffer too small");
    if (ctx->last_npairs == 0) return HB_OK;
    cudaSetDevice(ctx->device);
    std::vector<HbLocus> pr(ctx->last_npairs);
    std::vector<HbAln> al(ctx->last_npairs);
    HB_CUDA_OK(cudaMemcpyAsync(pr.data(), ctx->buf("p.pairs").p, sizeof(HbLocus) * pr.size(), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA_OK(cudaMemcpyAsync(al.data(), ctx->buf("p.alns").p, sizeof(HbAln) * al.size(), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < ctx->last_npairs; ++i) {
        hb_copy& o = out[i];
        o.cand = pr[i].cand; o.rank = (int32_t)pr[i].rank; o.rel = (int32_t)pr[i].rel; o.nseeds = (int32_t)pr[i].nseeds;
        o.score = al[i].score; o.i_s = al[i].i_s; o.i_e = al[i].i_e; o.keep = al[i].keep; o.row = al[i].row;
        o.contig = al[i].contig; o.j0 = pr[i].j0; o.j_s = al[i].j_s; o.j_e = al[i].j_e;
        o.start1 = al[i].start1; o.end1 = al[i].end1; o.ext_len = al[i].ext_len; o.trow = al[i].trow;
    }
    return HB_OK;
}

// Raw (pre-clean) matrices of the last serial hb_adjust, written on demand from the buffers it left on the device.
static int materialise_raw(hb_ctx* ctx) {
    HbLane& ln = ctx->lane[0];
    if (ln.raw_valid) return HB_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ln.stream;
    HB_CUDA_OK(ctx->buf("m.mat").ensure(ln.raw_bytes + 4));
    for (int trunc = 0; trunc <= (ln.raw_has_trunc ? 1 : 0); ++trunc)
        HB_CUDA_OK(hb_msa_write_launch(ctx->buf("p.pairs").as<HbLocus>(), ln.n_pairs, ctx->buf("p.alns").as<HbAln>(),
                                       ctx->buf("p.qr_off").as<uint32_t>(), ctx->buf("p.qrec").as<uint32_t>(),
                                       ctx->buf("m.n_cols").as<uint32_t>(), ctx->buf("m.col_off").as<uint32_t>(),
                                       ctx->buf("m.colq").as<uint32_t>(),
                                       ctx->buf(trunc ? "m.tmat_off" : "m.mat_off").as<uint64_t>(), ctx->buf("g.seq2").as<uint32_t>(),
                                       ctx->buf("g.nmask").as<uint32_t>(), ctx->g_tot, ln.raw_flank, trunc, ctx->buf("m.mat").as<uint8_t>(), s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    ln.raw_valid = true;
    return HB_OK;
}

extern "C" int hb_last_msa_trunc(hb_ctx* ctx, uint32_t cand, uint8_t* out_ascii, uint64_t cap, int32_t* R, int32_t* C) {
    if (!ctx || !R || !C) return HB_EINVAL;
    if (ctx->last_overlapped) return hb_fail_msg(ctx, HB_ESTATE, "hb_last_msa_trunc: needs a serial hb_adjust (hb_params.overlap < 2)");
    if (cand >= ctx->last_ncand || cand >= ctx->last_trows.size()) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_msa_trunc: candidate out of range");
    *R = ctx->last_rows[cand] ? (int32_t)ctx->last_trows[cand] : 0; *C = (int32_t)ctx->last_cols[cand];
    const uint64_t bytes = (uint64_t)*R * (uint64_t)*C;
    if (!out_ascii || bytes == 0) return HB_OK;
    if (cap < bytes) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_msa_trunc: buffer too small");
    { const int rc = materialise_raw(ctx); if (rc != HB_OK) return rc; }
    cudaSetDevice(ctx->device);
    HB_CUDA_OK(cudaMemcpyAsync(out_ascii, ctx->buf("m.mat").as<uint8_t>() + ctx->last_tmat_off[cand], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i < bytes; ++i) out_ascii[i] = (uint8_t)hb_ascii_of_sym(out_ascii[i]);
    return HB_OK;
}

extern "C" int hb_last_msa(hb_ctx* ctx, uint32_t cand, uint8_t* out_ascii, uint64_t cap, int32_t* R, int32_t* C) {
    if (!ctx || !R || !C) return HB_EINVAL;
    if (ctx->last_overlapped) return hb_fail_msg(ctx, HB_ESTATE, "hb_last_msa: needs a serial hb_adjust (hb_params.overlap < 2)");
    if (cand >= ctx->last_ncand) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_msa: candidate out of range");
    *R = (int32_t)ctx->last_rows[cand]; *C = (int32_t)ctx->last_cols[cand];
    const uint64_t bytes = (uint64_t)*R * (uint64_t)*C;
    if (!out_ascii || bytes == 0) return HB_OK;
    if (cap < bytes) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_msa: buffer too small");
    { const int rc = materialise_raw(ctx); if (rc != HB_OK) return rc; }
    cudaSetDevice(ctx->device);
    HB_CUDA_OK(cudaMemcpyAsync(out_ascii, ctx->buf("m.mat").as<uint8_t>() + ctx->last_mat_off[cand], bytes, cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i < bytes; ++i) out_ascii[i] = (uint8_t)hb_ascii_of_sym(out_ascii[i]);
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int hb_fmea_chain(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs, int skip_gap,
                             int max_len, int threads, hb_interval* out, uint64_t cap, uint64_t* n_out) {
    if (!q_off || !segs || !n_out || (n_seg && q_off[n_seg] && !hsp)) return HB_EINVAL;
    std::vector<hb_interval> res;
    hb_fmea_run(hsp, q_off, n_seg, segs, skip_gap, max_len, threads, res);
    *n_out = res.size();
    if (!out) return HB_OK;
    if (cap < res.size()) return HB_EINVAL;
    if (!res.empty()) memcpy(out, res.data(), sizeof(hb_interval) * res.size());
    return HB_OK;
}

extern "C" int hb_debug_fmea_blocks(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs,
                                    const uint32_t* file_of_seg, int skip_gap, int max_len, int threads, uint32_t n_blocks,
                                    hb_interval* out, uint64_t cap, uint64_t* n_out) {
    if (!q_off || !segs || !n_out || !file_of_seg || !n_blocks || (n_seg && q_off[n_seg] && !hsp)) return HB_EINVAL;
    for (uint32_t g = 1; g < n_seg; ++g) if (file_of_seg[g] < file_of_seg[g - 1]) return HB_EINVAL;
    std::vector<uint32_t> first_seg;                                   // first segment of every file (+ n_seg)
    for (uint32_t g = 0; g < n_seg; ++g) if (g == 0 || file_of_seg[g] != file_of_seg[g - 1]) first_seg.push_back(g);
    const uint32_t nf = (uint32_t)first_seg.size();
    first_seg.push_back(n_seg);
    HbFmeaPool* pool = hb_fmea_pool_start(n_seg, segs, file_of_seg, skip_gap, max_len, threads);
    std::vector<std::vector<uint64_t>> offs(n_blocks, std::vector<uint64_t>((size_t)n_seg + 1));
    for (uint32_t b = 0; b < n_blocks; ++b) {                          // block b: whole files, the other segments empty
        const uint32_t g_lo = first_seg[(size_t)((uint64_t)nf * b / n_blocks)], g_hi = first_seg[(size_t)((uint64_t)nf * (b + 1) / n_blocks)];
        for (uint32_t g = 0; g <= n_seg; ++g) offs[b][g] = q_off[std::min(std::max(g, g_lo), g_hi)];
        hb_fmea_pool_submit(pool, hsp, offs[b].data());
    }
    std::vector<hb_interval> res;
    hb_fmea_pool_finish(pool, res, nullptr, nullptr);
    *n_out = res.size();
    if (!out) return HB_OK;
    if (cap < res.size()) return HB_EINVAL;
    if (!res.empty()) memcpy(out, res.data(), sizeof(hb_interval) * res.size());
    return HB_OK;
}

extern "C" int hb_fmea_chain_files(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs,
                                   const uint32_t* file_of_seg, int skip_gap, int max_len, int threads, hb_interval* out,
                                   uint64_t cap, uint64_t* n_out) {
    if (!q_off || !segs || !n_out || !file_of_seg || (n_seg && q_off[n_seg] && !hsp)) return HB_EINVAL;
    for (uint32_t g = 1; g < n_seg; ++g) if (file_of_seg[g] < file_of_seg[g - 1]) return HB_EINVAL;
    std::vector<hb_interval> res;
    hb_fmea_run(hsp, q_off, n_seg, segs, skip_gap, max_len, threads, res, file_of_seg);
    *n_out = res.size();
    if (!out) return HB_OK;
    if (cap < res.size()) return HB_EINVAL;
    if (!res.empty()) memcpy(out, res.data(), sizeof(hb_interval) * res.size());
    return HB_OK;
}

extern "C" int hb_query_copies(const hb_hsp* hsp, const double* ident, uint32_t n, int query_len, const int32_t* subj_len,
                               double query_coverage, double subject_coverage, int qthr, int sthr, int max_copy_num,
                               hb_query_copy* out, uint32_t cap, uint32_t* n_out) {
    if (!n_out || (n && (!hsp || !ident)) || query_len <= 0) return HB_EINVAL;
    std::vector<hb_query_copy> res;
    hb_query_copies_run(hsp, ident, n, query_len, subj_len, query_coverage, subject_coverage, qthr, sthr, max_copy_num, res);
    *n_out = (uint32_t)res.size();
    if (!out) return HB_OK;
    if (cap < res.size()) return HB_EINVAL;
    if (!res.empty()) memcpy(out, res.data(), sizeof(hb_query_copy) * res.size());
    return HB_OK;
}

extern "C" int hb_full_length_copies(const hb_hsp* hsp, uint32_t n, int query_len, double full_length_threshold,
                                     hb_query_copy* out, uint32_t cap, uint32_t* n_out) {
    if ((!hsp && n) || !n_out || query_len < 1) return HB_EINVAL;
    for (uint32_t i = 0; i < n; ++i) if (hsp[i].subj < 0) return HB_EINVAL;
    std::vector<hb_query_copy> res;
    hb_full_length_copies_run(hsp, n, query_len, full_length_threshold, res);
    *n_out = (uint32_t)res.size();
    if (!out) return HB_OK;
    if (cap < res.size()) return HB_EINVAL;
    if (!res.empty()) memcpy(out, res.data(), sizeof(hb_query_copy) * res.size());
    return HB_OK;
}

// Batched chaining over the records of many queries (query g owns hsp[q_off[g] .. q_off[g+1])), `threads` host threads over
// queries; copies come back grouped by query, out_query[i] = query of out[i].  cap >= q_off[n_query] always suffices (every copy
// consumes at least one record of its own).
template <class F>
static int chain_batch(uint32_t n_query, int threads, hb_query_copy* out, uint32_t* out_query, uint64_t cap, uint64_t* n_out, F one) {
    const int T = std::max(1, std::min<int>(threads, 64));
    std::vector<std::vector<hb_query_copy>> res(n_query);
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        for (;;) {
            const uint32_t g0 = next.fetch_add(64);
            if (g0 >= n_query) break;
            for (uint32_t g = g0; g < std::min(n_query, g0 + 64); ++g) one(g, res[g]);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    uint64_t n = 0;
    for (uint32_t g = 0; g < n_query; ++g) n += res[g].size();
    *n_out = n;
    if (!out) return HB_OK;
    if (cap < n) return HB_EINVAL;
    uint64_t at = 0;
    for (uint32_t g = 0; g < n_query; ++g)
        for (const hb_query_copy& c : res[g]) { out[at] = c; if (out_query) out_query[at] = g; ++at; }
    return HB_OK;
}

extern "C" int hb_full_length_copies_batch(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_query, const int32_t* query_len,
                                           double full_length_threshold, int threads, hb_query_copy* out, uint32_t* out_query,
                                           uint64_t cap, uint64_t* n_out) {
    if (!q_off || !query_len || !n_out || (!hsp && n_query && q_off[n_query] > q_off[0])) return HB_EINVAL;
    for (uint64_t i = q_off[0]; i < q_off[n_query]; ++i) if (hsp[i].subj < 0) return HB_EINVAL;
    return chain_batch(n_query, threads, out, out_query, cap, n_out, [&](uint32_t g, std::vector<hb_query_copy>& r) {
        if (q_off[g + 1] > q_off[g] && query_len[g] > 0)
            hb_full_length_copies_run(hsp + q_off[g], (size_t)(q_off[g + 1] - q_off[g]), query_len[g], full_length_threshold, r);
    });
}

extern "C" int hb_query_copies_batch(const hb_hsp* hsp, const double* ident, const uint64_t* q_off, uint32_t n_query,
                                     const int32_t* query_len, double query_coverage, int qthr, int sthr, int max_copy_num,
                                     int threads, hb_query_copy* out, uint32_t* out_query, uint64_t cap, uint64_t* n_out) {
    if (!q_off || !query_len || !n_out || ((!hsp || !ident) && n_query && q_off[n_query] > q_off[0])) return HB_EINVAL;
    for (uint64_t i = q_off[0]; i < q_off[n_query]; ++i) if (hsp[i].subj < 0) return HB_EINVAL;
    return chain_batch(n_query, threads, out, out_query, cap, n_out, [&](uint32_t g, std::vector<hb_query_copy>& r) {
        if (q_off[g + 1] > q_off[g] && query_len[g] > 0)
            hb_query_copies_run(hsp + q_off[g], ident + q_off[g], (size_t)(q_off[g + 1] - q_off[g]), query_len[g], nullptr,
                                query_coverage, 0.0, qthr, sthr, max_copy_num, r);
    });
}

extern "C" int hb_genome_mask(hb_ctx* ctx, uint32_t n, const int32_t* contig, const int64_t* start, const int64_t* end,
                              uint64_t* n_masked) {
    if (!ctx || (n && (!contig || !start || !end))) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_mask: null argument");
    if (!ctx->have_genome) return hb_fail_msg(ctx, HB_ESTATE, "hb_genome_mask: call hb_genome_load first");
    for (uint32_t i = 0; i < n; ++i)
        if (contig[i] < 0 || (uint32_t)contig[i] >= ctx->g_ncontig)
            return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_mask: interval " + std::to_string(i) + " names an unknown contig");
    if (n_masked) *n_masked = 0;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    if (n == 0) return HB_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ENSURE("k.contig", sizeof(int32_t) * n); ENSURE("k.start", sizeof(int64_t) * n); ENSURE("k.end", sizeof(int64_t) * n);
    ENSURE("k.count", sizeof(unsigned long long));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("k.contig").p, contig, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("k.start").p, start, sizeof(int64_t) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("k.end").p, end, sizeof(int64_t) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemsetAsync(ctx->buf("k.count").p, 0, sizeof(unsigned long long), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    HB_CUDA_OK(hb_mask_fill_launch(ctx->buf("k.contig").as<int32_t>(), ctx->buf("k.start").as<int64_t>(), ctx->buf("k.end").as<int64_t>(), n,
                                   ctx->buf("g.goff").as<uint64_t>(), ctx->buf("g.glen").as<uint64_t>(), ctx->g_ncontig,
                                   ctx->buf("g.nmask").as<uint32_t>(), ctx->buf("k.count").as<unsigned long long>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    unsigned long long fresh = 0;
    HB_CUDA_OK(cudaMemcpyAsync(&fresh, ctx->buf("k.count").p, sizeof(fresh), cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    ctx->have_index = false;                       // the sketch of the unmasked genome is stale
    if (n_masked) *n_masked = fresh;
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    ctx->timing.k_ms[0] = ctx->timing.kernel_ms;
    ctx->timing.h2d_bytes = 20ull * n;
    ctx->timing.launches = 1;
    return HB_OK;
}

extern "C" int hb_genome_export(hb_ctx* ctx, uint32_t contig, uint8_t* out_ascii, uint64_t cap) {
    if (!ctx || !out_ascii) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_export: null argument");
    if (!ctx->have_genome) return hb_fail_msg(ctx, HB_ESTATE, "hb_genome_export: call hb_genome_load first");
    if (contig >= ctx->g_ncontig) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_export: contig out of range");
    const uint64_t len = ctx->g_glen[contig];
    if (cap < len) return hb_fail_msg(ctx, HB_EINVAL, "hb_genome_export: buffer too small");
    if (len == 0) return HB_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ENSURE("k.ascii", len + 16);
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    HB_CUDA_OK(hb_export_launch(ctx->buf("g.seq2").as<uint32_t>(), ctx->buf("g.nmask").as<uint32_t>(), ctx->g_goff[contig], len,
                                ctx->buf("k.ascii").as<uint8_t>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    HB_CUDA_OK(cudaMemcpyAsync(out_ascii, ctx->buf("k.ascii").p, len, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[1], ctx->ev[2]);
    ctx->timing.d2h_bytes = len;
    ctx->timing.launches = 1;
    return HB_OK;
}

extern "C" int hb_coarse_boundary_part(hb_ctx* ctx, int64_t seg_len, int skip_gap, int max_len, int occ_cap, int threads,
                                       uint32_t part, uint32_t n_parts, hb_interval* out, uint64_t cap, uint64_t* n_out) {
    if (!ctx || !n_out) return HB_EINVAL;
    if (n_parts == 0 || part >= n_parts) return hb_fail_msg(ctx, HB_EINVAL, "hb_coarse_boundary_part: need part < n_parts");
    if (!ctx->have_index) return hb_fail_msg(ctx, HB_ESTATE, "hb_coarse_boundary: call hb_genome_load + hb_index_build first");
    if (seg_len < 1000 || occ_cap < 1 || occ_cap > 4096) return hb_fail_msg(ctx, HB_EINVAL, "hb_coarse_boundary: parameter out of range");
    if (ctx->g_nmm >= (1ull << 29)) return hb_fail_msg(ctx, HB_EINVAL, "hb_coarse_boundary: more than 2^29 minimizers");
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    hb_timing& tm = ctx->timing;
    uint32_t launches = 0;
    // segments: seg_len pieces of every contig (split_genome_chunks.py:44-52)
    std::vector<hb_segment> segs;
    std::vector<uint64_t> seg_goff;
    for (uint32_t c = 0; c < ctx->g_ncontig; ++c)
        for (uint64_t o = 0; o < ctx->g_glen[c]; o += (uint64_t)seg_len) {
            segs.push_back(hb_segment{(int32_t)c, 0, (int64_t)o});
            seg_goff.push_back(ctx->g_goff[c] + o);
        }
    const int n_seg = (int)segs.size();
    const uint32_t n_mm = (uint32_t)ctx->g_nmm;
    // BLAST result files of the reference: split_and_store_sequences packs the segment records into a file until it holds
    // >= 1 000 000 bases (module/Util.py:4737, :4987-5012), so a contig's short tail shares a file -- and FMEA's
    // de-duplication state -- with the records that follow it.  A file is the unit a part (rank) owns.
    std::vector<uint32_t> file_of_seg((size_t)n_seg);
    uint32_t n_files = 0;
    {
        uint32_t f = 0;
        uint64_t count = 0;
        for (int g = 0; g < n_seg; ++g) {
            file_of_seg[(size_t)g] = f;
            count += std::min<uint64_t>((uint64_t)seg_len, ctx->g_glen[(size_t)segs[(size_t)g].chr] - (uint64_t)segs[(size_t)g].start);
            if (count >= 1000000ull) { ++f; count = 0; }
        }
        n_files = n_seg ? file_of_seg[(size_t)n_seg - 1] + 1 : 0;
    }
    // this part's query segments: the files [n_files * part / n_parts, n_files * (part + 1) / n_parts), a contiguous
    // range of the padded coordinate space
    const uint32_t f_lo = (uint32_t)((uint64_t)n_files * part / n_parts), f_hi = (uint32_t)((uint64_t)n_files * (part + 1) / n_parts);
    ENSURE("h.seg_goff", sizeof(uint64_t) * (n_seg + 1));
    ENSURE("h.seg_len", sizeof(uint64_t) * (n_seg + 1));
    std::vector<uint64_t> seg_lens((size_t)n_seg);
    for (int g = 0; g < n_seg; ++g)
        seg_lens[(size_t)g] = std::min<uint64_t>((uint64_t)seg_len, ctx->g_glen[(size_t)segs[(size_t)g].chr] - (uint64_t)segs[(size_t)g].start);
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("h.seg_goff").p, seg_goff.data(), sizeof(uint64_t) * n_seg, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("h.seg_len").p, seg_lens.data(), sizeof(uint64_t) * n_seg, cudaMemcpyHostToDevice, s));
    ENSURE("h.lo", sizeof(uint32_t) * (n_mm + 1)); ENSURE("h.cnt", sizeof(uint32_t) * (n_mm + 1)); ENSURE("h.off", sizeof(uint32_t) * (n_mm + 1));
    ENSURE("h.scan_tmp", sizeof(uint32_t) * hb_scan_tmp_words(n_mm + 1));
    ENSURE("h.total", sizeof(uint64_t));
    uint64_t n_pairs_total = 0;
    // ---- GPU half of one block of files [fb_lo, fb_hi): seed pairs -> sort -> runs -> base-level HSPs, read back ----
    auto gpu_block = [&](uint32_t fb_lo, uint32_t fb_hi, HbHostBuf& pin, uint64_t& n_recs) -> int {
        int g_lo = n_seg, g_hi = n_seg;
        for (int g = 0; g < n_seg; ++g) {
            if (file_of_seg[(size_t)g] >= fb_lo && g_lo == n_seg) g_lo = g;
            if (file_of_seg[(size_t)g] >= fb_hi) { g_hi = g; break; }
        }
        const uint64_t qpos_lo = g_lo < n_seg ? seg_goff[(size_t)g_lo] : ctx->g_tot;
        const uint64_t qpos_hi = g_hi < n_seg ? seg_goff[(size_t)g_hi] : ctx->g_tot;
        n_recs = 0;
        if (qpos_hi <= qpos_lo) return HB_OK;
        HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
        HB_CUDA_OK(hb_pair_count_launch(ctx->buf("g.mm_pos").as<uint64_t>(), n_mm, ctx->buf("g.idx_a").as<uint64_t>(),
                                        ctx->buf("g.bucket").as<uint32_t>(), occ_cap, qpos_lo, qpos_hi, ctx->buf("h.lo").as<uint32_t>(),
                                        ctx->buf("h.cnt").as<uint32_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(ctx->buf("h.cnt").as<uint32_t>(), ctx->buf("h.off").as<uint32_t>(), n_mm,
                                         ctx->buf("h.scan_tmp").as<uint32_t>(), ctx->buf("h.total").as<uint64_t>(), s));
        uint64_t n_pairs = 0;
        HB_CUDA_OK(cudaMemcpyAsync(&n_pairs, ctx->buf("h.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        launches += 4;
        if (n_pairs >= (1ull << 32)) return hb_fail_msg(ctx, HB_ENOMEM, "hb_coarse_boundary: more than 2^32 seed pairs; lower occ_cap");
        n_pairs_total += n_pairs;
        if (n_pairs == 0) return HB_OK;
        ENSURE("h.keys_a", sizeof(uint64_t) * n_pairs); ENSURE("h.keys_b", sizeof(uint64_t) * n_pairs);
        ENSURE("h.sort_tmp", sizeof(uint32_t) * hb_sort_tmp_words(n_pairs));
        HB_CUDA_OK(hb_pair_fill_launch(ctx->buf("g.mm_pos").as<uint64_t>(), n_mm, ctx->buf("g.idx_a").as<uint64_t>(), ctx->g_nmm,
                                       ctx->g_tot, ctx->k, occ_cap, ctx->buf("h.lo").as<uint32_t>(), ctx->buf("h.cnt").as<uint32_t>(),
                                       ctx->buf("h.off").as<uint32_t>(), ctx->buf("h.keys_a").as<uint64_t>(), s));
        HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
        uint64_t* keys = nullptr;
        HB_CUDA_OK(hb_radix_sort_u64(ctx->buf("h.keys_a").as<uint64_t>(), ctx->buf("h.keys_b").as<uint64_t>(), n_pairs, 0, 64,
                                     ctx->buf("h.sort_tmp").as<uint32_t>(), &keys, &launches, s));
        HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
        ENSURE("h.flag", sizeof(uint32_t) * n_pairs); ENSURE("h.rid", sizeof(uint32_t) * n_pairs);
        ENSURE("h.scan_tmp2", sizeof(uint32_t) * hb_scan_tmp_words(n_pairs));
        HB_CUDA_OK(hb_hsp_flag_launch(keys, n_pairs, ctx->buf("g.mm_pos").as<uint64_t>(), ctx->g_tot, ctx->k,
                                      ctx->buf("h.seg_goff").as<uint64_t>(), n_seg, ctx->buf("h.flag").as<uint32_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(ctx->buf("h.flag").as<uint32_t>(), ctx->buf("h.rid").as<uint32_t>(), n_pairs,
                                         ctx->buf("h.scan_tmp2").as<uint32_t>(), ctx->buf("h.total").as<uint64_t>(), s));
        uint64_t n_runs64 = 0;
        HB_CUDA_OK(cudaMemcpyAsync(&n_runs64, ctx->buf("h.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        if (n_runs64 >= (1ull << 31)) return hb_fail_msg(ctx, HB_ENOMEM, "hb_coarse_boundary: more than 2^31 seed runs; lower occ_cap");
        const uint32_t n_runs = (uint32_t)n_runs64;
        // base-level HSP ends (HB-hsp S11b): per run the block-grown segments of its first / last seed, duplicates dropped
        ENSURE("h.run_start", sizeof(uint32_t) * (n_runs + 1)); ENSURE("h.keep", sizeof(uint32_t) * (2 * (size_t)n_runs + 2));
        ENSURE("h.pos", sizeof(uint32_t) * (2 * (size_t)n_runs + 2)); ENSURE("h.segs", sizeof(uint4) * ((size_t)n_runs + 1));
        ENSURE("h.scan_tmp3", sizeof(uint32_t) * hb_scan_tmp_words(2 * (size_t)n_runs + 2));
        HB_CUDA_OK(hb_hsp_start_launch(ctx->buf("h.flag").as<uint32_t>(), ctx->buf("h.rid").as<uint32_t>(), n_pairs,
                                       ctx->buf("h.run_start").as<uint32_t>(), s));
        HB_CUDA_OK(hb_hsp_segment_launch(keys, n_pairs, ctx->buf("h.run_start").as<uint32_t>(), n_runs, ctx->buf("g.mm_pos").as<uint64_t>(),
                                         ctx->buf("g.seq2").as<uint32_t>(), ctx->buf("g.nmask").as<uint32_t>(), ctx->g_tot, ctx->k,
                                         ctx->buf("h.seg_goff").as<uint64_t>(), ctx->buf("h.seg_len").as<uint64_t>(), n_seg,
                                         ctx->buf("h.segs").as<uint4>(), ctx->buf("h.keep").as<uint32_t>(), s));
        HB_CUDA_OK(hb_exclusive_scan_u32(ctx->buf("h.keep").as<uint32_t>(), ctx->buf("h.pos").as<uint32_t>(), 2 * (size_t)n_runs,
                                         ctx->buf("h.scan_tmp3").as<uint32_t>(), ctx->buf("h.total").as<uint64_t>(), s));
        uint64_t n_hsp = 0;
        HB_CUDA_OK(cudaMemcpyAsync(&n_hsp, ctx->buf("h.total").p, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        ENSURE("h.out", sizeof(HbHspRec) * (n_hsp + 1));
        HB_CUDA_OK(hb_hsp_emit_launch(keys, n_pairs, ctx->buf("h.run_start").as<uint32_t>(), n_runs, ctx->buf("g.mm_pos").as<uint64_t>(),
                                      ctx->g_tot, ctx->k, ctx->buf("h.seg_goff").as<uint64_t>(), n_seg, ctx->buf("h.segs").as<uint4>(),
                                      ctx->buf("h.keep").as<uint32_t>(), ctx->buf("h.pos").as<uint32_t>(),
                                      ctx->buf("h.out").as<HbHspRec>(), s));
        HB_CUDA_OK(cudaEventRecord(ctx->ev[3], s));
        HB_CUDA_OK(pin.ensure(sizeof(HbHspRec) * (n_hsp + 1)));                        // pinned: the read-back runs at PCIe speed
        if (n_hsp) HB_CUDA_OK(cudaMemcpyAsync(pin.p, ctx->buf("h.out").p, sizeof(HbHspRec) * n_hsp, cudaMemcpyDeviceToHost, s));
        HB_CUDA_OK(cudaStreamSynchronize(s));
        n_recs = n_hsp;
        launches += 12;
        tm.d2h_bytes += sizeof(HbHspRec) * n_hsp;
        float a = 0.f, b = 0.f, c = 0.f;
        cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]);
        tm.k_ms[0] += a; tm.k_ms[1] += b; tm.k_ms[2] += c; tm.kernel_ms += a + b + c;
        return HB_OK;
    };
    // ---- host half of one block: HSPs grouped by query segment (stable: the "file order" of a query is the sorted-key
    // order), FMEA over the block's files on `threads` host threads ----
    // (the grouped HSPs live in the context and only ever grow: a fresh 100 MB vector per block costs more in page faults
    // than the chaining itself)
    struct Block { std::vector<hb_hsp>* hsps = nullptr; uint64_t n = 0; std::vector<uint64_t> off; };
    // stable counting sort by query segment on up to 8 host threads (thread t owns a contiguous slice of the records)
    auto group_block = [&](const HbHspRec* recs, uint64_t n, Block& B) {
        const int T = (int)std::max<uint64_t>(1, std::min<uint64_t>({(uint64_t)std::max(1, threads), 8ull, n / 65536 + 1}));
        std::vector<std::vector<uint64_t>> cnt((size_t)T, std::vector<uint64_t>((size_t)n_seg + 1, 0));
        auto slice = [&](int t) { return std::make_pair(n * (uint64_t)t / (uint64_t)T, n * (uint64_t)(t + 1) / (uint64_t)T); };
        auto par = [&](auto fn) {
            std::vector<std::thread> th;
            for (int t = 1; t < T; ++t) th.emplace_back(fn, t);
            fn(0);
            for (auto& x : th) x.join();
        };
        par([&](int t) { auto [a, b] = slice(t); for (uint64_t i = a; i < b; ++i) cnt[(size_t)t][(size_t)recs[i].qseg]++; });
        B.off.assign((size_t)n_seg + 1, 0);
        uint64_t run = 0;
        for (int g = 0; g < n_seg; ++g) {
            B.off[(size_t)g] = run;
            for (int t = 0; t < T; ++t) { const uint64_t c = cnt[(size_t)t][(size_t)g]; cnt[(size_t)t][(size_t)g] = run; run += c; }
        }
        B.off[(size_t)n_seg] = run;
        if (B.hsps->size() < n) B.hsps->resize(n + n / 8);
        B.n = n;
        hb_hsp* dst = B.hsps->data();
        par([&](int t) {
            auto [a, b] = slice(t);
            std::vector<uint64_t>& cur = cnt[(size_t)t];
            for (uint64_t i = a; i < b; ++i) { const HbHspRec& r = recs[i]; dst[cur[(size_t)r.qseg]++] = hb_hsp{r.subj, r.qs, r.qe, r.ss, r.se}; }
        });
    };
    // FMEA runs as a task graph on one pool of `threads` host threads for the whole part (hb_fmea.cu): a block is grouped and
    // handed over as soon as its read-back is there, its chaining overlaps the GPU half of the following blocks and the
    // order-dependent per-file de-duplication of earlier files overlaps the chaining of later ones.
    auto t_all = std::chrono::steady_clock::now();
    HbFmeaPool* pool = hb_fmea_pool_start((uint32_t)n_seg, segs.data(), file_of_seg.data(), skip_gap, max_len, threads);
    auto host_block = [&](const HbHspRec* recs, uint64_t n, Block& B) {
        group_block(recs, n, B);
        hb_fmea_pool_submit(pool, B.hsps->data(), B.off.data());
    };
    // The part's files in up to 8 blocks, software-pipelined: while the host threads group and chain block b (FMEA is the
    // longer half), the GPU already produces the HSPs of block b + 1 into the other pinned read-back buffer.  A block is made of
    // whole files, so its chaining is independent.
    const uint32_t nf = f_hi - f_lo;
    uint32_t n_blocks = nf >= 16 ? 8u : nf >= 8 ? 4u : 1u;
    if (const char* e = getenv("HB_COARSE_BLOCKS")) {       // test hook: force the block count (the result does not depend on it)
        const long v = atol(e);
        if (v >= 1) n_blocks = (uint32_t)std::min<long>(v, std::max<uint32_t>(1u, nf));
    }
    std::vector<Block> blocks(n_blocks);
    if (ctx->hsp_store.size() < n_blocks) ctx->hsp_store.resize(n_blocks);
    for (uint32_t b = 0; b < n_blocks; ++b) blocks[b].hsps = &ctx->hsp_store[b];
    std::thread worker;
    int rc_gpu = HB_OK;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const uint32_t fb_lo = f_lo + (uint32_t)((uint64_t)nf * b / n_blocks), fb_hi = f_lo + (uint32_t)((uint64_t)nf * (b + 1) / n_blocks);
        HbHostBuf& pin = ctx->h_hsp_pin[b & 1];           // block b - 2's worker was joined before block b - 1's was started
        uint64_t n_recs = 0;
        rc_gpu = gpu_block(fb_lo, fb_hi, pin, n_recs);
        if (rc_gpu != HB_OK) break;
        if (worker.joinable()) worker.join();
        worker = std::thread(host_block, pin.as<HbHspRec>(), n_recs, std::ref(blocks[b]));
    }
    if (worker.joinable()) worker.join();
    std::vector<hb_interval> res;
    double fmea_span_ms = 0.0, fmea_busy_ms = 0.0;
    hb_fmea_pool_finish(pool, res, &fmea_span_ms, &fmea_busy_ms);
    if (rc_gpu != HB_OK) return rc_gpu;
    // hb_last_hsps (a parity / debug getter) assembles the part's HSPs from the blocks on demand
    size_t n_all = 0;
    ctx->last_hsp_blocks.clear();
    for (Block& B : blocks) { n_all += B.n; ctx->last_hsp_blocks.emplace_back(B.n, std::move(B.off)); }
    ctx->last_hsps.clear(); ctx->last_hsp_off.clear();
    ctx->last_segs = segs;
    tm.k_ms[3] = (float)fmea_span_ms;                                               // host FMEA: first block handed over -> last task done
    tm.k_ms[5] = (float)(fmea_busy_ms / std::max(1, threads));                      // ... and its task time per host thread
    tm.k_ms[4] = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_all).count();   // wall of the pipelined stage
    tm.k_ms[6] = (float)n_all;
    tm.k_ms[7] = (float)((double)n_pairs_total * 1e-6);
    tm.launches = launches;
    ctx->last_intervals = res;                       // kept: hb_last_intervals fetches them without a second run
    *n_out = res.size();
    if (!out) return HB_OK;
    if (cap < res.size()) return hb_fail_msg(ctx, HB_EINVAL, "hb_coarse_boundary: output buffer too small (fetch with hb_last_intervals)");
    if (!res.empty()) memcpy(out, res.data(), sizeof(hb_interval) * res.size());
    return HB_OK;
}

extern "C" int hb_coarse_boundary(hb_ctx* ctx, int64_t seg_len, int skip_gap, int max_len, int occ_cap, int threads,
                                  hb_interval* out, uint64_t cap, uint64_t* n_out) {
    return hb_coarse_boundary_part(ctx, seg_len, skip_gap, max_len, occ_cap, threads, 0, 1, out, cap, n_out);
}

extern "C" int hb_last_intervals(hb_ctx* ctx, hb_interval* out, uint64_t cap, uint64_t* n) {
    if (!ctx || !n) return HB_EINVAL;
    *n = ctx->last_intervals.size();
    if (!out) return HB_OK;
    if (cap < ctx->last_intervals.size()) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_intervals: buffer too small");
    if (!ctx->last_intervals.empty()) memcpy(out, ctx->last_intervals.data(), sizeof(hb_interval) * ctx->last_intervals.size());
    return HB_OK;
}

extern "C" int hb_last_hsps(hb_ctx* ctx, hb_hsp* out, uint64_t cap, uint64_t* n, uint64_t* q_off, hb_segment* segs,
                            uint32_t seg_cap, uint32_t* n_seg) {
    if (!ctx || !n || !n_seg) return HB_EINVAL;
    if (ctx->last_hsp_off.empty() && !ctx->last_hsp_blocks.empty()) {              // grouped by query segment, blocks in order
        const size_t ns = ctx->last_segs.size();
        ctx->last_hsp_off.assign(ns + 1, 0);
        size_t n_all = 0;
        for (const auto& B : ctx->last_hsp_blocks) { n_all += B.first; for (size_t g = 0; g <= ns; ++g) ctx->last_hsp_off[g] += B.second[g]; }
        ctx->last_hsps.resize(n_all);
        std::vector<uint64_t> cur(ctx->last_hsp_off.begin(), ctx->last_hsp_off.end() - 1);
        for (size_t b = 0; b < ctx->last_hsp_blocks.size(); ++b) {
            const auto& B = ctx->last_hsp_blocks[b];
            const hb_hsp* src = ctx->hsp_store[b].data();
            for (size_t g = 0; g < ns; ++g)
                for (uint64_t i = B.second[g]; i < B.second[g + 1]; ++i) ctx->last_hsps[cur[g]++] = src[i];
        }
        ctx->last_hsp_blocks.clear();
    }
    *n = ctx->last_hsps.size();
    *n_seg = (uint32_t)ctx->last_segs.size();
    if (!out) return HB_OK;
    if (cap < ctx->last_hsps.size() || seg_cap < ctx->last_segs.size()) return hb_fail_msg(ctx, HB_EINVAL, "hb_last_hsps: buffer too small");
    if (!ctx->last_hsps.empty()) memcpy(out, ctx->last_hsps.data(), sizeof(hb_hsp) * ctx->last_hsps.size());
    if (q_off) memcpy(q_off, ctx->last_hsp_off.data(), sizeof(uint64_t) * ctx->last_hsp_off.size());
    if (segs) memcpy(segs, ctx->last_segs.data(), sizeof(hb_segment) * ctx->last_segs.size());
    return HB_OK;
}
