// hite_b200 -- C-ABI entry points (include/hite_b200.h).  Host-side C++ only; the kernels live in
// hb_vote.cu / hb_pack.cu / hb_index.cu / hb_extend.cu / hb_msa.cu.
#include "hb_ctx.cuh"

extern "C" size_t hb_vote_match_entry_size();
extern "C" size_t hb_vote_state_entries();
extern "C" size_t hb_vote_match_entries(int C);
extern "C" cudaError_t hb_vote_launch(const HbVoteTask* d_tasks, int n, const uint8_t* d_raw, uint8_t* d_clean,
                                      uint32_t* d_planes, void* d_matches, int* d_anchors, const uint8_t* d_cands,
                                      uint8_t* d_cons, HbVoteCall* d_calls, int te_type, int avg_rows, cudaStream_t stream);

static std::string g_last_global_error;

extern "C" int hb_version(void) { return 100; }

extern "C" int hb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* hb_last_error(hb_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_last_global_error.c_str();
}

extern "C" int hb_create(int device, hb_ctx** out) {
    if (!out) return HB_EINVAL;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        g_last_global_error = "hite_b200: no CUDA device available (there is no CPU fallback)";
        return HB_ENODEV;
    }
    if (device < 0 || device >= n) { g_last_global_error = "hite_b200: bad device ordinal"; return HB_EINVAL; }
    hb_ctx* ctx = new hb_ctx();
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_last_global_error = cudaGetErrorString(e); delete ctx; return HB_ECUDA; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_last_global_error = cudaGetErrorString(e); delete ctx; return HB_ECUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) {
        g_last_global_error = "hite_b200: kernels are built for sm_100a only; found compute capability " +
                              std::to_string(prop.major) + "." + std::to_string(prop.minor);
        delete ctx;
        return HB_ENODEV;
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_last_global_error = cudaGetErrorString(e); delete ctx; return HB_ECUDA;
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    // lane 0 = the serial lane on the context's stream; lanes 1, 2 = overlapped sub-batches
    ctx->lane[0].stream = ctx->stream;
    for (int l = 1; l < 3; ++l) {
        ctx->lane[l].prefix = l == 1 ? "L1:" : "L2:";
        if ((e = cudaStreamCreateWithFlags(&ctx->lane[l].stream, cudaStreamNonBlocking)) != cudaSuccess) {
            g_last_global_error = cudaGetErrorString(e); delete ctx; return HB_ECUDA;
        }
    }
    for (auto& ln : ctx->lane) for (auto& ev : ln.ev) cudaEventCreate(&ev);
    *out = ctx;
    return HB_OK;
}

extern "C" void hb_destroy(hb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (HbDevBuf* b : {&ctx->v_tasks, &ctx->v_raw, &ctx->v_clean, &ctx->v_planes, &ctx->v_matches, &ctx->v_cands,
                        &ctx->v_cons, &ctx->v_calls})
        b->release();
    for (auto& kv : ctx->bufs) kv.second.release();
    ctx->h_stage.release();
    ctx->h_cons.release();
    ctx->h_hsp_pin[0].release(); ctx->h_hsp_pin[1].release();
    for (auto& pc : ctx->part_cons) pc.release();
    for (auto& ln : ctx->lane) {
        cudaStreamSynchronize(ln.stream);
        ln.h_cons.release();
        for (auto& ev : ln.ev) if (ev) cudaEventDestroy(ev);
    }
    for (int l = 1; l < 3; ++l) if (ctx->lane[l].stream) cudaStreamDestroy(ctx->lane[l].stream);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int hb_get_timing(hb_ctx* ctx, hb_timing* out) {
    if (!ctx || !out) return HB_EINVAL;
    *out = ctx->timing;
    return HB_OK;
}

// ------------------------------------------------------------------------------------------
extern "C" int hb_vote_matrices(hb_ctx* ctx, uint32_t n, const uint8_t* rows_ascii, const uint64_t* mat_off,
                                const int32_t* R, const int32_t* C, const uint8_t* cand_ascii,
                                const uint64_t* cand_off, int te_type, int plant, int result_type,
                                hb_call* calls, uint8_t* cons_out, uint64_t cons_cap) {
    if (!ctx) return HB_EINVAL;
    if (n == 0) return HB_OK;
    if (!rows_ascii || !mat_off || !R || !C || !cand_ascii || !cand_off || !calls || !cons_out)
        return hb_fail_msg(ctx, HB_EINVAL, "hb_vote_matrices: null argument");
    if (te_type < HB_TE_TIR || te_type > HB_TE_NON_LTR)
        return hb_fail_msg(ctx, HB_EINVAL, "hb_vote_matrices: te_type must be HB_TE_TIR, HB_TE_HELITRON or HB_TE_NON_LTR");
    static_assert(sizeof(hb_call) == sizeof(HbVoteCall), "hb_call layout");
    cudaSetDevice(ctx->device);
    std::vector<HbVoteTask> tasks(n);
    uint64_t raw_total = 0, col_total = 0, match_total = 0, row_total = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (R[i] <= 0 || C[i] <= 0 || C[i] >= (1 << 26) || R[i] >= 65535)
            return hb_fail_msg(ctx, HB_EINVAL, "hb_vote_matrices: matrix " + std::to_string(i) + " has unsupported shape");
        HbVoteTask& t = tasks[i];
        memset(&t, 0, sizeof(t));
        t.raw_off = mat_off[i];
        t.clean_off = mat_off[i];
        t.plane_off = col_total * 24;
        t.cons_off = col_total + 2ull * i;          // capacity C + 2 (Helitron +-1 bp extension)
        t.match_off = match_total;
        t.anchor_off = 2 * row_total;
        t.cand_off = cand_off[i];
        t.cand_len = (int32_t)(cand_off[i + 1] - cand_off[i]);
        t.R = R[i]; t.C = C[i];
        t.te_type = te_type; t.plant = plant; t.result_type = result_type; t.raw_is_ascii = 1;
        raw_total = mat_off[i] + (uint64_t)R[i] * C[i] > raw_total ? mat_off[i] + (uint64_t)R[i] * C[i] : raw_total;
        col_total += (uint64_t)C[i];
        match_total += hb_vote_match_entries(C[i]);
        row_total += (uint64_t)R[i];
    }
    const uint64_t cons_total = col_total + 2ull * n;
    if (cons_cap < cons_total) return hb_fail_msg(ctx, HB_EINVAL, "hb_vote_matrices: cons_cap < sum(C) + 2n");
    const uint64_t cand_total = cand_off[n];
    HB_CUDA_OK(ctx->v_tasks.ensure(sizeof(HbVoteTask) * n));
    HB_CUDA_OK(ctx->v_raw.ensure(raw_total));
    HB_CUDA_OK(ctx->v_clean.ensure(raw_total));
    HB_CUDA_OK(ctx->v_planes.ensure(col_total * 24 * sizeof(uint32_t)));
    HB_CUDA_OK(ctx->v_matches.ensure(match_total * hb_vote_match_entry_size()));
    HB_CUDA_OK(ctx->v_cands.ensure(cand_total + 1));
    HB_CUDA_OK(ctx->v_cons.ensure(cons_total));
    HB_CUDA_OK(ctx->buf("v.anchors").ensure(sizeof(int) * (2 * row_total + 2)));
    HB_CUDA_OK(ctx->v_calls.ensure(sizeof(HbVoteCall) * n));
    cudaStream_t s = ctx->stream;
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->v_tasks.p, tasks.data(), sizeof(HbVoteTask) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->v_raw.p, rows_ascii, raw_total, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->v_cands.p, cand_ascii, cand_total, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemsetAsync(ctx->v_calls.p, 0, sizeof(HbVoteCall) * n, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    HB_CUDA_OK(hb_vote_launch(ctx->v_tasks.as<HbVoteTask>(), (int)n, ctx->v_raw.as<uint8_t>(), ctx->v_clean.as<uint8_t>(),
                              ctx->v_planes.as<uint32_t>(), ctx->v_matches.p, ctx->buf("v.anchors").as<int>(),
                              ctx->v_cands.as<uint8_t>(), ctx->v_cons.as<uint8_t>(), ctx->v_calls.as<HbVoteCall>(),
                              te_type, (int)(row_total / n), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    HB_CUDA_OK(cudaMemcpyAsync(calls, ctx->v_calls.p, sizeof(HbVoteCall) * n, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaMemcpyAsync(cons_out, ctx->v_cons.p, cons_total, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[3], s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    hb_timing& tm = ctx->timing;
    memset(&tm, 0, sizeof(tm));
    cudaEventElapsedTime(&tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&tm.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&tm.d2h_ms, ctx->ev[2], ctx->ev[3]);
    tm.h2d_bytes = sizeof(HbVoteTask) * n + raw_total + cand_total;
    tm.d2h_bytes = sizeof(HbVoteCall) * n + cons_total;
    tm.launches = 2;
    return HB_OK;
}

// ------------------------------------------------------------------------------------------------
// TSD-bounded variant generation (hb_tsd.cu)
#include "hb_tsd.cuh"
#include "hb_itr.cuh"
extern "C" int hb_itr_search(hb_ctx* ctx, uint32_t n, const uint8_t* seq_ascii, const uint64_t* seq_off, double min_identity,
                             int32_t min_len, int32_t max_window, hb_itr_rec* out) {
    if (!ctx || !seq_ascii || !seq_off || !out) return hb_fail_msg(ctx, HB_EINVAL, "hb_itr_search: null argument");
    if (max_window < 1 || max_window > 1000 || min_len < 1) return hb_fail_msg(ctx, HB_EINVAL, "hb_itr_search: need 1 <= max_window <= 1000, min_len >= 1");
    if (n == 0) return HB_OK;
    std::vector<uint64_t> off(n + 1), soff(n + 1);
    soff[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t len = seq_off[i + 1] - seq_off[i];
        if (len >= (1u << 20)) return hb_fail_msg(ctx, HB_EINVAL, "hb_itr_search: record " + std::to_string(i) + " is 2^20 bytes or longer");
        soff[i + 1] = soff[i] + 3 * (std::min<uint64_t>((uint64_t)max_window, len / 2) + 1);
    }
    for (uint32_t i = 0; i <= n; ++i) off[i] = seq_off[i] - seq_off[0];
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const uint64_t bytes = seq_off[n] - seq_off[0];
    HB_CUDA_OK(ctx->buf("i.seq").ensure(bytes + 16));
    HB_CUDA_OK(ctx->buf("i.off").ensure(sizeof(uint64_t) * (n + 1)));
    HB_CUDA_OK(ctx->buf("i.soff").ensure(sizeof(uint64_t) * (n + 1)));
    HB_CUDA_OK(ctx->buf("i.scratch").ensure(sizeof(long long) * (soff[n] + 1)));
    HB_CUDA_OK(ctx->buf("i.out").ensure(sizeof(hb_itr_rec) * (size_t)n));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("i.seq").p, seq_ascii + seq_off[0], bytes, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("i.off").p, off.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("i.soff").p, soff.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    HB_CUDA_OK(hb_itr_launch(ctx->buf("i.seq").as<uint8_t>(), ctx->buf("i.off").as<uint64_t>(), n, min_identity, min_len, max_window,
                             ctx->buf("i.scratch").as<long long>(), ctx->buf("i.soff").as<uint64_t>(), ctx->buf("i.out").as<hb_itr_rec>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    HB_CUDA_OK(cudaMemcpyAsync(out, ctx->buf("i.out").p, sizeof(hb_itr_rec) * (size_t)n, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[3], s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    ctx->timing.k_ms[0] = ctx->timing.kernel_ms;
    ctx->timing.h2d_bytes = bytes + 16ull * (n + 1);
    ctx->timing.d2h_bytes = sizeof(hb_itr_rec) * (uint64_t)n;
    ctx->timing.launches = 1;
    return HB_OK;
}

extern "C" int hb_tsd_candidates(hb_ctx* ctx, uint32_t n, const uint8_t* seq_ascii, const uint64_t* seq_off,
                                 const int32_t* raw_start1, const int32_t* raw_end1, int32_t d, int32_t plant,
                                 hb_tsd_rec* out, uint32_t* n_out) {
    if (!ctx || !seq_ascii || !seq_off || !raw_start1 || !raw_end1 || !out || !n_out)
        return hb_fail_msg(ctx, HB_EINVAL, "hb_tsd_candidates: null argument");
    if (d < 1 || d > 128) return hb_fail_msg(ctx, HB_EINVAL, "hb_tsd_candidates: tsd_search_distance must be in [1, 128]");
    if (n == 0) return HB_OK;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t len = seq_off[i + 1] - seq_off[i];
        if (len < 1 || len >= (1u << 20)) return hb_fail_msg(ctx, HB_EINVAL, "hb_tsd_candidates: candidate " + std::to_string(i) + " length must be in [1, 2^20)");
        if (raw_start1[i] < 1 || raw_end1[i] < 1 || (uint64_t)raw_start1[i] > len || (uint64_t)raw_end1[i] > len)
            return hb_fail_msg(ctx, HB_EINVAL, "hb_tsd_candidates: candidate " + std::to_string(i) + " boundary outside the sequence");
    }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const uint64_t bytes = seq_off[n] - seq_off[0];
    HB_CUDA_OK(ctx->buf("t.seq").ensure(bytes + 16));
    HB_CUDA_OK(ctx->buf("t.off").ensure(sizeof(uint64_t) * (n + 1)));
    HB_CUDA_OK(ctx->buf("t.rs").ensure(sizeof(int32_t) * n));
    HB_CUDA_OK(ctx->buf("t.re").ensure(sizeof(int32_t) * n));
    HB_CUDA_OK(ctx->buf("t.out").ensure(sizeof(hb_tsd_rec) * 100 * (size_t)n));
    HB_CUDA_OK(ctx->buf("t.n").ensure(sizeof(uint32_t) * n));
    std::vector<uint64_t> off(n + 1);
    for (uint32_t i = 0; i <= n; ++i) off[i] = seq_off[i] - seq_off[0];
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("t.seq").p, seq_ascii + seq_off[0], bytes, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("t.off").p, off.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("t.rs").p, raw_start1, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("t.re").p, raw_end1, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    HB_CUDA_OK(hb_tsd_launch(ctx->buf("t.seq").as<uint8_t>(), ctx->buf("t.off").as<uint64_t>(), ctx->buf("t.rs").as<int32_t>(),
                             ctx->buf("t.re").as<int32_t>(), n, d, plant, ctx->buf("t.out").as<hb_tsd_rec>(),
                             ctx->buf("t.n").as<uint32_t>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    HB_CUDA_OK(cudaMemcpyAsync(n_out, ctx->buf("t.n").p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaMemcpyAsync(out, ctx->buf("t.out").p, sizeof(hb_tsd_rec) * 100 * (size_t)n, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[3], s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    ctx->timing.k_ms[0] = ctx->timing.kernel_ms;
    ctx->timing.h2d_bytes = bytes + sizeof(uint64_t) * (n + 1) + 2 * sizeof(int32_t) * (uint64_t)n;
    ctx->timing.d2h_bytes = (sizeof(hb_tsd_rec) * 100 + sizeof(uint32_t)) * (uint64_t)n;
    ctx->timing.launches = 1;
    return HB_OK;
}

extern "C" int hb_nonltr_candidates(hb_ctx* ctx, uint32_t n, const uint8_t* seq_ascii, const uint64_t* seq_off, int32_t flank,
                                    hb_polya_rec* out) {
    if (!ctx || !seq_ascii || !seq_off || !out) return hb_fail_msg(ctx, HB_EINVAL, "hb_nonltr_candidates: null argument");
    if (flank < 0 || flank > 1000) return hb_fail_msg(ctx, HB_EINVAL, "hb_nonltr_candidates: flanking_len must be in [0, 1000]");
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    if (n == 0) return HB_OK;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t len = seq_off[i + 1] - seq_off[i];
        if (len < 1 || len >= (1u << 20)) return hb_fail_msg(ctx, HB_EINVAL, "hb_nonltr_candidates: candidate " + std::to_string(i) + " length must be in [1, 2^20)");
    }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const uint64_t bytes = seq_off[n] - seq_off[0];
    HB_CUDA_OK(ctx->buf("t.seq").ensure(bytes + 16));
    HB_CUDA_OK(ctx->buf("t.off").ensure(sizeof(uint64_t) * (n + 1)));
    HB_CUDA_OK(ctx->buf("t.polya").ensure(sizeof(hb_polya_rec) * (size_t)n));
    std::vector<uint64_t> off(n + 1);
    for (uint32_t i = 0; i <= n; ++i) off[i] = seq_off[i] - seq_off[0];
    HB_CUDA_OK(cudaEventRecord(ctx->ev[0], s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("t.seq").p, seq_ascii + seq_off[0], bytes, cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaMemcpyAsync(ctx->buf("t.off").p, off.data(), sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[1], s));
    HB_CUDA_OK(hb_polya_tsd_launch(ctx->buf("t.seq").as<uint8_t>(), ctx->buf("t.off").as<uint64_t>(), n, flank,
                                   ctx->buf("t.polya").as<hb_polya_rec>(), s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[2], s));
    HB_CUDA_OK(cudaMemcpyAsync(out, ctx->buf("t.polya").p, sizeof(hb_polya_rec) * (size_t)n, cudaMemcpyDeviceToHost, s));
    HB_CUDA_OK(cudaEventRecord(ctx->ev[3], s));
    HB_CUDA_OK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->timing.kernel_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->timing.d2h_ms, ctx->ev[2], ctx->ev[3]);
    ctx->timing.k_ms[0] = ctx->timing.kernel_ms;
    ctx->timing.h2d_bytes = bytes + sizeof(uint64_t) * (n + 1);
    ctx->timing.d2h_bytes = sizeof(hb_polya_rec) * (uint64_t)n;
    ctx->timing.launches = 1;
    return HB_OK;
}
