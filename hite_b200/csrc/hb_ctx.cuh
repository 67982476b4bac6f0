// hite_b200 -- context object behind the C-ABI (host side, C++).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/hite_b200.h"
#include "hb_common.cuh"

struct HbDevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

struct HbHostBuf {  // pinned staging
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return (T*)p; }
};

// One in-flight sub-batch of hb_adjust: its own stream, events, staging and result bookkeeping.
// Lane 0 is the serial lane (the context's stream, un-prefixed buffers); lanes 1 and 2 run the
// overlapped sub-batches of hb_params.overlap >= 2 on their own streams and "L1:" / "L2:" buffers.
struct HbLane {
    std::string prefix;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    hb_timing timing = {};
    HbHostBuf h_cons;              // pinned copy of the dense consensus bytes of this lane's last batch
    uint64_t cons_bytes = 0;
    uint32_t n_pairs = 0;
    uint64_t ext_steps = 0;
    bool raw_valid = false, raw_has_trunc = false;   // raw matrices of the last batch: materialised on demand (hb_last_msa)
    uint64_t raw_bytes = 0;
    int raw_flank = 50;
    std::vector<uint64_t> mat_off, tmat_off;
    std::vector<uint32_t> rows, cols, trows;
};

struct hb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[12] = {};
    std::string err;
    hb_timing timing = {};
    // vote path buffers
    HbDevBuf v_tasks, v_raw, v_clean, v_planes, v_matches, v_cands, v_cons, v_calls;
    HbHostBuf h_stage;
    // named device buffers of the genome / index / adjust pipeline (hb_pipeline.cu)
    std::map<std::string, HbDevBuf> bufs;
    std::mutex mu;                 // guards `bufs` (node insertion) and `err` when two lanes run
    HbDevBuf& buf(const char* name) { std::lock_guard<std::mutex> g(mu); return bufs[name]; }
    HbDevBuf& buf(const std::string& name) { std::lock_guard<std::mutex> g(mu); return bufs[name]; }
    HbLane lane[3];
    std::vector<HbHostBuf> part_cons;                                  // pinned consensus bytes per sub-batch
    std::vector<std::pair<const uint8_t*, uint64_t>> cons_parts;       // last hb_adjust: parts in candidate order
    std::vector<uint8_t> cons_stitched;                                // hb_last_cons_view of an overlapped run
    bool last_overlapped = false;  // the hb_last_copies / hb_last_msa getters need a serial hb_adjust
    // genome state (HB-align S1/S2)
    bool have_genome = false, have_index = false;
    uint64_t g_tot = 0, g_nmm = 0;
    uint32_t g_ncontig = 0;
    int k = 15, w = 10;
    std::vector<uint64_t> g_goff, g_glen;
    // results of the last hb_adjust kept for hb_last_* getters
    HbHostBuf h_cons;              // pinned copy of the dense consensus bytes of the last hb_adjust
    uint64_t last_cons_bytes = 0;
    uint32_t last_ncand = 0, last_npairs = 0;
    std::vector<uint64_t> last_mat_off, last_tmat_off;
    std::vector<uint32_t> last_trows;
    std::vector<hb_hsp> last_hsps;            // last hb_coarse_boundary: HSPs grouped by query segment
    std::vector<uint64_t> last_hsp_off;
    std::vector<std::pair<uint64_t, std::vector<uint64_t>>> last_hsp_blocks;   // ... per pipelined block (count, offsets), until asked for
    std::vector<std::vector<hb_hsp>> hsp_store;   // the blocks' grouped HSPs (grow-only, reused by every coarse call)
    HbHostBuf h_hsp_pin[2];                   // pinned read-back buffers of the coarse stage's HSP records (double-buffered)
    std::vector<hb_segment> last_segs;
    std::vector<hb_interval> last_intervals;  // last hb_coarse_boundary(_part): its result, for hb_last_intervals
    std::vector<uint32_t> last_rows, last_cols;
    std::vector<hb_qhsp> last_qhsps;          // last hb_query_hsps, ordered by (query, rank)
};

inline int hb_fail(hb_ctx* ctx, const char* what, cudaError_t e, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    if (ctx) { std::lock_guard<std::mutex> g(ctx->mu); ctx->err = buf; }
    return e == cudaErrorMemoryAllocation ? HB_ENOMEM : HB_ECUDA;
}
inline int hb_fail_msg(hb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) { std::lock_guard<std::mutex> g(ctx->mu); ctx->err = msg; }
    return code;
}
