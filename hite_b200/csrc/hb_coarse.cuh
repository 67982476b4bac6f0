// hite_b200 -- coarse-boundary HSP generation launchers (hb_coarse.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HB_COARSE_BUCKET_BITS 22

struct HbHspRec { int32_t qseg, subj, qs, qe, ss, se, pad; };

cudaError_t hb_pair_count_launch(const uint64_t* mm, uint32_t n_mm, const uint64_t* idx, const uint32_t* bucket,
                                 int occ_cap, uint64_t qpos_lo, uint64_t qpos_hi, uint32_t* lo, uint32_t* cnt, cudaStream_t s);
cudaError_t hb_pair_fill_launch(const uint64_t* mm, uint32_t n_mm, const uint64_t* idx, uint64_t n_idx, uint64_t tot, int k, int occ_cap,
                                const uint32_t* lo, const uint32_t* cnt, const uint32_t* off, uint64_t* keys,
                                cudaStream_t s);
cudaError_t hb_hsp_flag_launch(const uint64_t* keys, uint64_t n, const uint64_t* mm, uint64_t tot, int k,
                               const uint64_t* seg_goff, int n_seg, uint32_t* flag, cudaStream_t s);
cudaError_t hb_hsp_start_launch(const uint32_t* flag, const uint32_t* rid, uint64_t n, uint32_t* run_start, cudaStream_t s);
// base-level HSP ends: per run the segments of its first / last seed (segs) and the keep flags of its 2 slots
cudaError_t hb_hsp_segment_launch(const uint64_t* keys, uint64_t n, const uint32_t* run_start, uint32_t n_runs, const uint64_t* mm,
                                  const uint32_t* seq2, const uint32_t* nmask, uint64_t tot, int k, const uint64_t* seg_goff,
                                  const uint64_t* seg_len, int n_seg, uint4* segs, uint32_t* keep, cudaStream_t s);
// the kept segments as records at pos[] (exclusive scan of keep over the 2 * n_runs slots)
cudaError_t hb_hsp_emit_launch(const uint64_t* keys, uint64_t n, const uint32_t* run_start, uint32_t n_runs, const uint64_t* mm,
                               uint64_t tot, int k, const uint64_t* seg_goff, int n_seg, const uint4* segs, const uint32_t* keep,
                               const uint32_t* pos, HbHspRec* out, cudaStream_t s);
