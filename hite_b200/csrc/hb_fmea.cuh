// hite_b200 -- coarse-boundary records shared by hb_fmea.cu / hb_coarse.cu / the C-ABI
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/hite_b200.h"

int hb_fmea_run(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs, int skip_gap,
                int max_len, int threads, std::vector<hb_interval>& result, const uint32_t* file_of_seg = nullptr);
// the same chaining fed block by block while the GPU still produces HSPs (hb_coarse_boundary); see hb_fmea.cu
struct HbFmeaPool;
HbFmeaPool* hb_fmea_pool_start(uint32_t n_seg, const hb_segment* segs, const uint32_t* file_of_seg, int skip_gap, int max_len,
                               int threads);
void hb_fmea_pool_submit(HbFmeaPool* pool, const hb_hsp* hsp, const uint64_t* q_off);
void hb_fmea_pool_finish(HbFmeaPool* pool, std::vector<hb_interval>& result, double* span_ms, double* busy_ms);
int hb_query_copies_run(const hb_hsp* h, const double* ident, size_t n, int query_len, const int32_t* subj_len,
                        double query_coverage, double subject_coverage, int qthr, int sthr, int max_copy_num,
                        std::vector<hb_query_copy>& result);
int hb_full_length_copies_run(const hb_hsp* h, size_t n, int query_len, double full_length_threshold,
                              std::vector<hb_query_copy>& result);
