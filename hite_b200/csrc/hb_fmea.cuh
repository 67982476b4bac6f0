// hite_b200 -- coarse-boundary records shared by hb_fmea.cu / hb_coarse.cu / the C-ABI
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/hite_b200.h"

int hb_fmea_run(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs, int skip_gap,
                int max_len, int threads, std::vector<hb_interval>& result);
