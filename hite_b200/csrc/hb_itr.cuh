// hite_b200 -- HB-itr launcher (hb_itr.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hite_b200.h"

// scratch words (int64) record i needs: 3 * (min(max_w, len / 2) + 1); scratch_off = their exclusive prefix
cudaError_t hb_itr_launch(const uint8_t* seqs, const uint64_t* seq_off, uint32_t n, double min_ident, int min_len, int max_w,
                          long long* scratch, const uint64_t* scratch_off, hb_itr_rec* out, cudaStream_t s);
