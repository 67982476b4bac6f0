// hite_b200 -- kernel (ii) extension launchers (hb_extend.cu)
#pragma once
#include "hb_seed.cuh"

#define HB_EXTEND16_MAX_L 16000   // == hb16::HB16_MAX_L (hb_extend16.cuh); longer candidates take the 32-bit kernel

cudaError_t hb_pair_sizes_launch(const HbLocus* pairs, uint32_t n_pairs, const uint32_t* clen, uint32_t* tb_units,
                                 uint32_t* qrec_units, cudaStream_t s);
cudaError_t hb_extend_launch(const HbLocus* pairs, uint32_t pair_lo, uint32_t pair_hi, const uint32_t* gseq2,
                             const uint32_t* gnmask, uint64_t tot, const uint32_t* cseq2, const uint32_t* cnmask,
                             const uint64_t* cgoff, const uint32_t* clen, const uint32_t* tb_unit_off,
                             uint32_t tb_unit_base, uint32_t* tbuf, const uint32_t* qrec_unit_off, uint32_t* qrec,
                             HbAln* alns, uint32_t* order_ws, int any_long, uint32_t* launches, cudaStream_t s);
// words of scratch hb_extend_launch needs for a range of n_pairs pairs (order lists of the 16-bit kernels)
size_t hb_extend_order_words(uint32_t n_pairs);
