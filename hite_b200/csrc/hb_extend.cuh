// hite_b200 -- kernel (ii) extension launchers (hb_extend.cu)
#pragma once
#include "hb_seed.cuh"

cudaError_t hb_pair_sizes_launch(const HbLocus* pairs, uint32_t n_pairs, const uint32_t* clen, uint32_t* tb_units,
                                 uint32_t* qrec_units, cudaStream_t s);
cudaError_t hb_extend_launch(const HbLocus* pairs, uint32_t pair_lo, uint32_t pair_hi, const uint32_t* gseq2,
                             const uint32_t* gnmask, uint64_t tot, const uint32_t* cseq2, const uint32_t* cnmask,
                             const uint64_t* cgoff, const uint32_t* clen, const uint32_t* tb_unit_off,
                             uint32_t tb_unit_base, uint32_t* tbuf, const uint32_t* qrec_unit_off, uint32_t* qrec,
                             HbAln* alns, cudaStream_t s);
