// hite_b200 -- TSD-bounded variant generation launcher (hb_tsd.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/hite_b200.h"

size_t hb_tsd_smem_bytes(int d);
cudaError_t hb_tsd_launch(const uint8_t* seqs, const uint64_t* seq_off, const int32_t* raw_start1, const int32_t* raw_end1,
                          uint32_t n_cand, int d, int plant, hb_tsd_rec* out, uint32_t* n_out, cudaStream_t s);
cudaError_t hb_polya_tsd_launch(const uint8_t* seqs, const uint64_t* seq_off, uint32_t n_cand, int flank, hb_polya_rec* out,
                                cudaStream_t s);
