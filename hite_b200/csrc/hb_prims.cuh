// hite_b200 -- device-wide primitives (hb_prims.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

size_t hb_scan_tmp_words(size_t n);
cudaError_t hb_exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_tmp,
                                  uint64_t* d_total, cudaStream_t s);
size_t hb_sort_tmp_words(size_t n);
cudaError_t hb_radix_sort_u64(uint64_t* d_keys, uint64_t* d_alt, size_t n, int lo_bit, int hi_bit,
                              uint32_t* d_tmp, uint64_t** d_result, uint32_t* launches, cudaStream_t s);
