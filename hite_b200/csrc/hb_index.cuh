// hite_b200 -- kernel (i) launchers (hb_index.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HB_PAD 64
#define HB_BUCKET_BITS 22

cudaError_t hb_pack_launch(const uint8_t* d_ascii, const uint64_t* d_src_off, const uint64_t* d_goff,
                           const uint64_t* d_glen, uint32_t n_contig, uint64_t tot, uint32_t* d_seq2,
                           uint32_t* d_nmask, cudaStream_t s);
uint32_t hb_minimizer_ctas(uint64_t tot);
cudaError_t hb_minimizer_count_launch(const uint32_t* d_seq2, const uint32_t* d_nmask, uint64_t tot, int k, int w,
                                      uint32_t* d_cta_count, cudaStream_t s);
cudaError_t hb_minimizer_write_launch(const uint32_t* d_seq2, const uint32_t* d_nmask, uint64_t tot, int k, int w,
                                      const uint32_t* d_cta_base, uint64_t* d_out, cudaStream_t s);
// bucket[b] (b = 0 .. nb) = first entry of the sorted index whose key >> shift is >= b
cudaError_t hb_bucket_bounds_launch(const uint64_t* d_idx, uint64_t n, int shift, uint32_t nb, uint32_t* d_bucket, cudaStream_t s);
// N-fill of n intervals [start, end) (0-based, contig coordinates) in the packed genome's N mask; *d_n_new += newly masked bases
cudaError_t hb_mask_fill_launch(const int32_t* d_contig, const int64_t* d_start, const int64_t* d_end, uint32_t n,
                                const uint64_t* d_goff, const uint64_t* d_glen, uint32_t n_contig, uint32_t* d_nmask,
                                unsigned long long* d_n_new, cudaStream_t s);
// ASCII (A C G T N) of len bases of the padded space starting at g0
cudaError_t hb_export_launch(const uint32_t* d_seq2, const uint32_t* d_nmask, uint64_t g0, uint64_t len, uint8_t* d_out, cudaStream_t s);
