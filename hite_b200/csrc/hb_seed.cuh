// hite_b200 -- kernel (ii) seeding launchers (hb_seed.cu) and the records shared with hb_extend.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HB_SEED_BUCKET_BITS 22
#define HB_DIAG_OFF 65536
#define HB_DGAP 16
#define HB_MIN_SEEDS 3
#define HB_SPLIT_SPREAD 47

struct HbLocus {          // one ranked locus of a candidate (HB-align S6)
    uint32_t cand, rel, nseeds, rank;
    int64_t dmin, dmax, j0;
};

struct HbAln {            // per candidate x locus pair, result of the banded alignment
    int32_t score, i_s, i_e, keep;
    int64_t j_s, j_e;     // oriented global coordinates, inclusive
    int32_t contig, ext_len;
    int64_t start1, end1; // 1-based inclusive forward coordinates inside the contig
    int32_t row;          // row in the full MSA or -1
    int32_t trow;         // row in the truncated (first 500 + last 500) MSA or -1  (Util.py:8116-8124)
};

cudaError_t hb_seed_count_launch(const uint64_t* cm, uint32_t n_cm, const uint64_t* cgoff, uint32_t n_cand,
                                 const uint64_t* idx, uint64_t n_idx, const uint32_t* bucket, int occ_cap,
                                 uint32_t* m_cand, uint32_t* m_qidx, uint32_t* m_lo, uint32_t* m_cnt, cudaStream_t s);
cudaError_t hb_seed_fill_launch(const uint64_t* cm, uint32_t n_cm, const uint64_t* cgoff, const uint64_t* idx,
                                uint64_t tot, int k, const uint32_t* m_cand, const uint32_t* m_qidx,
                                const uint32_t* m_lo, const uint32_t* m_cnt, const uint32_t* m_off, uint64_t* hits,
                                cudaStream_t s);
cudaError_t hb_cluster_flag_launch(const uint64_t* hits, uint64_t n, uint32_t* flag, cudaStream_t s);
cudaError_t hb_cluster_split_launch(const uint64_t* hits, uint64_t n, const uint32_t* flag, const uint32_t* cid,
                                    const uint32_t* cl_start, uint32_t n_cl, uint32_t* flag2, cudaStream_t s);
cudaError_t hb_cluster_start_launch(const uint32_t* flag, const uint32_t* cid, uint64_t n, uint32_t* cl_start,
                                    cudaStream_t s);
cudaError_t hb_rank_select_launch(const uint64_t* hits, uint64_t n_hits, const uint32_t* cl_start, uint32_t n_cl,
                                  uint32_t* cand_first, uint32_t n_cand, int max_loci, HbLocus* loci,
                                  uint32_t* n_loci, cudaStream_t s);
cudaError_t hb_pair_emit_launch(const HbLocus* loci, const uint32_t* n_loci, const uint32_t* pair_off,
                                uint32_t n_cand, int max_loci, HbLocus* pairs, cudaStream_t s);
