// hite_b200 -- copy selection + star-MSA launchers (hb_msa.cu)
#pragma once
#include "hb_seed.cuh"
#include "hb_common.cuh"
#include "../../include/hite_b200.h"

cudaError_t hb_copy_filter_launch(const HbLocus* pairs, uint32_t n_pairs, const uint32_t* clen, const uint64_t* goff,
                                  const uint64_t* glen, uint32_t n_contig, uint64_t tot, int flank, HbAln* alns,
                                  cudaStream_t s);
// S8b: colinear partial alignments of one copy merged into the edit record of their best-ranked piece
cudaError_t hb_chain_launch(const HbLocus* pairs, const uint32_t* pair_off, uint32_t n_cand, const uint32_t* clen,
                            const uint32_t* qrec_unit_off, uint32_t* qrec, HbAln* alns, cudaStream_t s);
cudaError_t hb_copy_select_launch(const HbLocus* pairs, const uint32_t* pair_off, uint32_t n_cand, int max_rows,
                                  int trunc_gate, int rows_first, const int32_t* crank, HbAln* alns, uint32_t* n_rows,
                                  uint32_t* n_trows, cudaStream_t s);
cudaError_t hb_ins_max_launch(const HbLocus* pairs, uint32_t n_pairs, const HbAln* alns, const uint32_t* qrec_unit_off,
                              const uint32_t* qrec, const uint64_t* coff, uint32_t* ins, cudaStream_t s);
cudaError_t hb_layout_launch(const uint64_t* coff, const uint32_t* clen, uint32_t n_cand, int flank, uint32_t* ins,
                             uint32_t* n_cols, cudaStream_t s);
// column map ((q + flank) << 8 | slot per MSA column) of every candidate, at colq + col_off[c]
cudaError_t hb_colmap_launch(const uint64_t* coff, const uint32_t* clen, uint32_t n_cand, int flank, const uint32_t* pre,
                             const uint32_t* n_cols, const uint32_t* col_off, uint32_t* colq, cudaStream_t s);
cudaError_t hb_msa_write_launch(const HbLocus* pairs, uint32_t n_pairs, const HbAln* alns, const uint32_t* qrec_unit_off,
                                const uint32_t* qrec, const uint32_t* n_cols, const uint32_t* col_off,
                                const uint32_t* colq, const uint64_t* mat_off, const uint32_t* gseq2,
                                const uint32_t* gnmask, uint64_t tot, int flank, int trunc, uint8_t* mat,
                                cudaStream_t s);
// one hb_qhsp per pair (contig = -1 when the pair has no positive alignment inside one contig)
cudaError_t hb_qhsp_launch(const HbLocus* pairs, uint32_t n_pairs, const HbAln* alns, const uint32_t* qrec_unit_off,
                           const uint32_t* qrec, const uint32_t* cseq2, const uint32_t* cnmask, const uint64_t* cgoff,
                           const uint32_t* gseq2, const uint32_t* gnmask, const uint64_t* goff, const uint64_t* glen,
                           uint32_t n_contig, uint64_t tot, hb_qhsp* out, cudaStream_t s);
// star MSA + sparse-column removal fused: the clean matrices of all judge tasks straight from the edit records, a CTA per
// full task also producing its truncated-member task trunc_of_full[t] (-1: none); calls[].clean_cols is set for both.
// The raw matrix is never written.  planes != nullptr: the six bit-planes of every kept column over all rows of each matrix
// are written too (plane_off of the task; hb_vote_launch is then told they are prebuilt).
cudaError_t hb_msa_clean_launch(const HbVoteTask* tasks, const uint32_t* order /* n_full task ids, CTA b takes order[b] */,
                                const uint32_t* task_cand, const int32_t* trunc_of_full, uint32_t n_full,
                                const HbLocus* pairs, const uint32_t* pair_off, const HbAln* alns, const uint32_t* qrec_unit_off,
                                const uint32_t* qrec, const uint32_t* col_off, const uint32_t* colq, const uint32_t* gseq2,
                                const uint32_t* gnmask, uint64_t tot, int flank, uint8_t* clean, uint32_t* planes, HbVoteCall* calls,
                                cudaStream_t s);
