/*
 * hite_b200 -- C-ABI of libhite_b200.so (hand-written sm_100a CUDA kernels behind it).
 *
 * Drop-in boundary for HiTE's dynamic boundary-adjustment hot path.  The reference (pure Python +
 * external aligners) has no FFI of its own; every entry point below cites the reference
 * function(s) (paths relative to the reference repo root) whose work it replaces.  The Python
 * host side (hite_b200/) binds these with ctypes; INTEGRATION.md shows the binding a HiTE
 * maintainer would add to module/Util.py.
 *
 * Conventions: extern "C", plain pointers + sizes, caller owns every host buffer, the library
 * owns device memory inside the context.  All functions return 0 on success and a negative
 * HB_E* code on failure; hb_last_error() gives the message.  No CPU fallback exists: without a
 * CUDA device hb_create fails with HB_ENODEV.  A context is single-threaded and bound to one
 * device + one stream (one context per process per GPU, mirroring one Python process per stage
 * in the reference, main.py:532-628).
 */
#ifndef HITE_B200_H
#define HITE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_EINVAL (-1)
#define HB_ENODEV (-2)
#define HB_ECUDA (-3)
#define HB_ENOMEM (-4)
#define HB_ESTATE (-5)

#define HB_TE_TIR 0      /* judge_boundary_v5  module/Util.py:9145  */
#define HB_TE_HELITRON 1 /* judge_boundary_v6  module/Util.py:9821  */
#define HB_TE_NON_LTR 2  /* judge_boundary_v9  module/Util.py:9483  */

#define HB_INFO_NONE 0
#define HB_INFO_NB 1  /* 'nb'  anchor not found       module/Util.py:9187 */
#define HB_INFO_FL1 2 /* 'fl1' <=1 full-length copy   module/Util.py:9225 */

typedef struct hb_ctx hb_ctx;

/* Result of judging one candidate: mirrors the (is_TE, info, cons_seq, copy_num) tuple returned
 * by judge_boundary_v5/v6/v9 plus the intermediate boundaries for diagnostics/parity. */
typedef struct {
    int32_t is_te;
    int32_t info;
    int32_t copy_num;                /* row_num of the last judge call (module/Util.py:9480) */
    int32_t align_start, align_end;  /* anchor columns (module/Util.py:9179-9180), -1 if none */
    int32_t homo_start, homo_end;    /* search_boundary_homo_v3 results, -1 if rejected */
    int32_t cons_len;
    int32_t trim_lead, trim_trail;   /* TIR terminal trim chosen at module/Util.py:9406-9413 */
    int32_t clean_cols;              /* columns left by remove_sparse_col_in_align_file */
    int32_t flags;
    int64_t cons_off;                /* offset of the consensus bytes in the cons_out buffer */
} hb_call;

/* Timings of the most recent call, milliseconds, CUDA events on the context's stream. */
typedef struct {
    float h2d_ms, kernel_ms, d2h_ms;
    float k_ms[8];       /* per-kernel split; meaning documented per entry point */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t launches;   /* kernels launched by the call */
    uint32_t pad;
} hb_timing;

int hb_version(void);
int hb_device_count(void);
int hb_create(int device, hb_ctx** out);
void hb_destroy(hb_ctx* ctx);
const char* hb_last_error(hb_ctx* ctx);
int hb_get_timing(hb_ctx* ctx, hb_timing* out);

/*
 * hb_vote_matrices -- judge n pre-aligned matrices (the parity entry; kernel (iii)).
 * Replaces, per candidate, remove_sparse_col_in_align_file (module/Util.py:10344-10405) followed
 * by judge_boundary_v5 (module/Util.py:9145-9480; te_type HB_TE_TIR) or judge_boundary_v6
 * (module/Util.py:9821-10159; HB_TE_HELITRON, plant/result_type ignored), i.e. the body of
 * is_TE_from_align_file after mafft (module/Util.py:10419-10432).
 *
 *   rows_ascii : concatenated R_i x C_i row-major ASCII matrices ('-' = gap), matrix i at mat_off[i]
 *   cand_ascii : concatenated candidate sequences, candidate i at cand_off[i] .. cand_off[i+1]
 *   result_type: 1 = 'cons' (column vote), 0 = first full-length row
 *   calls      : n results
 *   cons_out   : receives consensus bytes; capacity cons_cap >= sum(C_i) + 2n; candidate i's consensus
 *                is cons_out[calls[i].cons_off .. +cons_len)
 */
int hb_vote_matrices(hb_ctx* ctx, uint32_t n, const uint8_t* rows_ascii, const uint64_t* mat_off,
                     const int32_t* R, const int32_t* C, const uint8_t* cand_ascii,
                     const uint64_t* cand_off, int te_type, int plant, int result_type,
                     hb_call* calls, uint8_t* cons_out, uint64_t cons_cap);

/* hb_params.flags: the FiLTR fork of the loop (bin/FiLTR-main/src/Util.py:956-1043, :1315-1339) keeps the FIRST 100 members of a
 * file (extract_copies :1306) instead of the 100 longest and stops at the alignment -- its own get_both_ends_frame reads it */
#define HB_PRM_ROWS_FIRST 1 /* S9: the first max_rows kept copies in rank order, not the longest */
#define HB_PRM_NO_JUDGE 2   /* stop after the star MSA: hb_last_copies / hb_last_msa / hb_last_msa_trunc hold the result */

/* hb_call.flags */
#define HB_FLAG_NO_COPIES 1 /* no full-length copy survived: the reference writes no member file for such a
                               query (module/Util.py:8125-8137) and it silently drops out of the result */

#define HB_FLAG_TRUNC_REJECT 2 /* verdict comes from the truncated-member pass (module/Util.py:10442-10446) */

/*
 * hb_genome_load -- kernel (i), part 1: upload the genome (the contigs of the `-r reference` FASTA,
 * upper/lower case ASCII, read_fasta semantics module/Util.py:1650-1670) and pack it to 2 bit/base
 * + N mask in one padded coordinate space.  contig_off has n_contig+1 entries into `ascii`.
 * Replaces reading the reference into a Python dict (module/Util.py:8073) and feeds what
 * `minimap2 -d` (module/Util.py:7941) indexes.
 */
int hb_genome_load(hb_ctx* ctx, const uint8_t* ascii, const uint64_t* contig_off, uint32_t n_contig);

/*
 * hb_genome_contig_rank -- the order of the contig NAMES, which the "100 longest members" rule needs: the reference
 * subsets a member file of more than 100 copies with `sh tools/ready_for_MSA.sh <file> 100 100` (module/Util.py:10409-10412),
 * whose `sort -nk 2 -r` on the samtools .fai lines (tools/ready_for_MSA.sh:63) breaks a length tie by a reversed whole-line
 * comparison, i.e. by the member name `contig:start-end(strand)` bytewise descending.  rank[c] = position of contig c when the
 * contigs are sorted by (name + ':') bytewise ascending.  Optional: after hb_genome_load the rank is the contig index
 * (correct whenever the FASTA lists its contigs in that order).  Call it after hb_genome_load, before hb_adjust.
 */
int hb_genome_contig_rank(hb_ctx* ctx, const int32_t* rank, uint32_t n_contig);

/*
 * hb_index_build -- kernel (i), part 2: (k,w)-minimizer sketch of both strands, hand-written LSD
 * radix sort with shared-memory binning, 22-bit bucket table.  Replaces `minimap2 -I 20G -d ref.mmi`
 * (module/Util.py:7938-7947).  8 <= k <= 15, 1 <= w <= 16; HiTE-equivalent default k=15, w=10.
 * Timing: k_ms[0] sketch, k_ms[1] sort, k_ms[2] bucket table.
 */
int hb_index_build(hb_ctx* ctx, int k, int w);
int hb_genome_info(hb_ctx* ctx, uint64_t* total_padded, uint64_t* n_minimizers, uint64_t* contig_goff, uint32_t cap);
/* Test hook: device exclusive scan of n copies of `value`: *total = exact 64-bit sum (the size guards of hb_adjust and
 * hb_coarse_boundary compare it with their limits before anything is sized from the 32-bit offsets), *last = last offset. */
int hb_debug_scan_total(hb_ctx* ctx, uint64_t n, uint32_t value, uint64_t* total, uint32_t* last);

typedef struct {
    int32_t occ_cap;      /* occurrences used per minimizer (first occ_cap in position order); 512 */
    int32_t max_loci;     /* loci aligned per candidate (minimap2 -N 300, module/Util.py:7955); 300 */
    int32_t flank;        /* flanking_len, module/Util.py:8103-8105; 50 */
    int32_t max_rows;     /* "100 longest" of ready_for_MSA.sh, module/Util.py:10409-10412; 100 */
    int32_t te_type;      /* HB_TE_* */
    int32_t plant;        /* --plant */
    int32_t result_type;  /* 1 = 'cons' */
    int32_t trunc_gate;   /* 1 = truncated-member pass first (module/Util.py:8116-8124, :10442-10446); 1 */
    uint64_t tb_budget_bytes; /* HBM budget of the traceback stream, 0 = 16 GiB */
    int32_t overlap;      /* 0 / 1 = one batch on the context's stream; N >= 2 = N sub-batches (contiguous candidate
                           * ranges) software-pipelined over two CUDA streams: same results, stages overlap */
    int32_t flags;        /* HB_PRM_* bits, 0 for the boundary-adjustment path */
} hb_params;

/* One candidate x locus pair: what get_copies_minimap2 (module/Util.py:7977-8030) returns per SAM
 * record -- (subject, start, end, aligned_len, strand) -- plus the alignment summary. */
typedef struct {
    uint32_t cand;
    int32_t rank, rel, nseeds; /* rel: 0 '+', 1 '-' */
    int32_t score, i_s, i_e;   /* query span of the local alignment, inclusive */
    int32_t keep, row, contig; /* keep: passed the reference's copy rules; row: MSA row or -1 */
    int64_t j0, j_s, j_e;      /* band origin / target span in oriented padded coordinates */
    int64_t start1, end1;      /* 1-based inclusive contig coordinates (module/Util.py:8017-8018) */
    int32_t ext_len, trow;     /* trow: row in the truncated-member MSA or -1 */
} hb_copy;

/*
 * hb_adjust -- one `flank_region_align_v5`-equivalent round over n_cand candidates (module/Util.py:
 * 8032-8299 up to the per-candidate verdict): seed lookup + diagonal clustering (replaces minimap2
 * seeding/chaining, :7952-7961), banded affine Smith-Waterman per candidate x locus (replaces
 * minimap2 base alignment), the copy rules of :8003-8025 and :8095-8115, the 100-longest subset of
 * :10409-10412, a candidate-anchored star MSA (replaces mafft, :10416), the truncated-member gate of
 * run_find_members_v8 (:10439-10449), then remove_sparse_col_in_align_file + judge_boundary_v5 / v6 / v9
 * by te_type (:10419-10432).
 * calls[i] mirrors is_TE_from_align_file's tuple for candidate i; consensus bytes are fetched with
 * hb_last_cons (calls[i].cons_off indexes that buffer); *cons_bytes = its size.
 * Timing: k_ms[0] pack+sketch candidates, [1] seed/sort/cluster/rank, [2] extension+traceback,
 * [3] copy rules + MSA, [4] clean + judge, [5] compaction, [6] #pairs, [7] million DP steps.
 */
int hb_adjust(hb_ctx* ctx, uint32_t n_cand, const uint8_t* cand_ascii, const uint64_t* cand_off,
              const hb_params* prm, hb_call* calls, uint64_t* cons_bytes);
int hb_last_cons(hb_ctx* ctx, uint8_t* out, uint64_t cap);
/* the same bytes without a copy: *ptr points into the context's page-locked result buffer (what the device wrote by DMA),
 * valid until the next call on this context.  The reference hands the consensus on as Python strings in memory
 * (module/Util.py:10434-10437); a binding that wraps the view skips one pass over the bytes. */
int hb_last_cons_view(hb_ctx* ctx, const uint8_t** ptr, uint64_t* nbytes);
/* Page-lock (and release) a caller-owned host buffer that is handed to hb_adjust / hb_query_hsps / hb_tsd_candidates again
 * and again, e.g. the flattened candidate file of a stage that runs several rounds over it (judge_TIR_transposons.py:29-38):
 * the host->device copy then runs as one DMA at link speed instead of through the driver's staging buffer.  Optional --
 * every entry point accepts pageable memory.  HB_ECUDA when the range cannot be locked (the buffer stays usable). */
int hb_host_register(void* ptr, uint64_t nbytes);
int hb_host_unregister(void* ptr);
/* device-resident copy of the same bytes (valid until the next hb_adjust on this context): lets the
 * multi-GPU merge all-gather the records over NVLink without a host round trip */
int hb_last_cons_device(hb_ctx* ctx, const void** dptr, uint64_t* nbytes);
/* copies of the last hb_adjust, ordered by (candidate, rank); pass out=NULL to query the count */
int hb_last_copies(hb_ctx* ctx, hb_copy* out, uint64_t cap, uint64_t* n);

/*
 * hb_query_hsps -- library-vs-genome local alignments for the homology non-LTR module (module/judge_Other_transposons.py:45
 * -> multi_process_align_and_get_copies, module/Util.py:8382-8430: makeblastdb + blastn -outfmt 6 of 50-query files,
 * :7213-7245).  Runs the seeding and banded-extension stages of hb_adjust (same kernels, same hb_params: occ_cap,
 * max_loci) with the library sequences as queries and reports every positive local alignment that lies inside one
 * contig as the fields get_copies_v1 reads from outfmt 6 (:7032-7056); pident = 100 * n_match / aln_len.  The records
 * are kept in the context, ordered by (query, locus rank); fetch them with hb_last_query_hsps (out = NULL: count only).
 * Chaining them into copies is hb_query_copies.  Timing: k_ms[0] sketch queries, [1] seed/sort/cluster/rank,
 * [2] extension + traceback, [3] HSP records, [6] #pairs, [7] million DP steps.
 */
typedef struct {
    uint32_t query;          /* index of the query in the call's input */
    int32_t contig;          /* subject: contig index of hb_genome_load */
    int32_t qs, qe;          /* 1-based inclusive query span */
    int32_t ss, se;          /* 1-based inclusive subject span; ss > se = minus strand */
    int32_t n_match, aln_len;
    int32_t score, rank;
} hb_qhsp;
int hb_query_hsps(hb_ctx* ctx, uint32_t n_query, const uint8_t* query_ascii, const uint64_t* query_off, const hb_params* prm,
                  uint64_t* n_out);
int hb_last_query_hsps(hb_ctx* ctx, hb_qhsp* out, uint64_t cap, uint64_t* n);

/* ---- coarse boundary (module/coarse_boundary.py -> determine_repeat_boundary_v5, module/Util.py:4637) ---- */
typedef struct { int32_t subj, qs, qe, ss, se; } hb_hsp;      /* one HSP of a query segment against segment `subj`;
                                                                  1-based inclusive, ss > se = minus strand: the fields
                                                                  get_longest_repeats_v4 reads from outfmt 6 (module/Util.py:4130-4137) */
typedef struct { int32_t chr, pad; int64_t start; } hb_segment; /* segment `chr$start` of split_genome_chunks.py:47-52 */
typedef struct { int32_t chr, qseg; int64_t start, end; } hb_interval; /* 'chr:start-end' (0-based, end exclusive; module/Util.py:4354-4363) */

/*
 * hb_fmea_chain -- FMEA interval chaining: get_longest_repeats_v4 + process_all_seqs (module/Util.py:
 * 4122-4400, 4529-4563) for n_seg query segments; query segment g owns hsp[q_off[g] .. q_off[g+1]) in
 * "file order".  Bit-identical to the reference on the same HSP list.  Two-call: out == NULL returns the
 * count in *n_out.  Host code (see hb_fmea.cu for why); `threads` host threads over query segments.
 */
int hb_fmea_chain(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs, int skip_gap,
                  int max_len, int threads, hb_interval* out, uint64_t cap, uint64_t* n_out);
/* The same with the reference's BLAST result files made explicit: the reference packs the segment records into query files of
 * >= 1 000 000 bases (split_and_store_sequences, module/Util.py:4737, :4987-5012) and get_longest_repeats_v4 keeps ONE
 * rounded-position dict and ONE name dict per file (:4152-4155), so the records of a file -- a contig's short tail segment and
 * whatever follows it -- de-duplicate against each other.  file_of_seg[g] = file of segment g (non-decreasing).
 * hb_fmea_chain is the case "one record per file"; hb_coarse_boundary derives the files from its segment lengths. */
int hb_fmea_chain_files(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs,
                        const uint32_t* file_of_seg, int skip_gap, int max_len, int threads, hb_interval* out,
                        uint64_t cap, uint64_t* n_out);
/* Test hook: hb_fmea_chain_files with the files handed to the chaining pool in n_blocks separate submissions, the way
 * hb_coarse_boundary hands over the blocks of its GPU pipeline; the result must not depend on n_blocks or on `threads`. */
int hb_debug_fmea_blocks(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_seg, const hb_segment* segs,
                         const uint32_t* file_of_seg, int skip_gap, int max_len, int threads, uint32_t n_blocks,
                         hb_interval* out, uint64_t cap, uint64_t* n_out);

/*
 * hb_query_copies -- get_query_copies (module/Util.py:6828-7030; used by judge_Other_transposons.py:45 through
 * Util.py:8382 and by pan_recover_low_copy_TEs): chain the HSPs of ONE query per subject into full-length
 * copies (gap thresholds qthr/sthr, 200 in the reference), keep the longest chain per subject cluster, sort by
 * length, apply query/subject coverage and max_copy_num.  ident[i] is the HSP's percent identity (part of the
 * reference's visited-key).  Bit-identical to the reference on the same HSP list.  Host code (hb_fmea.cu).
 */
typedef struct { int32_t subj, start, end, qlen, strand; } hb_query_copy; /* (subject, start, end, len, 0 '+' / 1 '-') :7015 */
int hb_query_copies(const hb_hsp* hsp, const double* ident, uint32_t n, int query_len, const int32_t* subj_len,
                    double query_coverage, double subject_coverage, int qthr, int sthr, int max_copy_num,
                    hb_query_copy* out, uint32_t cap, uint32_t* n_out);

/*
 * Known-TE genome masking: mask_genome_intactTE (module/Util.py:6389-6431; the step before the all-vs-all search in
 * determine_repeat_boundary_v5 :4654 and before LTR detection, judge_LTR_transposons.py:111).  The reference maps the
 * TE library to the genome (minimap2 -> PAF -> outfmt-6, :6318-6374), chains every query's records into full-length
 * copies (get_full_length_copies_from_blastn_v1, :5907-6128, threshold 0.95) and overwrites each copy with N (:6414-6420).
 * Here: hb_query_hsps gives the records, hb_full_length_copies is the chaining (host code, bit-identical to the reference
 * on the same HSP list: hsp[i].subj = subject index, spans 1-based inclusive, ss > se = minus strand as in outfmt 6),
 * hb_genome_mask N-fills [start, end) (0-based, end exclusive: the 'chr:start-end' of :6107) in the context's packed
 * genome -- clipped to the contig like the numpy slice -- and invalidates the index (call hb_index_build again);
 * *n_masked = bases that were not N before.  hb_genome_export returns a contig of the packed genome as ASCII
 * (A C G T N; letters outside ACGT were read as N by hb_genome_load).
 */
int hb_full_length_copies(const hb_hsp* hsp, uint32_t n, int query_len, double full_length_threshold,
                          hb_query_copy* out, uint32_t cap, uint32_t* n_out);
/* Batched forms over the records of many queries -- query g owns hsp[q_off[g] .. q_off[g+1]) and has length query_len[g];
 * `threads` host threads over queries.  Copies come back grouped by query (out_query[i] = query of out[i]); cap = number of
 * records always suffices.  hb_query_copies_batch chains with subject_coverage = 0 (what get_copies_v1 passes, :7032). */
int hb_full_length_copies_batch(const hb_hsp* hsp, const uint64_t* q_off, uint32_t n_query, const int32_t* query_len,
                                double full_length_threshold, int threads, hb_query_copy* out, uint32_t* out_query,
                                uint64_t cap, uint64_t* n_out);
int hb_query_copies_batch(const hb_hsp* hsp, const double* ident, const uint64_t* q_off, uint32_t n_query,
                          const int32_t* query_len, double query_coverage, int qthr, int sthr, int max_copy_num, int threads,
                          hb_query_copy* out, uint32_t* out_query, uint64_t cap, uint64_t* n_out);
int hb_genome_mask(hb_ctx* ctx, uint32_t n, const int32_t* contig, const int64_t* start, const int64_t* end, uint64_t* n_masked);
int hb_genome_export(hb_ctx* ctx, uint32_t contig, uint8_t* out_ascii, uint64_t cap);

/*
 * hb_coarse_boundary -- the all-vs-all step of process_blast_alignments (module/Util.py:4724-4780: n^2 blastn
 * runs of 1 Mbp segments) as one self-join of the minimizer index on the GPU (seed pairs -> radix sort ->
 * exact-diagonal runs = HSPs), followed by hb_fmea_chain.  Uses the genome + index of the context; segments
 * are seg_len-long pieces of each contig (chrom_seg_length, main.py:41).  Two-call like hb_fmea_chain.
 * Timing: k_ms[0] pair count+fill, [1] sort, [2] HSP extraction, [3] host FMEA (summed over its blocks), [4] wall of the
 * whole stage (the GPU half of block b + 1 runs while the host chains block b), [6] #HSPs, [7] million seed pairs.
 */
int hb_coarse_boundary(hb_ctx* ctx, int64_t seg_len, int skip_gap, int max_len, int occ_cap, int threads,
                       hb_interval* out, uint64_t cap, uint64_t* n_out);
/*
 * hb_coarse_boundary_part -- one part of the same stage sharded over n_parts processes / GPUs (the reference
 * fans the n^2 segment pairs out over a process pool, module/Util.py:4746-4748).  Every part holds the whole genome + index
 * and owns a contiguous block of the query-segment FILES (the unit FMEA de-duplicates in): the files
 * [F * part / n_parts, F * (part + 1) / n_parts).  Only the minimizers of those segments are paired (against the whole index) and
 * chained; the parts' interval lists concatenated in part order are exactly hb_coarse_boundary's list.  out = NULL (or a
 * short buffer) still runs the stage and keeps the result in the context: fetch it with hb_last_intervals, no second run.
 * Inside a part the files go through the GPU in up to 8 blocks, software-pipelined with the host chaining of the blocks already
 * read back; the result does not depend on the block count (test hook: the environment variable HB_COARSE_BLOCKS forces it).
 * Timing: k_ms[0] seed pairs, [1] sort, [2] runs + base-level HSPs, [3] host FMEA span (first block handed over -> last task
 * done), [4] wall of the pipelined stage, [5] FMEA task time per host thread, [6] #HSPs, [7] million seed pairs.
 */
int hb_coarse_boundary_part(hb_ctx* ctx, int64_t seg_len, int skip_gap, int max_len, int occ_cap, int threads,
                            uint32_t part, uint32_t n_parts, hb_interval* out, uint64_t cap, uint64_t* n_out);
int hb_last_intervals(hb_ctx* ctx, hb_interval* out, uint64_t cap, uint64_t* n);
/* the HSPs of the last hb_coarse_boundary, grouped by query segment (q_off has n_seg+1 entries) */
int hb_last_hsps(hb_ctx* ctx, hb_hsp* out, uint64_t cap, uint64_t* n, uint64_t* q_off, hb_segment* segs,
                 uint32_t seg_cap, uint32_t* n_seg);

/*
 * hb_tsd_candidates -- TSD-bounded variants of n flanked candidates, the step right before the boundary-adjustment
 * path in the TIR module: search_confident_tir_v4 (module/Util.py:7734-7845, driven per candidate by
 * search_confident_tir_batch :7358-7376 with raw_tir_start = flanking_len + 1, raw_tir_end = len - flanking_len,
 * tsd_search_distance = flanking_len) plus the classification of get_short_tir_contigs (:7297-7334).
 * Candidate i is seq_ascii[seq_off[i] .. seq_off[i+1]) (upper case, < 2^20 bytes); raw_start1 / raw_end1 are the
 * 1-based boundaries the reference receives; 1 <= tsd_search_distance <= 128.  out holds 100 records per candidate
 * (the reference's max_top_num), n_out[i] of them valid, in the reference's output order: variant = seq[start .. end]
 * (0-based, inclusive), TSD = the k bytes after `end`, distance = |start - (raw_start1-1)| + |end - (raw_end1-1)|,
 * idx = position in the distance-sorted candidate list (the '-C_<idx>-' of the reference's record name; among equal
 * distances the reference's order is Python set order, here (start, end, k)), short_class = 0 none / 1 hAT /
 * 2 Mutator / 3 CACTA / 4 CCC (which records get_short_tir_contigs keeps).  One warp per candidate.
 */
typedef struct { int32_t k, start, end, dist, idx, short_class; } hb_tsd_rec;
int hb_tsd_candidates(hb_ctx* ctx, uint32_t n, const uint8_t* seq_ascii, const uint64_t* seq_off, const int32_t* raw_start1,
                      const int32_t* raw_end1, int32_t tsd_search_distance, int32_t plant, hb_tsd_rec* out, uint32_t* n_out);

/*
 * hb_nonltr_candidates -- search_polyA_TSD (module/Util.py:10915-11006, with find_nearest_polyA / polyT :10865 / :10903 and
 * find_nearest_tandem :9772-9819) for n flanked candidates: the scan get_candidate_non_LTR (:11008-11045) runs on every
 * 100-8000 bp record of longest_repeats.flanked.fa right before the non-LTR boundary adjustment
 * (judge_Non_LTR_transposons.py:34).  end_5_window_size = 25 and TSD_sizes = 8..20 as that caller fixes them (:11009-11010).
 * Per candidate: found (a TSD of 8-20 bp, one edit allowed), direct 0 none / 1 '+' / 2 '-', the 3' / 5' ends, the TSD length
 * (TSD = seq[end_3 : end_3 + tsd_k] for '+', seq[end_3 - tsd_k : end_3] for '-') and non_ltr_seq = seq[seq_start : seq_end]
 * (to be reverse-complemented for '-').  Sequences are upper-case ASCII, < 2^20 bytes each; 0 <= flanking_len <= 1000.
 */
typedef struct { int32_t found, direct, end_5, end_3, tsd_k, seq_start, seq_end; } hb_polya_rec;
int hb_nonltr_candidates(hb_ctx* ctx, uint32_t n, const uint8_t* seq_ascii, const uint64_t* seq_off, int32_t flanking_len,
                         hb_polya_rec* out);

/*
 * hb_itr_search -- HB-itr, the terminal-inverted-repeat matcher that stands in for the closed binary `tools/itrsearch`
 * (run as `itrsearch -i 0.7 -l 7`, module/Util.py:216-224; callers search_confident_tir_batch_v1 :6563 / :6575 and
 * remove_no_tirs :13907).  Per record: the 5' end against the reverse complement of the 3' end, both anchored at the
 * record's ends, window min(max_window, len / 2) (the binary: 500), match +10, mismatch -16, gap of n bases -(32 + 32 n),
 * optimum over all end cells; found = score - p > 0 (p = 64 / 32 / 0 for len5 == / > / < len3, the binary's "Quality" as
 * fitted to its outputs) and m / (m + x) >= min_identity and len5 >= min_len.
 * len5 / len3 = bases of the 5' / 3' arm (the binary's header reads `ITR(1,len5)..(L,L-len3+1) Length itr=len5-1`).
 * The binary's own search is a heuristic below this optimum in places: parity unpinned; measured agreement in
 * tests/golden/itrsearch.json.gz.  Records of < 2^20 bytes; any case; N never matches.
 */
typedef struct { int32_t found, len5, len3, score, n_match, n_mismatch, n_gap; } hb_itr_rec;
int hb_itr_search(hb_ctx* ctx, uint32_t n, const uint8_t* seq_ascii, const uint64_t* seq_off, double min_identity, int32_t min_len,
                  int32_t max_window, hb_itr_rec* out);

/* loci aligned per candidate in the last hb_adjust (<= max_loci each): the cost proxy (x candidate length) a multi-GPU caller
 * rebalances the next round's shards with -- the stages run the same candidates through several rounds (judge_TIR_transposons.py:29-38) */
int hb_last_pair_counts(hb_ctx* ctx, uint32_t* out, uint32_t cap);

/* parity/debug getters: the sorted minimizer index (hash<<33 | pos<<1 | strand) and the raw
 * (pre-clean) MSA of one candidate of the last hb_adjust as ASCII */
int hb_debug_index(hb_ctx* ctx, uint64_t* keys_out, uint64_t cap);
int hb_last_msa(hb_ctx* ctx, uint32_t cand, uint8_t* out_ascii, uint64_t cap, int32_t* R, int32_t* C);
/* the truncated-member matrix of the candidate (rows = kept copies longer than 1000 bp, cells outside each row's first and
 * last 500 bases blanked; module/Util.py:8116-8124): what the reference aligns first when copies are long.  R = 0 if none. */
int hb_last_msa_trunc(hb_ctx* ctx, uint32_t cand, uint8_t* out_ascii, uint64_t cap, int32_t* R, int32_t* C);

#ifdef __cplusplus
}
#endif
#endif /* HITE_B200_H */
