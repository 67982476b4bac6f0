/* TEST INFRASTRUCTURE (oracle) -- NOT part of the product path.  See hbo_judge.c. */
#ifndef HBO_H
#define HBO_H
#ifdef __cplusplus
extern "C" {
#endif

#define HBO_INFO_NONE 0
#define HBO_INFO_NB 1  /* 'nb'  : anchor not found      (Util.py:9187) */
#define HBO_INFO_FL1 2 /* 'fl1' : <= 1 full-length copy (Util.py:9225) */

typedef struct {
    int is_te;
    int info;
    int copy_num;
    int align_start, align_end;
    int homo_start, homo_end;
    int cons_len;
    int trim_lead, trim_trail;
} hbo_call;

int hbo_levenshtein(const char *a, int la, const char *b, int lb);
int hbo_find_near_matches(const char *p, int m, const char *s, int n, int d,
                          int *first_start, int *first_end, int *last_start, int *last_end);
int hbo_remove_sparse_cols(const char *m, int R, int C, char *out, int *C_out);
int hbo_search_boundary_homo_v3(int valid_col_threshold, int pos, const char *const *rows, int R, int C,
                                int is_start, double thr, int int_win, int out_win, const void *colmaps);
int hbo_tsd_search_v5(const char *row, int C, int bstart, int bend, int plant);
int hbo_find_anchor(const char *cand, int cand_len, const char *m, int R, int C,
                    int *align_start, int *align_end);
int hbo_judge_v5(const char *cand, int cand_len, const char *m, int R, int C, int plant,
                 int result_type_cons, hbo_call *out, char *cons);

int hbo_search_boundary_homo_v4(int vthr, int pos, const char *const *rows, int R, int C, int is_start,
                                double thr, double int_thr, double out_thr, int int_win, int out_win,
                                const void *colmaps, int *boundary);
int hbo_judge_v6(const char *cand, int cand_len, const char *m, int R, int C, hbo_call *out, char *cons);

int hbo_judge_v9(const char *cand, int cand_len, const char *m, int R, int C, hbo_call *out, char *cons);

/* ---- HB-align restatement (hbo_align.c) ---- */
#include <stdint.h>
typedef struct hbo_genome hbo_genome;
typedef struct { int32_t score, i_s, i_e, pad; int64_t j_s, j_e; } hbo_aln;   /* j: oriented coords */
typedef struct { uint32_t cand, rel, nseeds, rank; int64_t dmin, dmax, j0; } hbo_locus;
typedef struct {
    int32_t rank, rel, score, i_s, i_e, contig;
    int64_t j_s, j_e;        /* oriented global coords, inclusive */
    int64_t start1, end1;    /* 1-based inclusive forward coords inside the contig (Util.py:8017-8018) */
    int32_t keep, ext_len;
} hbo_copy;

hbo_genome *hbo_genome_create(const uint8_t *ascii, const uint64_t *off, uint32_t n);
void hbo_genome_free(hbo_genome *g);
void hbo_genome_set_tie_rule(hbo_genome *g, int tie_rule, const int32_t *contig_rank);
uint64_t hbo_genome_total(const hbo_genome *g);
uint64_t hbo_genome_contig_off(const hbo_genome *g, uint32_t c);
uint64_t hbo_minimizers_codes(const uint8_t *code, uint64_t tot, int k, int w, uint64_t *out, uint64_t cap);
uint64_t hbo_minimizers(const hbo_genome *g, int k, int w, uint64_t *out, uint64_t cap);
void hbo_sort_u64(uint64_t *keys, uint64_t n);
int hbo_align_pair(const hbo_genome *g, const uint8_t *q, int L, int rel, int64_t j0, int half_band, hbo_aln *out, uint32_t *qrec);
uint64_t hbo_find_loci(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w,
                       const uint8_t *cand_codes, const uint64_t *coff, uint32_t n_cand,
                       int occ_cap, int max_loci, hbo_locus *out, uint64_t cap);
int hbo_candidate_msa(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w,
                      const uint8_t *cand_ascii, int L, int occ_cap, int max_loci, int flank, int max_rows,
                      hbo_copy *copies, int *n_copies, uint8_t *matrix, uint64_t matrix_cap, int *R_out, int *C_out);

/* ---- FMEA restatement (hbo_fmea.c) ---- */
typedef struct { int32_t subj, qs, qe, ss, se; } hbo_hsp;            /* 1-based inclusive, s_start > s_end = reverse */
typedef struct { int32_t chr, pad; int64_t start, end; } hbo_interval; /* 'chr:start-end', 0-based half open */
int hbo_fmea(const hbo_hsp *h, int n, int q_chr, int64_t q_seg_start, int self_subj,
             const int32_t *subj_chr, const int64_t *subj_seg_start, int n_subj,
             int skip_gap, int max_len, hbo_interval *out, int cap);

void *hbo_fmea_set_new(void);
void hbo_fmea_set_free(void *set);
int hbo_fmea_shared(const hbo_hsp *h, int n, int q_chr, int64_t q_seg_start, int self_subj,
                    const int32_t *subj_chr, const int64_t *subj_seg_start, int n_subj,
                    int skip_gap, int max_len, void *shared_set, hbo_interval *out, int cap);

/* ---- HB-hsp restatement (hbo_hsp.c) ---- */
typedef struct { int32_t qseg, subj, qs, qe, ss, se; } hbo_hsp_rec;
uint64_t hbo_self_hsps(const hbo_genome *g, const uint64_t *mm_pos, const uint64_t *idx, uint64_t n, uint64_t tot, int k,
                       const uint64_t *seg_goff, const uint64_t *seg_len, int n_seg, int occ_cap, hbo_hsp_rec *out, uint64_t cap);
int hbo_genome_code(const hbo_genome *g, uint64_t pos);     /* 0..3 ACGT, 4 N / pad / outside */

typedef struct { int32_t subj, start, end, qlen, strand; } hbo_copy_out;   /* (subject, start, end, len, '+'/'-') :7015 */
int hbo_query_copies(const hbo_hsp *h, const double *ident, int n, int query_len, const int32_t *subj_len, int n_subj,
                     double query_coverage, double subject_coverage, int qthr, int sthr, int max_copy_num,
                     hbo_copy_out *out, int cap);

int hbo_full_length_copies_v1(const hbo_hsp *h, int n, int query_len, double full_length_threshold, hbo_copy_out *out, int cap);

typedef struct { int32_t is_te, info, copy_num, cons_len; } hbo_result;
int hbo_run_batch(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w, const uint8_t *cand_ascii,
                  const uint64_t *coff, uint32_t n_cand, int occ_cap, int max_loci, int flank, int max_rows, int plant,
                  int threads, hbo_result *out);

/* ---- library-vs-genome HSPs (hbo_align.c): checker of hb_query_hsps ---- */
typedef struct { int32_t rank, contig, qs, qe, ss, se, n_match, aln_len, score; } hbo_qhsp;
int hbo_query_hsps(const hbo_genome *g, const uint64_t *idx, uint64_t n_idx, int k, int w, const uint8_t *query_ascii,
                   int L, int occ_cap, int max_loci, hbo_qhsp *out, int cap);

/* ---- TSD-bounded variant generation (hbo_tsd.c): search_confident_tir_v4 + get_short_tir_contigs ---- */
typedef struct { int32_t k, start, end, dist, idx, idx_lo, idx_hi, short_class; } hbo_tsd_rec;
int hbo_short_tir_class(const char *tir, int len, int tsd_len, int plant);
int hbo_search_confident_tir_v4(const char *seq, int n, int raw_start1, int raw_end1, int d, int plant,
                                hbo_tsd_rec *out, int cap);

/* search_polyA_TSD: direct 0 none / 1 '+' / 2 '-'; non_ltr_seq = seq[seq_start:seq_end] (reverse-complemented when '-');
   the TSD is seq[end_3 : end_3 + tsd_k] ('+') or seq[end_3 - tsd_k : end_3] ('-') */
typedef struct { int32_t found, direct, end_5, end_3, tsd_k, seq_start, seq_end; } hbo_polya_rec;
int hbo_search_polyA_TSD(const char *seq, int n, int flanking_len, hbo_polya_rec *out);

#ifdef __cplusplus
}
#endif


/* ---- HB-itr (hbo_tsd.c): checker of hb_itr_search; out[7] = found, len5, len3, score, n_match, n_mismatch, n_gap ---- */
void hbo_itr_search(const char *s, int len, double min_ident, int min_len, int max_w, int32_t *out);

#endif
