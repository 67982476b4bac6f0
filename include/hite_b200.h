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
 * by judge_boundary_v5 (module/Util.py:9145-9480; te_type HB_TE_TIR), i.e. the body of
 * is_TE_from_align_file after mafft (module/Util.py:10419-10432).
 *
 *   rows_ascii : concatenated R_i x C_i row-major ASCII matrices ('-' = gap), matrix i at mat_off[i]
 *   cand_ascii : concatenated candidate sequences, candidate i at cand_off[i] .. cand_off[i+1]
 *   result_type: 1 = 'cons' (column vote), 0 = first full-length row
 *   calls      : n results
 *   cons_out   : receives consensus bytes; capacity cons_cap >= sum(C_i); candidate i's consensus
 *                is cons_out[calls[i].cons_off .. +cons_len)
 */
int hb_vote_matrices(hb_ctx* ctx, uint32_t n, const uint8_t* rows_ascii, const uint64_t* mat_off,
                     const int32_t* R, const int32_t* C, const uint8_t* cand_ascii,
                     const uint64_t* cand_off, int te_type, int plant, int result_type,
                     hb_call* calls, uint8_t* cons_out, uint64_t cons_cap);

#ifdef __cplusplus
}
#endif
#endif /* HITE_B200_H */
