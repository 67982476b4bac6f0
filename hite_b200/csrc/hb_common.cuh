// hite_b200 -- shared device/host declarations for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define HB_FULL 0xffffffffu

// symbol codes used on device: 0..3 = ACGT, 4 = N (any other byte), 5 = gap '-'
#define HB_SYM_N 4
#define HB_SYM_GAP 5

__host__ __device__ __forceinline__ int hb_sym_of_ascii(uint8_t c) {
    if (c == '-') return HB_SYM_GAP;
    switch (c & 0xDF) {  // upper-case letters (read_fasta upper-cases, reference Util.py:1665)
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return HB_SYM_N;
    }
}
__host__ __device__ __forceinline__ char hb_ascii_of_sym(int s) {
    return "ACGTN-"[s];
}

// ---- per-candidate work descriptor of the vote/judge kernels (device resident) ------------
struct HbVoteTask {
    uint64_t raw_off;    // byte offset of the R x C input matrix (row-major, stride C)
    uint64_t clean_off;  // byte offset of the cleaned matrix (row-major, stride C, symbol codes)
    uint64_t plane_off;  // uint32 offset of the bit-planes  P[sym 0..4][word 0..3][col], stride C
    uint64_t cons_off;   // byte offset of the consensus output (capacity C)
    uint64_t match_off;  // entry offset of the anchor match scratch (capacity 2*(C+2))
    uint64_t cand_off;   // byte offset of the candidate sequence (ASCII)
    uint64_t anchor_off; // int offset of the per-row anchor scratch (capacity 2*R), Helitron only
    int32_t cand_len;
    int32_t R, C;
    int32_t te_type;     // 0 tir, 1 helitron, 2 non_ltr
    int32_t plant;
    int32_t result_type; // 1 = 'cons'
    int32_t raw_is_ascii;
    int32_t pad;
};

struct HbVoteCall {      // mirrors hb_call in include/hite_b200.h
    int32_t is_te;
    int32_t info;        // 0 '', 1 'nb', 2 'fl1'
    int32_t copy_num;
    int32_t align_start, align_end;
    int32_t homo_start, homo_end;
    int32_t cons_len;
    int32_t trim_lead, trim_trail;
    int32_t clean_cols;
    int32_t flags;
    int64_t cons_off;    // offset of the (trimmed) consensus in the caller's cons buffer
};

#define HB_CUDA_OK(expr)                                                        \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) return hb_fail(ctx, #expr, _e, __FILE__, __LINE__); \
    } while (0)
