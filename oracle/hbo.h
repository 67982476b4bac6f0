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

#ifdef __cplusplus
}
#endif
#endif
