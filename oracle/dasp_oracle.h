/*
 * oracle/dasp_oracle.h — CPU restatement of the DASP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libdasp_b200.so) never links or calls it.
 *
 * What it restates (citations are into /root/reference):
 *   - host preprocessing of spmv_all: src/dasp_f64.h:499-1157 (FP64), src/dasp_f16.h:1029-1443 (FP16)
 *   - the stable descending length sort:  src/utils.h:128-160,196-203
 *   - the in-place exclusive scan:        src/mmio_highlevel.h:10-25
 *   - the meaning of every packed slot as read by the kernels: src/dasp_f64.h:77-484
 *   - "the reference's serial CSR result": the reference ships no CPU SpMV (its comparator is
 *     cuSPARSE, src/main_f64.cu:19-100); the serial CSR loop below defines it.
 *
 * Parity pin: checked bit-for-bit against the reference's own host code compiled from
 * /root/reference (oracle/_ref, see oracle/Makefile) and against tests/golden/ (digests produced
 * by that build).  The reference holds no golden vectors of its own.
 */
#ifndef DASP_ORACLE_H
#define DASP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { DASP_ORACLE_F64 = 0, DASP_ORACLE_F16 = 1 };

typedef struct dasp_oracle_layout {
    /* ---- scalars (names follow the reference's locals) ---- */
    int dtype, m, n, nnz;
    int row_long, row_block, row_zero;
    int short_row_1, short_row_3, short_row_2, short_row_4; /* after 1&3 pairing */
    int common_13, short_row_34;
    int rowloop, blocknum;
    int warp_number, BlockNum_long, fill0_nnz_long;
    int fill0_nnz_reg, nnz_irreg, origin_nnz_reg;
    int fill0_nnz_short, fill0_nnz_short13, fill0_nnz_short34, fill0_nnz_short22;
    int threadblock13, threadblock34, threadblock22;
    int nnz_short, nnz_long;
    int BlockNum, BlockNum_short_1, BlockNum_all, sumBlockNum;
    int fill0_nnz_irreg; /* FP16: nnz_irreg rounded up to even (allocation of irreg_val) */
    /* ---- arrays (owned; free with dasp_oracle_free) ---- */
    int *order_rid;      /* [m]            permuted index -> original row */
    int *long_rpt_new;   /* [row_long+1]   warp offset of each long row   */
    void *long_val;      /* [fill0_nnz_long] */
    int *long_cid;
    int *blockPtr;       /* [blocknum+1] */
    int *irreg_rpt;      /* [row_block+1] */
    void *irreg_val;     /* [fill0_nnz_irreg] (FP64: nnz_irreg) */
    int *irreg_cid;      /* [nnz_irreg] */
    void *reg_val;       /* [fill0_nnz_reg] */
    int *reg_cid;
    void *short_val;     /* [fill0_nnz_short] */
    int *short_cid;
} dasp_oracle_layout;

/* Host preprocessing. val is double[nnz] (F64) or IEEE half bits uint16_t[nnz] (F16). Returns 0. */
int dasp_oracle_preprocess(int dtype, int m, int n, int nnz, const int *rowptr, const int *colidx,
                           const void *val, double threshold, int block_longest,
                           dasp_oracle_layout *out);
void dasp_oracle_free(dasp_oracle_layout *L);

/* y[i] = sum_j val[j] * x[col[j]], sequential in CSR order, one double accumulator per row. */
void dasp_oracle_csr_spmv_f64(int m, const int *rowptr, const int *colidx, const double *val,
                              const double *x, double *y);
/* Row-parallel variant of the same loop (same per-row order => identical bits), for the CPU
 * baseline on all host cores. nthreads<=0: use every core. */
void dasp_oracle_csr_spmv_f64_mt(int m, const int *rowptr, const int *colidx, const double *val,
                                 const double *x, double *y, int nthreads);
/* Half inputs (bit patterns), double accumulation, double result (rounded by the caller). */
void dasp_oracle_csr_spmv_f16(int m, const int *rowptr, const int *colidx, const uint16_t *val,
                              const uint16_t *x, double *y);

/* Evaluate y (permuted order, length m, double) directly from the packed layout, i.e. what the
 * reference kernels compute from these arrays (src/dasp_f64.h:90-483, K1-K7/K11 of SURVEY §8a). */
void dasp_oracle_layout_spmv(const dasp_oracle_layout *L, const void *x, double *y_perm);

double dasp_oracle_half_to_double(uint16_t h);
uint16_t dasp_oracle_double_to_half(double d); /* round-to-nearest-even */

#ifdef __cplusplus
}
#endif
#endif
