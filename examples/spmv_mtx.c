/*
 * examples/spmv_mtx.c — the reference's command line (./spmv_double A.mtx, ./spmv_half A.mtx: src/main_f64.cu:102-168)
 * written against the C ABI of this repository, as an integration example:
 *
 *   gcc -O2 -Iinclude examples/spmv_mtx.c -Ldasp_b200 -ldasp_b200 -Wl,-rpath,$PWD/dasp_b200 -lm -o spmv_mtx
 *   ./spmv_mtx A.mtx [reps] [-ones]
 *
 * Like the reference's main it reads the Matrix Market file (same reader semantics), sets x and, as the reference does
 * (src/main_f64.cu:131-132), optionally all matrix values to 1 (-ones), analyses, runs `reps` timed products and prints
 * the reference's "SpMV_X:" line and CSV record.  Unlike the reference it verifies y against the serial CSR loop
 * through order_rid (the reference's verify_new call is commented out, src/main_f64.cu:157).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dasp.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != DASP_OK) {                                                                    \
            fprintf(stderr, "%s -> %s: %s\n", #call, dasp_strerror(rc_), dasp_last_error());     \
            return 1;                                                                            \
        }                                                                                        \
    } while (0)

int main(int argc, char **argv)
{
    if (argc < 2) {
        printf("Run the code by './spmv_mtx matrix.mtx [reps] [-ones]'.\n");
        return 0;
    }
    const dasp_dtype dt = DASP_F64; /* this example drives the FP64 path; see dasp_spmv_all_f16 for FP16 */
    int reps = 1000, ones = 0;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "-ones")) ones = 1;
        else if (atoi(argv[i]) > 0) reps = atoi(argv[i]);
    }
    int m, n, sym, *rowptr, *colidx;
    int64_t nnz;
    void *valv;
    CHECK(dasp_read_mtx(argv[1], dt, &m, &n, &nnz, &sym, &rowptr, &colidx, &valv));
    double *val = (double *)valv;
    double *x = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *y = (double *)malloc(sizeof(double) * (size_t)(m > 0 ? m : 1));
    int *order = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    for (int j = 0; j < n; j++) x[j] = ones ? 1.0 : sin(0.37 * j) + 0.5;
    if (ones) for (int64_t k = 0; k < nnz; k++) val[k] = 1.0;
    printf("\n===%s===  %d x %d, %lld nnz%s\n\n", argv[1], m, n, (long long)nnz, sym ? " (symmetric, expanded)" : "");

    dasp_handle *h;
    CHECK(dasp_create(&h, dt, 0, m, n, nnz, rowptr, colidx, val, 0.75, 256));
    CHECK(dasp_spmv_host(h, x, y)); /* y in permuted order */
    CHECK(dasp_export(h, "order_rid", order, (int64_t)sizeof(int) * m, NULL));

    /* verify through order_rid against the serial CSR loop (what verify_new does against cuSPARSE) */
    double err2 = 0.0, ref2 = 0.0;
    for (int k = 0; k < m; k++) {
        const int i = order[k];
        double s = 0.0;
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++) s += val[p] * x[colidx[p]];
        err2 += (y[k] - s) * (y[k] - s);
        ref2 += s * s;
    }
    const double rel = ref2 > 0 ? sqrt(err2 / ref2) : sqrt(err2);
    printf("check: relative L2 error vs serial CSR = %.3e  %s\n", rel, rel <= 1e-12 ? "PASS" : "FAIL");

    /* the reference times kernels only (100 warm-up + 1000 launches on resident data, src/dasp_f64.h:1285-1320); from a
       host-only program the repeatable call is the host-buffer product, so this loop includes the PCIe copies of x and y.
       Kernel-only timing on device pointers is dasp_spmv_timed (see bench.py). */
    dasp_stats_t st;
    CHECK(dasp_stats(h, &st));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < reps; i++) CHECK(dasp_spmv_host(h, x, y));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double ms = ((t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6) / reps;
    printf("SpMV_X (host buffers, copies included):  %8.4lf ms, %8.4lf GFlop/s\n", ms, 2.0 * (double)nnz / (ms * 1e6));
    char rec[1024];
    int k = dasp_report(h, argv[1], ms, rec, sizeof(rec));
    if (k < 0) { fprintf(stderr, "dasp_report failed\n"); return 1; }
    printf("preprocessing on the GPU: %.3f ms; padding rate %.4f; long/medium/short rows %d/%d/%d\n", st.preprocess_ms, st.rate_fill0,
           st.row_long, st.row_block, st.short_row_1 + 2 * st.common_13 + st.short_row_34 + st.short_row_2);
    printf("record: %s\n", rec);
    dasp_destroy(h);
    dasp_free_host(rowptr); dasp_free_host(colidx); dasp_free_host(valv);
    free(x); free(y); free(order);
    return rel <= 1e-12 ? 0 : 2;
}
