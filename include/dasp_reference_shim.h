/*
 * include/dasp_reference_shim.h — the reference's spmv_all, implemented on libdasp_b200.so.
 *
 * Drop-in for src/dasp_f64.h / src/dasp_f16.h of SuperScientificSoftwareLaboratory/DASP at the reference's own call
 * sites (src/main_f64.cu:149, src/main_f16.cu:146): same name, same argument list (src/dasp_f64.h:486-487,
 * src/dasp_f16.h:1015-1016), host pointers in, Y_val in PERMUTED order and order_rid out, filename only a label, NUM unused.
 * It must be included where the reference includes its dasp_f64.h / dasp_f16.h, after (or instead of) the reference's
 * common.h and utils.h, which define MAT_VAL_TYPE / MAT_PTR_TYPE and the helpers the reference's main() uses.
 * Differences, deliberate: a failing call prints the library's error and exits (the reference ignores every CUDA status);
 * nothing is appended to the data/ CSV records by this function; the 100 + 1000 timing launches are not part of the call
 * (dasp_spmv_timed); FP16 accumulates in fp32.
 */
#ifndef DASP_REFERENCE_SHIM_H
#define DASP_REFERENCE_SHIM_H

#include <stdio.h>
#include <stdlib.h>

#include "dasp.h"

#ifndef MAT_VAL_TYPE
#error "include the reference's common.h (MAT_VAL_TYPE, MAT_PTR_TYPE) before dasp_reference_shim.h"
#endif

static inline void spmv_all(char *filename, MAT_VAL_TYPE *csrValA, MAT_PTR_TYPE *csrRowPtrA, int *csrColIdxA,
                            MAT_VAL_TYPE *X_val, MAT_VAL_TYPE *Y_val, int *order_rid, int rowA, int colA, MAT_PTR_TYPE nnzA,
                            int NUM, double threshold, int block_longest)
{
#ifdef f64
    int rc = dasp_spmv_all_f64(filename, csrValA, csrRowPtrA, csrColIdxA, X_val, Y_val, order_rid, rowA, colA, nnzA, NUM,
                               threshold, block_longest);
#else
    int rc = dasp_spmv_all_f16(filename, csrValA, csrRowPtrA, csrColIdxA, X_val, Y_val, order_rid, rowA, colA, nnzA, NUM,
                               threshold, block_longest);
#endif
    if (rc != DASP_OK) {
        fprintf(stderr, "dasp: %s: %s\n", dasp_strerror(rc), dasp_last_error());
        exit(1);
    }
}

#endif /* DASP_REFERENCE_SHIM_H */
