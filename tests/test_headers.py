"""CPU suite: the public headers are plain C (the boundary must be bindable from cgo/JNI/ctypes-style FFI)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile_c(src: str) -> None:
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "t.c")
        with open(path, "w") as f:
            f.write(src)
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                               "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), path])


def test_dasp_h_is_c99():
    _compile_c('#include "dasp.h"\n#include "dasp_synth.h"\n'
               "int use(dasp_handle *h, const void *x, void *y) { dasp_stats_t s; (void)dasp_stats(h, &s);"
               " return dasp_spmv(h, x, y, 0) + (int)sizeof(dasp_synth_spec); }\n")


def test_reference_shim_of_integration_md_compiles():
    """The spmv_all shim shown in INTEGRATION.md, compiled as C against include/dasp.h for both precisions."""
    shim = '''
#include "dasp.h"
#include <stdio.h>
#include <stdlib.h>
#define MAT_PTR_TYPE int
static void spmv_all(char *filename, MAT_VAL_TYPE *csrValA, MAT_PTR_TYPE *csrRowPtrA, int *csrColIdxA,
                     MAT_VAL_TYPE *X_val, MAT_VAL_TYPE *Y_val, int *order_rid,
                     int rowA, int colA, MAT_PTR_TYPE nnzA, int NUM, double threshold, int block_longest)
{
    int rc = SPMV_ALL(filename, csrValA, csrRowPtrA, csrColIdxA, X_val, Y_val, order_rid,
                      rowA, colA, nnzA, NUM, threshold, block_longest);
    if (rc != DASP_OK) { fprintf(stderr, "dasp: %s: %s\\n", dasp_strerror(rc), dasp_last_error()); exit(1); }
}
int main(void) { (void)spmv_all; return 0; }
'''
    _compile_c("#define MAT_VAL_TYPE double\n#define SPMV_ALL dasp_spmv_all_f64\n" + shim)
    _compile_c("#define MAT_VAL_TYPE unsigned short\n#define SPMV_ALL dasp_spmv_all_f16\n" + shim)


def test_oracle_header_is_c99():
    _compile_c('#include "dasp_oracle.h"\nint f(void) { return (int)sizeof(dasp_oracle_layout); }\n')
