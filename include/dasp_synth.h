/*
 * include/dasp_synth.h — synthetic CSR generators for the benchmark configurations of
 * BASELINE.json (libdasp_synth.so).  Bench/test tooling, not part of the SpMV product: the
 * reference reads Matrix Market files (src/mmio_highlevel.h:608) and ships no generator; a
 * 450 M-nnz text file is not a viable input, so the named shapes are generated directly in device
 * memory.  Every generator is a pure function of (parameters, seed, row): any rank can produce any
 * row slab [row0, row1) of the same global matrix, with GLOBAL column indices.
 *
 * Usage: call *_rowlen to fill len[row1-row0] (device), exclusive-scan it into rowptr (caller),
 * then call *_fill with that rowptr to write colidx/val (device).  Values are U(-1,1) from a
 * counter-based hash of (seed, row, k).  All pointers are device pointers; `stream` is a
 * cudaStream_t.  Returns 0 or a negative code (CUDA error string via dasp_synth_last_error).
 */
#ifndef DASP_SYNTH_H
#define DASP_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dasp_synth_spec {
    int kind;        /* 0 stencil27, 1 powerlaw, 2 skewed, 3 banded-symmetric (cop20k_A stand-in),
                        4 powerlaw_spec, 5 skewed_spec: the generators of SURVEY.md 8(d) taken literally — columns in
                        RANDOM order, distinct within a row (keyed bijections, evaluated per element):
                        4: same row lengths as kind 1; 9 of 10 elements uniform in the +-window window around i*n/m
                           (window enlarged to a power of two >= 2x the windowed count for rows that do not fit), every
                           10th element a uniform global column outside the window;
                        5: n_long rows of long_len entries at seeded positions (one per stride of m/n_long rows), columns
                           uniform over all n without replacement; the other rows have L uniform in {1,2,3,4} with distinct
                           random columns of the +-window window */
    int64_t m, n;    /* global rows / columns */
    uint64_t seed;
    /* stencil27: grid nx*ny*nz, natural ordering (x fastest), columns ascending */
    int nx, ny, nz;
    /* powerlaw: L = min(floor(u^(-1/alpha)), lmax), u ~ U(0,1]; 90 % of the entries of a row are distinct
       ASCENDING points of a window around i*n/m (half-width max(window, L), i.e. at most 50 % dense),
       10 % are uniform global columns left at their CSR position, so rows are not sorted */
    double alpha;
    int lmax, window;
    /* skewed: rows [0, n_long) have long_len entries, distinct and ascending, one per stripe of
       band/long_len columns of one common band (width band >= 2*long_len, starting at band_lo); the
       other rows have L uniform in {1,2,3,4} with ascending columns in a +-window window */
    int n_long, long_len;
    int64_t band_lo, band;
    /* banded-symmetric: structurally symmetric pattern, row lengths ~ mean_len */
    int mean_len;
} dasp_synth_spec;

int dasp_synth_rowlen(const dasp_synth_spec *spec, int64_t row0, int64_t row1, int *d_len, void *stream);
int dasp_synth_fill(const dasp_synth_spec *spec, int64_t row0, int64_t row1, const int *d_rowptr, int *d_colidx,
                    double *d_val, void *stream);
/* double -> IEEE half (round to nearest even), for the FP16 configurations */
int dasp_synth_to_half(const double *d_src, void *d_dst, int64_t count, void *stream);
/* writes more than the L2 capacity so the next kernel starts cold */
int dasp_synth_flush_l2(void *d_scratch, int64_t bytes, void *stream);
const char *dasp_synth_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
