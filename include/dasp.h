/*
 * include/dasp.h — C ABI of the B200-native DASP SpMV (libdasp_b200.so).
 *
 * This is the drop-in boundary for the one hot path of SuperScientificSoftwareLaboratory/DASP:
 *   CSR in -> long/medium/short row classification + block re-organisation -> repeatable y = A*x,
 *   FP64 and FP16.
 *
 * The reference exposes that path as ONE monolithic C++ function
 *     void spmv_all(char *filename, MAT_VAL_TYPE *csrValA, int *csrRowPtrA, int *csrColIdxA,
 *                   MAT_VAL_TYPE *X_val, MAT_VAL_TYPE *Y_val, int *order_rid,
 *                   int rowA, int colA, int nnzA, int NUM, double threshold, int block_longest);
 *     (reference: src/dasp_f64.h:486-487, src/dasp_f16.h:1015-1016; called from
 *      src/main_f64.cu:149, src/main_f16.cu:146)
 * which preprocesses on the host, uploads, times 1100 launches and downloads y in PERMUTED order
 * together with the permutation order_rid.  Here the same surface is split into analyse
 * (dasp_create) and execute (dasp_spmv) so the product is repeatable; dasp_spmv_all_f64/_f16 keep
 * the reference's one-shot signature on top of it.
 *
 * Conventions: plain pointers and sizes only; every entry returns 0 on success or a negative
 * dasp_status (the reference checks no CUDA status at all: SURVEY.md §5); a handle is used by one
 * host thread at a time and admits ONE in-flight product at a time (the merge scratch of split long
 * rows belongs to the handle: do not overlap two products of the same handle on different streams
 * or replay a captured graph concurrently with itself); distinct handles (one per GPU, or several
 * per GPU) are independent.  Every entry selects the handle's device itself and restores the caller's
 * current device before returning.  The library owns all
 * device memory of a handle, the caller owns x, y and the CSR arrays.  There is no CPU fallback:
 * without a CUDA device every entry fails with DASP_ERR_CUDA.
 */
#ifndef DASP_B200_H
#define DASP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dasp_handle dasp_handle;

/* value type: the reference selects it at compile time with -D f64 (src/common.h:21-25) */
typedef enum { DASP_F64 = 0, DASP_F16 = 1 } dasp_dtype;

typedef enum {
    DASP_OK = 0,
    DASP_ERR_INVALID = -1, /* bad argument (NULL, negative size, unknown name/dtype)      */
    DASP_ERR_CUDA = -2,    /* a CUDA call failed; dasp_last_error() has the CUDA string   */
    DASP_ERR_ALLOC = -3,   /* device or host allocation failed                            */
    DASP_ERR_RANGE = -4,   /* nnz or a padded size does not fit the 32-bit layout         */
    DASP_ERR_BUFFER = -5   /* dasp_export destination too small                           */
} dasp_status;

/* SpMV kernel variant per row category (north_star: keep the MMA tile variant only where it
 * wins on achieved HBM GB/s).  AUTO = the measured winner recorded in profiles/. */
typedef enum {
    DASP_VARIANT_AUTO = 0,
    DASP_VARIANT_CUDA_CORE = 1, /* per-lane 4-wide dot over the 8x4 tiles, 256/128-bit loads */
    DASP_VARIANT_MMA = 2,       /* tensor cores on the same tiles: FP64 mma.sync m8n8k4 (DMMA) as the reference does;
                                   FP16 mma.sync m16n8k16 f16 x f16 -> f32 (HMMA.16816.F32; the reference's m8n8k4.f16 is
                                   only emulated on sm_100a), medium and long rows */
    DASP_VARIANT_SPLIT = 3,     /* medium rows only: CUDA-core, four lanes per row (small, latency-bound matrices) */
    DASP_VARIANT_TMA = 4,       /* long rows only: CUDA-core fed by per-warp TMA bulk copies (cp.async.bulk + mbarrier ring) */
    DASP_VARIANT_BLOCKED = 5,   /* long rows only: column-blocked copy of the long part, x blocks staged in shared memory by
                                   TMA bulk copies, atomic merge of the split rows; AUTO picks it when the long rows'
                                   gathers are scattered (built on demand otherwise; needs row_long <= 65535) */
    DASP_VARIANT_BANDED = 6     /* medium rows: the 8 groups of a CTA (walked in order of original row id) gather from a
                                   64 KB window of x that a ninth warp stages in shared memory with TMA bulk copies; AUTO
                                   picks it when the gathers are scattered and mostly inside the window.
                                   short rows: the warp items of the four short segments walked by band of original
                                   rows, the band's window of x staged in shared memory by TMA bulk copies (double
                                   buffered), gathers served from there; AUTO picks it for large short parts whose entries
                                   lie inside their band's window (short_band_hit_rate) */
} dasp_variant;

/* The reference's locals that describe the layout (the 18 structure columns of its CSV record,
 * src/dasp_f64.h:1440-1441, plus the launch geometry of :1194-1214). All counts, not bytes. */
typedef struct dasp_stats_t {
    int dtype, m, n;
    int64_t nnz;
    int row_long, row_block, row_zero;
    int short_row_1, short_row_3, short_row_2, short_row_4; /* after the 1&3 pairing */
    int common_13, short_row_34;
    int rowloop, blocknum;
    int warp_number, BlockNum_long, fill0_nnz_long;
    int fill0_nnz_reg, nnz_irreg, origin_nnz_reg;
    int fill0_nnz_short, fill0_nnz_short13, fill0_nnz_short34, fill0_nnz_short22;
    int threadblock13, threadblock34, threadblock22;
    int nnz_short, nnz_long;
    int BlockNum, BlockNum_short_1, BlockNum_all, sumBlockNum;
    int fill0_nnz_irreg;
    double rate_fill0;       /* (padded slots - nnz) / nnz, src/dasp_f64.h:1159-1160           */
    int64_t data_X, data_X2; /* the reference's DASP byte accounting, src/dasp_f64.h:1162-1172 */
    int64_t data_origin1;    /* CSR ("algorithmic") bytes, src/main_f64.cu:143                 */
    double preprocess_ms;    /* device time of the GPU preprocessing inside dasp_create        */
    int64_t device_bytes;    /* device memory held by the handle                               */
    int col_min, col_max;    /* smallest / largest column index present (col_max = -1: no entries): the
                                only part of x a product reads; dasp_spmv_host uploads just that range */
    double long_gather_lines; /* diagnostic: estimated distinct 128-byte lines of x per 32-slot group of the long
                                part (1 = dense ascending columns, 32 = every gather its own line); above the
                                crossover AUTO uses the column-blocked long-row kernel (long_blocked != 0) */
    int long_blocked;         /* AUTO runs the long rows through the column-blocked kernel */
    int short_banded;         /* AUTO runs the short rows through the band kernel (x windows staged in shared memory) */
    double short_band_hit_rate; /* diagnostic: fraction of short-row entries whose column lies inside the window of
                                its row band (0 when the short part is too small for the band kernel) */
    double medium_gather_lines; /* diagnostic, filled when DASP_VARIANT_BANDED is selected for the medium rows: distinct
                                128-byte lines of x per 32-lane gather (1-3 for a stencil, ~32 for scattered columns) */
    double medium_band_hit_rate; /* same: fraction of medium-row entries inside the x window of their CTA */
    int medium_banded, reserved_; /* always 0: AUTO never picks the medium-band kernel (measured slower everywhere) */
} dasp_stats_t;

/* Analyse: run the DASP preprocessing on the GPU (replaces the host code src/dasp_f64.h:499-1157,
 * src/dasp_f16.h:1029-1443) and keep the packed layout resident.  rowptr/colidx/val may be host
 * or device pointers (detected); val is double[nnz] or IEEE-half[nnz] by dtype; columns need not
 * be sorted.  The CSR is validated (rowptr starts at 0, is non-decreasing and ends at nnz; columns in
 * [0, n)): DASP_ERR_INVALID otherwise.  Device-resident inputs must not be modified during the call
 * (the call waits for work already queued on the device before it reads them).  threshold/block_longest: the reference's run-time constants 0.75 / 256
 * (src/main_f64.cu:124-125).  The CSR is not retained. */
int dasp_create(dasp_handle **h, dasp_dtype dtype, int device, int m, int n, int64_t nnz,
                const int *rowptr, const int *colidx, const void *val, double threshold,
                int block_longest);

/* Execute y = A*x, asynchronously on `stream` (a cudaStream_t, NULL = default stream).
 * d_x: n values, d_y: m values, both device pointers of the handle's dtype.  y is produced in the
 * reference's PERMUTED order: y[k] belongs to original row order_rid[k] (K11 in SURVEY.md §8a;
 * kernels replaced: dasp_spmv2 + longPart_sum, src/dasp_f64.h:53-484, src/dasp_f16.h:106-590).
 * Rows without entries are written as 0 on every call. */
int dasp_spmv(dasp_handle *h, const void *d_x, void *d_y, void *stream);

/* Same product, y scattered to ORIGINAL row order (y[order_rid[k]]); for solvers that feed y back
 * as the next x (the power-iteration workload). */
int dasp_spmv_unpermuted(dasp_handle *h, const void *d_x, void *d_y, void *stream);

/* y = alpha*A*x + beta*y (y in permuted order when permuted != 0, original row order otherwise); the general
 * form solvers need (the reference computes alpha=1, beta=0 only; its cuSPARSE comparator passes alpha/beta,
 * src/main_f64.cu:23-24).  Rows without entries become beta*y. */
int dasp_spmv_axpby(dasp_handle *h, double alpha, const void *d_x, double beta, void *d_y, int permuted, void *stream);

/* FP16 matrix and x in, FP32 y out: the kernels accumulate an FP16 product in fp32 anyway; this entry
 * stores that sum unrounded (d_y: m floats; permuted != 0: the reference's permuted order, else original row order).  The
 * reference rounds to half (and accumulates in half, src/dasp_f16.h:121-126).  FP16 handles only. */
int dasp_spmv_f16_f32out(dasp_handle *h, const void *d_x, float *d_y, int permuted, void *stream);

/* Checkpoint of the preprocessed matrix (the layout is a pure function of the CSR): write every array and scalar of
 * the handle to one binary file / rebuild a handle from it on `device` without the CSR and without preprocessing. */
int dasp_save(const dasp_handle *h, const char *path);
int dasp_load(dasp_handle **h, const char *path, int device);

/* Host-buffer convenience with the reference's data movement (src/dasp_f64.h:1241,1402): upload x,
 * run, download y (permuted order), synchronous.  Only x[col_min .. col_max] (dasp_stats) is uploaded:
 * the product reads nothing else, and a row slab of a banded matrix touches a fraction of x. */
int dasp_spmv_host(dasp_handle *h, const void *x_host, void *y_host);

/* Host buffers, `count` independent products y_j = A*x_j (several right-hand sides, or a stream of requests):
 * product j uploads x_hosts[j], multiplies, downloads into y_hosts[j] (permuted order) exactly like dasp_spmv_host,
 * but the three phases run on three streams with double-buffered device staging, so the upload of product j+1 and
 * the download of product j-1 overlap the kernel of product j (both PCIe directions busy).  Host buffers should be
 * pinned (cudaHostAlloc / cudaHostRegister) for the copies to overlap.  Synchronous: returns when all y are in place. */
int dasp_spmv_host_batch(dasp_handle *h, const void *const *x_hosts, void *const *y_hosts, int count);

/* The reference's measurement loop (src/dasp_f64.h:1285-1320: warm-up launches, then `reps` launches
 * back to back) issued from C so that small matrices are not bound by the caller's launch rate, timed
 * with CUDA events on `stream`.  *total_ms receives the device time of the `reps` launches. */
int dasp_spmv_timed(dasp_handle *h, const void *d_x, void *d_y, void *stream, int warmup, int reps,
                    float *total_ms);

/* device pointer to order_rid[m] (permuted index -> original row), src/dasp_f64.h:960-976 */
int dasp_order(const dasp_handle *h, const int **d_order_rid);

int dasp_stats(const dasp_handle *h, dasp_stats_t *out);

/* The reference's reporting surface: the CSV record it appends to data/spmv_f64_record.csv / spmv_f16_record.csv
 * (src/dasp_f64.h:1440-1441, src/dasp_f16.h:1757-1758), same columns in the same order and printf formats:
 *   label,rowA,colA,nnzA,short_row_1,common_13,short_row_3,short_row_4,short_row_2,row_long,row_block,nnz_short,
 *   fill0_nnz_short,nnz_long,fill0_nnz_long,origin_nnz_reg,fill0_nnz_reg,nnz_irreg,rate_fill0,block_longest,data_X,
 *   [FP16: preprocessing ms,] time ms,GFlop/s,[FP16: time, GFlop/s again (the reference's "bypass" run),]GB/s(data_X),GB/s(data_X2),
 * for a measured time per SpMV of spmv_ms.  Nothing is written to disk.  Returns the record length (excluding
 * the NUL) or a negative status; DASP_ERR_BUFFER if cap is too small. */
int dasp_report(const dasp_handle *h, const char *label, double spmv_ms, char *out, int64_t cap);

/* Copy one preprocessing output to the host for bit-exact checks.  name is one of:
 * order_rid, long_rpt_new, long_val, long_cid, blockPtr, irreg_rpt, irreg_val, irreg_cid,
 * reg_val, reg_cid, short_val, short_cid.  *bytes receives the array size; host_dst may be NULL
 * to query it. */
int dasp_export(const dasp_handle *h, const char *name, void *host_dst, int64_t cap_bytes,
                int64_t *bytes);

/* medium: AUTO | CUDA_CORE | MMA | SPLIT | BANDED;  long_rows: AUTO | CUDA_CORE | MMA | TMA | BLOCKED;  short_rows: AUTO | CUDA_CORE | MMA | BANDED
 * (short-row MMA is FP64 only; values that do not apply to a category fall back to CUDA_CORE). */
int dasp_set_variant(dasp_handle *h, dasp_variant medium, dasp_variant long_rows, dasp_variant short_rows);

/* The medium-row kernels read a compact resident copy of the regular part's column indices (per 8x4 tile one
 * 32-bit base + 16-bit offsets, 10.1 instead of 12 bytes per FP64 slot) and fall back to reg_cid for blocks whose
 * tiles span >= 65535 columns.  reg_cid itself stays bit-exact and exportable.  on = 0 makes the kernels read
 * reg_cid everywhere (A/B measurements); default 1. */
int dasp_set_index_compression(dasp_handle *h, int on);

/* Profiling aid: restrict dasp_spmv to some row categories (bit 0 long, bit 1 medium, bit 2 short,
 * bit 3 empty rows; default 15 = all).  y entries of disabled categories are left untouched. */
int dasp_set_category_mask(dasp_handle *h, int mask);

/* number of kernel launches one dasp_spmv issues (for launch accounting) */
int dasp_launches_per_spmv(const dasp_handle *h);

int dasp_destroy(dasp_handle *h);

const char *dasp_strerror(int status);
const char *dasp_last_error(void); /* thread-local detail of the last failure */

/* One-shot calls with the reference's spmv_all argument list (host pointers, y permuted, order_rid
 * out).  `filename` is only a label, NUM is unused, exactly as in the reference. Returns a status
 * instead of void. */
int dasp_spmv_all_f64(const char *filename, const double *csrValA, const int *csrRowPtrA,
                      const int *csrColIdxA, const double *X_val, double *Y_val, int *order_rid,
                      int rowA, int colA, int nnzA, int NUM, double threshold, int block_longest);
int dasp_spmv_all_f16(const char *filename, const void *csrValA, const int *csrRowPtrA,
                      const int *csrColIdxA, const void *X_val, void *Y_val, int *order_rid,
                      int rowA, int colA, int nnzA, int NUM, double threshold, int block_longest);

/* Fused product + exchange for the row-partitioned iterated workload: this GPU's slab product, in ORIGINAL row order
 * and scaled by 1/sqrt(*d_norm2) when d_norm2 != NULL (device scalar, e.g. the all-reduced squared norm of the
 * previous iterate), is stored by the SpMV kernel itself at element offset row_offset of EVERY vector in d_dests
 * (1..8 device pointers): this GPU's copy of the next x and the peer GPUs' copies mapped into this process (CUDA IPC /
 * symmetric memory), or a single NVSwitch multicast mapping of all copies.  The stores travel over NVLink while the
 * rest of the slab is still being multiplied; no separate all-gather / broadcast is needed, only a barrier (the
 * all-reduce of the next norm) before the vectors are read.  FP64 and FP16. */
int dasp_spmv_scatter_to(dasp_handle *h, const void *d_x, void *const *d_dests, int n_dests, int64_t row_offset,
                         const double *d_norm2, void *stream);

/* The same fused product + exchange with the slab product left in PERMUTED order (no scatter: the stores of
 * neighbouring rows are contiguous): y_perm[k] / sqrt(*d_norm2) is stored at element row_offset + k of every vector in
 * d_dests.  Meant for the relabelled mode below, where the permuted product IS the next x. */
int dasp_spmv_permuted_to(dasp_handle *h, const void *d_x, void *const *d_dests, int n_dests, int64_t row_offset,
                          const double *d_norm2, void *stream);

/* Relabelled (P*A*P^T) mode for solvers that feed y back as the next x: the kernels' column indices
 * are replaced by d_new_index[column] (device array of n ints with values in [0, n_new)), so x is expected in the
 * relabelled index space (length n_new).  With d_new_index = the inverse permutation (dasp_inverse_order) of a square
 * matrix, x is expected in PERMUTED order — exactly the order dasp_spmv produces y in — and the iteration needs neither
 * dasp_spmv_unpermuted's scattered stores nor an un-permute pass; with row slabs on several GPUs, d_new_index[j] =
 * slab offset of the owner of row j + that slab's inverse permutation of j.  The reference arrays (dasp_export) are
 * untouched; compact indices and the column-blocked copy of the long part are rebuilt.  Synchronous; no product of the
 * handle may be in flight.  May be called again with another map (always relative to the ORIGINAL column indices). */
int dasp_relabel_columns(dasp_handle *h, const int *d_new_index, int n_new);
/* device pointer to the inverse of order_rid: original row -> permuted index, int[m] */
int dasp_inverse_order(const dasp_handle *h, const int **d_inv_order);

/* The exchange step of the iterated workload as ONE coalesced pass: y (this GPU's slab product in PERMUTED order, as
 * dasp_spmv leaves it) is read through the inverse permutation, scaled by 1/sqrt(*d_norm2) when d_norm2 != NULL and
 * stored in ORIGINAL row order, fully coalesced, at element offset row_offset of every vector in d_dests (local, peer
 * mapped, or one NVSwitch multicast mapping: then every 256-byte warp store is replicated to all GPUs by the switch).
 * Replaces un-permute + scale + all-gather/broadcast. */
int dasp_unpermute_to(dasp_handle *h, const void *d_y_perm, void *const *d_dests, int n_dests, int64_t row_offset,
                      const double *d_norm2, void *stream);

/* Vector helpers of the iterated (power-iteration) workload, x <- A x / ||A x||_2 (north_star; the
 * reference has no iterated driver).  All pointers are device pointers, FP64 only.
 *   dasp_sumsq:  *d_out = sum_i v[i]^2           (deterministic two-stage reduction)
 *   dasp_scale_rsqrt: v[i] *= 1/sqrt(*d_norm2)   (d_norm2 stays on the device: no host round trip) */
int dasp_sumsq(const double *d_v, int64_t count, double *d_out, void *stream);
int dasp_scale_rsqrt(double *d_v, int64_t count, const double *d_norm2, void *stream);
/*   dasp_scale_copy_to: dest_p[offset + i] = v[i] / sqrt(*d_norm2) for every destination p (1..8 device pointers:
 *   local, peer-mapped, or one NVSwitch multicast mapping) — scale + broadcast of a slab as one coalesced pass */
int dasp_scale_copy_to(const double *d_v, int64_t count, void *const *d_dests, int n_dests, int64_t offset,
                       const double *d_norm2, void *stream);

/* Matrix Market coordinate file -> host CSR, with the reference reader's exact semantics
 * (mmio_allinone, src/mmio_highlevel.h:608-774) so that a file produces the identical CSR and therefore the
 * identical DASP layout: 1-based -> 0-based; entries of a row kept in FILE order (columns not sorted, duplicates
 * kept); `symmetric` and `hermitian` mirrored (off-diagonal entries only, the mirror is appended to its row at the
 * moment the original is read); `skew-symmetric` NOT expanded; `pattern` -> 1.0; `integer` read as int; `complex`
 * -> real part.  val is double[nnz] or IEEE-half[nnz] (double rounded to nearest) by dtype.  The three arrays are
 * malloc'ed; release them with dasp_free_host.  Returns DASP_OK, DASP_ERR_INVALID (cannot open / not a coordinate
 * Matrix Market file / malformed entry) or DASP_ERR_RANGE (expanded nnz above 2^31-1). */
int dasp_read_mtx(const char *filename, dasp_dtype dtype, int *m, int *n, int64_t *nnz, int *is_symmetric,
                  int **rowptr, int **colidx, void **val);
void dasp_free_host(void *p);

/* nnz-balanced contiguous row partition for multi-GPU runs: cut[p] = smallest i
 * with rowptr[i] >= p*nnz/parts; cuts has parts+1 entries, cut[0]=0, cut[parts]=m. rowptr: host. */
int dasp_partition_rows(int m, const int *rowptr, int parts, int *cuts);

#ifdef __cplusplus
}
#endif
#endif /* DASP_B200_H */
