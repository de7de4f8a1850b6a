// dasp_internal.h — handle layout and helpers shared by the translation units of libdasp_b200.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dasp.h"

namespace dasp {

void set_error(const char *fmt, ...);

#define DASP_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            ::dasp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorMemoryAllocation) ? DASP_ERR_ALLOC : DASP_ERR_CUDA;           \
        }                                                                                        \
    } while (0)

#define DASP_TRY(expr)                   \
    do {                                 \
        int s_ = (expr);                 \
        if (s_ != DASP_OK) return s_;    \
    } while (0)

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards (one handle per GPU
// in one process: every entry that touches a handle's memory or launches on it goes through this).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        status = cudaGetDevice(&prev);
        if (status == cudaSuccess && prev != device) {
            status = cudaSetDevice(device);
            switched = status == cudaSuccess;
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define DASP_ON_DEVICE(dev)                                                                       \
    ::dasp::DeviceGuard device_guard_(dev);                                                       \
    if (device_guard_.status != cudaSuccess) {                                                    \
        ::dasp::set_error("cannot select device %d: %s", (dev), cudaGetErrorString(device_guard_.status)); \
        cudaGetLastError();                                                                       \
        return DASP_ERR_CUDA;                                                                     \
    }

// Tracks every device allocation of a handle so destroy/fail paths free them all.
struct DevicePool {
    std::vector<void *> ptrs;  // individually cudaMalloc'ed arrays
    std::vector<void *> slabs; // cudaMalloc'ed slabs (see reserve)
    int64_t bytes = 0;        // bytes handed out
    // One cudaMalloc that the following alloc() calls carve from (256-byte aligned pieces) while they fit: the ~40 arrays
    // of one analysis cost three driver allocations instead of one each (cudaMalloc / cudaFree of large blocks are the
    // bulk of dasp_create's time on the GPU).  Best effort: what does not fit is allocated on its own.
    char *slab = nullptr;
    size_t slab_size = 0, slab_used = 0;
    int reserve(size_t n);
    void close_slab() { slab = nullptr; slab_size = slab_used = 0; } // later alloc() calls allocate on their own (releasable)
    int alloc(void **p, size_t n);
    void release(void *p); // frees an individually allocated array now; pieces carved from a slab stay until free_all()
    void free_all();
};

// Work list entry of the long-row kernel: one warp reduces elements [begin, end) of long row `row`.
struct LongUnit {
    int row;
    int slot; // index into the partial-sum scratch (== unit id)
    int begin, end;
};

constexpr int SMQ_MAX = 512; // upper bound of the SM count the SM-affine queues are sized for

struct Layout {
    dasp_stats_t s{};
    size_t esz = 8;
    // ---- the reference's arrays (bit-exact) ----
    int *order_rid = nullptr;
    int *long_rpt_new = nullptr;
    void *long_val = nullptr;
    int *long_cid = nullptr;
    int *blockPtr = nullptr;
    int *irreg_rpt = nullptr;
    void *irreg_val = nullptr;
    int *irreg_cid = nullptr;
    void *reg_val = nullptr;
    int *reg_cid = nullptr;
    void *short_val = nullptr;
    int *short_cid = nullptr;
    // ---- column indices the kernels actually read: the reference arrays above, or — after dasp_relabel_columns — relabelled
    // copies (the reference arrays stay bit-exact and exportable) ----
    int *k_long_cid = nullptr, *k_reg_cid = nullptr, *k_irreg_cid = nullptr, *k_short_cid = nullptr;
    int relabelled = 0;
    int x_len = 0; // length of the x the kernels index (n, or the size of the relabelled index space)
    // ---- derived launch data of this implementation (not part of the reference layout) ----
    int n_long_units = 0;
    int long_unit_warps = 32;       // reference warps per long-row work unit: LONG_UNIT_WARPS, fewer when the long part is small (derive)
    int *long_unit_row = nullptr;   // [n_long_units] long row of each unit (execution order)
    int *long_unit_chunk = nullptr; // [n_long_units] which chunk of that row
    int *long_unit_first = nullptr; // [row_long+1]  first unit of each long row
    void *long_partial = nullptr;   // [n_long_units] double (f64) / float (f16) partial sums
    unsigned *long_done = nullptr;  // [row_long] arrival counters (self-resetting)
    // compact column indices of the regular part (resident form read by the kernels; reg_cid is kept for export and
    // for blocks that cannot be compressed): per 8x4 tile one 32-bit base + 32 16-bit offsets, 0xFFFF = column 0
    int *reg_cbase = nullptr;              // [fill0_nnz_reg / 32]
    unsigned short *reg_cdelta = nullptr;  // [fill0_nnz_reg]
    unsigned char *blk_wide = nullptr;     // [blocknum] 1: some tile of the block spans >= 65535 columns -> use reg_cid
    unsigned short *blk_live = nullptr;    // [blocknum] tiles of the block up to the last one holding a non-zero value
    int reg_compact_done = 0;              // the four arrays above were written by pack_reg (dasp_create); derive() skips compress_cid once
    int *long_cbase = nullptr;              // [fill0_nnz_long / 32] same compact form for the long part
    unsigned short *long_cdelta = nullptr;  // [fill0_nnz_long]
    unsigned char *long_wide = nullptr;     // [n_long_units] in execution order
    int *inv_order = nullptr;               // [m] inverse of order_rid (original row -> permuted index)
    unsigned char *med_has_irreg = nullptr; // [ceil(row_block/32)] 1 if any row of the 32-row group has an irregular tail
    // Locality-ordered work lists (large matrices): the fused kernel walks the medium 32-row groups, and the CTAs of the four
    // short segments interleaved, in order of the ORIGINAL id of their first row, so that at any time the resident CTAs
    // gather from one sliding window of x (every length class / short segment sweeps all of x on its own otherwise).
    int *med_order = nullptr;               // [blocknum / 4] group processed by the w-th medium warp
    int *smq_cnt = nullptr;                 // [SMQ_MAX + 1] per-SM chunk counters + arrivals of the SM-affine medium-row queues
                                            // (small matrices; self-resetting: the last CTA to take a chunk zeroes them)
    int *short_map = nullptr;               // [short_map_n] category << 28 | CTA index inside the category
    int short_map_n = 0;
    int short_ctas[4] = {0, 0, 0, 0};       // CTAs of singles / 1&3 / 3&4 / 2&2 the map was built for
    // Medium-band kernel: first column of the x window of every CTA (8 consecutive groups of the processing order), the
    // fraction of the medium entries inside their CTA's window, and the average number of distinct 128-byte lines one
    // 32-lane gather of the lane-per-row mapping touches (1-3 for a stencil, ~32 for scattered columns)
    int *mb_lo = nullptr;
    int mb_auto = 0;
    double mb_hit_rate = 0.0, med_gather_lines = 0.0;
    // Short-band kernel: warp items of the four short segments sorted by the band of original rows they start in, and the
    // x window of every band
    int *sb_item = nullptr;                 // [sb_nitems] segment (2..5) << 28 | warp item inside the segment
    int *sb_band_ptr = nullptr;             // [sb_nbands + 1]
    int *sb_lo = nullptr;                   // [sb_nbands] first column of the band's window (multiple of 8)
    int sb_nbands = 0, sb_nitems = 0, sb_auto = 0;
    double sb_hit_rate = 0.0;               // fraction of short-row entries whose column lies inside its band's window
    double long_lines_avg = 0.0;            // estimated distinct 128-byte lines of x per 32-slot group of the long part
    // Column-blocked copy of the long part ("LCB", built when the long rows gather x all over the place): the live
    // entries of all long rows sorted by (column block, long row); a CTA stages one block of x in shared memory with a
    // TMA bulk copy and gathers from there.  12 (f64) / 6 (f16) bytes per entry like the CSR itself.
    int lcb_bw_log2 = 0;                    // block width = 1 << lcb_bw_log2 columns
    int lcb_nblk = 0;                       // column blocks
    int lcb_live = 0;                       // entries (reference padding dropped, every block padded to a multiple of 4)
    int lcb_nctas = 0;                      // CTAs of the LCB kernel (LCB_PART entries each, never across blocks)
    void *lcb_val = nullptr;                // [lcb_live]
    unsigned *lcb_idx = nullptr;            // [lcb_live] long-row index << 16 | column - block * width (row_long <= 65535)
    // FP64: the same indices in 16 bits (what the kernel streams): 13 bits of column inside the block + 3 bits of row DELTA to
    // the previous entry; every chunk of 1024 entries (what one warp walks in a row) restarts from lcb_chunk_row; a chunk
    // with a delta > 7 is flagged in lcb_chunk_wide and read through lcb_idx
    unsigned short *lcb_idx16 = nullptr;    // [lcb_live]
    int *lcb_chunk_row = nullptr;           // [lcb_live / 1024]
    unsigned char *lcb_chunk_wide = nullptr; // [lcb_live / 1024]
    int *lcb_blk_ptr = nullptr;             // [lcb_nblk + 1] first entry of each block (multiples of 4)
    int *lcb_cta_first = nullptr;           // [lcb_nblk + 1] first CTA of each block
    void *lcb_acc = nullptr;                // [LCB_COPIES][lcb_acc_stride] accumulators (double / float), zero between launches
    unsigned *lcb_done = nullptr;           // [1] CTAs finished in the current launch (self-resetting)
};

} // namespace dasp

struct dasp_handle {
    int device = 0;
    int mb_attr_set = 0;
    int smq_attr_set = 0;
    int sb_attr_set = 0;
    int lcb_attr_set = 0; // lcb_kernel dynamic shared memory attribute set on this device
    int lcb_auto = 0; // AUTO uses the column-blocked long-row kernel (decided in derive() from long_lines_avg)
    dasp_dtype dtype = DASP_F64;
    double threshold = 0.75;
    int block_longest = 256;
    dasp::Layout L;
    dasp::DevicePool pool;
    int category_mask = 15;
    int index_compression = 1;
    int sm_count = 0;
    const void *carved_narrow = nullptr; // the same for the 128-thread small-matrix kernels
    const void *carved_kernel = nullptr; // kernel whose L1 carve-out preference was already set on this device
    dasp_variant var_medium = DASP_VARIANT_AUTO, var_long = DASP_VARIANT_AUTO, var_short = DASP_VARIANT_AUTO;
    // device staging of x / y for dasp_spmv_host (owned by pool)
    void *dx_stage = nullptr, *dy_stage = nullptr;
    cudaStream_t own_stream = nullptr;
    // column-blocked long rows run beside the fused kernel (launch_spmv): a side stream forked from / joined to the caller's
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // dasp_spmv_host_batch: upload / compute / download streams, double-buffered staging and hand-over events
    int batch_ready = 0;
    cudaStream_t batch_stream[3] = {nullptr, nullptr, nullptr};
    void *batch_dx[2] = {nullptr, nullptr}, *batch_dy[2] = {nullptr, nullptr};
    cudaEvent_t batch_ev[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
};

namespace dasp {
constexpr int SPMV_CTA = 256;           // threads per CTA of the fused kernel (bandwidth-bound form)
constexpr int SINGLES_PER_THREAD = 4;   // single-entry rows per thread
constexpr int SHORT_TILES_PER_WARP = 4; // 8x4 tiles of a short segment per warp
constexpr int LONG_UNIT_WARPS = 32; // one long-row work unit = at most 32 reference "warps" of 64 (f16: 256) slots (Layout::long_unit_warps)
#ifndef DASP_LCB_PART
#define DASP_LCB_PART 32768
#endif
#ifndef DASP_LCB_COPIES
#define DASP_LCB_COPIES 8
#endif
constexpr int LCB_PART = DASP_LCB_PART;     // most entries one CTA of the column-blocked long-row kernel takes; a block is cut into equal parts
constexpr int LCB_COPIES = DASP_LCB_COPIES;       // private copies of the long-row accumulators (CTA c adds into copy c % 8): with one copy
                                    // every atomic of the GPU lands on row_long * 8 bytes, a handful of L2 lines
constexpr int LCB_BYTES = 65536;    // shared-memory bytes of one staged block of x
constexpr int MB_WINDOW_BYTES = 65536; // bytes of x one CTA of the medium-band kernel stages (must match spmv.cu MB_WIN_BYTES)
constexpr int SB_BAND_ROWS = 16384; // original rows per band of the short-band kernel
constexpr int SB_WINDOW_BYTES = 196608; // bytes of x one band stages in shared memory: 24576 doubles = the band + 4096 columns either side

// preprocess.cu
int scan_inplace(DevicePool &tmp_pool, int *d, int count, cudaStream_t st);
int radix_sort_pairs(DevicePool &tmp, const int *keys_in, const int *vals_in, int *keys_out, int *vals_out, int n, int bits,
                     bool descending, cudaStream_t st);
// derive.cu: kernel-facing data derived from the reference layout (also after dasp_load); build_lcb on demand
int derive(dasp_handle *h, cudaStream_t st);
int build_lcb(dasp_handle *h, cudaStream_t st);
int build_short_bands(dasp_handle *h, cudaStream_t st, bool force);
int build_medium_bands(dasp_handle *h, cudaStream_t st);
// kernel-facing column indices := new_index[reference column]; compact indices and the column-blocked copy are rebuilt
int relabel_columns(dasp_handle *h, const int *d_new_index, int n_new, cudaStream_t st);
// range / monotonicity check of the offset and index arrays of a layout read from a file (dasp_load)
int validate_layout(dasp_handle *h, cudaStream_t st);
int preprocess(dasp_handle *h, int m, int n, int64_t nnz, const int *d_rowptr, const int *d_colidx,
               const void *d_val, cudaStream_t st);
// spmv.cu
// extra destinations of a fused-exchange product (dasp_spmv_scatter_to)
struct ScatterTo {
    void *extra[7];
    int n_extra;
    int64_t row_offset;
    const double *norm2;
};
int launch_spmv(dasp_handle *h, const void *d_x, void *d_y, const int *scatter, cudaStream_t st,
                const double *alpha_beta = nullptr, const ScatterTo *multi = nullptr, bool out_f32 = false);
int save_layout(const dasp_handle *h, const char *path);
int load_layout(dasp_handle *h, const char *path);
int launches_per_spmv(const dasp_handle *h);
int unpermute_to(dasp_handle *h, const void *d_y_perm, const ScatterTo &dst, void *first, cudaStream_t st);
int sumsq(const double *d_v, int64_t count, double *d_out, cudaStream_t st);
int scale_by_rsqrt(double *d_v, int64_t count, const double *d_norm2, cudaStream_t st);
int scale_copy_to(const double *d_v, int64_t count, void *const *dests, int n_dests, int64_t offset, const double *d_norm2,
                  cudaStream_t st);
} // namespace dasp
