// capi.cu — the C ABI of include/dasp.h on top of preprocess.cu / spmv.cu.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "dasp_internal.h"

namespace dasp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DevicePool::reserve(size_t n)
{
    if (n < (1u << 20)) return DASP_OK; // small analyses: individual allocations are cheap
    if (getenv("DASP_NO_SLAB")) return DASP_OK; // A/B aids
    if (const char *e = getenv("DASP_SLAB_MAX_MB")) { if (n > ((size_t)atol(e) << 20)) return DASP_OK; }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) { cudaGetLastError(); return DASP_OK; } // best effort: alloc() falls back to its own cudaMalloc
    slabs.push_back(p);
    slab = (char *)p; slab_size = n; slab_used = 0;
    return DASP_OK;
}

int DevicePool::alloc(void **p, size_t n)
{
    *p = nullptr;
    size_t want = n ? n : 256; // zero-length arrays still get a valid, aligned pointer
    const size_t carve = (want + 255) & ~(size_t)255;
    if (slab && slab_used + carve <= slab_size) {
        *p = slab + slab_used;
        slab_used += carve;
        bytes += (int64_t)want;
        return DASP_OK;
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) -> %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? DASP_ERR_ALLOC : DASP_ERR_CUDA;
    }
    ptrs.push_back(*p);
    bytes += (int64_t)want;
    return DASP_OK;
}

void DevicePool::release(void *p)
{
    if (!p) return;
    auto it = std::find(ptrs.begin(), ptrs.end(), p);
    if (it != ptrs.end()) { cudaFree(p); ptrs.erase(it); } // pieces carved from a slab stay until free_all()
}

void DevicePool::free_all()
{
    for (void *p : ptrs) cudaFree(p);
    for (void *p : slabs) cudaFree(p);
    ptrs.clear();
    slabs.clear();
    bytes = 0;
    slab = nullptr; slab_size = slab_used = 0;
}

static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// bring one CSR array to the device if it is a host pointer
static int stage(DevicePool &tmp, const void *src, size_t bytes, const void **dst, cudaStream_t st)
{
    if (bytes == 0 || is_device_ptr(src)) { *dst = src; return DASP_OK; }
    void *d = nullptr;
    DASP_TRY(tmp.alloc(&d, bytes));
    DASP_CUDA(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st));
    *dst = d;
    return DASP_OK;
}

} // namespace dasp

using namespace dasp;

extern "C" {

const char *dasp_strerror(int status)
{
    switch (status) {
    case DASP_OK: return "ok";
    case DASP_ERR_INVALID: return "invalid argument";
    case DASP_ERR_CUDA: return "CUDA error";
    case DASP_ERR_ALLOC: return "allocation failed";
    case DASP_ERR_RANGE: return "size exceeds the 32-bit layout";
    case DASP_ERR_BUFFER: return "destination buffer too small";
    default: return "unknown status";
    }
}

const char *dasp_last_error(void) { return g_err; }

int dasp_create(dasp_handle **out, dasp_dtype dtype, int device, int m, int n, int64_t nnz, const int *rowptr,
                const int *colidx, const void *val, double threshold, int block_longest)
{
    if (!out) { set_error("handle pointer is NULL"); return DASP_ERR_INVALID; }
    *out = nullptr;
    if (m < 0 || n < 0 || nnz < 0 || !rowptr || (nnz > 0 && (!colidx || !val)) || (dtype != DASP_F64 && dtype != DASP_F16) ||
        !(threshold > 0.0) || block_longest < 1) {
        set_error("dasp_create: bad argument (m=%d n=%d nnz=%lld threshold=%g block_longest=%d)", m, n, (long long)nnz,
                  threshold, block_longest);
        return DASP_ERR_INVALID;
    }
    if (nnz > INT32_MAX) { set_error("nnz=%lld does not fit the reference's 32-bit row pointers", (long long)nnz); return DASP_ERR_RANGE; }
    DASP_ON_DEVICE(device); // the caller's current device is restored on return
    // Device-resident CSR arrays may still be being written by work the caller queued on other streams; the
    // preprocessing runs on the handle's private non-blocking stream, which is not ordered against them.  dasp_create
    // is a one-time, synchronous analyse step: wait for the device once instead of asking for a stream.
    if (is_device_ptr(rowptr) || (nnz > 0 && (is_device_ptr(colidx) || is_device_ptr(val)))) DASP_CUDA(cudaDeviceSynchronize());
    dasp_handle *h = new (std::nothrow) dasp_handle();
    if (!h) { set_error("out of host memory"); return DASP_ERR_ALLOC; }
    h->device = device; h->dtype = dtype; h->threshold = threshold; h->block_longest = block_longest;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);

    int rc = DASP_OK;
    DevicePool staging;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        if ((rc = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) == cudaSuccess ? DASP_OK : DASP_ERR_CUDA)) {
            set_error("cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaStream_t st = h->own_stream;
        const size_t esz = dtype == DASP_F16 ? 2 : 8;
        const void *d_rowptr = nullptr, *d_colidx = nullptr, *d_val = nullptr;
        if ((rc = stage(staging, rowptr, sizeof(int) * ((size_t)m + 1), &d_rowptr, st))) break;
        if ((rc = stage(staging, colidx, sizeof(int) * (size_t)nnz, &d_colidx, st))) break;
        if ((rc = stage(staging, val, esz * (size_t)nnz, &d_val, st))) break;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        if ((rc = preprocess(h, m, n, nnz, (const int *)d_rowptr, (const int *)d_colidx, d_val, st))) break;
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { set_error("preprocessing failed: %s", cudaGetErrorString(cudaGetLastError())); rc = DASP_ERR_CUDA; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        h->L.s.preprocess_ms = ms;
        h->L.s.device_bytes = h->pool.bytes;
    } while (0);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    staging.free_all();
    if (rc != DASP_OK) { dasp_destroy(h); return rc; }
    *out = h;
    return DASP_OK;
}

int dasp_destroy(dasp_handle *h)
{
    if (!h) return DASP_OK;
    DeviceGuard guard(h->device);
    h->pool.free_all();
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int k = 0; k < 3; k++) {
        if (h->batch_stream[k]) cudaStreamDestroy(h->batch_stream[k]);
        for (int b = 0; b < 2; b++)
            if (h->batch_ev[k][b]) cudaEventDestroy(h->batch_ev[k][b]);
    }
    delete h;
    return DASP_OK;
}

int dasp_spmv(dasp_handle *h, const void *d_x, void *d_y, void *stream)
{
    if (!h || (!d_x && h->L.s.n > 0) || (!d_y && h->L.s.m > 0)) { set_error("dasp_spmv: NULL argument"); return DASP_ERR_INVALID; }
    DASP_ON_DEVICE(h->device);
    return launch_spmv(h, d_x, d_y, nullptr, (cudaStream_t)stream);
}

int dasp_spmv_unpermuted(dasp_handle *h, const void *d_x, void *d_y, void *stream)
{
    if (!h || (!d_x && h->L.s.n > 0) || (!d_y && h->L.s.m > 0)) { set_error("dasp_spmv_unpermuted: NULL argument"); return DASP_ERR_INVALID; }
    DASP_ON_DEVICE(h->device);
    return launch_spmv(h, d_x, d_y, h->L.order_rid, (cudaStream_t)stream);
}

int dasp_spmv_f16_f32out(dasp_handle *h, const void *d_x, float *d_y, int permuted, void *stream)
{
    if (!h || (!d_x && h->L.s.n > 0) || (!d_y && h->L.s.m > 0)) { set_error("dasp_spmv_f16_f32out: NULL argument"); return DASP_ERR_INVALID; }
    if (h->dtype != DASP_F16) { set_error("dasp_spmv_f16_f32out: the handle is not FP16"); return DASP_ERR_INVALID; }
    DASP_ON_DEVICE(h->device);
    return launch_spmv(h, d_x, d_y, permuted ? nullptr : h->L.order_rid, (cudaStream_t)stream, nullptr, nullptr, true);
}

int dasp_spmv_axpby(dasp_handle *h, double alpha, const void *d_x, double beta, void *d_y, int permuted, void *stream)
{
    if (!h || (!d_x && h->L.s.n > 0) || (!d_y && h->L.s.m > 0)) { set_error("dasp_spmv_axpby: NULL argument"); return DASP_ERR_INVALID; }
    const double ab[2] = {alpha, beta};
    DASP_ON_DEVICE(h->device);
    return launch_spmv(h, d_x, d_y, permuted ? nullptr : h->L.order_rid, (cudaStream_t)stream, ab);
}

int dasp_spmv_scatter_to(dasp_handle *h, const void *d_x, void *const *d_dests, int n_dests, int64_t row_offset,
                         const double *d_norm2, void *stream)
{
    if (!h || (!d_x && h->L.s.n > 0) || !d_dests || n_dests < 1 || n_dests > 8 || row_offset < 0) {
        set_error("dasp_spmv_scatter_to: bad argument (1..8 destinations)");
        return DASP_ERR_INVALID;
    }
    ScatterTo m{};
    for (int p = 0; p < n_dests; p++) {
        if (!d_dests[p]) { set_error("dasp_spmv_scatter_to: destination %d is NULL", p); return DASP_ERR_INVALID; }
        if (p > 0) m.extra[p - 1] = d_dests[p];
    }
    m.n_extra = n_dests - 1;
    m.row_offset = row_offset;
    m.norm2 = d_norm2;
    DASP_ON_DEVICE(h->device);
    return launch_spmv(h, d_x, d_dests[0], h->L.order_rid, (cudaStream_t)stream, nullptr, &m);
}

int dasp_spmv_permuted_to(dasp_handle *h, const void *d_x, void *const *d_dests, int n_dests, int64_t row_offset,
                          const double *d_norm2, void *stream)
{
    if (!h || (!d_x && h->L.s.n > 0) || !d_dests || n_dests < 1 || n_dests > 8 || row_offset < 0) {
        set_error("dasp_spmv_permuted_to: bad argument (1..8 destinations)");
        return DASP_ERR_INVALID;
    }
    ScatterTo m{};
    for (int p = 0; p < n_dests; p++) {
        if (!d_dests[p]) { set_error("dasp_spmv_permuted_to: destination %d is NULL", p); return DASP_ERR_INVALID; }
        if (p > 0) m.extra[p - 1] = d_dests[p];
    }
    m.n_extra = n_dests - 1;
    m.row_offset = row_offset;
    m.norm2 = d_norm2;
    DASP_ON_DEVICE(h->device);
    return launch_spmv(h, d_x, d_dests[0], nullptr, (cudaStream_t)stream, nullptr, &m);
}

int dasp_relabel_columns(dasp_handle *h, const int *d_new_index, int n_new)
{
    if (!h || (!d_new_index && h->L.s.n > 0) || n_new < 0) { set_error("dasp_relabel_columns: bad argument"); return DASP_ERR_INVALID; }
    if (!is_device_ptr(d_new_index) && h->L.s.n > 0) { set_error("dasp_relabel_columns: new_index must be a device pointer"); return DASP_ERR_INVALID; }
    DASP_ON_DEVICE(h->device);
    DASP_CUDA(cudaDeviceSynchronize()); // new_index may have been produced on another stream; no product may be in flight
    return relabel_columns(h, d_new_index, n_new, h->own_stream);
}

int dasp_inverse_order(const dasp_handle *h, const int **d_inv_order)
{
    if (!h || !d_inv_order) { set_error("dasp_inverse_order: NULL argument"); return DASP_ERR_INVALID; }
    *d_inv_order = h->L.inv_order;
    return DASP_OK;
}

int dasp_unpermute_to(dasp_handle *h, const void *d_y_perm, void *const *d_dests, int n_dests, int64_t row_offset,
                      const double *d_norm2, void *stream)
{
    if (!h || (!d_y_perm && h->L.s.m > 0) || !d_dests || n_dests < 1 || n_dests > 8 || row_offset < 0) {
        set_error("dasp_unpermute_to: bad argument (1..8 destinations)");
        return DASP_ERR_INVALID;
    }
    ScatterTo m{};
    for (int p = 0; p < n_dests; p++) {
        if (!d_dests[p]) { set_error("dasp_unpermute_to: destination %d is NULL", p); return DASP_ERR_INVALID; }
        if (p > 0) m.extra[p - 1] = d_dests[p];
    }
    m.n_extra = n_dests - 1;
    m.row_offset = row_offset;
    m.norm2 = d_norm2;
    DASP_ON_DEVICE(h->device);
    return unpermute_to(h, d_y_perm, m, d_dests[0], (cudaStream_t)stream);
}

// ---- checkpoint of the preprocessed layout ---------------------------------------------------------
// The file holds the reference layout only (the 12 bit-exact arrays + the scalars); everything else the kernels read is
// re-derived on the GPU by derive() after loading, from arrays that were range-checked first.
namespace {
constexpr uint64_t kMagic = 0x3230305f50534144ull; // "DASP_002"
constexpr int32_t kFormatVersion = 2;
constexpr int kFileArrays = 12;
struct ArrayRef { void **ptr; int64_t bytes; };

// the reference arrays of the layout with their sizes, in file order
void layout_arrays(dasp_handle *h, ArrayRef (&out)[kFileArrays])
{
    Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    const int64_t ev = (int64_t)L.esz, ei = sizeof(int);
    const ArrayRef a[kFileArrays] = {
        {(void **)&L.order_rid, ei * s.m},
        {(void **)&L.long_rpt_new, ei * ((int64_t)s.row_long + 1)},
        {&L.long_val, ev * s.fill0_nnz_long},
        {(void **)&L.long_cid, ei * s.fill0_nnz_long},
        {(void **)&L.blockPtr, ei * ((int64_t)s.blocknum + 1)},
        {(void **)&L.irreg_rpt, ei * ((int64_t)s.row_block + 1)},
        {&L.irreg_val, ev * s.fill0_nnz_irreg},
        {(void **)&L.irreg_cid, ei * s.nnz_irreg},
        {&L.reg_val, ev * s.fill0_nnz_reg},
        {(void **)&L.reg_cid, ei * s.fill0_nnz_reg},
        {&L.short_val, ev * s.fill0_nnz_short},
        {(void **)&L.short_cid, ei * s.fill0_nnz_short},
    };
    for (int i = 0; i < kFileArrays; i++) out[i] = a[i];
}

// scalars of a file: every count non-negative and consistent with m, n, nnz and with each other
bool stats_plausible(const dasp_stats_t &s, int dtype, int block_longest)
{
    const int f16 = dtype == DASP_F16;
    const int64_t ints[] = {s.m, s.n, s.row_long, s.row_block, s.row_zero, s.short_row_1, s.short_row_3, s.short_row_2, s.short_row_4,
                            s.common_13, s.short_row_34, s.blocknum, s.warp_number, s.fill0_nnz_long, s.fill0_nnz_reg, s.nnz_irreg,
                            s.fill0_nnz_short, s.fill0_nnz_short13, s.fill0_nnz_short34, s.fill0_nnz_short22, s.nnz_short,
                            s.nnz_long, s.fill0_nnz_irreg, s.origin_nnz_reg};
    for (int64_t v : ints)
        if (v < 0) return false;
    if (s.dtype != dtype || s.nnz < 0 || s.nnz > INT32_MAX || block_longest < 1) return false;
    if (s.rowloop != 1 && s.rowloop != 2 && s.rowloop != 4) return false;
    const int64_t rows = (int64_t)s.row_long + s.row_block + s.short_row_1 + 2 * (int64_t)s.common_13 + s.short_row_34 +
                         s.short_row_2 + s.row_zero;
    if (rows != s.m || s.short_row_34 != s.short_row_3 + s.short_row_4) return false;
    if ((int64_t)s.nnz_long + s.nnz_short + s.origin_nnz_reg + s.nnz_irreg != s.nnz) return false; // src/dasp_f64.h:1091
    if (s.blocknum % (4 * s.rowloop) || (int64_t)s.blocknum * 8 < s.row_block) return false;
    if (s.warp_number % 4 || (int64_t)s.warp_number * (f16 ? 256 : 64) != s.fill0_nnz_long) return false;
    if (s.fill0_nnz_reg % 32 || s.fill0_nnz_irreg != (f16 ? 2 * ((s.nnz_irreg + 1) / 2) : s.nnz_irreg)) return false;
    const int64_t singles = f16 ? 2 * (((int64_t)s.short_row_1 + 1) / 2) : s.short_row_1;
    if (singles + s.fill0_nnz_short13 + s.fill0_nnz_short34 + s.fill0_nnz_short22 != s.fill0_nnz_short) return false;
    if (s.fill0_nnz_short13 % 32 || s.fill0_nnz_short34 % 32 || s.fill0_nnz_short22 % 32) return false;
    if ((int64_t)s.fill0_nnz_short13 < 4 * (int64_t)s.common_13 || (int64_t)s.fill0_nnz_short34 < 4 * (int64_t)s.short_row_34 ||
        (int64_t)s.fill0_nnz_short22 < 2 * (int64_t)s.short_row_2)
        return false;
    return true;
}
} // namespace

int dasp_save(const dasp_handle *h, const char *path)
{
    if (!h || !path) { set_error("dasp_save: NULL argument"); return DASP_ERR_INVALID; }
    DASP_ON_DEVICE(h->device);
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("dasp_save: cannot open %s", path); return DASP_ERR_INVALID; }
    ArrayRef arr[kFileArrays];
    layout_arrays(const_cast<dasp_handle *>(h), arr);
    const int32_t head[6] = {(int32_t)h->dtype, h->block_longest, kFileArrays, (int32_t)sizeof(dasp_stats_t), kFormatVersion, 0};
    bool ok = fwrite(&kMagic, 8, 1, f) == 1 && fwrite(head, sizeof(head), 1, f) == 1 && fwrite(&h->threshold, 8, 1, f) == 1 &&
              fwrite(&h->L.s, sizeof(dasp_stats_t), 1, f) == 1;
    std::vector<char> buf;
    try {
        for (int i = 0; ok && i < kFileArrays; i++) {
            ok = fwrite(&arr[i].bytes, 8, 1, f) == 1;
            if (!ok || arr[i].bytes == 0) continue;
            buf.resize((size_t)arr[i].bytes);
            if (cudaMemcpy(buf.data(), *arr[i].ptr, buf.size(), cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
            ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
        }
    } catch (const std::bad_alloc &) { ok = false; }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { set_error("dasp_save: write to %s failed", path); return DASP_ERR_INVALID; }
    return DASP_OK;
}

int dasp_load(dasp_handle **out, const char *path, int device)
{
    if (!out || !path) { set_error("dasp_load: NULL argument"); return DASP_ERR_INVALID; }
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("dasp_load: cannot open %s", path); return DASP_ERR_INVALID; }
    struct Closer { FILE *f; ~Closer() { fclose(f); } } closer{f};
    uint64_t magic = 0;
    int32_t head[6] = {0, 0, 0, 0, 0, 0};
    double threshold = 0;
    dasp_stats_t st;
    const bool ok = fread(&magic, 8, 1, f) == 1 && magic == kMagic && fread(head, sizeof(head), 1, f) == 1 &&
                    head[2] == kFileArrays && head[3] == (int32_t)sizeof(dasp_stats_t) && head[4] == kFormatVersion &&
                    fread(&threshold, 8, 1, f) == 1 && fread(&st, sizeof(st), 1, f) == 1 &&
                    (head[0] == DASP_F64 || head[0] == DASP_F16);
    if (!ok) { set_error("dasp_load: %s is not a DASP layout file of format %d", path, kFormatVersion); return DASP_ERR_INVALID; }
    if (!(threshold > 0.0) || !stats_plausible(st, head[0], head[1])) {
        set_error("dasp_load: %s: inconsistent layout scalars", path);
        return DASP_ERR_INVALID;
    }
    DASP_ON_DEVICE(device);
    dasp_handle *h = new (std::nothrow) dasp_handle();
    if (!h) { set_error("out of host memory"); return DASP_ERR_ALLOC; }
    h->device = device; h->dtype = (dasp_dtype)head[0]; h->block_longest = head[1]; h->threshold = threshold;
    h->L.s = st; h->L.esz = head[0] == DASP_F16 ? 2 : 8;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    int rc = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) == cudaSuccess ? DASP_OK : DASP_ERR_CUDA;
    ArrayRef arr[kFileArrays];
    layout_arrays(h, arr);
    std::vector<char> buf;
    try {
        for (int i = 0; rc == DASP_OK && i < kFileArrays; i++) {
            int64_t bytes = -1;
            if (fread(&bytes, 8, 1, f) != 1 || bytes != arr[i].bytes) { set_error("dasp_load: %s is truncated or inconsistent", path); rc = DASP_ERR_INVALID; break; }
            if ((rc = h->pool.alloc(arr[i].ptr, (size_t)bytes)) != DASP_OK) break;
            if (bytes == 0) continue;
            buf.resize((size_t)bytes);
            if (fread(buf.data(), 1, buf.size(), f) != buf.size()) { set_error("dasp_load: %s is truncated", path); rc = DASP_ERR_INVALID; break; }
            if (cudaMemcpy(*arr[i].ptr, buf.data(), buf.size(), cudaMemcpyHostToDevice) != cudaSuccess) { set_error("dasp_load: upload failed: %s", cudaGetErrorString(cudaGetLastError())); rc = DASP_ERR_CUDA; }
        }
    } catch (const std::bad_alloc &) { set_error("dasp_load: out of host memory"); rc = DASP_ERR_ALLOC; }
    char extra;
    if (rc == DASP_OK && fread(&extra, 1, 1, f) == 1) { set_error("dasp_load: %s has trailing data", path); rc = DASP_ERR_INVALID; }
    // the offsets and indices the kernels will follow: checked on the device before anything is derived from them
    if (rc == DASP_OK) rc = validate_layout(h, h->own_stream);
    if (rc == DASP_OK) rc = derive(h, h->own_stream);
    if (rc != DASP_OK) { dasp_destroy(h); return rc; }
    h->L.s.device_bytes = h->pool.bytes;
    *out = h;
    return DASP_OK;
}

int dasp_spmv_timed(dasp_handle *h, const void *d_x, void *d_y, void *stream, int warmup, int reps, float *total_ms)
{
    if (!h || !total_ms || reps < 1 || warmup < 0) { set_error("dasp_spmv_timed: bad argument"); return DASP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    DASP_ON_DEVICE(h->device);
    for (int i = 0; i < warmup; i++) DASP_TRY(launch_spmv(h, d_x, d_y, nullptr, st));
    cudaEvent_t e0, e1;
    DASP_CUDA(cudaEventCreate(&e0));
    DASP_CUDA(cudaEventCreate(&e1));
    DASP_CUDA(cudaEventRecord(e0, st));
    int rc = DASP_OK;
    for (int i = 0; i < reps && rc == DASP_OK; i++) rc = launch_spmv(h, d_x, d_y, nullptr, st);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaEventSynchronize(e1);
    if (rc == DASP_OK && e != cudaSuccess) { set_error("dasp_spmv_timed: %s", cudaGetErrorString(e)); rc = DASP_ERR_CUDA; }
    if (rc == DASP_OK) cudaEventElapsedTime(total_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

int dasp_spmv_host(dasp_handle *h, const void *x_host, void *y_host)
{
    if (!h || (!x_host && h->L.s.n > 0) || (!y_host && h->L.s.m > 0)) { set_error("dasp_spmv_host: NULL argument"); return DASP_ERR_INVALID; }
    DASP_ON_DEVICE(h->device);
    const size_t esz = h->dtype == DASP_F16 ? 2 : 8;
    const size_t xb = esz * (size_t)h->L.s.n, yb = esz * (size_t)h->L.s.m;
    if (!h->dx_stage) { // device-side staging buffers, allocated once; committed only when both exist
        void *dx = nullptr, *dy = nullptr;
        DASP_TRY(h->pool.alloc(&dx, xb));
        int rc = h->pool.alloc(&dy, yb);
        if (rc != DASP_OK) { h->pool.release(dx); return rc; }
        h->dx_stage = dx; h->dy_stage = dy;
    }
    cudaStream_t st = h->own_stream;
    // only the columns that occur in the matrix are read by the product: upload x[col_min .. col_max] (a row slab of a
    // banded / stencil matrix touches 1/P of x plus a halo)
    const size_t xoff = esz * (size_t)h->L.s.col_min;
    const size_t xlen = h->L.s.col_max >= h->L.s.col_min ? esz * ((size_t)h->L.s.col_max - h->L.s.col_min + 1) : 0;
    if (xlen) DASP_CUDA(cudaMemcpyAsync((char *)h->dx_stage + xoff, (const char *)x_host + xoff, xlen, cudaMemcpyHostToDevice, st));
    DASP_TRY(launch_spmv(h, h->dx_stage, h->dy_stage, nullptr, st));
    DASP_CUDA(cudaMemcpyAsync(y_host, h->dy_stage, yb, cudaMemcpyDeviceToHost, st));
    DASP_CUDA(cudaStreamSynchronize(st));
    return DASP_OK;
}

int dasp_spmv_host_batch(dasp_handle *h, const void *const *x_hosts, void *const *y_hosts, int count)
{
    if (!h || count < 0 || (count > 0 && (!x_hosts || !y_hosts))) { set_error("dasp_spmv_host_batch: bad argument"); return DASP_ERR_INVALID; }
    if (count == 0) return DASP_OK;
    DASP_ON_DEVICE(h->device);
    const size_t esz = h->dtype == DASP_F16 ? 2 : 8;
    const size_t xb = esz * (size_t)h->L.s.n, yb = esz * (size_t)h->L.s.m;
    const size_t xoff = esz * (size_t)h->L.s.col_min; // only the column range of the matrix is uploaded (see dasp_spmv_host)
    const size_t xlen = h->L.s.col_max >= h->L.s.col_min ? esz * ((size_t)h->L.s.col_max - h->L.s.col_min + 1) : 0;
    if (!h->batch_ready) { // streams, events and staging are committed only when all of them exist
        for (int k = 0; k < 3; k++)
            if (!h->batch_stream[k]) DASP_CUDA(cudaStreamCreateWithFlags(&h->batch_stream[k], cudaStreamNonBlocking));
        for (int k = 0; k < 3; k++)
            for (int b = 0; b < 2; b++)
                if (!h->batch_ev[k][b]) DASP_CUDA(cudaEventCreateWithFlags(&h->batch_ev[k][b], cudaEventDisableTiming));
        for (int b = 0; b < 2; b++) {
            if (!h->batch_dx[b]) DASP_TRY(h->pool.alloc(&h->batch_dx[b], xb));
            if (!h->batch_dy[b]) DASP_TRY(h->pool.alloc(&h->batch_dy[b], yb));
        }
        h->batch_ready = 1;
    }
    cudaStream_t up = h->batch_stream[0], comp = h->batch_stream[1], down = h->batch_stream[2];
    cudaEvent_t(&ev)[3][2] = h->batch_ev; // [0] upload done, [1] kernel done, [2] download done, per staging buffer
    for (int i = 0; i < count; i++) {
        const int b = i & 1;
        if (!x_hosts[i] && xb) { set_error("dasp_spmv_host_batch: x_hosts[%d] is NULL", i); return DASP_ERR_INVALID; }
        if (!y_hosts[i] && yb) { set_error("dasp_spmv_host_batch: y_hosts[%d] is NULL", i); return DASP_ERR_INVALID; }
        if (i >= 2) DASP_CUDA(cudaStreamWaitEvent(up, ev[1][b], 0)); // x staging b free once product i-2 was multiplied
        if (xlen) DASP_CUDA(cudaMemcpyAsync((char *)h->batch_dx[b] + xoff, (const char *)x_hosts[i] + xoff, xlen, cudaMemcpyHostToDevice, up));
        DASP_CUDA(cudaEventRecord(ev[0][b], up));
        DASP_CUDA(cudaStreamWaitEvent(comp, ev[0][b], 0));
        if (i >= 2) DASP_CUDA(cudaStreamWaitEvent(comp, ev[2][b], 0)); // y staging b free once product i-2 was downloaded
        DASP_TRY(launch_spmv(h, h->batch_dx[b], h->batch_dy[b], nullptr, comp));
        DASP_CUDA(cudaEventRecord(ev[1][b], comp));
        DASP_CUDA(cudaStreamWaitEvent(down, ev[1][b], 0));
        DASP_CUDA(cudaMemcpyAsync(y_hosts[i], h->batch_dy[b], yb, cudaMemcpyDeviceToHost, down));
        DASP_CUDA(cudaEventRecord(ev[2][b], down));
    }
    DASP_CUDA(cudaStreamSynchronize(down));
    DASP_CUDA(cudaStreamSynchronize(up));
    DASP_CUDA(cudaStreamSynchronize(comp));
    return DASP_OK;
}

int dasp_order(const dasp_handle *h, const int **d_order_rid)
{
    if (!h || !d_order_rid) { set_error("dasp_order: NULL argument"); return DASP_ERR_INVALID; }
    *d_order_rid = h->L.order_rid;
    return DASP_OK;
}

int dasp_stats(const dasp_handle *h, dasp_stats_t *out)
{
    if (!h || !out) { set_error("dasp_stats: NULL argument"); return DASP_ERR_INVALID; }
    *out = h->L.s;
    return DASP_OK;
}

int dasp_report(const dasp_handle *h, const char *label, double spmv_ms, char *out, int64_t cap)
{
    if (!h || !out || cap <= 0 || !(spmv_ms > 0.0)) { set_error("dasp_report: bad argument"); return DASP_ERR_INVALID; }
    const dasp_stats_t &s = h->L.s;
    const double gflops = (double)(s.nnz * 2) / (spmv_ms * 1e6);
    const double bw1 = (double)s.data_X / (spmv_ms * 1e6), bw2 = (double)s.data_X2 / (spmv_ms * 1e6);
    int k = snprintf(out, (size_t)cap, "%s,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,", label ? label : "", s.m, s.n,
                     (int)s.nnz, s.short_row_1, s.common_13, s.short_row_3, s.short_row_4, s.short_row_2, s.row_long,
                     s.row_block, s.nnz_short, s.fill0_nnz_short, s.nnz_long, s.fill0_nnz_long, s.origin_nnz_reg,
                     s.fill0_nnz_reg, s.nnz_irreg);
    if (k < 0 || k >= cap) { set_error("dasp_report: buffer too small"); return DASP_ERR_BUFFER; }
    int k2;
    if (h->dtype == DASP_F64)
        k2 = snprintf(out + k, (size_t)(cap - k), "%lf,%d,%lld,%lf,%lf,%lf,%lf,", s.rate_fill0, h->block_longest,
                      (long long)s.data_X, spmv_ms, gflops, bw1, bw2);
    else
        k2 = snprintf(out + k, (size_t)(cap - k), "%lf,%d,%lld,%lf,%lf,%lf,%lf,%lf,%lf,%lf,", s.rate_fill0, h->block_longest,
                      (long long)s.data_X, s.preprocess_ms, spmv_ms, gflops, spmv_ms, gflops, bw1, bw2);
    if (k2 < 0 || k2 >= cap - k) { set_error("dasp_report: buffer too small"); return DASP_ERR_BUFFER; }
    return k + k2;
}

int dasp_export(const dasp_handle *h, const char *name, void *host_dst, int64_t cap_bytes, int64_t *bytes)
{
    if (!h || !name) { set_error("dasp_export: NULL argument"); return DASP_ERR_INVALID; }
    const Layout &L = h->L;
    const dasp_stats_t &s = L.s;
    const int64_t ev = (int64_t)L.esz, ei = sizeof(int);
    struct Entry { const char *name; const void *ptr; int64_t bytes; };
    const Entry table[] = {
        {"order_rid", L.order_rid, ei * s.m},
        {"long_rpt_new", L.long_rpt_new, ei * ((int64_t)s.row_long + 1)},
        {"long_val", L.long_val, ev * s.fill0_nnz_long},
        {"long_cid", L.long_cid, ei * s.fill0_nnz_long},
        {"blockPtr", L.blockPtr, ei * ((int64_t)s.blocknum + 1)},
        {"irreg_rpt", L.irreg_rpt, ei * ((int64_t)s.row_block + 1)},
        {"irreg_val", L.irreg_val, ev * s.fill0_nnz_irreg},
        {"irreg_cid", L.irreg_cid, ei * s.nnz_irreg},
        {"reg_val", L.reg_val, ev * s.fill0_nnz_reg},
        {"reg_cid", L.reg_cid, ei * s.fill0_nnz_reg},
        {"short_val", L.short_val, ev * s.fill0_nnz_short},
        {"short_cid", L.short_cid, ei * s.fill0_nnz_short},
    };
    for (const Entry &e : table) {
        if (strcmp(e.name, name)) continue;
        if (bytes) *bytes = e.bytes;
        if (!host_dst) return DASP_OK;
        if (cap_bytes < e.bytes) { set_error("dasp_export(%s): need %lld bytes, got %lld", name, (long long)e.bytes, (long long)cap_bytes); return DASP_ERR_BUFFER; }
        DASP_ON_DEVICE(h->device);
        if (e.bytes > 0) DASP_CUDA(cudaMemcpy(host_dst, e.ptr, (size_t)e.bytes, cudaMemcpyDeviceToHost));
        return DASP_OK;
    }
    set_error("dasp_export: unknown array '%s'", name);
    return DASP_ERR_INVALID;
}

int dasp_set_variant(dasp_handle *h, dasp_variant medium, dasp_variant long_rows, dasp_variant short_rows)
{
    if (!h) { set_error("dasp_set_variant: NULL handle"); return DASP_ERR_INVALID; }
    h->var_medium = medium; h->var_long = long_rows; h->var_short = short_rows;
    if (long_rows == DASP_VARIANT_BLOCKED && !h->L.lcb_val) { // build the column-blocked copy on demand
        DASP_ON_DEVICE(h->device);
        DASP_TRY(build_lcb(h, h->own_stream));
    }
    if (short_rows == DASP_VARIANT_BANDED && !h->L.sb_item) { // and the short-band work list
        DASP_ON_DEVICE(h->device);
        DASP_TRY(build_short_bands(h, h->own_stream, true));
    }
    if (medium == DASP_VARIANT_BANDED && !h->L.mb_lo) { // and the medium-band windows
        DASP_ON_DEVICE(h->device);
        DASP_TRY(build_medium_bands(h, h->own_stream));
    }
    return DASP_OK;
}

int dasp_set_index_compression(dasp_handle *h, int on)
{
    if (!h) { set_error("dasp_set_index_compression: NULL handle"); return DASP_ERR_INVALID; }
    h->index_compression = on ? 1 : 0;
    return DASP_OK;
}

int dasp_set_category_mask(dasp_handle *h, int mask)
{
    if (!h) { set_error("dasp_set_category_mask: NULL handle"); return DASP_ERR_INVALID; }
    h->category_mask = mask & 15;
    return DASP_OK;
}

int dasp_launches_per_spmv(const dasp_handle *h) { return h ? launches_per_spmv(h) : 0; }

static int spmv_all_impl(dasp_dtype dt, const void *val, const int *rowptr, const int *colidx, const void *x, void *y,
                         int *order_rid, int m, int n, int nnz, double threshold, int block_longest)
{
    dasp_handle *h = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError())); return DASP_ERR_CUDA; }
    DASP_TRY(dasp_create(&h, dt, dev, m, n, nnz, rowptr, colidx, val, threshold, block_longest));
    int rc = dasp_spmv_host(h, x, y);
    if (rc == DASP_OK && order_rid) rc = dasp_export(h, "order_rid", order_rid, (int64_t)sizeof(int) * m, nullptr);
    dasp_destroy(h);
    return rc;
}

int dasp_spmv_all_f64(const char *, const double *csrValA, const int *csrRowPtrA, const int *csrColIdxA, const double *X_val,
                      double *Y_val, int *order_rid, int rowA, int colA, int nnzA, int, double threshold, int block_longest)
{
    return spmv_all_impl(DASP_F64, csrValA, csrRowPtrA, csrColIdxA, X_val, Y_val, order_rid, rowA, colA, nnzA, threshold,
                         block_longest);
}

int dasp_spmv_all_f16(const char *, const void *csrValA, const int *csrRowPtrA, const int *csrColIdxA, const void *X_val,
                      void *Y_val, int *order_rid, int rowA, int colA, int nnzA, int, double threshold, int block_longest)
{
    return spmv_all_impl(DASP_F16, csrValA, csrRowPtrA, csrColIdxA, X_val, Y_val, order_rid, rowA, colA, nnzA, threshold,
                         block_longest);
}

int dasp_sumsq(const double *d_v, int64_t count, double *d_out, void *stream)
{
    if ((!d_v && count > 0) || count < 0 || !d_out) { set_error("dasp_sumsq: bad argument"); return DASP_ERR_INVALID; }
    return sumsq(d_v, count, d_out, (cudaStream_t)stream);
}

int dasp_scale_rsqrt(double *d_v, int64_t count, const double *d_norm2, void *stream)
{
    if ((!d_v && count > 0) || count < 0 || !d_norm2) { set_error("dasp_scale_rsqrt: bad argument"); return DASP_ERR_INVALID; }
    return scale_by_rsqrt(d_v, count, d_norm2, (cudaStream_t)stream);
}

int dasp_scale_copy_to(const double *d_v, int64_t count, void *const *d_dests, int n_dests, int64_t offset, const double *d_norm2,
                       void *stream)
{
    if ((!d_v && count > 0) || count < 0 || !d_dests || n_dests < 1 || n_dests > 8 || offset < 0) {
        set_error("dasp_scale_copy_to: bad argument (1..8 destinations)");
        return DASP_ERR_INVALID;
    }
    for (int p = 0; p < n_dests; p++)
        if (!d_dests[p]) { set_error("dasp_scale_copy_to: destination %d is NULL", p); return DASP_ERR_INVALID; }
    return scale_copy_to(d_v, count, d_dests, n_dests, offset, d_norm2, (cudaStream_t)stream);
}

int dasp_partition_rows(int m, const int *rowptr, int parts, int *cuts)
{
    if (m < 0 || !rowptr || parts < 1 || !cuts) { set_error("dasp_partition_rows: bad argument"); return DASP_ERR_INVALID; }
    const int64_t nnz = rowptr[m] - (int64_t)rowptr[0];
    cuts[0] = 0;
    for (int p = 1; p < parts; p++) {
        const int64_t target = rowptr[0] + nnz * p / parts;
        const int *it = std::lower_bound(rowptr + cuts[p - 1], rowptr + m + 1, target,
                                         [](int v, int64_t t) { return (int64_t)v < t; });
        int c = (int)(it - rowptr);
        cuts[p] = c > m ? m : c;
    }
    cuts[parts] = m;
    return DASP_OK;
}

} // extern "C"
