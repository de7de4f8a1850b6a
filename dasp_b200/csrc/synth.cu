// synth.cu — device-side generators for the benchmark matrices (include/dasp_synth.h).
// Bench/test tooling only; not linked into libdasp_b200.so.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/dasp_synth.h"

namespace {

thread_local char g_err[256] = "";
int fail(cudaError_t e, const char *what)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return -2;
}
#define SYNTH_LAUNCH_CHECK(what)                          \
    do {                                                  \
        cudaError_t e_ = cudaGetLastError();              \
        if (e_ != cudaSuccess) return fail(e_, what);     \
    } while (0)

// splitmix64 finaliser as a counter-based generator
__host__ __device__ inline uint64_t mix(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t h3(uint64_t seed, uint64_t a, uint64_t b) { return mix(mix(seed ^ mix(a)) + b); }
__host__ __device__ inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); } // [0,1)
__host__ __device__ inline double sym(uint64_t h) { return 2.0 * u01(h) - 1.0; }                              // [-1,1)

__device__ inline int64_t next_pow2(int64_t v)
{
    int64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------- stencil27
__device__ inline int cnt1(int c, int n) { return 1 + (c > 0) + (c < n - 1); }

__global__ void stencil_len(dasp_synth_spec s, int64_t row0, int64_t rows, int *len)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= rows) return;
    int64_t i = row0 + t;
    int x = (int)(i % s.nx), y = (int)((i / s.nx) % s.ny), z = (int)(i / ((int64_t)s.nx * s.ny));
    len[t] = cnt1(x, s.nx) * cnt1(y, s.ny) * cnt1(z, s.nz);
}

__global__ void stencil_fill(dasp_synth_spec s, int64_t row0, int64_t rows, const int *rowptr, int *colidx, double *val)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= rows) return;
    int64_t i = row0 + t;
    int x = (int)(i % s.nx), y = (int)((i / s.nx) % s.ny), z = (int)(i / ((int64_t)s.nx * s.ny));
    int64_t p = rowptr[t];
    int k = 0;
    for (int dz = -1; dz <= 1; dz++) {
        int z2 = z + dz;
        if (z2 < 0 || z2 >= s.nz) continue;
        for (int dy = -1; dy <= 1; dy++) {
            int y2 = y + dy;
            if (y2 < 0 || y2 >= s.ny) continue;
            for (int dx = -1; dx <= 1; dx++) {
                int x2 = x + dx;
                if (x2 < 0 || x2 >= s.nx) continue;
                colidx[p] = (int)(x2 + (int64_t)s.nx * (y2 + (int64_t)s.ny * z2));
                val[p] = sym(h3(s.seed, (uint64_t)i, (uint64_t)k));
                p++; k++;
            }
        }
    }
}

// ---------------------------------------------------------------- powerlaw / skewed (element-parallel fill)
__device__ inline int powerlaw_len(const dasp_synth_spec &s, int64_t i)
{
    double u = 1.0 - u01(h3(s.seed, (uint64_t)i, 0x51ull)); // (0,1]
    double l = floor(pow(u, -1.0 / s.alpha));
    if (!(l < (double)s.lmax)) l = (double)s.lmax;
    return l < 0 ? 0 : (int)l;
}

__device__ inline int skewed_len(const dasp_synth_spec &s, int64_t i)
{
    if (i < s.n_long) return s.long_len;
    return 1 + (int)(h3(s.seed, (uint64_t)i, 0x52ull) & 3);
}

// ---- SURVEY.md §8(d)-literal generators (kinds 4, 5): columns in RANDOM order, distinct within a row ----
// A keyed bijection of [0, 2^bits): odd multiplications, additions and xor-shifts are each invertible on a fixed
// width, so distinct inputs give distinct outputs ("uniform without replacement", evaluated per element).
__device__ inline uint64_t perm_pow2(uint64_t x, int bits, uint64_t key)
{
    const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
    const int sh = bits > 1 ? bits / 2 : 1;
    const uint64_t k1 = mix(key), k2 = mix(k1), k3 = mix(k2);
    x = (x * (k1 | 1) + (k1 >> 32)) & mask;
    x ^= x >> sh;
    x = (x * (k2 | 1) + (k2 >> 32)) & mask;
    x ^= x >> sh;
    x = (x * (k3 | 1) + (k3 >> 32)) & mask;
    x ^= x >> sh;
    return x;
}
// the same on [0, size), size <= 2^bits: cycle walking (expected < 2 rounds)
__device__ inline int64_t perm_range(int64_t j, int64_t size, uint64_t key)
{
    int bits = 1;
    while ((1ll << bits) < size) bits++;
    uint64_t x = (uint64_t)j;
    do x = perm_pow2(x, bits, key); while ((int64_t)x >= size);
    return (int64_t)x;
}
// window of `width` columns centred on the diagonal position of row i, clipped to [0, n)
__device__ inline int64_t window_lo(const dasp_synth_spec &s, int64_t i, int64_t width)
{
    int64_t c = (int64_t)((double)i * (double)s.n / (double)s.m);
    int64_t lo = c - width / 2;
    if (lo > s.n - width) lo = s.n - width;
    if (lo < 0) lo = 0;
    return lo;
}

// kind 5: one long row per stride of m / n_long rows, at a hashed position inside its stride
__device__ inline bool skewed_spec_is_long(const dasp_synth_spec &s, int64_t i)
{
    if (s.n_long <= 0) return false;
    const int64_t stride = s.m / s.n_long, j = i / stride;
    if (j >= s.n_long) return false;
    return i == j * stride + (int64_t)(h3(s.seed, (uint64_t)j, 0x10c5ull) % (uint64_t)stride);
}
__device__ inline int spec_len(const dasp_synth_spec &s, int64_t i)
{
    if (s.kind == 4) return powerlaw_len(s, i);
    if (skewed_spec_is_long(s, i)) return s.long_len;
    return 1 + (int)(h3(s.seed, (uint64_t)i, 0x52ull) & 3);
}
// column of element k of row i (length len)
__device__ inline int64_t spec_col(const dasp_synth_spec &s, int64_t i, int64_t k, int len)
{
    const uint64_t key = h3(s.seed ^ 0xc01ull, (uint64_t)i, 0x77ull);
    if (s.kind == 5 && len == s.long_len && skewed_spec_is_long(s, i)) return perm_range(k, s.n, key); // uniform over all columns
    if (s.kind == 5) { // short rows: distinct columns of the +-window window
        int64_t D = 2 * (int64_t)s.window;
        if (D > s.n) D = s.n;
        return window_lo(s, i, D) + perm_range(k, D, key);
    }
    // kind 4: every 10th element is a global column outside the window, the others are distinct columns of the window
    // (2*window wide; enlarged to the next power of two >= twice the windowed count for rows that would not fit)
    const int64_t lg = len / 10, lw = len - lg;
    int64_t D = 2 * (int64_t)s.window;
    while (D < 2 * lw) D <<= 1;
    if (D > s.n) D = s.n;
    const int64_t lo = window_lo(s, i, D);
    const bool glob = (k % 10 == 9) && (s.n - D >= 2 * lg);
    if (glob) {
        const int64_t q = perm_range(k / 10, s.n - D, key ^ 0x9e37ull);
        return q < lo ? q : q + D;
    }
    // windowed element index: elements that are not global keep their rank; if the global part had to be folded into
    // the window (tiny n), every element is windowed and k itself is the rank
    const int64_t j = (s.n - D >= 2 * lg) ? k - k / 10 : k;
    return lo + perm_range(j, D, key);
}

__global__ void plsk_len(dasp_synth_spec s, int64_t row0, int64_t rows, int *len)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= rows) return;
    len[t] = s.kind == 1 ? powerlaw_len(s, row0 + t) : (s.kind == 2 ? skewed_len(s, row0 + t) : spec_len(s, row0 + t));
}

// Windowed columns of a row are an ASCENDING sequence of distinct points of a window centred on the
// row's diagonal position (CSR built from sorted input has ascending columns): the window holds
// width = max(2*halfw, 2*len) columns, element k lands in stripe k of width/len columns at a hashed
// offset inside the first half of the stripe (so neighbours never collide).
__device__ inline int64_t window_col(const dasp_synth_spec &s, int64_t i, int64_t k, int len, int64_t halfw)
{
    int64_t W = 2 * (halfw > len ? halfw : (int64_t)len);
    if (W > s.n) W = s.n;
    int64_t c = (int64_t)((double)i * (double)s.n / (double)s.m);
    int64_t lo = c - W / 2;
    if (lo < 0) lo = 0;
    if (lo > s.n - W) lo = s.n - W;
    double stripe = (double)W / (double)len; // >= 2 unless the window was clipped to n
    double u = 0.5 * u01(h3(s.seed ^ 0x5eedull, (uint64_t)i, (uint64_t)k));
    int64_t off = (int64_t)(((double)k + u) * stripe);
    if (off >= W) off = W - 1;
    return lo + off;
}

__global__ void plsk_fill(dasp_synth_spec s, int64_t row0, int64_t rows, const int *rowptr, int64_t nnz, int *colidx,
                          double *val)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    // row of element e: last t with rowptr[t] <= e
    int64_t lo = 0, hi = rows;
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] <= e) lo = mid; else hi = mid;
    }
    const int64_t t = lo, i = row0 + t, k = e - rowptr[t];
    const int len = rowptr[t + 1] - rowptr[t];
    int64_t col;
    if (s.kind >= 4) {
        col = spec_col(s, i, k, len);
    } else if (s.kind == 2 && i < s.n_long) {
        // long rows of the skewed matrix: ascending, distinct, one entry per stripe of band/len columns of a
        // band shared by all long rows (each row starts at its own hashed shift inside the first stripe)
        double stripe = (double)s.band / (double)len;
        double u = 0.5 * u01(h3(s.seed ^ 0x5eedull, (uint64_t)i, (uint64_t)k));
        int64_t off = (int64_t)(((double)k + u) * stripe);
        if (off >= s.band) off = s.band - 1;
        col = s.band_lo + off;
    } else {
        uint64_t hk = h3(s.seed ^ 0xabcdefull, (uint64_t)i, (uint64_t)k);
        if (s.kind == 1 && hk % 10 == 0) col = (int64_t)((hk >> 8) % (uint64_t)s.n); // 10 % global
        else col = window_col(s, i, k, len, s.window);
    }
    colidx[e] = (int)col;
    val[e] = sym(h3(s.seed, (uint64_t)i, (uint64_t)k));
}

// ---------------------------------------------------------------- banded symmetric (cop20k_A stand-in)
// entry (i,j), j != i, exists iff hash(min,max) < p; p = (mean_len-1)/(2*window); diagonal always present.
__device__ inline bool band_edge(const dasp_synth_spec &s, int64_t i, int64_t j)
{
    if (j < 0 || j >= s.n) return false;
    if (i == j) return true;
    int64_t a = i < j ? i : j, b = i < j ? j : i;
    double p = (double)(s.mean_len - 1) / (2.0 * s.window);
    return u01(h3(s.seed, (uint64_t)a, (uint64_t)b)) < p;
}

template <bool FILL>
__global__ void band_kernel(dasp_synth_spec s, int64_t row0, int64_t rows, int *len, const int *rowptr, int *colidx,
                            double *val)
{
    const int lane = threadIdx.x & 31;
    int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; // one warp per row
    if (t >= rows) return;
    const int64_t i = row0 + t;
    int count = 0;
    int64_t p = FILL ? rowptr[t] : 0;
    for (int64_t j0 = i - s.window; j0 <= i + s.window; j0 += 32) {
        int64_t j = j0 + lane;
        bool ok = j <= i + s.window && band_edge(s, i, j);
        unsigned b = __ballot_sync(0xffffffffu, ok);
        if (FILL && ok) {
            int64_t q = p + count + __popc(b & ((1u << lane) - 1u));
            colidx[q] = (int)j;
            int64_t lo = i < j ? i : j, hi = i < j ? j : i;
            val[q] = sym(h3(s.seed ^ 0x77ull, (uint64_t)lo, (uint64_t)hi));
        }
        count += __popc(b);
    }
    if (!FILL && lane == 0) len[t] = count;
}

__global__ void to_half_kernel(const double *src, __half *dst, int64_t n)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) dst[t] = __double2half(src[t]);
}

__global__ void flush_kernel(uint4 *p, int64_t n)
{
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) p[t] = make_uint4((unsigned)t, 1u, 2u, 3u);
}

inline unsigned blocks(int64_t n, int th) { return (unsigned)((n + th - 1) / th); }

} // namespace

extern "C" {

const char *dasp_synth_last_error(void) { return g_err; }

int dasp_synth_rowlen(const dasp_synth_spec *spec, int64_t row0, int64_t row1, int *d_len, void *stream)
{
    if (!spec || row1 < row0 || !d_len) { snprintf(g_err, sizeof(g_err), "bad argument"); return -1; }
    const int64_t rows = row1 - row0;
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (spec->kind) {
    case 0: stencil_len<<<blocks(rows, 256), 256, 0, st>>>(*spec, row0, rows, d_len); break;
    case 1:
    case 2:
    case 4:
    case 5: plsk_len<<<blocks(rows, 256), 256, 0, st>>>(*spec, row0, rows, d_len); break;
    case 3: band_kernel<false><<<blocks(rows * 32, 256), 256, 0, st>>>(*spec, row0, rows, d_len, nullptr, nullptr, nullptr); break;
    default: snprintf(g_err, sizeof(g_err), "unknown kind %d", spec->kind); return -1;
    }
    SYNTH_LAUNCH_CHECK("rowlen");
    return 0;
}

int dasp_synth_fill(const dasp_synth_spec *spec, int64_t row0, int64_t row1, const int *d_rowptr, int *d_colidx,
                    double *d_val, void *stream)
{
    if (!spec || row1 < row0 || !d_rowptr) { snprintf(g_err, sizeof(g_err), "bad argument"); return -1; }
    const int64_t rows = row1 - row0;
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (spec->kind) {
    case 0: stencil_fill<<<blocks(rows, 128), 128, 0, st>>>(*spec, row0, rows, d_rowptr, d_colidx, d_val); break;
    case 1:
    case 2:
    case 4:
    case 5: {
        int nnz = 0;
        cudaError_t e = cudaMemcpyAsync(&nnz, d_rowptr + rows, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(e, "read nnz");
        if (nnz > 0) plsk_fill<<<blocks(nnz, 256), 256, 0, st>>>(*spec, row0, rows, d_rowptr, nnz, d_colidx, d_val);
        break;
    }
    case 3: band_kernel<true><<<blocks(rows * 32, 256), 256, 0, st>>>(*spec, row0, rows, nullptr, d_rowptr, d_colidx, d_val); break;
    default: snprintf(g_err, sizeof(g_err), "unknown kind %d", spec->kind); return -1;
    }
    SYNTH_LAUNCH_CHECK("fill");
    return 0;
}

int dasp_synth_to_half(const double *d_src, void *d_dst, int64_t count, void *stream)
{
    if (count > 0) to_half_kernel<<<blocks(count, 256), 256, 0, (cudaStream_t)stream>>>(d_src, (__half *)d_dst, count);
    SYNTH_LAUNCH_CHECK("to_half");
    return 0;
}

int dasp_synth_flush_l2(void *d_scratch, int64_t bytes, void *stream)
{
    int64_t n = bytes / 16;
    if (n > 0) flush_kernel<<<blocks(n, 256), 256, 0, (cudaStream_t)stream>>>((uint4 *)d_scratch, n);
    SYNTH_LAUNCH_CHECK("flush_l2");
    return 0;
}

} // extern "C"
