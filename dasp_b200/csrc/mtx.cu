// mtx.cu — Matrix Market coordinate reader with the semantics of the reference's mmio_allinone
// (src/mmio_highlevel.h:608-774, banner/size parsing of src/mmio.h), host only.  Row "next (f)-1" of the
// scope table: needed only so that file inputs give the identical CSR, hence identical DASP layouts.
// Differences in mechanism, not in result: the file is read in one piece and tokenised with strtol/strtod
// instead of one fscanf per entry.
#include <ctype.h>
#include <cuda_fp16.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "dasp_internal.h"

namespace {

std::string lower(std::string s)
{
    for (char &c : s) c = (char)tolower((unsigned char)c);
    return s;
}

// next whitespace-separated token as integer / double; advances p; false at end of data or on garbage
bool next_long(const char *&p, const char *end, long &out)
{
    while (p < end && isspace((unsigned char)*p)) p++;
    if (p >= end) return false;
    char *q;
    errno = 0;
    out = strtol(p, &q, 10);
    if (q == p) return false;
    p = q;
    return true;
}
bool next_double(const char *&p, const char *end, double &out)
{
    while (p < end && isspace((unsigned char)*p)) p++;
    if (p >= end) return false;
    char *q;
    out = strtod(p, &q);
    if (q == p) return false;
    p = q;
    return true;
}

template <typename T> inline T cast_val(double v);
template <> inline double cast_val<double>(double v) { return v; }
template <> inline unsigned short cast_val<unsigned short>(double v)
{
    __half h = __double2half(v); // round to nearest even, as the reference's implicit double -> half conversion
    unsigned short b;
    memcpy(&b, &h, 2);
    return b;
}

template <typename T>
int build(const std::vector<int> &ri, const std::vector<int> &cj, const std::vector<double> &vv, int m, bool sym,
          int64_t *nnz_out, int **rowptr, int **colidx, void **val)
{
    const size_t k = ri.size();
    std::vector<int64_t> cnt((size_t)m + 1, 0);
    for (size_t i = 0; i < k; i++) {
        cnt[ri[i]]++;
        if (sym && ri[i] != cj[i]) cnt[cj[i]]++;
    }
    int64_t run = 0;
    for (int r = 0; r <= m; r++) { int64_t c = cnt[r]; cnt[r] = run; run += c; } // cnt[r] = row start, cnt[m] = nnz
    const int64_t nnz = cnt[m];
    if (nnz > INT32_MAX) { dasp::set_error("expanded nnz %lld exceeds 32-bit row pointers", (long long)nnz); return DASP_ERR_RANGE; }
    int *rp = (int *)malloc(sizeof(int) * ((size_t)m + 1));
    int *ci = (int *)malloc(sizeof(int) * (size_t)(nnz ? nnz : 1));
    T *va = (T *)malloc(sizeof(T) * (size_t)(nnz ? nnz : 1));
    if (!rp || !ci || !va) { free(rp); free(ci); free(va); dasp::set_error("out of host memory"); return DASP_ERR_ALLOC; }
    for (int r = 0; r <= m; r++) rp[r] = (int)cnt[r];
    std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1); // next free slot of each row
    for (size_t i = 0; i < k; i++) {
        const T v = cast_val<T>(vv[i]);
        int64_t o = fill[ri[i]]++;
        ci[o] = cj[i]; va[o] = v;
        if (sym && ri[i] != cj[i]) { // the mirror goes to its row right now: file order is preserved per row
            o = fill[cj[i]]++;
            ci[o] = ri[i]; va[o] = v;
        }
    }
    *nnz_out = nnz; *rowptr = rp; *colidx = ci; *val = va;
    return DASP_OK;
}

} // namespace

extern "C" {

void dasp_free_host(void *p) { free(p); }

static int read_mtx_impl(const char *filename, dasp_dtype dtype, int *m, int *n, int64_t *nnz, int *is_symmetric, int **rowptr,
                         int **colidx, void **val)
{
    if (!filename || !m || !n || !nnz || !rowptr || !colidx || !val || (dtype != DASP_F64 && dtype != DASP_F16)) {
        dasp::set_error("dasp_read_mtx: bad argument");
        return DASP_ERR_INVALID;
    }
    FILE *f = fopen(filename, "rb");
    if (!f) { dasp::set_error("cannot open %s: %s", filename, strerror(errno)); return DASP_ERR_INVALID; }
    std::string data;
    {
        char buf[1 << 16];
        size_t got;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, got);
        fclose(f);
    }
    // banner: %%MatrixMarket matrix coordinate <field> <symmetry>   (src/mmio.h mm_read_banner; case-insensitive)
    size_t eol = data.find('\n');
    const std::string first = data.substr(0, eol == std::string::npos ? data.size() : eol);
    const std::string banner = lower(first);
    char w0[64], w1[64], w2[64], w3[64], w4[64];
    // the "%%MatrixMarket" token itself is compared case-sensitively by the reference (src/mmio.h:450)
    if (sscanf(banner.c_str(), "%63s %63s %63s %63s %63s", w0, w1, w2, w3, w4) != 5 ||
        strncmp(first.c_str() + strspn(first.c_str(), " \t"), "%%MatrixMarket", 14) || strcmp(w1, "matrix")) {
        dasp::set_error("%s: not a Matrix Market banner", filename);
        return DASP_ERR_INVALID;
    }
    if (strcmp(w2, "coordinate")) { dasp::set_error("%s: only coordinate format is supported (as in the reference)", filename); return DASP_ERR_INVALID; }
    const bool is_real = !strcmp(w3, "real"), is_complex = !strcmp(w3, "complex"), is_int = !strcmp(w3, "integer"),
               is_pat = !strcmp(w3, "pattern");
    if (!is_real && !is_complex && !is_int && !is_pat) { dasp::set_error("%s: unknown field '%s'", filename, w3); return DASP_ERR_INVALID; }
    const bool sym = !strcmp(w4, "symmetric") || !strcmp(w4, "hermitian"); // skew-symmetric is NOT expanded (:642)
    if (!sym && strcmp(w4, "general") && strcmp(w4, "skew-symmetric")) { dasp::set_error("%s: unknown symmetry '%s'", filename, w4); return DASP_ERR_INVALID; }
    // size line: skip comment lines (mm_read_mtx_crd_size)
    const char *p = data.c_str() + (eol == std::string::npos ? data.size() : eol + 1), *end = data.c_str() + data.size();
    while (p < end) {
        const char *q = p;
        while (q < end && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
        if (q < end && (*q == '%' || *q == '\n')) { while (p < end && *p != '\n') p++; if (p < end) p++; }
        else break;
    }
    long M, N, K;
    if (!next_long(p, end, M) || !next_long(p, end, N) || !next_long(p, end, K) || M < 0 || N < 0 || K < 0 || M > INT32_MAX ||
        N > INT32_MAX) {
        dasp::set_error("%s: bad size line", filename);
        return DASP_ERR_INVALID;
    }
    // the mirror of a symmetric / hermitian entry (i, j) is stored in row j: that row must exist (the reference writes
    // out of bounds there, src/mmio_highlevel.h:722-745)
    if (sym && M != N) { dasp::set_error("%s: symmetric matrix with %ld rows and %ld columns", filename, M, N); return DASP_ERR_INVALID; }
    // every entry needs at least "i j" plus a separator: a size line that promises more entries than the file can hold is
    // rejected before anything is allocated for it
    if ((unsigned long)K > (unsigned long)(end - p) / 4 + 1) { dasp::set_error("%s: size line announces %ld entries, the file is too short for that", filename, K); return DASP_ERR_INVALID; }
    std::vector<int> ri((size_t)K), cj((size_t)K);
    std::vector<double> vv((size_t)K);
    for (long i = 0; i < K; i++) {
        long a, b;
        double v = 1.0, im;
        bool ok = next_long(p, end, a) && next_long(p, end, b);
        if (ok && is_real) ok = next_double(p, end, v);
        else if (ok && is_complex) ok = next_double(p, end, v) && next_double(p, end, im);
        else if (ok && is_int) { long iv; ok = next_long(p, end, iv); v = (double)(int)iv; }
        if (!ok || a < 1 || a > M || b < 1 || b > N) {
            dasp::set_error("%s: malformed or out-of-range entry %ld", filename, i + 1);
            return DASP_ERR_INVALID;
        }
        ri[i] = (int)a - 1; cj[i] = (int)b - 1; vv[i] = v;
    }
    *m = (int)M; *n = (int)N;
    if (is_symmetric) *is_symmetric = sym ? 1 : 0;
    if (dtype == DASP_F16) return build<unsigned short>(ri, cj, vv, (int)M, sym, nnz, rowptr, colidx, val);
    return build<double>(ri, cj, vv, (int)M, sym, nnz, rowptr, colidx, val);
}

// no exception crosses the C ABI: allocation failures of the parser become a status code
int dasp_read_mtx(const char *filename, dasp_dtype dtype, int *m, int *n, int64_t *nnz, int *is_symmetric, int **rowptr,
                  int **colidx, void **val)
{
    try {
        return read_mtx_impl(filename, dtype, m, n, nnz, is_symmetric, rowptr, colidx, val);
    } catch (const std::bad_alloc &) {
        dasp::set_error("dasp_read_mtx: out of host memory");
        return DASP_ERR_ALLOC;
    } catch (const std::exception &e) {
        dasp::set_error("dasp_read_mtx: %s", e.what());
        return DASP_ERR_ALLOC;
    }
}

} // extern "C"
