// tools/compare.cu — comparison harness of SURVEY.md §8(f)-3: for one benchmark configuration, cuSPARSE CSR SpMV called
// exactly as the reference's comparator calls it (src/main_f64.cu:50-79: CUSPARSE_SPMV_ALG_DEFAULT, 32-bit indices,
// CUDA_R_64F; src/main_f16.cu:52-59: FP16 storage with CUDA_R_32F compute, float alpha/beta) next to this library
// (dasp_create + dasp_spmv through the C ABI), same device, same inputs, same timing protocol (warm-up launches, then
// launches back to back between two CUDA events), and the reference's verification of one against the other through
// order_rid (verify_new, src/main_f64.cu:3-16: |y_cusparse[order_rid[i]] - y_dasp[i]| <= 1e-5, FP16: <= 1.0 in float,
// src/main_f16.cu:3-18).  Bench tooling: links cuSPARSE, which the product library never does.
//
//   compare <c1|c2|c3|c3_spec|c4|c5|c5_spec> [scale=1.0] [grid=256]      -> one JSON line on stdout
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cusparse.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/dasp.h"
#include "../include/dasp_synth.h"

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(2); } \
    } while (0)
#define CS(call)                                                                                          \
    do {                                                                                                  \
        cusparseStatus_t s_ = (call);                                                                     \
        if (s_ != CUSPARSE_STATUS_SUCCESS) { fprintf(stderr, "%s:%d cusparse status %d\n", __FILE__, __LINE__, (int)s_); exit(2); } \
    } while (0)

static dasp_synth_spec make_spec(const char *w, double scale, int grid)
{
    dasp_synth_spec s;
    memset(&s, 0, sizeof(s));
    if (!strcmp(w, "c4")) {
        s.kind = 0; s.nx = s.ny = s.nz = grid; s.m = s.n = (int64_t)grid * grid * grid; s.seed = 20240004;
    } else if (!strcmp(w, "c1") || !strcmp(w, "c2")) {
        s.kind = 3; s.m = s.n = 121192; s.seed = 7; s.mean_len = 22; s.window = 2048;
    } else if (!strcmp(w, "c3") || !strcmp(w, "c3_spec")) {
        s.kind = strcmp(w, "c3") ? 4 : 1; s.m = s.n = (int64_t)(10000000 * scale); s.seed = 20240001; s.alpha = 0.95;
        s.lmax = (int)(s.m / 2 < 1000000 ? s.m / 2 : 1000000); s.window = 4096;
    } else if (!strcmp(w, "c5") || !strcmp(w, "c5_spec")) {
        const int64_t nshort = (int64_t)(50000000 * scale);
        s.n_long = (int)(1000 * scale) > 0 ? (int)(1000 * scale) : 1;
        s.m = s.n = s.n_long + nshort; s.seed = 20240005; s.window = 4096;
        s.long_len = (int)(s.m / 2 < 1000000 ? s.m / 2 : 1000000);
        if (!strcmp(w, "c5")) {
            s.kind = 2; s.long_len = 1000000;
            int64_t band = 1, top = 1;
            while (band < 2 * (int64_t)s.long_len) band <<= 1;
            while (top * 2 <= s.m) top <<= 1;
            s.band = band < top ? band : top; s.band_lo = (s.m - s.band) / 2;
        } else
            s.kind = 5;
    } else {
        fprintf(stderr, "unknown workload %s\n", w);
        exit(2);
    }
    return s;
}

__global__ void fill_x(double *x, int64_t n, uint64_t seed)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t z = seed + 0x9e3779b97f4a7c15ull * (uint64_t)(i + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    x[i] = 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: compare <c1|c2|c3|c3_spec|c4|c5|c5_spec> [scale] [grid]\n"); return 2; }
    const char *w = argv[1];
    const double scale = argc > 2 ? atof(argv[2]) : 1.0;
    const int grid = argc > 3 ? atoi(argv[3]) : 256;
    const bool half = !strcmp(w, "c2");
    const dasp_synth_spec spec = make_spec(w, scale, grid);
    const int64_t m = spec.m, n = spec.n;
    const size_t esz = half ? 2 : 8;

    // ---- the matrix, generated on the device (row lengths -> host prefix sum -> fill) ----
    int *d_len, *d_rp;
    CK(cudaMalloc(&d_len, sizeof(int) * (size_t)m));
    CK(cudaMalloc(&d_rp, sizeof(int) * (size_t)(m + 1)));
    if (dasp_synth_rowlen(&spec, 0, m, d_len, nullptr)) { fprintf(stderr, "synth: %s\n", dasp_synth_last_error()); return 2; }
    std::vector<int> len((size_t)m), rp((size_t)m + 1);
    CK(cudaMemcpy(len.data(), d_len, sizeof(int) * (size_t)m, cudaMemcpyDeviceToHost));
    int64_t run = 0;
    for (int64_t i = 0; i < m; i++) { rp[(size_t)i] = (int)run; run += len[(size_t)i]; }
    rp[(size_t)m] = (int)run;
    const int64_t nnz = run;
    if (nnz > 2147483647LL) { fprintf(stderr, "nnz exceeds 32-bit row pointers\n"); return 2; }
    CK(cudaMemcpy(d_rp, rp.data(), sizeof(int) * (size_t)(m + 1), cudaMemcpyHostToDevice));
    CK(cudaFree(d_len));
    int *d_ci;
    double *d_v64;
    void *d_val;
    CK(cudaMalloc(&d_ci, sizeof(int) * (size_t)nnz));
    CK(cudaMalloc(&d_v64, sizeof(double) * (size_t)nnz));
    if (dasp_synth_fill(&spec, 0, m, d_rp, d_ci, d_v64, nullptr)) { fprintf(stderr, "synth: %s\n", dasp_synth_last_error()); return 2; }
    double *d_x64;
    CK(cudaMalloc(&d_x64, sizeof(double) * (size_t)n));
    fill_x<<<(unsigned)((n + 255) / 256), 256>>>(d_x64, n, 7);
    void *d_x, *d_y_cu, *d_y_da;
    if (half) {
        CK(cudaMalloc(&d_val, 2 * (size_t)nnz));
        CK(cudaMalloc(&d_x, 2 * (size_t)n));
        dasp_synth_to_half(d_v64, d_val, nnz, nullptr);
        dasp_synth_to_half(d_x64, d_x, n, nullptr);
        CK(cudaDeviceSynchronize());
        CK(cudaFree(d_v64));
        CK(cudaFree(d_x64));
    } else {
        d_val = d_v64;
        d_x = d_x64;
    }
    CK(cudaMalloc(&d_y_cu, esz * (size_t)m));
    CK(cudaMalloc(&d_y_da, esz * (size_t)m));
    CK(cudaMemset(d_y_cu, 0, esz * (size_t)m));
    CK(cudaDeviceSynchronize());

    const double b_alg = (double)nnz * (esz + 4) + (double)(m + 1) * 4 + (double)n * esz + (double)m * esz; // src/main_f64.cu:143
    const bool small = b_alg < 256e6;
    const int warm = small ? 100 : 5, reps = small ? 1000 : 20; // the reference's 100 + 1000 for matrices of its own size
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));

    // ---- cuSPARSE, the reference's calls ----
    cusparseHandle_t handle;
    cusparseSpMatDescr_t matA;
    cusparseDnVecDescr_t vecX, vecY;
    const cudaDataType vt = half ? CUDA_R_16F : CUDA_R_64F, ct = half ? CUDA_R_32F : CUDA_R_64F;
    const double alpha_d = 1.0, beta_d = 0.0;
    const float alpha_f = 1.0f, beta_f = 0.0f;
    const void *alpha = half ? (const void *)&alpha_f : (const void *)&alpha_d, *beta = half ? (const void *)&beta_f : (const void *)&beta_d;
    CS(cusparseCreate(&handle));
    CS(cusparseCreateCsr(&matA, m, n, nnz, d_rp, d_ci, d_val, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, vt));
    CS(cusparseCreateDnVec(&vecX, n, d_x, vt));
    CS(cusparseCreateDnVec(&vecY, m, d_y_cu, vt));
    size_t buf_bytes = 0;
    void *d_buf = nullptr;
    CS(cusparseSpMV_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, matA, vecX, beta, vecY, ct, CUSPARSE_SPMV_ALG_DEFAULT, &buf_bytes));
    CK(cudaMalloc(&d_buf, buf_bytes ? buf_bytes : 16));
    for (int i = 0; i < warm; i++)
        CS(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, matA, vecX, beta, vecY, ct, CUSPARSE_SPMV_ALG_DEFAULT, d_buf));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++)
        CS(cusparseSpMV(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, alpha, matA, vecX, beta, vecY, ct, CUSPARSE_SPMV_ALG_DEFAULT, d_buf));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float cu_ms = 0.f;
    CK(cudaEventElapsedTime(&cu_ms, e0, e1));
    cu_ms /= reps;

    // ---- this library through its C ABI ----
    dasp_handle *h = nullptr;
    int rc = dasp_create(&h, half ? DASP_F16 : DASP_F64, 0, (int)m, (int)n, nnz, d_rp, d_ci, d_val, 0.75, 256);
    if (rc) { fprintf(stderr, "dasp_create: %s: %s\n", dasp_strerror(rc), dasp_last_error()); return 2; }
    float da_total = 0.f;
    rc = dasp_spmv_timed(h, d_x, d_y_da, nullptr, warm, reps, &da_total);
    if (rc) { fprintf(stderr, "dasp_spmv_timed: %s: %s\n", dasp_strerror(rc), dasp_last_error()); return 2; }
    const float da_ms = da_total / reps;

    // ---- verify_new: cuSPARSE result against ours through order_rid ----
    std::vector<int> order((size_t)m);
    rc = dasp_export(h, "order_rid", order.data(), (int64_t)sizeof(int) * m, nullptr);
    if (rc) { fprintf(stderr, "dasp_export: %s\n", dasp_last_error()); return 2; }
    std::vector<unsigned char> ycu(esz * (size_t)m), yda(esz * (size_t)m);
    CK(cudaMemcpy(ycu.data(), d_y_cu, ycu.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(yda.data(), d_y_da, yda.size(), cudaMemcpyDeviceToHost));
    double max_abs = 0.0, num = 0.0, den = 0.0;
    long bad = 0;
    const double thr = half ? 1.0 : 1e-5; // src/main_f16.cu:10, src/main_f64.cu:8
    for (int64_t i = 0; i < m; i++) {
        double a, b;
        if (half) {
            a = (double)__half2float(reinterpret_cast<const __half *>(ycu.data())[order[(size_t)i]]);
            b = (double)__half2float(reinterpret_cast<const __half *>(yda.data())[i]);
        } else {
            a = reinterpret_cast<const double *>(ycu.data())[order[(size_t)i]];
            b = reinterpret_cast<const double *>(yda.data())[i];
        }
        const double d = fabs(a - b);
        if (d > max_abs) max_abs = d;
        if (!(d <= thr)) bad++;
        num += d * d; den += a * a;
    }
    dasp_stats_t st;
    dasp_stats(h, &st);
    printf("{\"workload\": \"%s\", \"dtype\": \"%s\", \"m\": %lld, \"nnz\": %lld, \"warmup\": %d, \"reps\": %d, "
           "\"cusparse_ms\": %.6f, \"cusparse_gflops\": %.2f, \"cusparse_hbm_gbs\": %.1f, \"cusparse_buffer_bytes\": %zu, "
           "\"dasp_ms\": %.6f, \"dasp_gflops\": %.2f, \"dasp_hbm_gbs\": %.1f, \"speedup_vs_cusparse\": %.3f, "
           "\"verify_new\": {\"threshold\": %g, \"rows_over_threshold\": %ld, \"max_abs_diff\": %.3e, \"rel_l2\": %.3e}, "
           "\"dasp_preprocess_ms\": %.3f, \"l2\": \"%s\"}\n",
           w, half ? "f16" : "f64", (long long)m, (long long)nnz, warm, reps, cu_ms, 2.0 * nnz / (cu_ms * 1e6), b_alg / (cu_ms * 1e6),
           buf_bytes, da_ms, 2.0 * nnz / (da_ms * 1e6), b_alg / (da_ms * 1e6), cu_ms / da_ms, thr, bad, max_abs,
           den > 0 ? sqrt(num / den) : 0.0, st.preprocess_ms, small ? "warm (back-to-back launches on an L2-resident matrix)" : "inputs exceed L2");
    dasp_destroy(h);
    cusparseDestroySpMat(matA);
    cusparseDestroyDnVec(vecX);
    cusparseDestroyDnVec(vecY);
    cusparseDestroy(handle);
    return bad ? 1 : 0;
}
