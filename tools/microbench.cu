// tools/microbench.cu — ceilings on the box: launch rate, streaming-read bandwidth
// for L2-resident and HBM-resident buffers with the load flavours the SpMV kernels use.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void empty_kernel() {}

template <int MODE> // 0: plain 128-bit, 1: nc no_allocate 256-bit, 2: nc no_allocate evict_first 256-bit
__global__ void __launch_bounds__(256) read_kernel(const double *__restrict__ p, size_t n4, double *out)
{
    double acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        double a, b, c, d;
        if (MODE == 0) {
            double2 u = *reinterpret_cast<const double2 *>(p + 4 * i), v = *reinterpret_cast<const double2 *>(p + 4 * i + 2);
            a = u.x; b = u.y; c = v.x; d = v.y;
        } else if (MODE == 1) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + 4 * i));
        } else {
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + 4 * i));
        }
        acc += a + b + c + d;
    }
    if (acc == 1.2345e300) *out = acc;
}

template <typename F> float time_loop(F f, int warm, int reps)
{
    for (int i = 0; i < warm; i++) f();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps * 1e3f; // us
}

int main()
{
    double *out; cudaMalloc(&out, 8);
    printf("empty kernel launch, back to back: %.2f us\n", time_loop([&] { empty_kernel<<<1, 32>>>(); }, 100, 2000));
    printf("empty kernel 475x256:             %.2f us\n", time_loop([&] { empty_kernel<<<475, 256>>>(); }, 100, 2000));
    size_t sizes[] = {17u << 20, 34u << 20, 68u << 20, 4096ull << 20};
    for (size_t bytes : sizes) {
        double *p; cudaMalloc(&p, bytes); cudaMemset(p, 0, bytes);
        size_t n4 = bytes / 32;
        int grids[] = {148, 148 * 4, 148 * 8, 148 * 16, (int)((n4 + 255) / 256)};
        for (int g : grids) {
            int reps = bytes > (1u << 30) ? 20 : 2000;
            float t0 = time_loop([&] { read_kernel<0><<<g, 256>>>(p, n4, out); }, 20, reps);
            float t1 = time_loop([&] { read_kernel<1><<<g, 256>>>(p, n4, out); }, 20, reps);
            float t2 = time_loop([&] { read_kernel<2><<<g, 256>>>(p, n4, out); }, 20, reps);
            printf("read %6zu MB grid %8d: plain128 %8.2f us %7.0f GB/s | nc.na.256 %8.2f us %7.0f GB/s | +evict_first %8.2f us %7.0f GB/s\n",
                   bytes >> 20, g, t0, bytes / t0 / 1e3, t1, bytes / t1 / 1e3, t2, bytes / t2 / 1e3);
        }
        cudaFree(p);
    }
    return 0;
}
