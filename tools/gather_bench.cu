// tools/gather_bench.cu — the gather floor of the box: how many scattered 8-byte (and 2-byte) loads per clock one SM can
// retire, from L1-resident, L2-resident and shared-memory-resident x.  Every lane of a warp instruction reads its own
// random element, so a global gather costs the L1 one wavefront per lane; the shared-memory form costs bank-conflict
// cycles instead.  This is the bound of every DASP row category whose columns do not coalesce (short rows and
// length-sorted medium rows of the power-law / skewed shapes): see profiles/r02/README.md.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/gather_bench tools/gather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ unsigned mix(unsigned z)
{
    z ^= z >> 16; z *= 0x7feb352du; z ^= z >> 15; z *= 0x846ca68bu; z ^= z >> 16;
    return z;
}

template <typename T, bool SMEM>
__global__ void __launch_bounds__(256) gather_kernel(const T *__restrict__ x, unsigned mask, int iters, double *out)
{
    extern __shared__ unsigned char sm[];
    T *xs = reinterpret_cast<T *>(sm);
    if (SMEM) {
        for (unsigned i = threadIdx.x; i <= mask; i += blockDim.x) xs[i] = x[i];
        __syncthreads();
    }
    unsigned s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    double acc = 0;
    for (int it = 0; it < iters; it += 8) {
        T v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            s = s * 1664525u + 1013904223u;
            const unsigned c = mix(s) & mask;
            v[j] = SMEM ? xs[c] : __ldg(x + c);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc += (double)v[j];
    }
    if (acc == 1.2345e300) *out = acc;
}

template <typename T, bool SMEM> void run(const char *what, size_t elems, int sms, double ghz)
{
    T *x;
    double *out;
    cudaMalloc(&x, elems * sizeof(T));
    cudaMemset(x, 0, elems * sizeof(T));
    cudaMalloc(&out, 8);
    const int iters = 4096, grid = sms * 8;
    const size_t smem = SMEM ? elems * sizeof(T) : 0;
    if (SMEM) cudaFuncSetAttribute(gather_kernel<T, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) gather_kernel<T, SMEM><<<grid, 256, smem>>>(x, (unsigned)elems - 1, iters, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; r++) gather_kernel<T, SMEM><<<grid, 256, smem>>>(x, (unsigned)elems - 1, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gathers = 10.0 * grid * 256.0 * iters, rate = gathers / (ms * 1e-3);
    printf("%-44s %8.1f G gathers/s  %6.3f per clock per SM  (%s)\n", what, rate / 1e9, rate / (ghz * 1e9 * sms),
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(x); cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const double ghz = p.clockRate / 1e6;
    printf("%s, %d SMs, %.3f GHz (nominal max)\n", p.name, sms, ghz);
    run<double, false>("f64 global, x = 64 KB (L1-resident)", 8192, sms, ghz);
    run<double, false>("f64 global, x = 8 MB (L2-resident)", 1u << 20, sms, ghz);
    run<double, false>("f64 global, x = 64 MB (L2-resident)", 8u << 20, sms, ghz);
    run<double, false>("f64 global, x = 512 MB (HBM)", 64u << 20, sms, ghz);
    run<double, true>("f64 shared memory, x = 64 KB", 8192, sms, ghz);
    run<float, false>("4-byte global, x = 8 MB (L2-resident)", 2u << 20, sms, ghz);
    return 0;
}
