// microbench_gather.cu -- how fast can a B200 SM do divergent 4-byte gathers?  Decides the strategy for the
// power-law configuration (R-MAT, BASELINE config 3), where x[col] is a random access into a 64 MB vector.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench_gather tools/microbench_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ float ld_nc_na(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int MODE, int U> // 0: __ldg, 1: L1::no_allocate, 2: shared table
__global__ void __launch_bounds__(256) gather_kernel(const int *__restrict__ idx, long long n, const float *__restrict__ x, int table, float *out)
{
    extern __shared__ float sx[];
    if(MODE == 2)
    {
        for(int i = threadIdx.x; i < table; i += blockDim.x)
            sx[i] = x[i];
        __syncthreads();
    }
    float           acc    = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x * U;
    for(long long i = (long long)blockIdx.x * blockDim.x * U + threadIdx.x; i < n; i += stride)
    {
        int c[U];
#pragma unroll
        for(int u = 0; u < U; ++u)
            c[u] = (i + u * blockDim.x < n) ? idx[i + u * blockDim.x] : 0;
#pragma unroll
        for(int u = 0; u < U; ++u)
        {
            if(MODE == 0)
                acc += __ldg(x + c[u]);
            else if(MODE == 1)
                acc += ld_nc_na(x + c[u]);
            else
                acc += sx[c[u]];
        }
    }
    if(acc == 123.456f)
        out[0] = acc;
}

int main()
{
    const long long   n = 1LL << 28; // gathers per launch
    std::vector<int>  h(n);
    int              *d_idx;
    float            *d_x, *d_out;
    cudaMalloc(&d_idx, n * 4);
    cudaMalloc(&d_x, 1LL << 30);
    cudaMalloc(&d_out, 4);
    cudaMemset(d_x, 0, 1LL << 30);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(gather_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("%-28s %14s %12s %12s\n", "case", "footprint", "Ggather/s", "per SM/clk@1.9GHz");
    const long long foots[] = {16LL << 10, 64LL << 10, 1LL << 20, 16LL << 20, 64LL << 20, 256LL << 20};
    for(long long fb : foots)
    {
        const long long entries = fb / 4;
        unsigned long long s = 88172645463325252ull;
        for(long long i = 0; i < n; ++i)
        {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            h[i] = (int)(s % (unsigned long long)entries);
        }
        cudaMemcpy(d_idx, h.data(), n * 4, cudaMemcpyHostToDevice);
        for(int mode = 0; mode < 3; ++mode)
        {
            if(mode == 2 && fb > (128LL << 10))
                continue;
            float best = 1e30f;
            for(int rep = 0; rep < 4; ++rep)
            {
                cudaEventRecord(e0);
                if(mode == 0)
                    gather_kernel<0, 8><<<148 * 8, 256>>>(d_idx, n, d_x, 0, d_out);
                else if(mode == 1)
                    gather_kernel<1, 8><<<148 * 8, 256>>>(d_idx, n, d_x, 0, d_out);
                else
                    gather_kernel<2, 8><<<148 * 1, 256, fb>>>(d_idx, n, d_x, (int)entries, d_out);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if(ms < best)
                    best = ms;
            }
            const double g = (double)n / (best * 1e-3) / 1e9;
            printf("%-28s %11lld KB %12.1f %12.3f  %s\n", mode == 0 ? "ldg (L1 allocate)" : (mode == 1 ? "ld.nc L1::no_allocate" : "shared table (1 CTA/SM)"),
                   fb >> 10, g, g / (148 * 1.9), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
