// microbench_rowload.cu -- how fast does a B200 SM deliver 256-byte rows of a dense operand (csrmm, n = 32 doubles) out
// of L1 / L2 / shared memory, depending on how a warp's lanes are laid over the rows?  Decides the load shape of the
// row-major csrmm kernel (BASELINE config 4), which sits at ~66 B/clk/SM with 8 lanes x LDG.128 per row.
//   A: 8 lanes per row, 2 x LDG.128 per lane  (warp instruction spans 4 rows = 4 lines)        [the round-1 kernel]
//   B: 32 lanes per row, LDG.64 per lane      (warp instruction = one row = 2 lines)
//   C: 16 lanes per row, LDG.128 per lane     (warp instruction spans 2 rows = 4 lines)
//   D: 32 lanes per row, LDG.32 x 2           (warp instruction = half a row = 1 line)
//   S: B's shape out of shared memory (LDS.64), rows staged once per CTA
// Row indices come from a table in shared memory (like the staged column indices); `rows` bounds the footprint
// (256 rows = 64 KB: L1 hits; 1 M rows = 256 MB: L2 / HBM).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench_rowload tools/microbench_rowload.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int NT = 256, NIDX = 2048, ROWB = 256;

template <int MODE>
__global__ void __launch_bounds__(NT) rowload_kernel(const int *__restrict__ idx, int iters, const double *__restrict__ B, double *out, int srows)
{
    __shared__ int sidx[NIDX];
    extern __shared__ __align__(16) double sB[];
    for(int i = threadIdx.x; i < NIDX; i += NT)
        sidx[i] = idx[(blockIdx.x * NIDX + i) % (NIDX * 64)];
    if(MODE == 4)
        for(int i = threadIdx.x; i < srows * 32; i += NT)
            sB[i] = B[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double    a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for(int it = 0; it < iters; ++it)
    {
        // every warp walks its own 256 of the 2048 staged indices, 4 rows per step
        for(int j = warp * 256; j < warp * 256 + 256; j += 8)
        {
            if(MODE == 0)
            {
#pragma unroll
                for(int u = 0; u < 2; ++u)
                {
                    const int     r = sidx[j + u * 4 + (lane >> 3)];
                    const double *p = B + (long long)r * 32 + (lane & 7) * 2;
                    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
                    const double2 w = __ldg(reinterpret_cast<const double2 *>(p + 16));
                    a0 += v.x; a1 += v.y; a2 += w.x; a3 += w.y;
                }
            }
            else if(MODE == 1)
            {
#pragma unroll
                for(int u = 0; u < 8; ++u)
                {
                    const int r = sidx[j + u];
                    a0 += __ldg(B + (long long)r * 32 + lane);
                }
            }
            else if(MODE == 2)
            {
#pragma unroll
                for(int u = 0; u < 4; ++u)
                {
                    const int     r = sidx[j + u * 2 + (lane >> 4)];
                    const double2 v = __ldg(reinterpret_cast<const double2 *>(B + (long long)r * 32 + (lane & 15) * 2));
                    a0 += v.x; a1 += v.y;
                }
            }
            else if(MODE == 3)
            {
#pragma unroll
                for(int u = 0; u < 8; ++u)
                {
                    const int    r = sidx[j + u];
                    const float *p = reinterpret_cast<const float *>(B + (long long)r * 32);
                    a0 += __ldg(p + lane);
                    a1 += __ldg(p + 32 + lane);
                }
            }
            else
            {
#pragma unroll
                for(int u = 0; u < 8; ++u)
                {
                    const int r = sidx[j + u] % srows;
                    a0 += sB[r * 32 + lane];
                }
            }
        }
    }
    if(a0 + a1 + a2 + a3 == 123.456)
        out[0] = a0;
}

int main()
{
    const long long total_rows = 1 << 20;
    double         *dB, *dout;
    int            *didx;
    cudaMalloc(&dB, total_rows * ROWB);
    cudaMemset(dB, 0, total_rows * ROWB);
    cudaMalloc(&dout, 8);
    cudaMalloc(&didx, NIDX * 64 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char *names[5] = {"A  8 lanes/row 2xLDG.128", "B 32 lanes/row LDG.64", "C 16 lanes/row LDG.128", "D 32 lanes/row 2xLDG.32", "S shared LDS.64"};
    printf("%-28s %10s %8s %10s %12s\n", "shape", "rows", "CTAs/SM", "GB/s", "B/clk/SM@1.965");
    for(long long rows : {256LL, 4096LL, 1LL << 20})
    {
        std::vector<int> h(NIDX * 64);
        unsigned long long s = 88172645463325252ULL;
        for(auto &v : h)
        {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            v = (int)(s % (unsigned long long)rows);
        }
        cudaMemcpy(didx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        for(int mode = 0; mode < 5; ++mode)
        {
            if(mode == 4 && rows != 256)
                continue;
            for(int occ : {4, 8})
            {
                const int    grid  = 148 * occ;
                const int    iters = rows > 100000 ? 4 : 40;
                const int    srows = 128; // 32 KB of staged rows per CTA
                const size_t smem  = mode == 4 ? (size_t)srows * ROWB : 0;
                auto         launch = [&]() {
                    switch(mode)
                    {
                    case 0: rowload_kernel<0><<<grid, NT, smem>>>(didx, iters, dB, dout, srows); break;
                    case 1: rowload_kernel<1><<<grid, NT, smem>>>(didx, iters, dB, dout, srows); break;
                    case 2: rowload_kernel<2><<<grid, NT, smem>>>(didx, iters, dB, dout, srows); break;
                    case 3: rowload_kernel<3><<<grid, NT, smem>>>(didx, iters, dB, dout, srows); break;
                    default: rowload_kernel<4><<<grid, NT, smem>>>(didx, iters, dB, dout, srows); break;
                    }
                };
                launch();
                cudaDeviceSynchronize();
                cudaEventRecord(e0);
                for(int r = 0; r < 5; ++r)
                    launch();
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0;
                cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = 5.0 * grid * (double)iters * NIDX * ROWB;
                const double gbs   = bytes / (ms * 1e-3) / 1e9;
                printf("%-28s %10lld %8d %10.0f %12.1f  %s\n", names[mode], rows, occ, gbs, gbs * 1e9 / (148 * 1.965e9),
                       cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
