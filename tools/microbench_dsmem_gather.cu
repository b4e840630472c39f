// microbench_dsmem_gather.cu -- how fast can a B200 SM do divergent 4-byte gathers out of a table that is spread over the
// shared memory of a thread-block cluster (ld.shared::cluster), alone and while the same threads also gather from a
// 64 MB vector in global memory (L1 misses, served by L2)?  Decides whether a cluster-wide hot-column table can help
// the power-law configuration (BASELINE config 3).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/microbench_dsmem_gather tools/microbench_dsmem_gather.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
namespace cg = cooperative_groups;

__device__ __forceinline__ float ld_cluster(unsigned addr)
{
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// idx[i] >= 0: gather x[idx[i]] from global memory; idx[i] < 0: slot = idx[i] & 0x7fffffff of the cluster table, which is
// interleaved over the CTAs of the cluster (rank = slot % CLUSTER, offset = slot / CLUSTER; CLUSTER a power of two)
template <int U>
__global__ void __launch_bounds__(1024, 1) gather_kernel(int cluster_log2, int entries, const int *__restrict__ idx, long long n,
                                                         const float *__restrict__ x, float *out)
{
    extern __shared__ float table[];
    cg::cluster_group cl = cg::this_cluster();
    for(int i = threadIdx.x; i < entries; i += blockDim.x)
        table[i] = (float)i;
    cl.sync();
    const unsigned base  = (unsigned)__cvta_generic_to_shared(table);
    const unsigned rmask = (1u << cluster_log2) - 1u;
    float           acc    = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x * U;
    for(long long i = (long long)blockIdx.x * blockDim.x * U + threadIdx.x; i < n; i += stride)
    {
        int c[U];
#pragma unroll
        for(int u = 0; u < U; ++u)
            c[u] = (i + u * blockDim.x < n) ? idx[i + u * blockDim.x] : 0;
        float v[U];
#pragma unroll
        for(int u = 0; u < U; ++u)
        {
            if(c[u] < 0)
            {
                const unsigned s = (unsigned)c[u] & 0x7fffffffu;
                unsigned       a;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(base + (s >> cluster_log2) * 4u), "r"(s & rmask));
                v[u] = ld_cluster(a);
            }
            else
                v[u] = __ldg(x + c[u]);
        }
#pragma unroll
        for(int u = 0; u < U; ++u)
            acc += v[u];
    }
    if(acc == 123.456f)
        out[0] = acc;
    cl.sync(); // nobody leaves while its table may still be read
}

int main()
{
    float *d_x, *d_out;
    const long long n_x = 1LL << 24; // 64 MB of floats
    cudaMalloc(&d_x, n_x * 4);
    cudaMalloc(&d_out, 4);
    cudaMemset(d_x, 0, n_x * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto kern = gather_kernel<8>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const long long n = 1LL << 25; // gathers per launch
    std::vector<int> h(n);
    int             *d_idx;
    cudaMalloc(&d_idx, n * 4);
    printf("%-8s %-10s %-8s %8s %12s %14s %14s  %s\n", "cluster", "table/CTA", "hot %", "CTAs", "ms", "Ggather/s", "per SM/clk", "status");
    const int clusters[] = {1, 2, 4, 8, 16};
    const int table_kb[] = {48, 96};
    const int hots[]     = {256, 128, 84, 64, 0}; // 100 %, 50 %, 33 %, 25 %, 0 % of the gathers go to the table
    for(int tk : table_kb)
        for(int cs : clusters)
        {
            const int entries = tk * 1024 / 4;
            const size_t smem = (size_t)entries * 4;
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim           = dim3(1024);
            cfg.dynamicSmemBytes   = smem;
            cudaLaunchAttribute at[1];
            at[0].id               = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs              = at;
            cfg.numAttrs           = 1;
            cfg.gridDim            = dim3(cs);
            int max_clusters       = 0;
            cudaError_t oe         = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
            if(oe != cudaSuccess || max_clusters < 1)
            {
                printf("%-8d %-10d cluster size not schedulable (%s)\n", cs, tk, cudaGetErrorString(oe));
                cudaGetLastError();
                continue;
            }
            const int ctas = max_clusters * cs; // one wave
            cfg.gridDim    = dim3(ctas);
            int cs_log2 = 0;
            while((1 << cs_log2) < cs)
                ++cs_log2;
            for(int hot : hots)
            {
                {
                    unsigned long long sd = 88172645463325252ull;
                    const unsigned     slots = (unsigned)cs * (unsigned)entries;
                    for(long long i = 0; i < n; ++i)
                    {
                        sd ^= sd << 13;
                        sd ^= sd >> 7;
                        sd ^= sd << 17;
                        const unsigned r = (unsigned)(sd >> 32);
                        h[i] = ((int)(sd & 255u) < hot) ? (int)(0x80000000u | (r % slots)) : (int)(r & (unsigned)(n_x - 1));
                    }
                    cudaMemcpy(d_idx, h.data(), n * 4, cudaMemcpyHostToDevice);
                }
                float best = 1e30f;
                cudaError_t le = cudaSuccess;
                for(int rep = 0; rep < 3; ++rep)
                {
                    cudaEventRecord(e0);
                    le = cudaLaunchKernelEx(&cfg, kern, cs_log2, entries, (const int *)d_idx, n, (const float *)d_x, d_out);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if(ms < best)
                        best = ms;
                }
                const double total = (double)n;
                const double g     = total / (best * 1e-3) / 1e9;
                printf("%-8d %-10d %-8.1f %8d %12.4f %14.1f %14.3f  %s %s\n", cs, tk, hot / 2.56, ctas, best, g, g / (ctas * 1.9),
                       cudaGetErrorString(le), cudaGetErrorString(cudaGetLastError()));
            }
        }
    return 0;
}
