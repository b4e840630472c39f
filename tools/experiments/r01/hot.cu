// hot.cu -- analysis side of the hot-column table (kernel: spmv_hot.cuh).  Part of aoclsparse_optimize for a
// general, non-transposed mv hint on a matrix with skewed row lengths (memory policy permitting):
//   1. count how often every column occurs (one atomic per stored entry);
//   2. sort the counts (CUB radix sort, analysis time only) and keep the K most frequent columns, K chosen so that
//      the table of x values fills 96 KB of shared memory;
//   3. if those K columns cover at least 15 % of all stored entries, write a second column array in which they are
//      replaced by (HOT_BIT | slot); otherwise the matrix keeps using the ordinary kernel.
#include "spmv_hot.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

namespace b200
{
    namespace
    {
        inline unsigned grid_for(long long n, int tpb)
        {
            long long b = (n + tpb - 1) / tpb;
            if(b > 148LL * 32)
                b = 148LL * 32;
            return (unsigned)(b < 1 ? 1 : b);
        }
        __global__ void col_hist_kernel(long long nnz, const int *__restrict__ col, unsigned *cnt)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < nnz; i += (long long)gridDim.x * blockDim.x)
                atomicAdd(cnt + col[i], 1u);
        }
        __global__ void iota_u_kernel(long long n, int *out)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < n; i += (long long)gridDim.x * blockDim.x)
                out[i] = (int)i;
        }
        __global__ void slot_scatter_kernel(int k, const int *__restrict__ hot_cols, int *slot_of)
        {
            int i = blockIdx.x * blockDim.x + threadIdx.x;
            if(i < k)
                slot_of[hot_cols[i]] = i;
        }
        __global__ void remap_kernel(long long nnz, const int *__restrict__ col, const int *__restrict__ slot_of, int *col_hot)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < nnz; i += (long long)gridDim.x * blockDim.x)
            {
                const int c = col[i], s = slot_of[c];
                col_hot[i]  = s >= 0 ? (HOT_BIT | s) : c;
            }
        }
    }

    aoclsparse_status build_hot_table(dev_csr &A, size_t elem_size, cudaStream_t st)
    {
        row_block_plan &P = A.plan;
        P.hot_entries     = 0;
        if(A.nnz < (1 << 18) || A.n < (1 << 14))
            return aoclsparse_status_success; // small problems: x lives in L1/L2 anyway
        const bool two_ctas = getenv("AOCLSPARSE_B200_HOT_CTAS") && atoi(getenv("AOCLSPARSE_B200_HOT_CTAS")) == 2;
        P.hot_stages        = two_ctas ? 1 : 2;
        int K = (int)((two_ctas ? 40960 : 98304) / elem_size);
        if(K > A.n)
            K = A.n;
        const long long n = A.n, nnz = A.nnz;
        dev_buf cnt, cnt_sorted, ids, ids_sorted, temp, slot_of;
        B200_TRY(cnt.alloc(4 * (size_t)n));
        B200_TRY(cnt_sorted.alloc(4 * (size_t)n));
        B200_TRY(ids.alloc(4 * (size_t)n));
        B200_TRY(ids_sorted.alloc(4 * (size_t)n));
        B200_CUDA(cudaMemsetAsync(cnt.p, 0, 4 * (size_t)n, st));
        col_hist_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, A.col_idx.as<int>(), cnt.as<unsigned>());
        B200_LAUNCHED();
        iota_u_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, ids.as<int>());
        B200_LAUNCHED();
        size_t tb = 0;
        B200_CUDA(cub::DeviceRadixSort::SortPairsDescending(
            nullptr, tb, cnt.as<unsigned>(), cnt_sorted.as<unsigned>(), ids.as<int>(), ids_sorted.as<int>(), (int)n, 0, 32, st));
        B200_TRY(temp.alloc(tb));
        B200_CUDA(cub::DeviceRadixSort::SortPairsDescending(
            temp.p, tb, cnt.as<unsigned>(), cnt_sorted.as<unsigned>(), ids.as<int>(), ids_sorted.as<int>(), (int)n, 0, 32, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        std::vector<unsigned> top((size_t)K);
        B200_CUDA(cudaMemcpyAsync(top.data(), cnt_sorted.p, 4 * (size_t)K, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        long long mass = 0;
        for(unsigned c : top)
            mass += c;
        if(mass * 100 < 15 * nnz)
            return aoclsparse_status_success;
        B200_TRY(P.hot_cols.alloc(4 * (size_t)K));
        B200_TRY(P.col_hot.alloc(4 * (size_t)nnz));
        B200_TRY(slot_of.alloc(4 * (size_t)n));
        B200_CUDA(cudaMemcpyAsync(P.hot_cols.p, ids_sorted.p, 4 * (size_t)K, cudaMemcpyDeviceToDevice, st));
        B200_CUDA(cudaMemsetAsync(slot_of.p, 0xff, 4 * (size_t)n, st));
        slot_scatter_kernel<<<(K + 255) / 256, 256, 0, st>>>(K, P.hot_cols.as<int>(), slot_of.as<int>());
        B200_LAUNCHED();
        remap_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, A.col_idx.as<int>(), slot_of.as<int>(), P.col_hot.as<int>());
        B200_LAUNCHED();
        B200_CUDA(cudaStreamSynchronize(st));
        if(const char *e = getenv("AOCLSPARSE_B200_HOT_MODE"))
            P.hot_mode = atoi(e) == 2 ? 2 : 1;
        P.hot_entries = K;
        P.hot_mass    = (double)mass / (double)nnz;
        return aoclsparse_status_success;
    }

    template <typename T>
    aoclsparse_status launch_hot(const dev_csr &A, const T *x, T *y, T alpha, T beta, cudaStream_t st)
    {
        const row_block_plan &P    = A.plan;
        const int             cap  = P.block_nnz + 8;
        const size_t          smem = hot_smem_bytes(sizeof(T), P.block_nnz, P.hot_entries, P.hot_stages);
        static std::atomic<size_t> configured{0};
        if(configured.load() < smem)
        {
            B200_CUDA(cudaFuncSetAttribute(spmv_hot_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            B200_CUDA(cudaFuncSetAttribute(spmv_hot_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured.store(smem);
        }
        int grid = P.hot_stages == 1 ? 296 : 148;
        if(grid * HOT_GROUPS > P.n_blocks)
            grid = (P.n_blocks + HOT_GROUPS - 1) / HOT_GROUPS;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim            = dim3((unsigned)grid);
        cfg.blockDim           = dim3(HOT_THREADS);
        cfg.dynamicSmemBytes   = smem;
        cfg.stream             = st;
        cudaLaunchAttribute attr[1];
        attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs                                          = attr;
        cfg.numAttrs                                       = P.pdl ? 1 : 0;
        B200_CUDA(cudaLaunchKernelEx(&cfg,
                                     P.hot_stages == 1 ? spmv_hot_kernel<T, 1> : spmv_hot_kernel<T, 2>,
                                     (const int4 *)P.desc.as<int4>(),
                                     (const int *)P.kind.as<int>(),
                                     (int)P.n_blocks,
                                     cap,
                                     (const aoclsparse_int *)A.row_ptr.as<aoclsparse_int>(),
                                     (const aoclsparse_int *)P.col_hot.as<aoclsparse_int>(),
                                     (const T *)A.val.as<T>(),
                                     x,
                                     y,
                                     alpha,
                                     beta,
                                     is_zero(beta) ? 1 : 0,
                                     P.partials.as<T>(),
                                     (const aoclsparse_int *)P.hot_cols.as<aoclsparse_int>(),
                                     (int)P.hot_entries));
        B200_LAUNCHED();
        return aoclsparse_status_success;
    }

    template aoclsparse_status launch_hot<float>(const dev_csr &, const float *, float *, float, float, cudaStream_t);
    template aoclsparse_status launch_hot<double>(const dev_csr &, const double *, double *, double, double, cudaStream_t);
    template aoclsparse_status launch_hot<float2>(const dev_csr &, const float2 *, float2 *, float2, float2, cudaStream_t);
    template aoclsparse_status launch_hot<double2>(const dev_csr &, const double2 *, double2 *, double2, double2, cudaStream_t);
}
