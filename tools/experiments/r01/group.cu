// group.cu -- "row-grouped" copy of a sorted CSR matrix for the row-major csrmm kernel (csrmm.cu).
//
// Why: C = A B with a row-major B of n columns reads one B row (n * sizeof(T) bytes) per stored entry of A.  On the
// 27-point stencil x 32 doubles that is 14.3 GB of L1/L2 traffic for 1.75 GB of HBM traffic, and the kernel is bound by
// L1 delivery, not by HBM (profiles/r01_summary.md, C4).  Neighbouring rows of A mostly name the same columns, so K
// consecutive rows are taken together: the union of their column indices is walked once, each B row is loaded once into
// registers and multiplied into up to K accumulators.  Stored per group entry: the column index with a K-bit row mask in
// bits 27.. (which of the K rows really hold that column -- absent entries are NOT multiplied, so Inf / NaN in B
// propagate exactly as in the row-by-row product) and K values (zero where absent).
//
// The reference has no such format; its csrmm kernels (library/src/level3/aoclsparse_csrmm_kt.cpp:31-363) walk plain
// CSR.  The closest thing is the analysis-time copy making of aoclsparse_optimize (library/src/analysis/
// aoclsparse_analysis.cpp:426-566): like those copies this one is built when a mm hint was given, lives in the handle,
// and is dropped when values change.
//
// Built only when it pays: rows fully sorted (the union is a K-way merge), n < 2^27 (room for the mask), no group longer
// than a row block, and the union at most ~70 % of the entries it replaces.
#include "common.hpp"

#include <cub/device/device_scan.cuh>

namespace b200
{
    namespace
    {
        constexpr int COL_BITS = 27;

        // one thread per group: K-way merge of the sorted rows; FILL = false counts the union, true writes it
        template <typename T, int K, bool FILL>
        __global__ void group_walk_kernel(int m,
                                          int n_groups,
                                          const int *__restrict__ rp,
                                          const int *__restrict__ col,
                                          const T *__restrict__ val,
                                          int *__restrict__ counts,
                                          const int *__restrict__ gptr,
                                          int *__restrict__ gcol,
                                          T *__restrict__ gval)
        {
            long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; g < n_groups; g += (long long)gridDim.x * blockDim.x)
            {
                int p[K], e[K];
#pragma unroll
                for(int i = 0; i < K; ++i)
                {
                    const long long r = g * K + i;
                    p[i]              = r < m ? rp[r] : 0;
                    e[i]              = r < m ? rp[r + 1] : 0;
                }
                int out = FILL ? gptr[g] : 0;
                while(true)
                {
                    int cmin = INT_MAX;
#pragma unroll
                    for(int i = 0; i < K; ++i)
                        if(p[i] < e[i])
                            cmin = min(cmin, col[p[i]]);
                    if(cmin == INT_MAX)
                        break;
                    int mask = 0;
#pragma unroll
                    for(int i = 0; i < K; ++i)
                    {
                        // a row consumes ONE entry per step, so a repeated column index in a row simply opens another
                        // group entry and the products of that row keep their storage order
                        const bool hit = p[i] < e[i] && col[p[i]] == cmin;
                        if(FILL)
                            gval[(size_t)out * K + i] = hit ? val[p[i]] : vt<T>::zero();
                        if(hit)
                        {
                            mask |= 1 << i;
                            ++p[i];
                        }
                    }
                    if(FILL)
                        gcol[out] = cmin | (mask << COL_BITS);
                    ++out;
                }
                if(!FILL)
                    counts[g] = out;
            }
        }

        template <typename T, int K>
        aoclsparse_status build_grouped_t(const dev_csr &A, dev_csr &G, cudaStream_t st)
        {
            const int n_groups = (int)(((long long)A.m + K - 1) / K);
            G.m                = n_groups;
            G.n                = A.n;
            dev_buf counts, temp;
            B200_TRY(counts.alloc(sizeof(int) * ((size_t)n_groups + 1)));
            B200_TRY(G.row_ptr.alloc(sizeof(int) * ((size_t)n_groups + 1)));
            B200_CUDA(cudaMemsetAsync(counts.p, 0, sizeof(int) * ((size_t)n_groups + 1), st));
            long long blocks = ((long long)n_groups + 127) / 128;
            if(blocks > 148 * 64)
                blocks = 148 * 64;
            group_walk_kernel<T, K, false><<<(unsigned)blocks, 128, 0, st>>>(A.m,
                                                                            n_groups,
                                                                            A.row_ptr.as<int>(),
                                                                            A.col_idx.as<int>(),
                                                                            A.val.as<T>(),
                                                                            counts.as<int>(),
                                                                            nullptr,
                                                                            nullptr,
                                                                            nullptr);
            B200_LAUNCHED();
            size_t temp_bytes = 0;
            B200_CUDA(cub::DeviceScan::ExclusiveSum(
                nullptr, temp_bytes, counts.as<int>(), G.row_ptr.as<int>(), n_groups + 1, st));
            B200_TRY(temp.alloc(temp_bytes));
            B200_CUDA(cub::DeviceScan::ExclusiveSum(
                temp.p, temp_bytes, counts.as<int>(), G.row_ptr.as<int>(), n_groups + 1, st));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            int total = 0;
            B200_CUDA(cudaMemcpyAsync(&total, G.row_ptr.as<int>() + n_groups, sizeof(int), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            G.nnz = total;
            // worth it only when clearly fewer B rows are loaded (and the A stream does not grow out of proportion)
            if((double)total > 0.7 * (double)A.nnz)
                return aoclsparse_status_success; // G.val stays empty: "not beneficial"
            B200_TRY(G.col_idx.alloc(sizeof(int) * (size_t)total));
            B200_TRY(G.val.alloc(sizeof(T) * (size_t)total * K));
            group_walk_kernel<T, K, true><<<(unsigned)blocks, 128, 0, st>>>(A.m,
                                                                           n_groups,
                                                                           A.row_ptr.as<int>(),
                                                                           A.col_idx.as<int>(),
                                                                           A.val.as<T>(),
                                                                           nullptr,
                                                                           G.row_ptr.as<int>(),
                                                                           G.col_idx.as<int>(),
                                                                           G.val.as<T>());
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }

        template <typename T>
        aoclsparse_status build_grouped_k(const dev_csr &A, int k, dev_csr &G, cudaStream_t st)
        {
            return k == 2 ? build_grouped_t<T, 2>(A, G, st) : build_grouped_t<T, 4>(A, G, st);
        }
    }

    // Attaches the grouped copy to A->mats[0] when it is eligible and beneficial; a no-op otherwise (remembered in
    // group_k = -1 so the analysis is not repeated).  Caller holds the handle's write lock.
    aoclsparse_status ensure_grouped(aoclsparse_matrix A, cudaStream_t st)
    {
        dev_csr &M = *A->mats[0];
        if(M.group_k != 0)
            return aoclsparse_status_success;
        // OFF by default: measured slower than the plain kernel on the 27-point stencil x 32 doubles (1.00 ms K=2,
        // 1.08 ms K=4 against 0.81 ms, profiles/r01_sweep_group.txt) -- the zero-padded groups are only 50-75 % full and
        // the masked multiply-adds cost more issue slots than the saved B-row loads.  AOCLSPARSE_B200_MM_GROUP=2|4
        // enables it for experiments and for tests/test_parity_gpu.py::test_csrmm_row_grouped_copy.
        const int        env_k = getenv("AOCLSPARSE_B200_MM_GROUP") ? atoi(getenv("AOCLSPARSE_B200_MM_GROUP")) : 0;
        const int        k     = env_k == 2 ? 2 : (env_k == 4 ? 4 : 0);
        M.group_k              = -1;
        if(k == 0 || A->sort != aoclsparse_fully_sorted || A->mem_policy != aoclsparse_memory_usage_unrestricted
           || M.n >= (1 << COL_BITS) || M.m < 4096 || (long long)M.nnz < 2LL * M.m || A->win_hi >= 0)
            return aoclsparse_status_success;
        std::unique_ptr<dev_csr> G(new(std::nothrow) dev_csr);
        if(!G)
            return aoclsparse_status_memory_error;
        aoclsparse_status s;
        switch(A->val_type)
        {
        case aoclsparse_smat:
            s = build_grouped_k<float>(M, k, *G, st);
            break;
        case aoclsparse_dmat:
            s = build_grouped_k<double>(M, k, *G, st);
            break;
        case aoclsparse_cmat:
            s = build_grouped_k<float2>(M, k, *G, st);
            break;
        default:
            s = build_grouped_k<double2>(M, k, *G, st);
            break;
        }
        if(s != aoclsparse_status_success)
            return s == aoclsparse_status_memory_error ? aoclsparse_status_success : s; // no room: stay on plain CSR
        if(G->val.p == nullptr)
            return aoclsparse_status_success;
        // row blocks over the groups: about one group per sub-warp slot of a CTA (32 slots), capped by the staging
        // budget of ~56 KB per CTA (4 CTAs per SM)
        const size_t   entry_bytes = value_size(A->val_type) * k + sizeof(int);
        const double   mean        = (double)G->nnz / (double)G->m;
        aoclsparse_int T           = (aoclsparse_int)(mean * 29.0);
        const int env_t = getenv("AOCLSPARSE_B200_MM_GROUP_NNZ") ? atoi(getenv("AOCLSPARSE_B200_MM_GROUP_NNZ")) : 0;
        if(env_t > 0)
            T = env_t;
        const aoclsparse_int t_max = (aoclsparse_int)((56 * 1024) / entry_bytes);
        T                          = T > t_max ? t_max : T;
        T                          = (T / 64) * 64;
        if(T < 256)
            T = 256;
        s = build_plan(*G, entry_bytes - sizeof(int), -1, -1, std::vector<aoclsparse_int>(), st, T);
        if(s != aoclsparse_status_success)
            return s;
        if(G->plan.n_long_rows > 0)
            return aoclsparse_status_success; // a group does not fit one row block: not worth a second code path
        M.group_k = k;
        M.grouped = std::move(G);
        return aoclsparse_status_success;
    }
}
