// transpose.cu -- device CSR -> transposed CSR (optionally conjugated), used by aoclsparse_optimize
// when a transposed mv / mm hint was given and the memory policy allows an extra copy.
//
// The reference materialises the same kind of copy on the host in aoclsparse_matrix_transform
// (library/src/analysis/aoclsparse_csr_util.hpp:512-756, via csr2csc, library/src/conversion/
// aoclsparse_convert.cpp:831-946).  Here: a stable radix sort of the entry indices by column (so the
// rows of every output row stay ascending), then two gathers.  Analysis-time only, never on the
// multiply path; the sort is CUB's (shipped with the CUDA toolkit).
#include "common.hpp"

#include <cub/device/device_radix_sort.cuh>

namespace b200
{
    namespace
    {
        __global__ void iota_kernel(long long n, int *out)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < n; i += (long long)gridDim.x * blockDim.x)
                out[i] = (int)i;
        }

        // row_of[p] = row owning entry p (one warp per row writes its span)
        __global__ void expand_rows_kernel(int m, const int *__restrict__ rp, int *__restrict__ row_of)
        {
            const int lane = threadIdx.x & 31;
            long long w    = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
            const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
            for(; w < m; w += nw)
                for(int p = rp[w] + lane; p < rp[w + 1]; p += 32)
                    row_of[p] = (int)w;
        }

        // new_rp[c] = number of sorted keys < c  (lower bound), c in [0, n]
        __global__ void bounds_kernel(int n, long long nnz, const int *__restrict__ keys, int *__restrict__ new_rp)
        {
            long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; c <= n; c += (long long)gridDim.x * blockDim.x)
            {
                long long lo = 0, hi = nnz;
                while(lo < hi)
                {
                    long long mid = lo + (hi - lo) / 2;
                    if(keys[mid] < (int)c)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                new_rp[c] = (int)lo;
            }
        }

        template <typename T>
        __global__ void permute_kernel(long long nnz,
                                       const int *__restrict__ perm,
                                       const int *__restrict__ row_of,
                                       const T *__restrict__ val,
                                       int conj,
                                       int *__restrict__ out_col,
                                       T *__restrict__ out_val)
        {
            long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; q < nnz; q += (long long)gridDim.x * blockDim.x)
            {
                const int p = perm[q];
                out_col[q]  = row_of[p];
                T v         = val[p];
                out_val[q]  = conj ? cj(v) : v;
            }
        }

        inline unsigned grid_for(long long n, int tpb)
        {
            long long b = (n + tpb - 1) / tpb;
            if(b > 148LL * 32)
                b = 148LL * 32;
            return (unsigned)(b < 1 ? 1 : b);
        }
    }

    aoclsparse_status transpose_csr(const dev_csr &A, int val_type, bool conj, dev_csr &out, cudaStream_t st)
    {
        const long long nnz = A.nnz;
        const size_t    es  = value_size(val_type);
        out.m               = A.n;
        out.n               = A.m;
        out.nnz             = A.nnz;
        B200_TRY(out.row_ptr.alloc(sizeof(int) * ((size_t)A.n + 1)));
        B200_TRY(out.col_idx.alloc(sizeof(int) * (size_t)nnz));
        B200_TRY(out.val.alloc(es * (size_t)nnz));
        if(nnz == 0)
        {
            B200_CUDA(cudaMemsetAsync(out.row_ptr.p, 0, sizeof(int) * ((size_t)A.n + 1), st));
            return aoclsparse_status_success;
        }
        dev_buf row_of, idx_in, idx_out, keys_out, temp;
        B200_TRY(row_of.alloc(sizeof(int) * (size_t)nnz));
        B200_TRY(idx_in.alloc(sizeof(int) * (size_t)nnz));
        B200_TRY(idx_out.alloc(sizeof(int) * (size_t)nnz));
        B200_TRY(keys_out.alloc(sizeof(int) * (size_t)nnz));
        iota_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, idx_in.as<int>());
        B200_LAUNCHED();
        expand_rows_kernel<<<grid_for((long long)A.m * 32, 256), 256, 0, st>>>(A.m, A.row_ptr.as<int>(), row_of.as<int>());
        B200_LAUNCHED();

        int end_bit = 1;
        while(end_bit < 31 && (1LL << end_bit) < (long long)A.n)
            ++end_bit;
        size_t temp_bytes = 0;
        B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr,
                                                  temp_bytes,
                                                  A.col_idx.as<int>(),
                                                  keys_out.as<int>(),
                                                  idx_in.as<int>(),
                                                  idx_out.as<int>(),
                                                  (int)nnz,
                                                  0,
                                                  end_bit,
                                                  st));
        B200_TRY(temp.alloc(temp_bytes));
        B200_CUDA(cub::DeviceRadixSort::SortPairs(temp.p,
                                                  temp_bytes,
                                                  A.col_idx.as<int>(),
                                                  keys_out.as<int>(),
                                                  idx_in.as<int>(),
                                                  idx_out.as<int>(),
                                                  (int)nnz,
                                                  0,
                                                  end_bit,
                                                  st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        bounds_kernel<<<grid_for((long long)A.n + 1, 256), 256, 0, st>>>(A.n, nnz, keys_out.as<int>(), out.row_ptr.as<int>());
        B200_LAUNCHED();
        const unsigned g = grid_for(nnz, 256);
        switch(val_type)
        {
        case aoclsparse_dmat:
            permute_kernel<double><<<g, 256, 0, st>>>(
                nnz, idx_out.as<int>(), row_of.as<int>(), A.val.as<double>(), 0, out.col_idx.as<int>(), out.val.as<double>());
            break;
        case aoclsparse_smat:
            permute_kernel<float><<<g, 256, 0, st>>>(
                nnz, idx_out.as<int>(), row_of.as<int>(), A.val.as<float>(), 0, out.col_idx.as<int>(), out.val.as<float>());
            break;
        case aoclsparse_cmat:
            permute_kernel<float2><<<g, 256, 0, st>>>(
                nnz, idx_out.as<int>(), row_of.as<int>(), A.val.as<float2>(), conj ? 1 : 0, out.col_idx.as<int>(), out.val.as<float2>());
            break;
        default:
            permute_kernel<double2><<<g, 256, 0, st>>>(
                nnz, idx_out.as<int>(), row_of.as<int>(), A.val.as<double2>(), conj ? 1 : 0, out.col_idx.as<int>(), out.val.as<double2>());
            break;
        }
        B200_LAUNCHED();
        B200_CUDA(cudaStreamSynchronize(st)); // temporaries are freed on return
        return aoclsparse_status_success;
    }
}
