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

// ------------------------------------------------------------------------------------------------
// aoclsparse_?csr2csc (aoclsparse_convert.h:430-530; aoclsparse_csr2csc_template, library/src/conversion/
// aoclsparse_convert.hpp:553-660): array-level CSR -> CSC conversion.  Input and output arrays may be host or device
// memory; the work is the device transposition above (stable: row indices ascend inside every column and repeated
// entries keep their order, as the reference's counting sort does).
// ------------------------------------------------------------------------------------------------
namespace b200
{
    namespace
    {
        __global__ void add_const_kernel(long long n, int *v, int c)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < n; i += (long long)gridDim.x * blockDim.x)
                v[i] += c;
        }

        template <typename T>
        aoclsparse_status csr2csc_t(aoclsparse_int             m,
                                    aoclsparse_int             n,
                                    aoclsparse_int             nnz,
                                    const aoclsparse_mat_descr descr,
                                    aoclsparse_index_base      baseCSC,
                                    const aoclsparse_int      *csr_row_ptr,
                                    const aoclsparse_int      *csr_col_ind,
                                    const T                   *csr_val,
                                    aoclsparse_int            *csc_row_ind,
                                    aoclsparse_int            *csc_col_ptr,
                                    T                         *csc_val,
                                    int                        val_type)
        {
            if(descr == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(m < 0 || n < 0 || nnz < 0)
                return aoclsparse_status_invalid_size;
            cudaStream_t st = current_stream();
            if(m == 0 || n == 0 || nnz == 0)
            {
                if(csc_col_ptr == nullptr)
                    return aoclsparse_status_invalid_pointer; // (the reference writes through it unchecked)
                std::vector<aoclsparse_int> fill((size_t)n + 1, (aoclsparse_int)baseCSC);
                B200_CUDA(cudaMemcpyAsync(csc_col_ptr, fill.data(), sizeof(aoclsparse_int) * ((size_t)n + 1), cudaMemcpyDefault, st));
                B200_CUDA(cudaStreamSynchronize(st));
                return aoclsparse_status_success;
            }
            const aoclsparse_index_base baseCSR = descr->base;
            if((baseCSR != aoclsparse_index_base_zero && baseCSR != aoclsparse_index_base_one)
               || (baseCSC != aoclsparse_index_base_zero && baseCSC != aoclsparse_index_base_one))
                return aoclsparse_status_invalid_value;
            if(!csr_val || !csr_row_ptr || !csr_col_ind || !csc_val || !csc_row_ind || !csc_col_ptr)
                return aoclsparse_status_invalid_pointer;
            dev_csr A, Tr;
            A.m   = m;
            A.n   = n;
            A.nnz = nnz;
            B200_TRY(A.row_ptr.alloc(sizeof(int) * ((size_t)m + 1)));
            B200_TRY(A.col_idx.alloc(sizeof(int) * (size_t)nnz));
            B200_TRY(A.val.alloc(sizeof(T) * (size_t)nnz));
            B200_CUDA(cudaMemcpyAsync(A.row_ptr.p, csr_row_ptr, sizeof(int) * ((size_t)m + 1), cudaMemcpyDefault, st));
            B200_CUDA(cudaMemcpyAsync(A.col_idx.p, csr_col_ind, sizeof(int) * (size_t)nnz, cudaMemcpyDefault, st));
            B200_CUDA(cudaMemcpyAsync(A.val.p, csr_val, sizeof(T) * (size_t)nnz, cudaMemcpyDefault, st));
            if(baseCSR == aoclsparse_index_base_one)
                B200_TRY(rebase_to_zero(m, nnz, A.row_ptr.as<aoclsparse_int>(), A.col_idx.as<aoclsparse_int>(), st));
            B200_TRY(transpose_csr(A, val_type, false, Tr, st));
            if(baseCSC == aoclsparse_index_base_one)
            {
                add_const_kernel<<<grid_for((long long)n + 1, 256), 256, 0, st>>>((long long)n + 1, Tr.row_ptr.as<int>(), 1);
                B200_LAUNCHED();
                add_const_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, Tr.col_idx.as<int>(), 1);
                B200_LAUNCHED();
            }
            B200_CUDA(cudaMemcpyAsync(csc_col_ptr, Tr.row_ptr.p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDefault, st));
            B200_CUDA(cudaMemcpyAsync(csc_row_ind, Tr.col_idx.p, sizeof(int) * (size_t)nnz, cudaMemcpyDefault, st));
            B200_CUDA(cudaMemcpyAsync(csc_val, Tr.val.p, sizeof(T) * (size_t)nnz, cudaMemcpyDefault, st));
            B200_CUDA(cudaStreamSynchronize(st));
            return aoclsparse_status_success;
        }
    }
}

extern "C" {
aoclsparse_status aoclsparse_scsr2csc(aoclsparse_int m, aoclsparse_int n, aoclsparse_int nnz, const aoclsparse_mat_descr descr, aoclsparse_index_base baseCSC, const aoclsparse_int *csr_row_ptr, const aoclsparse_int *csr_col_ind, const float *csr_val, aoclsparse_int *csc_row_ind, aoclsparse_int *csc_col_ptr, float *csc_val)
{
    return b200::csr2csc_t<float>(m, n, nnz, descr, baseCSC, csr_row_ptr, csr_col_ind, csr_val, csc_row_ind, csc_col_ptr, csc_val, aoclsparse_smat);
}
aoclsparse_status aoclsparse_dcsr2csc(aoclsparse_int m, aoclsparse_int n, aoclsparse_int nnz, const aoclsparse_mat_descr descr, aoclsparse_index_base baseCSC, const aoclsparse_int *csr_row_ptr, const aoclsparse_int *csr_col_ind, const double *csr_val, aoclsparse_int *csc_row_ind, aoclsparse_int *csc_col_ptr, double *csc_val)
{
    return b200::csr2csc_t<double>(m, n, nnz, descr, baseCSC, csr_row_ptr, csr_col_ind, csr_val, csc_row_ind, csc_col_ptr, csc_val, aoclsparse_dmat);
}
aoclsparse_status aoclsparse_ccsr2csc(aoclsparse_int m, aoclsparse_int n, aoclsparse_int nnz, const aoclsparse_mat_descr descr, aoclsparse_index_base baseCSC, const aoclsparse_int *csr_row_ptr, const aoclsparse_int *csr_col_ind, const aoclsparse_float_complex *csr_val, aoclsparse_int *csc_row_ind, aoclsparse_int *csc_col_ptr, aoclsparse_float_complex *csc_val)
{
    return b200::csr2csc_t<float2>(m, n, nnz, descr, baseCSC, csr_row_ptr, csr_col_ind, reinterpret_cast<const float2 *>(csr_val), csc_row_ind, csc_col_ptr, reinterpret_cast<float2 *>(csc_val), aoclsparse_cmat);
}
aoclsparse_status aoclsparse_zcsr2csc(aoclsparse_int m, aoclsparse_int n, aoclsparse_int nnz, const aoclsparse_mat_descr descr, aoclsparse_index_base baseCSC, const aoclsparse_int *csr_row_ptr, const aoclsparse_int *csr_col_ind, const aoclsparse_double_complex *csr_val, aoclsparse_int *csc_row_ind, aoclsparse_int *csc_col_ptr, aoclsparse_double_complex *csc_val)
{
    return b200::csr2csc_t<double2>(m, n, nnz, descr, baseCSC, csr_row_ptr, csr_col_ind, reinterpret_cast<const double2 *>(csr_val), csc_row_ind, csc_col_ptr, reinterpret_cast<double2 *>(csc_val), aoclsparse_zmat);
}
}
