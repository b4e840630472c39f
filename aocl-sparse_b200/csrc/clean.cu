// clean.cu -- device restatement of the reference's "clean CSR" analysis product:
//   aoclsparse_csr_csc_optimize<T>       library/src/analysis/aoclsparse_csr_util.hpp:765-967
//   aoclsparse_csr_csc_check_sort_diag   library/src/analysis/aoclsparse_csr_util.cpp:290-364
//   aoclsparse_sort_idx_val              library/src/analysis/aoclsparse_csr_util.hpp:99-160
//   aoclsparse_csr_csc_fill_diag         library/src/analysis/aoclsparse_csr_util.hpp:166-279
//   aoclsparse_csr_csc_indices           library/src/analysis/aoclsparse_csr_util.cpp:389-458
// i.e. a CSR whose rows are grouped lower | diagonal | upper with every diagonal entry of rows i < n present
// (explicit zeros inserted where missing), plus idiag[i] / iurow[i] = position of the diagonal / of the first
// strictly-upper entry of row i.  This is integer metadata the reference's tests pin bit-exactly
// (tests/unit_tests/hint_tests.cpp:72-140); the multiply kernels here do not need it (they mask by comparing
// col with row), it is produced for callers / tests through aoclsparse_b200_get_clean_csr.
//
// Decision tree as in the reference: rows already group-ordered and full diagonal -> the input is the clean
// matrix (the caller's base is kept); otherwise a base-0 copy is made, rows are sorted by column if they were
// not group-ordered, and missing diagonals are inserted in front of the first upper entry.
#include "common.hpp"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

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

        __global__ void row_keys_kernel(int m, const int *__restrict__ rp, const int *__restrict__ col, unsigned long long *keys, int *idx)
        {
            const int lane = threadIdx.x & 31;
            long long w    = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
            const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
            for(; w < m; w += nw)
                for(int p = rp[w] + lane; p < rp[w + 1]; p += 32)
                {
                    keys[p] = ((unsigned long long)w << 32) | (unsigned)col[p];
                    idx[p]  = p;
                }
        }

        template <typename T>
        __global__ void apply_perm_kernel(long long nnz, const unsigned long long *__restrict__ keys, const int *__restrict__ perm, const T *__restrict__ val, int *col_out, T *val_out)
        {
            long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; q < nnz; q += (long long)gridDim.x * blockDim.x)
            {
                col_out[q] = (int)(keys[q] & 0xffffffffull);
                val_out[q] = val[perm[q]];
            }
        }

        // missing[i] = 1 if row i < n has no diagonal entry; pos[i] = first entry with col >= i (or row end)
        __global__ void diag_scan_kernel(int m, int n, const int *__restrict__ rp, const int *__restrict__ col, int *missing, int *pos)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < m; i += (long long)gridDim.x * blockDim.x)
            {
                int p = rp[i];
                const int e = rp[i + 1];
                while(p < e && col[p] < (int)i)
                    ++p;
                pos[i]     = p;
                missing[i] = ((p == e || col[p] != (int)i) && i < n) ? 1 : 0;
            }
            if(blockIdx.x == 0 && threadIdx.x == 0)
                missing[m] = 0;
        }

        template <typename T>
        __global__ void fill_diag_kernel(int m,
                                         const int *__restrict__ rp,
                                         const int *__restrict__ col,
                                         const T *__restrict__ val,
                                         const int *__restrict__ missing,
                                         const int *__restrict__ shift, // exclusive scan of missing
                                         const int *__restrict__ pos,
                                         int *rp_out,
                                         int *col_out,
                                         T   *val_out)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < m; i += (long long)gridDim.x * blockDim.x)
            {
                const int s = rp[i], e = rp[i + 1], sh = shift[i];
                rp_out[i] = s + sh;
                int q     = s + sh;
                for(int p = s; p < e; ++p)
                {
                    if(missing[i] && p == pos[i])
                    {
                        col_out[q] = (int)i;
                        val_out[q] = vt<T>::zero();
                        ++q;
                    }
                    col_out[q] = col[p];
                    val_out[q] = val[p];
                    ++q;
                }
                if(missing[i] && pos[i] == e)
                {
                    col_out[q] = (int)i;
                    val_out[q] = vt<T>::zero();
                }
                if(i == m - 1)
                    rp_out[m] = e + sh + missing[i];
            }
        }

        __global__ void diag_index_kernel(int m, int base_out, const int *__restrict__ rp, const int *__restrict__ col, int *idiag, int *iurow)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < m; i += (long long)gridDim.x * blockDim.x)
            {
                int       p = rp[i];
                const int e = rp[i + 1];
                while(p < e && col[p] < (int)i)
                    ++p;
                idiag[i] = p + base_out;
                iurow[i] = ((p < e && col[p] == (int)i) ? p + 1 : p) + base_out;
            }
        }

        template <typename T>
        aoclsparse_status build_clean(aoclsparse_matrix A, clean_csr &out, cudaStream_t st)
        {
            const dev_csr &M   = *A->mats[0];
            const int      m   = M.m;
            const bool grouped = A->sort != aoclsparse_unsorted;
            out                = clean_csr();
            B200_TRY(out.idiag.alloc(sizeof(int) * (size_t)(m > 0 ? m : 1)));
            B200_TRY(out.iurow.alloc(sizeof(int) * (size_t)(m > 0 ? m : 1)));
            if(grouped && A->fulldiag)
            {
                // the input already is the clean matrix; positions are reported in the caller's base
                out.is_internal = false;
                out.nnz         = M.nnz;
                if(m > 0)
                {
                    diag_index_kernel<<<grid_for(m, 128), 128, 0, st>>>(
                        m, (int)A->base, M.row_ptr.as<int>(), M.col_idx.as<int>(), out.idiag.as<int>(), out.iurow.as<int>());
                    B200_LAUNCHED();
                }
                out.valid = true;
                return aoclsparse_status_success;
            }
            out.is_internal = true;
            const long long nnz = M.nnz;
            // 1. rows sorted by column when the input was not group-ordered
            dev_buf s_col, s_val;
            const int *cur_col = M.col_idx.as<int>();
            const T   *cur_val = M.val.as<T>();
            if(!grouped && nnz > 0)
            {
                dev_buf keys_in, keys_out, idx_in, idx_out, temp;
                B200_TRY(keys_in.alloc(8 * (size_t)nnz));
                B200_TRY(keys_out.alloc(8 * (size_t)nnz));
                B200_TRY(idx_in.alloc(4 * (size_t)nnz));
                B200_TRY(idx_out.alloc(4 * (size_t)nnz));
                row_keys_kernel<<<grid_for((long long)m * 32, 256), 256, 0, st>>>(
                    m, M.row_ptr.as<int>(), M.col_idx.as<int>(), keys_in.as<unsigned long long>(), idx_in.as<int>());
                B200_LAUNCHED();
                size_t tb = 0;
                B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr,
                                                          tb,
                                                          keys_in.as<unsigned long long>(),
                                                          keys_out.as<unsigned long long>(),
                                                          idx_in.as<int>(),
                                                          idx_out.as<int>(),
                                                          (int)nnz,
                                                          0,
                                                          64,
                                                          st));
                B200_TRY(temp.alloc(tb));
                B200_CUDA(cub::DeviceRadixSort::SortPairs(temp.p,
                                                          tb,
                                                          keys_in.as<unsigned long long>(),
                                                          keys_out.as<unsigned long long>(),
                                                          idx_in.as<int>(),
                                                          idx_out.as<int>(),
                                                          (int)nnz,
                                                          0,
                                                          64,
                                                          st));
                g_launches.fetch_add(1, std::memory_order_relaxed);
                B200_TRY(s_col.alloc(4 * (size_t)nnz));
                B200_TRY(s_val.alloc(sizeof(T) * (size_t)nnz));
                apply_perm_kernel<T><<<grid_for(nnz, 256), 256, 0, st>>>(
                    nnz, keys_out.as<unsigned long long>(), idx_out.as<int>(), M.val.as<T>(), s_col.as<int>(), s_val.as<T>());
                B200_LAUNCHED();
                B200_CUDA(cudaStreamSynchronize(st));
                cur_col = s_col.as<int>();
                cur_val = s_val.as<T>();
            }
            // 2. missing diagonals
            dev_buf missing, shift, pos;
            B200_TRY(missing.alloc(4 * ((size_t)m + 1)));
            B200_TRY(shift.alloc(4 * ((size_t)m + 1)));
            B200_TRY(pos.alloc(4 * ((size_t)m + 1)));
            int n_missing = 0;
            if(m > 0)
            {
                diag_scan_kernel<<<grid_for(m, 128), 128, 0, st>>>(
                    m, M.n, M.row_ptr.as<int>(), cur_col, missing.as<int>(), pos.as<int>());
                B200_LAUNCHED();
                size_t tb = 0;
                B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, missing.as<int>(), shift.as<int>(), m + 1, st));
                dev_buf t;
                B200_TRY(t.alloc(tb));
                B200_CUDA(cub::DeviceScan::ExclusiveSum(t.p, tb, missing.as<int>(), shift.as<int>(), m + 1, st));
                g_launches.fetch_add(1, std::memory_order_relaxed);
                B200_CUDA(cudaMemcpyAsync(&n_missing, shift.as<int>() + m, 4, cudaMemcpyDeviceToHost, st));
                B200_CUDA(cudaStreamSynchronize(st));
            }
            out.nnz = (aoclsparse_int)(nnz + n_missing);
            B200_TRY(out.row_ptr.alloc(4 * ((size_t)m + 1)));
            B200_TRY(out.col_idx.alloc(4 * (size_t)out.nnz));
            B200_TRY(out.val.alloc(sizeof(T) * (size_t)out.nnz));
            if(m > 0)
            {
                fill_diag_kernel<T><<<grid_for(m, 128), 128, 0, st>>>(m,
                                                                     M.row_ptr.as<int>(),
                                                                     cur_col,
                                                                     cur_val,
                                                                     missing.as<int>(),
                                                                     shift.as<int>(),
                                                                     pos.as<int>(),
                                                                     out.row_ptr.as<int>(),
                                                                     out.col_idx.as<int>(),
                                                                     out.val.as<T>());
                B200_LAUNCHED();
                diag_index_kernel<<<grid_for(m, 128), 128, 0, st>>>(
                    m, 0, out.row_ptr.as<int>(), out.col_idx.as<int>(), out.idiag.as<int>(), out.iurow.as<int>());
                B200_LAUNCHED();
            }
            else
                B200_CUDA(cudaMemsetAsync(out.row_ptr.p, 0, 4, st));
            B200_CUDA(cudaStreamSynchronize(st));
            out.valid = true;
            return aoclsparse_status_success;
        }
    }

    aoclsparse_status ensure_clean(aoclsparse_matrix A, cudaStream_t st)
    {
        std::unique_lock<std::shared_mutex> wl(A->guard);
        if(A->clean.valid)
            return aoclsparse_status_success;
        switch(A->val_type)
        {
        case aoclsparse_dmat:
            return build_clean<double>(A, A->clean, st);
        case aoclsparse_smat:
            return build_clean<float>(A, A->clean, st);
        case aoclsparse_cmat:
            return build_clean<float2>(A, A->clean, st);
        default:
            return build_clean<double2>(A, A->clean, st);
        }
    }
}

using namespace b200;

extern "C" aoclsparse_status aoclsparse_b200_get_clean_csr(aoclsparse_matrix A,
                                                           aoclsparse_int   *nnz,
                                                           int              *is_internal,
                                                           aoclsparse_int   *row_ptr,
                                                           aoclsparse_int   *col_idx,
                                                           void             *val,
                                                           aoclsparse_int   *idiag,
                                                           aoclsparse_int   *iurow)
{
    if(!A || !nnz)
        return aoclsparse_status_invalid_pointer;
    if(A->mats.empty() || !A->mats[0])
        return aoclsparse_status_invalid_pointer;
    cudaStream_t st = current_stream();
    B200_TRY(ensure_clean(A, st));
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const clean_csr                    &Cn = A->clean;
    const dev_csr                      &M  = *A->mats[0];
    *nnz                                   = Cn.nnz;
    if(is_internal)
        *is_internal = Cn.is_internal ? 1 : 0;
    const size_t es = value_size(A->val_type);
    const int    m  = A->m;
    // when the input itself is the clean matrix the caller's own (based) arrays are what the reference exposes
    const void *d_rp  = Cn.is_internal ? Cn.row_ptr.p : M.row_ptr.p;
    const void *d_col = Cn.is_internal ? Cn.col_idx.p : M.col_idx.p;
    const void *d_val = Cn.is_internal ? Cn.val.p : M.val.p;
    const int   shift = Cn.is_internal ? 0 : (int)A->base;
    if(row_ptr)
    {
        B200_CUDA(cudaMemcpyAsync(row_ptr, d_rp, 4 * ((size_t)m + 1), cudaMemcpyDeviceToHost, st));
    }
    if(col_idx && Cn.nnz > 0)
        B200_CUDA(cudaMemcpyAsync(col_idx, d_col, 4 * (size_t)Cn.nnz, cudaMemcpyDeviceToHost, st));
    if(val && Cn.nnz > 0)
        B200_CUDA(cudaMemcpyAsync(val, d_val, es * (size_t)Cn.nnz, cudaMemcpyDeviceToHost, st));
    if(idiag && m > 0)
        B200_CUDA(cudaMemcpyAsync(idiag, Cn.idiag.p, 4 * (size_t)m, cudaMemcpyDeviceToHost, st));
    if(iurow && m > 0)
        B200_CUDA(cudaMemcpyAsync(iurow, Cn.iurow.p, 4 * (size_t)m, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if(shift)
    {
        if(row_ptr)
            for(int i = 0; i <= m; ++i)
                row_ptr[i] += shift;
        if(col_idx)
            for(aoclsparse_int i = 0; i < Cn.nnz; ++i)
                col_idx[i] += shift;
    }
    return aoclsparse_status_success;
}
