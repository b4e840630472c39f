// expand.cu -- builds, on the device, the full general CSR that a symmetric / hermitian descriptor
// stands for, so that sparse x dense products (and repeated mv calls after aoclsparse_optimize) run
// on the streaming gather kernels instead of a gather + atomic scatter pair.
//
// Device analogue of the symmetric / hermitian branch of aoclsparse_matrix_transform
// (library/src/analysis/aoclsparse_csr_util.hpp:620-745).  Given the stored triangle T (strict part),
// the diagonal D' selected by the diag type (stored / ones / absent) and the mirror S:
//     symmetric: F = T + D' + T^T ; op H -> conj of everything
//     hermitian: F = T + D' + T^H ; op T -> conj(T) + D' + T^T   (diagonal as stored, like the
//                                     reference kernels, aoclsparse_csrmv_kr.hpp:398-425)
// Row i of F is laid out [lower part | diagonal | upper part]; the mirrored part comes from a stable
// transpose of the strict triangle, so rows of a sorted input stay sorted and the layout is
// deterministic.  Scans and the sort inside transpose_csr are CUB's; analysis-time only.
#include "common.hpp"

#include <cub/device/device_scan.cuh>

namespace b200
{
    namespace
    {
        // per row: number of kept strict-triangle entries and position of the stored diagonal (-1 if none)
        __global__ void tri_count_kernel(int m, int lower, const int *__restrict__ rp, const int *__restrict__ col, int *cnt, int *diag_pos)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < m; i += (long long)gridDim.x * blockDim.x)
            {
                int c = 0, dp = -1;
                for(int p = rp[i]; p < rp[i + 1]; ++p)
                {
                    const int j = col[p];
                    if(j == (int)i)
                        dp = p;
                    else if(lower ? (j < (int)i) : (j > (int)i))
                        ++c;
                }
                cnt[i]      = c;
                diag_pos[i] = dp;
            }
        }

        template <typename T>
        __global__ void tri_fill_kernel(int m,
                                        int lower,
                                        int conj,
                                        const int *__restrict__ rp,
                                        const int *__restrict__ col,
                                        const T *__restrict__ val,
                                        const int *__restrict__ out_rp,
                                        int *out_col,
                                        T   *out_val)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < m; i += (long long)gridDim.x * blockDim.x)
            {
                int q = out_rp[i];
                for(int p = rp[i]; p < rp[i + 1]; ++p)
                {
                    const int j = col[p];
                    if(j != (int)i && (lower ? (j < (int)i) : (j > (int)i)))
                    {
                        out_col[q] = j;
                        out_val[q] = conj ? cj(val[p]) : val[p];
                        ++q;
                    }
                }
            }
        }

        __global__ void full_count_kernel(int m, int diag_mode, const int *__restrict__ ts_rp, const int *__restrict__ tt_rp, const int *__restrict__ diag_pos, int *cnt)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < m; i += (long long)gridDim.x * blockDim.x)
            {
                int d = 0;
                if(diag_mode == 1)
                    d = 1; // unit
                else if(diag_mode == 0 && diag_pos[i] >= 0)
                    d = 1; // stored
                cnt[i] = (ts_rp[i + 1] - ts_rp[i]) + (tt_rp[i + 1] - tt_rp[i]) + d;
            }
        }

        // row i of F: [lower | diag | upper]
        template <typename T>
        __global__ void assemble_kernel(int m,
                                        int lower,      // stored triangle is the lower one
                                        int diag_mode,  // 0 stored, 1 unit, 2 absent
                                        int conj_diag,
                                        int conj_mirror, // mirror values need one more conjugation relative to Ts values
                                        const int *__restrict__ ts_rp,
                                        const int *__restrict__ ts_col,
                                        const T *__restrict__ ts_val,
                                        const int *__restrict__ tt_rp,
                                        const int *__restrict__ tt_col,
                                        const T *__restrict__ tt_val,
                                        const int *__restrict__ diag_pos,
                                        const T *__restrict__ val,
                                        const int *__restrict__ out_rp,
                                        int *out_col,
                                        T   *out_val)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < m; i += (long long)gridDim.x * blockDim.x)
            {
                int q = out_rp[i];
                // first part: entries with col < i
                const int *a_rp = lower ? ts_rp : tt_rp, *b_rp = lower ? tt_rp : ts_rp;
                const int *a_col = lower ? ts_col : tt_col, *b_col = lower ? tt_col : ts_col;
                const T   *a_val = lower ? ts_val : tt_val, *b_val = lower ? tt_val : ts_val;
                const int  a_cj = lower ? 0 : conj_mirror, b_cj = lower ? conj_mirror : 0;
                for(int p = a_rp[i]; p < a_rp[i + 1]; ++p, ++q)
                {
                    out_col[q] = a_col[p];
                    out_val[q] = a_cj ? cj(a_val[p]) : a_val[p];
                }
                if(diag_mode == 1)
                {
                    out_col[q] = (int)i;
                    out_val[q] = vt<T>::one();
                    ++q;
                }
                else if(diag_mode == 0 && diag_pos[i] >= 0)
                {
                    out_col[q] = (int)i;
                    T v        = val[diag_pos[i]];
                    out_val[q] = conj_diag ? cj(v) : v;
                    ++q;
                }
                for(int p = b_rp[i]; p < b_rp[i + 1]; ++p, ++q)
                {
                    out_col[q] = b_col[p];
                    out_val[q] = b_cj ? cj(b_val[p]) : b_val[p];
                }
            }
        }

        inline unsigned grid_for(long long n, int tpb)
        {
            long long b = (n + tpb - 1) / tpb;
            if(b > 148LL * 32)
                b = 148LL * 32;
            return (unsigned)(b < 1 ? 1 : b);
        }

        aoclsparse_status exclusive_scan(const int *in, int *out, int count, cudaStream_t st)
        {
            size_t tb = 0;
            B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, count, st));
            dev_buf t;
            B200_TRY(t.alloc(tb));
            B200_CUDA(cub::DeviceScan::ExclusiveSum(t.p, tb, in, out, count, st));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            B200_CUDA(cudaStreamSynchronize(st));
            return aoclsparse_status_success;
        }

        template <typename T>
        aoclsparse_status build_expanded(const dev_csr &A, int val_type, int lower, int diag_mode, int cg, int cs, int cd, dev_csr &F, cudaStream_t st)
        {
            const int m = A.m;
            dev_buf   cnt, diag_pos;
            B200_TRY(cnt.alloc(sizeof(int) * ((size_t)m + 1)));
            B200_TRY(diag_pos.alloc(sizeof(int) * ((size_t)m + 1)));
            B200_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int) * ((size_t)m + 1), st));
            tri_count_kernel<<<grid_for(m, 128), 128, 0, st>>>(
                m, lower, A.row_ptr.as<int>(), A.col_idx.as<int>(), cnt.as<int>(), diag_pos.as<int>());
            B200_LAUNCHED();

            // strict triangle Ts (values already carry the gather-side conjugation cg)
            dev_csr Ts;
            Ts.m = m;
            Ts.n = A.n;
            B200_TRY(Ts.row_ptr.alloc(sizeof(int) * ((size_t)m + 1)));
            B200_TRY(exclusive_scan(cnt.as<int>(), Ts.row_ptr.as<int>(), m + 1, st));
            int ts_nnz = 0;
            B200_CUDA(cudaMemcpyAsync(&ts_nnz, Ts.row_ptr.as<int>() + m, sizeof(int), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            Ts.nnz = ts_nnz;
            B200_TRY(Ts.col_idx.alloc(sizeof(int) * (size_t)ts_nnz));
            B200_TRY(Ts.val.alloc(sizeof(T) * (size_t)ts_nnz));
            tri_fill_kernel<T><<<grid_for(m, 128), 128, 0, st>>>(m,
                                                                lower,
                                                                cg,
                                                                A.row_ptr.as<int>(),
                                                                A.col_idx.as<int>(),
                                                                A.val.as<T>(),
                                                                Ts.row_ptr.as<int>(),
                                                                Ts.col_idx.as<int>(),
                                                                Ts.val.as<T>());
            B200_LAUNCHED();

            // mirror = stable transpose of Ts
            dev_csr Tt;
            B200_TRY(transpose_csr(Ts, val_type, false, Tt, st));

            B200_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int) * ((size_t)m + 1), st));
            full_count_kernel<<<grid_for(m, 256), 256, 0, st>>>(
                m, diag_mode, Ts.row_ptr.as<int>(), Tt.row_ptr.as<int>(), diag_pos.as<int>(), cnt.as<int>());
            B200_LAUNCHED();
            F.m = m;
            F.n = A.n;
            B200_TRY(F.row_ptr.alloc(sizeof(int) * ((size_t)m + 1)));
            B200_TRY(exclusive_scan(cnt.as<int>(), F.row_ptr.as<int>(), m + 1, st));
            int f_nnz = 0;
            B200_CUDA(cudaMemcpyAsync(&f_nnz, F.row_ptr.as<int>() + m, sizeof(int), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            F.nnz = f_nnz;
            B200_TRY(F.col_idx.alloc(sizeof(int) * (size_t)f_nnz));
            B200_TRY(F.val.alloc(sizeof(T) * (size_t)f_nnz));
            // Ts holds op_g(T); the mirror must hold op_s(T)^T, i.e. one more conjugation iff cg != cs
            assemble_kernel<T><<<grid_for(m, 128), 128, 0, st>>>(m,
                                                                lower,
                                                                diag_mode,
                                                                cd,
                                                                cg != cs ? 1 : 0,
                                                                Ts.row_ptr.as<int>(),
                                                                Ts.col_idx.as<int>(),
                                                                Ts.val.as<T>(),
                                                                Tt.row_ptr.as<int>(),
                                                                Tt.col_idx.as<int>(),
                                                                Tt.val.as<T>(),
                                                                diag_pos.as<int>(),
                                                                A.val.as<T>(),
                                                                F.row_ptr.as<int>(),
                                                                F.col_idx.as<int>(),
                                                                F.val.as<T>());
            B200_LAUNCHED();
            B200_CUDA(cudaStreamSynchronize(st));
            return aoclsparse_status_success;
        }
    }

    aoclsparse_status get_expanded_copy(aoclsparse_matrix            A,
                                        const _aoclsparse_mat_descr &descr_in,
                                        aoclsparse_operation         op,
                                        const dev_csr              *&out,
                                        cudaStream_t                 st)
    {
        const bool            cplx  = A->val_type == aoclsparse_cmat || A->val_type == aoclsparse_zmat;
        _aoclsparse_mat_descr descr = descr_in;
        if(A->is_csc)
        {
            // the stored arrays are those of A^T: the named triangle sits on the other side, and for a hermitian
            // matrix A^T = conj(A), i.e. op none <-> op transpose (trans_doid, csrmm.hpp:542-555)
            descr.fill_mode = descr_in.fill_mode == aoclsparse_fill_mode_lower ? aoclsparse_fill_mode_upper
                                                                               : aoclsparse_fill_mode_lower;
            if(cplx && descr.type == aoclsparse_matrix_type_hermitian)
                op = op == aoclsparse_operation_transpose ? aoclsparse_operation_none : aoclsparse_operation_transpose;
        }
        const int  d_id = get_doid(cplx, descr.type, descr.fill_mode, op);
        if(d_id < 4 || d_id > 11)
            return aoclsparse_status_invalid_value;
        // cache tag: doid in the low bits, diagonal type above
        const int tag = d_id | ((int)descr.diag_type << 8) | (1 << 16);
        std::unique_lock<std::shared_mutex> wl(A->guard);
        for(size_t i = 1; i < A->mats.size(); ++i)
            if(A->mats[i]->doid == tag && A->mats[i]->plan.valid)
            {
                out = A->mats[i];
                return aoclsparse_status_success;
            }
        const bool herm    = cplx && descr.type == aoclsparse_matrix_type_hermitian;
        const bool conj_op = cplx && op == aoclsparse_operation_conjugate_transpose;
        int        cg, cs, cd;
        if(!herm)
            cg = cs = cd = conj_op ? 1 : 0;
        else
        {
            const bool t = op == aoclsparse_operation_transpose;
            cg           = t ? 1 : 0;
            cs           = t ? 0 : 1;
            cd           = 0;
        }
        const int lower     = descr.fill_mode == aoclsparse_fill_mode_lower ? 1 : 0;
        const int diag_mode = descr.diag_type == aoclsparse_diag_type_unit ? 1 : (descr.diag_type == aoclsparse_diag_type_zero ? 2 : 0);
        dev_csr  *F         = new(std::nothrow) dev_csr;
        if(!F)
            return aoclsparse_status_memory_error;
        aoclsparse_status s;
        switch(A->val_type)
        {
        case aoclsparse_dmat:
            s = build_expanded<double>(*A->mats[0], A->val_type, lower, diag_mode, cg, cs, cd, *F, st);
            break;
        case aoclsparse_smat:
            s = build_expanded<float>(*A->mats[0], A->val_type, lower, diag_mode, cg, cs, cd, *F, st);
            break;
        case aoclsparse_cmat:
            s = build_expanded<float2>(*A->mats[0], A->val_type, lower, diag_mode, cg, cs, cd, *F, st);
            break;
        default:
            s = build_expanded<double2>(*A->mats[0], A->val_type, lower, diag_mode, cg, cs, cd, *F, st);
            break;
        }
        if(s == aoclsparse_status_success)
            s = build_plan(*F, value_size(A->val_type), -1, -1, std::vector<aoclsparse_int>(), st);
        if(s != aoclsparse_status_success)
        {
            delete F;
            return s;
        }
        F->doid = tag;
        A->mats.push_back(F);
        out = F;
        return aoclsparse_status_success;
    }
}
