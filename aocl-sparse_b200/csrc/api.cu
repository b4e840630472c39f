// api.cu -- handles, descriptors, hints, aoclsparse_optimize and the read-back extensions.
//
// Reference counterparts:
//   descriptor functions   library/src/extra/aoclsparse_auxiliary.cpp:191-362
//   create_csr<T>          library/src/create/aoclsparse_create.cpp:33-95
//   aoclsparse_destroy     library/src/extra/aoclsparse_auxiliary.cpp:657-670
//   set_hint / optimize    library/src/analysis/aoclsparse_analysis.cpp:426-747
//   update_values          library/src/extra/aoclsparse_auxiliary.hpp:215-272
//   get_doid               library/src/include/aoclsparse_mtx_dispatcher.hpp:79-143
#include "common.hpp"

#include <cstdlib>
#include <new>
#include <string>

namespace b200
{
    std::atomic<unsigned long long> g_launches{0};

    namespace
    {
        thread_local cudaStream_t tl_stream = nullptr;
        thread_local std::string  tl_error;
    }

    cudaStream_t current_stream()
    {
        return tl_stream;
    }

    void note_cuda_error(cudaError_t e, const char *where)
    {
        tl_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " @ " + where;
        cudaGetLastError(); // clear the sticky-less error so later calls are judged on their own
    }

    bool is_device_accessible(const void *p)
    {
        cudaPointerAttributes at;
        cudaError_t           e = cudaPointerGetAttributes(&at, p);
        if(e != cudaSuccess)
        {
            cudaGetLastError();
            return false;
        }
        return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
    }

    // device-side address of page-locked (pinned / registered) host memory, nullptr for anything else
    void *pinned_host_device_ptr(const void *p)
    {
        cudaPointerAttributes at;
        cudaError_t           e = cudaPointerGetAttributes(&at, p);
        if(e != cudaSuccess)
        {
            cudaGetLastError();
            return nullptr;
        }
        return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
    }

    // Same table as aoclsparse::get_doid<T>: [group:3][op:2], op bits 0 = conjugate, 1 = transpose
    // (for symmetric / hermitian bit 1 = upper triangle).
    int get_doid(bool complex_type, int descr_type, int fill_mode, int op)
    {
        int op_v = op - 111; // 0 none, 1 transpose, 2 conjugate transpose
        if(op_v < 0 || op_v > 2)
            return DOID_LEN;
        int t = descr_type;
        if(!complex_type)
        {
            if(op_v == 2)
                op_v = 1;
            if(t == aoclsparse_matrix_type_hermitian)
                t = aoclsparse_matrix_type_symmetric;
        }
        if(t == aoclsparse_matrix_type_symmetric && op_v == 1)
            op_v = 0;
        else if(t == aoclsparse_matrix_type_hermitian && op_v == 2)
            op_v = 0;
        static const int op_bits[3] = {0, 2, 3};
        switch(t)
        {
        case aoclsparse_matrix_type_general:
            return op_bits[op_v];
        case aoclsparse_matrix_type_symmetric:
            return 4 + 2 * fill_mode + (op_v >> 1);
        case aoclsparse_matrix_type_hermitian:
            return 8 + 2 * fill_mode + (op_v ^ fill_mode);
        case aoclsparse_matrix_type_triangular:
            return 12 + 4 * fill_mode + op_bits[op_v];
        default:
            return DOID_LEN;
        }
    }

    aoclsparse_status ensure_plan(aoclsparse_matrix A, cudaStream_t st)
    {
        {
            std::shared_lock<std::shared_mutex> rl(A->guard);
            if(A->mats[0]->plan.valid && !A->mats[0]->plan.ecodes_stale)
                return aoclsparse_status_success;
        }
        std::unique_lock<std::shared_mutex> wl(A->guard);
        dev_csr                            &M = *A->mats[0];
        if(M.plan.valid && M.plan.ecodes_stale)
        {
            // the stored values changed (aoclsparse_?update_values, aoclsparse_?set_value) since the entry-code copy was
            // built: encode them again; if they no longer fit a 256-entry table the multiply goes back to the kernel that
            // streams the values (diagonal-code copy, the handle's main plan)
            B200_TRY(build_entry_codes(M, value_size(A->val_type), A->max_row_nnz, A->row_cuts, st));
        }
        if(M.plan.valid)
            return aoclsparse_status_success;
        return build_plan(M, value_size(A->val_type), A->max_row_nnz, -1, A->row_cuts, st);
    }

    namespace
    {
        // copies `bytes` from a host or device source into device memory on `st`
        aoclsparse_status upload(void *dst, const void *src, size_t bytes, cudaStream_t st)
        {
            if(bytes == 0)
                return aoclsparse_status_success;
            B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st));
            return aoclsparse_status_success;
        }

        aoclsparse_status fetch_int(const aoclsparse_int *src, aoclsparse_int &out, cudaStream_t st)
        {
            if(is_device_accessible(src))
            {
                B200_CUDA(cudaMemcpyAsync(&out, src, sizeof(out), cudaMemcpyDeviceToHost, st));
                B200_CUDA(cudaStreamSynchronize(st));
            }
            else
                out = *src;
            return aoclsparse_status_success;
        }

        template <typename T>
        aoclsparse_status create_csr(aoclsparse_matrix    *mat,
                                     aoclsparse_index_base base,
                                     aoclsparse_int        M,
                                     aoclsparse_int        N,
                                     aoclsparse_int        nnz,
                                     const aoclsparse_int *row_ptr,
                                     const aoclsparse_int *col_idx,
                                     const void           *val,
                                     int                   val_type,
                                     bool                  validate = true,
                                     bool                  csc      = false)
        {
            // CSC input (aoclsparse_create_csc_t, library/src/extra/aoclsparse_auxiliary.cpp:1030-1089): the arrays
            // are the CSR of the transpose; they are validated and stored as such (N rows, M columns)
            const aoclsparse_int M_user = M, N_user = N;
            if(csc)
            {
                M = N_user;
                N = M_user;
            }
            if(!mat)
                return aoclsparse_status_invalid_pointer;
            *mat = nullptr;
            if(row_ptr == nullptr || col_idx == nullptr || val == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(M < 0 || N < 0 || nnz < 0)
                return aoclsparse_status_invalid_size;

            cudaStream_t   st = current_stream();
            aoclsparse_int first = 0, last = 0;
            B200_TRY(fetch_int(row_ptr, first, st));
            if(first - (aoclsparse_int)base != 0)
                return aoclsparse_status_invalid_value;
            B200_TRY(fetch_int(row_ptr + M, last, st));
            if(last - (aoclsparse_int)base != nnz)
                return aoclsparse_status_invalid_value;

            _aoclsparse_matrix *A = new(std::nothrow) _aoclsparse_matrix;
            b200::dev_csr      *C = new(std::nothrow) b200::dev_csr;
            if(!A || !C)
            {
                delete A;
                delete C;
                return aoclsparse_status_memory_error;
            }
            A->mats.push_back(C);
            auto fail = [&](aoclsparse_status s) {
                delete A;
                return s;
            };
            aoclsparse_status s;
            if((s = C->row_ptr.alloc(sizeof(aoclsparse_int) * ((size_t)M + 1))) != aoclsparse_status_success)
                return fail(s);
            if((s = C->col_idx.alloc(sizeof(aoclsparse_int) * (size_t)nnz)) != aoclsparse_status_success)
                return fail(s);
            if((s = C->val.alloc(sizeof(T) * (size_t)nnz)) != aoclsparse_status_success)
                return fail(s);
            if((s = upload(C->row_ptr.p, row_ptr, sizeof(aoclsparse_int) * ((size_t)M + 1), st)) != aoclsparse_status_success)
                return fail(s);
            if((s = upload(C->col_idx.p, col_idx, sizeof(aoclsparse_int) * (size_t)nnz, st)) != aoclsparse_status_success)
                return fail(s);
            if((s = upload(C->val.p, val, sizeof(T) * (size_t)nnz, st)) != aoclsparse_status_success)
                return fail(s);

            b200::check_result cr;
            if((s = b200::check_csr_device(
                    M, N, nnz, (int)base, C->row_ptr.as<aoclsparse_int>(), C->col_idx.as<aoclsparse_int>(), cr, st))
               != aoclsparse_status_success)
                return fail(s);
            if(cr.status != aoclsparse_status_success && (validate || cr.status != aoclsparse_status_invalid_value))
                return fail(cr.status); // (the handle-free legacy entry tolerates duplicate diagonals)
            if(base == aoclsparse_index_base_one)
            {
                if((s = b200::rebase_to_zero(M, nnz, C->row_ptr.as<aoclsparse_int>(), C->col_idx.as<aoclsparse_int>(), st))
                   != aoclsparse_status_success)
                    return fail(s);
            }
            // the caller may free its arrays as soon as we return
            cudaError_t e = cudaStreamSynchronize(st);
            if(e != cudaSuccess)
                return fail(b200::cuda_status(e, "create sync"));

            C->m = M;
            C->n = N;
            C->nnz = nnz;
            C->doid = b200::DOID_GN;
            A->m = M_user;
            A->n = N_user;
            A->is_csc = csc;
            A->nnz = nnz;
            A->base = base;
            A->val_type = (aoclsparse_matrix_data_type)val_type;
            A->input_format = aoclsparse_csr_mat;
            A->sort = (aoclsparse_matrix_sort)cr.sort;
            A->fulldiag = cr.fulldiag != 0;
            A->min_col = cr.min_col;
            A->max_col = cr.max_col;
            A->max_row_nnz = cr.max_row_nnz;
            cudaGetDevice(&A->device);
            *mat = A;
            return aoclsparse_status_success;
        }

    }

    aoclsparse_status create_temp_csr(aoclsparse_matrix    *mat,
                                      int                   val_type,
                                      aoclsparse_index_base base,
                                      aoclsparse_int        M,
                                      aoclsparse_int        N,
                                      aoclsparse_int        nnz,
                                      const aoclsparse_int *row_ptr,
                                      const aoclsparse_int *col_idx,
                                      const void           *val)
    {
        if(val_type == aoclsparse_smat)
            return create_csr<float>(mat, base, M, N, nnz, row_ptr, col_idx, val, val_type, false);
        return create_csr<double>(mat, base, M, N, nnz, row_ptr, col_idx, val, val_type, false);
    }

    void drop_derived_copies(aoclsparse_matrix A)
    {
        for(size_t i = 1; i < A->mats.size(); ++i)
            delete A->mats[i];
        A->mats.resize(1);
        A->mats[0]->tiles = b200::mesh_tiles(); // holds a copy of the values
        if(A->mats[0]->plan.n_ecodes > 0)
            A->mats[0]->plan.ecodes_stale = true; // the entry-code copy is re-encoded before the next multiply (ensure_plan)
        A->clean = b200::clean_csr();
        for(auto &h : A->hints)
            h.done = false;
    }

    namespace
    {
        template <typename T>
        aoclsparse_status update_values(aoclsparse_matrix A, aoclsparse_int len, const void *val)
        {
            if(A == nullptr || val == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(A->mats.empty() || A->mats[0] == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(len != A->nnz)
                return aoclsparse_status_invalid_size;
            if(A->val_type != b200::vt<T>::data_type)
                return aoclsparse_status_wrong_type;
            cudaStream_t                        st = current_stream();
            std::unique_lock<std::shared_mutex> wl(A->guard);
            B200_TRY(upload(A->mats[0]->val.p, val, sizeof(T) * (size_t)len, st));
            B200_CUDA(cudaStreamSynchronize(st));
            // derived copies hold stale values: drop them (the reference does the same,
            // aoclsparse_auxiliary.hpp:260-272); the row-block plan depends on the pattern only and stays
            drop_derived_copies(A);
            return aoclsparse_status_success;
        }

        aoclsparse_status set_hint(aoclsparse_matrix          mat,
                                   int                        act,
                                   aoclsparse_operation       trans,
                                   const aoclsparse_mat_descr descr,
                                   aoclsparse_int             calls,
                                   aoclsparse_int             kid)
        {
            if(!mat || mat->mats.empty() || mat->mats[0] == nullptr || descr == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(descr->base != aoclsparse_index_base_zero && descr->base != aoclsparse_index_base_one)
                return aoclsparse_status_invalid_value;
            if(descr->base != mat->base)
                return aoclsparse_status_invalid_value;
            if(trans != aoclsparse_operation_none && trans != aoclsparse_operation_transpose
               && trans != aoclsparse_operation_conjugate_transpose)
                return aoclsparse_status_invalid_value;
            if(descr->fill_mode != aoclsparse_fill_mode_lower && descr->fill_mode != aoclsparse_fill_mode_upper)
                return aoclsparse_status_invalid_value;
            if(descr->diag_type != aoclsparse_diag_type_non_unit && descr->diag_type != aoclsparse_diag_type_unit
               && descr->diag_type != aoclsparse_diag_type_zero)
                return aoclsparse_status_invalid_value;
            if(descr->type != aoclsparse_matrix_type_general && descr->type != aoclsparse_matrix_type_symmetric
               && descr->type != aoclsparse_matrix_type_triangular && descr->type != aoclsparse_matrix_type_hermitian)
                return aoclsparse_status_invalid_value;
            if(calls < 0 || (calls == 0 && kid == -1))
                return aoclsparse_status_invalid_value;
            b200::hint h;
            h.act       = act;
            h.trans     = trans;
            h.type      = descr->type;
            h.fill_mode = descr->fill_mode;
            h.nop       = calls;
            h.kid       = kid;
            const bool cplx = mat->val_type == aoclsparse_cmat || mat->val_type == aoclsparse_zmat;
            h.doid      = b200::get_doid(cplx, descr->type, descr->fill_mode, trans);
            std::unique_lock<std::shared_mutex> wl(mat->guard);
            mat->hints.insert(mat->hints.begin(), h);
            return aoclsparse_status_success;
        }
    }
}

using namespace b200;

extern "C" {

const char *aoclsparse_get_version(void)
{
    return "AOCL-Sparse 5.3.2 (b200-native CSR SpMV/SpMM path, sm_100a)";
}

aoclsparse_status aoclsparse_b200_set_stream(void *cuda_stream)
{
    tl_stream = static_cast<cudaStream_t>(cuda_stream);
    return aoclsparse_status_success;
}

// device selection for callers that do not link the CUDA runtime themselves (a C program driving several GPUs)
aoclsparse_status aoclsparse_b200_device_count(int *count)
{
    if(!count)
        return aoclsparse_status_invalid_pointer;
    B200_CUDA(cudaGetDeviceCount(count));
    return aoclsparse_status_success;
}
aoclsparse_status aoclsparse_b200_set_device(int device)
{
    B200_CUDA(cudaSetDevice(device));
    return aoclsparse_status_success;
}

void *aoclsparse_b200_get_stream(void)
{
    return tl_stream;
}

const char *aoclsparse_b200_last_error(void)
{
    return tl_error.c_str();
}

unsigned long long aoclsparse_b200_launch_count(void)
{
    return g_launches.load();
}

// ------------------------------------------------------------------ descriptor
aoclsparse_status aoclsparse_create_mat_descr(aoclsparse_mat_descr *descr)
{
    if(descr == nullptr)
        return aoclsparse_status_invalid_pointer;
    *descr = new(std::nothrow) _aoclsparse_mat_descr;
    return *descr ? aoclsparse_status_success : aoclsparse_status_memory_error;
}

aoclsparse_status aoclsparse_copy_mat_descr(aoclsparse_mat_descr dest, const aoclsparse_mat_descr src)
{
    if(dest == nullptr || src == nullptr)
        return aoclsparse_status_invalid_pointer;
    *dest = *src;
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_destroy_mat_descr(aoclsparse_mat_descr descr)
{
    delete descr;
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_set_mat_index_base(aoclsparse_mat_descr descr, aoclsparse_index_base base)
{
    if(descr == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(base != aoclsparse_index_base_zero && base != aoclsparse_index_base_one)
        return aoclsparse_status_invalid_value;
    descr->base = base;
    return aoclsparse_status_success;
}

aoclsparse_index_base aoclsparse_get_mat_index_base(const aoclsparse_mat_descr descr)
{
    return descr ? descr->base : aoclsparse_index_base_zero;
}

aoclsparse_status aoclsparse_set_mat_type(aoclsparse_mat_descr descr, aoclsparse_matrix_type type)
{
    if(descr == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(type != aoclsparse_matrix_type_general && type != aoclsparse_matrix_type_symmetric
       && type != aoclsparse_matrix_type_hermitian && type != aoclsparse_matrix_type_triangular)
        return aoclsparse_status_invalid_value;
    descr->type = type;
    return aoclsparse_status_success;
}

aoclsparse_matrix_type aoclsparse_get_mat_type(const aoclsparse_mat_descr descr)
{
    return descr ? descr->type : aoclsparse_matrix_type_general;
}

aoclsparse_status aoclsparse_set_mat_fill_mode(aoclsparse_mat_descr descr, aoclsparse_fill_mode fill_mode)
{
    if(descr == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(fill_mode != aoclsparse_fill_mode_lower && fill_mode != aoclsparse_fill_mode_upper)
        return aoclsparse_status_invalid_value;
    descr->fill_mode = fill_mode;
    return aoclsparse_status_success;
}

aoclsparse_fill_mode aoclsparse_get_mat_fill_mode(const aoclsparse_mat_descr descr)
{
    return descr ? descr->fill_mode : aoclsparse_fill_mode_lower;
}

aoclsparse_status aoclsparse_set_mat_diag_type(aoclsparse_mat_descr descr, aoclsparse_diag_type diag_type)
{
    if(descr == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(diag_type != aoclsparse_diag_type_unit && diag_type != aoclsparse_diag_type_non_unit
       && diag_type != aoclsparse_diag_type_zero)
        return aoclsparse_status_invalid_value;
    descr->diag_type = diag_type;
    return aoclsparse_status_success;
}

aoclsparse_diag_type aoclsparse_get_mat_diag_type(const aoclsparse_mat_descr descr)
{
    return descr ? descr->diag_type : aoclsparse_diag_type_non_unit;
}

// ------------------------------------------------------------------ matrix handle
aoclsparse_status aoclsparse_create_scsr(aoclsparse_matrix    *mat,
                                         aoclsparse_index_base base,
                                         aoclsparse_int        M,
                                         aoclsparse_int        N,
                                         aoclsparse_int        nnz,
                                         aoclsparse_int       *row_ptr,
                                         aoclsparse_int       *col_idx,
                                         float                *val)
{
    return create_csr<float>(mat, base, M, N, nnz, row_ptr, col_idx, val, aoclsparse_smat);
}
aoclsparse_status aoclsparse_create_dcsr(aoclsparse_matrix    *mat,
                                         aoclsparse_index_base base,
                                         aoclsparse_int        M,
                                         aoclsparse_int        N,
                                         aoclsparse_int        nnz,
                                         aoclsparse_int       *row_ptr,
                                         aoclsparse_int       *col_idx,
                                         double               *val)
{
    return create_csr<double>(mat, base, M, N, nnz, row_ptr, col_idx, val, aoclsparse_dmat);
}
aoclsparse_status aoclsparse_create_ccsr(aoclsparse_matrix        *mat,
                                         aoclsparse_index_base     base,
                                         aoclsparse_int            M,
                                         aoclsparse_int            N,
                                         aoclsparse_int            nnz,
                                         aoclsparse_int           *row_ptr,
                                         aoclsparse_int           *col_idx,
                                         aoclsparse_float_complex *val)
{
    return create_csr<float2>(mat, base, M, N, nnz, row_ptr, col_idx, val, aoclsparse_cmat);
}
aoclsparse_status aoclsparse_create_zcsr(aoclsparse_matrix         *mat,
                                         aoclsparse_index_base      base,
                                         aoclsparse_int             M,
                                         aoclsparse_int             N,
                                         aoclsparse_int             nnz,
                                         aoclsparse_int            *row_ptr,
                                         aoclsparse_int            *col_idx,
                                         aoclsparse_double_complex *val)
{
    return create_csr<double2>(mat, base, M, N, nnz, row_ptr, col_idx, val, aoclsparse_zmat);
}

aoclsparse_status aoclsparse_create_scsc(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *col_ptr, aoclsparse_int *row_idx, float *val)
{
    return create_csr<float>(mat, base, M, N, nnz, col_ptr, row_idx, val, aoclsparse_smat, true, true);
}
aoclsparse_status aoclsparse_create_dcsc(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *col_ptr, aoclsparse_int *row_idx, double *val)
{
    return create_csr<double>(mat, base, M, N, nnz, col_ptr, row_idx, val, aoclsparse_dmat, true, true);
}
aoclsparse_status aoclsparse_create_ccsc(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *col_ptr, aoclsparse_int *row_idx, aoclsparse_float_complex *val)
{
    return create_csr<float2>(mat, base, M, N, nnz, col_ptr, row_idx, val, aoclsparse_cmat, true, true);
}
aoclsparse_status aoclsparse_create_zcsc(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *col_ptr, aoclsparse_int *row_idx, aoclsparse_double_complex *val)
{
    return create_csr<double2>(mat, base, M, N, nnz, col_ptr, row_idx, val, aoclsparse_zmat, true, true);
}

aoclsparse_status aoclsparse_destroy(aoclsparse_matrix *mat)
{
    if(mat && *mat)
    {
        delete *mat;
        *mat = nullptr;
    }
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_supdate_values(aoclsparse_matrix A, aoclsparse_int len, float *val)
{
    return update_values<float>(A, len, val);
}
aoclsparse_status aoclsparse_dupdate_values(aoclsparse_matrix A, aoclsparse_int len, double *val)
{
    return update_values<double>(A, len, val);
}
aoclsparse_status aoclsparse_cupdate_values(aoclsparse_matrix A, aoclsparse_int len, aoclsparse_float_complex *val)
{
    return update_values<float2>(A, len, val);
}
aoclsparse_status aoclsparse_zupdate_values(aoclsparse_matrix A, aoclsparse_int len, aoclsparse_double_complex *val)
{
    return update_values<double2>(A, len, val);
}

// ------------------------------------------------------------------ hints + optimize
aoclsparse_status aoclsparse_set_mv_hint(aoclsparse_matrix          mat,
                                         aoclsparse_operation       trans,
                                         const aoclsparse_mat_descr descr,
                                         aoclsparse_int             expected_no_of_calls)
{
    return set_hint(mat, 1, trans, descr, expected_no_of_calls, -1);
}
aoclsparse_status aoclsparse_set_mv_hint_kid(aoclsparse_matrix          mat,
                                             aoclsparse_operation       trans,
                                             const aoclsparse_mat_descr descr,
                                             aoclsparse_int             expected_no_of_calls,
                                             aoclsparse_int             kid)
{
    return set_hint(mat, 1, trans, descr, expected_no_of_calls, kid);
}
aoclsparse_status aoclsparse_set_mm_hint(aoclsparse_matrix          mat,
                                         aoclsparse_operation       trans,
                                         const aoclsparse_mat_descr descr,
                                         aoclsparse_int             expected_no_of_calls)
{
    return set_hint(mat, 3, trans, descr, expected_no_of_calls, -1);
}
// Hints for operations next to the path (aoclsparse_analysis.h:102-161,202-206): recorded with the same validation
// (aoclsparse_analysis.cpp:568-670; action numbering of aoclsparse_hinted_action, aoclsparse_mat_structures.hpp:32-48).
// A dotmv hint is a multiply hint for aoclsparse_optimize; the solver-side ones (triangular solve, smoothers, sparse
// products) only take their place in the list -- those operations are outside this library's path or need no copy.
aoclsparse_status aoclsparse_set_sv_hint(aoclsparse_matrix mat, aoclsparse_operation trans, const aoclsparse_mat_descr descr, aoclsparse_int expected_no_of_calls)
{
    return set_hint(mat, 2, trans, descr, expected_no_of_calls, -1);
}
aoclsparse_status aoclsparse_set_2m_hint(aoclsparse_matrix mat, aoclsparse_operation trans, const aoclsparse_mat_descr descr, aoclsparse_int expected_no_of_calls)
{
    return set_hint(mat, 4, trans, descr, expected_no_of_calls, -1);
}
aoclsparse_status aoclsparse_set_lu_smoother_hint(aoclsparse_matrix mat, aoclsparse_operation trans, const aoclsparse_mat_descr descr, aoclsparse_int expected_no_of_calls)
{
    return set_hint(mat, 5, trans, descr, expected_no_of_calls, -1);
}
aoclsparse_status aoclsparse_set_sm_hint(aoclsparse_matrix mat, aoclsparse_operation trans, const aoclsparse_mat_descr descr, const aoclsparse_order order, const aoclsparse_int expected_no_of_calls)
{
    if(order != aoclsparse_order_row && order != aoclsparse_order_column)
        return aoclsparse_status_invalid_value;
    return set_hint(mat, order == aoclsparse_order_row ? 6 : 7, trans, descr, expected_no_of_calls, -1);
}
aoclsparse_status aoclsparse_set_dotmv_hint(aoclsparse_matrix mat, aoclsparse_operation trans, const aoclsparse_mat_descr descr, aoclsparse_int expected_no_of_calls)
{
    return set_hint(mat, 8, trans, descr, expected_no_of_calls, -1);
}
aoclsparse_status aoclsparse_set_symgs_hint(aoclsparse_matrix mat, aoclsparse_operation trans, const aoclsparse_mat_descr descr, aoclsparse_int expected_no_of_calls)
{
    return set_hint(mat, 9, trans, descr, expected_no_of_calls, -1);
}
aoclsparse_status aoclsparse_set_memory_hint(aoclsparse_matrix mat, const aoclsparse_memory_usage policy)
{
    if(mat == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(policy != aoclsparse_memory_usage_minimal && policy != aoclsparse_memory_usage_unrestricted)
        return aoclsparse_status_invalid_value;
    mat->mem_policy = policy;
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_optimize(aoclsparse_matrix A)
{
    if(!A)
        return aoclsparse_status_invalid_pointer;
    if(A->m < 0 || A->n < 0 || A->nnz < 0)
        return aoclsparse_status_invalid_size;
    if(A->mats.empty() || A->mats[0] == nullptr)
        return aoclsparse_status_invalid_pointer;
    cudaStream_t                        st = current_stream();
    std::unique_lock<std::shared_mutex> wl(A->guard);
    dev_csr                            &M = *A->mats[0];

    // a kid >= 0 on a plain general mv hint forces the row strategy of every block
    aoclsparse_int forced = -1;
    for(const hint &h : A->hints)
        if(h.act == 1 && h.doid == DOID_GN && h.kid >= 0 && h.kid <= 2)
        {
            forced = h.kid;
            break;
        }
    if(const char *e = getenv("AOCLSPARSE_B200_FORCE_KID")) // tuning knob for experiments
        forced = atoi(e);
    // a multiply was hinted and memory is not restricted: banded / stencil matrices also get the diagonal-code copy of
    // their column indices (one byte per entry; the reference's optimize builds format copies at this point too,
    // aoclsparse_analysis.cpp:192-385), and a plan whose block size suits it
    const bool want_codes = A->mem_policy == aoclsparse_memory_usage_unrestricted && !A->hints.empty();
    if(want_codes && (!M.plan.valid || forced >= 0 || M.plan.code_state == 0))
        B200_TRY(build_plan_with_codes(M, value_size(A->val_type), A->max_row_nnz, forced, A->row_cuts, st));
    else if(!M.plan.valid || forced >= 0)
        B200_TRY(build_plan(M, value_size(A->val_type), A->max_row_nnz, forced, A->row_cuts, st));

    // transposed device copies for general transposed mv / mm hints (memory policy permitting):
    // they turn the atomic scatter into a streaming gather
    if(A->mem_policy == aoclsparse_memory_usage_unrestricted && A->win_hi < 0)
    {
        for(hint &h : A->hints)
        {
            if(h.done)
                continue;
            // a CSC handle stores the transpose: its plain product is the transposed product of what is stored
            const int want = A->is_csc ? (h.doid == DOID_GN ? DOID_GT : DOID_LEN) : h.doid;
            if((h.act == 1 || h.act == 3 || h.act == 8) && (want == DOID_GT || want == DOID_GH))
            {
                bool have = false;
                for(size_t i = 1; i < A->mats.size(); ++i)
                    have = have || A->mats[i]->doid == want;
                if(!have)
                {
                    dev_csr *C = new(std::nothrow) dev_csr;
                    if(!C)
                        return aoclsparse_status_memory_error;
                    aoclsparse_status s = transpose_csr(M, A->val_type, want == DOID_GH, *C, st);
                    if(s == aoclsparse_status_success)
                        s = build_plan_with_codes(*C, value_size(A->val_type), -1, -1, std::vector<aoclsparse_int>(), st);
                    if(s != aoclsparse_status_success)
                    {
                        delete C;
                        return s;
                    }
                    C->doid = want;
                    A->mats.push_back(C);
                }
            }
            h.done = true;
        }
    }
    B200_CUDA(cudaStreamSynchronize(st));
    return aoclsparse_status_success;
}

// ------------------------------------------------------------------ read-back extensions
aoclsparse_status aoclsparse_b200_get_matrix_info(const aoclsparse_matrix A, aoclsparse_b200_matrix_info *info)
{
    if(!A || !info)
        return aoclsparse_status_invalid_pointer;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    memset(info, 0, sizeof(*info));
    info->m           = A->m;
    info->n           = A->n;
    info->nnz         = A->nnz;
    info->base        = A->base;
    info->val_type    = A->val_type;
    info->sort        = A->sort;
    info->fulldiag    = A->fulldiag ? 1 : 0;
    info->min_col     = A->min_col;
    info->max_col     = A->max_col;
    info->max_row_nnz = A->max_row_nnz;
    info->n_hints     = (int)A->hints.size();
    info->n_copies    = (int)A->mats.size();
    const row_block_plan &P = A->mats[0]->plan;
    info->optimized         = P.valid ? 1 : 0;
    if(P.valid)
    {
        info->block_nnz        = P.block_nnz;
        info->block_rows       = P.block_rows;
        info->n_blocks         = P.n_blocks;
        info->n_thread_blocks  = P.n_strat[STRAT_THREAD];
        info->n_warp_blocks    = P.n_strat[STRAT_WARP];
        info->n_product_blocks = P.n_strat[STRAT_PRODUCT];
        info->n_long_segments  = P.n_long_segments;
        info->n_long_rows      = P.n_long_rows;
        info->n_diag_codes     = P.n_codes;
        info->n_entry_codes    = P.ecodes_stale ? 0 : P.n_ecodes;
        if(P.eplan && P.n_ecodes > 0)
        {
            info->e_block_nnz  = P.eplan->block_nnz;
            info->e_block_rows = P.eplan->block_rows;
            info->e_n_blocks   = P.eplan->n_blocks;
        }
    }
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_get_plan(const aoclsparse_matrix A,
                                           aoclsparse_int          capacity,
                                           aoclsparse_int         *block_desc,
                                           aoclsparse_int         *block_kind,
                                           aoclsparse_int         *n_blocks)
{
    if(!A || !n_blocks)
        return aoclsparse_status_invalid_pointer;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const row_block_plan               &P = A->mats[0]->plan;
    if(!P.valid)
        return aoclsparse_status_invalid_operation;
    *n_blocks = P.n_blocks;
    if(capacity < P.n_blocks)
        return (block_desc || block_kind) ? aoclsparse_status_invalid_size : aoclsparse_status_success;
    cudaStream_t st = current_stream();
    if(block_desc && P.n_blocks > 0)
        B200_CUDA(cudaMemcpyAsync(block_desc, P.desc.p, sizeof(int4) * (size_t)P.n_blocks, cudaMemcpyDeviceToHost, st));
    if(block_kind && P.n_blocks > 0)
        B200_CUDA(cudaMemcpyAsync(block_kind, P.kind.p, sizeof(int) * (size_t)P.n_blocks, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_get_entry_plan(const aoclsparse_matrix A,
                                                 aoclsparse_int          capacity,
                                                 aoclsparse_int         *block_desc,
                                                 aoclsparse_int         *block_kind,
                                                 aoclsparse_int         *n_blocks)
{
    if(!A || !n_blocks)
        return aoclsparse_status_invalid_pointer;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const row_block_plan               &M = A->mats[0]->plan;
    *n_blocks                             = 0;
    if(!M.valid || !M.eplan || M.n_ecodes <= 0)
        return aoclsparse_status_success;
    const row_block_plan &P = *M.eplan;
    *n_blocks               = P.n_blocks;
    if(capacity < P.n_blocks)
        return (block_desc || block_kind) ? aoclsparse_status_invalid_size : aoclsparse_status_success;
    cudaStream_t st = current_stream();
    if(block_desc && P.n_blocks > 0)
        B200_CUDA(cudaMemcpyAsync(block_desc, P.desc.p, sizeof(int4) * (size_t)P.n_blocks, cudaMemcpyDeviceToHost, st));
    if(block_kind && P.n_blocks > 0)
        B200_CUDA(cudaMemcpyAsync(block_kind, P.kind.p, sizeof(int) * (size_t)P.n_blocks, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_get_diag_codes(const aoclsparse_matrix A,
                                                 aoclsparse_int         *n_codes,
                                                 aoclsparse_int         *offsets,
                                                 unsigned char          *codes)
{
    if(!A || !n_codes)
        return aoclsparse_status_invalid_pointer;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const row_block_plan               &P = A->mats[0]->plan;
    *n_codes                              = P.valid ? P.n_codes : 0;
    if(*n_codes == 0 || (!offsets && !codes))
        return aoclsparse_status_success;
    cudaStream_t st = current_stream();
    if(offsets)
        B200_CUDA(cudaMemcpyAsync(offsets, P.code_offsets.p, sizeof(int) * (size_t)P.n_codes, cudaMemcpyDeviceToHost, st));
    if(codes)
        B200_CUDA(cudaMemcpyAsync(codes, P.codes.p, (size_t)A->mats[0]->nnz, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_get_entry_codes(const aoclsparse_matrix A,
                                                  aoclsparse_int         *n_pairs,
                                                  aoclsparse_int         *offsets,
                                                  void                   *values,
                                                  unsigned char          *ecodes)
{
    if(!A || !n_pairs)
        return aoclsparse_status_invalid_pointer;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const row_block_plan               &P = A->mats[0]->plan;
    *n_pairs                              = (P.valid && !P.ecodes_stale) ? P.n_ecodes : 0;
    if(*n_pairs == 0 || (!offsets && !values && !ecodes))
        return aoclsparse_status_success;
    cudaStream_t st = current_stream();
    if(offsets)
        B200_CUDA(cudaMemcpyAsync(offsets, P.etab_off.p, sizeof(int) * (size_t)P.n_ecodes, cudaMemcpyDeviceToHost, st));
    if(values)
        B200_CUDA(cudaMemcpyAsync(values, P.etab_val.p, value_size(A->val_type) * (size_t)P.n_ecodes, cudaMemcpyDeviceToHost, st));
    if(ecodes)
        B200_CUDA(cudaMemcpyAsync(ecodes, P.ecodes.p, (size_t)A->mats[0]->nnz, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return aoclsparse_status_success;
}

int aoclsparse_b200_value_type(const aoclsparse_matrix A)
{
    return A ? (int)A->val_type : -1;
}

int aoclsparse_b200_doid(const aoclsparse_mat_descr descr, aoclsparse_operation op, int val_type)
{
    if(!descr)
        return DOID_LEN;
    const bool cplx = val_type == aoclsparse_cmat || val_type == aoclsparse_zmat;
    return get_doid(cplx, descr->type, descr->fill_mode, op);
}

// ids are [family:3][variant:2]; family 0 general, 1 symmetric, 2 hermitian, 3 / 4 triangular lower / upper
int aoclsparse_b200_effective_doid(int mat_doid, int req_doid)
{
    if(mat_doid < 0 || mat_doid >= DOID_LEN || req_doid < 0 || req_doid >= DOID_LEN)
        return DOID_LEN;
    const int mf = mat_doid >> 2, rf = req_doid >> 2, mv = mat_doid & 3, rv = req_doid & 3;
    if(mf == rf)
    {
        // same family: what is left to do is the difference of the two variants; a symmetric /
        // hermitian copy cannot be turned into the other triangle
        const bool tri_or_gen = (mf == 0 || mf >= 3);
        return (tri_or_gen || (mv ^ rv) <= 1) ? (mv ^ rv) : DOID_LEN;
    }
    if(mf != 0)
        return DOID_LEN; // only a general copy can serve another family
    if(rf >= 3)
    {
        // a stored transpose swaps the triangle and the transpose bit, a stored conjugate flips bit 0
        static const int flip[4] = {0, 1, 6, 7};
        return 12 + ((req_doid - 12) ^ flip[mv]);
    }
    return req_doid ^ mv;
}

aoclsparse_status aoclsparse_b200_set_x_window(aoclsparse_matrix A, aoclsparse_int col_lo, aoclsparse_int col_hi)
{
    if(!A)
        return aoclsparse_status_invalid_pointer;
    if(A->is_csc)
        return aoclsparse_status_not_implemented;
    if(col_lo < 0 || col_hi > A->n || col_lo > col_hi)
        return aoclsparse_status_invalid_size;
    std::unique_lock<std::shared_mutex> wl(A->guard);
    if(col_lo == 0 && col_hi == A->n)
    {
        A->win_lo = 0;
        A->win_hi = -1;
        return aoclsparse_status_success;
    }
    if(A->nnz > 0 && (A->min_col < col_lo || A->max_col >= col_hi))
        return aoclsparse_status_invalid_index_value;
    A->win_lo = col_lo;
    A->win_hi = col_hi;
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_set_row_cuts(aoclsparse_matrix A, aoclsparse_int n_cuts, const aoclsparse_int *cuts)
{
    if(!A || (n_cuts > 0 && !cuts))
        return aoclsparse_status_invalid_pointer;
    if(n_cuts < 0)
        return aoclsparse_status_invalid_size;
    for(aoclsparse_int i = 0; i < n_cuts; ++i)
        if(cuts[i] <= 0 || cuts[i] >= A->m || (i > 0 && cuts[i] <= cuts[i - 1]))
            return aoclsparse_status_invalid_value;
    std::unique_lock<std::shared_mutex> wl(A->guard);
    A->row_cuts.assign(cuts, cuts + n_cuts);
    A->mats[0]->plan.valid = false;
    return aoclsparse_status_success;
}

}
