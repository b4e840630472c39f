// cxx_api.cu -- the C++ template front ends the reference exports from its shared library
// (library/include/aoclsparse.hpp:85-147): aoclsparse::mv<T>, aoclsparse::create_csr<T>, aoclsparse::sp2m<T> for
// float, double, std::complex<float>, std::complex<double>.  Same mangled names; each forwards to the C entry.
#include "aoclsparse.hpp"

namespace
{
    template <typename T>
    struct cxx_type;
    template <>
    struct cxx_type<float>
    {
        static constexpr aoclsparse_matrix_data_type id = aoclsparse_smat;
    };
    template <>
    struct cxx_type<double>
    {
        static constexpr aoclsparse_matrix_data_type id = aoclsparse_dmat;
    };
    template <>
    struct cxx_type<std::complex<float>>
    {
        static constexpr aoclsparse_matrix_data_type id = aoclsparse_cmat;
    };
    template <>
    struct cxx_type<std::complex<double>>
    {
        static constexpr aoclsparse_matrix_data_type id = aoclsparse_zmat;
    };
}

extern "C" int aoclsparse_b200_value_type(const aoclsparse_matrix A); // api.cu: -1 for a NULL handle

namespace aoclsparse
{
    template <>
    DLL_PUBLIC aoclsparse_status mv<float>(aoclsparse_operation op, const float *alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const float *x, const float *beta, float *y)
    {
        return aoclsparse_smv(op, alpha, A, descr, x, beta, y);
    }
    template <>
    DLL_PUBLIC aoclsparse_status mv<double>(aoclsparse_operation op, const double *alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const double *x, const double *beta, double *y)
    {
        return aoclsparse_dmv(op, alpha, A, descr, x, beta, y);
    }
    template <>
    DLL_PUBLIC aoclsparse_status mv<std::complex<float>>(aoclsparse_operation op, const std::complex<float> *alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const std::complex<float> *x, const std::complex<float> *beta, std::complex<float> *y)
    {
        return aoclsparse_cmv(op,
                              reinterpret_cast<const aoclsparse_float_complex *>(alpha),
                              A,
                              descr,
                              reinterpret_cast<const aoclsparse_float_complex *>(x),
                              reinterpret_cast<const aoclsparse_float_complex *>(beta),
                              reinterpret_cast<aoclsparse_float_complex *>(y));
    }
    template <>
    DLL_PUBLIC aoclsparse_status mv<std::complex<double>>(aoclsparse_operation op, const std::complex<double> *alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const std::complex<double> *x, const std::complex<double> *beta, std::complex<double> *y)
    {
        return aoclsparse_zmv(op,
                              reinterpret_cast<const aoclsparse_double_complex *>(alpha),
                              A,
                              descr,
                              reinterpret_cast<const aoclsparse_double_complex *>(x),
                              reinterpret_cast<const aoclsparse_double_complex *>(beta),
                              reinterpret_cast<aoclsparse_double_complex *>(y));
    }

    template <>
    DLL_PUBLIC aoclsparse_status create_csr<float>(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *row_ptr, aoclsparse_int *col_idx, float *val, bool)
    {
        return aoclsparse_create_scsr(mat, base, M, N, nnz, row_ptr, col_idx, val);
    }
    template <>
    DLL_PUBLIC aoclsparse_status create_csr<double>(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *row_ptr, aoclsparse_int *col_idx, double *val, bool)
    {
        return aoclsparse_create_dcsr(mat, base, M, N, nnz, row_ptr, col_idx, val);
    }
    template <>
    DLL_PUBLIC aoclsparse_status create_csr<std::complex<float>>(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *row_ptr, aoclsparse_int *col_idx, std::complex<float> *val, bool)
    {
        return aoclsparse_create_ccsr(mat, base, M, N, nnz, row_ptr, col_idx, reinterpret_cast<aoclsparse_float_complex *>(val));
    }
    template <>
    DLL_PUBLIC aoclsparse_status create_csr<std::complex<double>>(aoclsparse_matrix *mat, aoclsparse_index_base base, aoclsparse_int M, aoclsparse_int N, aoclsparse_int nnz, aoclsparse_int *row_ptr, aoclsparse_int *col_idx, std::complex<double> *val, bool)
    {
        return aoclsparse_create_zcsr(mat, base, M, N, nnz, row_ptr, col_idx, reinterpret_cast<aoclsparse_double_complex *>(val));
    }

    // csr2m.cpp:603-628: pointer checks, then both operands must hold values of type T
    template <typename T>
    static aoclsparse_status sp2m_typed(aoclsparse_operation opA, const aoclsparse_mat_descr descrA, const aoclsparse_matrix A, aoclsparse_operation opB, const aoclsparse_mat_descr descrB, const aoclsparse_matrix B, aoclsparse_request request, aoclsparse_matrix *C)
    {
        if(descrA == nullptr || descrB == nullptr || A == nullptr || B == nullptr || C == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(aoclsparse_b200_value_type(A) != (int)cxx_type<T>::id || aoclsparse_b200_value_type(B) != (int)cxx_type<T>::id)
        {
            if(request != aoclsparse_stage_finalize)
                *C = nullptr;
            return aoclsparse_status_wrong_type;
        }
        return aoclsparse_sp2m(opA, descrA, A, opB, descrB, B, request, C);
    }
    template <>
    DLL_PUBLIC aoclsparse_status sp2m<float>(aoclsparse_operation opA, const aoclsparse_mat_descr descrA, const aoclsparse_matrix A, aoclsparse_operation opB, const aoclsparse_mat_descr descrB, const aoclsparse_matrix B, aoclsparse_request request, aoclsparse_matrix *C)
    {
        return sp2m_typed<float>(opA, descrA, A, opB, descrB, B, request, C);
    }
    template <>
    DLL_PUBLIC aoclsparse_status sp2m<double>(aoclsparse_operation opA, const aoclsparse_mat_descr descrA, const aoclsparse_matrix A, aoclsparse_operation opB, const aoclsparse_mat_descr descrB, const aoclsparse_matrix B, aoclsparse_request request, aoclsparse_matrix *C)
    {
        return sp2m_typed<double>(opA, descrA, A, opB, descrB, B, request, C);
    }
    template <>
    DLL_PUBLIC aoclsparse_status sp2m<std::complex<float>>(aoclsparse_operation opA, const aoclsparse_mat_descr descrA, const aoclsparse_matrix A, aoclsparse_operation opB, const aoclsparse_mat_descr descrB, const aoclsparse_matrix B, aoclsparse_request request, aoclsparse_matrix *C)
    {
        return sp2m_typed<std::complex<float>>(opA, descrA, A, opB, descrB, B, request, C);
    }
    template <>
    DLL_PUBLIC aoclsparse_status sp2m<std::complex<double>>(aoclsparse_operation opA, const aoclsparse_mat_descr descrA, const aoclsparse_matrix A, aoclsparse_operation opB, const aoclsparse_mat_descr descrB, const aoclsparse_matrix B, aoclsparse_request request, aoclsparse_matrix *C)
    {
        return sp2m_typed<std::complex<double>>(opA, descrA, A, opB, descrB, B, request, C);
    }
}
