/* C++ front ends of the path, for callers written against the reference's aoclsparse.hpp
 * (library/include/aoclsparse.hpp:85-147).  In the reference these templates are instantiated inside the shared
 * library for float, double, std::complex<float> and std::complex<double> (library/src/level2/aoclsparse_mv.cpp:351-360,
 * library/src/create/aoclsparse_create.cpp:97-109, library/src/level3/aoclsparse_csr2m.cpp:862-873); this library exports
 * the same twelve symbols (same mangled names, checked by tests/test_abi.py against tests/golden/cxx_symbols.json), each
 * forwarding to the C entry of the matching precision. */
#ifndef AOCLSPARSE_HPP_B200
#define AOCLSPARSE_HPP_B200

#include "aoclsparse.h"

#include <complex>

namespace aoclsparse
{
    /* y = alpha op(A) x + beta y -- aoclsparse_{s,d,c,z}mv */
    template <typename T>
    aoclsparse_status mv(aoclsparse_operation       op,
                         const T                   *alpha,
                         aoclsparse_matrix          A,
                         const aoclsparse_mat_descr descr,
                         const T                   *x,
                         const T                   *beta,
                         T                         *y);

    /* aoclsparse_create_{s,d,c,z}csr.  fast_chck (skip the full validation in the reference) is accepted and ignored:
     * the validation here is one GPU pass over arrays that are being uploaded anyway. */
    template <typename T>
    aoclsparse_status create_csr(aoclsparse_matrix    *mat,
                                 aoclsparse_index_base base,
                                 aoclsparse_int        M,
                                 aoclsparse_int        N,
                                 aoclsparse_int        nnz,
                                 aoclsparse_int       *row_ptr,
                                 aoclsparse_int       *col_idx,
                                 T                    *val,
                                 bool                  fast_chck = false);

    /* C = op(A) op(B), sparse x sparse -- aoclsparse_sp2m; wrong_type unless A and B hold values of type T */
    template <typename T>
    aoclsparse_status sp2m(aoclsparse_operation       opA,
                           const aoclsparse_mat_descr descrA,
                           const aoclsparse_matrix    A,
                           aoclsparse_operation       opB,
                           const aoclsparse_mat_descr descrB,
                           const aoclsparse_matrix    B,
                           aoclsparse_request         request,
                           aoclsparse_matrix         *C);
}

#endif
