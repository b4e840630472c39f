// Compile + link check for include/aoclsparse.hpp (run by tests/test_abi.py): a caller written against the reference's
// C++ front ends must link against libaoclsparse_b200.so unchanged.  Nothing here touches the GPU: the calls fail
// argument validation (NULL handles) before any device work.
#include "aoclsparse.hpp"

#include <complex>
#include <cstdio>

int main()
{
    double                    one = 1.0, x[1] = {0}, y[1] = {0};
    std::complex<float>       cone(1.0f, 0.0f), cx[1], cy[1];
    aoclsparse_matrix         C = nullptr;
    const aoclsparse_status   s1 = aoclsparse::mv<double>(aoclsparse_operation_none, &one, nullptr, nullptr, x, &one, y);
    const aoclsparse_status   s2 = aoclsparse::mv<std::complex<float>>(aoclsparse_operation_none, &cone, nullptr, nullptr, cx, &cone, cy);
    const aoclsparse_status   s3 = aoclsparse::create_csr<float>(nullptr, aoclsparse_index_base_zero, 1, 1, 1, nullptr, nullptr, (float *)nullptr);
    const aoclsparse_status   s4 = aoclsparse::sp2m<std::complex<double>>(aoclsparse_operation_none, nullptr, nullptr, aoclsparse_operation_none, nullptr, nullptr, aoclsparse_stage_full_computation, &C);
    std::printf("%d %d %d %d\n", (int)s1, (int)s2, (int)s3, (int)s4);
    return (s1 == aoclsparse_status_invalid_pointer && s2 == aoclsparse_status_invalid_pointer
            && s3 == aoclsparse_status_invalid_pointer && s4 == aoclsparse_status_invalid_pointer)
               ? 0
               : 1;
}
