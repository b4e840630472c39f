/* TEST INFRASTRUCTURE ONLY.  Stand-in for AOCL-BLIS's C++ header "cblas.hh", which the reference's iterative
 * solvers include through library/src/extra/aoclsparse_lapack.hpp:37 and which is not vendored in the reference tree
 * (AOCL-BLIS is an external dependency found through AOCL_ROOT, cmake/Dependencies.cmake:139-154).
 *
 * Only what library/src/solvers/aoclsparse_itsol_functions.hpp names is declared.  The conjugate-gradient loop -- the
 * one solver used as a parity oracle here -- calls exactly one BLAS routine, blis::cblas_nrm2 (:679,704,843); it is
 * written out below as the textbook two-norm.  dot / dotc / axpby / scal are only referenced by GMRES, which is
 * compiled so that the library links but is NOT used as an oracle.
 */
#ifndef ORACLE_SHIM_CBLAS_HH
#define ORACLE_SHIM_CBLAS_HH

#include <cmath>
#include <complex>

typedef int f77_int;

namespace blis
{
    template <typename T>
    inline T cblas_nrm2(f77_int n, const T *x, f77_int incx)
    {
        long double s = 0;
        for(f77_int i = 0; i < n; ++i)
            s += (long double)x[i * incx] * (long double)x[i * incx];
        return (T)std::sqrt(s);
    }
    template <typename R>
    inline R cblas_nrm2(f77_int n, const std::complex<R> *x, f77_int incx)
    {
        long double s = 0;
        for(f77_int i = 0; i < n; ++i)
            s += (long double)std::norm(x[i * incx]);
        return (R)std::sqrt(s);
    }
    template <typename T>
    inline T cblas_dot(f77_int n, const T *x, f77_int incx, const T *y, f77_int incy)
    {
        T s = T(0);
        for(f77_int i = 0; i < n; ++i)
            s += x[i * incx] * y[i * incy];
        return s;
    }
    template <typename T>
    inline T cblas_dotc(f77_int n, const T *x, f77_int incx, const T *y, f77_int incy)
    {
        T s = T(0);
        for(f77_int i = 0; i < n; ++i)
            s += std::conj(x[i * incx]) * y[i * incy];
        return s;
    }
    template <typename T>
    struct same
    {
        typedef T type;
    };
    template <typename T>
    inline void cblas_axpby(
        f77_int n, typename same<T>::type alpha, const T *x, f77_int incx, typename same<T>::type beta, T *y, f77_int incy)
    {
        for(f77_int i = 0; i < n; ++i)
            y[i * incy] = alpha * x[i * incx] + beta * y[i * incy];
    }
    template <typename T, typename S>
    inline void cblas_scal(f77_int n, S alpha, T *x, f77_int incx)
    {
        for(f77_int i = 0; i < n; ++i)
            x[i * incx] = x[i * incx] * T(alpha);
    }
}
#endif
