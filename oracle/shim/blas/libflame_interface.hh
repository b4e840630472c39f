/* TEST INFRASTRUCTURE ONLY.  Stand-in for AOCL-libFLAME's "libflame_interface.hh" (external, not vendored), included by
 * library/src/extra/aoclsparse_lapack.hpp:38.  The reference uses one routine from it, libflame::lartg (generation of a
 * plane rotation), and only inside GMRES (aoclsparse_itsol_functions.hpp:1151-1170).  GMRES is not used as an oracle; the
 * routine is stated in its textbook form so that the solvers translation unit links. */
#ifndef ORACLE_SHIM_LIBFLAME_HH
#define ORACLE_SHIM_LIBFLAME_HH

#include <cmath>
#include <complex>

typedef int integer;
typedef struct
{
    float real, imag;
} scomplex;
typedef struct
{
    double real, imag;
} dcomplex;

namespace libflame
{
    /* [c s; -s c] [f; g] = [r; 0] */
    template <typename T>
    inline void lartg(T *f, T *g, T *c, T *s, T *r)
    {
        if(*g == T(0))
        {
            *c = T(1);
            *s = T(0);
            *r = *f;
            return;
        }
        if(*f == T(0))
        {
            *c = T(0);
            *s = T(1);
            *r = *g;
            return;
        }
        const T h = std::hypot(*f, *g);
        *r        = std::copysign(h, *f);
        *c        = *f / *r;
        *s        = *g / *r;
    }
    template <typename CT, typename R>
    inline void lartg(CT *f, CT *g, R *c, CT *s, CT *r)
    {
        const std::complex<R> ff(f->real, f->imag), gg(g->real, g->imag);
        const R               af = std::abs(ff), ag = std::abs(gg);
        if(ag == R(0))
        {
            *c = R(1);
            s->real = s->imag = R(0);
            *r                = *f;
            return;
        }
        if(af == R(0))
        {
            *c = R(0);
            const std::complex<R> sv = std::conj(gg) / ag;
            s->real = sv.real(), s->imag = sv.imag();
            r->real = ag, r->imag = R(0);
            return;
        }
        const R               h  = std::hypot(af, ag);
        const std::complex<R> ph = ff / af;
        *c                       = af / h;
        const std::complex<R> sv = ph * std::conj(gg) / h;
        const std::complex<R> rv = ph * h;
        s->real = sv.real(), s->imag = sv.imag();
        r->real = rv.real(), r->imag = rv.imag();
    }
}
#endif
