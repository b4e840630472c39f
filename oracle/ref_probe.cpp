// ref_probe.cpp -- TEST INFRASTRUCTURE ONLY (fixture generation in this container).
//
// The reference keeps some integer facts of the path inside its opaque handle and only its white-box
// unit tests read them (createcsr_tests.cpp:77-99 reads A->sort / A->fulldiag; get_doid is an inline
// template).  This probe is compiled AGAINST the reference's own internal headers, where they lie
// under /root/reference, and linked to oracle/_ref/libaoclsparse_ref.so, so that
// tests/golden/make_golden.py can record those facts as golden fixtures.  Contains no reference code.
#include "aoclsparse.h"
#include "aoclsparse_descr.h"
#include "aoclsparse_mat_structures.hpp"
#include "aoclsparse_mtx_dispatcher.hpp"

#include <complex>

extern "C" {

__attribute__((visibility("default"))) int probe_matrix_facts(aoclsparse_matrix A, int *sort, int *fulldiag, int *n_mats)
{
    if(!A)
        return -1;
    *sort     = (int)A->sort;
    *fulldiag = A->fulldiag ? 1 : 0;
    *n_mats   = (int)A->mats.size();
    return 0;
}

// clean CSR produced by aoclsparse_optimize / csr_csc_optimize: the first csr in A->mats with is_optimized set
// (what tests/unit_tests/hint_tests.cpp:179-197 reads).  Arrays may be NULL to query sizes only.
__attribute__((visibility("default"))) int probe_clean_csr(aoclsparse_matrix A, int *nnz, int *is_internal, int *base,
                                                           int *ptr, int *ind, double *val, int *idiag, int *iurow)
{
    if(!A)
        return -1;
    for(auto *mat : A->mats)
    {
        auto *c = dynamic_cast<aoclsparse::csr *>(mat);
        if(c && c->is_optimized)
        {
            const int nz = c->ptr[A->m] - (int)c->base;
            *nnz         = nz;
            *is_internal = (mat != A->mats[0]) ? 1 : 0; // a copy was made (hint_tests.cpp opt_csr_is_internal)
            *base        = (int)c->base;
            if(ptr)
                for(int i = 0; i <= A->m; ++i)
                    ptr[i] = c->ptr[i];
            if(ind)
                for(int i = 0; i < nz; ++i)
                    ind[i] = c->ind[i];
            if(val)
                for(int i = 0; i < nz; ++i)
                    val[i] = ((double *)c->val)[i];
            if(idiag)
                for(int i = 0; i < A->m; ++i)
                    idiag[i] = c->idiag[i];
            if(iurow)
                for(int i = 0; i < A->m; ++i)
                    iurow[i] = c->iurow[i];
            return 0;
        }
    }
    return 1;
}

__attribute__((visibility("default"))) int probe_get_doid(int is_complex, int type, int fill, int op)
{
    _aoclsparse_mat_descr d;
    d.type      = (aoclsparse_matrix_type)type;
    d.fill_mode = (aoclsparse_fill_mode)fill;
    d.diag_type = aoclsparse_diag_type_non_unit;
    d.base      = aoclsparse_index_base_zero;
    if(is_complex)
        return (int)aoclsparse::get_doid<std::complex<double>>(&d, (aoclsparse_operation)op);
    return (int)aoclsparse::get_doid<double>(&d, (aoclsparse_operation)op);
}

__attribute__((visibility("default"))) int probe_effective_doid(int mat, int req)
{
    return (int)aoclsparse::get_effective_doid((aoclsparse::doid)mat, (aoclsparse::doid)req);
}
}
