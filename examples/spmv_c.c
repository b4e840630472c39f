/* A C caller of the kept API, written the way a user of the reference writes one (create handle and descriptor, hint,
 * optimize, multiply, destroy) -- the same call sequence as the reference's tests/examples/sample_spmv_c.c, on the
 * matrix and vectors of that sample, so the expected result is the sample's: y = {9, 6, 12, 69, 40}.
 * Builds against include/aoclsparse.h and links against libaoclsparse_b200.so without any change a reference user
 * would not also make for libaoclsparse.so:
 *     cc examples/spmv_c.c -Iinclude -Laocl-sparse_b200 -laoclsparse_b200 -Wl,-rpath,$PWD/aocl-sparse_b200 -o spmv_c
 * Host arrays go in, the result comes back on the host; the product itself runs on the GPU.
 */
#include "aoclsparse.h"

#include <stdio.h>

int main(void)
{
    aoclsparse_int    row_ptr[] = {0, 2, 3, 4, 7, 8};
    aoclsparse_int    col_idx[] = {0, 3, 1, 2, 1, 3, 4, 4};
    double            val[]     = {1, 2, 3, 4, 5, 6, 7, 8};
    double            x[]       = {1, 2, 3, 4, 5};
    double            y[5]      = {0, 0, 0, 0, 0};
    const double      want[5]   = {9, 6, 12, 69, 40};
    const double      alpha = 1.0, beta = 0.0;
    aoclsparse_matrix A     = NULL;
    aoclsparse_mat_descr descr = NULL;
    aoclsparse_status    st;

    printf("%s\n", aoclsparse_get_version());
    if((st = aoclsparse_create_dcsr(&A, aoclsparse_index_base_zero, 5, 5, 8, row_ptr, col_idx, val)) != aoclsparse_status_success)
        return printf("create failed: %d\n", (int)st), 1;
    if((st = aoclsparse_create_mat_descr(&descr)) != aoclsparse_status_success)
        return printf("descr failed: %d\n", (int)st), 1;
    aoclsparse_set_mat_index_base(descr, aoclsparse_index_base_zero);
    if((st = aoclsparse_set_mv_hint(A, aoclsparse_operation_none, descr, 1)) != aoclsparse_status_success)
        return printf("hint failed: %d\n", (int)st), 1;
    if((st = aoclsparse_optimize(A)) != aoclsparse_status_success)
        return printf("optimize failed: %d\n", (int)st), 1;
    if((st = aoclsparse_dmv(aoclsparse_operation_none, &alpha, A, descr, x, &beta, y)) != aoclsparse_status_success)
        return printf("mv failed: %d\n", (int)st), 1;
    int bad = 0;
    for(int i = 0; i < 5; ++i)
    {
        printf("y[%d] = %g\n", i, y[i]);
        bad += y[i] != want[i];
    }
    aoclsparse_destroy_mat_descr(descr);
    aoclsparse_destroy(&A);
    return bad ? 2 : 0;
}
