/* shard_check.c -- the row-sharded iteration driven from plain C, no Python and no collective library in the loop:
 * one process, `world` shards (on `world` GPUs when the box has them, else several shards share a device), each rank's
 * slab of a 3-D 7-point stencil, x <- A x / 12 iterated; the links are exchanged by copying 320-byte records in memory.
 * Checks (exit code != 0 on any mismatch, like every example of the reference, tests/examples/sample_spmv_c.c:100-109):
 *   1. the gathered result equals, bit for bit, the same iteration on ONE shard (world = 1, plain aoclsparse_dmv);
 *   2. both equal the host evaluation of the recurrence within 1e-12 * iterations * sum|a||x|.
 * Build: gcc -O2 -I include tests/shard_check.c -o shard_check aocl-sparse_b200/libaoclsparse_b200.so -lm */
#include "aoclsparse.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(call)                                                                    \
    do                                                                                 \
    {                                                                                  \
        aoclsparse_status s__ = (call);                                                \
        if(s__ != aoclsparse_status_success)                                           \
        {                                                                              \
            fprintf(stderr, "%s:%d: %s -> status %d (%s)\n", __FILE__, __LINE__, #call, (int)s__, aoclsparse_b200_last_error()); \
            return 2;                                                                  \
        }                                                                              \
    } while(0)

static int nx = 24, ny = 20, nz;

/* rows [lo, hi) of the 7-point stencil, global column indices, columns ascending */
static void slab(long lo, long hi, aoclsparse_int **rp, aoclsparse_int **col, double **val, aoclsparse_int *nnz)
{
    const long plane = (long)nx * ny;
    *rp  = malloc(sizeof(aoclsparse_int) * (size_t)(hi - lo + 1));
    *col = malloc(sizeof(aoclsparse_int) * (size_t)(hi - lo) * 7);
    *val = malloc(sizeof(double) * (size_t)(hi - lo) * 7);
    aoclsparse_int p = 0;
    for(long r = lo; r < hi; ++r)
    {
        const long x = r % nx, y = (r / nx) % ny, z = r / plane;
        (*rp)[r - lo] = p;
        if(z > 0) { (*col)[p] = (aoclsparse_int)(r - plane); (*val)[p++] = -1.0; }
        if(y > 0) { (*col)[p] = (aoclsparse_int)(r - nx); (*val)[p++] = -1.0; }
        if(x > 0) { (*col)[p] = (aoclsparse_int)(r - 1); (*val)[p++] = -1.0; }
        (*col)[p] = (aoclsparse_int)r; (*val)[p++] = 6.0;
        if(x < nx - 1) { (*col)[p] = (aoclsparse_int)(r + 1); (*val)[p++] = -1.0; }
        if(y < ny - 1) { (*col)[p] = (aoclsparse_int)(r + nx); (*val)[p++] = -1.0; }
        if(z < nz - 1) { (*col)[p] = (aoclsparse_int)(r + plane); (*val)[p++] = -1.0; }
    }
    (*rp)[hi - lo] = p;
    *nnz           = p;
}

static int run(int world, int iters, const double *x0, double *out)
{
    const long plane = (long)nx * ny, n = plane * nz;
    int        ndev  = 1;
    CHECK(aoclsparse_b200_device_count(&ndev));
    aoclsparse_matrix     *A = calloc((size_t)world, sizeof(*A));
    aoclsparse_b200_shard *S = calloc((size_t)world, sizeof(*S));
    unsigned char         *links = calloc((size_t)world, AOCLSPARSE_B200_SHARD_LINK_BYTES);
    aoclsparse_mat_descr   descr;
    CHECK(aoclsparse_create_mat_descr(&descr));
    long *lo = malloc(sizeof(long) * (size_t)(world + 1));
    for(int r = 0; r <= world; ++r)
        lo[r] = (long)nz * r / world * plane;
    for(int r = 0; r < world; ++r)
    {
        CHECK(aoclsparse_b200_set_device(r % ndev));
        aoclsparse_int *rp, *col, nnz;
        double         *val;
        slab(lo[r], lo[r + 1], &rp, &col, &val, &nnz);
        CHECK(aoclsparse_create_dcsr(&A[r], aoclsparse_index_base_zero, (aoclsparse_int)(lo[r + 1] - lo[r]), (aoclsparse_int)n, nnz, rp, col, val));
        free(rp); free(col); free(val); /* the handle owns device copies */
        CHECK(aoclsparse_b200_shard_create(&S[r], A[r], descr, r, world, (aoclsparse_int)lo[r], (aoclsparse_int)plane));
        CHECK(aoclsparse_b200_shard_export(S[r], links + (size_t)r * AOCLSPARSE_B200_SHARD_LINK_BYTES));
    }
    for(int r = 0; r < world; ++r)
        CHECK(aoclsparse_b200_shard_connect(S[r], r > 0 ? links + (size_t)(r - 1) * AOCLSPARSE_B200_SHARD_LINK_BYTES : NULL,
                                            r < world - 1 ? links + (size_t)(r + 1) * AOCLSPARSE_B200_SHARD_LINK_BYTES : NULL));
    for(int r = 0; r < world; ++r)
        CHECK(aoclsparse_b200_shard_set_x(S[r], x0 + lo[r]));
    for(int r = 0; r < world; ++r)
        CHECK(aoclsparse_b200_shard_publish(S[r]));
    for(int done = 0; done < iters; done += 5) /* one host thread drives every shard: short turns */
        for(int r = 0; r < world; ++r)
            CHECK(aoclsparse_b200_shard_iterate(S[r], 1.0 / 12.0, iters - done < 5 ? iters - done : 5));
    for(int r = 0; r < world; ++r)
        CHECK(aoclsparse_b200_shard_get_x(S[r], out + lo[r]));
    for(int r = 0; r < world; ++r)
    {
        CHECK(aoclsparse_b200_shard_destroy(&S[r]));
        CHECK(aoclsparse_destroy(&A[r]));
    }
    aoclsparse_destroy_mat_descr(descr);
    free(A); free(S); free(links); free(lo);
    return 0;
}

int main(int argc, char **argv)
{
    const int world = argc > 1 ? atoi(argv[1]) : 2, iters = argc > 2 ? atoi(argv[2]) : 25;
    nz              = 6 * world;
    const long plane = (long)nx * ny, n = plane * nz;
    double    *x0 = malloc(sizeof(double) * (size_t)n), *sharded = malloc(sizeof(double) * (size_t)n),
           *single = malloc(sizeof(double) * (size_t)n), *host = malloc(sizeof(double) * (size_t)n),
           *scale = malloc(sizeof(double) * (size_t)n), *tmp = malloc(sizeof(double) * (size_t)n);
    unsigned long long s = 88172645463325252ULL;
    for(long i = 0; i < n; ++i)
    {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        x0[i] = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    }
    if(run(world, iters, x0, sharded) || run(1, iters, x0, single))
        return 2;
    /* host evaluation of the recurrence and of its error scale */
    memcpy(host, x0, sizeof(double) * (size_t)n);
    for(long i = 0; i < n; ++i)
        scale[i] = fabs(x0[i]);
    for(int it = 0; it < iters; ++it)
    {
        for(int pass = 0; pass < 2; ++pass)
        {
            double *v = pass ? scale : host;
            for(long r = 0; r < n; ++r)
            {
                const long x = r % nx, y = (r / nx) % ny, z = r / plane;
                const double off = pass ? 1.0 : -1.0;
                double       acc = 0.0;
                if(z > 0) acc += off * v[r - plane];
                if(y > 0) acc += off * v[r - nx];
                if(x > 0) acc += off * v[r - 1];
                acc += 6.0 * v[r];
                if(x < nx - 1) acc += off * v[r + 1];
                if(y < ny - 1) acc += off * v[r + nx];
                if(z < nz - 1) acc += off * v[r + plane];
                tmp[r] = acc / 12.0;
            }
            memcpy(v, tmp, sizeof(double) * (size_t)n);
        }
    }
    long   mism = 0;
    double worst = 0.0;
    for(long i = 0; i < n; ++i)
    {
        if(memcmp(&sharded[i], &single[i], sizeof(double)) != 0)
            ++mism;
        const double e = fabs(sharded[i] - host[i]) / scale[i];
        if(e > worst)
            worst = e;
    }
    printf("world %d, %d iterations, %ld rows: %ld entries differ from the one-shard run, worst error vs host %.3e\n",
           world, iters, n, mism, worst);
    if(mism != 0 || !(worst <= 1e-12 * iters))
    {
        printf("SHARD_CHECK_FAILED\n");
        return 1;
    }
    printf("SHARD_CHECK_OK\n");
    return 0;
}
