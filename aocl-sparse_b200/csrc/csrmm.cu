// csrmm.cu -- C = alpha * op(A) * B + beta * C, A sparse CSR on the device, B and C dense.
//
// Replaces (front end)  aoclsparse_csrmm_t<T>            library/src/level3/aoclsparse_csrmm.hpp:429-838
//          (kernels)    csrmm_row_kt / csrmm_col_kt       library/src/level3/aoclsparse_csrmm_kt.cpp:31-363
//                       aoclsparse_csrmm_{row,col}_major_ref, *_sym_*_ref, scale_dense_matrix
//                                                         library/src/level3/aoclsparse_csrmm.hpp:36-427
//
// Decomposition (not the reference's): the row blocks of the SpMV plan are reused, so one CTA again
// streams a fixed-size slice of val/col into shared memory with TMA bulk copies.
//   row-major   : one warp per row, lanes own columns of B/C, so every non-zero turns into one
//                 coalesced read of a B row segment (n = 32 doubles: one 256-byte line pair) and the
//                 C row is written once, coalesced;
//   column-major: one thread per row (lanes = consecutive rows, B column gathers coalesce exactly as x
//                 does in SpMV), a register tile of 4 columns of B per pass over the staged row.
// Rows split across CTAs (longer than the block capacity) accumulate with atomics into a pre-scaled C.
// beta == 0 overwrites C without reading it.
#include "spmv_kernels.cuh"

#include <cstdint>
#include <cstdlib>

namespace b200
{
    namespace
    {
        constexpr int MM_THREADS = 256;
        constexpr int COL_TILE   = 8;

        // ROW_MAJOR: B is (k x n) with row stride ldb, C is (m x n) with row stride ldc
        template <typename T, bool CONJ>
        __global__ void __launch_bounds__(MM_THREADS) csrmm_row_major_kernel(const int4 *__restrict__ desc,
                                                                            const int *__restrict__ kind,
                                                                            int cap,
                                                                            const aoclsparse_int *__restrict__ rp,
                                                                            const aoclsparse_int *__restrict__ col,
                                                                            const T *__restrict__ val,
                                                                            const T *__restrict__ B,
                                                                            long long ldb,
                                                                            T *__restrict__ C,
                                                                            long long ldc,
                                                                            int       n,
                                                                            T         alpha,
                                                                            T         beta,
                                                                            int       beta_zero)
        {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            uint64_t       *bar  = reinterpret_cast<uint64_t *>(smem_raw);
            T              *sval = reinterpret_cast<T *>(smem_raw + SMEM_HEADER);
            aoclsparse_int *scol = reinterpret_cast<aoclsparse_int *>(smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T));

            const int  tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            const int4 d     = desc[blockIdx.x];
            const int  strat = kind[blockIdx.x] & 15;
            const int  ns = d.z, ne = d.w;
            const int  a   = ns & ~3;
            const int  cnt = ((ne - a) + 3) & ~3;
            if(tid == 0)
            {
                mbar_init(bar, 1);
                mbar_init_fence();
                if(cnt > 0)
                {
                    mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                    bulk_load(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                    bulk_load(scol, col + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar);
                }
            }
            __syncthreads();
            if(cnt > 0)
                mbar_wait(bar, 0);

            if(strat != STRAT_LONG)
            {
                for(int r = d.x + warp; r < d.y; r += MM_THREADS / 32)
                {
                    const int s = rp[r] - a, e = rp[r + 1] - a;
                    for(int c0 = 0; c0 < n; c0 += 128)
                    {
                        // lanes own columns c0+lane+32q of B / C; 4 non-zeros (4 independent B-row reads per
                        // owned column) are in flight before the first multiply-add retires
                        T acc[4];
#pragma unroll
                        for(int q = 0; q < 4; ++q)
                            acc[q] = vt<T>::zero();
                        const bool q1 = c0 + lane + 32 < n, q2 = c0 + lane + 64 < n, q3 = c0 + lane + 96 < n;
                        const bool q0 = c0 + lane < n;
                        int        j  = s;
                        for(; j + 4 <= e; j += 4)
                        {
                            T        v[4];
                            const T *br[4];
#pragma unroll
                            for(int u = 0; u < 4; ++u)
                            {
                                v[u] = sval[j + u];
                                if(CONJ)
                                    v[u] = cj(v[u]);
                                br[u] = B + (long long)scol[j + u] * ldb + c0 + lane;
                            }
                            T b0[4], b1[4], b2[4], b3[4];
#pragma unroll
                            for(int u = 0; u < 4; ++u)
                            {
                                b0[u] = q0 ? ldg_ro(br[u]) : vt<T>::zero();
                                b1[u] = q1 ? ldg_ro(br[u] + 32) : vt<T>::zero();
                                b2[u] = q2 ? ldg_ro(br[u] + 64) : vt<T>::zero();
                                b3[u] = q3 ? ldg_ro(br[u] + 96) : vt<T>::zero();
                            }
#pragma unroll
                            for(int u = 0; u < 4; ++u)
                            {
                                acc[0] = mad(v[u], b0[u], acc[0]);
                                acc[1] = mad(v[u], b1[u], acc[1]);
                                acc[2] = mad(v[u], b2[u], acc[2]);
                                acc[3] = mad(v[u], b3[u], acc[3]);
                            }
                        }
                        for(; j < e; ++j)
                        {
                            T v = sval[j];
                            if(CONJ)
                                v = cj(v);
                            const T *brow = B + (long long)scol[j] * ldb + c0 + lane;
                            if(q0)
                                acc[0] = mad(v, ldg_ro(brow), acc[0]);
                            if(q1)
                                acc[1] = mad(v, ldg_ro(brow + 32), acc[1]);
                            if(q2)
                                acc[2] = mad(v, ldg_ro(brow + 64), acc[2]);
                            if(q3)
                                acc[3] = mad(v, ldg_ro(brow + 96), acc[3]);
                        }
                        T *crow = C + (long long)r * ldc + c0 + lane;
#pragma unroll
                        for(int q = 0; q < 4; ++q)
                            if(c0 + lane + 32 * q < n)
                                crow[32 * q] = axpby_out(alpha, acc[q], beta, beta_zero != 0, crow + 32 * q);
                    }
                }
            }
            else
            {
                // one segment of a long row: warps split the entries, results are added atomically to
                // the row of C that scale_dense_kernel already multiplied by beta
                const int r = d.x, first = ns - a, total = ne - ns;
                for(int c0 = 0; c0 < n; c0 += 32)
                {
                    T acc = vt<T>::zero();
                    if(c0 + lane < n)
                        for(int i = warp; i < total; i += MM_THREADS / 32)
                        {
                            T v = sval[first + i];
                            if(CONJ)
                                v = cj(v);
                            acc = mad(v, ldg_ro(B + (long long)scol[first + i] * ldb + c0 + lane), acc);
                        }
                    if(c0 + lane < n)
                        atomic_accumulate(C + (long long)r * ldc + c0 + lane, mul(alpha, acc));
                }
            }
        }

        // 16-byte vectors of T
        template <typename T>
        struct vec16
        {
            static constexpr int N = 16 / sizeof(T);
            T                    v[N];
        };
        template <typename T>
        __device__ __forceinline__ vec16<T> load_vec(const T *p)
        {
            const int4 raw = __ldg(reinterpret_cast<const int4 *>(p));
            vec16<T>   r;
            memcpy(&r, &raw, 16);
            return r;
        }
        template <typename T>
        __device__ __forceinline__ void store_vec(T *p, const vec16<T> &v)
        {
            int4 raw;
            memcpy(&raw, &v, 16);
            *reinterpret_cast<int4 *>(p) = raw;
        }

        // ROW MAJOR, vectorised: LPR lanes share one row of A and together own LPR*2 16-byte vectors of the
        // B / C row (n = 32 doubles: 8 lanes x 2 vectors x 2 doubles), so a warp advances 32/LPR rows of A at a
        // time and every non-zero costs two 128-bit loads per lane instead of one 64-bit load per column.
        // Needs 16-byte aligned B / C rows and n a multiple of the vector width (checked by the launcher).
        template <typename T, bool CONJ, int LPR, int U>
        __global__ void __launch_bounds__(MM_THREADS) csrmm_row_major_vec_kernel(const int4 *__restrict__ desc,
                                                                                const int *__restrict__ kind,
                                                                                int cap,
                                                                                const aoclsparse_int *__restrict__ rp,
                                                                                const aoclsparse_int *__restrict__ col,
                                                                                const T *__restrict__ val,
                                                                                const T *__restrict__ B,
                                                                                long long ldb,
                                                                                T *__restrict__ C,
                                                                                long long ldc,
                                                                                int       n,
                                                                                T         alpha,
                                                                                T         beta,
                                                                                int       beta_zero)
        {
            constexpr int VEC = vec16<T>::N;
            constexpr int RPW = 32 / LPR;        // rows of A per warp pass
            constexpr int CPP = LPR * 2 * VEC;   // columns of B per pass
            extern __shared__ __align__(16) unsigned char smem_raw[];
            uint64_t       *bar  = reinterpret_cast<uint64_t *>(smem_raw);
            T              *sval = reinterpret_cast<T *>(smem_raw + SMEM_HEADER);
            aoclsparse_int *scol = reinterpret_cast<aoclsparse_int *>(smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T));

            const int  tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            const int  sub = lane / LPR, sl = lane % LPR;
            const int4 d     = desc[blockIdx.x];
            const int  strat = kind[blockIdx.x] & 15;
            const int  ns = d.z, ne = d.w;
            const int  a   = ns & ~3;
            const int  cnt = ((ne - a) + 3) & ~3;
            if(tid == 0)
            {
                mbar_init(bar, 1);
                mbar_init_fence();
                if(cnt > 0)
                {
                    mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                    bulk_load_stream(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                    bulk_load_stream(scol, col + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar);
                }
            }
            __syncthreads();
            if(cnt > 0)
                mbar_wait(bar, 0);

            if(strat != STRAT_LONG)
            {
                for(int rb = d.x + warp * RPW; rb < d.y; rb += (MM_THREADS / 32) * RPW)
                {
                    const int  r     = rb + sub;
                    const bool valid = r < d.y;
                    int        s = 0, e = 0;
                    if(valid)
                    {
                        s = rp[r] - a;
                        e = rp[r + 1] - a;
                    }
                    for(int c0 = 0; c0 < n; c0 += CPP)
                    {
                        const int  ca = c0 + sl * VEC, cb = c0 + (LPR + sl) * VEC;
                        const bool pa = ca < n, pb = cb < n;
                        vec16<T>   acc_a, acc_b;
#pragma unroll
                        for(int q = 0; q < VEC; ++q)
                            acc_a.v[q] = acc_b.v[q] = vt<T>::zero();
                        int j = s;
                        for(; j + U <= e; j += U)
                        {
                            T        v[U];
                            const T *bp[U];
#pragma unroll
                            for(int u = 0; u < U; ++u)
                            {
                                v[u] = sval[j + u];
                                if(CONJ)
                                    v[u] = cj(v[u]);
                                bp[u] = B + (long long)scol[j + u] * ldb;
                            }
                            vec16<T> xa[U], xb[U];
                            if(pa)
                            {
#pragma unroll
                                for(int u = 0; u < U; ++u)
                                    xa[u] = load_vec(bp[u] + ca);
                            }
                            if(pb)
                            {
#pragma unroll
                                for(int u = 0; u < U; ++u)
                                    xb[u] = load_vec(bp[u] + cb);
                            }
                            if(pa)
                            {
#pragma unroll
                                for(int u = 0; u < U; ++u)
#pragma unroll
                                    for(int q = 0; q < VEC; ++q)
                                        acc_a.v[q] = mad(v[u], xa[u].v[q], acc_a.v[q]);
                            }
                            if(pb)
                            {
#pragma unroll
                                for(int u = 0; u < U; ++u)
#pragma unroll
                                    for(int q = 0; q < VEC; ++q)
                                        acc_b.v[q] = mad(v[u], xb[u].v[q], acc_b.v[q]);
                            }
                        }
                        for(; j < e; ++j)
                        {
                            T v0 = sval[j];
                            if(CONJ)
                                v0 = cj(v0);
                            const T *b0 = B + (long long)scol[j] * ldb;
                            if(pa)
                            {
                                const vec16<T> x0a = load_vec(b0 + ca);
#pragma unroll
                                for(int q = 0; q < VEC; ++q)
                                    acc_a.v[q] = mad(v0, x0a.v[q], acc_a.v[q]);
                            }
                            if(pb)
                            {
                                const vec16<T> x0b = load_vec(b0 + cb);
#pragma unroll
                                for(int q = 0; q < VEC; ++q)
                                    acc_b.v[q] = mad(v0, x0b.v[q], acc_b.v[q]);
                            }
                        }
                        if(valid)
                        {
                            T *crow = C + (long long)r * ldc;
                            if(pa)
                            {
                                vec16<T> o;
                                if(!beta_zero)
                                    o = load_vec(crow + ca);
#pragma unroll
                                for(int q = 0; q < VEC; ++q)
                                    o.v[q] = axpby_out(alpha, acc_a.v[q], beta, beta_zero != 0, &o.v[q]);
                                store_vec(crow + ca, o);
                            }
                            if(pb)
                            {
                                vec16<T> o;
                                if(!beta_zero)
                                    o = load_vec(crow + cb);
#pragma unroll
                                for(int q = 0; q < VEC; ++q)
                                    o.v[q] = axpby_out(alpha, acc_b.v[q], beta, beta_zero != 0, &o.v[q]);
                                store_vec(crow + cb, o);
                            }
                        }
                    }
                }
            }
            else
            {
                const int r = d.x, first = ns - a, total = ne - ns;
                for(int c0 = 0; c0 < n; c0 += 32)
                {
                    T acc = vt<T>::zero();
                    if(c0 + lane < n)
                        for(int i = warp; i < total; i += MM_THREADS / 32)
                        {
                            T v = sval[first + i];
                            if(CONJ)
                                v = cj(v);
                            acc = mad(v, ldg_ro(B + (long long)scol[first + i] * ldb + c0 + lane), acc);
                        }
                    if(c0 + lane < n)
                        atomic_accumulate(C + (long long)r * ldc + c0 + lane, mul(alpha, acc));
                }
            }
        }

        // COLUMN MAJOR: B is (k x n) with column stride ldb, C is (m x n) with column stride ldc
        template <typename T, bool CONJ>
        __global__ void __launch_bounds__(MM_THREADS) csrmm_col_major_kernel(const int4 *__restrict__ desc,
                                                                            const int *__restrict__ kind,
                                                                            int cap,
                                                                            const aoclsparse_int *__restrict__ rp,
                                                                            const aoclsparse_int *__restrict__ col,
                                                                            const T *__restrict__ val,
                                                                            const T *__restrict__ B,
                                                                            long long ldb,
                                                                            T *__restrict__ C,
                                                                            long long ldc,
                                                                            int       n,
                                                                            T         alpha,
                                                                            T         beta,
                                                                            int       beta_zero)
        {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            uint64_t       *bar  = reinterpret_cast<uint64_t *>(smem_raw);
            T              *sval = reinterpret_cast<T *>(smem_raw + SMEM_HEADER);
            aoclsparse_int *scol = reinterpret_cast<aoclsparse_int *>(smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T));

            const int  tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            const int4 d     = desc[blockIdx.x];
            const int  strat = kind[blockIdx.x] & 15;
            const int  ns = d.z, ne = d.w;
            const int  a   = ns & ~3;
            const int  cnt = ((ne - a) + 3) & ~3;
            if(tid == 0)
            {
                mbar_init(bar, 1);
                mbar_init_fence();
                if(cnt > 0)
                {
                    mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                    bulk_load(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                    bulk_load(scol, col + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar);
                }
            }
            __syncthreads();
            if(cnt > 0)
                mbar_wait(bar, 0);

            if(strat == STRAT_THREAD)
            {
                // work item = (row, tile of COL_TILE columns of B / C); consecutive threads take consecutive rows of the
                // same tile, so the gathers B[col + j*ldb] of a warp fall into a few lines (banded matrices) and every
                // thread of the CTA is busy even when the block holds fewer rows than threads
                const int nrows = d.y - d.x;
                const int tiles = (n + COL_TILE - 1) / COL_TILE;
                for(int item = tid; item < nrows * tiles; item += MM_THREADS)
                {
                    const int r  = d.x + item % nrows;
                    const int j0 = (item / nrows) * COL_TILE;
                    const int s = rp[r] - a, e = rp[r + 1] - a;
                    T         acc[COL_TILE];
#pragma unroll
                    for(int q = 0; q < COL_TILE; ++q)
                        acc[q] = vt<T>::zero();
                    if(j0 + COL_TILE <= n)
                    {
                        for(int j = s; j < e; ++j)
                        {
                            T v = sval[j];
                            if(CONJ)
                                v = cj(v);
                            const T *bp = B + scol[j] + (long long)j0 * ldb;
#pragma unroll
                            for(int q = 0; q < COL_TILE; ++q)
                                acc[q] = mad(v, ldg_ro(bp + (long long)q * ldb), acc[q]);
                        }
                    }
                    else
                    {
                        for(int j = s; j < e; ++j)
                        {
                            T v = sval[j];
                            if(CONJ)
                                v = cj(v);
                            const T *bp = B + scol[j] + (long long)j0 * ldb;
#pragma unroll
                            for(int q = 0; q < COL_TILE; ++q)
                                if(j0 + q < n)
                                    acc[q] = mad(v, ldg_ro(bp + (long long)q * ldb), acc[q]);
                        }
                    }
#pragma unroll
                    for(int q = 0; q < COL_TILE; ++q)
                        if(j0 + q < n)
                        {
                            T *cp = C + r + (long long)(j0 + q) * ldc;
                            *cp   = axpby_out(alpha, acc[q], beta, beta_zero != 0, cp);
                        }
                }
            }
            else if(strat != STRAT_LONG)
            {
                // longer / irregular rows: one warp per row, lanes stride the entries, one column at a time
                for(int r = d.x + warp; r < d.y; r += MM_THREADS / 32)
                {
                    const int s = rp[r] - a, e = rp[r + 1] - a;
                    for(int j0 = 0; j0 < n; ++j0)
                    {
                        T acc = vt<T>::zero();
                        for(int j = s + lane; j < e; j += 32)
                        {
                            T v = sval[j];
                            if(CONJ)
                                v = cj(v);
                            acc = mad(v, ldg_ro(B + scol[j] + (long long)j0 * ldb), acc);
                        }
                        acc = warp_sum(acc);
                        if(lane == 0)
                        {
                            T *cp = C + r + (long long)j0 * ldc;
                            *cp   = axpby_out(alpha, acc, beta, beta_zero != 0, cp);
                        }
                    }
                }
            }
            else
            {
                const int r = d.x, first = ns - a, total = ne - ns;
                for(int j0 = warp; j0 < n; j0 += MM_THREADS / 32)
                {
                    T acc = vt<T>::zero();
                    for(int i = lane; i < total; i += 32)
                    {
                        T v = sval[first + i];
                        if(CONJ)
                            v = cj(v);
                        acc = mad(v, ldg_ro(B + scol[first + i] + (long long)j0 * ldb), acc);
                    }
                    acc = warp_sum(acc);
                    if(lane == 0)
                        atomic_accumulate(C + r + (long long)j0 * ldc, mul(alpha, acc));
                }
            }
        }

        // C(rows x cols, leading dimension ld) *= beta (beta == 0: zero fill); padding untouched.
        // only_rows != nullptr restricts the operation to the listed rows (long rows of the plan).
        template <typename T>
        __global__ void scale_dense_kernel(int        row_major,
                                           long long  rows,
                                           long long  cols,
                                           long long  ld,
                                           T         *C,
                                           T          beta,
                                           int        beta_zero,
                                           const int4 *only_rows,
                                           int        n_only)
        {
            const long long total  = (only_rows ? (long long)n_only : rows) * cols;
            long long       i      = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            const long long stride = (long long)gridDim.x * blockDim.x;
            for(; i < total; i += stride)
            {
                long long r, c;
                if(row_major || only_rows)
                {
                    r = i / cols;
                    c = i % cols;
                }
                else
                {
                    c = i / rows;
                    r = i % rows;
                }
                if(only_rows)
                    r = only_rows[r].x;
                T *p = row_major ? C + r * ld + c : C + r + c * ld;
                *p   = beta_zero ? vt<T>::zero() : mul(beta, *p);
            }
        }

        template <typename T>
        aoclsparse_status scale_dense(aoclsparse_order order, T *C, long long rows, long long cols, long long ld, T beta, cudaStream_t st)
        {
            const long long total = rows * cols;
            if(total <= 0)
                return aoclsparse_status_success;
            long long blocks = (total + 255) / 256;
            if(blocks > 148 * 32)
                blocks = 148 * 32;
            scale_dense_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(
                order == aoclsparse_order_row ? 1 : 0, rows, cols, ld, C, beta, is_zero(beta) ? 1 : 0, nullptr, 0);
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }

        template <typename T, bool CONJ>
        aoclsparse_status launch_mm(const dev_csr   &A,
                                    aoclsparse_order order,
                                    const T         *B,
                                    long long        ldb,
                                    T               *C,
                                    long long        ldc,
                                    int              n,
                                    T                alpha,
                                    T                beta,
                                    cudaStream_t     st)
        {
            const row_block_plan &P = A.plan;
            if(P.n_blocks <= 0)
                return aoclsparse_status_success;
            const int    cap  = P.block_nnz + 8;
            const size_t smem = spmv_smem_bytes(sizeof(T), P.block_nnz);
            const int    bz   = is_zero(beta) ? 1 : 0;
            if(P.n_long_rows > 0)
            {
                const long long total  = (long long)P.n_long_rows * n;
                long long       blocks = (total + 255) / 256;
                if(blocks > 148 * 32)
                    blocks = 148 * 32;
                scale_dense_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(order == aoclsparse_order_row ? 1 : 0,
                                                                         A.m,
                                                                         n,
                                                                         ldc,
                                                                         C,
                                                                         beta,
                                                                         bz,
                                                                         P.long_rows.as<int4>(),
                                                                         P.n_long_rows);
                B200_LAUNCHED();
            }
            constexpr int VEC = 16 / (int)sizeof(T);
            const bool    vec_ok = order == aoclsparse_order_row && (n % VEC) == 0 && (ldb % VEC) == 0 && (ldc % VEC) == 0
                                && ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0;
            // stencil / grid matrices: box tiles with their distinct B rows staged once in shared memory (mesh_tiles.cu).
            // AOCLSPARSE_B200_MM_TILES: 0 never, 1 (default) matrices of at least 32768 entries, 2 any size
            const size_t row_bytes = (size_t)n * sizeof(T);
            if(vec_ok && !CONJ && P.n_long_rows == 0 && (row_bytes == 128 || row_bytes == 256 || row_bytes == 512))
            {
                const char *e    = getenv("AOCLSPARSE_B200_MM_TILES");
                const int   mode = e ? atoi(e) : 1;
                if(mode >= 2 || (mode == 1 && A.nnz >= 32768))
                {
                    {
                        std::lock_guard<std::mutex> lk(A.tiles_mu);
                        if(A.tiles.state == 0)
                            B200_TRY(build_mesh_tiles(A, sizeof(T), row_bytes, st));
                    }
                    if(A.tiles.state == 2 && mesh_tiles_smem(A.tiles, row_bytes, sizeof(T)) <= (size_t)227 * 1024)
                        return launch_mm_tiles<T>(A, B, ldb, C, ldc, n, alpha, beta, st);
                }
            }
            if(vec_ok)
            {
                const int vecs = n / VEC; // 16-byte vectors per row; LPR lanes x 2 vectors per pass
                int       lpr  = vecs <= 8 ? 4 : (vecs <= 16 ? 8 : (vecs <= 32 ? 16 : 32));
                static const int env_lpr = getenv("AOCLSPARSE_B200_MM_LPR") ? atoi(getenv("AOCLSPARSE_B200_MM_LPR")) : 0;
                static const int env_u   = getenv("AOCLSPARSE_B200_MM_UNROLL") ? atoi(getenv("AOCLSPARSE_B200_MM_UNROLL")) : 2;
                if(env_lpr == 4 || env_lpr == 8 || env_lpr == 16 || env_lpr == 32)
                    if(env_lpr * 2 >= vecs)
                        lpr = env_lpr;
#define B200_MM_VEC(L, UU)                                                                                               \
    {                                                                                                                \
        static std::atomic<size_t> cfg{0};                                                                           \
        if(cfg.load() < smem)                                                                                        \
        {                                                                                                            \
            B200_CUDA(cudaFuncSetAttribute(                                                                          \
                csrmm_row_major_vec_kernel<T, CONJ, L, UU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            cfg.store(smem);                                                                                         \
        }                                                                                                            \
        csrmm_row_major_vec_kernel<T, CONJ, L, UU><<<P.n_blocks, MM_THREADS, smem, st>>>(P.desc.as<int4>(),        \
                                                                                      P.kind.as<int>(),             \
                                                                                      cap,                          \
                                                                                      A.row_ptr.as<aoclsparse_int>(), \
                                                                                      A.col_idx.as<aoclsparse_int>(), \
                                                                                      A.val.as<T>(),                \
                                                                                      B,                            \
                                                                                      ldb,                          \
                                                                                      C,                            \
                                                                                      ldc,                          \
                                                                                      n,                            \
                                                                                      alpha,                        \
                                                                                      beta,                         \
                                                                                      bz);                          \
    }
                if(env_u == 4)
                {
                    if(lpr == 4)
                        B200_MM_VEC(4, 4)
                    else if(lpr == 8)
                        B200_MM_VEC(8, 4)
                    else if(lpr == 16)
                        B200_MM_VEC(16, 4)
                    else
                        B200_MM_VEC(32, 4)
                }
                else
                {
                    if(lpr == 4)
                        B200_MM_VEC(4, 2)
                    else if(lpr == 8)
                        B200_MM_VEC(8, 2)
                    else if(lpr == 16)
                        B200_MM_VEC(16, 2)
                    else
                        B200_MM_VEC(32, 2)
                }
#undef B200_MM_VEC
            }
            else if(order == aoclsparse_order_row)
            {
                static std::atomic<size_t> cfg{0};
                if(cfg.load() < smem)
                {
                    B200_CUDA(cudaFuncSetAttribute(
                        csrmm_row_major_kernel<T, CONJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    cfg.store(smem);
                }
                csrmm_row_major_kernel<T, CONJ><<<P.n_blocks, MM_THREADS, smem, st>>>(P.desc.as<int4>(),
                                                                                       P.kind.as<int>(),
                                                                                       cap,
                                                                                       A.row_ptr.as<aoclsparse_int>(),
                                                                                       A.col_idx.as<aoclsparse_int>(),
                                                                                       A.val.as<T>(),
                                                                                       B,
                                                                                       ldb,
                                                                                       C,
                                                                                       ldc,
                                                                                       n,
                                                                                       alpha,
                                                                                       beta,
                                                                                       bz);
            }
            else
            {
                static std::atomic<size_t> cfg{0};
                if(cfg.load() < smem)
                {
                    B200_CUDA(cudaFuncSetAttribute(
                        csrmm_col_major_kernel<T, CONJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    cfg.store(smem);
                }
                csrmm_col_major_kernel<T, CONJ><<<P.n_blocks, MM_THREADS, smem, st>>>(P.desc.as<int4>(),
                                                                                       P.kind.as<int>(),
                                                                                       cap,
                                                                                       A.row_ptr.as<aoclsparse_int>(),
                                                                                       A.col_idx.as<aoclsparse_int>(),
                                                                                       A.val.as<T>(),
                                                                                       B,
                                                                                       ldb,
                                                                                       C,
                                                                                       ldc,
                                                                                       n,
                                                                                       alpha,
                                                                                       beta,
                                                                                       bz);
            }
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }

        inline bool int_product_overflow(aoclsparse_int a, aoclsparse_int b)
        {
            if(a == 0 || b == 0)
                return false;
            const long long p = (long long)a * (long long)b;
            const long long lim = sizeof(aoclsparse_int) == 4 ? 0x7fffffffLL : 0x7fffffffffffffffLL;
            if(sizeof(aoclsparse_int) == 8)
                return (a > 0 ? a : -a) > lim / (b > 0 ? b : -b);
            return p > lim || p < -lim - 1;
        }

        struct mm_staging
        {
            dev_buf B, C;
        };
        mm_staging &tls_mm_staging()
        {
            static thread_local mm_staging s;
            return s;
        }
    }

    // symmetric / hermitian descriptors: multiply with the expanded general copy (expand.cu)
    template <typename T>
    aoclsparse_status csrmm_symmetric(aoclsparse_operation         op,
                                      T                            alpha,
                                      const aoclsparse_matrix      A,
                                      const _aoclsparse_mat_descr &descr,
                                      aoclsparse_order             order,
                                      const T                     *B,
                                      long long                    ldb,
                                      T                            beta,
                                      T                           *C,
                                      long long                    ldc,
                                      int                          n,
                                      cudaStream_t                 st)
    {
        if(!vt<T>::is_complex && descr.type == aoclsparse_matrix_type_hermitian)
        {
            // real data: hermitian == symmetric (get_doid folds it, aoclsparse_mtx_dispatcher.hpp:93-99)
        }
        const dev_csr *F = nullptr;
        B200_TRY(get_expanded_copy(A, descr, op, F, st));
        std::shared_lock<std::shared_mutex> rl(A->guard);
        return launch_mm<T, false>(*F, order, B, ldb, C, ldc, n, alpha, beta, st);
    }

    template <typename T>
    aoclsparse_status csrmm_entry(aoclsparse_operation       op,
                                  T                          alpha,
                                  const aoclsparse_matrix    A,
                                  const aoclsparse_mat_descr descr,
                                  aoclsparse_order           order,
                                  const T                   *B,
                                  aoclsparse_int             n,
                                  aoclsparse_int             ldb,
                                  T                          beta,
                                  T                         *C,
                                  aoclsparse_int             ldc,
                                  aoclsparse_int             kid)
    {
        // ---- validation, in the reference's order (csrmm.hpp:447-618)
        if(A == nullptr || B == nullptr || C == nullptr || descr == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(A->input_format != aoclsparse_csr_mat)
            return aoclsparse_status_not_implemented;
        if(op != aoclsparse_operation_none && op != aoclsparse_operation_transpose
           && op != aoclsparse_operation_conjugate_transpose)
            return aoclsparse_status_invalid_value;
        if(descr->type != aoclsparse_matrix_type_general && descr->type != aoclsparse_matrix_type_symmetric
           && descr->type != aoclsparse_matrix_type_hermitian)
            return aoclsparse_status_not_implemented;
        if((descr->type == aoclsparse_matrix_type_symmetric || descr->type == aoclsparse_matrix_type_hermitian)
           && A->m != A->n)
            return aoclsparse_status_invalid_size;
        if(order != aoclsparse_order_row && order != aoclsparse_order_column)
            return aoclsparse_status_invalid_value;
        if(A->val_type != vt<T>::data_type)
            return aoclsparse_status_wrong_type;
        if(descr->base != A->base)
            return aoclsparse_status_invalid_value;
        if(A->mats.empty() || A->mats[0] == nullptr)
            return aoclsparse_status_not_implemented;

        const aoclsparse_int m = A->m, k = A->n;
        if(m < 0 || n < 0 || k < 0)
            return aoclsparse_status_invalid_size;
        if(m == 0 || n == 0 || k == 0)
            return aoclsparse_status_success;
        if(is_zero(alpha) && is_one(beta))
            return aoclsparse_status_success;

        const bool           none    = op == aoclsparse_operation_none;
        const aoclsparse_int chk_ldb = none ? (order == aoclsparse_order_column ? k : n)
                                            : (order == aoclsparse_order_column ? m : n);
        if(ldb < (chk_ldb > 1 ? chk_ldb : 1))
            return aoclsparse_status_invalid_size;
        const aoclsparse_int chk_ldc = none ? (order == aoclsparse_order_column ? m : n)
                                            : (order == aoclsparse_order_column ? k : n);
        if(ldc < (chk_ldc > 1 ? chk_ldc : 1))
            return aoclsparse_status_invalid_size;
        const aoclsparse_int m_c = none ? m : k, b_rows = none ? k : m;
        {
            const aoclsparse_int c_dim = order == aoclsparse_order_column ? n : m_c;
            const aoclsparse_int b_dim = order == aoclsparse_order_column ? n : b_rows;
            if(int_product_overflow(c_dim, ldc) || int_product_overflow(b_dim, ldb))
                return aoclsparse_status_invalid_size;
        }
        if(kid > 3)
            return aoclsparse_status_invalid_kid;

        cudaStream_t st = current_stream();

        // ---- B / C residency
        const long long b_elems = order == aoclsparse_order_column ? (long long)ldb * n : (long long)ldb * b_rows;
        const long long c_elems = order == aoclsparse_order_column ? (long long)ldc * n : (long long)ldc * m_c;
        const bool      b_dev = is_device_accessible(B), c_dev = is_device_accessible(C);
        const T        *dB = B;
        T              *dC = C;
        mm_staging     &sg = tls_mm_staging();
        // host operands: only the rows x columns the product names travel (pitched copies), so neither the bytes after
        // the last row of a sub-matrix view are read (a BLAS-style caller only guarantees (rows-1)*ld + cols elements)
        // nor is C's padding touched; with beta == 0 C is not read, so it is not uploaded either
        const long long b_lines = order == aoclsparse_order_column ? n : b_rows, b_inner = order == aoclsparse_order_column ? b_rows : n;
        const long long c_lines = order == aoclsparse_order_column ? n : m_c, c_inner = order == aoclsparse_order_column ? m_c : n;
        if(!b_dev)
        {
            if(sg.B.bytes < (size_t)b_elems * sizeof(T))
                B200_TRY(sg.B.alloc((size_t)b_elems * sizeof(T)));
            if(b_lines > 0 && b_inner > 0)
                B200_CUDA(cudaMemcpy2DAsync(sg.B.p, (size_t)ldb * sizeof(T), B, (size_t)ldb * sizeof(T), (size_t)b_inner * sizeof(T),
                                            (size_t)b_lines, cudaMemcpyHostToDevice, st));
            dB = sg.B.as<T>();
        }
        if(!c_dev)
        {
            if(sg.C.bytes < (size_t)c_elems * sizeof(T))
                B200_TRY(sg.C.alloc((size_t)c_elems * sizeof(T)));
            if(!is_zero(beta) && c_lines > 0 && c_inner > 0)
                B200_CUDA(cudaMemcpy2DAsync(sg.C.p, (size_t)ldc * sizeof(T), C, (size_t)ldc * sizeof(T), (size_t)c_inner * sizeof(T),
                                            (size_t)c_lines, cudaMemcpyHostToDevice, st));
            dC = sg.C.as<T>();
        }

        aoclsparse_status status = aoclsparse_status_success;
        if(is_zero(alpha))
            status = scale_dense<T>(order, dC, m_c, n, ldc, beta, st); // csrmm.hpp:613-618
        else
        {
            status = ensure_plan(A, st);
            if(status == aoclsparse_status_success)
            {
                const bool cplx    = vt<T>::is_complex;
                const bool conj_op = cplx && op == aoclsparse_operation_conjugate_transpose;
                // on the stored matrix (a CSC handle stores the transpose, csrmm.hpp:496-504)
                const bool trans_s = none == A->is_csc;
                if(descr->type != aoclsparse_matrix_type_general)
                    status = csrmm_symmetric<T>(op, alpha, A, *descr, order, dB, ldb, beta, dC, ldc, n, st);
                else if(!trans_s)
                {
                    std::shared_lock<std::shared_mutex> rl(A->guard);
                    if(conj_op)
                        status = launch_mm<T, true>(*A->mats[0], order, dB, ldb, dC, ldc, n, alpha, beta, st);
                    else
                        status = launch_mm<T, false>(*A->mats[0], order, dB, ldb, dC, ldc, n, alpha, beta, st);
                }
                else
                {
                    // transposed product: run the gather kernels on a transposed device copy (kept in the
                    // handle when the memory policy allows; the reference transposes on every call,
                    // csrmm.hpp:737-771)
                    const int      want = conj_op ? DOID_GH : DOID_GT;
                    const dev_csr *Tm   = nullptr;
                    dev_csr        temp;
                    {
                        std::unique_lock<std::shared_mutex> wl(A->guard);
                        for(size_t i = 1; i < A->mats.size(); ++i)
                            if(A->mats[i]->doid == want && A->mats[i]->plan.valid)
                                Tm = A->mats[i];
                        if(!Tm)
                        {
                            dev_csr *Cn = (A->mem_policy == aoclsparse_memory_usage_unrestricted) ? new(std::nothrow) dev_csr
                                                                                                  : &temp;
                            if(!Cn)
                                return aoclsparse_status_memory_error;
                            status = transpose_csr(*A->mats[0], A->val_type, conj_op, *Cn, st);
                            if(status == aoclsparse_status_success)
                                status = build_plan(*Cn, sizeof(T), -1, -1, std::vector<aoclsparse_int>(), st);
                            Cn->doid = want;
                            if(status == aoclsparse_status_success && Cn != &temp)
                                A->mats.push_back(Cn);
                            else if(Cn != &temp)
                                delete Cn;
                            if(status == aoclsparse_status_success)
                                Tm = Cn;
                        }
                    }
                    if(status == aoclsparse_status_success)
                    {
                        status = launch_mm<T, false>(*Tm, order, dB, ldb, dC, ldc, n, alpha, beta, st);
                        if(Tm == &temp)
                            cudaStreamSynchronize(st); // temp is destroyed on return
                    }
                }
            }
        }
        if(status != aoclsparse_status_success)
            return status;
        if(!c_dev)
        {
            if(c_lines > 0 && c_inner > 0)
                B200_CUDA(cudaMemcpy2DAsync(C, (size_t)ldc * sizeof(T), dC, (size_t)ldc * sizeof(T), (size_t)c_inner * sizeof(T),
                                            (size_t)c_lines, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
        }
        else if(!b_dev)
            B200_CUDA(cudaStreamSynchronize(st));
        return aoclsparse_status_success;
    }
}

using namespace b200;

extern "C" {

#define B200_CSRMM(SUF, CT, DT)                                                                                  \
    aoclsparse_status aoclsparse_##SUF##csrmm(aoclsparse_operation       op,                                     \
                                              const CT                   alpha,                                  \
                                              const aoclsparse_matrix    A,                                      \
                                              const aoclsparse_mat_descr descr,                                  \
                                              aoclsparse_order           order,                                  \
                                              const CT                  *B,                                      \
                                              aoclsparse_int             n,                                      \
                                              aoclsparse_int             ldb,                                    \
                                              const CT                   beta,                                   \
                                              CT                        *C,                                      \
                                              aoclsparse_int             ldc)                                    \
    {                                                                                                            \
        DT a_, b_;                                                                                               \
        memcpy(&a_, &alpha, sizeof(DT));                                                                         \
        memcpy(&b_, &beta, sizeof(DT));                                                                          \
        return csrmm_entry<DT>(                                                                                  \
            op, a_, A, descr, order, reinterpret_cast<const DT *>(B), n, ldb, b_, reinterpret_cast<DT *>(C), ldc, -1); \
    }                                                                                                            \
    aoclsparse_status aoclsparse_##SUF##csrmm_kid(aoclsparse_operation       op,                                 \
                                                  const CT                   alpha,                              \
                                                  const aoclsparse_matrix    A,                                  \
                                                  const aoclsparse_mat_descr descr,                              \
                                                  aoclsparse_order           order,                              \
                                                  const CT                  *B,                                  \
                                                  aoclsparse_int             n,                                  \
                                                  aoclsparse_int             ldb,                                \
                                                  const CT                   beta,                               \
                                                  CT                        *C,                                  \
                                                  aoclsparse_int             ldc,                                \
                                                  const aoclsparse_int       kid)                                \
    {                                                                                                            \
        DT a_, b_;                                                                                               \
        memcpy(&a_, &alpha, sizeof(DT));                                                                         \
        memcpy(&b_, &beta, sizeof(DT));                                                                          \
        return csrmm_entry<DT>(                                                                                  \
            op, a_, A, descr, order, reinterpret_cast<const DT *>(B), n, ldb, b_, reinterpret_cast<DT *>(C), ldc, kid); \
    }

B200_CSRMM(s, float, float)
B200_CSRMM(d, double, double)
B200_CSRMM(c, aoclsparse_float_complex, float2)
B200_CSRMM(z, aoclsparse_double_complex, double2)
}
