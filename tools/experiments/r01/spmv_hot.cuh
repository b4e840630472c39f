// spmv_hot.cuh -- SpMV for matrices whose columns are hit very unevenly (power-law graphs, BASELINE config 3).
//
// On such a matrix x[col] is a random 4/8-byte gather; the multiply is bound by how many L1-MISSING gathers an SM
// retires (~0.9 per clock, profiles/r01_microbench_gather.txt), not by HBM.  The hardware L1 is of little help: it
// allocates 128-byte lines, so ~100 KB hold only ~800 distinct hot columns (ncu: 6 % hit rate on R-MAT scale 24).
// aoclsparse_optimize therefore counts how often every column occurs, takes the K most frequent ones and writes a
// second column array in which those columns are replaced by (HOT_BIT | slot).  This kernel keeps the K hot values of
// x in shared memory -- element granularity, so 24 K floats fit where L1 held 800 -- and only the remaining gathers
// go to L1/L2.
//
// Shape: one CTA of 1024 threads per SM, resident for the whole launch; the table is filled once per CTA.  The CTA
// is four independent groups of 256 threads; each group walks its own sequence of row blocks with its own
// double-buffered TMA staging (bulk copies of val / col for block i+1 are in flight while block i is reduced) and
// reduces a block exactly like spmv_row_blocks_kernel does (same strategies, same summation order), with named
// barriers instead of __syncthreads().
#pragma once
#include "spmv_kernels.cuh"

namespace b200
{
    constexpr int HOT_GROUPS  = 4;
    constexpr int HOT_GT      = 256; // threads per group
    constexpr int HOT_THREADS = HOT_GROUPS * HOT_GT;
    constexpr int HOT_BIT     = (int)0x80000000;

    inline size_t hot_smem_bytes(size_t elem_size, aoclsparse_int block_nnz, int table_entries, int stages)
    {
        // header (8 mbarriers) | table | 4 groups x stages x (val[cap] + col[cap])
        return 128 + (((size_t)table_entries * elem_size + 15) & ~(size_t)15)
               + (size_t)HOT_GROUPS * stages * (size_t)(block_nnz + 8) * (elem_size + 4);
    }

    template <typename T>
    __device__ __forceinline__ T hot_get(const T *__restrict__ x, const T *xs, int c)
    {
        return c < 0 ? xs[c & 0x7fffffff] : ldg_ro(x + c);
    }

    // four gathers at once: the L1/L2 loads of the non-hot columns are issued together (predicated off for hot
    // columns), then the hot ones are taken from the table
    template <typename T>
    __device__ __forceinline__ void hot_get4(const T *__restrict__ x, const T *xs, const int c[4], const bool ok[4], T out[4])
    {
#pragma unroll
        for(int u = 0; u < 4; ++u)
            out[u] = (ok[u] && c[u] >= 0) ? ldg_ro(x + c[u]) : vt<T>::zero();
#pragma unroll
        for(int u = 0; u < 4; ++u)
            if(ok[u] && c[u] < 0)
                out[u] = xs[c[u] & 0x7fffffff];
    }

    __device__ __forceinline__ void group_sync(int g)
    {
        asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(HOT_GT) : "memory");
    }

    // STAGES = 2: one CTA per SM, copies of block i+1 in flight while block i is reduced
    // STAGES = 1: two CTAs per SM (half-size table each), overlap comes from the 8 groups resident per SM
    template <typename T, int STAGES>
    __global__ void __launch_bounds__(HOT_THREADS, (STAGES == 1 ? 2 : 1)) spmv_hot_kernel(const int4 *__restrict__ desc,
                                                                     const int *__restrict__ kind,
                                                                     int n_blocks,
                                                                     int cap,
                                                                     const aoclsparse_int *__restrict__ rp,
                                                                     const aoclsparse_int *__restrict__ col_hot,
                                                                     const T *__restrict__ val,
                                                                     const T *__restrict__ x,
                                                                     T *__restrict__ y,
                                                                     T   alpha,
                                                                     T   beta,
                                                                     int beta_zero,
                                                                     T  *partials,
                                                                     const aoclsparse_int *__restrict__ hot_cols,
                                                                     int table_entries)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        uint64_t     *bars = reinterpret_cast<uint64_t *>(smem_raw); // [group][stage]
        T            *xs   = reinterpret_cast<T *>(smem_raw + 128);
        const size_t  stage_bytes = (size_t)cap * (sizeof(T) + sizeof(aoclsparse_int));
        unsigned char *ring = smem_raw + 128 + (((size_t)table_entries * sizeof(T) + 15) & ~(size_t)15);
        __shared__ T   s_part[HOT_GROUPS][HOT_GT / 32];

        const int tid_all = threadIdx.x;
        const int g       = tid_all / HOT_GT; // group
        const int tid     = tid_all % HOT_GT;
        const int lane = tid & 31, warp = tid >> 5;
        constexpr int NT = HOT_GT;

        if(tid_all == 0)
        {
            for(int i = 0; i < HOT_GROUPS * 2; ++i)
                mbar_init(&bars[i], 1);
            mbar_init_fence();
        }
        asm volatile("griddepcontrol.launch_dependents;");
        __syncthreads();

        const int first = blockIdx.x * HOT_GROUPS + g, step = gridDim.x * HOT_GROUPS;
        unsigned char *my_ring = ring + (size_t)g * STAGES * stage_bytes;
        auto           issue   = [&](int b, int s) {
            const int4 d   = desc[b];
            const int  a   = d.z & ~3;
            const int  cnt = ((d.w - a) + 3) & ~3;
            uint64_t  *bar = &bars[g * 2 + s];
            T              *sv = reinterpret_cast<T *>(my_ring + (size_t)s * stage_bytes);
            aoclsparse_int *sc = reinterpret_cast<aoclsparse_int *>(my_ring + (size_t)s * stage_bytes + (size_t)cap * sizeof(T));
            if(cnt > 0)
            {
                mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                bulk_load_stream(sv, val + a, (unsigned)(cnt * sizeof(T)), bar);
                bulk_load_stream(sc, col_hot + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar);
            }
            else
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        };
        if(tid == 0 && first < n_blocks)
            issue(first, 0);

        // x may be the previous launch's output: everything that reads it comes after this point
        asm volatile("griddepcontrol.wait;" ::: "memory");
        for(int s = tid_all; s < table_entries; s += HOT_THREADS)
            xs[s] = ldg_ro(x + hot_cols[s]);
        __syncthreads();

        int i = 0;
        for(int b = first; b < n_blocks; b += step, ++i)
        {
            const int s = STAGES == 2 ? (i & 1) : 0;
            // next block's copies go into the other stage; every thread of the group left it at the end of the
            // previous iteration (group_sync below)
            if(STAGES == 2)
            {
                if(tid == 0 && b + step < n_blocks)
                    issue(b + step, s ^ 1);
            }
            else if(i > 0 && tid == 0)
                issue(b, 0);
            const int4 d     = desc[b];
            const int  k     = kind[b];
            const int  strat = k & 15;
            const int  ns = d.z, ne = d.w;
            const int  a = ns & ~3;
            T              *sval = reinterpret_cast<T *>(my_ring + (size_t)s * stage_bytes);
            aoclsparse_int *scol = reinterpret_cast<aoclsparse_int *>(my_ring + (size_t)s * stage_bytes + (size_t)cap * sizeof(T));
            mbar_wait(&bars[g * 2 + s], (unsigned)((STAGES == 2 ? (i >> 1) : i) & 1));

            if(strat == STRAT_THREAD)
            {
                for(int r = d.x + tid; r < d.y; r += NT)
                {
                    int       j   = rp[r] - a;
                    const int e   = rp[r + 1] - a;
                    T         acc = vt<T>::zero();
                    for(; j < e; ++j)
                        acc = mad(sval[j], hot_get(x, xs, scol[j]), acc);
                    y[r] = axpby_out(alpha, acc, beta, beta_zero != 0, y + r);
                }
            }
            else if(strat == STRAT_WARP)
            {
                for(int r = d.x + warp; r < d.y; r += NT / 32)
                {
                    const int s0 = rp[r] - a, e = rp[r + 1] - a;
                    T         acc = vt<T>::zero();
                    for(int j = s0 + lane; j < e; j += 128)
                    {
                        int  c[4];
                        bool ok[4];
                        T    xv[4];
#pragma unroll
                        for(int u = 0; u < 4; ++u)
                        {
                            ok[u] = j + 32 * u < e;
                            c[u]  = ok[u] ? scol[j + 32 * u] : 0;
                        }
                        hot_get4(x, xs, c, ok, xv);
#pragma unroll
                        for(int u = 0; u < 4; ++u)
                            if(ok[u])
                                acc = mad(sval[j + 32 * u], xv[u], acc);
                    }
                    acc = warp_sum(acc);
                    if(lane == 0)
                        y[r] = axpby_out(alpha, acc, beta, beta_zero != 0, y + r);
                }
            }
            else if(strat == STRAT_PRODUCT)
            {
                const int f0 = ns - a, total = ne - ns;
                for(int q = tid; q < total; q += 4 * NT)
                {
                    int  c[4];
                    bool ok[4];
                    T    xv[4];
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                    {
                        ok[u] = q + u * NT < total;
                        c[u]  = ok[u] ? scol[f0 + q + u * NT] : 0;
                    }
                    hot_get4(x, xs, c, ok, xv);
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                        if(ok[u])
                            sval[f0 + q + u * NT] = mul(sval[f0 + q + u * NT], xv[u]);
                }
                group_sync(g);
                for(int rb = d.x + warp * 32; rb < d.y; rb += NT)
                {
                    const int  r     = rb + lane;
                    const bool valid = r < d.y;
                    int        s0 = 0, e = 0;
                    if(valid)
                    {
                        s0 = rp[r] - a;
                        e  = rp[r + 1] - a;
                    }
                    T acc = vt<T>::zero();
                    if(e - s0 <= SHORT_ROW)
                        for(int j = s0; j < e; ++j)
                            acc = add(acc, sval[j]);
                    unsigned pending = __ballot_sync(0xffffffffu, valid && (e - s0 > SHORT_ROW));
                    while(pending)
                    {
                        const int src = __ffs(pending) - 1;
                        pending &= pending - 1;
                        const int ss = __shfl_sync(0xffffffffu, s0, src);
                        const int ee = __shfl_sync(0xffffffffu, e, src);
                        T         part = vt<T>::zero();
                        for(int j = ss + lane; j < ee; j += 32)
                            part = add(part, sval[j]);
                        part = warp_sum(part);
                        if(lane == src)
                            acc = part;
                    }
                    if(valid)
                        y[r] = axpby_out(alpha, acc, beta, beta_zero != 0, y + r);
                }
            }
            else
            {
                const int f0 = ns - a, total = ne - ns;
                T         acc = vt<T>::zero();
                for(int q = tid; q < total; q += 4 * NT)
                {
                    int  c[4];
                    bool ok[4];
                    T    xv[4];
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                    {
                        ok[u] = q + u * NT < total;
                        c[u]  = ok[u] ? scol[f0 + q + u * NT] : 0;
                    }
                    hot_get4(x, xs, c, ok, xv);
#pragma unroll
                    for(int u = 0; u < 4; ++u)
                        if(ok[u])
                            acc = mad(sval[f0 + q + u * NT], xv[u], acc);
                }
                acc = warp_sum(acc);
                if(lane == 0)
                    s_part[g][warp] = acc;
                group_sync(g);
                if(tid == 0)
                {
                    T tot = s_part[g][0];
#pragma unroll
                    for(int w = 1; w < NT / 32; ++w)
                        tot = add(tot, s_part[g][w]);
                    partials[k >> 4] = tot;
                }
            }
            group_sync(g); // the stage (and s_part) may be overwritten from here on
        }
    }
}
