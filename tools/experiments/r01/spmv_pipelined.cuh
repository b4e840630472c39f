// spmv_pipelined.cuh -- persistent, warp-specialised variant of the row-block SpMV for plans whose blocks are
// all binned thread-per-row (stencils, banded matrices: BASELINE configs 1, 2, 5).
//
// Why a second kernel: spmv_row_blocks_kernel runs one row block per CTA and relies on ~8 resident CTAs per SM
// being in different phases (bulk copy in flight / reducing) to keep HBM busy.  On a matrix with only a couple of
// waves of CTAs (config 1: 2478 blocks, 2.1 waves) all CTAs start together, so the chip alternates between
// "everybody waits for its copy" and "everybody reduces" and HBM idles half the time (ncu: 42 % DRAM throughput).
// Here a CTA stays resident and walks blocks b = blockIdx.x, + gridDim.x, ...:
//   * warp NCW (one elected lane) is the PRODUCER: for each of its blocks it waits for a free stage, then issues
//     the two TMA bulk copies (val, col_idx) into that stage of a ring in shared memory (full/empty mbarriers);
//   * warps 0..NCW-1 are CONSUMERS: they wait for the stage to fill, reduce the block thread-per-row out of shared
//     memory, and hand the stage back.  The descriptor of the block two ahead and the row bounds (and y, beta != 0)
//     of the next block are prefetched into registers while the current block is reduced, so no dependent global
//     load sits between a stage filling and its x gathers.
// So up to NSTAGES slices per CTA are in flight at all times, independent of what the consumers are doing.
#pragma once
#include "spmv_kernels.cuh"

namespace b200
{
    constexpr int PIPE_CONSUMER_WARPS = 8;
    constexpr int PIPE_THREADS        = 32 * (PIPE_CONSUMER_WARPS + 1);
    constexpr int PIPE_MAX_STAGES     = 8;
    constexpr int PIPE_HEADER         = 16 * PIPE_MAX_STAGES; // full[8] + empty[8] mbarriers

    inline size_t pipe_smem_bytes(size_t elem_size, aoclsparse_int block_nnz, int stages)
    {
        return PIPE_HEADER + (size_t)stages * (size_t)(block_nnz + 8) * (elem_size + 4);
    }

    __device__ __forceinline__ void mbar_arrive(uint64_t *bar)
    {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    }

    template <typename T>
    __global__ void __launch_bounds__(PIPE_THREADS) spmv_thread_pipelined_kernel(const int4 *__restrict__ desc,
                                                                                int block_first,
                                                                                int block_end,
                                                                                int cap,
                                                                                int stages,
                                                                                const aoclsparse_int *__restrict__ rp,
                                                                                const aoclsparse_int *__restrict__ col,
                                                                                const T *__restrict__ val,
                                                                                const T *__restrict__ x,
                                                                                T *__restrict__ y,
                                                                                T   alpha,
                                                                                T   beta,
                                                                                int beta_zero,
                                                                                T  *push_dst,
                                                                                int push_row0)
    {
        constexpr int NTC = 32 * PIPE_CONSUMER_WARPS;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        uint64_t     *full  = reinterpret_cast<uint64_t *>(smem_raw);
        uint64_t     *empty = full + PIPE_MAX_STAGES;
        const size_t  stage_bytes = (size_t)cap * (sizeof(T) + sizeof(aoclsparse_int));
        unsigned char *ring = smem_raw + PIPE_HEADER;

        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        if(tid == 0)
        {
            for(int s = 0; s < stages; ++s)
            {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], PIPE_CONSUMER_WARPS);
            }
            mbar_init_fence();
        }
        __syncthreads();

        const int first = block_first + blockIdx.x, step = gridDim.x;

        if(warp == PIPE_CONSUMER_WARPS)
        {
            // ---------------- producer ----------------
            if(lane == 0)
            {
                int  i = 0;
                int4 d = (first < block_end) ? desc[first] : make_int4(0, 0, 0, 0);
                for(int b = first; b < block_end; b += step, ++i)
                {
                    const int  s     = i % stages;
                    const int  round = i / stages;
                    const int4 dn    = (b + step < block_end) ? desc[b + step] : make_int4(0, 0, 0, 0);
                    if(round > 0)
                        mbar_wait(&empty[s], (unsigned)((round - 1) & 1));
                    const int a   = d.z & ~3;
                    const int cnt = ((d.w - a) + 3) & ~3;
                    T              *sval = reinterpret_cast<T *>(ring + (size_t)s * stage_bytes);
                    aoclsparse_int *scol = reinterpret_cast<aoclsparse_int *>(ring + (size_t)s * stage_bytes + (size_t)cap * sizeof(T));
                    if(cnt > 0)
                    {
                        mbar_expect_tx(&full[s], (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                        bulk_load_stream(sval, val + a, (unsigned)(cnt * sizeof(T)), &full[s]);
                        bulk_load_stream(scol, col + a, (unsigned)(cnt * sizeof(aoclsparse_int)), &full[s]);
                    }
                    else
                        mbar_arrive(&full[s]);
                    d = dn;
                }
            }
            return;
        }

        // ---------------- consumers ----------------
        int4 d_cur = (first < block_end) ? desc[first] : make_int4(0, 0, 0, 0);
        int4 d_nxt = (first + step < block_end) ? desc[first + step] : make_int4(0, 0, 0, 0);
        // row bounds of this thread's first two rows of the current block (later rows are read in the loop)
        int s0 = 0, e0 = 0, s1 = 0, e1 = 0;
        T   y0 = vt<T>::zero(), y1 = vt<T>::zero();
        {
            const int r0 = d_cur.x + tid, r1 = r0 + NTC;
            if(r0 < d_cur.y)
            {
                s0 = rp[r0];
                e0 = rp[r0 + 1];
                if(!beta_zero)
                    y0 = y[r0];
            }
            if(r1 < d_cur.y)
            {
                s1 = rp[r1];
                e1 = rp[r1 + 1];
                if(!beta_zero)
                    y1 = y[r1];
            }
        }
        int i = 0;
        for(int b = first; b < block_end; b += step, ++i)
        {
            const int s = i % stages, round = i / stages;
            // prefetch: descriptor two blocks ahead, bounds / y of the next block
            const int4 d_nn = (b + 2 * step < block_end) ? desc[b + 2 * step] : make_int4(0, 0, 0, 0);
            int        ns0 = 0, ne0 = 0, ns1 = 0, ne1 = 0;
            T          ny0 = vt<T>::zero(), ny1 = vt<T>::zero();
            {
                const int r0 = d_nxt.x + tid, r1 = r0 + NTC;
                if(r0 < d_nxt.y)
                {
                    ns0 = rp[r0];
                    ne0 = rp[r0 + 1];
                    if(!beta_zero)
                        ny0 = y[r0];
                }
                if(r1 < d_nxt.y)
                {
                    ns1 = rp[r1];
                    ne1 = rp[r1 + 1];
                    if(!beta_zero)
                        ny1 = y[r1];
                }
            }
            const T              *sval = reinterpret_cast<const T *>(ring + (size_t)s * stage_bytes);
            const aoclsparse_int *scol
                = reinterpret_cast<const aoclsparse_int *>(ring + (size_t)s * stage_bytes + (size_t)cap * sizeof(T));
            const int a = d_cur.z & ~3;
            mbar_wait(&full[s], (unsigned)(round & 1));

            int pass = 0;
            for(int r = d_cur.x + tid; r < d_cur.y; r += NTC, ++pass)
            {
                int j, e;
                T   yin;
                if(pass == 0)
                {
                    j   = s0 - a;
                    e   = e0 - a;
                    yin = y0;
                }
                else if(pass == 1)
                {
                    j   = s1 - a;
                    e   = e1 - a;
                    yin = y1;
                }
                else
                {
                    j   = rp[r] - a;
                    e   = rp[r + 1] - a;
                    yin = beta_zero ? vt<T>::zero() : y[r];
                }
                T acc = vt<T>::zero();
                for(; j + 4 <= e; j += 4)
                {
                    const int c0 = scol[j], c1 = scol[j + 1], c2 = scol[j + 2], c3 = scol[j + 3];
                    const T   x0 = ldg_ro(x + c0), x1 = ldg_ro(x + c1), x2 = ldg_ro(x + c2), x3 = ldg_ro(x + c3);
                    acc          = mad(sval[j], x0, acc);
                    acc          = mad(sval[j + 1], x1, acc);
                    acc          = mad(sval[j + 2], x2, acc);
                    acc          = mad(sval[j + 3], x3, acc);
                }
                for(; j < e; ++j)
                    acc = mad(sval[j], ldg_ro(x + scol[j]), acc);
                T out = mul(alpha, acc);
                if(!beta_zero)
                    out = mad(beta, yin, out);
                y[r] = out;
                if(push_dst)
                    push_dst[r - push_row0] = out;
            }
            __syncwarp();
            if(lane == 0)
                mbar_arrive(&empty[s]);
            d_cur = d_nxt;
            d_nxt = d_nn;
            s0 = ns0;
            e0 = ne0;
            s1 = ns1;
            e1 = ne1;
            y0 = ny0;
            y1 = ny1;
        }
    }
}
