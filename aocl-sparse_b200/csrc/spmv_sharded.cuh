// spmv_sharded.cuh -- one launch per iteration of the row-sharded product x_{k+1} = alpha * A_local * x_k:
// compute AND halo exchange in the same kernel (BASELINE config 5).
//
// A rank's rows are cut (aoclsparse_b200_set_row_cuts) into  [first boundary | interior | last boundary]; only the
// boundary rows read halo entries of x and only their results are needed by the neighbours.  The grid is ordered
// boundary blocks first:
//   * a boundary CTA first waits (thread 0 spins on a flag in this GPU's memory that the neighbour writes over NVLink)
//     until that neighbour's facing boundary has completed iteration k-1.  That one event means both "my halo of x_k
//     has been delivered" and "the neighbour no longer reads the halo I am about to overwrite" (only its facing
//     boundary rows read it).  Then the CTA reduces its rows like any other block and stores every result twice: to
//     the local y and, over NVLink, into the neighbour's halo of x_{k+1}.  The last boundary CTA of a side to finish
//     publishes "boundary done k" to that neighbour.
//   * interior CTAs need no remote data, take part in no counter or fence, and run meanwhile: the transfer overlaps
//     the bulk of the multiply.
// Counters only ever count up (target = k * number of CTAs), so nothing is reset between iterations.  Every block
// must be binned thread-per-row (stencil / banded matrices); otherwise the caller uses the multi-launch path.
#pragma once
#include "spmv_kernels.cuh"

namespace b200
{
    struct halo_ctl
    {
        // waits (addresses in THIS GPU's memory, written by the neighbours); nullptr = no neighbour on that side
        const unsigned *left_done, *right_done;
        // signals (addresses in the NEIGHBOURS' memory)
        unsigned *to_left_done, *to_right_done;
        unsigned *counters; // [0] first-boundary CTAs done, [1] last-boundary CTAs done, [3] a flag wait gave up
        unsigned  k;        // iteration number, 1-based: the value the flags carry
        unsigned  kc;       // launches of this kernel so far, this one included: the counters' target is kc * CTAs
        int       n_first, n_last, n_blocks; // blocks in the first boundary, the last boundary, in total
        int       last_begin;                // index of the first block of the last boundary
        int       first_rows, last_row0;     // rows in the first boundary; first row of the last boundary
    };

    __device__ __forceinline__ bool spin_until(const unsigned *flag, unsigned value, unsigned *timeout)
    {
        const volatile unsigned *f = flag;
        long long                spins = 0;
        while((int)(*f - value) < 0)
        {
            __nanosleep(100);
            if(++spins > 20000000LL)
            {
                *timeout = 1;
                return false;
            }
        }
        __threadfence_system();
        return true;
    }

    // CODED: the column stream is the diagonal-code copy (one byte per entry, col = row + code_off[code]; see
    // spmv_row_blocks_kernel)
    template <typename T, bool CODED = false>
    __global__ void __launch_bounds__(256) spmv_sharded_step_kernel(const int4 *__restrict__ desc,
                                                                   int cap,
                                                                   const aoclsparse_int *__restrict__ rp,
                                                                   const aoclsparse_int *__restrict__ col,
                                                                   const T *__restrict__ val,
                                                                   const T *__restrict__ x,
                                                                   T *__restrict__ y,
                                                                   T        alpha,
                                                                   T       *push_left,  // neighbour's halo for my first rows
                                                                   T       *push_right, // neighbour's halo for my last rows
                                                                   halo_ctl hc,
                                                                   const unsigned char *__restrict__ codes = nullptr,
                                                                   const int *__restrict__ code_off = nullptr)
    {
        constexpr int NT = 256;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        uint64_t            *bar   = reinterpret_cast<uint64_t *>(smem_raw);
        T                   *sval  = reinterpret_cast<T *>(smem_raw + SMEM_HEADER);
        aoclsparse_int      *scol  = reinterpret_cast<aoclsparse_int *>(smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T));
        const unsigned char *scode = smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T);
        int                 *soff  = reinterpret_cast<int *>(smem_raw + SMEM_HEADER + (size_t)cap * (sizeof(T) + 1));

        const int tid = threadIdx.x;
        // launch order -> block: first boundary, last boundary, then the interior
        const int bid = blockIdx.x;
        int       b, side; // side 0 first boundary, 1 last boundary, 2 interior
        if(bid < hc.n_first)
        {
            b    = bid;
            side = 0;
        }
        else if(bid < hc.n_first + hc.n_last)
        {
            b    = hc.last_begin + (bid - hc.n_first);
            side = 1;
        }
        else
        {
            b    = hc.n_first + (bid - hc.n_first - hc.n_last);
            side = 2;
        }
        const int4 d = desc[b];
        asm volatile("griddepcontrol.launch_dependents;");
        constexpr int GR  = CODED ? 16 : 4; // entries per 16-byte granule of the narrowest staged array
        const int     a   = d.z & ~(GR - 1);
        const int     cnt = ((d.w - a) + GR - 1) & ~(GR - 1);
        if(tid == 0)
        {
            mbar_init(bar, 1);
            mbar_init_fence();
            if(cnt > 0)
            {
                if constexpr(CODED)
                {
                    mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + 1)));
                    bulk_load_stream(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                    bulk_load_stream(const_cast<unsigned char *>(scode), codes + a, (unsigned)cnt, bar);
                }
                else
                {
                    mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                    bulk_load_stream(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                    bulk_load_stream(scol, col + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar);
                }
            }
        }
        if constexpr(CODED)
            for(int i = tid; i < CODE_TABLE; i += NT)
                soff[i] = code_off[i];
        __syncthreads();
        int pre_s = 0, pre_e = 0;
        if(d.x + tid < d.y)
        {
            pre_s = rp[d.x + tid];
            pre_e = rp[d.x + tid + 1];
        }
        // x is the previous launch's y: wait for that grid (programmatic dependent launch)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if(side != 2)
        {
            // halo of x_k present, and the neighbour done with the buffer the results go to
            if(tid == 0)
            {
                const unsigned *done = side == 0 ? hc.left_done : hc.right_done;
                if(done)
                    spin_until(done, hc.k - 1, hc.counters + 3);
            }
            __syncthreads();
        }
        if(cnt > 0)
            mbar_wait(bar, 0);

        // column of staged entry j of row r
        auto col_at = [&](int r, int j) -> int {
            if constexpr(CODED)
                return r + soff[scode[j]];
            else
                return scol[j];
        };
        T *push = side == 0 ? push_left : (side == 1 ? push_right : nullptr);
        const int push_row0 = side == 0 ? 0 : hc.last_row0;
        if(side == 2)
        {
            // interior rows: every x entry they name was written by the previous LOCAL launch, which has completed
            // (griddepcontrol.wait) -> read-only path through L1
            for(int r = d.x + tid; r < d.y; r += NT)
            {
                const bool first = r == d.x + tid;
                int        j     = (first ? pre_s : rp[r]) - a;
                const int  e     = (first ? pre_e : rp[r + 1]) - a;
                T          acc   = vt<T>::zero();
                for(; j + 4 <= e; j += 4)
                {
                    const int c0 = col_at(r, j), c1 = col_at(r, j + 1), c2 = col_at(r, j + 2), c3 = col_at(r, j + 3);
                    const T   x0 = ldg_ro(x + c0), x1 = ldg_ro(x + c1), x2 = ldg_ro(x + c2), x3 = ldg_ro(x + c3);
                    acc          = mad(sval[j], x0, acc);
                    acc          = mad(sval[j + 1], x1, acc);
                    acc          = mad(sval[j + 2], x2, acc);
                    acc          = mad(sval[j + 3], x3, acc);
                }
                for(; j < e; ++j)
                    acc = mad(sval[j], ldg_ro(x + col_at(r, j)), acc);
                y[r] = mul(alpha, acc);
            }
        }
        else
        {
            // boundary rows: the halo part of x is stored by the PEER GPU while this grid is already running, so it
            // must not go through the non-coherent path (ld.global.nc requires the data to be read-only for the
            // kernel's lifetime): ld.global.cg reads at the L2, the coherence point the peer's stores arrive at
            for(int r = d.x + tid; r < d.y; r += NT)
            {
                const bool first = r == d.x + tid;
                int        j     = (first ? pre_s : rp[r]) - a;
                const int  e     = (first ? pre_e : rp[r + 1]) - a;
                T          acc   = vt<T>::zero();
                for(; j + 4 <= e; j += 4)
                {
                    const int c0 = col_at(r, j), c1 = col_at(r, j + 1), c2 = col_at(r, j + 2), c3 = col_at(r, j + 3);
                    const T   x0 = __ldcg(x + c0), x1 = __ldcg(x + c1), x2 = __ldcg(x + c2), x3 = __ldcg(x + c3);
                    acc          = mad(sval[j], x0, acc);
                    acc          = mad(sval[j + 1], x1, acc);
                    acc          = mad(sval[j + 2], x2, acc);
                    acc          = mad(sval[j + 3], x3, acc);
                }
                for(; j < e; ++j)
                    acc = mad(sval[j], __ldcg(x + col_at(r, j)), acc);
                const T out = mul(alpha, acc);
                y[r]        = out;
                if(push)
                    push[r - push_row0] = out;
            }
        }

        // ---- completion bookkeeping (boundary CTAs only): the last one of a side tells that neighbour
        if(side != 2)
        {
            __syncthreads();
            if(tid == 0)
            {
                __threadfence_system(); // this CTA's stores (peer stores included) before the counter / flag
                const unsigned n_side = side == 0 ? (unsigned)hc.n_first : (unsigned)hc.n_last;
                const unsigned done   = atomicAdd(hc.counters + side, 1u) + 1u;
                if(done == hc.kc * n_side)
                {
                    unsigned *flag = side == 0 ? hc.to_left_done : hc.to_right_done;
                    if(flag)
                    {
                        __threadfence_system();
                        *reinterpret_cast<volatile unsigned *>(flag) = hc.k;
                    }
                }
            }
        }
    }
}
