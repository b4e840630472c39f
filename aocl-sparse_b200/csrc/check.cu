// check.cu -- GPU restatement of the CSR validation / classification the reference performs,
// serially on the host, inside aoclsparse_create_?csr:
//   aoclsparse_mat_check_internal  library/src/analysis/aoclsparse_csr_util.cpp:124-279
//
// What must come out bit-exact:
//   * the status code, with the reference's precedence: a decreasing row_ptr anywhere beats every
//     index error; otherwise the FIRST offending entry in storage order decides between
//     invalid_index_value (column outside [0,n)) and invalid_value (second diagonal entry of a row);
//   * sort  in {fully_sorted, partially_sorted, unsorted};
//   * fulldiag.
//
// The serial scan carries state along a row (running maximum column, "seen diagonal", "seen upper").
// Each of its outcomes is restated as an order-independent property of the row so that a group of
// lanes can evaluate it with min/max reductions:
//   partially sorted  <=> some adjacent pair in a row descends (col[k-1] > col[k])
//                         [the serial test "running max > col[k]" fires iff a descent exists]
//   unsorted          <=> an entry with col <= row sits after an entry with col > row, or an entry
//                         with col < row sits after the diagonal entry
//                         <=> lastLE > firstU  or  lastL > firstD   (positions within the row)
//   duplicate diagonal at position p <=> p is the second-smallest position with col == row
//   first error       = min over all rows of (position*2 + kind), kind 0 = range, 1 = dup diag
// The extra facts min_col / max_col / max_row_nnz feed the multiply plan and the x-window check.
#include "common.hpp"

namespace b200
{
    namespace
    {
        struct check_acc
        {
            unsigned long long first_err; // position*2 + kind, ~0 if none
            int                flags;     // bit0 descent, bit1 group-order violation, bit2 missing diag, bit3 ptr decreases
            int                min_col;
            int                max_col;
            int                max_row_nnz;
        };

        constexpr int F_DESCENT = 1, F_GROUP = 2, F_NODIAG = 4, F_PTRDEC = 8;
        constexpr int BIG = 0x7fffffff;

        template <int G>
        __global__ void __launch_bounds__(256) check_rows_kernel(aoclsparse_int m,
                                                                 aoclsparse_int n,
                                                                 aoclsparse_int nnz,
                                                                 int            base,
                                                                 const aoclsparse_int *__restrict__ rp,
                                                                 const aoclsparse_int *__restrict__ col,
                                                                 check_acc *acc)
        {
            const int      lane_in_group = threadIdx.x % G;
            const long long gid          = (long long)blockIdx.x * (blockDim.x / G) + threadIdx.x / G;
            const long long ngroups      = (long long)gridDim.x * (blockDim.x / G);

            unsigned long long first_err = ~0ull;
            int                flags = 0, mincol = BIG, maxcol = -1, maxlen = 0;

            // the trip count is made warp-uniform (the group shuffles below use the full mask)
            const long long warp_first = gid - (threadIdx.x % 32) / G;
            for(long long it = 0; warp_first + it * ngroups < m; ++it)
            {
                const long long row    = gid + it * ngroups;
                const bool      active = row < m;
                long long       s = 0, e = 0;
                if(active)
                {
                    s = (long long)rp[row] - base;
                    e = (long long)rp[row + 1] - base;
                    if(s > e)
                    {
                        flags |= F_PTRDEC;
                        e = s; // skip the entries, keep the lanes converged
                    }
                }
                // a decreasing row_ptr elsewhere may push this row outside the arrays; the verdict is
                // invalid_value in that case anyway, so only memory safety matters here
                if(s < 0)
                    s = 0;
                if(e > nnz)
                    e = nnz;
                const int i = (int)row;
                int firstU = BIG, firstD = BIG, secondD = BIG, lastLE = -1, lastL = -1, oor = BIG;
                for(long long p = s + lane_in_group; p < e; p += G)
                {
                    const int j = col[p] - base;
                    const int q = (int)(p - s); // position within the row
                    if(j < 0 || j >= n)
                    {
                        oor = min(oor, q);
                        continue; // the serial scan returns here; nothing after matters
                    }
                    mincol = min(mincol, j);
                    maxcol = max(maxcol, j);
                    if(q > 0 && col[p - 1] > col[p])
                        flags |= F_DESCENT;
                    if(j > i)
                        firstU = min(firstU, q);
                    else
                    {
                        lastLE = max(lastLE, q);
                        if(j == i)
                        {
                            if(q < firstD)
                            {
                                secondD = firstD;
                                firstD  = q;
                            }
                            else if(q < secondD)
                                secondD = q;
                        }
                        else
                            lastL = max(lastL, q);
                    }
                }
                // combine the lanes of the group
#pragma unroll
                for(int off = G / 2; off > 0; off >>= 1)
                {
                    firstU   = min(firstU, __shfl_xor_sync(0xffffffffu, firstU, off, G));
                    lastLE   = max(lastLE, __shfl_xor_sync(0xffffffffu, lastLE, off, G));
                    lastL    = max(lastL, __shfl_xor_sync(0xffffffffu, lastL, off, G));
                    oor      = min(oor, __shfl_xor_sync(0xffffffffu, oor, off, G));
                    int oD1  = __shfl_xor_sync(0xffffffffu, firstD, off, G);
                    int oD2  = __shfl_xor_sync(0xffffffffu, secondD, off, G);
                    // two smallest of {firstD, secondD, oD1, oD2}
                    int lo   = min(firstD, oD1);
                    int hi   = max(firstD, oD1);
                    secondD  = min(hi, min(secondD, oD2));
                    firstD   = lo;
                }
                if(lane_in_group == 0 && active)
                {
                    maxlen = max(maxlen, (int)(e - s));
                    if(firstU != BIG && lastLE > firstU)
                        flags |= F_GROUP;
                    if(firstD != BIG && lastL > firstD)
                        flags |= F_GROUP;
                    if(firstD == BIG && i < n)
                        flags |= F_NODIAG;
                    if(oor != BIG)
                        first_err = min(first_err, ((unsigned long long)(s + oor) << 1) | 0ull);
                    if(secondD != BIG)
                        first_err = min(first_err, ((unsigned long long)(s + secondD) << 1) | 1ull);
                }
            }

            // CTA-level combine, then one set of atomics per CTA
            __shared__ unsigned long long s_err;
            __shared__ int                s_flags, s_min, s_max, s_len;
            if(threadIdx.x == 0)
            {
                s_err   = ~0ull;
                s_flags = 0;
                s_min   = BIG;
                s_max   = -1;
                s_len   = 0;
            }
            __syncthreads();
            // warp combine first
            for(int off = 16; off > 0; off >>= 1)
            {
                flags |= __shfl_xor_sync(0xffffffffu, flags, off);
                mincol = min(mincol, __shfl_xor_sync(0xffffffffu, mincol, off));
                maxcol = max(maxcol, __shfl_xor_sync(0xffffffffu, maxcol, off));
                maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, off));
                unsigned long long o = __shfl_xor_sync(0xffffffffu, first_err, off);
                first_err            = min(first_err, o);
            }
            if((threadIdx.x & 31) == 0)
            {
                if(first_err != ~0ull)
                    atomicMin(&s_err, first_err);
                if(flags)
                    atomicOr(&s_flags, flags);
                atomicMin(&s_min, mincol);
                atomicMax(&s_max, maxcol);
                atomicMax(&s_len, maxlen);
            }
            __syncthreads();
            if(threadIdx.x == 0)
            {
                if(s_err != ~0ull)
                    atomicMin(&acc->first_err, s_err);
                if(s_flags)
                    atomicOr(&acc->flags, s_flags);
                atomicMin(&acc->min_col, s_min);
                atomicMax(&acc->max_col, s_max);
                atomicMax(&acc->max_row_nnz, s_len);
            }
        }

        __global__ void rebase_kernel(long long n_ptr, long long nnz, aoclsparse_int *rp, aoclsparse_int *col)
        {
            long long i      = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            long long stride = (long long)gridDim.x * blockDim.x;
            for(long long k = i; k < n_ptr; k += stride)
                rp[k] -= 1;
            for(long long k = i; k < nnz; k += stride)
                col[k] -= 1;
        }
    }

    aoclsparse_status check_csr_device(aoclsparse_int        m,
                                       aoclsparse_int        n,
                                       aoclsparse_int        nnz,
                                       int                   base,
                                       const aoclsparse_int *d_row_ptr,
                                       const aoclsparse_int *d_col,
                                       check_result         &out,
                                       cudaStream_t          st)
    {
        out.status      = aoclsparse_status_success;
        out.sort        = aoclsparse_fully_sorted;
        out.fulldiag    = 1;
        out.min_col     = n;
        out.max_col     = -1;
        out.max_row_nnz = 0;
        if(m == 0)
            return aoclsparse_status_success;

        dev_buf acc_buf;
        B200_TRY(acc_buf.alloc(sizeof(check_acc)));
        check_acc init;
        init.first_err   = ~0ull;
        init.flags       = 0;
        init.min_col     = BIG;
        init.max_col     = -1;
        init.max_row_nnz = 0;
        B200_CUDA(cudaMemcpyAsync(acc_buf.p, &init, sizeof(init), cudaMemcpyHostToDevice, st));

        // lanes per row follow the mean row length
        const double    mean   = (double)nnz / (double)m;
        const int       G      = mean <= 6.0 ? 4 : (mean <= 24.0 ? 8 : 32);
        const long long groups = (long long)m;
        const int       tpb    = 256;
        long long       blocks = (groups * G + tpb - 1) / tpb;
        if(blocks > 148LL * 64)
            blocks = 148LL * 64;
        if(blocks < 1)
            blocks = 1;
        check_acc *acc = acc_buf.as<check_acc>();
        if(G == 4)
            check_rows_kernel<4><<<(unsigned)blocks, tpb, 0, st>>>(m, n, nnz, base, d_row_ptr, d_col, acc);
        else if(G == 8)
            check_rows_kernel<8><<<(unsigned)blocks, tpb, 0, st>>>(m, n, nnz, base, d_row_ptr, d_col, acc);
        else
            check_rows_kernel<32><<<(unsigned)blocks, tpb, 0, st>>>(m, n, nnz, base, d_row_ptr, d_col, acc);
        B200_LAUNCHED();

        check_acc res;
        B200_CUDA(cudaMemcpyAsync(&res, acc_buf.p, sizeof(res), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));

        if(res.flags & F_PTRDEC)
        {
            out.status = aoclsparse_status_invalid_value;
            return aoclsparse_status_success;
        }
        if(res.first_err != ~0ull)
        {
            out.status = (res.first_err & 1ull) ? aoclsparse_status_invalid_value
                                                : aoclsparse_status_invalid_index_value;
            return aoclsparse_status_success;
        }
        out.sort = (res.flags & F_GROUP)
                       ? aoclsparse_unsorted
                       : ((res.flags & F_DESCENT) ? aoclsparse_partially_sorted : aoclsparse_fully_sorted);
        out.fulldiag    = (res.flags & F_NODIAG) ? 0 : 1;
        out.min_col     = (res.min_col == BIG) ? n : res.min_col;
        out.max_col     = res.max_col;
        out.max_row_nnz = res.max_row_nnz;
        return aoclsparse_status_success;
    }

    aoclsparse_status rebase_to_zero(aoclsparse_int  m,
                                     aoclsparse_int  nnz,
                                     aoclsparse_int *d_row_ptr,
                                     aoclsparse_int *d_col,
                                     cudaStream_t    st)
    {
        long long work   = (long long)((m + 1) > nnz ? (m + 1) : nnz);
        long long blocks = (work + 255) / 256;
        if(blocks > 148 * 32)
            blocks = 148 * 32;
        if(blocks < 1)
            blocks = 1;
        rebase_kernel<<<(unsigned)blocks, 256, 0, st>>>((long long)m + 1, (long long)nnz, d_row_ptr, d_col);
        B200_LAUNCHED();
        return aoclsparse_status_success;
    }
}
