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
        int       own_lo, own_hi;            // columns of x this rank computes itself; the rest of its window is halo
        int       cta_fence_gpu;             // see boundary_release
    };

    // acquire / release fence at system scope (lighter than the sequentially consistent __threadfence_system())
    __device__ __forceinline__ void fence_acq_rel_sys()
    {
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    }

    // What a boundary CTA does between its last store and its arrival on the side's counter.  The peer must see every
    // boundary row before it sees the flag.  cta_fence_gpu == 0: every CTA fences at system scope.  cta_fence_gpu == 1:
    // a CTA only releases at GPU scope (fence + relaxed atomic = release pattern, observed by the last CTA's atomic +
    // its fence.acq_rel.sys = acquire pattern at a scope that includes both); the last CTA's system-scope fence before
    // the flag store is then cumulative over the stores of all CTAs it synchronised with (PTX memory model, causality
    // order) -- one system-scope fence per side and iteration instead of one per CTA.
    __device__ __forceinline__ void boundary_release(int cta_fence_gpu)
    {
        if(cta_fence_gpu)
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        else
            fence_acq_rel_sys();
    }

    // x[c] for a BOUNDARY row: halo entries (columns outside the rank's own rows [own_lo, own_hi)) are stored by the peer
    // GPU while this grid runs -> read at L2 (ld.global.cg), the coherence point those stores arrive at; own entries
    // take the cached path like interior rows do
    template <typename T, bool PERSISTENT>
    __device__ __forceinline__ T boundary_x(const T *x, int c, int own_lo, int own_hi)
    {
        if(c < own_lo || c >= own_hi)
            return __ldcg(x + c);
        if constexpr(PERSISTENT)
            return __ldca(x + c);
        else
            return ldg_ro(x + c);
    }

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
        fence_acq_rel_sys(); // acquire: the peer's stores that preceded its flag store are visible from here on
        return true;
    }

    // CODED: the column stream is the diagonal-code copy (one byte per entry, col = row + code_off[code]; see
    // spmv_row_blocks_kernel)
    // EC (instead of CODED): the entry-code copy (one byte per entry indexing the table of the matrix's distinct
    // (col - row, value) pairs; see spmv_row_blocks_kernel): `codes` is that copy, `code_off` / `code_val` the table
    template <typename T, bool CODED = false, bool EC = false>
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
                                                                   const int *__restrict__ code_off = nullptr,
                                                                   const T *__restrict__ code_val = nullptr,
                                                                   int n_table = CODE_TABLE_MAX)
    {
        constexpr int NT = 256;
        static_assert(!(CODED && EC), "one code stream at a time");
        using pair_t = entry_pair<T>;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        uint64_t            *bar    = reinterpret_cast<uint64_t *>(smem_raw);
        T                   *sval   = reinterpret_cast<T *>(smem_raw + SMEM_HEADER);
        aoclsparse_int      *scol   = reinterpret_cast<aoclsparse_int *>(smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T));
        const unsigned char *scode  = smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T);
        int                 *soff   = reinterpret_cast<int *>(smem_raw + SMEM_HEADER + (size_t)cap * (sizeof(T) + 1));
        const unsigned char *secode = smem_raw + SMEM_HEADER;                                          // EC
        pair_t              *stab   = reinterpret_cast<pair_t *>(smem_raw + SMEM_HEADER + (size_t)cap); // EC

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
        constexpr int GR  = (CODED || EC) ? 16 : 4; // entries per 16-byte granule of the narrowest staged array
        const int     a   = d.z & ~(GR - 1);
        const int     cnt = ((d.w - a) + GR - 1) & ~(GR - 1);
        if(tid == 0)
        {
            mbar_init(bar, 1);
            mbar_init_fence();
            if(cnt > 0)
            {
                if constexpr(EC)
                {
                    mbar_expect_tx(bar, (unsigned)cnt);
                    bulk_load_stream(const_cast<unsigned char *>(secode), codes + a, (unsigned)cnt, bar);
                }
                else if constexpr(CODED)
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
            for(int i = tid; i < n_table; i += NT)
                soff[i] = code_off[i];
        if constexpr(EC)
            for(int i = tid; i < n_table; i += NT)
            {
                pair_t pr;
                pr.v    = code_val[i];
                pr.off  = code_off[i];
                stab[i] = pr;
            }
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

        // column and value of staged entry j of row r
        auto entry_at = [&](int r, int j, int &c, T &v) {
            if constexpr(EC)
            {
                const pair_t pr = stab[secode[j]];
                c               = r + pr.off;
                v               = pr.v;
            }
            else if constexpr(CODED)
            {
                c = r + soff[scode[j]];
                v = sval[j];
            }
            else
            {
                c = scol[j];
                v = sval[j];
            }
        };
        T *push = side == 0 ? push_left : (side == 1 ? push_right : nullptr);
        const int push_row0 = side == 0 ? 0 : hc.last_row0;
        if(side == 2)
        {
            // interior rows: every x entry they name was written by the previous LOCAL launch, which has completed
            // (griddepcontrol.wait) -> read-only path through L1
            // the bounds of a thread's NEXT row are requested before its current row is reduced
            int cur_s = pre_s, cur_e = pre_e;
            for(int r = d.x + tid; r < d.y; r += NT)
            {
                int nxt_s = 0, nxt_e = 0;
                if(r + NT < d.y)
                {
                    nxt_s = rp[r + NT];
                    nxt_e = rp[r + NT + 1];
                }
                int       j = cur_s - a;
                const int e = cur_e - a;
                cur_s       = nxt_s;
                cur_e       = nxt_e;
                T          acc   = vt<T>::zero();
                for(; j + 4 <= e; j += 4)
                {
                    int c0, c1, c2, c3;
                    T   a0, a1, a2, a3;
                    entry_at(r, j, c0, a0);
                    entry_at(r, j + 1, c1, a1);
                    entry_at(r, j + 2, c2, a2);
                    entry_at(r, j + 3, c3, a3);
                    const T x0 = ldg_ro(x + c0), x1 = ldg_ro(x + c1), x2 = ldg_ro(x + c2), x3 = ldg_ro(x + c3);
                    acc        = mad(a0, x0, acc);
                    acc        = mad(a1, x1, acc);
                    acc        = mad(a2, x2, acc);
                    acc        = mad(a3, x3, acc);
                }
                if(j < e)
                {
                    // 1-3 entries left: issued together as well (an absent entry re-reads the row's first one and is not added)
                    const int n = e - j;
                    int       c0, c1, c2;
                    T         a0, a1, a2;
                    entry_at(r, j, c0, a0);
                    entry_at(r, n > 1 ? j + 1 : j, c1, a1);
                    entry_at(r, n > 2 ? j + 2 : j, c2, a2);
                    const T x0 = ldg_ro(x + c0), x1 = ldg_ro(x + c1), x2 = ldg_ro(x + c2);
                    acc        = mad(a0, x0, acc);
                    if(n > 1)
                        acc = mad(a1, x1, acc);
                    if(n > 2)
                        acc = mad(a2, x2, acc);
                }
                y[r] = mul(alpha, acc);
            }
        }
        else
        {
            // boundary rows: the halo part of x is stored by the PEER GPU while this grid is already running, so it
            // must not go through the non-coherent path (ld.global.nc requires the data to be read-only for the
            // kernel's lifetime): ld.global.cg reads at the L2, the coherence point the peer's stores arrive at
            // the bounds of a thread's NEXT row are requested before its current row is reduced
            int cur_s = pre_s, cur_e = pre_e;
            for(int r = d.x + tid; r < d.y; r += NT)
            {
                int nxt_s = 0, nxt_e = 0;
                if(r + NT < d.y)
                {
                    nxt_s = rp[r + NT];
                    nxt_e = rp[r + NT + 1];
                }
                int       j = cur_s - a;
                const int e = cur_e - a;
                cur_s       = nxt_s;
                cur_e       = nxt_e;
                T          acc   = vt<T>::zero();
                for(; j + 4 <= e; j += 4)
                {
                    int c0, c1, c2, c3;
                    T   a0, a1, a2, a3;
                    entry_at(r, j, c0, a0);
                    entry_at(r, j + 1, c1, a1);
                    entry_at(r, j + 2, c2, a2);
                    entry_at(r, j + 3, c3, a3);
                    const T x0 = boundary_x<T, false>(x, c0, hc.own_lo, hc.own_hi), x1 = boundary_x<T, false>(x, c1, hc.own_lo, hc.own_hi), x2 = boundary_x<T, false>(x, c2, hc.own_lo, hc.own_hi), x3 = boundary_x<T, false>(x, c3, hc.own_lo, hc.own_hi);
                    acc        = mad(a0, x0, acc);
                    acc        = mad(a1, x1, acc);
                    acc        = mad(a2, x2, acc);
                    acc        = mad(a3, x3, acc);
                }
                if(j < e)
                {
                    // 1-3 entries left: issued together as well (an absent entry re-reads the row's first one and is not added)
                    const int n = e - j;
                    int       c0, c1, c2;
                    T         a0, a1, a2;
                    entry_at(r, j, c0, a0);
                    entry_at(r, n > 1 ? j + 1 : j, c1, a1);
                    entry_at(r, n > 2 ? j + 2 : j, c2, a2);
                    const T x0 = boundary_x<T, false>(x, c0, hc.own_lo, hc.own_hi), x1 = boundary_x<T, false>(x, c1, hc.own_lo, hc.own_hi), x2 = boundary_x<T, false>(x, c2, hc.own_lo, hc.own_hi);
                    acc        = mad(a0, x0, acc);
                    if(n > 1)
                        acc = mad(a1, x1, acc);
                    if(n > 2)
                        acc = mad(a2, x2, acc);
                }
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
                boundary_release(hc.cta_fence_gpu); // this CTA's stores (peer stores included) before the counter / flag
                const unsigned n_side = side == 0 ? (unsigned)hc.n_first : (unsigned)hc.n_last;
                const unsigned done   = atomicAdd(hc.counters + side, 1u) + 1u;
                if(done == hc.kc * n_side)
                {
                    unsigned *flag = side == 0 ? hc.to_left_done : hc.to_right_done;
                    if(flag)
                    {
                        fence_acq_rel_sys();
                        *reinterpret_cast<volatile unsigned *>(flag) = hc.k;
                    }
                }
            }
        }
    }

    // ---------------------------------------------------------------------------------------------------------------
    // k iterations in ONE launch: a persistent grid (every CTA resident: cooperative launch) walks the row blocks of
    // every iteration itself, so an iteration costs neither a launch nor the drain / refill of the device between two
    // grids.  Per iteration:
    //   * CTA c takes blocks c, c+G, c+2G, ... of the order [first boundary | last boundary | interior]; the boundary
    //     blocks behave exactly as in spmv_sharded_step_kernel (flag wait, peer stores, last one publishes the flag);
    //   * the matrix slice of the NEXT block -- also when it belongs to the next iteration -- is requested (TMA bulk
    //     copies) as soon as the staging buffer is free, i.e. before the grid barrier: the barrier's latency and the
    //     refill of the pipeline overlap;
    //   * between iterations the grid meets at one counter in device memory (arrive: fence + atomicAdd; wait: spin on a
    //     volatile load, then fence.acq_rel.gpu, which also invalidates this SM's L1).  x of iteration k+1 is what other
    //     CTAs stored during iteration k, so the gathers use ld.global.ca (coherent after the acquire) instead of the
    //     non-coherent path; halo entries, stored by the peer GPU while this iteration runs, are read at L2 (ld.cg).
    // Results are bit-identical to k launches of the step kernel (same blocks, same per-row order of operations).
    // Plain and diagonal-coded streams only.  ENTRY-CODED shards (one staged byte per entry) run one step kernel per
    // iteration instead: that multiply is bound by the L1 pipeline, not by HBM, and hardware-scheduled CTAs of the step
    // kernel beat this statically scheduled grid by 25-35 % there (2 GPUs, 7-point 512^3: 0.440 ms per iteration against
    // 0.551-0.677 ms; 64 planes per rank: 0.122 against 0.145 ms; profiles/r02_sharded_entry_codes.txt).
    struct iterate_ctl
    {
        const unsigned *left_done, *right_done;
        unsigned       *to_left_done, *to_right_done;
        unsigned       *counters; // [0] / [1] boundary CTAs done, [2] grid barrier arrivals, [3] a wait gave up
        unsigned        k0;       // flag value of this launch's first iteration
        unsigned        kc0;      // iterations run on these counters before this launch
        unsigned        bar0;     // value of counters[2] before this launch
        int             iters;
        int             n_first, n_last, n_blocks, last_begin, last_row0;
        int             own_lo, own_hi; // columns of x this rank computes itself; the rest of its window is halo
        int             cta_fence_gpu;  // see boundary_release
    };

    template <typename T, bool CODED>
    __global__ void __launch_bounds__(256) spmv_sharded_iterate_kernel(const int4 *__restrict__ desc,
                                                                      int cap,
                                                                      const aoclsparse_int *__restrict__ rp,
                                                                      const aoclsparse_int *__restrict__ col,
                                                                      const T *__restrict__ val,
                                                                      const T *x0, // x of even iterations (0, 2, ...)
                                                                      const T *x1, // x of odd iterations
                                                                      T       *y0, // y of even iterations (own rows of x1's window)
                                                                      T       *y1,
                                                                      T        alpha,
                                                                      T       *push_left0,
                                                                      T       *push_left1,
                                                                      T       *push_right0,
                                                                      T       *push_right1,
                                                                      iterate_ctl hc,
                                                                      const unsigned char *__restrict__ codes,
                                                                      const int *__restrict__ code_off,
                                                                      int n_table)
    {
        constexpr int NT = 256;
        extern __shared__ __align__(16) unsigned char smem_raw[];
        uint64_t            *bar    = reinterpret_cast<uint64_t *>(smem_raw);
        T                   *sval   = reinterpret_cast<T *>(smem_raw + SMEM_HEADER);
        aoclsparse_int      *scol   = reinterpret_cast<aoclsparse_int *>(smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T));
        const unsigned char *scode  = smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T);
        int                 *soff   = reinterpret_cast<int *>(smem_raw + SMEM_HEADER + (size_t)cap * (sizeof(T) + 1));
        constexpr int        GR     = CODED ? 16 : 4;

        const int tid = threadIdx.x;
        const int G   = (int)gridDim.x;

        // launch order -> (block, side); side 0 first boundary, 1 last boundary, 2 interior
        auto locate = [&](int bid, int &b, int &side) {
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
        };
        // thread 0: request the matrix slice of block descriptor d into the staging buffer
        auto request = [&](const int4 &d) {
            const int a   = d.z & ~(GR - 1);
            const int cnt = ((d.w - a) + GR - 1) & ~(GR - 1);
            if(cnt <= 0)
                return;
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
        };
        auto entry_at = [&](int r, int j, int &c, T &v) {
            if constexpr(CODED)
            {
                c = r + soff[scode[j]];
                v = sval[j];
            }
            else
            {
                c = scol[j];
                v = sval[j];
            }
        };

        if(tid == 0)
        {
            mbar_init(bar, 1);
            mbar_init_fence();
        }
        if constexpr(CODED)
            for(int i = tid; i < n_table; i += NT)
                soff[i] = code_off[i];
        __syncthreads();
        if((int)blockIdx.x >= hc.n_blocks || hc.iters <= 0)
            return; // host launches G <= n_blocks
        // This CTA's blocks of one iteration, in launch-order numbering (see locate): c, c+G, c+2G, ...  (Giving every CTA one
        // contiguous chunk of interior blocks instead -- for L1 reuse of x between consecutive grid lines -- measured SLOWER,
        // 0.600 against 0.575 ms per iteration on 2 GPUs: a thousand scattered streams cost more in HBM than the reuse saves.)
        const int c_id   = (int)blockIdx.x;
        const int n_mine = hc.n_blocks > c_id ? (hc.n_blocks - c_id + G - 1) / G : 0;
        auto      nth    = [&](int i) -> int { return c_id + i * G; };
        // (a CTA without blocks -- possible only when there are about as many CTAs as blocks -- still takes part in
        // every grid barrier below)
        unsigned parity = 0;
        int      b = 0, side = 2;
        int4     d = make_int4(0, 0, 0, 0);
        if(n_mine > 0)
        {
            locate(nth(0), b, side);
            d = desc[b];
            if(tid == 0)
                request(d);
        }

        for(int it = 0; it < hc.iters; ++it)
        {
            const T *x          = (it & 1) ? x1 : x0;
            T       *y          = (it & 1) ? y1 : y0;
            T       *push_left  = (it & 1) ? push_left1 : push_left0;
            T       *push_right = (it & 1) ? push_right1 : push_right0;
            const unsigned k    = hc.k0 + (unsigned)it;
            for(int ib = 0; ib < n_mine; ++ib)
            {
                // (b, side, d) describe this CTA's ib-th block; its slice has been requested
                const int a   = d.z & ~(GR - 1);
                const int cnt = ((d.w - a) + GR - 1) & ~(GR - 1);
                // next block of this CTA: in this iteration, or the first one of the next iteration
                const bool wraps = ib + 1 >= n_mine;
                const bool more  = !wraps || it + 1 < hc.iters;
                int        bn = b, side_n = side;
                int4       dn = d;
                if(more)
                {
                    locate(wraps ? nth(0) : nth(ib + 1), bn, side_n);
                    dn = desc[bn];
                }
                int       pre_s = 0, pre_e = 0;
                if(d.x + tid < d.y)
                {
                    pre_s = rp[d.x + tid];
                    pre_e = rp[d.x + tid + 1];
                }
                if(side != 2)
                {
                    if(tid == 0)
                    {
                        const unsigned *done = side == 0 ? hc.left_done : hc.right_done;
                        if(done)
                            spin_until(done, k - 1, hc.counters + 3);
                    }
                    __syncthreads();
                }
                if(cnt > 0)
                {
                    mbar_wait(bar, parity);
                    parity ^= 1u;
                }
                if(side == 2)
                {
                    // the bounds of a thread's NEXT row are requested before its current row is reduced
                    int cur_s = pre_s, cur_e = pre_e;
                    for(int r = d.x + tid; r < d.y; r += NT)
                    {
                        int nxt_s = 0, nxt_e = 0;
                        if(r + NT < d.y)
                        {
                            nxt_s = rp[r + NT];
                            nxt_e = rp[r + NT + 1];
                        }
                        int       j = cur_s - a;
                        const int e = cur_e - a;
                        cur_s       = nxt_s;
                        cur_e       = nxt_e;
                        T          acc   = vt<T>::zero();
                        for(; j + 4 <= e; j += 4)
                        {
                            int c0, c1, c2, c3;
                            T   a0, a1, a2, a3;
                            entry_at(r, j, c0, a0);
                            entry_at(r, j + 1, c1, a1);
                            entry_at(r, j + 2, c2, a2);
                            entry_at(r, j + 3, c3, a3);
                            const T x0 = __ldca(x + c0), x1 = __ldca(x + c1), x2 = __ldca(x + c2), x3 = __ldca(x + c3);
                            acc        = mad(a0, x0, acc);
                            acc        = mad(a1, x1, acc);
                            acc        = mad(a2, x2, acc);
                            acc        = mad(a3, x3, acc);
                        }
                        if(j < e)
                        {
                            // 1-3 entries left: issued together as well (an absent entry re-reads the row's first one and is not added)
                            const int n = e - j;
                            int       c0, c1, c2;
                            T         a0, a1, a2;
                            entry_at(r, j, c0, a0);
                            entry_at(r, n > 1 ? j + 1 : j, c1, a1);
                            entry_at(r, n > 2 ? j + 2 : j, c2, a2);
                            const T x0 = __ldca(x + c0), x1 = __ldca(x + c1), x2 = __ldca(x + c2);
                            acc        = mad(a0, x0, acc);
                            if(n > 1)
                                acc = mad(a1, x1, acc);
                            if(n > 2)
                                acc = mad(a2, x2, acc);
                        }
                        y[r] = mul(alpha, acc);
                    }
                }
                else
                {
                    T        *push      = side == 0 ? push_left : push_right;
                    const int push_row0 = side == 0 ? 0 : hc.last_row0;
                    // the bounds of a thread's NEXT row are requested before its current row is reduced
                    int cur_s = pre_s, cur_e = pre_e;
                    for(int r = d.x + tid; r < d.y; r += NT)
                    {
                        int nxt_s = 0, nxt_e = 0;
                        if(r + NT < d.y)
                        {
                            nxt_s = rp[r + NT];
                            nxt_e = rp[r + NT + 1];
                        }
                        int       j = cur_s - a;
                        const int e = cur_e - a;
                        cur_s       = nxt_s;
                        cur_e       = nxt_e;
                        T          acc   = vt<T>::zero();
                        for(; j + 4 <= e; j += 4)
                        {
                            int c0, c1, c2, c3;
                            T   a0, a1, a2, a3;
                            entry_at(r, j, c0, a0);
                            entry_at(r, j + 1, c1, a1);
                            entry_at(r, j + 2, c2, a2);
                            entry_at(r, j + 3, c3, a3);
                            const T x0 = boundary_x<T, true>(x, c0, hc.own_lo, hc.own_hi), x1 = boundary_x<T, true>(x, c1, hc.own_lo, hc.own_hi), x2 = boundary_x<T, true>(x, c2, hc.own_lo, hc.own_hi), x3 = boundary_x<T, true>(x, c3, hc.own_lo, hc.own_hi);
                            acc        = mad(a0, x0, acc);
                            acc        = mad(a1, x1, acc);
                            acc        = mad(a2, x2, acc);
                            acc        = mad(a3, x3, acc);
                        }
                        if(j < e)
                        {
                            // 1-3 entries left: issued together as well (an absent entry re-reads the row's first one and is not added)
                            const int n = e - j;
                            int       c0, c1, c2;
                            T         a0, a1, a2;
                            entry_at(r, j, c0, a0);
                            entry_at(r, n > 1 ? j + 1 : j, c1, a1);
                            entry_at(r, n > 2 ? j + 2 : j, c2, a2);
                            const T x0 = boundary_x<T, true>(x, c0, hc.own_lo, hc.own_hi), x1 = boundary_x<T, true>(x, c1, hc.own_lo, hc.own_hi), x2 = boundary_x<T, true>(x, c2, hc.own_lo, hc.own_hi);
                            acc        = mad(a0, x0, acc);
                            if(n > 1)
                                acc = mad(a1, x1, acc);
                            if(n > 2)
                                acc = mad(a2, x2, acc);
                        }
                        const T out = mul(alpha, acc);
                        y[r]        = out;
                        if(push)
                            push[r - push_row0] = out;
                    }
                }
                __syncthreads(); // the staging buffer is free; this CTA's stores are ordered before thread 0's fences
                const int  done_side = side;
                if(more)
                {
                    b    = bn;
                    side = side_n;
                    d    = dn;
                    if(tid == 0)
                        request(d);
                }
                if(done_side != 2 && tid == 0)
                {
                    boundary_release(hc.cta_fence_gpu); // this CTA's stores (peer stores included) before the counter / flag
                    const unsigned n_side = done_side == 0 ? (unsigned)hc.n_first : (unsigned)hc.n_last;
                    const unsigned cntd   = atomicAdd(hc.counters + done_side, 1u) + 1u;
                    if(cntd == (hc.kc0 + (unsigned)it + 1u) * n_side)
                    {
                        unsigned *flag = done_side == 0 ? hc.to_left_done : hc.to_right_done;
                        if(flag)
                        {
                            fence_acq_rel_sys();
                            *reinterpret_cast<volatile unsigned *>(flag) = k;
                        }
                    }
                }
            }
            if(it + 1 < hc.iters)
            {
                // grid barrier: every y of this iteration is stored before any CTA gathers it as the next x
                if(tid == 0)
                {
                    __threadfence();
                    atomicAdd(hc.counters + 2, 1u);
                    const unsigned           target = hc.bar0 + (unsigned)(it + 1) * (unsigned)G;
                    const volatile unsigned *c      = hc.counters + 2;
                    long long                spins  = 0;
                    while((int)(*c - target) < 0)
                    {
                        __nanosleep(32);
                        if(++spins > 100000000LL)
                        {
                            hc.counters[3] = 1;
                            break;
                        }
                    }
                    __threadfence(); // acquire: also drops this SM's L1 lines of the buffer that was just rewritten
                }
                __syncthreads();
            }
        }
    }
}
