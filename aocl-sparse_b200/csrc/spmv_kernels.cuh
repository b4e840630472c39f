// spmv_kernels.cuh -- sm_100a kernels for y = alpha * op(A) * x + beta * y on a device CSR.
//
// They replace the reference's OpenMP row-parallel CPU kernels
//   ref_csrmv_gn / aoclsparse_csrmv_vectorized{,_avx2}   library/src/level2/aoclsparse_csrmv_kr.hpp:448-513,734-831,949-1040
//   aoclsparse_csrmv_vectorized_avx512                    library/src/level2/aoclsparse_csrmv_avx512.cpp:36-134
//   csrmv_kt / csrmvt_kt / csrmv_symm_kt                  library/src/level2/aoclsparse_csrmv_kt.cpp:30-329
//   ref_csrmv_th / ref_csrmv_tri / ref_csrmv_tri_th       library/src/level2/aoclsparse_csrmv_kr.hpp:520-728
// with a different decomposition (this is not a translation of them):
//
//   * one CTA per ROW BLOCK of the plan (plan.cu): consecutive rows holding <= T stored entries, or
//     one T-entry segment of a row longer than T.  Every CTA therefore streams the same number of
//     bytes whatever the row-length distribution is.
//   * the block's slice of val[] and col_idx[] is brought into shared memory by two TMA bulk copies
//     (cp.async.bulk, 16-byte granules, completion on an mbarrier): HBM is read in full, aligned,
//     contiguous bursts that are independent of how the rows are later walked.
//   * rows are then reduced out of shared memory with the strategy the analysis binned the block
//     into: thread-per-row (lanes = consecutive rows, so x[col] of a banded matrix is read
//     coalesced), warp-per-row, or CTA-wide products followed by per-row sums (short rows by one
//     lane, long rows by the whole warp).  Sums are formed in a fixed order: results are
//     run-to-run reproducible.
//   * a row split over several CTAs leaves one partial sum per segment; finish_long_rows_kernel adds
//     them in segment order and applies alpha / beta.
//   * x is read through the read-only path (ld.global.nc); it stays L2 resident while val/col stream.
//   * beta == 0 never reads y (the reference's NaN-in-y rule, aoclsparse_csrmv_kr.hpp:504-509).
#pragma once
#include "common.hpp"

namespace b200
{
    // ---- element masks for triangular / symmetric / hermitian descriptors ------------------
    enum : int
    {
        MASK_NONE  = 0,
        MASK_LOWER = 1, // keep col <= row
        MASK_UPPER = 2, // keep col >= row
        MASK_DIAG  = 3  // keep col == row
    };
    enum : int
    {
        DIAG_KEEP = 0, // stored diagonal entries take part
        DIAG_UNIT = 1, // stored diagonal entries are skipped, 1 is used instead
        DIAG_ZERO = 2  // stored diagonal entries are skipped
    };
    struct elem_rule
    {
        int mask; // MASK_*
        int diag; // DIAG_*
        int conj; // conjugate stored off-diagonal values
        int conj_diag; // conjugate stored diagonal values
    };

    __device__ __forceinline__ bool keep_entry(const elem_rule &rl, int row, int col)
    {
        if(rl.mask == MASK_LOWER && col > row)
            return false;
        if(rl.mask == MASK_UPPER && col < row)
            return false;
        if(rl.mask == MASK_DIAG && col != row)
            return false;
        if(rl.diag != DIAG_KEEP && col == row)
            return false;
        return true;
    }

    // ---- TMA bulk copy + mbarrier primitives (PTX ISA: cp.async.bulk, mbarrier) ---------------
    __device__ __forceinline__ uint32_t smem_u32(const void *p)
    {
        return (uint32_t)__cvta_generic_to_shared(p);
    }
    __device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    }
    __device__ __forceinline__ void mbar_init_fence()
    {
        // make the initialised barrier visible to the async (TMA) proxy
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                     : "memory");
    }
    __device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
    {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(dst_smem)),
                     "l"(src_gmem),
                     "r"(bytes),
                     "r"(smem_u32(bar))
                     : "memory");
    }
    // same copy, tagged evict-first in L2: the matrix stream is read once per multiply and must not push
    // x (which every CTA re-reads) out of the 126 MB L2
    __device__ __forceinline__ void bulk_load_stream(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar)
    {
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                         smem_u32(dst_smem)),
                     "l"(src_gmem),
                     "r"(bytes),
                     "r"(smem_u32(bar)),
                     "l"(policy)
                     : "memory");
    }
    __device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
    {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "WAIT_LOOP:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                     "@p bra WAIT_DONE;\n"
                     "bra WAIT_LOOP;\n"
                     "WAIT_DONE:\n"
                     "}\n" ::"r"(smem_u32(bar)),
                     "r"(parity)
                     : "memory");
    }

    template <typename T>
    __device__ __forceinline__ T ldg_ro(const T *p)
    {
        return __ldg(p);
    }

    // acc + op(v) * x[col] if the entry (row, col) takes part under `rl`, else acc
    template <typename T>
    __device__ __forceinline__ T generic_term(const elem_rule &rl, int row, int col, T v, const T *__restrict__ x, T acc)
    {
        if(!keep_entry(rl, row, col))
            return acc;
        if(col == row ? rl.conj_diag : rl.conj)
            v = cj(v);
        return mad(v, ldg_ro(x + col), acc);
    }

    template <typename T>
    __device__ __forceinline__ T warp_sum(T v)
    {
#pragma unroll
        for(int off = 16; off > 0; off >>= 1)
        {
            if constexpr(vt<T>::is_complex)
            {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
                v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
            }
            else
                v += __shfl_xor_sync(0xffffffffu, v, off);
        }
        return v;
    }

    template <typename T>
    __device__ __forceinline__ T axpby_out(T alpha, T acc, T beta, bool beta_zero, const T *y_in)
    {
        T r = mul(alpha, acc);
        if(!beta_zero)
            r = mad(beta, *y_in, r);
        return r;
    }

    constexpr int SPMV_THREADS = 256;
    constexpr int SMEM_HEADER  = 16; // mbarrier + padding, keeps the staged arrays 16-byte aligned
    constexpr int SHORT_ROW    = 32; // PRODUCT strategy: rows up to this length are summed by one lane

    inline size_t spmv_smem_bytes(size_t elem_size, aoclsparse_int block_nnz)
    {
        return SMEM_HEADER + (size_t)(block_nnz + 8) * (elem_size + 4);
    }
    // CODED variant: values, one code byte per entry (16-entry alignment slack at both ends), the offset table
    constexpr int CODE_TABLE = CODE_TABLE_MAX;
    inline size_t spmv_coded_smem_bytes(size_t elem_size, aoclsparse_int block_nnz)
    {
        return SMEM_HEADER + (size_t)(block_nnz + 32) * (elem_size + 1) + CODE_TABLE * sizeof(int);
    }

    // ECODED variant (see spmv_row_blocks_kernel): one code byte per entry, then the table of (value, col - row) pairs
    template <typename T>
    struct alignas(sizeof(T) >= 8 ? 16 : 8) entry_pair
    {
        T   v;
        int off;
    };
    inline int spmv_ecoded_cap(aoclsparse_int block_nnz)
    {
        return ((int)block_nnz + 32 + 15) & ~15;
    }
    inline size_t spmv_ecoded_smem_bytes(size_t elem_size, aoclsparse_int block_nnz)
    {
        return SMEM_HEADER + (size_t)spmv_ecoded_cap(block_nnz) + (size_t)CODE_TABLE * (elem_size >= 8 ? 16 : 8);
    }

    // GENERIC == false: general matrix, no conjugation (the measured hot path)
    // GENERIC == true : entries filtered / conjugated by `rule` (triangular, symmetric, hermitian parts)
    // PUSH == true: every computed y[r] is also stored to push_dst[r - push_row0], a buffer that may live in a
    //                PEER GPU's memory (mapped over NVLink): the boundary rows of a row-sharded matrix write the
    //                neighbour's halo of the next x directly from this epilogue -- no separate copy or collective.
    // CODED == true (never with GENERIC; every block thread-per-row): the column stream is the DIAGONAL-CODE copy
    //                built by aoclsparse_optimize (plan.cu, build_diag_codes): one byte per stored entry, an index
    //                into the table of the matrix's distinct (col - row) offsets, so col = row + code_off[code] --
    //                the same column, bit for bit, from a quarter of the bytes.  `col` is then unused.
    // ECODED == true (instead of CODED; 4- and 8-byte value types): the ENTRY-CODE copy (plan.cu, build_entry_codes): ONE
    //                byte per stored entry, an index into the table of the matrix's distinct (col - row, value) pairs
    //                (<= 256; a constant-coefficient stencil has as many as it has points), so col = row + pair.off and
    //                val = pair.v -- the same column and the same value, bit for bit, from 1 instead of 4 + sizeof(T)
    //                bytes.  `codes` is then that copy, `code_off` / `code_val` the table; `col` and `val` are unused.
    template <typename T, bool GENERIC, int NT, bool PUSH = false, bool CODED = false, bool ECODED = false>
    __global__ void __launch_bounds__(NT, (ECODED && NT <= 256 && std::is_same<T, double>::value) ? 2048 / NT : 0) spmv_row_blocks_kernel(const int4 *__restrict__ desc,
                                                                          const int *__restrict__ kind,
                                                                          int block_first,
                                                                          int cap, // staged capacity in entries
                                                                          const aoclsparse_int *__restrict__ rp,
                                                                          const aoclsparse_int *__restrict__ col,
                                                                          const T *__restrict__ val,
                                                                          const T *__restrict__ x,
                                                                          T *__restrict__ y,
                                                                          T         alpha,
                                                                          T         beta,
                                                                          int       beta_zero,
                                                                          T        *partials,
                                                                          elem_rule rule,
                                                                          int       n_cols,
                                                                          int       stream_hint,
                                                                          T        *push_dst  = nullptr,
                                                                          int       push_row0 = 0,
                                                                          const unsigned char *__restrict__ codes = nullptr,
                                                                          const int *__restrict__ code_off = nullptr,
                                                                          const T *__restrict__ code_val = nullptr,
                                                                          int n_table = CODE_TABLE_MAX)
    {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        uint64_t       *bar  = reinterpret_cast<uint64_t *>(smem_raw);
        T              *sval = reinterpret_cast<T *>(smem_raw + SMEM_HEADER);
        aoclsparse_int *scol = reinterpret_cast<aoclsparse_int *>(smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T));

        if constexpr(ECODED)
        {
            static_assert(!GENERIC && !CODED && sizeof(T) <= 8, "entry codes: plain general product, 4- / 8-byte values");
            using pair_t                = entry_pair<T>;
            const unsigned char *secode = smem_raw + SMEM_HEADER;
            pair_t              *stab   = reinterpret_cast<pair_t *>(smem_raw + SMEM_HEADER + (size_t)cap);
            const int            tid    = threadIdx.x;
            const int4           d      = desc[blockIdx.x + block_first];
            asm volatile("griddepcontrol.launch_dependents;");
            // staged window [a, a+cnt): 16-entry granules keep the bulk copy 16-byte aligned
            const int a   = d.z & ~15;
            const int cnt = ((d.w - a) + 15) & ~15;
            if(tid == 0)
            {
                mbar_init(bar, 1);
                mbar_init_fence();
                if(cnt > 0)
                {
                    mbar_expect_tx(bar, (unsigned)cnt);
                    if(stream_hint)
                        bulk_load_stream(const_cast<unsigned char *>(secode), codes + a, (unsigned)cnt, bar);
                    else
                        bulk_load(const_cast<unsigned char *>(secode), codes + a, (unsigned)cnt, bar);
                }
            }
            for(int i = tid; i < n_table; i += NT)
            {
                pair_t pr;
                pr.v    = code_val[i];
                pr.off  = code_off[i];
                stab[i] = pr;
            }
            __syncthreads();
            // bounds (and y for beta != 0) of a thread's NEXT row are requested before its current row is reduced
            int r     = d.x + tid;
            int cur_s = 0, cur_e = 0;
            T   cur_y = vt<T>::zero();
            if(r < d.y)
            {
                cur_s = rp[r];
                cur_e = rp[r + 1];
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            if(r < d.y && !beta_zero)
                cur_y = y[r];
            if(cnt > 0)
                mbar_wait(bar, 0);
            while(r < d.y)
            {
                const int rn    = r + NT;
                int       nxt_s = 0, nxt_e = 0;
                T         nxt_y = vt<T>::zero();
                if(rn < d.y)
                {
                    nxt_s = rp[rn];
                    nxt_e = rp[rn + 1];
                    if(!beta_zero)
                        nxt_y = y[rn];
                }
                int       j   = cur_s - a;
                const int e   = cur_e - a;
                T         acc = vt<T>::zero();
                const T  *xr  = x + r; // x[col] = xr[col - row]
                // Rows of >= 12 entries read their code bytes as aligned 32-bit words (the staged buffer is 16-byte aligned
                // and padded): four codes = funnel shift of two neighbouring words by the row's byte phase, the second
                // word of one group being the first of the next -- one word load per four entries instead of four byte
                // loads: 27-point 128^3 50.8 -> 42.7 us.  Short rows keep the byte loads (the shift chain costs a 5- or
                // 7-entry row more latency than the loads it saves: 7-point 512^3 0.903 -> 0.986 ms;
                // profiles/r02_entry_codes.txt).  Same codes either way, hence the same bits.  (8-byte values only: with
                // 4-byte values both paths do not fit the 32 registers that 8 resident CTAs leave.)
                if(sizeof(T) == 8 && e - j >= 12)
                {
                    const unsigned *cw = reinterpret_cast<const unsigned *>(secode) + (j >> 2);
                    const unsigned  sh = (unsigned)(j & 3) * 8u;
                    unsigned        lo = cw[0];
                    for(; j + 4 <= e; j += 4)
                    {
                        const unsigned hi = *++cw;
                        const unsigned c4 = __funnelshift_r(lo, hi, sh);
                        lo                = hi;
                        const pair_t p0 = stab[c4 & 255u], p1 = stab[(c4 >> 8) & 255u], p2 = stab[(c4 >> 16) & 255u], p3 = stab[c4 >> 24];
                        const T      x0 = ldg_ro(xr + p0.off), x1 = ldg_ro(xr + p1.off), x2 = ldg_ro(xr + p2.off), x3 = ldg_ro(xr + p3.off);
                        acc             = mad(p0.v, x0, acc);
                        acc             = mad(p1.v, x1, acc);
                        acc             = mad(p2.v, x2, acc);
                        acc             = mad(p3.v, x3, acc);
                    }
                }
                for(; j + 4 <= e; j += 4)
                {
                    const pair_t p0 = stab[secode[j]], p1 = stab[secode[j + 1]], p2 = stab[secode[j + 2]], p3 = stab[secode[j + 3]];
                    const T      x0 = ldg_ro(xr + p0.off), x1 = ldg_ro(xr + p1.off), x2 = ldg_ro(xr + p2.off), x3 = ldg_ro(xr + p3.off);
                    acc             = mad(p0.v, x0, acc);
                    acc             = mad(p1.v, x1, acc);
                    acc             = mad(p2.v, x2, acc);
                    acc             = mad(p3.v, x3, acc);
                }
                if(j < e)
                {
                    // 1-3 entries left: issued together as well
                    const int    n  = e - j;
                    const pair_t p0 = stab[secode[j]];
                    const pair_t p1 = stab[n > 1 ? secode[j + 1] : 0];
                    const pair_t p2 = stab[n > 2 ? secode[j + 2] : 0];
                    const T      x0 = ldg_ro(xr + p0.off);
                    const T      x1 = n > 1 ? ldg_ro(xr + p1.off) : vt<T>::zero();
                    const T      x2 = n > 2 ? ldg_ro(xr + p2.off) : vt<T>::zero();
                    acc             = mad(p0.v, x0, acc);
                    if(n > 1)
                        acc = mad(p1.v, x1, acc);
                    if(n > 2)
                        acc = mad(p2.v, x2, acc);
                }
                const T out = axpby_out(alpha, acc, beta, beta_zero != 0, &cur_y);
                y[r]        = out;
                if constexpr(PUSH)
                    push_dst[r - push_row0] = out;
                r     = rn;
                cur_s = nxt_s;
                cur_e = nxt_e;
                cur_y = nxt_y;
            }
            return;
        }

        if constexpr(CODED)
        {
            static_assert(!GENERIC, "the diagonal-code copy serves the plain general product only");
            const unsigned char *scode = smem_raw + SMEM_HEADER + (size_t)cap * sizeof(T);
            int                 *soff  = reinterpret_cast<int *>(smem_raw + SMEM_HEADER + (size_t)cap * (sizeof(T) + 1));
            const int            tid   = threadIdx.x;
            const int4           d     = desc[blockIdx.x + block_first];
            asm volatile("griddepcontrol.launch_dependents;");
            // staged window [a, a+cnt): 16-entry granules keep both bulk copies 16-byte aligned
            const int a   = d.z & ~15;
            const int cnt = ((d.w - a) + 15) & ~15;
            if(tid == 0)
            {
                mbar_init(bar, 1);
                mbar_init_fence();
                if(cnt > 0)
                {
                    mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + 1)));
                    if(stream_hint)
                    {
                        bulk_load_stream(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                        bulk_load_stream(const_cast<unsigned char *>(scode), codes + a, (unsigned)cnt, bar);
                    }
                    else
                    {
                        bulk_load(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                        bulk_load(const_cast<unsigned char *>(scode), codes + a, (unsigned)cnt, bar);
                    }
                }
            }
            for(int i = tid; i < n_table; i += NT)
                soff[i] = code_off[i];
            __syncthreads();
            int pre_s = 0, pre_e = 0;
            T   pre_y = vt<T>::zero();
            if(d.x + tid < d.y)
            {
                pre_s = rp[d.x + tid];
                pre_e = rp[d.x + tid + 1];
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            if(d.x + tid < d.y && !beta_zero)
                pre_y = y[d.x + tid];
            if(cnt > 0)
                mbar_wait(bar, 0);
            for(int r = d.x + tid; r < d.y; r += NT)
            {
                const bool first = r == d.x + tid;
                int        j     = (first ? pre_s : rp[r]) - a;
                const int  e     = (first ? pre_e : rp[r + 1]) - a;
                T          acc   = vt<T>::zero();
                const T   *xr    = x + r; // x[col] = xr[col - row]
                for(; j + 4 <= e; j += 4)
                {
                    const int o0 = soff[scode[j]], o1 = soff[scode[j + 1]], o2 = soff[scode[j + 2]], o3 = soff[scode[j + 3]];
                    const T   x0 = ldg_ro(xr + o0), x1 = ldg_ro(xr + o1), x2 = ldg_ro(xr + o2), x3 = ldg_ro(xr + o3);
                    acc          = mad(sval[j], x0, acc);
                    acc          = mad(sval[j + 1], x1, acc);
                    acc          = mad(sval[j + 2], x2, acc);
                    acc          = mad(sval[j + 3], x3, acc);
                }
                for(; j < e; ++j)
                    acc = mad(sval[j], ldg_ro(xr + soff[scode[j]]), acc);
                const T out = axpby_out(alpha, acc, beta, beta_zero != 0, first ? &pre_y : y + r);
                y[r]        = out;
                if constexpr(PUSH)
                    push_dst[r - push_row0] = out;
            }
            return;
        }

        const int  tid  = threadIdx.x;
        const int  lane = tid & 31;
        const int  warp = tid >> 5;
        const int  b    = blockIdx.x + block_first;
        const int4 d    = desc[b];
        const int  k    = kind[b];
        const int  strat = k & 15;

        // programmatic dependent launch: let the next kernel in the stream start its own prologue (descriptor
        // read, bulk copies of ITS matrix slice -- none of which depends on this kernel's output) while this grid
        // drains; see griddepcontrol.wait below
        asm volatile("griddepcontrol.launch_dependents;");

        // staged window: [a, a+cnt) with a 16-byte aligned for both arrays
        const int ns = d.z, ne = d.w;
        const int a   = ns & ~3;
        const int cnt = ((ne - a) + 3) & ~3;

        if(tid == 0)
        {
            mbar_init(bar, 1);
            mbar_init_fence();
            if(cnt > 0)
            {
                mbar_expect_tx(bar, (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                if(stream_hint)
                {
                    bulk_load_stream(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                    bulk_load_stream(scol, col + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar);
                }
                else
                {
                    bulk_load(sval, val + a, (unsigned)(cnt * sizeof(T)), bar);
                    bulk_load(scol, col + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar);
                }
            }
        }
        __syncthreads();
        // the first row's bounds (and its y for beta != 0) are fetched while the bulk copies are in flight
        int  pre_s = 0, pre_e = 0;
        T    pre_y = vt<T>::zero();
        if(strat == STRAT_THREAD && d.x + tid < d.y)
        {
            pre_s = rp[d.x + tid];
            pre_e = rp[d.x + tid + 1];
        }
        // everything above touched only the matrix; x and y may be the previous kernel's output (iterated
        // products), so wait here until the grid this launch depends on has completed and flushed
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if(strat == STRAT_THREAD && d.x + tid < d.y && !beta_zero)
            pre_y = y[d.x + tid];
        if(cnt > 0)
            mbar_wait(bar, 0);

        if(strat == STRAT_THREAD)
        {
            for(int r = d.x + tid; r < d.y; r += NT)
            {
                const bool first = r == d.x + tid;
                int        j     = (first ? pre_s : rp[r]) - a;
                const int  e     = (first ? pre_e : rp[r + 1]) - a;
                T          acc   = vt<T>::zero();
                if constexpr(!GENERIC)
                {
                    for(; j + 4 <= e; j += 4)
                    {
                        const int c0 = scol[j], c1 = scol[j + 1], c2 = scol[j + 2], c3 = scol[j + 3];
                        const T   x0 = ldg_ro(x + c0), x1 = ldg_ro(x + c1), x2 = ldg_ro(x + c2),
                                  x3 = ldg_ro(x + c3);
                        acc          = mad(sval[j], x0, acc);
                        acc          = mad(sval[j + 1], x1, acc);
                        acc          = mad(sval[j + 2], x2, acc);
                        acc          = mad(sval[j + 3], x3, acc);
                    }
                    for(; j < e; ++j)
                        acc = mad(sval[j], ldg_ro(x + scol[j]), acc);
                }
                else
                {
                    for(; j < e; ++j)
                        acc = generic_term(rule, r, scol[j], sval[j], x, acc);
                    if(rule.diag == DIAG_UNIT && r < n_cols)
                        acc = add(acc, ldg_ro(x + r));
                }
                const T out = axpby_out(alpha, acc, beta, beta_zero != 0, first ? &pre_y : y + r);
                y[r]        = out;
                if constexpr(PUSH)
                    push_dst[r - push_row0] = out;
            }
        }
        else if(strat == STRAT_WARP)
        {
            for(int r = d.x + warp; r < d.y; r += NT / 32)
            {
                const int s = rp[r] - a, e = rp[r + 1] - a;
                T         acc = vt<T>::zero();
                for(int j = s + lane; j < e; j += 32)
                {
                    const int c = scol[j];
                    if constexpr(!GENERIC)
                        acc = mad(sval[j], ldg_ro(x + c), acc);
                    else
                        acc = generic_term(rule, r, c, sval[j], x, acc);
                }
                acc = warp_sum(acc);
                if(lane == 0)
                {
                    if constexpr(GENERIC)
                        if(rule.diag == DIAG_UNIT && r < n_cols)
                            acc = add(acc, ldg_ro(x + r));
                    const T out = axpby_out(alpha, acc, beta, beta_zero != 0, y + r);
                    y[r]        = out;
                    if constexpr(PUSH)
                        push_dst[r - push_row0] = out;
                }
            }
        }
        else if(strat == STRAT_PRODUCT)
        {
            // phase A: every staged entry becomes its product with x, in place
            // (the filtered variant needs the row of an entry to decide, so it multiplies in phase B)
            if constexpr(!GENERIC)
            {
                const int first = ns - a, total = ne - ns;
                for(int i = tid; i < total; i += NT)
                {
                    const int j = first + i;
                    sval[j]     = mul(sval[j], ldg_ro(x + scol[j]));
                }
                __syncthreads();
            }
            // phase B: per-row sums; 32 consecutive rows per warp pass
            for(int rb = d.x + warp * 32; rb < d.y; rb += NT)
            {
                const int  r     = rb + lane;
                const bool valid = r < d.y;
                int        s = 0, e = 0;
                if(valid)
                {
                    s = rp[r] - a;
                    e = rp[r + 1] - a;
                }
                T acc = vt<T>::zero();
                if(e - s <= SHORT_ROW)
                {
                    for(int j = s; j < e; ++j)
                    {
                        if constexpr(GENERIC)
                            acc = generic_term(rule, r, scol[j], sval[j], x, acc);
                        else
                            acc = add(acc, sval[j]);
                    }
                }
                unsigned pending = __ballot_sync(0xffffffffu, valid && (e - s > SHORT_ROW));
                while(pending)
                {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    const int ss = __shfl_sync(0xffffffffu, s, src);
                    const int ee = __shfl_sync(0xffffffffu, e, src);
                    const int rr = rb + src;
                    T         part = vt<T>::zero();
                    for(int j = ss + lane; j < ee; j += 32)
                    {
                        if constexpr(GENERIC)
                            part = generic_term(rule, rr, scol[j], sval[j], x, part);
                        else
                            part = add(part, sval[j]);
                    }
                    part = warp_sum(part);
                    if(lane == src)
                        acc = part;
                }
                if(valid)
                {
                    if constexpr(GENERIC)
                        if(rule.diag == DIAG_UNIT && r < n_cols)
                            acc = add(acc, ldg_ro(x + r));
                    const T out = axpby_out(alpha, acc, beta, beta_zero != 0, y + r);
                    y[r]        = out;
                    if constexpr(PUSH)
                        push_dst[r - push_row0] = out;
                }
            }
        }
        else // STRAT_LONG: one segment of a row split across CTAs -> one partial sum
        {
            const int first = ns - a, total = ne - ns;
            const int r     = d.x;
            T         acc   = vt<T>::zero();
            for(int i = tid; i < total; i += NT)
            {
                const int j = first + i;
                const int c = scol[j];
                if constexpr(!GENERIC)
                    acc = mad(sval[j], ldg_ro(x + c), acc);
                else
                    acc = generic_term(rule, r, c, sval[j], x, acc);
            }
            acc = warp_sum(acc);
            __shared__ T s_part[NT / 32];
            if(lane == 0)
                s_part[warp] = acc;
            __syncthreads();
            if(tid == 0)
            {
                T tot = s_part[0];
#pragma unroll
                for(int w = 1; w < NT / 32; ++w)
                    tot = add(tot, s_part[w]);
                partials[k >> 4] = tot;
            }
        }
    }

    // one warp per long row: partial sums added in segment order, then alpha / beta
    template <typename T>
    __global__ void finish_long_rows_kernel(int n_long,
                                            const int4 *__restrict__ long_rows,
                                            const T *__restrict__ partials,
                                            const T *__restrict__ x,
                                            T *__restrict__ y,
                                            T   alpha,
                                            T   beta,
                                            int beta_zero,
                                            int unit_diag,
                                            int n_cols,
                                            int row_lo,
                                            int row_hi,
                                            T  *push_dst  = nullptr,
                                            int push_row0 = 0)
    {
        const int lane = threadIdx.x & 31;
        const int w    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if(w >= n_long)
            return;
        const int4 lr = long_rows[w];
        if(lr.x < row_lo || lr.x >= row_hi)
            return;
        T acc = vt<T>::zero();
        // fixed association: lane-strided running sums, then the xor tree
        for(int s = lane; s < lr.z; s += 32)
            acc = add(acc, partials[lr.y + s]);
        acc = warp_sum(acc);
        if(lane == 0)
        {
            if(unit_diag && lr.x < n_cols)
                acc = add(acc, x[lr.x]);
            const T out = axpby_out(alpha, acc, beta, beta_zero != 0, y + lr.x);
            y[lr.x]     = out;
            if(push_dst)
                push_dst[lr.x - push_row0] = out;
        }
    }

    // y = beta * y (beta == 0: overwrite with zeros), optionally + alpha * x on the first n_unit entries
    template <typename T>
    __global__ void scale_vector_kernel(long long len, T *__restrict__ y, T beta, int beta_zero, T alpha, const T *__restrict__ x, long long n_unit)
    {
        long long i      = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        long long stride = (long long)gridDim.x * blockDim.x;
        for(; i < len; i += stride)
        {
            T r = beta_zero ? vt<T>::zero() : mul(beta, y[i]);
            if(i < n_unit)
                r = mad(alpha, x[i], r);
            y[i] = r;
        }
    }

    // cross-GPU flags for the row-sharded iteration: a rank publishes "iteration k done" into a word that lives in
    // (or is mapped from) its neighbour's memory; the neighbour's stream spins on it before it reuses the buffers
    static __global__ void signal_flag_kernel(volatile unsigned *flag, unsigned value)
    {
        __threadfence_system(); // everything this stream wrote before (peer stores included) is visible first
        *flag = value;
        __threadfence_system();
    }
    static __global__ void wait_flag_kernel(const volatile unsigned *flag, unsigned value, unsigned *timed_out)
    {
        // counts up (iteration numbers), so ">=" tolerates a neighbour that is already further ahead
        long long spins = 0;
        while((int)(*flag - value) < 0)
        {
            __nanosleep(200);
            if(++spins > 10000000LL) // a few seconds: a lost peer must not hang the GPU
            {
                if(timed_out)
                    *timed_out = 1;
                return;
            }
        }
        __threadfence_system();
    }

    __device__ __forceinline__ void atomic_accumulate(float *p, float v)
    {
        atomicAdd(p, v);
    }
    __device__ __forceinline__ void atomic_accumulate(double *p, double v)
    {
        atomicAdd(p, v);
    }
    __device__ __forceinline__ void atomic_accumulate(float2 *p, float2 v)
    {
        atomicAdd(&p->x, v.x);
        atomicAdd(&p->y, v.y);
    }
    __device__ __forceinline__ void atomic_accumulate(double2 *p, double2 v)
    {
        atomicAdd(&p->x, v.x);
        atomicAdd(&p->y, v.y);
    }

    // y[col] += alpha * op(a_row,col) * x[row] for the entries kept by `rule`
    // (transposed / conjugate-transposed products and the mirrored triangle of symmetric / hermitian
    // matrices).  One CTA per row block; lanes walk the entries, the owning row is found by a binary
    // search over the block's slice of row_ptr.
    template <typename T>
    __global__ void __launch_bounds__(SPMV_THREADS) spmv_scatter_kernel(const int4 *__restrict__ desc,
                                                                       const aoclsparse_int *__restrict__ rp,
                                                                       const aoclsparse_int *__restrict__ col,
                                                                       const T *__restrict__ val,
                                                                       const T *__restrict__ x,
                                                                       T *__restrict__ y,
                                                                       T         alpha,
                                                                       elem_rule rule)
    {
        const int4 d = desc[blockIdx.x];
        for(int p = d.z + threadIdx.x; p < d.w; p += SPMV_THREADS)
        {
            // largest row r in [d.x, d.y) with rp[r] <= p
            int lo = d.x, hi = d.y - 1;
            while(lo < hi)
            {
                const int mid = lo + (hi - lo + 1) / 2;
                if(rp[mid] <= p)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            const int r = lo, c = col[p];
            if(!keep_entry(rule, r, c))
                continue;
            T v = val[p];
            if(c == r ? rule.conj_diag : rule.conj)
                v = cj(v);
            atomic_accumulate(y + c, mul(mul(alpha, ldg_ro(x + r)), v));
        }
    }
}
