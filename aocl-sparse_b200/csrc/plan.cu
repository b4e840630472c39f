// plan.cu -- the GPU analysis pass behind aoclsparse_optimize (and behind the first multiply of a
// matrix that was never optimised).
//
// The reference's aoclsparse_optimize (library/src/analysis/aoclsparse_analysis.cpp:426-566) picks
// one of several CPU storage formats (blocked CSR, "br4", ELL-T hybrid).  None of that carries over.
// What a B200 needs from the analysis is a way to give every CTA the same number of bytes to stream
// no matter how skewed the row lengths are, and to know per CTA how its rows should be reduced.
//
// SPEC (tests/ and oracle/csr_oracle.c restate exactly this, and must agree bit for bit):
//   T = block_nnz, R = block_rows (plan_parameters below), S = 64*T.
//   Segment boundaries: row 0, row m, every forced row cut, and for k = 1.. the smallest row r with
//   row_ptr[r] >= k*S.  Sorted, duplicates removed.
//   Inside a segment [sa, sb), starting at r = sa and until r == sb:
//     len = row_ptr[r+1]-row_ptr[r]
//     if len > T: the row is LONG; emit ceil(len/T) blocks, block q covering entries
//                 [row_ptr[r]+q*T, min(row_ptr[r]+(q+1)*T, row_ptr[r+1])); r += 1
//     else      : r1 = largest row in (r, min(sb, r+R)] with row_ptr[r1]-row_ptr[r] <= T;
//                 emit block rows [r, r1); r = r1
//   Strategy of a non-LONG block with nr rows, nz entries and longest row L:
//     forced strategy if one was given (kid hint), else
//     THREAD  if L <= 64 and L*nr <= 2*nz + nr        (rows short and of similar length)
//     WARP    if nz >= 48*nr and L*nr <= 4*nz         (rows long and of similar length)
//     PRODUCT otherwise
//   LONG segments get consecutive partial-sum slots in block order; long rows are listed in row order.
#include "common.hpp"

#include <algorithm>
#include <cstdlib>

namespace b200
{
    long long ctas_per_wave(size_t elem_size, aoclsparse_int T, int coded);

    void plan_parameters(size_t          elem_size,
                         aoclsparse_int  m,
                         aoclsparse_int  nnz,
                         aoclsparse_int  max_row_nnz,
                         aoclsparse_int &block_nnz,
                         aoclsparse_int &block_rows,
                         int             coded)
    {
        if(coded == 2)
        {
            // Entry codes: 1 staged byte per entry, so ~20 KB hold 20 224 entries and a block is bounded by its ROWS; T only
            // caps blocks of long rows.  Measured on the 7-point 512^3 matrix (profiles/r02_entry_codes.txt): 1.17 ms with
            // 512 rows per block, 0.95 ms with 2048 -- a CTA's prologue (descriptor, bulk copy, table, barrier) is latency
            // that only larger blocks amortise -- while matrices of a few waves want the wave-fitted ~400-500 rows
            // (build_plan).  So: as many rows as still leave 8 waves of CTAs, within [512, 2048]; matrices of less than one
            // wave: one block per resident CTA (at least 64 rows).
            block_nnz             = (aoclsparse_int)((24576 - 4096 - 32) / 256 * 256);
            const long long slots = ctas_per_wave(elem_size, block_nnz, 2);
            long long       r     = (long long)m / (8 * slots) / 64 * 64;
            r                     = r < 512 ? 512 : (r > 2048 ? 2048 : r);
            if((long long)m < slots * 512)
            {
                r = (((long long)m + slots - 1) / slots + 31) / 32 * 32;
                r = r < 64 ? 64 : r;
            }
            block_rows = (aoclsparse_int)r;
            if(const char *e = getenv("AOCLSPARSE_B200_BLOCK_ROWS")) // tuning knob for experiments
            {
                const long v = atol(e);
                if(v >= 32 && v <= 4096)
                    block_rows = (aoclsparse_int)v;
            }
            return;
        }
        // staged bytes per entry = elem_size + 4 (column index); ~24 KB per CTA keeps 8 CTAs resident per SM,
        // which measured best on the 27-point stencil (profiles/r01_sweep_c2.txt): 2048 entries for 8-byte
        // values, 3072 for 4-byte, 1024 for 16-byte
        aoclsparse_int T = (aoclsparse_int)((24576 / (elem_size + 4)) / 512 * 512);
        // 16-byte values (double complex): 1536 entries (30 KB) measured best on the 27-point stencil, together with
        // 128-thread CTAs (build_plan): 4 109 -> 5 779 GB/s (tools/z_sweep.py, profiles/r01_summary.md)
        // with the diagonal-code copy an entry stages elem_size + 1 bytes: the same ~24 KB per CTA hold more entries,
        // which keeps as many bytes in flight per SM as the 32-bit column stream did (2048 entries of 9 bytes left
        // HBM under-used: 8 TB/s-class bandwidth needs ~180 KB in flight per SM; profiles/r02_dcc_sweep.txt)
        if(coded)
            T = (aoclsparse_int)((24576 / (elem_size + 1) - 32) / 256 * 256);
        if(elem_size >= 16)
            T = 1536;
        // skewed row lengths (longest row > 16x the mean): x[col] is a random gather and the multiply is bound by
        // L1 misses (profiles/r01_microbench_gather.txt: 0.9 sectors/clk/SM on a miss, 2.7 on a hit), so leave
        // more of the 228 KB to L1: ~16 KB staged per CTA (profiles/r01_sweep_c3.txt: 1.09 ms vs 2.03 ms on R-MAT)
        const long long mean = m > 0 ? (long long)nnz / m : 0;
        // Measured on R-MAT scale 24 (profiles/r02_c3_l1_lines.txt): every gather that misses L1 holds an L1 line until its
        // sector arrives, so the gathers an SM keeps in flight -- hence its gather rate -- scale with the L1 that the
        // shared-memory carve-out leaves.  8 CTAs of <= 7 KB (+1 KB the system reserves each) fit the 64 KB carve-out
        // and leave 192 KB of L1: 768 entries for 4-byte values, 512 for 8-byte ones (1.011 ms against 1.092 ms with
        // 16 KB per CTA and 2.03 ms with 24 KB).
        if((long long)max_row_nnz > 16 * (mean > 1 ? mean : 1))
            T = (aoclsparse_int)(((7152 / (elem_size + 4)) - 8) / 128 * 128);
        if(T < 512)
            T = 512;
        // small matrices: keep at least ~8 CTAs per SM in the grid
        while(T >= 1024 && (long long)nnz < (long long)T * 148 * 8)
            T -= 512;
        // tuning knob for experiments (never set in tests / bench defaults)
        if(const char *e = getenv("AOCLSPARSE_B200_BLOCK_NNZ"))
        {
            const long v = atol(e);
            if(v >= 256 && v <= 16384)
                T = (aoclsparse_int)(v & ~3L);
        }
        block_nnz  = T;
        block_rows = 1024;
    }

    namespace
    {
        __global__ void grid_rows_kernel(aoclsparse_int m,
                                         const aoclsparse_int *__restrict__ rp,
                                         long long       S,
                                         int             ngrid,
                                         aoclsparse_int *out)
        {
            int k = blockIdx.x * blockDim.x + threadIdx.x;
            if(k >= ngrid)
                return;
            long long target = (long long)(k + 1) * S;
            // smallest r in [0, m] with rp[r] >= target
            aoclsparse_int lo = 0, hi = m;
            while(lo < hi)
            {
                aoclsparse_int mid = lo + (hi - lo) / 2;
                if((long long)rp[mid] >= target)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            out[k] = lo;
        }

        // FILL == false: count blocks / long rows / long segments of each segment
        // FILL == true : write the block descriptors at the scanned offsets
        template <bool FILL>
        __global__ void walk_segments_kernel(int nseg,
                                             const aoclsparse_int *__restrict__ seg_start,
                                             const aoclsparse_int *__restrict__ rp,
                                             aoclsparse_int T,
                                             aoclsparse_int R,
                                             int3          *counts,    // per segment (FILL=false: out)
                                             const int3    *offsets,   // per segment (FILL=true: in)
                                             int4          *desc,
                                             int           *kind,
                                             int4          *long_rows)
        {
            int sidx = blockIdx.x * blockDim.x + threadIdx.x;
            if(sidx >= nseg)
                return;
            const aoclsparse_int sa = seg_start[sidx], sb = seg_start[sidx + 1];
            int nb = 0, nlr = 0, nls = 0;
            int ob = 0, olr = 0, ols = 0;
            if(FILL)
            {
                ob  = offsets[sidx].x;
                olr = offsets[sidx].y;
                ols = offsets[sidx].z;
            }
            aoclsparse_int r = sa;
            while(r < sb)
            {
                const aoclsparse_int p0  = rp[r];
                const aoclsparse_int len = rp[r + 1] - p0;
                if(len > T)
                {
                    const int q = (int)(((long long)len + T - 1) / T);
                    if(FILL)
                    {
                        long_rows[olr + nlr] = make_int4(r, ols + nls, q, 0);
                        for(int s = 0; s < q; ++s)
                        {
                            long long a = (long long)p0 + (long long)s * T;
                            long long b = a + T;
                            if(b > (long long)p0 + len)
                                b = (long long)p0 + len;
                            desc[ob + nb + s] = make_int4(r, r + 1, (int)a, (int)b);
                            kind[ob + nb + s] = STRAT_LONG | ((ols + nls + s) << 4);
                        }
                    }
                    nb += q;
                    nls += q;
                    nlr += 1;
                    r += 1;
                }
                else
                {
                    aoclsparse_int hi = (sb - r > R) ? r + R : sb;
                    // largest r1 in (r, hi] with rp[r1] - p0 <= T ; rp[r+1]-p0 = len <= T holds
                    aoclsparse_int lo = r + 1;
                    const long long lim = (long long)p0 + T;
                    while(lo < hi)
                    {
                        aoclsparse_int mid = lo + (hi - lo + 1) / 2;
                        if((long long)rp[mid] <= lim)
                            lo = mid;
                        else
                            hi = mid - 1;
                    }
                    if(FILL)
                    {
                        desc[ob + nb] = make_int4(r, lo, p0, rp[lo]);
                        kind[ob + nb] = -1; // classified afterwards
                    }
                    nb += 1;
                    r = lo;
                }
            }
            if(!FILL)
                counts[sidx] = make_int3(nb, nlr, nls);
        }

        // one warp per block: longest row -> strategy
        __global__ void classify_kernel(int nblocks,
                                        const aoclsparse_int *__restrict__ rp,
                                        const int4 *__restrict__ desc,
                                        int *kind,
                                        int  forced,
                                        int *strat_count)
        {
            const int lane = threadIdx.x & 31;
            const int b    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
            if(b >= nblocks)
                return;
            int k = kind[b];
            if(k < 0)
            {
                const int4 d  = desc[b];
                int        L  = 0;
                for(int r = d.x + lane; r < d.y; r += 32)
                    L = max(L, rp[r + 1] - rp[r]);
                for(int off = 16; off > 0; off >>= 1)
                    L = max(L, __shfl_xor_sync(0xffffffffu, L, off));
                const long long nr = d.y - d.x, nz = d.w - d.z;
                if(forced >= 0)
                    k = forced;
                else if(L <= 64 && (long long)L * nr <= 2 * nz + nr)
                    k = STRAT_THREAD;
                else if(nz >= 48 * nr && (long long)L * nr <= 4 * nz)
                    k = STRAT_WARP;
                else
                    k = STRAT_PRODUCT;
                if(lane == 0)
                {
                    kind[b] = k;
                    atomicMax(&strat_count[8], d.y - d.x);
                    atomicMax(&strat_count[9], d.w - d.z);
                }
            }
            if(lane == 0)
                atomicAdd(&strat_count[k & 15], 1);
        }
    }

    // CTAs of the multiply kernel that are resident at once on the whole chip for block size T
    long long ctas_per_wave(size_t elem_size, aoclsparse_int T, int coded)
    {
        const long long smem = coded == 2 ? 16 + (long long)((T + 32 + 15) & ~15) + 256 * (elem_size >= 8 ? 16LL : 8LL) + 1024
                               : coded    ? 16 + (long long)(T + 32) * (long long)(elem_size + 1) + 1024 + 1024
                                          : 16 + (long long)(T + 8) * (long long)(elem_size + 4) + 1024; // + 1 KB reserved per CTA
        long long       c    = 232448 / smem;
        if(c > 8)
            c = 8; // 2048 threads per SM / 256
        if(c < 1)
            c = 1;
        return 148 * c;
    }
    bool wave_search_applies(size_t elem_size, aoclsparse_int nnz, aoclsparse_int T, int coded)
    {
        return (long long)nnz < 8 * ctas_per_wave(elem_size, T, coded) * (long long)T
               && (long long)nnz >= ctas_per_wave(elem_size, T, coded) * (long long)T;
    }
    // value-coded plans are bounded by rows: the same test on rows
    bool row_wave_search_applies(size_t elem_size, aoclsparse_int m, aoclsparse_int T, aoclsparse_int R)
    {
        return R >= 512 && (long long)m < 8 * ctas_per_wave(elem_size, T, 2) * (long long)R && (long long)m >= ctas_per_wave(elem_size, T, 2) * (long long)R;
    }

    namespace
    {
        // segment boundaries for block size T and, per segment, the number of blocks / long rows / long segments
        aoclsparse_status count_pass(const dev_csr                     &A,
                                     aoclsparse_int                     T,
                                     aoclsparse_int                     R,
                                     const std::vector<aoclsparse_int> &row_cuts,
                                     cudaStream_t                       st,
                                     std::vector<aoclsparse_int>       &bounds,
                                     std::vector<int3>                 &counts,
                                     dev_buf                           &d_seg,
                                     dev_buf                           &d_grid,
                                     dev_buf                           &d_counts)
        {
            const long long       S  = 64LL * T;
            const aoclsparse_int *rp = A.row_ptr.as<aoclsparse_int>();
            bounds.clear();
        // ---- segment boundaries: nnz grid (device lower bounds) merged with the forced row cuts
        const int ngrid = (int)(((long long)A.nnz + S - 1) / S) - 1 > 0 ? (int)(((long long)A.nnz + S - 1) / S) - 1 : 0;
        bounds.push_back(0);
        if(ngrid > 0)
        {
            if(d_grid.bytes < sizeof(aoclsparse_int) * (size_t)ngrid)
                B200_TRY(d_grid.alloc(sizeof(aoclsparse_int) * (size_t)ngrid));
            grid_rows_kernel<<<(ngrid + 127) / 128, 128, 0, st>>>(A.m, rp, S, ngrid, d_grid.as<aoclsparse_int>());
            B200_LAUNCHED();
            std::vector<aoclsparse_int> h((size_t)ngrid);
            B200_CUDA(cudaMemcpyAsync(h.data(), d_grid.p, sizeof(aoclsparse_int) * (size_t)ngrid, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            bounds.insert(bounds.end(), h.begin(), h.end());
        }
        for(aoclsparse_int c : row_cuts)
            bounds.push_back(c);
        bounds.push_back(A.m);
        std::sort(bounds.begin(), bounds.end());
        bounds.erase(std::unique(bounds.begin(), bounds.end()), bounds.end());
        const int nseg = (int)bounds.size() - 1;

        if(d_seg.bytes < sizeof(aoclsparse_int) * bounds.size())
            B200_TRY(d_seg.alloc(sizeof(aoclsparse_int) * bounds.size() * 2));
        if(d_counts.bytes < sizeof(int3) * (size_t)nseg)
            B200_TRY(d_counts.alloc(sizeof(int3) * (size_t)nseg * 2));
        B200_CUDA(cudaMemcpyAsync(d_seg.p, bounds.data(), sizeof(aoclsparse_int) * bounds.size(), cudaMemcpyHostToDevice, st));

        walk_segments_kernel<false><<<(nseg + 63) / 64, 64, 0, st>>>(
            nseg, d_seg.as<aoclsparse_int>(), rp, T, R, d_counts.as<int3>(), nullptr, nullptr, nullptr, nullptr);
        B200_LAUNCHED();
        counts.assign((size_t)nseg, make_int3(0, 0, 0));
        B200_CUDA(cudaMemcpyAsync(counts.data(), d_counts.p, sizeof(int3) * (size_t)nseg, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
            return aoclsparse_status_success;
        }
    }

    aoclsparse_status build_plan(dev_csr                           &A,
                                 size_t                             elem_size,
                                 aoclsparse_int                     max_row_nnz,
                                 aoclsparse_int                     forced_strategy,
                                 const std::vector<aoclsparse_int> &row_cuts,
                                 cudaStream_t                       st,
                                 aoclsparse_int                     block_nnz_override,
                                 int                                coded,
                                 row_block_plan                    *target)
    {
        row_block_plan &P = target ? *target : A.plan;
        P                 = row_block_plan();
        plan_parameters(elem_size, A.m, A.nnz, max_row_nnz, P.block_nnz, P.block_rows, coded);
        if(block_nnz_override > 0)
            P.block_nnz = block_nnz_override;
        // few rows per block (long rows or wide values): the thread-per-row strategy keeps only one lane per row busy, so
        // smaller CTAs put more of them -- hence more rows in flight -- on an SM
        {
            const long long mean = A.m > 0 ? (long long)A.nnz / A.m : 0;
            if(mean > 0 && (long long)P.block_nnz / mean <= 64)
                P.threads = 128;
        }
        if(const char *e = getenv("AOCLSPARSE_B200_THREADS"))
        {
            const int v = atoi(e);
            if(v == 128 || v == 256 || v == 512)
                P.threads = v;
        }
        if(const char *e = getenv("AOCLSPARSE_B200_L2HINT"))
            P.stream_hint = atoi(e) ? 1 : 0;
        if(const char *e = getenv("AOCLSPARSE_B200_PDL"))
            P.pdl = atoi(e) ? 1 : 0;
        const aoclsparse_int *rp = A.row_ptr.as<aoclsparse_int>();

        if(A.m == 0)
        {
            P.valid = true;
            return aoclsparse_status_success;
        }

        // ---- wave-aware block size (small matrices only): with fewer than ~8 waves of CTAs the partial last wave is
        // a visible fraction of the run (2D Laplacian 1000^2: 2478 blocks = 2.09 waves), so candidates T0 + 32k are
        // tried and the one minimising ceil(blocks / CTAs per wave) * T is kept (profiles/r01_summary.md)
        std::vector<aoclsparse_int> bounds;
        std::vector<int3>           counts;
        dev_buf                     d_seg, d_grid, d_cnt;
        aoclsparse_int              T = P.block_nnz;
        aoclsparse_int              R = P.block_rows;
        if(coded == 2 && block_nnz_override <= 0 && row_wave_search_applies(elem_size, A.m, T, R) && !getenv("AOCLSPARSE_B200_BLOCK_ROWS")
           && !getenv("AOCLSPARSE_B200_BLOCK_NNZ"))
        {
            // value-coded plan: blocks end at R rows, so the candidates are row counts R0 - 8k (same cost function)
            long long      best_cost = -1;
            aoclsparse_int best_R    = R;
            for(int k = 0; k <= 24; ++k)
            {
                const aoclsparse_int Rk = R - 8 * k;
                B200_TRY(count_pass(A, T, Rk, row_cuts, st, bounds, counts, d_seg, d_grid, d_cnt));
                long long nbk = 0;
                for(const int3 &c : counts)
                    nbk += c.x;
                const long long wave = ctas_per_wave(elem_size, T, 2);
                const long long cost = (((nbk * 203 + 199) / 200 + wave - 1) / wave) * (long long)Rk;
                if(best_cost < 0 || cost < best_cost)
                {
                    best_cost = cost;
                    best_R    = Rk;
                }
            }
            R = P.block_rows = best_R;
        }
        else if(coded != 2 && block_nnz_override <= 0 && wave_search_applies(elem_size, A.nnz, T, coded) && !getenv("AOCLSPARSE_B200_BLOCK_NNZ"))
        {
            long long      best_cost = -1;
            aoclsparse_int best_T    = T;
            for(int k = 0; k <= 16; ++k)
            {
                const aoclsparse_int Tk = T + 32 * k;
                B200_TRY(count_pass(A, Tk, R, row_cuts, st, bounds, counts, d_seg, d_grid, d_cnt));
                long long nbk = 0;
                for(const int3 &c : counts)
                    nbk += c.x;
                const long long wave  = ctas_per_wave(elem_size, Tk, coded);
                // 1.5 % slack: a last wave that is only just full still ends late (measured, profiles/r01_summary.md)
                const long long cost  = (((nbk * 203 + 199) / 200 + wave - 1) / wave) * (long long)Tk;
                if(best_cost < 0 || cost < best_cost)
                {
                    best_cost = cost;
                    best_T    = Tk;
                }
            }
            T = P.block_nnz = best_T;
        }
        B200_TRY(count_pass(A, T, R, row_cuts, st, bounds, counts, d_seg, d_grid, d_cnt));
        const int nseg = (int)bounds.size() - 1;
        dev_buf   d_offsets;
        B200_TRY(d_offsets.alloc(sizeof(int3) * (size_t)nseg));
        std::vector<int3> offsets((size_t)nseg);
        long long nb = 0, nlr = 0, nls = 0;
        for(int s = 0; s < nseg; ++s)
        {
            offsets[s] = make_int3((int)nb, (int)nlr, (int)nls);
            nb += counts[s].x;
            nlr += counts[s].y;
            nls += counts[s].z;
        }
        if(nb > 0x7fffffffLL / 16)
            return aoclsparse_status_internal_error;
        P.n_blocks        = (aoclsparse_int)nb;
        P.n_long_rows     = (aoclsparse_int)nlr;
        P.n_long_segments = (aoclsparse_int)nls;

        // where each row cut begins in block numbering (cuts are segment boundaries)
        P.cut_block.clear();
        for(aoclsparse_int c : row_cuts)
        {
            size_t sidx = std::lower_bound(bounds.begin(), bounds.end(), c) - bounds.begin();
            P.cut_block.push_back(sidx < (size_t)nseg ? offsets[sidx].x : (aoclsparse_int)nb);
        }

        B200_TRY(P.desc.alloc(sizeof(int4) * (size_t)std::max<long long>(nb, 1)));
        B200_TRY(P.kind.alloc(sizeof(int) * (size_t)std::max<long long>(nb, 1)));
        B200_TRY(P.long_rows.alloc(sizeof(int4) * (size_t)std::max<long long>(nlr, 1)));
        B200_TRY(P.partials.alloc(16 * (size_t)std::max<long long>(nls, 1)));
        B200_CUDA(cudaMemcpyAsync(d_offsets.p, offsets.data(), sizeof(int3) * (size_t)nseg, cudaMemcpyHostToDevice, st));
        walk_segments_kernel<true><<<(nseg + 63) / 64, 64, 0, st>>>(nseg,
                                                                   d_seg.as<aoclsparse_int>(),
                                                                   rp,
                                                                   T,
                                                                   R,
                                                                   nullptr,
                                                                   d_offsets.as<int3>(),
                                                                   P.desc.as<int4>(),
                                                                   P.kind.as<int>(),
                                                                   P.long_rows.as<int4>());
        B200_LAUNCHED();

        dev_buf d_sc;
        B200_TRY(d_sc.alloc(sizeof(int) * 16));
        B200_CUDA(cudaMemsetAsync(d_sc.p, 0, sizeof(int) * 16, st));
        if(nb > 0)
        {
            const long long threads = nb * 32;
            classify_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
                (int)nb, rp, P.desc.as<int4>(), P.kind.as<int>(), (int)forced_strategy, d_sc.as<int>());
            B200_LAUNCHED();
        }
        int sc[16];
        B200_CUDA(cudaMemcpyAsync(sc, d_sc.p, sizeof(sc), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        for(int i = 0; i < 4; ++i)
            P.n_strat[i] = sc[i];
        P.max_block_rows = sc[8] > 1 ? sc[8] : 1;
        P.max_block_nnz  = sc[9] > 0 ? sc[9] : 0;

        P.valid = true;
        return aoclsparse_status_success;
    }

    // ------------------------------------------------------------------------------------------------------------
    // Diagonal-code copy.  SPEC (restated by oracle/csr_oracle.c::oracle_diag_codes, compared bit for bit):
    //   D = sorted (ascending) set of the distinct values col_idx[p] - r over all stored entries p of all rows r
    //   if 1 <= |D| <= 256: codes[p] = index of (col_idx[p] - r) in D, code_offsets[i] = D[i] (i >= |D|: D[|D|-1])
    //   else: not applicable (n_codes = 0).
    // ------------------------------------------------------------------------------------------------------------
    namespace
    {
        constexpr int OFFSET_SLOTS = 1024;
        constexpr int OFFSET_EMPTY = (int)0x80000000;

        // ctl[0] = distinct offsets inserted so far, ctl[1] = 1 once there are more than 256 (everyone gives up)
        __global__ void collect_offsets_kernel(aoclsparse_int m,
                                               const aoclsparse_int *__restrict__ rp,
                                               const aoclsparse_int *__restrict__ col,
                                               int *table,
                                               int *ctl)
        {
            const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if(r >= m || *(volatile int *)(ctl + 1))
                return;
            int last = OFFSET_EMPTY; // consecutive rows of a stencil repeat the same offsets: skip the re-probe
            for(aoclsparse_int p = rp[r]; p < rp[r + 1]; ++p)
            {
                const int off = col[p] - (int)r;
                if(off == last)
                    continue;
                last       = off;
                unsigned h = ((unsigned)off * 2654435761u) >> 22; // 10 bits
                for(int probe = 0; probe < OFFSET_SLOTS; ++probe, h = (h + 1) & (OFFSET_SLOTS - 1))
                {
                    int cur = *(volatile int *)(table + h);
                    if(cur == off)
                        break;
                    if(cur == OFFSET_EMPTY)
                    {
                        cur = atomicCAS(table + h, OFFSET_EMPTY, off);
                        if(cur == OFFSET_EMPTY)
                        {
                            if(atomicAdd(ctl, 1) + 1 > CODE_TABLE_MAX)
                                atomicExch(ctl + 1, 1);
                            break;
                        }
                        if(cur == off)
                            break;
                    }
                }
                if(*(volatile int *)(ctl + 1))
                    return;
            }
        }

        __global__ void encode_offsets_kernel(aoclsparse_int m,
                                              const aoclsparse_int *__restrict__ rp,
                                              const aoclsparse_int *__restrict__ col,
                                              const int *__restrict__ sorted_off,
                                              int            n_off,
                                              unsigned char *codes)
        {
            __shared__ int so[CODE_TABLE_MAX];
            for(int i = threadIdx.x; i < CODE_TABLE_MAX; i += blockDim.x)
                so[i] = sorted_off[i];
            __syncthreads();
            const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if(r >= m)
                return;
            for(aoclsparse_int p = rp[r]; p < rp[r + 1]; ++p)
            {
                const int off = col[p] - (int)r;
                int       lo = 0, hi = n_off - 1;
                while(lo < hi)
                {
                    const int mid = (lo + hi) >> 1;
                    if(so[mid] < off)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                codes[p] = (unsigned char)lo;
            }
        }
    }

    // the sorted distinct (col - row) offsets of A, empty when there are none or more than 256 (or the knob is off)
    aoclsparse_status probe_diag_offsets(const dev_csr &A, std::vector<int> &offs, cudaStream_t st)
    {
        offs.clear();
        if(A.nnz <= 0 || A.m <= 0)
            return aoclsparse_status_success;
        if(const char *e = getenv("AOCLSPARSE_B200_DIAG_CODES")) // A/B knob
            if(atoi(e) == 0)
                return aoclsparse_status_success;
        dev_buf work;
        B200_TRY(work.alloc(sizeof(int) * (OFFSET_SLOTS + 2)));
        int *table = work.as<int>(), *ctl = table + OFFSET_SLOTS;
        std::vector<int> h((size_t)OFFSET_SLOTS + 2, OFFSET_EMPTY);
        h[OFFSET_SLOTS] = h[OFFSET_SLOTS + 1] = 0;
        B200_CUDA(cudaMemcpyAsync(table, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, st));
        const unsigned grid = (unsigned)(((long long)A.m + 255) / 256);
        collect_offsets_kernel<<<grid, 256, 0, st>>>(A.m, A.row_ptr.as<aoclsparse_int>(), A.col_idx.as<aoclsparse_int>(), table, ctl);
        B200_LAUNCHED();
        B200_CUDA(cudaMemcpyAsync(h.data(), table, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        if(h[OFFSET_SLOTS + 1] != 0 || h[OFFSET_SLOTS] <= 0 || h[OFFSET_SLOTS] > CODE_TABLE_MAX)
            return aoclsparse_status_success; // too many distinct diagonals: keep the 32-bit column stream
        for(int i = 0; i < OFFSET_SLOTS; ++i)
            if(h[i] != OFFSET_EMPTY)
                offs.push_back(h[i]);
        if((int)offs.size() != h[OFFSET_SLOTS])
        {
            offs.clear();
            return aoclsparse_status_internal_error;
        }
        std::sort(offs.begin(), offs.end());
        return aoclsparse_status_success;
    }

    aoclsparse_status build_diag_codes(dev_csr &A, cudaStream_t st)
    {
        row_block_plan &P = A.plan;
        P.n_codes         = 0;
        P.code_state      = 1;
        P.codes.release();
        P.code_offsets.release();
        if(!P.valid || P.n_blocks <= 0 || A.nnz <= 0 || P.n_strat[STRAT_THREAD] != P.n_blocks)
            return aoclsparse_status_success;
        std::vector<int> offs;
        B200_TRY(probe_diag_offsets(A, offs, st));
        if(offs.empty())
            return aoclsparse_status_success;
        const unsigned grid  = (unsigned)(((long long)A.m + 255) / 256);
        const int      n_off = (int)offs.size();
        offs.resize(CODE_TABLE_MAX, offs.back());
        B200_TRY(P.code_offsets.alloc(sizeof(int) * CODE_TABLE_MAX));
        B200_TRY(P.codes.alloc((size_t)A.nnz));
        B200_CUDA(cudaMemcpyAsync(P.code_offsets.p, offs.data(), sizeof(int) * CODE_TABLE_MAX, cudaMemcpyHostToDevice, st));
        encode_offsets_kernel<<<grid, 256, 0, st>>>(A.m,
                                                   A.row_ptr.as<aoclsparse_int>(),
                                                   A.col_idx.as<aoclsparse_int>(),
                                                   P.code_offsets.as<int>(),
                                                   n_off,
                                                   P.codes.as<unsigned char>());
        B200_LAUNCHED();
        B200_CUDA(cudaStreamSynchronize(st)); // offs (host) is read by the copy above
        P.n_codes = n_off;
        return aoclsparse_status_success;
    }

    // ------------------------------------------------------------------------------------------------------------
    // Entry-code copy.  SPEC (restated by oracle/csr_oracle.c::oracle_entry_codes, compared bit for bit):
    //   needs the diagonal-code copy (offset table D, at most 256 distinct col - row offsets) and a 4- or 8-byte value type
    //   V = sorted (ascending, as unsigned integers) set of the distinct BIT PATTERNS of val[p] over all stored entries
    //       (-0.0 and 0.0, and different NaNs, are different patterns); not applicable if |V| > 256 or a value has the
    //       all-ones pattern
    //   Q = sorted set of the distinct pairs (index of col[p] - r in D, index of val[p] in V), ascending by the first,
    //       then by the second component; not applicable if |Q| > 256
    //   ecodes[p] = index of p's pair in Q;  etab_off[i] = D[Q[i].first], etab_val[i] = V[Q[i].second]  (i < |Q|)
    //   else: not applicable (n_ecodes = 0).
    // ------------------------------------------------------------------------------------------------------------
    namespace
    {
        constexpr int                VALUE_SLOTS = 1024;
        constexpr unsigned long long VALUE_EMPTY = 0xffffffffffffffffull; // a value with this pattern: not applicable

        template <typename U> // unsigned / unsigned long long: the value's bit pattern
        __global__ void collect_values_kernel(long long nnz, const U *__restrict__ val, unsigned long long *table, int *ctl)
        {
            long long       i      = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            const long long stride = (long long)gridDim.x * blockDim.x;
            unsigned long long last0 = VALUE_EMPTY, last1 = VALUE_EMPTY; // the patterns seen last: skip the re-probe
            for(; i < nnz; i += stride)
            {
                const unsigned long long v = (unsigned long long)val[i];
                if(v == last0 || v == last1)
                    continue;
                if(*(volatile int *)(ctl + 1))
                    return;
                if(v == VALUE_EMPTY)
                {
                    atomicExch(ctl + 1, 1);
                    return;
                }
                last1      = last0;
                last0      = v;
                unsigned h = (unsigned)((v * 0x9e3779b97f4a7c15ull) >> 54); // 10 bits
                for(int probe = 0; probe < VALUE_SLOTS; ++probe, h = (h + 1) & (VALUE_SLOTS - 1))
                {
                    unsigned long long cur = *(volatile unsigned long long *)(table + h);
                    if(cur == v)
                        break;
                    if(cur == VALUE_EMPTY)
                    {
                        cur = atomicCAS(table + h, VALUE_EMPTY, v);
                        if(cur == VALUE_EMPTY)
                        {
                            if(atomicAdd(ctl, 1) + 1 > CODE_TABLE_MAX)
                                atomicExch(ctl + 1, 1);
                            break;
                        }
                        if(cur == v)
                            break;
                    }
                }
            }
        }

        __device__ __forceinline__ int value_index(const unsigned long long *sv, int n_vals, unsigned long long v)
        {
            int lo = 0, hi = n_vals - 1;
            while(lo < hi)
            {
                const int mid = (lo + hi) >> 1;
                if(sv[mid] < v)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            return lo;
        }

        // pass 1 (rank == nullptr): seen[diag code << 8 | value index] = 1 for every stored entry
        // pass 2: ecodes[p] = rank[diag code << 8 | value index]
        template <typename U>
        __global__ void entry_pairs_kernel(long long nnz,
                                           const U *__restrict__ val,
                                           const unsigned char *__restrict__ dcodes,
                                           const unsigned long long *__restrict__ sorted_vals,
                                           int                  n_vals,
                                           unsigned char       *seen,
                                           const unsigned char *__restrict__ rank,
                                           unsigned char       *ecodes)
        {
            __shared__ unsigned long long sv[CODE_TABLE_MAX];
            for(int i = threadIdx.x; i < CODE_TABLE_MAX; i += blockDim.x)
                sv[i] = sorted_vals[i];
            __syncthreads();
            long long       i      = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            const long long stride = (long long)gridDim.x * blockDim.x;
            for(; i < nnz; i += stride)
            {
                const int pair = ((int)dcodes[i] << 8) | value_index(sv, n_vals, (unsigned long long)val[i]);
                if(rank)
                    ecodes[i] = rank[pair];
                else if(!seen[pair])
                    seen[pair] = 1;
            }
        }

        inline unsigned value_grid(long long nnz)
        {
            long long b = (nnz + 255) / 256;
            if(b > 148LL * 16)
                b = 148LL * 16;
            return (unsigned)(b < 1 ? 1 : b);
        }
    }

    aoclsparse_status probe_values(const dev_csr &A, size_t elem_size, std::vector<unsigned long long> &vals, cudaStream_t st)
    {
        vals.clear();
        if(A.nnz <= 0 || (elem_size != 4 && elem_size != 8))
            return aoclsparse_status_success;
        if(const char *e = getenv("AOCLSPARSE_B200_ENTRY_CODES")) // A/B knob
            if(atoi(e) == 0)
                return aoclsparse_status_success;
        dev_buf work;
        B200_TRY(work.alloc(sizeof(unsigned long long) * (VALUE_SLOTS + 1)));
        unsigned long long *table = work.as<unsigned long long>();
        int                *ctl   = reinterpret_cast<int *>(table + VALUE_SLOTS);
        std::vector<unsigned long long> h((size_t)VALUE_SLOTS + 1, VALUE_EMPTY);
        h[VALUE_SLOTS] = 0; // ctl[0] = distinct patterns so far, ctl[1] = give up
        B200_CUDA(cudaMemcpyAsync(table, h.data(), sizeof(unsigned long long) * h.size(), cudaMemcpyHostToDevice, st));
        if(elem_size == 8)
            collect_values_kernel<unsigned long long><<<value_grid(A.nnz), 256, 0, st>>>(A.nnz, A.val.as<unsigned long long>(), table, ctl);
        else
            collect_values_kernel<unsigned><<<value_grid(A.nnz), 256, 0, st>>>(A.nnz, A.val.as<unsigned>(), table, ctl);
        B200_LAUNCHED();
        B200_CUDA(cudaMemcpyAsync(h.data(), table, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        const int n_seen = (int)(h[VALUE_SLOTS] & 0xffffffffull), gave_up = (int)(h[VALUE_SLOTS] >> 32);
        if(gave_up != 0 || n_seen <= 0 || n_seen > CODE_TABLE_MAX)
            return aoclsparse_status_success; // too many distinct values: keep the value stream
        for(int i = 0; i < VALUE_SLOTS; ++i)
            if(h[i] != VALUE_EMPTY)
                vals.push_back(h[i]);
        if((int)vals.size() != n_seen)
        {
            vals.clear();
            return aoclsparse_status_internal_error;
        }
        std::sort(vals.begin(), vals.end());
        return aoclsparse_status_success;
    }

    aoclsparse_status build_entry_codes(dev_csr &A, size_t elem_size, aoclsparse_int max_row_nnz, const std::vector<aoclsparse_int> &row_cuts, cudaStream_t st)
    {
        row_block_plan &P = A.plan;
        P.n_ecodes        = 0;
        P.ecodes_stale    = false;
        P.ecodes.release();
        P.etab_off.release();
        P.etab_val.release();
        if(!P.valid || P.n_codes <= 0 || A.nnz <= 0)
        {
            P.eplan.reset();
            return aoclsparse_status_success;
        }
        // the block plan of the entry-coded kernels depends on the pattern only: kept across re-encodings
        if(!P.eplan)
        {
            std::unique_ptr<row_block_plan> E(new(std::nothrow) row_block_plan);
            if(!E)
                return aoclsparse_status_memory_error;
            B200_TRY(build_plan(A, elem_size, max_row_nnz, -1, row_cuts, st, 0, 2, E.get()));
            if(E->n_blocks <= 0 || E->n_strat[STRAT_THREAD] != E->n_blocks)
                return aoclsparse_status_success; // some block is not thread-per-row at that block size
            P.eplan = std::move(E);
        }
        std::vector<unsigned long long> vals;
        B200_TRY(probe_values(A, elem_size, vals, st));
        if(vals.empty())
            return aoclsparse_status_success;
        const int n_vals = (int)vals.size();
        vals.resize(CODE_TABLE_MAX, vals.back());
        dev_buf d_vals, d_seen, d_rank;
        B200_TRY(d_vals.alloc(8 * CODE_TABLE_MAX));
        B200_TRY(d_seen.alloc(65536));
        B200_TRY(d_rank.alloc(65536));
        B200_CUDA(cudaMemcpyAsync(d_vals.p, vals.data(), 8 * CODE_TABLE_MAX, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemsetAsync(d_seen.p, 0, 65536, st));
        auto pass = [&](const unsigned char *rank, unsigned char *out) -> aoclsparse_status {
            if(elem_size == 8)
                entry_pairs_kernel<unsigned long long><<<value_grid(A.nnz), 256, 0, st>>>(A.nnz, A.val.as<unsigned long long>(), P.codes.as<unsigned char>(),
                                                                                      d_vals.as<unsigned long long>(), n_vals, d_seen.as<unsigned char>(), rank, out);
            else
                entry_pairs_kernel<unsigned><<<value_grid(A.nnz), 256, 0, st>>>(A.nnz, A.val.as<unsigned>(), P.codes.as<unsigned char>(),
                                                                            d_vals.as<unsigned long long>(), n_vals, d_seen.as<unsigned char>(), rank, out);
            B200_LAUNCHED();
            return aoclsparse_status_success;
        };
        B200_TRY(pass(nullptr, nullptr));
        std::vector<unsigned char> seen(65536), rank(65536, 0);
        std::vector<int>           offs(CODE_TABLE_MAX);
        B200_CUDA(cudaMemcpyAsync(seen.data(), d_seen.p, 65536, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(offs.data(), P.code_offsets.p, sizeof(int) * CODE_TABLE_MAX, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        std::vector<int> pairs; // ascending pair ids = ascending (offset, value pattern)
        for(int q = 0; q < 65536; ++q)
            if(seen[q])
                pairs.push_back(q);
        if(pairs.empty() || pairs.size() > (size_t)CODE_TABLE_MAX)
            return aoclsparse_status_success; // more than 256 distinct (offset, value) pairs
        std::vector<int>                t_off(CODE_TABLE_MAX);
        std::vector<unsigned long long> t_val(CODE_TABLE_MAX);
        for(size_t i = 0; i < (size_t)CODE_TABLE_MAX; ++i)
        {
            const int q = pairs[i < pairs.size() ? i : pairs.size() - 1];
            t_off[i]    = offs[q >> 8];
            t_val[i]    = vals[q & 255];
            if(i < pairs.size())
                rank[q] = (unsigned char)i;
        }
        B200_TRY(P.etab_off.alloc(sizeof(int) * CODE_TABLE_MAX));
        B200_TRY(P.etab_val.alloc(elem_size * CODE_TABLE_MAX));
        B200_TRY(P.ecodes.alloc((size_t)A.nnz));
        std::vector<unsigned> t_val32(CODE_TABLE_MAX);
        for(int i = 0; i < CODE_TABLE_MAX; ++i)
            t_val32[i] = (unsigned)t_val[i];
        B200_CUDA(cudaMemcpyAsync(d_rank.p, rank.data(), 65536, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(P.etab_off.p, t_off.data(), sizeof(int) * CODE_TABLE_MAX, cudaMemcpyHostToDevice, st));
        if(elem_size == 8)
            B200_CUDA(cudaMemcpyAsync(P.etab_val.p, t_val.data(), 8 * CODE_TABLE_MAX, cudaMemcpyHostToDevice, st));
        else
            B200_CUDA(cudaMemcpyAsync(P.etab_val.p, t_val32.data(), 4 * CODE_TABLE_MAX, cudaMemcpyHostToDevice, st));
        B200_TRY(pass(d_rank.as<unsigned char>(), P.ecodes.as<unsigned char>()));
        B200_CUDA(cudaStreamSynchronize(st)); // host staging vectors are read by the copies above
        P.n_ecodes = (aoclsparse_int)pairs.size();
        return aoclsparse_status_success;
    }

    aoclsparse_status build_plan_with_codes(dev_csr                           &A,
                                            size_t                             elem_size,
                                            aoclsparse_int                     max_row_nnz,
                                            aoclsparse_int                     forced_strategy,
                                            const std::vector<aoclsparse_int> &row_cuts,
                                            cudaStream_t                       st)
    {
        std::vector<int>                offs;
        std::vector<unsigned long long> vals;
        if(forced_strategy < 0 || forced_strategy == STRAT_THREAD)
            B200_TRY(probe_diag_offsets(A, offs, st));
        if(!offs.empty())
            B200_TRY(probe_values(A, elem_size, vals, st));
        B200_TRY(build_plan(A, elem_size, max_row_nnz, forced_strategy, row_cuts, st, 0, offs.empty() ? 0 : 1));
        if(offs.empty())
        {
            A.plan.code_state = 1;
            return aoclsparse_status_success;
        }
        B200_TRY(build_diag_codes(A, st));
        if(A.plan.n_codes == 0) // some block is not thread-per-row: the plain plan with its own block size
        {
            B200_TRY(build_plan(A, elem_size, max_row_nnz, forced_strategy, row_cuts, st));
            A.plan.code_state = 1;
            return aoclsparse_status_success;
        }
        if(!vals.empty())
            B200_TRY(build_entry_codes(A, elem_size, max_row_nnz, row_cuts, st));
        return aoclsparse_status_success;
    }
}
