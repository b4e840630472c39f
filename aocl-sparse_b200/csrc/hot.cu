// hot.cu -- SpMV for matrices whose columns are hit very unevenly (power-law graphs, BASELINE config 3): the analysis
// that finds the hot columns, and the persistent kernel that keeps their x entries in shared memory.
//
// Replaces, for that class of matrices, the same reference kernels as spmv_kernels.cuh
// (aoclsparse_csrmv_vectorized<float>, library/src/level2/aoclsparse_csrmv_kr.hpp:734-831: an OpenMP static row split whose
// threads gather x through the CPU cache hierarchy).
//
// Why.  On R-MAT scale 24 the row-block kernel is not HBM-bound: every x[col] is a 4-byte gather that misses L1 and
// moves a 32-byte sector out of L2, and an SM retires only ~0.9 such gathers per clock (profiles/r01_microbench_gather.txt);
// 263 M gathers are ~1.0 ms whatever else the kernel does (profiles/r01_ncu_c3.txt, r02_c3_l1_lines.txt).  Sorting
// entries does not make the gathers share sectors (tools/experiments/c3_sector_model.py), and the hardware L1 cannot
// keep the popular columns either (128-byte lines thrashed by the cold gathers).  But column popularity is very uneven
// -- the 16 K most frequent columns carry ~28 % of all stored entries -- and a table in shared memory holds exactly
// those, at element granularity and without eviction: a gather that finds its column in the table never enters the
// L1 miss path.
//
// Two earlier attempts at this (tools/experiments/r01/spmv_hot.cuh, r02/hot_pipeline.cu) lost because they staged
// 100-190 KB of matrix slices next to the table: every gather that misses holds an L1 line until its sector arrives, so
// the gather rate of an SM scales with the L1 the carve-out leaves (profiles/r02_c3_l1_lines.txt).  This version is
// sized the other way round: the staging is what the row-block kernel uses on such matrices (8 x ~6 KB), the table
// takes a further 32-64 KB, and >= 124 KB stay L1.
//
// aoclsparse_optimize (general non-transposed mv hint, skewed row lengths, memory not restricted) or
// aoclsparse_b200_set_hot_table:
//   1. column histogram (one atomic per stored entry), radix sort of the counts (CUB, analysis time only);
//   2. the K most frequent columns become the table (K from the shared memory the kernel can spare);
//   3. a second column array in which those columns are replaced by HOT_BIT | slot (the stored matrix is untouched).
// Kernel: ONE CTA of 1024 threads per SM, resident for the whole launch; the table is filled once per launch from x.
// The CTA is 8 (4) independent TEAMS of 128 (256) threads; a team is what a CTA is in spmv_row_blocks_kernel: it owns a
// staging buffer + mbarrier, walks the plan's row blocks with the stride of all teams, and per block does bulk copy ->
// wait -> reduce by the block's strategy, synchronising with a named barrier of its own.  Sums are formed in the same
// order as in the row-block kernel, from the same products: the results are bit-identical to it.
#include "spmv_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

namespace b200
{
    namespace
    {
        constexpr int HT_THREADS = 1024;
        constexpr int HT_HEADER  = 1024; // 2 x 16 mbarriers (256 B) | partial sums of split-row segments at +256 (2 x 32 x 8 B)
        constexpr int HOT_BIT    = (int)0x80000000;
        constexpr int HT_U       = 8; // entries of a block per thread of its team (all gathered at once)

        inline unsigned grid_for(long long n, int tpb)
        {
            long long b = (n + tpb - 1) / tpb;
            if(b > 148LL * 32)
                b = 148LL * 32;
            return (unsigned)(b < 1 ? 1 : b);
        }
        __global__ void col_hist_kernel(long long nnz, const int *__restrict__ col, unsigned *cnt)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < nnz; i += (long long)gridDim.x * blockDim.x)
                atomicAdd(cnt + col[i], 1u);
        }
        __global__ void hot_iota_kernel(long long n, int *out)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < n; i += (long long)gridDim.x * blockDim.x)
                out[i] = (int)i;
        }
        __global__ void slot_scatter_kernel(int k, const int *__restrict__ hot_cols, int *slot_of)
        {
            int i = blockIdx.x * blockDim.x + threadIdx.x;
            if(i < k)
                slot_of[hot_cols[i]] = i;
        }
        __global__ void remap_kernel(long long nnz, const int *__restrict__ col, const int *__restrict__ slot_of, int *col_hot)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < nnz; i += (long long)gridDim.x * blockDim.x)
            {
                const int c = col[i], s = slot_of[c];
                col_hot[i]  = s >= 0 ? (HOT_BIT | s) : c;
            }
        }

        template <int TEAM>
        __device__ __forceinline__ void team_sync(int team)
        {
            // named barriers 1..15 (0 is __syncthreads); 16 teams of 64 threads also use 0, after the kernel's only __syncthreads
            asm volatile("bar.sync %0, %1;" ::"r"(TEAM == 64 ? team : team + 1), "n"(TEAM) : "memory");
        }
        // generic-proxy accesses to the staging buffer (the in-place products are writes) are ordered before the bulk
        // copy (async proxy) that refills it
        __device__ __forceinline__ void fence_proxy_async_smem()
        {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }

        inline size_t ht_stage_bytes(size_t elem, int cap)
        {
            return (size_t)cap * (elem + 4);
        }
        // two stages per team: the next block's slice arrives while the current one is reduced
        inline size_t ht_smem_bytes(size_t elem, int entries, int teams, int cap)
        {
            return HT_HEADER + ((((size_t)entries * elem) + 15) & ~(size_t)15) + (size_t)teams * 2 * ht_stage_bytes(elem, cap);
        }

        template <typename T, int TEAM>
        __global__ void __launch_bounds__(HT_THREADS, 1) spmv_hot_teams_kernel(const int4 *__restrict__ desc,
                                                                              const int *__restrict__ kind,
                                                                              int n_blocks,
                                                                              int cap, // staged capacity in entries
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
            constexpr int NTEAMS = HT_THREADS / TEAM;
            constexpr int TW     = TEAM / 32;
            constexpr int U      = HT_U; // gathers a thread keeps in flight = entries per thread of one block
            extern __shared__ __align__(16) unsigned char smem_raw[];
            uint64_t      *bars        = reinterpret_cast<uint64_t *>(smem_raw);                // [team][stage]
            T             *table       = reinterpret_cast<T *>(smem_raw + HT_HEADER);
            const size_t   table_bytes = (((size_t)table_entries * sizeof(T)) + 15) & ~(size_t)15;
            const int      tid = threadIdx.x, team = tid / TEAM, t = tid % TEAM, lane = t & 31, warp = t >> 5;
            const size_t   stage_bytes = (size_t)cap * (sizeof(T) + 4);
            unsigned char *stage0      = smem_raw + HT_HEADER + table_bytes + (size_t)team * 2 * stage_bytes;
            uint64_t      *bar         = bars + 2 * team;
            T             *s_part      = reinterpret_cast<T *>(smem_raw + 256) + team * 2 * TW; // [parity][warp]

            const int total_teams = gridDim.x * NTEAMS;
            const int b0          = blockIdx.x * NTEAMS + team;

            auto issue = [&](const int4 &d, int s) {
                const int a   = d.z & ~3;
                const int cnt = ((d.w - a) + 3) & ~3;
                if(cnt > 0)
                {
                    unsigned char *st = stage0 + (size_t)s * stage_bytes;
                    mbar_expect_tx(bar + s, (unsigned)(cnt * (sizeof(T) + sizeof(aoclsparse_int))));
                    bulk_load_stream(st, val + a, (unsigned)(cnt * sizeof(T)), bar + s);
                    bulk_load_stream(st + (size_t)cap * sizeof(T), col_hot + a, (unsigned)(cnt * sizeof(aoclsparse_int)), bar + s);
                }
            };
            const int4 zero4 = make_int4(0, 0, 0, 0);
            // block whose products are being formed (`n`: next) and block whose rows are being summed (`c`: current)
            int  bn = b0;
            int4 dn = bn < n_blocks ? desc[bn] : zero4;
            int  kn = bn < n_blocks ? kind[bn] : 0;
            int4 dc = zero4;
            int  kc = 0;
            bool have_c = false;
            if(t == 0)
            {
                mbar_init(bar, 1);
                mbar_init(bar + 1, 1);
                mbar_init_fence();
                if(bn < n_blocks)
                    issue(dn, 0); // the matrix slices do not depend on x: they fly while the table is filled
                if(bn + total_teams < n_blocks)
                    issue(desc[bn + total_teams], 1);
            }
            for(int i = tid; i < table_entries; i += HT_THREADS)
                table[i] = ldg_ro(x + hot_cols[i]);
            __syncthreads();

            auto xv = [&](int c) -> T { return c < 0 ? table[c & 0x7fffffff] : ldg_ro(x + c); };

            unsigned ph = 0; // phase parity of the two stage barriers (bit s)
            int      sn = 0; // stage of block bn (block bc sits in the other one)
            unsigned it = 0;
            // Software pipeline over the team's blocks: the gathers of block bn are issued, the rows of the previous
            // block bc are summed while they fly, then the products of bn are written -- ONE team barrier per block,
            // and a warp has gathers outstanding for most of its time.
            while(bn < n_blocks || have_c)
            {
                const bool have_n = bn < n_blocks;
                // descriptor of the block after bn: its bulk copy is issued at the end of this iteration
                int4 dnn = zero4;
                int  knn = 0;
                if(have_n && bn + total_teams < n_blocks)
                {
                    dnn = desc[bn + total_teams];
                    knn = kind[bn + total_teams];
                }
                // ---- (1) block bn: wait for its slice, read the columns, issue the gathers
                const int a_n     = dn.z & ~3;
                const int first_n = dn.z - a_n, total_n = have_n ? dn.w - dn.z : 0;
                T        *sval_n  = reinterpret_cast<T *>(stage0 + (size_t)sn * stage_bytes);
                int      *scol_n  = reinterpret_cast<int *>(stage0 + (size_t)sn * stage_bytes + (size_t)cap * sizeof(T));
                T         xx[U];
                if(have_n)
                {
                    if((((dn.w - a_n) + 3) & ~3) > 0)
                    {
                        mbar_wait(bar + sn, (ph >> sn) & 1u);
                        ph ^= 1u << sn;
                    }
                    int c[U];
#pragma unroll
                    for(int u = 0; u < U; ++u)
                        c[u] = t + u * TEAM < total_n ? scol_n[first_n + t + u * TEAM] : HOT_BIT;
#pragma unroll
                    for(int u = 0; u < U; ++u)
                        xx[u] = xv(c[u]);
                }
                // ---- (2) block bc: per-row sums of its finished products; 32 consecutive rows per warp pass
                if(have_c && (kc & 15) != STRAT_LONG)
                {
                    const int a_c    = dc.z & ~3;
                    const T  *sval_c = reinterpret_cast<const T *>(stage0 + (size_t)(sn ^ 1) * stage_bytes);
                    for(int rb = dc.x + warp * 32; rb < dc.y; rb += TEAM)
                    {
                        const int  r     = rb + lane;
                        const bool valid = r < dc.y;
                        int        ss = 0, ee = 0;
                        if(valid)
                        {
                            ss = rp[r] - a_c;
                            ee = rp[r + 1] - a_c;
                        }
                        T acc = vt<T>::zero();
                        if(ee - ss <= SHORT_ROW)
                            for(int j = ss; j < ee; ++j)
                                acc = add(acc, sval_c[j]);
                        unsigned pending = __ballot_sync(0xffffffffu, valid && (ee - ss > SHORT_ROW));
                        while(pending)
                        {
                            const int src = __ffs(pending) - 1;
                            pending &= pending - 1;
                            const int js   = __shfl_sync(0xffffffffu, ss, src);
                            const int je   = __shfl_sync(0xffffffffu, ee, src);
                            T         part = vt<T>::zero();
                            for(int j = js + lane; j < je; j += 32)
                                part = add(part, sval_c[j]);
                            part = warp_sum(part);
                            if(lane == src)
                                acc = part;
                        }
                        if(valid)
                            y[r] = axpby_out(alpha, acc, beta, beta_zero != 0, y + r);
                    }
                }
                // ---- (3) block bn: products in place (or, for a segment of a split row, the segment's partial sum)
                const bool long_n = have_n && (kn & 15) == STRAT_LONG;
                if(have_n && !long_n)
                {
#pragma unroll
                    for(int u = 0; u < U; ++u)
                        if(t + u * TEAM < total_n)
                            sval_n[first_n + t + u * TEAM] = mul(sval_n[first_n + t + u * TEAM], xx[u]);
                }
                else if(long_n)
                {
                    T acc = vt<T>::zero();
#pragma unroll
                    for(int u = 0; u < U; ++u)
                        if(t + u * TEAM < total_n)
                            acc = mad(sval_n[first_n + t + u * TEAM], xx[u], acc);
                    acc = warp_sum(acc);
                    if(lane == 0)
                        s_part[(it & 1u) * TW + warp] = acc;
                }
                // ---- (4) the products of bn are visible to the team, everybody is done with bc's stage
                fence_proxy_async_smem();
                team_sync<TEAM>(team);
                if(t == 0)
                {
                    if(have_c && have_n && bn + total_teams < n_blocks)
                        issue(dnn, sn ^ 1); // block after bn goes where bc was
                    if(long_n)
                    {
                        T tot = s_part[(it & 1u) * TW];
#pragma unroll
                        for(int w = 1; w < TW; ++w)
                            tot = add(tot, s_part[(it & 1u) * TW + w]);
                        partials[kn >> 4] = tot;
                    }
                }
                dc     = dn;
                kc     = kn;
                have_c = have_n;
                dn     = dnn;
                kn     = knn;
                bn     = have_n ? bn + total_teams : bn;
                sn ^= 1;
                ++it;
            }
        }

        template <typename T, int TEAM>
        aoclsparse_status launch_hot_teams(const dev_csr &A, const T *x, T *y, T alpha, T beta, cudaStream_t st)
        {
            const row_block_plan &P      = A.plan;
            constexpr int         NTEAMS = HT_THREADS / TEAM;
            const int             cap    = P.block_nnz + 8;
            const size_t          smem   = ht_smem_bytes(sizeof(T), P.hot_entries, NTEAMS, cap);
            static std::atomic<size_t> configured{0};
            if(configured.load(std::memory_order_acquire) != smem)
            {
                B200_CUDA(cudaFuncSetAttribute(spmv_hot_teams_kernel<T, TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                // the smallest carve-out that holds one CTA: everything else stays L1 (see the header)
                int pct = (int)((smem + 1024 + 2047) * 100 / (228 * 1024)) + 1;
                if(pct > 100)
                    pct = 100;
                B200_CUDA(cudaFuncSetAttribute(spmv_hot_teams_kernel<T, TEAM>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
                configured.store(smem, std::memory_order_release);
            }
            int sms = 148, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int want = (P.n_blocks + NTEAMS - 1) / NTEAMS;
            const int grid = want < sms ? want : sms;
            spmv_hot_teams_kernel<T, TEAM><<<grid, HT_THREADS, smem, st>>>(P.desc.as<int4>(),
                                                                            P.kind.as<int>(),
                                                                            (int)P.n_blocks,
                                                                            cap,
                                                                            A.row_ptr.as<aoclsparse_int>(),
                                                                            P.col_hot.as<aoclsparse_int>(),
                                                                            A.val.as<T>(),
                                                                            x,
                                                                            y,
                                                                            alpha,
                                                                            beta,
                                                                            is_zero(beta) ? 1 : 0,
                                                                            P.partials.as<T>(),
                                                                            P.hot_cols.as<aoclsparse_int>(),
                                                                            (int)P.hot_entries);
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }
    }

    // entries <= 0: as many as fit HOT_TABLE_BYTES of shared memory; team_threads 128 or 256
    aoclsparse_status build_hot_table(dev_csr &A, size_t elem_size, long long entries, int team_threads, bool force, cudaStream_t st)
    {
        row_block_plan &P = A.plan;
        P.hot_entries     = 0;
        P.hot_mass        = 0.0;
        P.hot_cols.release();
        P.col_hot.release();
        P.hot_state = 1;
        if(!P.valid || elem_size > 8 || A.nnz <= 0 || A.n <= 0)
            return aoclsparse_status_success;
        if(!force && (A.nnz < (1 << 22) || (size_t)A.n * elem_size < (size_t)(8u << 20)))
            return aoclsparse_status_success; // small problems: x lives in L1 / L2 lines that are re-used anyway
        if(team_threads != 64 && team_threads != 128 && team_threads != 256)
            team_threads = 128;
        // a thread gathers all of its entries of a block at once: HT_U per thread
        while(team_threads < 256 && P.block_nnz > HT_U * team_threads)
            team_threads *= 2;
        if(P.block_nnz > HT_U * team_threads)
            return aoclsparse_status_success;
        const int    teams = HT_THREADS / team_threads;
        const int    cap   = P.block_nnz + 8;
        const size_t fixed = ht_smem_bytes(elem_size, 0, teams, cap);
        const size_t limit = 232448 - 1024; // one CTA per SM
        if(fixed + 4096 > limit)
            return aoclsparse_status_success;
        long long K = entries > 0 ? entries : (long long)(HOT_TABLE_BYTES / elem_size);
        if(K > (long long)((limit - fixed) / elem_size))
            K = (long long)((limit - fixed) / elem_size);
        if(K > A.n)
            K = A.n;
        K &= ~3LL;
        if(K < 4)
            return aoclsparse_status_success;
        const long long n = A.n, nnz = A.nnz;
        dev_buf         cnt, cnt_sorted, ids, ids_sorted, temp, slot_of;
        B200_TRY(cnt.alloc(4 * (size_t)n));
        B200_TRY(cnt_sorted.alloc(4 * (size_t)n));
        B200_TRY(ids.alloc(4 * (size_t)n));
        B200_TRY(ids_sorted.alloc(4 * (size_t)n));
        B200_CUDA(cudaMemsetAsync(cnt.p, 0, 4 * (size_t)n, st));
        col_hist_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, A.col_idx.as<int>(), cnt.as<unsigned>());
        B200_LAUNCHED();
        hot_iota_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, ids.as<int>());
        B200_LAUNCHED();
        size_t tb = 0;
        B200_CUDA(cub::DeviceRadixSort::SortPairsDescending(
            nullptr, tb, cnt.as<unsigned>(), cnt_sorted.as<unsigned>(), ids.as<int>(), ids_sorted.as<int>(), (int)n, 0, 32, st));
        B200_TRY(temp.alloc(tb));
        B200_CUDA(cub::DeviceRadixSort::SortPairsDescending(
            temp.p, tb, cnt.as<unsigned>(), cnt_sorted.as<unsigned>(), ids.as<int>(), ids_sorted.as<int>(), (int)n, 0, 32, st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        std::vector<unsigned> top((size_t)K);
        B200_CUDA(cudaMemcpyAsync(top.data(), cnt_sorted.p, 4 * (size_t)K, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        long long mass = 0;
        for(unsigned c : top)
            mass += c;
        if(!force && mass * 100 < (long long)HOT_MIN_MASS_PCT * nnz)
            return aoclsparse_status_success; // flat column distribution: nothing worth keeping on chip
        B200_TRY(P.hot_cols.alloc(4 * (size_t)K));
        B200_TRY(P.col_hot.alloc(4 * (size_t)nnz));
        B200_TRY(slot_of.alloc(4 * (size_t)n));
        B200_CUDA(cudaMemcpyAsync(P.hot_cols.p, ids_sorted.p, 4 * (size_t)K, cudaMemcpyDeviceToDevice, st));
        B200_CUDA(cudaMemsetAsync(slot_of.p, 0xff, 4 * (size_t)n, st));
        slot_scatter_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>((int)K, P.hot_cols.as<int>(), slot_of.as<int>());
        B200_LAUNCHED();
        remap_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, A.col_idx.as<int>(), slot_of.as<int>(), P.col_hot.as<int>());
        B200_LAUNCHED();
        B200_CUDA(cudaStreamSynchronize(st));
        P.hot_entries = (aoclsparse_int)K;
        P.hot_team    = team_threads;
        P.hot_mass    = (double)mass / (double)nnz;
        return aoclsparse_status_success;
    }

    template <typename T>
    aoclsparse_status launch_hot(const dev_csr &A, const T *x, T *y, T alpha, T beta, cudaStream_t st)
    {
        if(A.plan.hot_team == 256)
            return launch_hot_teams<T, 256>(A, x, y, alpha, beta, st);
        if(A.plan.hot_team == 64)
            return launch_hot_teams<T, 64>(A, x, y, alpha, beta, st);
        return launch_hot_teams<T, 128>(A, x, y, alpha, beta, st);
    }

    template aoclsparse_status launch_hot<float>(const dev_csr &, const float *, float *, float, float, cudaStream_t);
    template aoclsparse_status launch_hot<double>(const dev_csr &, const double *, double *, double, double, cudaStream_t);
}
