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

        // 4-byte values: only the column slice is staged; the gatherers read the values straight from global memory
        // (coalesced) and a product takes the place of its column.  8-byte values: values + columns are staged, the
        // product takes the place of the value.
        __host__ __device__ inline size_t ht_stage_bytes(size_t elem, int cap)
        {
            return (size_t)cap * (elem == 4 ? 4 : elem + 4);
        }
        template <typename T>
        __device__ __forceinline__ T ldg_stream(const T *p, uint64_t policy);
        template <>
        __device__ __forceinline__ float ldg_stream<float>(const float *p, uint64_t policy)
        {
            float v;
            asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy));
            return v;
        }
        template <>
        __device__ __forceinline__ double ldg_stream<double>(const double *p, uint64_t policy)
        {
            double v;
            asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
            return v;
        }
        // two stages per team: the next block's slice arrives while the current one is reduced
        inline size_t ht_smem_bytes(size_t elem, int entries, int teams, int cap)
        {
            return HT_HEADER + ((((size_t)entries * elem) + 15) & ~(size_t)15) + (size_t)teams * 2 * ht_stage_bytes(elem, cap);
        }

        __device__ __forceinline__ void mbar_arrive(uint64_t *bar)
        {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        }

        // One CTA of 1024 threads per SM = 8 teams of 4 warps.  Warp 0 of a team is its REDUCER (it also issues the team's
        // bulk copies), warps 1..3 are its GATHERERS; the team walks the plan's row blocks b = first, first + stride, ...
        // through a ring of two stages:
        //     reducer, lane 0 : bulk copy of block i+2's val / col_hot slice into stage i & 1  -> full[s]
        //     gatherers       : wait full[s]; every staged entry becomes its product with x in place, all of a thread's
        //                       (<= HT_U) gathers issued together, hot columns from the table  -> gathered[s] (3 arrivals)
        //     reducer         : wait gathered[s]; per-row sums, alpha / beta, y; then refills the stage
        // No barrier couples the two roles: the gatherers of a team have gathers outstanding for as long as the reducer
        // keeps up, and the 8 teams are independent of each other.
        template <typename T>
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
            constexpr int TEAM   = 128;
            constexpr int NTEAMS = HT_THREADS / TEAM;
            constexpr int GT     = TEAM - 32; // gather threads of a team
            constexpr int U      = HT_U;
            constexpr bool COLONLY = sizeof(T) == 4;
            extern __shared__ __align__(16) unsigned char smem_raw[];
            uint64_t      *bars        = reinterpret_cast<uint64_t *>(smem_raw); // [team]: full[2], gathered[2]
            T             *table       = reinterpret_cast<T *>(smem_raw + HT_HEADER);
            const size_t   table_bytes = (((size_t)table_entries * sizeof(T)) + 15) & ~(size_t)15;
            const int      tid = threadIdx.x, team = tid / TEAM, t = tid % TEAM, lane = t & 31, warp = t >> 5;
            const size_t   stage_bytes = ht_stage_bytes(sizeof(T), cap);
            const size_t   col_off     = COLONLY ? 0 : (size_t)cap * sizeof(T); // the columns inside a stage
            unsigned char *stage0      = smem_raw + HT_HEADER + table_bytes + (size_t)team * 2 * stage_bytes;
            uint64_t      *full        = bars + 4 * team;
            uint64_t      *gathered    = full + 2;
            T             *s_part      = reinterpret_cast<T *>(smem_raw + 512) + team * 8; // [stage][gather warp]

            const int total_teams = gridDim.x * NTEAMS;
            const int b0          = blockIdx.x * NTEAMS + team;

            auto issue = [&](const int4 &d, int s) {
                const int a   = d.z & ~3;
                const int cnt = ((d.w - a) + 3) & ~3;
                if(cnt > 0)
                {
                    unsigned char *st = stage0 + (size_t)s * stage_bytes;
                    mbar_expect_tx(full + s, (unsigned)(cnt * ((COLONLY ? 0 : sizeof(T)) + sizeof(aoclsparse_int))));
                    if(!COLONLY)
                        bulk_load_stream(st, val + a, (unsigned)(cnt * sizeof(T)), full + s);
                    bulk_load_stream(st + col_off, col_hot + a, (unsigned)(cnt * sizeof(aoclsparse_int)), full + s);
                }
                else
                    mbar_arrive(full + s); // nothing to copy: the phase still completes
            };
            if(t == 0)
            {
                mbar_init(full, 1);
                mbar_init(full + 1, 1);
                mbar_init(gathered, 3);
                mbar_init(gathered + 1, 3);
                mbar_init_fence();
                // the matrix slices do not depend on x: they fly while the table is filled
                if(b0 < n_blocks)
                    issue(desc[b0], 0);
                if(b0 + total_teams < n_blocks)
                    issue(desc[b0 + total_teams], 1);
            }
            for(int i = tid; i < table_entries; i += HT_THREADS)
                table[i] = ldg_ro(x + hot_cols[i]);
            __syncthreads();

            if(warp != 0)
            {
                // ------------------------------------------------------------------ gatherers
                const int gt = t - 32, gw = warp - 1;
                uint64_t  policy;
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
                int4      d  = b0 < n_blocks ? desc[b0] : make_int4(0, 0, 0, 0);
                int       k  = b0 < n_blocks ? kind[b0] : 0;
                unsigned  i  = 0;
                for(int b = b0; b < n_blocks; b += total_teams, ++i)
                {
                    int4 dn = make_int4(0, 0, 0, 0);
                    int  kn = 0;
                    if(b + total_teams < n_blocks)
                    {
                        dn = desc[b + total_teams];
                        kn = kind[b + total_teams];
                    }
                    const int s = i & 1u;
                    const int first = d.z - (d.z & ~3), total = d.w - d.z;
                    T         vv[U];
                    if constexpr(COLONLY)
                    {
                        // the block's values, coalesced, straight from global memory (read once: evict-first in L2)
#pragma unroll
                        for(int u = 0; u < U; ++u)
                            vv[u] = gt + u * GT < total ? ldg_stream(val + d.z + gt + u * GT, policy) : vt<T>::zero();
                    }
                    mbar_wait(full + s, (i >> 1) & 1u);
                    T        *sval  = reinterpret_cast<T *>(stage0 + (size_t)s * stage_bytes); // products (COLONLY: over the columns)
                    int      *scol  = reinterpret_cast<int *>(stage0 + (size_t)s * stage_bytes + col_off);
                    int       c[U];
                    T         xx[U];
#pragma unroll
                    for(int u = 0; u < U; ++u)
                        c[u] = gt + u * GT < total ? scol[first + gt + u * GT] : HOT_BIT;
#pragma unroll
                    for(int u = 0; u < U; ++u)
                        xx[u] = c[u] < 0 ? table[c[u] & 0x7fffffff] : ldg_ro(x + c[u]);
                    if constexpr(!COLONLY)
                    {
#pragma unroll
                        for(int u = 0; u < U; ++u)
                            vv[u] = gt + u * GT < total ? sval[first + gt + u * GT] : vt<T>::zero();
                    }
                    if((k & 15) != STRAT_LONG)
                    {
#pragma unroll
                        for(int u = 0; u < U; ++u)
                            if(gt + u * GT < total)
                                sval[first + gt + u * GT] = mul(vv[u], xx[u]);
                    }
                    else
                    {
                        T acc = vt<T>::zero();
#pragma unroll
                        for(int u = 0; u < U; ++u)
                            if(gt + u * GT < total)
                                acc = mad(vv[u], xx[u], acc);
                        acc = warp_sum(acc);
                        if(lane == 0)
                            s_part[s * 4 + gw] = acc;
                    }
                    __syncwarp();
                    if(lane == 0)
                        mbar_arrive(gathered + s);
                    d = dn;
                    k = kn;
                }
            }
            else
            {
                // ------------------------------------------------------------------ reducer (+ the team's bulk copies)
                int4     d = b0 < n_blocks ? desc[b0] : make_int4(0, 0, 0, 0);
                int      k = b0 < n_blocks ? kind[b0] : 0;
                unsigned i = 0;
                for(int b = b0; b < n_blocks; b += total_teams, ++i)
                {
                    int4 dn = make_int4(0, 0, 0, 0);
                    int  kn = 0;
                    if(b + total_teams < n_blocks)
                    {
                        dn = desc[b + total_teams];
                        kn = kind[b + total_teams];
                    }
                    int4 dnn = make_int4(0, 0, 0, 0);
                    if(b + 2 * total_teams < n_blocks)
                        dnn = desc[b + 2 * total_teams];
                    const int  s       = i & 1u;
                    const int  a       = d.z & ~3;
                    const bool is_long = (k & 15) == STRAT_LONG;
                    // bounds of the first 32 rows, requested before the wait
                    int s0 = 0, e0 = 0;
                    if(!is_long && d.x + lane < d.y)
                    {
                        s0 = rp[d.x + lane] - a;
                        e0 = rp[d.x + lane + 1] - a;
                    }
                    mbar_wait(gathered + s, (i >> 1) & 1u);
                    const T *sval = reinterpret_cast<const T *>(stage0 + (size_t)s * stage_bytes);
                    if(!is_long)
                    {
                        for(int rb = d.x; rb < d.y; rb += 32)
                        {
                            const int  r     = rb + lane;
                            const bool valid = r < d.y;
                            int        ss = s0, ee = e0;
                            if(rb != d.x)
                            {
                                ss = ee = 0;
                                if(valid)
                                {
                                    ss = rp[r] - a;
                                    ee = rp[r + 1] - a;
                                }
                            }
                            T acc = vt<T>::zero();
                            if(ee - ss <= SHORT_ROW)
                                for(int j = ss; j < ee; ++j)
                                    acc = add(acc, sval[j]);
                            unsigned pending = __ballot_sync(0xffffffffu, valid && (ee - ss > SHORT_ROW));
                            while(pending)
                            {
                                const int src = __ffs(pending) - 1;
                                pending &= pending - 1;
                                const int js   = __shfl_sync(0xffffffffu, ss, src);
                                const int je   = __shfl_sync(0xffffffffu, ee, src);
                                T         part = vt<T>::zero();
                                for(int j = js + lane; j < je; j += 32)
                                    part = add(part, sval[j]);
                                part = warp_sum(part);
                                if(lane == src)
                                    acc = part;
                            }
                            if(valid)
                                y[r] = axpby_out(alpha, acc, beta, beta_zero != 0, y + r);
                        }
                    }
                    else if(lane == 0)
                        partials[k >> 4] = add(add(s_part[s * 4], s_part[s * 4 + 1]), s_part[s * 4 + 2]);
                    // the stage is free: refill it with the block after next
                    fence_proxy_async_smem();
                    __syncwarp();
                    if(lane == 0 && b + 2 * total_teams < n_blocks)
                        issue(dnn, s);
                    d = dn;
                    k = kn;
                }
            }
        }

        template <typename T>
        aoclsparse_status launch_hot_teams(const dev_csr &A, const T *x, T *y, T alpha, T beta, cudaStream_t st)
        {
            const row_block_plan &P      = A.plan;
            constexpr int         NTEAMS = HT_THREADS / 128;
            const int             cap    = P.block_nnz + 8;
            const size_t          smem   = ht_smem_bytes(sizeof(T), P.hot_entries, NTEAMS, cap);
            static std::atomic<size_t> configured{0};
            if(configured.load(std::memory_order_acquire) != smem)
            {
                B200_CUDA(cudaFuncSetAttribute(spmv_hot_teams_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                // the smallest carve-out that holds one CTA: everything else stays L1 (see the header)
                int pct = (int)((smem + 1024 + 2047) * 100 / (228 * 1024)) + 1;
                if(pct > 100)
                    pct = 100;
                B200_CUDA(cudaFuncSetAttribute(spmv_hot_teams_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
                configured.store(smem, std::memory_order_release);
            }
            int sms = 148, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int want = (P.n_blocks + NTEAMS - 1) / NTEAMS;
            const int grid = want < sms ? want : sms;
            spmv_hot_teams_kernel<T><<<grid, HT_THREADS, smem, st>>>(P.desc.as<int4>(),
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
        team_threads = 128; // 1 reducer warp + 3 gather warps
        // a gather thread takes all of its entries of a block at once: at most HT_U
        if(P.block_nnz > HT_U * (team_threads - 32))
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
        return launch_hot_teams<T>(A, x, y, alpha, beta, st);
    }

    template aoclsparse_status launch_hot<float>(const dev_csr &, const float *, float *, float, float, cudaStream_t);
    template aoclsparse_status launch_hot<double>(const dev_csr &, const double *, double *, double, double, cudaStream_t);
}
