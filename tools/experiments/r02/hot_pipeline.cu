// hot.cu -- SpMV for matrices whose columns are hit very unevenly (power-law graphs, BASELINE config 3): analysis
// (hot-column table) and the persistent, warp-specialised kernel that uses it.
//
// Replaces, for that class of matrices, the same reference kernels as spmv_kernels.cuh
// (aoclsparse_csrmv_vectorized<float>, library/src/level2/aoclsparse_csrmv_kr.hpp:734-831: an OpenMP static row split whose
// threads gather x through the CPU cache hierarchy).
//
// Why: on R-MAT scale 24 the row-block kernel is bound by L2 -> SM traffic, not by HBM: every x[col] is a 4-byte gather
// that moves a 32-byte sector (263 M sectors = 8.4 GB per multiply against 2.3 GB of compulsory HBM traffic; ncu: L1 hit
// rate 6 %, profiles/r01_ncu_c3.txt).  Sorting a block's entries by column does not help -- 32 consecutive sorted
// entries of a 2048-entry block still touch 30 distinct sectors (tools/experiments/c3_sector_model.py) -- but the column
// popularity is very uneven: the 36 K most frequent columns carry ~39 % of all stored entries.  The hardware L1 cannot
// exploit that (128-byte lines, LRU thrashed by the 60 % cold gathers), a table in shared memory can: element
// granularity and no eviction.
//
// aoclsparse_optimize (general, non-transposed mv hint, skewed row lengths, memory not restricted):
//   1. column histogram (one atomic per stored entry), radix sort of the counts (CUB, analysis time only);
//   2. the K most frequent columns fill the shared memory one persistent CTA per SM can spare for the table;
//   3. a second column array in which those columns are replaced by HOT_BIT | slot.
// Kernel: ONE CTA of 1024 threads per SM, resident for the whole launch, table loaded once.  Three roles, connected by
// mbarriers over a ring of S staged row blocks (this is the shape the round-1 experiment lacked: its four thread groups
// each did load -> products -> barrier -> row sums in sequence and spent most of their time waiting):
//   * warp 0, one lane : TMA producer.  Bulk copies of the block's val / col_hot slice and of its row_ptr window into
//                        the next free stage (evict-first in L2), descriptor alongside.
//   * warps 9..31      : gatherers.  Every staged entry becomes its product with x in place; hot columns read the
//                        table (LDS), the others go to L1/L2 with four gathers in flight per thread.
//   * warps 1..8       : reducers.  Per-row sums out of the finished stage (short rows by one lane, long rows by a
//                        warp, one segment of a split row by all eight warps), alpha / beta, y.
// Gathering block i+1 overlaps reducing block i and loading block i+2.  Sums are formed in a fixed order (run-to-run
// reproducible); rows split across blocks leave partial sums that finish_long_rows_kernel adds, as in the row-block
// kernel.
#include "spmv_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

namespace b200
{
    namespace
    {
        constexpr int HP_THREADS      = 1024;
        constexpr int HP_REDUCE_WARPS = 8;                                 // warps 1..8
        constexpr int HP_GATHER_WARP0 = 1 + HP_REDUCE_WARPS;               // warps 9..31
        constexpr int HP_GATHER_WARPS = HP_THREADS / 32 - HP_GATHER_WARP0; // 23
        constexpr int HP_U            = 8;                                 // gathers in flight per gather thread
        constexpr int HP_RT           = HP_REDUCE_WARPS * 32;              // 256 reduce threads
        constexpr int HP_MAX_STAGES   = 8;
        constexpr int HP_HEADER       = 512; // 3 x 8 mbarriers (192 B), 8 partial sums (64 B), 8 descriptors, 8 kinds
        constexpr int HOT_BIT         = (int)0x80000000;

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
        __global__ void iota_kernel(long long n, int *out)
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

        __device__ __forceinline__ void mbar_arrive(uint64_t *bar)
        {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        __device__ __forceinline__ void reducers_sync()
        {
            asm volatile("bar.sync 1, %0;" ::"n"(HP_RT) : "memory");
        }

        // per-stage sizes in bytes (all multiples of 16)
        __host__ __device__ inline size_t hp_stage_bytes(size_t elem, int cap, int rcap)
        {
            return (size_t)cap * elem + (size_t)cap * 4 + (size_t)rcap * 4;
        }

        template <typename T>
        __global__ void __launch_bounds__(HP_THREADS, 1) spmv_hot_pipeline_kernel(const int4 *__restrict__ desc,
                                                                                 const int *__restrict__ kind,
                                                                                 int n_blocks,
                                                                                 int cap,  // staged entries per stage
                                                                                 int rcap, // staged row_ptr entries per stage
                                                                                 int n_stages,
                                                                                 int n_teams, // gather teams: blocks in the gather phase at once
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
            uint64_t *full     = reinterpret_cast<uint64_t *>(smem_raw);
            uint64_t *gathered = full + HP_MAX_STAGES;
            uint64_t *freeb    = gathered + HP_MAX_STAGES;
            T        *s_part   = reinterpret_cast<T *>(smem_raw + 192); // 8 partial sums of a split-row segment
            int4     *s_desc   = reinterpret_cast<int4 *>(smem_raw + 256); // [stage]: the block's descriptor (128 bytes)
            int      *s_kind   = reinterpret_cast<int *>(smem_raw + 384); // [stage]
            T        *xs       = reinterpret_cast<T *>(smem_raw + HP_HEADER + 16);
            const size_t   table_bytes = (((size_t)table_entries * sizeof(T)) + 15) & ~(size_t)15;
            unsigned char *ring        = smem_raw + HP_HEADER + 16 + table_bytes;
            const size_t   stage_bytes = hp_stage_bytes(sizeof(T), cap, rcap);

            const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            if(tid == 0)
            {
                for(int s = 0; s < n_stages; ++s)
                {
                    mbar_init(&full[s], 1);
                    mbar_init(&gathered[s], HP_GATHER_WARPS / n_teams);
                    mbar_init(&freeb[s], HP_REDUCE_WARPS);
                }
                mbar_init_fence();
            }
            // x may be the previous launch's output: the table is the first thing that reads it
            for(int s = tid; s < table_entries; s += HP_THREADS)
                xs[s] = ldg_ro(x + hot_cols[s]);
            __syncthreads();

            auto stage_val = [&](int s) { return reinterpret_cast<T *>(ring + (size_t)s * stage_bytes); };
            auto stage_col = [&](int s) { return reinterpret_cast<aoclsparse_int *>(ring + (size_t)s * stage_bytes + (size_t)cap * sizeof(T)); };
            auto stage_rp  = [&](int s) {
                return reinterpret_cast<aoclsparse_int *>(ring + (size_t)s * stage_bytes + (size_t)cap * (sizeof(T) + 4));
            };

            if(warp == 0)
            {
                // ------------------------------------------------------------------ producer
                if(lane != 0)
                    return;
                int i = 0;
                for(int b = blockIdx.x; b < n_blocks; b += gridDim.x, ++i)
                {
                    const int s = i % n_stages, use = i / n_stages;
                    if(use > 0)
                        mbar_wait(&freeb[s], (unsigned)((use - 1) & 1)); // the reducers are done with its previous block
                    const int4 d   = desc[b];
                    const int  k   = kind[b];
                    const int  a   = d.z & ~3;
                    const int  cnt = ((d.w - a) + 3) & ~3;
                    const int  ra  = d.x & ~3;
                    const int  rcnt = (k & 15) == STRAT_LONG ? 0 : (((d.y + 1 - ra) + 3) & ~3);
                    s_desc[s]       = d;
                    s_kind[s]       = k;
                    const unsigned bytes = (unsigned)(cnt * (sizeof(T) + 4) + rcnt * 4);
                    if(bytes == 0)
                    {
                        mbar_arrive(&full[s]);
                        continue;
                    }
                    // the stage was last written and read through the generic proxy (products in place, row sums): order
                    // those accesses before the bulk copies (async proxy) that overwrite it
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&full[s], bytes);
                    if(cnt > 0)
                    {
                        bulk_load_stream(stage_val(s), val + a, (unsigned)(cnt * sizeof(T)), &full[s]);
                        bulk_load_stream(stage_col(s), col_hot + a, (unsigned)(cnt * 4), &full[s]);
                    }
                    if(rcnt > 0)
                        bulk_load(stage_rp(s), rp + ra, (unsigned)(rcnt * 4), &full[s]);
                }
            }
            else if(warp >= HP_GATHER_WARP0)
            {
                // ------------------------------------------------------------------ gatherers
                // n_teams teams of wpt warps; team g takes the blocks i = g, g + n_teams, ... so that several blocks are
                // in their gather phase at once (the gathers are latency-bound: what counts is how many are in flight)
                const int wpt  = HP_GATHER_WARPS / n_teams;
                const int team = (warp - HP_GATHER_WARP0) / wpt;
                if(team >= n_teams)
                    return;
                const int gt    = tid - (HP_GATHER_WARP0 + team * wpt) * 32;
                const int HP_GT = wpt * 32;
                int       i     = 0;
                for(int b = blockIdx.x; b < n_blocks; b += gridDim.x, ++i)
                {
                    if(i % n_teams != team)
                        continue;
                    const int s = i % n_stages, use = i / n_stages;
                    mbar_wait(&full[s], (unsigned)(use & 1));
                    const int4            d    = s_desc[s];
                    T                    *sval = stage_val(s);
                    const aoclsparse_int *scol = stage_col(s);
                    const int             f0 = d.z - (d.z & ~3), total = d.w - d.z;
                    for(int q = gt; q < total; q += HP_U * HP_GT)
                    {
                        int  c[HP_U];
                        bool ok[HP_U];
                        T    xv[HP_U];
#pragma unroll
                        for(int u = 0; u < HP_U; ++u)
                        {
                            ok[u] = q + u * HP_GT < total;
                            c[u]  = ok[u] ? scol[f0 + q + u * HP_GT] : 0;
                        }
                        // the L1/L2 gathers of the cold columns are issued together, then the table reads
#pragma unroll
                        for(int u = 0; u < HP_U; ++u)
                            xv[u] = (ok[u] && c[u] >= 0) ? ldg_ro(x + c[u]) : vt<T>::zero();
#pragma unroll
                        for(int u = 0; u < HP_U; ++u)
                            if(ok[u] && c[u] < 0)
                                xv[u] = xs[c[u] & 0x7fffffff];
#pragma unroll
                        for(int u = 0; u < HP_U; ++u)
                            if(ok[u])
                                sval[f0 + q + u * HP_GT] = mul(sval[f0 + q + u * HP_GT], xv[u]);
                    }
                    __syncwarp();
                    if(lane == 0)
                        mbar_arrive(&gathered[s]);
                }
            }
            else
            {
                // ------------------------------------------------------------------ reducers (warps 1..8)
                const int rw = warp - 1, rt = tid - 32;
                int       i  = 0;
                for(int b = blockIdx.x; b < n_blocks; b += gridDim.x, ++i)
                {
                    const int s = i % n_stages, use = i / n_stages;
                    mbar_wait(&gathered[s], (unsigned)(use & 1));
                    const int4            d    = s_desc[s];
                    const int             k    = s_kind[s];
                    const T              *sval = stage_val(s);
                    const aoclsparse_int *srp  = stage_rp(s);
                    const int             a = d.z & ~3;
                    if((k & 15) != STRAT_LONG)
                    {
                        const int ra = d.x & ~3;
                        for(int rb = d.x + rw * 32; rb < d.y; rb += HP_RT)
                        {
                            const int  r     = rb + lane;
                            const bool valid = r < d.y;
                            int        s0 = 0, e = 0;
                            if(valid)
                            {
                                s0 = srp[r - ra] - a;
                                e  = srp[r + 1 - ra] - a;
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
                        const int f0 = d.z - a, total = d.w - d.z;
                        T         acc = vt<T>::zero();
                        for(int q = rt; q < total; q += HP_RT)
                            acc = add(acc, sval[f0 + q]);
                        acc = warp_sum(acc);
                        if(lane == 0)
                            s_part[rw] = acc;
                        reducers_sync();
                        if(rt == 0)
                        {
                            T tot = s_part[0];
#pragma unroll
                            for(int w = 1; w < HP_REDUCE_WARPS; ++w)
                                tot = add(tot, s_part[w]);
                            partials[k >> 4] = tot;
                        }
                        reducers_sync(); // s_part is reused by the next split-row segment
                    }
                    __syncwarp();
                    if(lane == 0)
                        mbar_arrive(&freeb[s]);
                }
            }
        }
    }

    // shared memory the kernel needs for a table of `entries` values and `stages` stages
    static size_t hp_smem_bytes(size_t elem, int entries, int stages, int cap, int rcap)
    {
        return HP_HEADER + 16 + ((((size_t)entries * elem) + 15) & ~(size_t)15) + (size_t)stages * hp_stage_bytes(elem, cap, rcap);
    }

    aoclsparse_status build_hot_table(dev_csr &A, size_t elem_size, cudaStream_t st)
    {
        row_block_plan &P = A.plan;
        P.hot_entries     = 0;
        P.hot_cols.release();
        P.col_hot.release();
        if(!P.valid || A.nnz < (1 << 22) || A.n < (1 << 16) || elem_size > 8)
            return aoclsparse_status_success; // small problems: x lives in L1/L2 anyway
        int stages = 5, teams = 2;
        if(const char *e = getenv("AOCLSPARSE_B200_HOT_STAGES"))
            stages = atoi(e) < 2 ? 2 : (atoi(e) > HP_MAX_STAGES ? HP_MAX_STAGES : atoi(e));
        if(const char *e = getenv("AOCLSPARSE_B200_HOT_TEAMS"))
            teams = atoi(e) < 1 ? 1 : (atoi(e) > 4 ? 4 : atoi(e));
        if(stages < teams + 1)
            stages = teams + 1;
        const int cap = P.block_nnz + 8, rcap = P.block_rows + 8;
        // the table gets what one CTA per SM can spare: 227 KB - header - stages
        const size_t budget = 232448 - 1024;
        const size_t fixed  = hp_smem_bytes(elem_size, 0, stages, cap, rcap);
        if(fixed + 16384 > budget)
            return aoclsparse_status_success;
        long long K = (long long)((budget - fixed) / elem_size) & ~3LL;
        if(const char *e = getenv("AOCLSPARSE_B200_HOT_TABLE"))
            if(atoll(e) >= 1024 && atoll(e) < K)
                K = atoll(e) & ~3LL;
        if(K > A.n)
            K = A.n & ~3LL;
        const long long n = A.n, nnz = A.nnz;
        dev_buf cnt, cnt_sorted, ids, ids_sorted, temp, slot_of;
        B200_TRY(cnt.alloc(4 * (size_t)n));
        B200_TRY(cnt_sorted.alloc(4 * (size_t)n));
        B200_TRY(ids.alloc(4 * (size_t)n));
        B200_TRY(ids_sorted.alloc(4 * (size_t)n));
        B200_CUDA(cudaMemsetAsync(cnt.p, 0, 4 * (size_t)n, st));
        col_hist_kernel<<<grid_for(nnz, 256), 256, 0, st>>>(nnz, A.col_idx.as<int>(), cnt.as<unsigned>());
        B200_LAUNCHED();
        iota_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, ids.as<int>());
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
        if(mass * 100 < 20 * nnz)
            return aoclsparse_status_success; // flat column distribution: nothing to keep on chip
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
        P.hot_stages  = stages;
        P.hot_teams   = teams;
        P.hot_mass    = (double)mass / (double)nnz;
        return aoclsparse_status_success;
    }

    template <typename T>
    aoclsparse_status launch_hot(const dev_csr &A, const T *x, T *y, T alpha, T beta, cudaStream_t st)
    {
        const row_block_plan &P    = A.plan;
        const int             cap  = P.block_nnz + 8, rcap = P.block_rows + 8;
        const size_t          smem = hp_smem_bytes(sizeof(T), P.hot_entries, P.hot_stages, cap, rcap);
        static std::atomic<size_t> configured{0};
        if(configured.load() < smem)
        {
            B200_CUDA(cudaFuncSetAttribute(spmv_hot_pipeline_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured.store(smem);
        }
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = P.n_blocks < sms ? P.n_blocks : sms;
        spmv_hot_pipeline_kernel<T><<<grid, HP_THREADS, smem, st>>>(P.desc.as<int4>(),
                                                                      P.kind.as<int>(),
                                                                      (int)P.n_blocks,
                                                                      cap,
                                                                      rcap,
                                                                      P.hot_stages,
                                                                      P.hot_teams,
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

    template aoclsparse_status launch_hot<float>(const dev_csr &, const float *, float *, float, float, cudaStream_t);
    template aoclsparse_status launch_hot<double>(const dev_csr &, const double *, double *, double, double, cudaStream_t);
}
