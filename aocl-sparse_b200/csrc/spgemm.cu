// spgemm.cu -- sparse x sparse product C = op(A) op(B) on the device: aoclsparse_sp2m / aoclsparse_spmm
// (SURVEY.md 8(f) row 4), plus the two entries a caller needs around it: aoclsparse_export_?csr and aoclsparse_order_mat.
//
// Reference: aoclsparse::sp2m<T> (library/src/level3/aoclsparse_csr2m.cpp:592-860) validates, resolves CSR / CSC
// storage and the two operations into "transpose this operand?" + "conjugate its values?", and runs Gustavson's
// row-by-row algorithm with a dense marker array of n entries per thread: aoclsparse_csr2m_nnz_count (:46-305) counts
// the distinct columns of every row of C, aoclsparse_csr2m_finalize (:310-540) accumulates the values.  C is a new
// library-owned CSR matrix, always base 0, structural (numerically cancelling entries stay).
//
// B200: a dense marker per row does not fit a GPU (thousands of rows in flight), so the marker becomes a hash table
// sized from an upper bound of the row (the number of scalar products of that row):
//   tier WARP    <=   96 products   table  128 in shared memory, one warp per row, lanes over the entries of A's row
//   tier CTA_S   <=  768 products   table 1024 in shared memory, 128 threads per row, warps over A's entries
//   tier CTA_L   <= 6144 products   table 8192 in shared memory, 256 threads per row
//   tier GLOBAL  beyond             table of 2 x min(products, n) [symbolic] / 2 x nnz(row) [numeric] slots in global memory
// Pass 1 (nnz count) inserts column indices only, an exclusive scan gives row_ptr; pass 2 (finalize) inserts again with
// atomic accumulation of the values, compacts every table into its row of C and a segmented sort by column puts the
// rows in ascending order (the reference leaves them in first-touch order; both are valid CSR, ours is reproducible in
// structure).  Value sums are formed by atomic adds, so their rounding can differ from run to run in the last bits.
// Integer work (row_ptr, sorted col_idx) is bit-exact against the reference; tests/test_parity_gpu.py.
#include "common.hpp"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>

#include <algorithm>
#include <chrono>

namespace b200
{
    namespace
    {
        // AOCLSPARSE_B200_SPGEMM_TRACE=1: per-phase wall times (stream-synchronised) on stderr, for profiles/
        struct phase_trace
        {
            bool                                           on;
            cudaStream_t                                   st;
            std::chrono::time_point<std::chrono::steady_clock> t0;
            explicit phase_trace(cudaStream_t s)
                : on(getenv("AOCLSPARSE_B200_SPGEMM_TRACE") && atoi(getenv("AOCLSPARSE_B200_SPGEMM_TRACE")) != 0)
                , st(s)
            {
                if(on)
                {
                    cudaStreamSynchronize(st);
                    t0 = std::chrono::steady_clock::now();
                }
            }
            void mark(const char *what)
            {
                if(!on)
                    return;
                cudaStreamSynchronize(st);
                const auto t1 = std::chrono::steady_clock::now();
                fprintf(stderr, "[spgemm] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
                t0 = t1;
            }
        };

        constexpr int TIER_NONE = 0, TIER_WARP = 1, TIER_CTA_S = 2, TIER_CTA_L = 3, TIER_GLOBAL = 4, N_TIERS = 5;
        constexpr int TAB_WARP = 128, UB_WARP = 96;
        constexpr int TAB_S = 1024, UB_S = 768, THREADS_S = 128;
        constexpr int THREADS_L = 256;
        // large CTA tier: 8192 slots for values of up to 8 bytes, 4096 for 16-byte values (shared memory budget)
        template <typename T>
        struct large_tier
        {
            static constexpr int TAB = sizeof(T) > 8 ? 4096 : 8192;
            static constexpr int UB  = TAB / 4 * 3;
        };
        constexpr int RANK_LIMIT = 1024; // rows up to this length are ordered by counting ranks, longer ones bitonically
        constexpr int THREADS_W = 128; // four rows per CTA in the warp tier

        __device__ __forceinline__ unsigned first_slot(int key, unsigned mask)
        {
            return (((unsigned)key * 2654435761u) >> 7) & mask;
        }

        // returns the slot of key; *fresh = 1 when this call created it
        __device__ __forceinline__ unsigned insert_key(int *tab, unsigned mask, int key, int *fresh)
        {
            unsigned h = first_slot(key, mask);
            while(true)
            {
                int cur = tab[h];
                if(cur == key)
                {
                    *fresh = 0;
                    return h;
                }
                if(cur == -1)
                {
                    cur = atomicCAS(&tab[h], -1, key);
                    if(cur == -1)
                    {
                        *fresh = 1;
                        return h;
                    }
                    if(cur == key)
                    {
                        *fresh = 0;
                        return h;
                    }
                }
                h = (h + 1) & mask;
            }
        }

        __device__ __forceinline__ void atomic_add_val(float *p, float v)
        {
            atomicAdd(p, v);
        }
        __device__ __forceinline__ void atomic_add_val(double *p, double v)
        {
            atomicAdd(p, v);
        }
        __device__ __forceinline__ void atomic_add_val(float2 *p, float2 v)
        {
            atomicAdd(&p->x, v.x);
            atomicAdd(&p->y, v.y);
        }
        __device__ __forceinline__ void atomic_add_val(double2 *p, double2 v)
        {
            atomicAdd(&p->x, v.x);
            atomicAdd(&p->y, v.y);
        }

        // products per row of C (upper bound of its length), tier of the row, tier histogram
        __global__ void spgemm_bound_kernel(int m,
                                            const int *__restrict__ rpA,
                                            const int *__restrict__ colA,
                                            const int *__restrict__ rpB,
                                            int ub_large,
                                            int *__restrict__ ub,
                                            unsigned char *__restrict__ tier,
                                            int *__restrict__ hist)
        {
            __shared__ int h[N_TIERS];
            if(threadIdx.x < N_TIERS)
                h[threadIdx.x] = 0;
            __syncthreads();
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if(i < m)
            {
                long long s = 0;
                for(int j = rpA[i]; j < rpA[i + 1]; ++j)
                {
                    const int c = colA[j];
                    s += rpB[c + 1] - rpB[c];
                }
                const int t = s == 0 ? TIER_NONE
                                     : (s <= UB_WARP ? TIER_WARP : (s <= UB_S ? TIER_CTA_S : (s <= ub_large ? TIER_CTA_L : TIER_GLOBAL)));
                ub[i]   = s > 0x7fffffffLL ? 0x7fffffff : (int)s;
                tier[i] = (unsigned char)t;
                atomicAdd(&h[t], 1);
            }
            __syncthreads();
            if(threadIdx.x < N_TIERS && h[threadIdx.x])
                atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
        }

        // rows of every tier, contiguous per tier (order inside a tier is irrelevant to the result)
        __global__ void spgemm_place_kernel(int m, const unsigned char *__restrict__ tier, const int *__restrict__ start,
                                            int *__restrict__ cursor, int *__restrict__ rows)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if(i >= m)
                return;
            const int t = tier[i];
            if(t == TIER_NONE)
                return;
            // one atomic per (warp, tier)
            const unsigned peers = __match_any_sync(__activemask(), t);
            const int      lead  = __ffs(peers) - 1;
            const int      rank  = __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
            int            base  = 0;
            if((threadIdx.x & 31) == lead)
                base = atomicAdd(&cursor[t], __popc(peers));
            base = __shfl_sync(peers, base, lead);
            rows[start[t] + base + rank] = (int)i;
        }

        template <int TPR>
        __device__ __forceinline__ void group_sync()
        {
            if(TPR == 32)
                __syncwarp();
            else
                __syncthreads();
        }

        // One row of C per group of TPR threads.  NUMERIC = false: count distinct columns.  NUMERIC = true: accumulate
        // values and compact the table into the row's slice of C (unsorted).  GLOBAL: the tables live in global memory
        // (gkeys / gvals, per-row offset and size), otherwise in shared memory (TABLE slots per group).
        template <typename T, int TPR, int TABLE, bool NUMERIC, bool GLOBAL>
        __global__ void __launch_bounds__(TPR == 32 ? THREADS_W : TPR)
            spgemm_row_kernel(const int *__restrict__ rows,
                              int nrows,
                              const int *__restrict__ rpA,
                              const int *__restrict__ colA,
                              const T *__restrict__ valA,
                              const int *__restrict__ rpB,
                              const int *__restrict__ colB,
                              const T *__restrict__ valB,
                              int conjA,
                              int conjB,
                              int *__restrict__ nnz_row,     // symbolic: out
                              const int *__restrict__ rpC,   // numeric: in
                              int *__restrict__ colC,
                              T *__restrict__ valC,
                              int *__restrict__ gkeys,
                              T *__restrict__ gvals,
                              const long long *__restrict__ goff,
                              const int *__restrict__ gsize,
                              int *__restrict__ err)
        {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            constexpr int GROUPS = (TPR == 32 ? THREADS_W : TPR) / TPR;
            __shared__ int counter[GROUPS];
            const int gl = threadIdx.x / TPR, lane = threadIdx.x % TPR;
            const int gid = blockIdx.x * GROUPS + gl;
            const bool live = gid < nrows;

            int     *keys;
            T       *vals = nullptr;
            unsigned mask;
            int      size;
            if(GLOBAL)
            {
                size = live ? gsize[gid] : 0;
                keys = gkeys + (live ? goff[gid] : 0);
                if(NUMERIC)
                    vals = gvals + (live ? goff[gid] : 0);
                mask = (unsigned)size - 1u; // tables arrive cleared (memset by the host)
            }
            else
            {
                size = TABLE;
                keys = reinterpret_cast<int *>(smem_raw) + gl * TABLE;
                if(NUMERIC)
                    vals = reinterpret_cast<T *>(smem_raw + (size_t)GROUPS * TABLE * sizeof(int)) + gl * TABLE;
                mask = TABLE - 1;
                for(int s = lane; s < TABLE; s += TPR)
                {
                    keys[s] = -1;
                    if(NUMERIC)
                        vals[s] = vt<T>::zero();
                }
            }
            if(lane == 0)
                counter[gl] = 0;
            group_sync<TPR>();

            const int i   = live ? rows[gid] : 0;
            int       cnt = 0;
            if(live)
            {
                // warp tier: a lane per entry of A's row; CTA tiers: a warp per entry, lanes over B's row
                const int jstep = TPR == 32 ? 32 : TPR / 32;
                const int j0    = TPR == 32 ? lane : lane / 32;
                const int kstep = TPR == 32 ? 1 : 32;
                const int k0    = TPR == 32 ? 0 : lane % 32;
                for(int j = rpA[i] + j0; j < rpA[i + 1]; j += jstep)
                {
                    const int c = colA[j];
                    T         a = vt<T>::zero();
                    if(NUMERIC)
                    {
                        a = valA[j];
                        if(conjA)
                            a = cj(a);
                    }
                    for(int k = rpB[c] + k0; k < rpB[c + 1]; k += kstep)
                    {
                        int            fresh;
                        const unsigned h = insert_key(keys, mask, colB[k], &fresh);
                        cnt += fresh;
                        if(NUMERIC)
                        {
                            T b = valB[k];
                            if(conjB)
                                b = cj(b);
                            atomic_add_val(&vals[h], mul(a, b));
                        }
                    }
                }
            }
            if(!NUMERIC)
            {
#pragma unroll
                for(int o = 16; o > 0; o >>= 1)
                    cnt += __shfl_down_sync(0xffffffffu, cnt, o);
                if((lane & 31) == 0 && cnt)
                    atomicAdd(&counter[gl], cnt);
                group_sync<TPR>();
                if(lane == 0 && live)
                    nnz_row[i] = counter[gl];
                return;
            }
            if(GLOBAL)
            {
                // global tables: compact in slot order; these (few, long) rows are ordered afterwards by a segmented sort
                __threadfence();
                group_sync<TPR>();
                if(live)
                {
                    const int base = rpC[i], len = rpC[i + 1] - base;
                    for(int s = lane; s < size; s += TPR)
                    {
                        const int key = keys[s];
                        if(key != -1)
                        {
                            const int p = atomicAdd(&counter[gl], 1);
                            if(p < len)
                            {
                                colC[base + p] = key;
                                valC[base + p] = vals[s];
                            }
                        }
                    }
                    group_sync<TPR>();
                    if(lane == 0 && counter[gl] != len)
                        atomicExch(err, 1); // the pattern changed between the two stages (csr2m.cpp:521-522)
                }
                else
                    group_sync<TPR>();
                return;
            }
            // shared-memory tables: compact (key, slot) pairs, then write the row in ascending column order
            int            *ckey  = reinterpret_cast<int *>(smem_raw + (size_t)GROUPS * TABLE * (sizeof(int) + sizeof(T))) + gl * TABLE;
            unsigned short *cslot = reinterpret_cast<unsigned short *>(smem_raw + (size_t)GROUPS * TABLE * (2 * sizeof(int) + sizeof(T)))
                                    + gl * TABLE;
            group_sync<TPR>();
            for(int s = lane; s < TABLE; s += TPR)
            {
                const int key = keys[s];
                if(key != -1)
                {
                    const int p = atomicAdd(&counter[gl], 1);
                    ckey[p]     = key;
                    cslot[p]    = (unsigned short)s;
                }
            }
            group_sync<TPR>();
            const int L    = counter[gl];
            const int base = live ? rpC[i] : 0;
            const int len  = live ? rpC[i + 1] - base : 0;
            if(live && lane == 0 && L != len)
                atomicExch(err, 1); // the pattern changed between the two stages (csr2m.cpp:521-522)
            if(L != len)
                return; // uniform over the group
            if(L <= RANK_LIMIT)
            {
                // distinct keys: the rank of a key is the number of smaller ones
                for(int e = lane; e < L; e += TPR)
                {
                    const int k    = ckey[e];
                    int       rank = 0;
                    for(int f = 0; f < L; ++f)
                        rank += ckey[f] < k ? 1 : 0;
                    colC[base + rank] = k;
                    valC[base + rank] = vals[cslot[e]];
                }
            }
            else
            {
                // bitonic sort of the pairs, padded with +inf keys to a power of two (<= TABLE)
                int P = 1;
                while(P < L)
                    P <<= 1;
                for(int e = L + lane; e < P; e += TPR)
                    ckey[e] = 0x7fffffff;
                group_sync<TPR>();
                for(int k2 = 2; k2 <= P; k2 <<= 1)
                    for(int j = k2 >> 1; j > 0; j >>= 1)
                    {
                        for(int e = lane; e < P; e += TPR)
                        {
                            const int x = e ^ j;
                            if(x > e)
                            {
                                const int  ka = ckey[e], kb = ckey[x];
                                const bool up = (e & k2) == 0;
                                if((ka > kb) == up)
                                {
                                    ckey[e]                 = kb;
                                    ckey[x]                 = ka;
                                    const unsigned short sa = cslot[e];
                                    cslot[e]                = cslot[x];
                                    cslot[x]                = sa;
                                }
                            }
                        }
                        group_sync<TPR>();
                    }
                for(int e = lane; e < L; e += TPR)
                {
                    colC[base + e] = ckey[e];
                    valC[base + e] = vals[cslot[e]];
                }
            }
        }

        __global__ void sum64_kernel(int m, const int *__restrict__ v, unsigned long long *out)
        {
            unsigned long long s = 0;
            for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
                s += (unsigned)v[i];
#pragma unroll
            for(int o = 16; o > 0; o >>= 1)
                s += __shfl_down_sync(0xffffffffu, s, o);
            if((threadIdx.x & 31) == 0 && s)
                atomicAdd(out, s);
        }

        __global__ void gather_int_kernel(int n, const int *__restrict__ idx, const int *__restrict__ src, int *__restrict__ dst)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if(i < n)
                dst[i] = src[idx[i]];
        }

        __global__ void row_len_kernel(int m, const int *__restrict__ rp, int *__restrict__ out)
        {
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if(i < m)
                out[i] = rp[i + 1] - rp[i];
        }

        __global__ void iota_int_kernel(long long n, int *out)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < n; i += (long long)gridDim.x * blockDim.x)
                out[i] = (int)i;
        }

        template <typename T>
        __global__ void permute_vals_kernel(long long n, const int *__restrict__ perm, const T *__restrict__ src, T *__restrict__ dst)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < n; i += (long long)gridDim.x * blockDim.x)
                dst[i] = src[perm[i]];
        }

        inline unsigned blocks_for(long long n, int tpb)
        {
            long long b = (n + tpb - 1) / tpb;
            if(b > 148LL * 64)
                b = 148LL * 64;
            return (unsigned)(b < 1 ? 1 : b);
        }

        inline int pow2_at_least(long long v)
        {
            long long p = 64;
            while(p < v && p < (1LL << 30))
                p <<= 1;
            return (int)p;
        }

        struct tier_lists
        {
            dev_buf ub, tier, rows;
            int     count[N_TIERS] = {0, 0, 0, 0, 0};
            int     start[N_TIERS] = {0, 0, 0, 0, 0};
        };

        aoclsparse_status make_tiers(const dev_csr &A, const dev_csr &B, int ub_large, tier_lists &L, cudaStream_t st)
        {
            const int m = A.m;
            dev_buf   hist;
            B200_TRY(L.ub.alloc(sizeof(int) * (size_t)(m > 0 ? m : 1)));
            B200_TRY(L.tier.alloc((size_t)(m > 0 ? m : 1)));
            B200_TRY(L.rows.alloc(sizeof(int) * (size_t)(m > 0 ? m : 1)));
            B200_TRY(hist.alloc(sizeof(int) * 3 * N_TIERS));
            B200_CUDA(cudaMemsetAsync(hist.p, 0, sizeof(int) * 3 * N_TIERS, st));
            const unsigned grid = (unsigned)(((long long)m + 255) / 256);
            if(m > 0)
            {
                spgemm_bound_kernel<<<grid, 256, 0, st>>>(
                    m, A.row_ptr.as<int>(), A.col_idx.as<int>(), B.row_ptr.as<int>(), ub_large, L.ub.as<int>(),
                    L.tier.as<unsigned char>(), hist.as<int>());
                B200_LAUNCHED();
            }
            B200_CUDA(cudaMemcpyAsync(L.count, hist.p, sizeof(int) * N_TIERS, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            int run = 0;
            for(int t = 1; t < N_TIERS; ++t)
            {
                L.start[t] = run;
                run += L.count[t];
            }
            B200_CUDA(cudaMemcpyAsync(hist.as<int>() + N_TIERS, L.start, sizeof(int) * N_TIERS, cudaMemcpyHostToDevice, st));
            if(m > 0)
            {
                spgemm_place_kernel<<<grid, 256, 0, st>>>(
                    m, L.tier.as<unsigned char>(), hist.as<int>() + N_TIERS, hist.as<int>() + 2 * N_TIERS, L.rows.as<int>());
                B200_LAUNCHED();
            }
            B200_CUDA(cudaStreamSynchronize(st)); // L.start (host) was the source of an async copy
            return aoclsparse_status_success;
        }

        template <typename T, int TPR, int TABLE, bool NUMERIC>
        aoclsparse_status launch_smem_tier(const int *rows, int nrows, const dev_csr &A, const dev_csr &B, int conjA, int conjB,
                                           int *nnz_row, const int *rpC, int *colC, T *valC, int *err, cudaStream_t st)
        {
            if(nrows <= 0)
                return aoclsparse_status_success;
            constexpr int THREADS = TPR == 32 ? THREADS_W : TPR;
            constexpr int GROUPS  = THREADS / TPR;
            // keys [+ values + compacted keys + compacted slot ids]
            const size_t  smem    = (size_t)GROUPS * TABLE
                                 * (sizeof(int) + (NUMERIC ? sizeof(T) + sizeof(int) + sizeof(unsigned short) : 0));
            auto          kern    = spgemm_row_kernel<T, TPR, TABLE, NUMERIC, false>;
            if(smem > 48 * 1024)
                B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)((nrows + GROUPS - 1) / GROUPS), THREADS, smem, st>>>(rows,
                                                                                  nrows,
                                                                                  A.row_ptr.as<int>(),
                                                                                  A.col_idx.as<int>(),
                                                                                  A.val.as<T>(),
                                                                                  B.row_ptr.as<int>(),
                                                                                  B.col_idx.as<int>(),
                                                                                  B.val.as<T>(),
                                                                                  conjA,
                                                                                  conjB,
                                                                                  nnz_row,
                                                                                  rpC,
                                                                                  colC,
                                                                                  valC,
                                                                                  nullptr,
                                                                                  nullptr,
                                                                                  nullptr,
                                                                                  nullptr,
                                                                                  err);
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }

        // rows whose table does not fit shared memory: global tables, processed in batches under a scratch budget
        template <typename T, bool NUMERIC>
        aoclsparse_status launch_global_tier(const int *rows, int nrows, const int *bound /* per listed row */, long long n_cols,
                                             const dev_csr &A, const dev_csr &B, int conjA, int conjB, int *nnz_row,
                                             const int *rpC, int *colC, T *valC, int *err, cudaStream_t st)
        {
            if(nrows <= 0)
                return aoclsparse_status_success;
            std::vector<int> hb((size_t)nrows);
            B200_CUDA(cudaMemcpyAsync(hb.data(), bound, sizeof(int) * (size_t)nrows, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            const long long budget = 64LL << 20; // table slots per batch
            int             b0     = 0;
            while(b0 < nrows)
            {
                std::vector<long long> off;
                std::vector<int>       size;
                long long              total = 0;
                int                    b1    = b0;
                while(b1 < nrows)
                {
                    const long long want = 2 * std::min<long long>(hb[(size_t)b1], n_cols);
                    if(want > (1LL << 30))
                        return aoclsparse_status_invalid_size; // a single row of C beyond 2^29 entries
                    const int       sz   = pow2_at_least(want > 0 ? want : 1);
                    if(b1 > b0 && total + sz > budget)
                        break;
                    off.push_back(total);
                    size.push_back(sz);
                    total += sz;
                    ++b1;
                }
                const int nb = b1 - b0;
                dev_buf   gkeys, gvals, goff, gsize;
                B200_TRY(gkeys.alloc(sizeof(int) * (size_t)total));
                if(NUMERIC)
                    B200_TRY(gvals.alloc(sizeof(T) * (size_t)total));
                B200_TRY(goff.alloc(sizeof(long long) * (size_t)nb));
                B200_TRY(gsize.alloc(sizeof(int) * (size_t)nb));
                B200_CUDA(cudaMemsetAsync(gkeys.p, 0xff, sizeof(int) * (size_t)total, st));
                if(NUMERIC)
                    B200_CUDA(cudaMemsetAsync(gvals.p, 0, sizeof(T) * (size_t)total, st));
                B200_CUDA(cudaMemcpyAsync(goff.p, off.data(), sizeof(long long) * (size_t)nb, cudaMemcpyHostToDevice, st));
                B200_CUDA(cudaMemcpyAsync(gsize.p, size.data(), sizeof(int) * (size_t)nb, cudaMemcpyHostToDevice, st));
                spgemm_row_kernel<T, THREADS_L, 1, NUMERIC, true><<<(unsigned)nb, THREADS_L, 0, st>>>(rows + b0,
                                                                                                     nb,
                                                                                                     A.row_ptr.as<int>(),
                                                                                                     A.col_idx.as<int>(),
                                                                                                     A.val.as<T>(),
                                                                                                     B.row_ptr.as<int>(),
                                                                                                     B.col_idx.as<int>(),
                                                                                                     B.val.as<T>(),
                                                                                                     conjA,
                                                                                                     conjB,
                                                                                                     nnz_row,
                                                                                                     rpC,
                                                                                                     colC,
                                                                                                     valC,
                                                                                                     gkeys.as<int>(),
                                                                                                     gvals.as<T>(),
                                                                                                     goff.as<long long>(),
                                                                                                     gsize.as<int>(),
                                                                                                     err);
                B200_LAUNCHED();
                B200_CUDA(cudaStreamSynchronize(st)); // off / size / scratch go out of scope
                b0 = b1;
            }
            return aoclsparse_status_success;
        }

        template <typename T, bool NUMERIC>
        aoclsparse_status run_tiers(const dev_csr &A, const dev_csr &B, int conjA, int conjB, tier_lists &L, int *nnz_row,
                                    const int *rpC, int *colC, T *valC, int *err, cudaStream_t st)
        {
            const int *rows = L.rows.as<int>();
            B200_TRY((launch_smem_tier<T, 32, TAB_WARP, NUMERIC>(
                rows + L.start[TIER_WARP], L.count[TIER_WARP], A, B, conjA, conjB, nnz_row, rpC, colC, valC, err, st)));
            B200_TRY((launch_smem_tier<T, THREADS_S, TAB_S, NUMERIC>(
                rows + L.start[TIER_CTA_S], L.count[TIER_CTA_S], A, B, conjA, conjB, nnz_row, rpC, colC, valC, err, st)));
            B200_TRY((launch_smem_tier<T, THREADS_L, large_tier<T>::TAB, NUMERIC>(
                rows + L.start[TIER_CTA_L], L.count[TIER_CTA_L], A, B, conjA, conjB, nnz_row, rpC, colC, valC, err, st)));
            const int ng = L.count[TIER_GLOBAL];
            if(ng > 0)
            {
                // table size from the product count (symbolic) or from the exact row length (numeric)
                dev_buf bound;
                B200_TRY(bound.alloc(sizeof(int) * (size_t)ng));
                gather_int_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(
                    ng, rows + L.start[TIER_GLOBAL], NUMERIC ? nnz_row : L.ub.as<int>(), bound.as<int>());
                B200_LAUNCHED();
                B200_TRY((launch_global_tier<T, NUMERIC>(rows + L.start[TIER_GLOBAL], ng, bound.as<int>(), B.n, A, B, conjA,
                                                         conjB, nnz_row, rpC, colC, valC, err, st)));
            }
            return aoclsparse_status_success;
        }

        // pass 1: P.row_ptr (m+1, exclusive scan of the row lengths), P.nnz; col / val allocated, not filled
        template <typename T>
        aoclsparse_status symbolic(const dev_csr &A, const dev_csr &B, dev_csr &P, tier_lists &L, cudaStream_t st)
        {
            const int m = A.m;
            P.m         = m;
            P.n         = B.n;
            phase_trace tr(st);
            B200_TRY(make_tiers(A, B, large_tier<T>::UB, L, st));
            tr.mark("count: bounds + tiers");
            dev_buf nnz_row, total, temp;
            B200_TRY(nnz_row.alloc(sizeof(int) * ((size_t)m + 1)));
            B200_TRY(total.alloc(sizeof(unsigned long long)));
            B200_CUDA(cudaMemsetAsync(nnz_row.p, 0, sizeof(int) * ((size_t)m + 1), st));
            B200_CUDA(cudaMemsetAsync(total.p, 0, sizeof(unsigned long long), st));
            B200_TRY((run_tiers<T, false>(A, B, 0, 0, L, nnz_row.as<int>(), nullptr, nullptr, nullptr, nullptr, st)));
            tr.mark("count: hash kernels");
            sum64_kernel<<<blocks_for(m, 256), 256, 0, st>>>(m, nnz_row.as<int>(), total.as<unsigned long long>());
            B200_LAUNCHED();
            unsigned long long h_total = 0;
            B200_CUDA(cudaMemcpyAsync(&h_total, total.p, sizeof(h_total), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            if(h_total > 0x7fffffffULL)
                return aoclsparse_status_invalid_size; // csr2m.cpp:236-241
            B200_TRY(P.row_ptr.alloc(sizeof(int) * ((size_t)m + 1)));
            size_t temp_bytes = 0;
            B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, nnz_row.as<int>(), P.row_ptr.as<int>(), m + 1, st));
            B200_TRY(temp.alloc(temp_bytes));
            B200_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, temp_bytes, nnz_row.as<int>(), P.row_ptr.as<int>(), m + 1, st));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            P.nnz = (aoclsparse_int)h_total;
            B200_TRY(P.col_idx.alloc(sizeof(int) * (size_t)P.nnz));
            B200_TRY(P.val.alloc(sizeof(T) * (size_t)P.nnz));
            B200_CUDA(cudaStreamSynchronize(st));
            tr.mark("count: scan + allocate C");
            return aoclsparse_status_success;
        }

        // segmented sort of (col, val) by column inside every row [rp[i], rp[i+1])
        template <typename T>
        aoclsparse_status sort_rows(int m, long long nnz, const int *rp, dev_buf &col, dev_buf &val, cudaStream_t st)
        {
            if(nnz <= 0 || m <= 0)
                return aoclsparse_status_success;
            dev_buf col_out, val_out, idx_in, idx_out, temp;
            B200_TRY(col_out.alloc(sizeof(int) * (size_t)nnz));
            B200_TRY(val_out.alloc(sizeof(T) * (size_t)nnz));
            B200_TRY(idx_in.alloc(sizeof(int) * (size_t)nnz));
            B200_TRY(idx_out.alloc(sizeof(int) * (size_t)nnz));
            iota_int_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, idx_in.as<int>());
            B200_LAUNCHED();
            size_t temp_bytes = 0;
            B200_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr,
                                                                temp_bytes,
                                                                col.as<int>(),
                                                                col_out.as<int>(),
                                                                idx_in.as<int>(),
                                                                idx_out.as<int>(),
                                                                (int)nnz,
                                                                m,
                                                                rp,
                                                                rp + 1,
                                                                st));
            B200_TRY(temp.alloc(temp_bytes));
            B200_CUDA(cub::DeviceSegmentedSort::StableSortPairs(temp.p,
                                                                temp_bytes,
                                                                col.as<int>(),
                                                                col_out.as<int>(),
                                                                idx_in.as<int>(),
                                                                idx_out.as<int>(),
                                                                (int)nnz,
                                                                m,
                                                                rp,
                                                                rp + 1,
                                                                st));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            permute_vals_kernel<T><<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, idx_out.as<int>(), val.as<T>(), val_out.as<T>());
            B200_LAUNCHED();
            B200_CUDA(cudaStreamSynchronize(st));
            col = std::move(col_out);
            val = std::move(val_out);
            return aoclsparse_status_success;
        }

        __global__ void segment_bounds_kernel(int n, const int *__restrict__ rows, const int *__restrict__ rp,
                                              int *__restrict__ beg, int *__restrict__ end)
        {
            const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if(g < n)
            {
                beg[g] = rp[rows[g]];
                end[g] = rp[rows[g] + 1];
            }
        }

        // one CTA per listed row: phase 0 gathers the permuted values into vtmp, phase 1 copies keys / values back
        template <typename T>
        __global__ void segment_apply_kernel(const int *__restrict__ beg, const int *__restrict__ end, int phase,
                                             const int *__restrict__ col_sorted, const int *__restrict__ perm,
                                             int *__restrict__ col, T *__restrict__ val, T *__restrict__ vtmp)
        {
            const int b = beg[blockIdx.x], e = end[blockIdx.x];
            for(int p = b + threadIdx.x; p < e; p += blockDim.x)
            {
                if(phase == 0)
                    vtmp[p] = val[perm[p]];
                else
                {
                    col[p] = col_sorted[p];
                    val[p] = vtmp[p];
                }
            }
        }

        // ascending column order for the listed rows only (the rows of the global-table tier)
        template <typename T>
        aoclsparse_status sort_row_subset(const int *rows, int n_rows, long long nnz, const int *rp, dev_buf &col, dev_buf &val,
                                          cudaStream_t st)
        {
            if(n_rows <= 0 || nnz <= 0)
                return aoclsparse_status_success;
            dev_buf beg, end, col_out, idx_in, idx_out, vtmp, temp;
            B200_TRY(beg.alloc(sizeof(int) * (size_t)n_rows));
            B200_TRY(end.alloc(sizeof(int) * (size_t)n_rows));
            B200_TRY(col_out.alloc(sizeof(int) * (size_t)nnz));
            B200_TRY(idx_in.alloc(sizeof(int) * (size_t)nnz));
            B200_TRY(idx_out.alloc(sizeof(int) * (size_t)nnz));
            B200_TRY(vtmp.alloc(sizeof(T) * (size_t)nnz));
            segment_bounds_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(n_rows, rows, rp, beg.as<int>(), end.as<int>());
            B200_LAUNCHED();
            iota_int_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(nnz, idx_in.as<int>());
            B200_LAUNCHED();
            size_t temp_bytes = 0;
            B200_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr, temp_bytes, col.as<int>(), col_out.as<int>(),
                                                                idx_in.as<int>(), idx_out.as<int>(), (int)nnz, n_rows,
                                                                beg.as<int>(), end.as<int>(), st));
            B200_TRY(temp.alloc(temp_bytes));
            B200_CUDA(cub::DeviceSegmentedSort::StableSortPairs(temp.p, temp_bytes, col.as<int>(), col_out.as<int>(),
                                                                idx_in.as<int>(), idx_out.as<int>(), (int)nnz, n_rows,
                                                                beg.as<int>(), end.as<int>(), st));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            for(int phase = 0; phase < 2; ++phase)
            {
                segment_apply_kernel<T><<<(unsigned)n_rows, 256, 0, st>>>(beg.as<int>(), end.as<int>(), phase, col_out.as<int>(),
                                                                         idx_out.as<int>(), col.as<int>(), val.as<T>(), vtmp.as<T>());
                B200_LAUNCHED();
            }
            B200_CUDA(cudaStreamSynchronize(st));
            return aoclsparse_status_success;
        }

        // pass 2: fills P.col_idx / P.val for the row_ptr of pass 1; rows sorted by column
        template <typename T>
        aoclsparse_status numeric(const dev_csr &A, const dev_csr &B, int conjA, int conjB, dev_csr &P, tier_lists *have,
                                  cudaStream_t st)
        {
            const int m = A.m;
            if(P.nnz == 0 || m == 0)
                return aoclsparse_status_success;
            phase_trace tr(st);
            tier_lists  own;
            if(!have) // the finalize stage on its own: the tiers depend on the patterns only, recompute them
            {
                B200_TRY(make_tiers(A, B, large_tier<T>::UB, own, st));
                tr.mark("fill: bounds + tiers");
            }
            tier_lists &L = have ? *have : own;
            dev_buf nnz_row, err;
            B200_TRY(err.alloc(sizeof(int)));
            B200_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), st));
            B200_TRY(nnz_row.alloc(sizeof(int) * (size_t)m));
            row_len_kernel<<<(unsigned)(((long long)m + 255) / 256), 256, 0, st>>>(m, P.row_ptr.as<int>(), nnz_row.as<int>());
            B200_LAUNCHED();
            B200_TRY((run_tiers<T, true>(A, B, conjA, conjB, L, nnz_row.as<int>(), P.row_ptr.as<int>(), P.col_idx.as<int>(),
                                         P.val.as<T>(), err.as<int>(), st)));
            int h_err = 0;
            B200_CUDA(cudaMemcpyAsync(&h_err, err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            if(h_err)
                return aoclsparse_status_internal_error;
            tr.mark("fill: hash kernels");
            // the shared-memory tiers wrote their rows in order; only the global-table rows are still in slot order
            const aoclsparse_status ss = sort_row_subset<T>(L.rows.as<int>() + L.start[TIER_GLOBAL], L.count[TIER_GLOBAL], P.nnz,
                                                            P.row_ptr.as<int>(), P.col_idx, P.val, st);
            tr.mark("fill: order global-tier rows");
            return ss;
        }

        // a general handle around device arrays that are already in place (base 0)
        aoclsparse_status finish_handle(aoclsparse_matrix C, bool classify, cudaStream_t st)
        {
            dev_csr &M = *C->mats[0];
            C->nnz     = M.nnz;
            if(classify)
            {
                check_result cr;
                B200_TRY(check_csr_device(M.m, M.n, M.nnz, 0, M.row_ptr.as<aoclsparse_int>(), M.col_idx.as<aoclsparse_int>(), cr, st));
                if(cr.status != aoclsparse_status_success)
                    return aoclsparse_status_internal_error;
                C->sort        = (aoclsparse_matrix_sort)cr.sort;
                C->fulldiag    = cr.fulldiag != 0;
                C->min_col     = cr.min_col;
                C->max_col     = cr.max_col;
                C->max_row_nnz = cr.max_row_nnz;
            }
            M.plan.valid = false;
            return aoclsparse_status_success;
        }

        aoclsparse_status new_result_handle(aoclsparse_matrix *C, int val_type, aoclsparse_int m, aoclsparse_int n)
        {
            _aoclsparse_matrix *H = new(std::nothrow) _aoclsparse_matrix;
            dev_csr            *M = new(std::nothrow) dev_csr;
            if(!H || !M)
            {
                delete H;
                delete M;
                return aoclsparse_status_memory_error;
            }
            H->mats.push_back(M);
            H->m = M->m = m;
            H->n = M->n = n;
            H->nnz      = 0;
            H->base     = aoclsparse_index_base_zero;
            H->val_type = (aoclsparse_matrix_data_type)val_type;
            H->sort     = aoclsparse_unknown_sort;
            cudaGetDevice(&H->device);
            *C = H;
            return aoclsparse_status_success;
        }

        inline bool valid_op(aoclsparse_operation op)
        {
            return op == aoclsparse_operation_none || op == aoclsparse_operation_transpose
                   || op == aoclsparse_operation_conjugate_transpose;
        }

        // the stored matrix of X, or a transposed temporary of it, as the left / right factor
        struct factor
        {
            const dev_csr *M = nullptr;
            dev_csr        temp;
            int            conj = 0;
        };

        template <typename T>
        aoclsparse_status sp2m_t(aoclsparse_operation       opA,
                                 const aoclsparse_mat_descr descrA,
                                 const aoclsparse_matrix    A,
                                 aoclsparse_operation       opB,
                                 const aoclsparse_mat_descr descrB,
                                 const aoclsparse_matrix    B,
                                 aoclsparse_request         request,
                                 aoclsparse_matrix         *C)
        {
            // ---- validation in the reference's order (csr2m.cpp:603-700)
            if(descrA == nullptr || descrB == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(A == nullptr || B == nullptr || C == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(request != aoclsparse_stage_finalize)
                *C = nullptr;
            if(A->input_format != aoclsparse_csr_mat || B->input_format != aoclsparse_csr_mat)
                return aoclsparse_status_not_implemented;
            if(A->val_type != vt<T>::data_type || B->val_type != vt<T>::data_type)
                return aoclsparse_status_wrong_type;
            if(A->mats.empty() || !A->mats[0] || B->mats.empty() || !B->mats[0])
                return aoclsparse_status_invalid_pointer;
            if((descrA->base != aoclsparse_index_base_zero && descrA->base != aoclsparse_index_base_one)
               || (descrB->base != aoclsparse_index_base_zero && descrB->base != aoclsparse_index_base_one))
                return aoclsparse_status_invalid_value;
            if(A->base != descrA->base || B->base != descrB->base)
                return aoclsparse_status_invalid_value;
            if(descrA->type != aoclsparse_matrix_type_general || descrB->type != aoclsparse_matrix_type_general)
                return aoclsparse_status_not_implemented;
            if(!valid_op(opA) || !valid_op(opB))
                return aoclsparse_status_invalid_value;
            const bool           tA  = opA != aoclsparse_operation_none, tB = opB != aoclsparse_operation_none;
            const aoclsparse_int m_a = tA ? A->n : A->m, n_a = tA ? A->m : A->n;
            const aoclsparse_int m_b = tB ? B->n : B->m, n_b = tB ? B->m : B->n;
            if(n_a != m_b)
                return aoclsparse_status_invalid_size;
            cudaStream_t st = current_stream();
            if(m_a == 0 || n_a == 0 || n_b == 0 || A->nnz == 0 || B->nnz == 0)
            {
                if(*C == nullptr)
                {
                    B200_TRY(new_result_handle(C, vt<T>::data_type, m_a, n_b));
                    dev_csr &M = *(*C)->mats[0];
                    aoclsparse_status s = M.row_ptr.alloc(sizeof(int) * ((size_t)m_a + 1));
                    if(s == aoclsparse_status_success)
                        s = M.col_idx.alloc(0);
                    if(s == aoclsparse_status_success)
                        s = M.val.alloc(0);
                    if(s != aoclsparse_status_success)
                    {
                        aoclsparse_destroy(C);
                        return s;
                    }
                    B200_CUDA(cudaMemsetAsync(M.row_ptr.p, 0, sizeof(int) * ((size_t)m_a + 1), st));
                    B200_CUDA(cudaStreamSynchronize(st));
                    (*C)->sort = aoclsparse_fully_sorted;
                }
                return aoclsparse_status_success;
            }
            if(request != aoclsparse_stage_nnz_count && request != aoclsparse_stage_finalize
               && request != aoclsparse_stage_full_computation)
                return aoclsparse_status_invalid_value;

            // ---- which stored matrix, transposed or not, conjugated or not (csr2m.cpp:713-810).  Stored doid: gn for a
            // CSR handle, gt for a CSC one; general ids are [transpose:1][conjugate:1], so the remaining work is the xor
            const bool cplx  = vt<T>::is_complex;
            const int  reqA  = get_doid(cplx, aoclsparse_matrix_type_general, descrA->fill_mode, opA);
            const int  reqB  = get_doid(cplx, aoclsparse_matrix_type_general, descrB->fill_mode, opB);
            const int  effA  = (A->is_csc ? DOID_GT : DOID_GN) ^ reqA;
            const int  effB  = (B->is_csc ? DOID_GT : DOID_GN) ^ reqB;
            const bool trA = (effA & 2) != 0, trB = (effB & 2) != 0;
            const int  opflag = (trA ? 1 : 0) | (trB ? 2 : 0);

            std::shared_lock<std::shared_mutex> la(A->guard);
            std::shared_lock<std::shared_mutex> lb;
            if(B != A)
                lb = std::shared_lock<std::shared_mutex>(B->guard);
            factor L, R;
            if(opflag == 3)
            {
                // op(A) op(B) = (B_s A_s)^T on the stored matrices: multiply in swapped roles, transpose the product
                L.M    = B->mats[0];
                L.conj = effB & 1;
                R.M    = A->mats[0];
                R.conj = effA & 1;
            }
            else
            {
                L.M    = A->mats[0];
                L.conj = effA & 1;
                R.M    = B->mats[0];
                R.conj = effB & 1;
                if(trA)
                {
                    B200_TRY(transpose_csr(*A->mats[0], A->val_type, false, L.temp, st));
                    L.M = &L.temp;
                }
                if(trB)
                {
                    B200_TRY(transpose_csr(*B->mats[0], B->val_type, false, R.temp, st));
                    R.M = &R.temp;
                }
            }

            // ---- stages
            tier_lists tiers;
            bool       have_tiers = false;
            if(request == aoclsparse_stage_finalize)
            {
                if(*C == nullptr || (*C)->mats.empty() || !(*C)->mats[0])
                    return aoclsparse_status_invalid_pointer;
                if((*C)->val_type != vt<T>::data_type)
                    return aoclsparse_status_wrong_type;
            }
            else
            {
                B200_TRY(new_result_handle(C, vt<T>::data_type, m_a, n_b));
                dev_csr          *P = opflag == 3 ? new(std::nothrow) dev_csr : (*C)->mats[0];
                aoclsparse_status s = P ? symbolic<T>(*L.M, *R.M, *P, tiers, st) : aoclsparse_status_memory_error;
                have_tiers          = s == aoclsparse_status_success;
                if(s == aoclsparse_status_success && opflag == 3)
                {
                    // the product B A is kept in the handle (its own field: value updates of C drop the derived copies
                    // in mats[1..], not this) until and across finalize calls; the result's pattern is its transpose,
                    // so row_ptr / col_idx are already the final ones after the count stage (values: finalize)
                    (*C)->sp2m_product.reset(P);
                    P->doid = DOID_GT;
                    dev_csr Tr;
                    s = transpose_csr(*P, (*C)->val_type, false, Tr, st);
                    if(s == aoclsparse_status_success)
                    {
                        dev_csr &M = *(*C)->mats[0];
                        M.row_ptr  = std::move(Tr.row_ptr);
                        M.col_idx  = std::move(Tr.col_idx);
                        M.val      = std::move(Tr.val);
                        M.nnz      = P->nnz;
                    }
                }
                else if(opflag == 3)
                    delete P;
                if(s != aoclsparse_status_success)
                {
                    aoclsparse_destroy(C);
                    return s;
                }
                (*C)->nnz = (*C)->mats[0]->nnz;
                if(request == aoclsparse_stage_nnz_count)
                {
                    B200_CUDA(cudaStreamSynchronize(st));
                    return aoclsparse_status_success;
                }
            }

            aoclsparse_matrix H = *C;
            if(opflag == 3)
            {
                if(!H->sp2m_product)
                    return aoclsparse_status_invalid_pointer;
                dev_csr &P = *H->sp2m_product;
                B200_TRY(numeric<T>(*L.M, *R.M, L.conj, R.conj, P, have_tiers ? &tiers : nullptr, st));
                dev_csr Tr;
                B200_TRY(transpose_csr(P, H->val_type, false, Tr, st));
                dev_csr &M = *H->mats[0];
                M.row_ptr  = std::move(Tr.row_ptr);
                M.col_idx  = std::move(Tr.col_idx);
                M.val      = std::move(Tr.val);
                M.nnz      = P.nnz;
            }
            else
            {
                dev_csr &M = *H->mats[0];
                if(M.m != L.M->m || M.n != R.M->n || !M.row_ptr.p)
                    return aoclsparse_status_invalid_pointer;
                B200_TRY(numeric<T>(*L.M, *R.M, L.conj, R.conj, M, have_tiers ? &tiers : nullptr, st));
            }
            phase_trace tr(st);
            {
                // a repeated finalize is the documented way to refresh the values of C: copies derived from the old
                // values (transposed / expanded / clean, built by earlier mv / csrmm calls on C) must not survive it
                std::unique_lock<std::shared_mutex> wl(H->guard);
                drop_derived_copies(H);
            }
            B200_TRY(finish_handle(H, true, st));
            B200_CUDA(cudaStreamSynchronize(st));
            tr.mark("classify result (check.cu)");
            return aoclsparse_status_success;
        }

        // host mirror of a handle's CSR arrays in the handle's own base (aoclsparse_export_?csr)
        template <typename T>
        aoclsparse_status export_csr_t(const aoclsparse_matrix mat,
                                       aoclsparse_index_base  *base,
                                       aoclsparse_int         *m,
                                       aoclsparse_int         *n,
                                       aoclsparse_int         *nnz,
                                       aoclsparse_int        **row_ptr,
                                       aoclsparse_int        **col_ind,
                                       T                     **val)
        {
            if(!mat || !base || !m || !n || !nnz || !row_ptr || !col_ind || !val)
                return aoclsparse_status_invalid_pointer;
            if(mat->val_type != vt<T>::data_type)
                return aoclsparse_status_wrong_type;
            if(mat->mats.empty() || !mat->mats[0])
                return aoclsparse_status_invalid_pointer;
            if(mat->is_csc)
                return aoclsparse_status_invalid_value; // no CSR representation held (auxiliary.cpp:1326-1341)
            cudaStream_t                        st = current_stream();
            std::unique_lock<std::shared_mutex> wl(mat->guard);
            const dev_csr                      &M = *mat->mats[0];
            const size_t                        cnt = (size_t)M.nnz;
            mat->host_row_ptr.resize((size_t)M.m + 1);
            mat->host_col.resize(cnt ? cnt : 1);
            mat->host_val.resize((cnt ? cnt : 1) * sizeof(T));
            B200_CUDA(cudaMemcpyAsync(
                mat->host_row_ptr.data(), M.row_ptr.p, sizeof(int) * ((size_t)M.m + 1), cudaMemcpyDeviceToHost, st));
            if(cnt && M.col_idx.p && M.val.p)
            {
                B200_CUDA(cudaMemcpyAsync(mat->host_col.data(), M.col_idx.p, sizeof(int) * cnt, cudaMemcpyDeviceToHost, st));
                B200_CUDA(cudaMemcpyAsync(mat->host_val.data(), M.val.p, sizeof(T) * cnt, cudaMemcpyDeviceToHost, st));
            }
            B200_CUDA(cudaStreamSynchronize(st));
            if(mat->base == aoclsparse_index_base_one) // device copies are kept zero-based
            {
                for(auto &v : mat->host_row_ptr)
                    v += 1;
                for(size_t i = 0; i < cnt; ++i)
                    mat->host_col[i] += 1;
            }
            *base    = mat->base;
            *m       = mat->m;
            *n       = mat->n;
            *nnz     = mat->host_row_ptr[(size_t)M.m] - (aoclsparse_int)mat->base;
            *row_ptr = mat->host_row_ptr.data();
            *col_ind = mat->host_col.data();
            *val     = reinterpret_cast<T *>(mat->host_val.data());
            return aoclsparse_status_success;
        }
    }
}

using namespace b200;

extern "C" {
aoclsparse_status aoclsparse_sp2m(aoclsparse_operation       opA,
                                  const aoclsparse_mat_descr descrA,
                                  const aoclsparse_matrix    A,
                                  aoclsparse_operation       opB,
                                  const aoclsparse_mat_descr descrB,
                                  const aoclsparse_matrix    B,
                                  const aoclsparse_request   request,
                                  aoclsparse_matrix         *C)
{
    // aoclsparse_sp2m.cpp: pointer checks, then dispatch on the value type of A (a differing B is wrong_type)
    if(A == nullptr || B == nullptr || C == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(A->val_type != B->val_type)
        return aoclsparse_status_wrong_type;
    switch(A->val_type)
    {
    case aoclsparse_smat:
        return sp2m_t<float>(opA, descrA, A, opB, descrB, B, request, C);
    case aoclsparse_dmat:
        return sp2m_t<double>(opA, descrA, A, opB, descrB, B, request, C);
    case aoclsparse_cmat:
        return sp2m_t<float2>(opA, descrA, A, opB, descrB, B, request, C);
    case aoclsparse_zmat:
        return sp2m_t<double2>(opA, descrA, A, opB, descrB, B, request, C);
    default:
        return aoclsparse_status_wrong_type;
    }
}

// aoclsparse_spmm.cpp:27-67: general descriptors in the matrices' own bases, B not transposed, single stage
aoclsparse_status aoclsparse_spmm(aoclsparse_operation opA, const aoclsparse_matrix A, const aoclsparse_matrix B, aoclsparse_matrix *C)
{
    if(A == nullptr || B == nullptr || C == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(A->mats.empty() || !A->mats[0] || B->mats.empty() || !B->mats[0])
        return aoclsparse_status_invalid_pointer;
    _aoclsparse_mat_descr dA, dB;
    dA.type = aoclsparse_matrix_type_general;
    dA.base = A->base;
    dB.type = aoclsparse_matrix_type_general;
    dB.base = B->base;
    if(A->val_type != B->val_type)
        return aoclsparse_status_wrong_type;
    return aoclsparse_sp2m(opA, &dA, A, aoclsparse_operation_none, &dB, B, aoclsparse_stage_full_computation, C);
}

aoclsparse_status aoclsparse_export_scsr(const aoclsparse_matrix mat, aoclsparse_index_base *base, aoclsparse_int *m, aoclsparse_int *n, aoclsparse_int *nnz, aoclsparse_int **row_ptr, aoclsparse_int **col_ind, float **val)
{
    return export_csr_t<float>(mat, base, m, n, nnz, row_ptr, col_ind, val);
}
aoclsparse_status aoclsparse_export_dcsr(const aoclsparse_matrix mat, aoclsparse_index_base *base, aoclsparse_int *m, aoclsparse_int *n, aoclsparse_int *nnz, aoclsparse_int **row_ptr, aoclsparse_int **col_ind, double **val)
{
    return export_csr_t<double>(mat, base, m, n, nnz, row_ptr, col_ind, val);
}
aoclsparse_status aoclsparse_export_ccsr(const aoclsparse_matrix mat, aoclsparse_index_base *base, aoclsparse_int *m, aoclsparse_int *n, aoclsparse_int *nnz, aoclsparse_int **row_ptr, aoclsparse_int **col_ind, aoclsparse_float_complex **val)
{
    return export_csr_t<float2>(mat, base, m, n, nnz, row_ptr, col_ind, reinterpret_cast<float2 **>(val));
}
aoclsparse_status aoclsparse_export_zcsr(const aoclsparse_matrix mat, aoclsparse_index_base *base, aoclsparse_int *m, aoclsparse_int *n, aoclsparse_int *nnz, aoclsparse_int **row_ptr, aoclsparse_int **col_ind, aoclsparse_double_complex **val)
{
    return export_csr_t<double2>(mat, base, m, n, nnz, row_ptr, col_ind, reinterpret_cast<double2 **>(val));
}

// device view of the same arrays (always zero-based), for consumers that stay on the GPU
aoclsparse_status aoclsparse_b200_export_device_csr(const aoclsparse_matrix mat, aoclsparse_int *m, aoclsparse_int *n, aoclsparse_int *nnz, const aoclsparse_int **row_ptr, const aoclsparse_int **col_ind, const void **val)
{
    if(!mat || !m || !n || !nnz || !row_ptr || !col_ind || !val)
        return aoclsparse_status_invalid_pointer;
    if(mat->mats.empty() || !mat->mats[0])
        return aoclsparse_status_invalid_pointer;
    const dev_csr &M = *mat->mats[0];
    *m               = M.m;
    *n               = M.n;
    *nnz             = M.nnz;
    *row_ptr         = M.row_ptr.as<aoclsparse_int>();
    *col_ind         = M.col_idx.as<aoclsparse_int>();
    *val             = M.val.p;
    return aoclsparse_status_success;
}

// aoclsparse_order_mat (auxiliary.cpp:840-878): ascending column indices inside every row (row indices inside every
// column for a CSC handle -- the stored arrays are the rows of the transpose, so it is the same sort)
aoclsparse_status aoclsparse_order_mat(aoclsparse_matrix mat)
{
    if(!mat)
        return aoclsparse_status_invalid_pointer;
    if(mat->m < 0 || mat->n < 0 || mat->nnz < 0)
        return aoclsparse_status_invalid_value;
    if(mat->input_format != aoclsparse_csr_mat)
        return aoclsparse_status_not_implemented;
    if(mat->m == 0 || mat->n == 0 || mat->nnz == 0)
        return aoclsparse_status_success;
    if(mat->mats.empty() || !mat->mats[0])
        return aoclsparse_status_invalid_pointer;
    cudaStream_t                        st = current_stream();
    std::unique_lock<std::shared_mutex> wl(mat->guard);
    dev_csr                            &M = *mat->mats[0];
    aoclsparse_status                   s;
    switch(mat->val_type)
    {
    case aoclsparse_smat:
        s = sort_rows<float>(M.m, M.nnz, M.row_ptr.as<int>(), M.col_idx, M.val, st);
        break;
    case aoclsparse_dmat:
        s = sort_rows<double>(M.m, M.nnz, M.row_ptr.as<int>(), M.col_idx, M.val, st);
        break;
    case aoclsparse_cmat:
        s = sort_rows<float2>(M.m, M.nnz, M.row_ptr.as<int>(), M.col_idx, M.val, st);
        break;
    case aoclsparse_zmat:
        s = sort_rows<double2>(M.m, M.nnz, M.row_ptr.as<int>(), M.col_idx, M.val, st);
        break;
    default:
        return aoclsparse_status_wrong_type;
    }
    if(s != aoclsparse_status_success)
        return s;
    // derived copies and analyses follow the old entry order
    drop_derived_copies(mat);
    M.plan.valid  = false;
    mat->sort = aoclsparse_fully_sorted;
    return aoclsparse_status_success;
}
}
