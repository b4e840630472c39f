// spmv.cu -- host front end of aoclsparse_{s,d,c,z}mv and the row-range extension.
//
// Mirrors the argument checking, quick returns and operation / descriptor handling of
//   aoclsparse::mv<T>            library/src/level2/aoclsparse_mv.cpp:41-349
//   aoclsparse_csrmv_t<T,false>  library/src/level2/aoclsparse_csrmv.hpp:32-450
// and launches the sm_100a kernels of spmv_kernels.cuh.  There is no CPU path: if a launch fails the
// call returns internal_error.
#include "spmv_sharded.cuh"

#include <chrono>
#include <cstdlib>
#include <map>
#include <unordered_map>

namespace b200
{
    namespace
    {
        // per-thread staging buffers for host-resident x / y (grow-only)
        struct staging
        {
            dev_buf x, y;
        };
        staging &tls_staging()
        {
            static thread_local staging s;
            return s;
        }

        template <typename T>
        aoclsparse_status scale_vector(T *y, long long len, T beta, T alpha, const T *x, long long n_unit, cudaStream_t st)
        {
            if(len <= 0)
                return aoclsparse_status_success;
            long long blocks = (len + 255) / 256;
            if(blocks > 148 * 16)
                blocks = 148 * 16;
            scale_vector_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(len, y, beta, is_zero(beta) ? 1 : 0, alpha, x, n_unit);
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }

        // CTAs of `kernel` that can be resident on the whole device at once (per kernel / shared-memory size, cached)
        template <typename K>
        long long resident_ctas(K kernel, int threads, size_t smem)
        {
            static std::mutex                       mu;
            static std::map<std::pair<const void *, size_t>, long long> cache;
            std::lock_guard<std::mutex>             lk(mu);
            const auto                              key = std::make_pair(reinterpret_cast<const void *>(kernel), smem);
            auto                                    it  = cache.find(key);
            if(it != cache.end())
                return it->second;
            int per_sm = 0, sms = 148, dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1)
            {
                cudaGetLastError();
                per_sm = 32; // unknown: assume the hardware maximum, i.e. be conservative about overlapping launches
            }
            return cache[key] = (long long)per_sm * sms;
        }

        template <typename T, bool GENERIC, int NT, bool PUSH = false, bool CODED = false, bool ECODED = false>
        aoclsparse_status launch_row_blocks(const dev_csr &A,
                                            int            b0,
                                            int            b1,
                                            const T       *x,
                                            T             *y,
                                            T              alpha,
                                            T              beta,
                                            elem_rule      rule,
                                            cudaStream_t   st,
                                            T             *push_dst  = nullptr,
                                            int            push_row0 = 0)
        {
            const row_block_plan &M    = A.plan;                       // owns the code arrays
            const row_block_plan &P    = ECODED ? *A.plan.eplan : A.plan; // the blocks this launch walks
            // (entry codes: blocks end at a row count, so the buffer is sized by the largest block, not by the capacity)
            const aoclsparse_int  e_nnz = ECODED ? (P.max_block_nnz > 0 ? P.max_block_nnz : P.block_nnz) : 0;
            const int             cap  = ECODED ? spmv_ecoded_cap(e_nnz) : P.block_nnz + (CODED ? 32 : 8);
            const size_t          smem = ECODED ? spmv_ecoded_smem_bytes(sizeof(T), e_nnz)
                                                : (CODED ? spmv_coded_smem_bytes(sizeof(T), P.block_nnz) : spmv_smem_bytes(sizeof(T), P.block_nnz));
            static std::atomic<size_t> configured{0};
            if(configured.load(std::memory_order_acquire) < smem)
            {
                B200_CUDA(cudaFuncSetAttribute(
                    spmv_row_blocks_kernel<T, GENERIC, NT, PUSH, CODED, ECODED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured.store(smem, std::memory_order_release);
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim            = dim3((unsigned)(b1 - b0));
            cfg.blockDim           = dim3(NT);
            cfg.dynamicSmemBytes   = smem;
            cfg.stream             = st;
            cudaLaunchAttribute attr[1];
            attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs                                          = attr;
            // Programmatic dependent launch only for grids of more than one wave.  The kernel signals
            // launch_dependents at entry, so with a sub-wave grid three consecutive launches of an iterated product
            // could be resident at once, and launch k-2 (still reading the buffer launch k-1 is about to overwrite)
            // could leave lines in an SM's L1 that launch k then hits through ld.global.nc after its
            // griddepcontrol.wait.  With more CTAs than fit the chip, the last CTA of launch k-1 cannot start before
            // launch k-2 has completed, so launch k (which starts after it) never overlaps k-2.
            cfg.numAttrs = (P.pdl && (long long)(b1 - b0) > resident_ctas(spmv_row_blocks_kernel<T, GENERIC, NT, PUSH, CODED, ECODED>, NT, smem)) ? 1 : 0;
            B200_CUDA(cudaLaunchKernelEx(&cfg,
                                         spmv_row_blocks_kernel<T, GENERIC, NT, PUSH, CODED, ECODED>,
                                         (const int4 *)P.desc.as<int4>(),
                                         (const int *)P.kind.as<int>(),
                                         b0,
                                         cap,
                                         (const aoclsparse_int *)A.row_ptr.as<aoclsparse_int>(),
                                         (const aoclsparse_int *)A.col_idx.as<aoclsparse_int>(),
                                         (const T *)A.val.as<T>(),
                                         x,
                                         y,
                                         alpha,
                                         beta,
                                         is_zero(beta) ? 1 : 0,
                                         P.partials.as<T>(),
                                         rule,
                                         (int)A.n,
                                         P.stream_hint,
                                         push_dst,
                                         push_row0,
                                         (const unsigned char *)(ECODED ? M.ecodes.as<unsigned char>() : M.codes.as<unsigned char>()),
                                         (const int *)(ECODED ? M.etab_off.as<int>() : M.code_offsets.as<int>()),
                                         (const T *)M.etab_val.as<T>(),
                                         (int)(ECODED ? M.n_ecodes : M.n_codes)));
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }

        // gather pass over blocks [b0, b1) of A's plan
        template <typename T>
        aoclsparse_status launch_gather(const dev_csr &A,
                                        int            b0,
                                        int            b1,
                                        int            row_lo,
                                        int            row_hi,
                                        const T       *x,
                                        T             *y,
                                        T              alpha,
                                        T              beta,
                                        bool           generic,
                                        elem_rule      rule,
                                        cudaStream_t   st,
                                        T             *push_dst = nullptr)
        {
            const row_block_plan &P = A.plan;
            if(b1 <= b0)
                return aoclsparse_status_success;
            const int bz = is_zero(beta) ? 1 : 0;
            // diagonal-code copy (aoclsparse_optimize on a banded / stencil matrix): 1 instead of 4 index bytes per entry
            const bool coded  = !generic && P.n_codes > 0 && P.n_strat[STRAT_THREAD] == P.n_blocks;
            // entry-code copy (constant-coefficient stencils): 1 byte per entry instead of 4 + sizeof(T)
            // (whole-matrix launches only: the entry-coded kernels walk a block plan of their own, P.eplan)
            const bool ecoded = coded && P.n_ecodes > 0 && !P.ecodes_stale && P.eplan && sizeof(T) <= 8 && b0 == 0 && b1 == P.n_blocks && !push_dst;
            if constexpr(sizeof(T) <= 8)
            {
                if(ecoded && P.eplan->threads == 128)
                    B200_TRY((launch_row_blocks<T, false, 128, false, false, true>(A, 0, P.eplan->n_blocks, x, y, alpha, beta, rule, st)));
                else if(ecoded)
                    B200_TRY((launch_row_blocks<T, false, 256, false, false, true>(A, 0, P.eplan->n_blocks, x, y, alpha, beta, rule, st)));
            }
            if(ecoded)
                ;
            else if(coded && push_dst)
                B200_TRY((launch_row_blocks<T, false, 256, true, true>(A, b0, b1, x, y, alpha, beta, rule, st, push_dst, row_lo)));
            else if(coded && P.threads == 128)
                B200_TRY((launch_row_blocks<T, false, 128, false, true>(A, b0, b1, x, y, alpha, beta, rule, st)));
            else if(coded)
                B200_TRY((launch_row_blocks<T, false, 256, false, true>(A, b0, b1, x, y, alpha, beta, rule, st)));
            else if(push_dst)
                B200_TRY((launch_row_blocks<T, false, 256, true>(A, b0, b1, x, y, alpha, beta, rule, st, push_dst, row_lo)));
            else if(generic)
                B200_TRY((launch_row_blocks<T, true, 256>(A, b0, b1, x, y, alpha, beta, rule, st)));
            else if(P.threads == 128)
                B200_TRY((launch_row_blocks<T, false, 128>(A, b0, b1, x, y, alpha, beta, rule, st)));
            else if(P.threads == 512)
                B200_TRY((launch_row_blocks<T, false, 512>(A, b0, b1, x, y, alpha, beta, rule, st)));
            else
                B200_TRY((launch_row_blocks<T, false, 256>(A, b0, b1, x, y, alpha, beta, rule, st)));
            if(P.n_long_rows > 0)
            {
                const long long threads = (long long)P.n_long_rows * 32;
                finish_long_rows_kernel<T><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(
                    P.n_long_rows,
                    P.long_rows.as<int4>(),
                    P.partials.as<T>(),
                    x,
                    y,
                    alpha,
                    beta,
                    bz,
                    (generic && rule.diag == DIAG_UNIT) ? 1 : 0,
                    A.n,
                    row_lo,
                    row_hi,
                    push_dst,
                    row_lo);
                B200_LAUNCHED();
            }
            return aoclsparse_status_success;
        }

        template <typename T>
        aoclsparse_status launch_scatter(const dev_csr &A, const T *x, T *y, T alpha, elem_rule rule, cudaStream_t st)
        {
            const row_block_plan &P = A.plan;
            if(P.n_blocks <= 0)
                return aoclsparse_status_success;
            spmv_scatter_kernel<T><<<P.n_blocks, SPMV_THREADS, 0, st>>>(P.desc.as<int4>(),
                                                                        A.row_ptr.as<aoclsparse_int>(),
                                                                        A.col_idx.as<aoclsparse_int>(),
                                                                        A.val.as<T>(),
                                                                        x,
                                                                        y,
                                                                        alpha,
                                                                        rule);
            B200_LAUNCHED();
            return aoclsparse_status_success;
        }

        int diag_rule(aoclsparse_diag_type d)
        {
            return d == aoclsparse_diag_type_unit ? DIAG_UNIT : (d == aoclsparse_diag_type_zero ? DIAG_ZERO : DIAG_KEEP);
        }

        inline bool valid_op(aoclsparse_operation op)
        {
            return op == aoclsparse_operation_none || op == aoclsparse_operation_transpose
                   || op == aoclsparse_operation_conjugate_transpose;
        }
        inline bool valid_type(aoclsparse_matrix_type t)
        {
            return t == aoclsparse_matrix_type_general || t == aoclsparse_matrix_type_symmetric
                   || t == aoclsparse_matrix_type_hermitian || t == aoclsparse_matrix_type_triangular;
        }
    }

    // The device-side multiply, x and y already device accessible.
    // calls of an un-hinted transposed / symmetric / hermitian product on one handle after which the derived copy is built.
    // Building it costs about as much as a hundred products save (a device sort of all entries, ~25 ms on 56 M entries
    // against 0.2-0.3 ms saved per product), so only a handle that is clearly being iterated on gets one.
    constexpr int LAZY_COPY_AFTER = 32;

    template <typename T>
    aoclsparse_status mv_device(aoclsparse_operation       op,
                                T                          alpha,
                                aoclsparse_matrix          A,
                                const _aoclsparse_mat_descr &descr,
                                const T                   *x,
                                T                          beta,
                                T                         *y,
                                cudaStream_t               st)
    {
        B200_TRY(ensure_plan(A, st));
        // symmetric / hermitian descriptor that was hinted (aoclsparse_set_mv_hint + aoclsparse_optimize): multiply with
        // the expanded general copy (the reference materialises the same copy in aoclsparse_matrix_transform,
        // csr_util.hpp:620-745) -- a plain streaming gather instead of gather + atomic scatter
        if((descr.type == aoclsparse_matrix_type_symmetric || descr.type == aoclsparse_matrix_type_hermitian)
           && A->mem_policy == aoclsparse_memory_usage_unrestricted && A->win_hi < 0)
        {
            const int d_id  = get_doid(vt<T>::is_complex, descr.type, descr.fill_mode, op);
            bool      hinted = false;
            {
                std::shared_lock<std::shared_mutex> rl0(A->guard);
                for(const hint &h : A->hints)
                    hinted = hinted || ((h.act == 1 || h.act == 8) && h.doid == d_id && h.done);
            }
            // no hint, but the same kind of product keeps coming: build the copy after a few calls, as the reference's
            // mv lazily builds its optimised CSR (csr_util.hpp:812-825).  Gather + atomic scatter costs ~3.7x the
            // streaming gather on the 27-point stencil (profiles/r01_mm_sweep.txt).
            if(!hinted && A->lazy_copy_calls.fetch_add(1, std::memory_order_relaxed) + 1 >= LAZY_COPY_AFTER)
                hinted = true;
            if(hinted)
            {
                const dev_csr *F = nullptr;
                B200_TRY(get_expanded_copy(A, descr, op, F, st));
                std::shared_lock<std::shared_mutex> rl1(A->guard);
                elem_rule none_rule{MASK_NONE, DIAG_KEEP, 0, 0};
                return launch_gather<T>(*F, 0, F->plan.n_blocks, 0, F->m, x, y, alpha, beta, false, none_rule, st);
            }
        }
        // un-hinted transposed general products: same lazy policy, a transposed copy turns the atomic scatter into a gather
        if(descr.type == aoclsparse_matrix_type_general && A->mem_policy == aoclsparse_memory_usage_unrestricted && A->win_hi < 0
           && (op != aoclsparse_operation_none) != A->is_csc)
        {
            const int want = (vt<T>::is_complex && op == aoclsparse_operation_conjugate_transpose) ? DOID_GH : DOID_GT;
            bool      have = false;
            {
                std::shared_lock<std::shared_mutex> rl0(A->guard);
                for(size_t i = 1; i < A->mats.size(); ++i)
                    have = have || (A->mats[i]->doid == want && A->mats[i]->plan.valid);
            }
            if(!have && A->lazy_copy_calls.fetch_add(1, std::memory_order_relaxed) + 1 >= LAZY_COPY_AFTER)
            {
                std::unique_lock<std::shared_mutex> wl(A->guard);
                for(size_t i = 1; i < A->mats.size(); ++i)
                    have = have || (A->mats[i]->doid == want && A->mats[i]->plan.valid);
                if(!have)
                {
                    dev_csr *C = new(std::nothrow) dev_csr;
                    if(C)
                    {
                        aoclsparse_status s = transpose_csr(*A->mats[0], A->val_type, want == DOID_GH, *C, st);
                        if(s == aoclsparse_status_success)
                            s = build_plan(*C, sizeof(T), -1, -1, std::vector<aoclsparse_int>(), st);
                        if(s == aoclsparse_status_success)
                        {
                            C->doid = want;
                            A->mats.push_back(C);
                        }
                        else
                            delete C; // no room: keep scattering
                    }
                }
            }
        }
        std::shared_lock<std::shared_mutex> rl(A->guard);
        const dev_csr                      &M = *A->mats[0];
        const bool                          cplx = vt<T>::is_complex;

        if(A->win_hi >= 0)
        {
            // windowed x: only the plain general product is defined on a row shard
            if(descr.type != aoclsparse_matrix_type_general || op != aoclsparse_operation_none)
                return aoclsparse_status_not_implemented;
            x = x - A->win_lo;
        }

        // what has to be applied to the STORED matrix S: a CSC handle stores S = A^T, so the transposition flips and the
        // stored triangle is the other one; conjugation is unaffected (op H on a CSC handle = conj(S) without transpose)
        const bool conj_op = cplx && op == aoclsparse_operation_conjugate_transpose;
        bool       trans   = op != aoclsparse_operation_none;
        int        fill    = descr.fill_mode;
        if(A->is_csc)
        {
            trans = !trans;
            fill  = fill == aoclsparse_fill_mode_lower ? aoclsparse_fill_mode_upper : aoclsparse_fill_mode_lower;
        }
        elem_rule none_rule{MASK_NONE, DIAG_KEEP, 0, 0};
        const int cj_flag = conj_op ? 1 : 0;

        switch(descr.type)
        {
        case aoclsparse_matrix_type_general:
            if(!trans)
            {
                if(!conj_op)
                    return launch_gather<T>(M, 0, M.plan.n_blocks, 0, M.m, x, y, alpha, beta, false, none_rule, st);
                elem_rule r{MASK_NONE, DIAG_KEEP, 1, 1}; // conj(S) x
                return launch_gather<T>(M, 0, M.plan.n_blocks, 0, M.m, x, y, alpha, beta, true, r, st);
            }
            else
            {
                // an explicitly transposed copy built by aoclsparse_optimize turns this into a gather
                for(size_t i = 1; i < A->mats.size(); ++i)
                {
                    const dev_csr &C = *A->mats[i];
                    if(C.doid == (conj_op ? DOID_GH : DOID_GT) && C.plan.valid)
                        return launch_gather<T>(C, 0, C.plan.n_blocks, 0, C.m, x, y, alpha, beta, false, none_rule, st);
                }
                B200_TRY(scale_vector<T>(y, M.n, beta, alpha, x, 0, st));
                elem_rule r{MASK_NONE, DIAG_KEEP, cj_flag, cj_flag};
                return launch_scatter<T>(M, x, y, alpha, r, st);
            }
        case aoclsparse_matrix_type_triangular:
        {
            const int mask = fill == aoclsparse_fill_mode_lower ? MASK_LOWER : MASK_UPPER;
            const int dg   = diag_rule(descr.diag_type);
            if(!trans)
            {
                elem_rule r{mask, dg, cj_flag, cj_flag};
                return launch_gather<T>(M, 0, M.plan.n_blocks, 0, M.m, x, y, alpha, beta, true, r, st);
            }
            // y (length n) = beta*y [+ alpha*x on the unit diagonal], then scatter the kept entries
            const long long n_unit = (dg == DIAG_UNIT) ? (M.m < M.n ? M.m : M.n) : 0;
            B200_TRY(scale_vector<T>(y, M.n, beta, alpha, x, n_unit, st));
            elem_rule r{mask, dg, cj_flag, cj_flag};
            return launch_scatter<T>(M, x, y, alpha, r, st);
        }
        case aoclsparse_matrix_type_symmetric:
        case aoclsparse_matrix_type_hermitian:
        {
            const bool herm = descr.type == aoclsparse_matrix_type_hermitian;
            const int  mask = fill == aoclsparse_fill_mode_lower ? MASK_LOWER : MASK_UPPER;
            const int  dg   = diag_rule(descr.diag_type);
            // stored triangle T of S, mirror:  symmetric T^T, hermitian T^H;  F_S = T + D + mirror.
            //   symmetric: F_A = F_S whether CSR or CSC; op none/T -> F_S ; op H -> conj of all of it
            //   hermitian: CSR F_A = F_S, CSC F_A = conj(F_S);  op T conjugates once more, op H does nothing
            //              => conj(T) + D + T^T  when (CSC xor op == T), else T + D + conj(T)^T
            // (D as stored; the reference does not conjugate a hermitian diagonal,
            //  aoclsparse_csrmv_kr.hpp:398-401,422-425, but does conjugate a symmetric one for op H, :217-218)
            int cg, cs, cd;
            if(!herm)
            {
                cg = cs = cd = cj_flag;
            }
            else
            {
                const bool t = cplx && ((op == aoclsparse_operation_transpose) != A->is_csc);
                cg           = t ? 1 : 0;
                cs           = t ? 0 : 1;
                cd           = 0;
            }
            // pass 1 (gather): y = beta*y + alpha*(T_strict + D) x
            elem_rule rg{mask, dg, cg, cd};
            B200_TRY(launch_gather<T>(M, 0, M.plan.n_blocks, 0, M.m, x, y, alpha, beta, true, rg, st));
            // pass 2 (scatter): y[col] += alpha * op(a) * x[row] over the strict triangle
            elem_rule rs{mask, DIAG_ZERO, cs, cs};
            return launch_scatter<T>(M, x, y, alpha, rs, st);
        }
        }
        return aoclsparse_status_invalid_value;
    }

    namespace
    {
        __global__ void block_maxcol_kernel(int nblocks, const int4 *__restrict__ desc, const aoclsparse_int *__restrict__ col, int *out)
        {
            const int lane = threadIdx.x & 31;
            const int b    = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
            if(b >= nblocks)
                return;
            const int4 d = desc[b];
            int        mx = -1;
            for(int p = d.z + lane; p < d.w; p += 32)
                mx = max(mx, col[p]);
            for(int off = 16; off > 0; off >>= 1)
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            if(lane == 0)
                out[b] = mx;
        }

        constexpr int HOST_CHUNKS_MAX = 32;
        // builds (once per plan) the chunk table of the host-staged pipeline
        aoclsparse_status ensure_host_chunks(aoclsparse_matrix A, size_t elem_size, cudaStream_t st)
        {
            std::unique_lock<std::shared_mutex> wl(A->guard);
            dev_csr                            &M = *A->mats[0];
            row_block_plan                     &P = M.plan;
            if(P.host_chunks_ready)
                return aoclsparse_status_success;
            P.host_chunks.clear();
            P.host_chunks_ready = true;
            // only plans without split rows (their partial sums are finished per launch) and big enough to matter
            const size_t vec_bytes = ((size_t)M.m + (size_t)M.n) * elem_size;
            if(P.n_long_rows > 0 || P.n_blocks < 64 || vec_bytes < (size_t)(4u << 20))
                return aoclsparse_status_success;
            dev_buf d_mx;
            B200_TRY(d_mx.alloc(sizeof(int) * (size_t)P.n_blocks));
            const long long threads = (long long)P.n_blocks * 32;
            block_maxcol_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
                P.n_blocks, P.desc.as<int4>(), M.col_idx.as<aoclsparse_int>(), d_mx.as<int>());
            B200_LAUNCHED();
            std::vector<int>  mx((size_t)P.n_blocks);
            std::vector<int4> hd((size_t)P.n_blocks);
            B200_CUDA(cudaMemcpyAsync(mx.data(), d_mx.p, sizeof(int) * mx.size(), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaMemcpyAsync(hd.data(), P.desc.p, sizeof(int4) * hd.size(), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            // ~4 MB of vector traffic per chunk, 2..8 chunks (measured on C2: 2 chunks 0.594 ms, 8 chunks 0.546 ms,
            // 16 chunks 0.62 ms; un-pipelined 0.755 ms)
            int n_chunks = (int)(vec_bytes / (size_t)(4u << 20));
            n_chunks     = n_chunks < 2 ? 2 : (n_chunks > 8 ? 8 : n_chunks);
            // very long vectors (config 5: 2 x 1 GB): the copies dominate and the un-overlapped tail is the last chunk's
            // kernel + read-back, so cut finer
            if(vec_bytes >= ((size_t)512u << 20))
                n_chunks = 32;
            if(const char *e = getenv("AOCLSPARSE_B200_HOST_CHUNKS"))
                n_chunks = atoi(e) < 1 ? 1 : (atoi(e) > HOST_CHUNKS_MAX ? HOST_CHUNKS_MAX : atoi(e));
            int run_max  = -1;
            for(int c = 0; c < n_chunks; ++c)
            {
                row_block_plan::host_chunk h;
                h.b0 = (aoclsparse_int)((long long)P.n_blocks * c / n_chunks);
                h.b1 = (aoclsparse_int)((long long)P.n_blocks * (c + 1) / n_chunks);
                if(h.b1 <= h.b0)
                    continue;
                for(int b = h.b0; b < h.b1; ++b)
                    run_max = mx[b] > run_max ? mx[b] : run_max;
                h.row0 = hd[h.b0].x;
                h.row1 = hd[h.b1 - 1].y;
                h.x_hi = run_max + 1;
                P.host_chunks.push_back(h);
            }
            if(P.host_chunks.size() < 2)
                P.host_chunks.clear();
            return aoclsparse_status_success;
        }

        struct host_pipe
        {
            cudaStream_t h2d = nullptr, d2h = nullptr;
            cudaEvent_t  ev_x[HOST_CHUNKS_MAX] = {}, ev_k[HOST_CHUNKS_MAX] = {}, ev_start = nullptr;
            cudaEvent_t  tr_x[HOST_CHUNKS_MAX] = {}, tr_k[HOST_CHUNKS_MAX] = {}, tr_y[HOST_CHUNKS_MAX] = {}, tr_0 = nullptr; // AOCLSPARSE_B200_HOST_TRACE=1 only
            bool         ready = false, trace = false;
            aoclsparse_status init()
            {
                if(ready)
                    return aoclsparse_status_success;
                B200_CUDA(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
                B200_CUDA(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
                for(int i = 0; i < HOST_CHUNKS_MAX; ++i)
                {
                    B200_CUDA(cudaEventCreateWithFlags(&ev_x[i], cudaEventDisableTiming));
                    B200_CUDA(cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming));
                }
                B200_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
                trace = getenv("AOCLSPARSE_B200_HOST_TRACE") && atoi(getenv("AOCLSPARSE_B200_HOST_TRACE")) != 0;
                if(trace)
                {
                    B200_CUDA(cudaEventCreate(&tr_0));
                    for(int i = 0; i < HOST_CHUNKS_MAX; ++i)
                    {
                        B200_CUDA(cudaEventCreate(&tr_x[i]));
                        B200_CUDA(cudaEventCreate(&tr_k[i]));
                        B200_CUDA(cudaEventCreate(&tr_y[i]));
                    }
                }
                ready = true;
                return aoclsparse_status_success;
            }
        };
        host_pipe &tls_pipe()
        {
            static thread_local host_pipe p;
            return p;
        }

        // y = alpha*A*x + beta*y with x and y in HOST memory, general matrix, op = none: x arrives in pieces, each
        // chunk of row blocks starts as soon as the prefix of x it reads is on the device, and its slice of y leaves
        // while later chunks still compute -- H2D, kernels and D2H overlap (PCIe is full duplex)
        template <typename T>
        aoclsparse_status mv_host_pipelined(aoclsparse_matrix A, T alpha, const T *hx, T beta, T *hy, cudaStream_t st, bool &done)
        {
            done = false;
            B200_TRY(ensure_plan(A, st));
            B200_TRY(ensure_host_chunks(A, sizeof(T), st));
            std::shared_lock<std::shared_mutex> rl(A->guard);
            const dev_csr                      &M = *A->mats[0];
            const row_block_plan               &P = M.plan;
            if(P.host_chunks.empty())
                return aoclsparse_status_success;
            host_pipe &hp = tls_pipe();
            B200_TRY(hp.init());
            staging &sg = tls_staging();
            if(sg.x.bytes < (size_t)M.n * sizeof(T))
                B200_TRY(sg.x.alloc((size_t)M.n * sizeof(T)));
            if(sg.y.bytes < (size_t)M.m * sizeof(T))
                B200_TRY(sg.y.alloc((size_t)M.m * sizeof(T)));
            T         *dx = sg.x.as<T>(), *dy = sg.y.as<T>();
            const bool bz = is_zero(beta);
            elem_rule  none_rule{MASK_NONE, DIAG_KEEP, 0, 0};
            // experiment, OFF by default (AOCLSPARSE_B200_HOST_DIRECT=1): with page-locked y and beta == 0 the kernels
            // store y straight into host memory instead of D2H copies.  Measured SLOWER on C2 (0.82 ms against 0.58 ms
            // with the copies): SM-issued posted writes over PCIe run well below the copy engine's 53 GB/s and hold the
            // CTAs resident while they drain.
            const bool        direct_ok = getenv("AOCLSPARSE_B200_HOST_DIRECT") && atoi(getenv("AOCLSPARSE_B200_HOST_DIRECT")) != 0;
            T                *y_direct  = (bz && direct_ok) ? static_cast<T *>(pinned_host_device_ptr(hy)) : nullptr;
            // the side streams start after whatever the caller's stream was doing
            const auto host_t0 = std::chrono::steady_clock::now();
            if(hp.trace)
                B200_CUDA(cudaEventRecord(hp.tr_0, st));
            B200_CUDA(cudaEventRecord(hp.ev_start, st));
            B200_CUDA(cudaStreamWaitEvent(hp.h2d, hp.ev_start, 0));
            B200_CUDA(cudaStreamWaitEvent(hp.d2h, hp.ev_start, 0));
            const int nc    = (int)P.host_chunks.size();
            int       x_lo  = 0;
            // one loop: the copy of chunk c's slice of x is followed at once by chunk c's kernel and read-back, so the
            // first kernel is already queued when its data lands (enqueueing all copies first cost ~30 us of start-up;
            // profiles/r01_summary.md, host-resident vectors)
            for(int c = 0; c < nc; ++c)
            {
                const auto &h   = P.host_chunks[c];
                int         xhi = c == nc - 1 ? M.n : (h.x_hi > x_lo ? h.x_hi : x_lo); // the tail of x goes with the last chunk
                if(xhi > x_lo)
                    B200_CUDA(cudaMemcpyAsync(dx + x_lo, hx + x_lo, (size_t)(xhi - x_lo) * sizeof(T), cudaMemcpyHostToDevice, hp.h2d));
                x_lo = xhi;
                if(!bz && h.row1 > h.row0)
                    B200_CUDA(cudaMemcpyAsync(
                        dy + h.row0, hy + h.row0, (size_t)(h.row1 - h.row0) * sizeof(T), cudaMemcpyHostToDevice, hp.h2d));
                B200_CUDA(cudaEventRecord(hp.ev_x[c], hp.h2d));
                if(hp.trace)
                    B200_CUDA(cudaEventRecord(hp.tr_x[c], hp.h2d));
                B200_CUDA(cudaStreamWaitEvent(st, hp.ev_x[c], 0));
                if(y_direct)
                {
                    B200_TRY(launch_gather<T>(M, h.b0, h.b1, h.row0, h.row1, dx, y_direct, alpha, beta, false, none_rule, st));
                    continue;
                }
                B200_TRY(launch_gather<T>(M, h.b0, h.b1, h.row0, h.row1, dx, dy, alpha, beta, false, none_rule, st));
                B200_CUDA(cudaEventRecord(hp.ev_k[c], st));
                if(hp.trace)
                    B200_CUDA(cudaEventRecord(hp.tr_k[c], st));
                B200_CUDA(cudaStreamWaitEvent(hp.d2h, hp.ev_k[c], 0));
                if(h.row1 > h.row0)
                    B200_CUDA(cudaMemcpyAsync(
                        hy + h.row0, dy + h.row0, (size_t)(h.row1 - h.row0) * sizeof(T), cudaMemcpyDeviceToHost, hp.d2h));
                if(hp.trace)
                    B200_CUDA(cudaEventRecord(hp.tr_y[c], hp.d2h));
            }
            const auto host_t1 = std::chrono::steady_clock::now();
            if(!y_direct)
                B200_CUDA(cudaStreamSynchronize(hp.d2h));
            B200_CUDA(cudaStreamSynchronize(st));
            if(hp.trace && !y_direct)
            {
                const auto host_t2 = std::chrono::steady_clock::now();
                fprintf(stderr, "[host pipe] enqueue %.1f us, total %.1f us on the host; device times since start (us):\n",
                        std::chrono::duration<double, std::micro>(host_t1 - host_t0).count(),
                        std::chrono::duration<double, std::micro>(host_t2 - host_t0).count());
                for(int c = 0; c < nc; ++c)
                {
                    float tx = 0, tk = 0, ty = 0;
                    cudaEventElapsedTime(&tx, hp.tr_0, hp.tr_x[c]);
                    cudaEventElapsedTime(&tk, hp.tr_0, hp.tr_k[c]);
                    cudaEventElapsedTime(&ty, hp.tr_0, hp.tr_y[c]);
                    fprintf(stderr, "   chunk %d: x arrived %7.1f  kernel done %7.1f  y delivered %7.1f\n", c, tx * 1e3, tk * 1e3, ty * 1e3);
                }
            }
            done = true;
            return aoclsparse_status_success;
        }
    }

    // Common validation + staging.  Mirrors aoclsparse::mv<T> (mv.cpp:55-121).
    template <typename T>
    aoclsparse_status mv_entry(aoclsparse_operation       op,
                               const T                   *alpha,
                               aoclsparse_matrix          A,
                               const aoclsparse_mat_descr descr,
                               const T                   *x,
                               const T                   *beta,
                               T                         *y)
    {
        if(alpha == nullptr || beta == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(A == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(descr == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(x == nullptr || y == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(A->mats.empty() || A->mats[0] == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(descr->base != A->base)
            return aoclsparse_status_invalid_value;
        if(!valid_op(op))
            return aoclsparse_status_invalid_value;
        if(A->val_type != vt<T>::data_type)
            return aoclsparse_status_wrong_type;
        if(!valid_type(descr->type))
            return aoclsparse_status_invalid_value;
        if((descr->type == aoclsparse_matrix_type_symmetric || descr->type == aoclsparse_matrix_type_hermitian)
           && A->m != A->n)
            return aoclsparse_status_invalid_size;
        if(!vt<T>::is_complex)
        {
            if(op == aoclsparse_operation_conjugate_transpose)
                op = aoclsparse_operation_transpose;
            if(descr->type == aoclsparse_matrix_type_hermitian)
                return aoclsparse_status_not_implemented;
        }
        // Reference quirk kept for drop-in fidelity: a GENERAL descriptor whose diag_type is unit / zero
        // makes the plain product fail with invalid_pointer (mv.cpp:221-226 calls aoclsparse_set_mat_diag on
        // a matrix that has no diagonal bookkeeping, csr_util.hpp:478-480); transposed products ignore it.
        // (for a CSC handle the stored matrix is the transpose, so it is the TRANSPOSED product that trips it)
        if(descr->type == aoclsparse_matrix_type_general
           && op == (A->is_csc ? aoclsparse_operation_transpose : aoclsparse_operation_none)
           && descr->diag_type != aoclsparse_diag_type_non_unit && !(A->m == 0 || A->n == 0 || A->nnz == 0))
            return aoclsparse_status_invalid_pointer;

        cudaStream_t    st    = current_stream();
        const long long x_len = (A->win_hi >= 0) ? (long long)(A->win_hi - A->win_lo)
                                                 : (op == aoclsparse_operation_none ? A->n : A->m);
        const long long y_len = op == aoclsparse_operation_none ? A->m : A->n;

        // requested kernel id (aoclsparse_set_mv_hint_kid): -1 auto, 0..2 force a row strategy
        {
            std::shared_lock<std::shared_mutex> rl(A->guard);
            const int d_id = get_doid(vt<T>::is_complex, descr->type, descr->fill_mode, op);
            for(const hint &h : A->hints)
                if(h.act == 1 && h.doid == d_id)
                {
                    if(h.kid > 2)
                        return aoclsparse_status_invalid_kid;
                    break;
                }
        }

        const bool x_dev = is_device_accessible(x), y_dev = is_device_accessible(y);
        const bool empty = A->m == 0 || A->n == 0 || (A->nnz == 0 && descr->type == aoclsparse_matrix_type_general);

        // Reference behaviour kept (tests/golden/ref_csc_sweep.json): the conjugate-transposed product of a complex
        // CSC handle needs conj(S) x without a transposition, which its general / triangular kernels do not provide
        // (get_effective_doid -> gc; aoclsparse_csrmv.hpp dispatch) -> not_implemented.  mv_device can compute it
        // (the gather kernel's conjugating element rule); set AOCLSPARSE_B200_CSC_CONJ=1 to get the product instead.
        if(A->is_csc && vt<T>::is_complex && op == aoclsparse_operation_conjugate_transpose && !empty
           && (descr->type == aoclsparse_matrix_type_general || descr->type == aoclsparse_matrix_type_triangular))
        {
            static const bool allow = [] {
                const char *e = getenv("AOCLSPARSE_B200_CSC_CONJ");
                return e && atoi(e) != 0;
            }();
            if(!allow)
                return aoclsparse_status_not_implemented;
        }

        // both vectors on the host, plain general product: chunked pipeline over three streams
        if(!x_dev && !y_dev && !empty && descr->type == aoclsparse_matrix_type_general && op == aoclsparse_operation_none
           && A->win_hi < 0 && !A->is_csc)
        {
            bool              done = false;
            aoclsparse_status ps   = mv_host_pipelined<T>(A, *alpha, x, *beta, y, st, done);
            if(ps != aoclsparse_status_success)
                return ps;
            if(done)
                return aoclsparse_status_success;
        }

        const T *dx = x;
        T       *dy = y;
        staging &sg = tls_staging();
        if(!y_dev)
        {
            if(sg.y.bytes < (size_t)y_len * sizeof(T))
                B200_TRY(sg.y.alloc((size_t)y_len * sizeof(T)));
            dy = sg.y.as<T>();
            if(!is_zero(*beta) && y_len > 0)
                B200_CUDA(cudaMemcpyAsync(dy, y, (size_t)y_len * sizeof(T), cudaMemcpyHostToDevice, st));
        }
        if(!x_dev && !empty)
        {
            if(sg.x.bytes < (size_t)x_len * sizeof(T))
                B200_TRY(sg.x.alloc((size_t)x_len * sizeof(T)));
            B200_CUDA(cudaMemcpyAsync(sg.x.p, x, (size_t)x_len * sizeof(T), cudaMemcpyHostToDevice, st));
            dx = sg.x.as<T>();
        }

        aoclsparse_status status;
        if(empty)
            status = scale_vector<T>(dy, y_len, *beta, *alpha, nullptr, 0, st); // mv.cpp:116-121
        else
            status = mv_device<T>(op, *alpha, A, *descr, dx, *beta, dy, st);
        if(status != aoclsparse_status_success)
            return status;

        if(!y_dev)
        {
            if(y_len > 0)
                B200_CUDA(cudaMemcpyAsync(y, dy, (size_t)y_len * sizeof(T), cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
        }
        else if(!x_dev)
            B200_CUDA(cudaStreamSynchronize(st)); // staging buffer may be reused by the next call
        return aoclsparse_status_success;
    }

    template <typename T>
    aoclsparse_status mv_rows_entry(const T                   *alpha,
                                    aoclsparse_matrix          A,
                                    const aoclsparse_mat_descr descr,
                                    const T                   *x,
                                    const T                   *beta,
                                    T                         *y,
                                    aoclsparse_int             row_begin,
                                    aoclsparse_int             row_end,
                                    T                         *push_dst = nullptr)
    {
        if(!alpha || !beta || !A || !descr || !x || !y)
            return aoclsparse_status_invalid_pointer;
        if(A->mats.empty() || A->mats[0] == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(descr->base != A->base)
            return aoclsparse_status_invalid_value;
        if(A->val_type != vt<T>::data_type)
            return aoclsparse_status_wrong_type;
        if(descr->type != aoclsparse_matrix_type_general || A->is_csc)
            return aoclsparse_status_not_implemented;
        if(!is_device_accessible(x) || !is_device_accessible(y))
            return aoclsparse_status_invalid_pointer;
        if(row_begin < 0 || row_end > A->m || row_begin > row_end)
            return aoclsparse_status_invalid_size;
        cudaStream_t st = current_stream();
        B200_TRY(ensure_plan(A, st));
        std::shared_lock<std::shared_mutex> rl(A->guard);
        const dev_csr                      &M = *A->mats[0];
        auto block_of = [&](aoclsparse_int row, int &blk) -> bool {
            if(row == 0)
            {
                blk = 0;
                return true;
            }
            if(row == A->m)
            {
                blk = M.plan.n_blocks;
                return true;
            }
            for(size_t i = 0; i < A->row_cuts.size(); ++i)
                if(A->row_cuts[i] == row && i < M.plan.cut_block.size())
                {
                    blk = M.plan.cut_block[i];
                    return true;
                }
            return false;
        };
        int b0, b1;
        if(!block_of(row_begin, b0) || !block_of(row_end, b1))
            return aoclsparse_status_invalid_value;
        if(A->win_hi >= 0)
            x = x - A->win_lo;
        elem_rule none_rule{MASK_NONE, DIAG_KEEP, 0, 0};
        return launch_gather<T>(M, b0, b1, row_begin, row_end, x, y, *alpha, *beta, false, none_rule, st, push_dst);
    }
}

namespace b200
{
    // Handle-free legacy entry: aoclsparse_csrmv_t<T, true> (csrmv.hpp:63-110 for the checks).
    template <typename T>
    aoclsparse_status csrmv_legacy(aoclsparse_operation       trans,
                                   const T                   *alpha,
                                   aoclsparse_int             m,
                                   aoclsparse_int             n,
                                   aoclsparse_int             nnz,
                                   const T                   *csr_val,
                                   const aoclsparse_int      *csr_col_ind,
                                   const aoclsparse_int      *csr_row_ptr,
                                   const aoclsparse_mat_descr descr,
                                   const T                   *x,
                                   const T                   *beta,
                                   T                         *y)
    {
        if(alpha == nullptr || beta == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(descr == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(descr->base != aoclsparse_index_base_zero && descr->base != aoclsparse_index_base_one)
            return aoclsparse_status_invalid_value;
        if(!valid_type(descr->type))
            return aoclsparse_status_invalid_value;
        if(!valid_op(trans))
            return aoclsparse_status_invalid_value;
        if(descr->type != aoclsparse_matrix_type_general && descr->type != aoclsparse_matrix_type_symmetric)
            return aoclsparse_status_not_implemented;
        if(descr->type == aoclsparse_matrix_type_symmetric && m != n)
            return aoclsparse_status_invalid_size;
        if(m < 0 || n < 0 || nnz < 0)
            return aoclsparse_status_invalid_size;
        if(csr_val == nullptr || csr_row_ptr == nullptr || csr_col_ind == nullptr || x == nullptr || y == nullptr)
            return aoclsparse_status_invalid_pointer;

        aoclsparse_matrix A = nullptr;
        B200_TRY(create_temp_csr(&A, vt<T>::data_type, descr->base, m, n, nnz, csr_row_ptr, csr_col_ind, csr_val));
        _aoclsparse_mat_descr d = *descr;
        if(d.type == aoclsparse_matrix_type_symmetric)
        {
            // the legacy symmetric kernel (aoclsparse_csrmv_symm, csrmv_kr.hpp:41-91) reads the stored LOWER
            // triangle with its diagonal and ignores fill_mode / diag_type
            d.fill_mode = aoclsparse_fill_mode_lower;
            d.diag_type = aoclsparse_diag_type_non_unit;
        }
        else
            d.diag_type = aoclsparse_diag_type_non_unit;
        aoclsparse_status s = mv_entry<T>(trans, alpha, A, &d, x, beta, y);
        cudaStreamSynchronize(current_stream());
        delete A;
        return s;
    }
}

namespace b200
{
    namespace
    {
        // d = sum_i conj(x_i) * y_i, deterministic two-level reduction (aoclsparse::dense_dot,
        // library/src/level1/aoclsparse_dense_dot_kt.cpp:30-66)
        template <typename T>
        __global__ void dotc_partial_kernel(long long n, const T *__restrict__ x, const T *__restrict__ y, T *partial)
        {
            __shared__ T sh[8];
            T            acc = vt<T>::zero();
            for(long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
                acc = mad(cj(x[i]), y[i], acc);
            acc = warp_sum(acc);
            if((threadIdx.x & 31) == 0)
                sh[threadIdx.x >> 5] = acc;
            __syncthreads();
            if(threadIdx.x == 0)
            {
                T t = sh[0];
                for(int w = 1; w < (int)(blockDim.x >> 5); ++w)
                    t = add(t, sh[w]);
                partial[blockIdx.x] = t;
            }
        }
        template <typename T>
        __global__ void dotc_final_kernel(int n, const T *__restrict__ partial, T *out)
        {
            T acc = vt<T>::zero();
            for(int i = threadIdx.x; i < n; i += 32)
                acc = add(acc, partial[i]);
            acc = warp_sum(acc);
            if(threadIdx.x == 0)
                *out = acc;
        }
    }

    // y = alpha*op(A)*x + beta*y, d = x^H y over the first min(m,n) entries: aoclsparse_dotmv_t
    // (library/src/level2/aoclsparse_dotmv.hpp:30-62)
    template <typename T>
    aoclsparse_status dotmv_entry(aoclsparse_operation       op,
                                  T                          alpha,
                                  aoclsparse_matrix          A,
                                  const aoclsparse_mat_descr descr,
                                  const T                   *x,
                                  T                          beta,
                                  T                         *y,
                                  T                         *d)
    {
        if(d == nullptr || A == nullptr)
            return aoclsparse_status_invalid_pointer;
        B200_TRY(mv_entry<T>(op, &alpha, A, descr, x, &beta, y));
        cudaStream_t    st  = current_stream();
        const long long len = A->m < A->n ? A->m : A->n;
        const bool      x_dev = is_device_accessible(x), y_dev = is_device_accessible(y), d_dev = is_device_accessible(d);
        dev_buf         tx, ty, part;
        const T        *dx = x;
        const T        *dy = y;
        if(!x_dev && len > 0)
        {
            B200_TRY(tx.alloc(sizeof(T) * (size_t)len));
            B200_CUDA(cudaMemcpyAsync(tx.p, x, sizeof(T) * (size_t)len, cudaMemcpyHostToDevice, st));
            dx = tx.as<T>();
        }
        if(!y_dev && len > 0)
        {
            B200_TRY(ty.alloc(sizeof(T) * (size_t)len));
            B200_CUDA(cudaMemcpyAsync(ty.p, y, sizeof(T) * (size_t)len, cudaMemcpyHostToDevice, st));
            dy = ty.as<T>();
        }
        const int nb = 148 * 4;
        B200_TRY(part.alloc(sizeof(T) * (size_t)(nb + 1)));
        dotc_partial_kernel<T><<<nb, 256, 0, st>>>(len, dx, dy, part.as<T>());
        B200_LAUNCHED();
        T *dout = d_dev ? d : part.as<T>() + nb;
        dotc_final_kernel<T><<<1, 32, 0, st>>>(nb, part.as<T>(), dout);
        B200_LAUNCHED();
        if(!d_dev)
            B200_CUDA(cudaMemcpyAsync(d, dout, sizeof(T), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st)); // temporaries are freed on return
        return aoclsparse_status_success;
    }

    // aoclsparse_set_value_t (library/src/extra/aoclsparse_auxiliary.hpp:388-473): overwrite one stored entry
    namespace
    {
        template <typename T>
        __global__ void set_value_kernel(const aoclsparse_int *__restrict__ rp, const aoclsparse_int *__restrict__ col, T *val, int row, int c, T v, int *found)
        {
            // first match in storage order, like the reference's scan
            int hit = -1;
            for(int p = rp[row]; p < rp[row + 1]; ++p)
                if(col[p] == c)
                {
                    hit = p;
                    break;
                }
            if(hit >= 0)
                val[hit] = v;
            *found = hit >= 0 ? 1 : 0;
        }
    }

    template <typename T>
    aoclsparse_status set_value_entry(aoclsparse_matrix A, aoclsparse_int row_idx, aoclsparse_int col_idx, T v)
    {
        if(A == nullptr)
            return aoclsparse_status_invalid_pointer;
        if(A->mats.empty() || A->mats[0] == nullptr)
            return aoclsparse_status_invalid_pointer;
        const aoclsparse_int base = A->base;
        if(A->m + base <= row_idx || row_idx < base || A->n + base <= col_idx || col_idx < base)
            return aoclsparse_status_invalid_value;
        if(A->is_csc) // the stored arrays are those of the transpose (auxiliary.hpp:444-447)
        {
            const aoclsparse_int t = row_idx;
            row_idx                = col_idx;
            col_idx                = t;
        }
        if(A->val_type != vt<T>::data_type)
            return aoclsparse_status_wrong_type;
        cudaStream_t                        st = current_stream();
        std::unique_lock<std::shared_mutex> wl(A->guard);
        dev_csr                            &M = *A->mats[0];
        dev_buf                             flag;
        B200_TRY(flag.alloc(sizeof(int)));
        set_value_kernel<T><<<1, 1, 0, st>>>(M.row_ptr.as<aoclsparse_int>(),
                                             M.col_idx.as<aoclsparse_int>(),
                                             M.val.as<T>(),
                                             (int)(row_idx - base),
                                             (int)(col_idx - base),
                                             v,
                                             flag.as<int>());
        B200_LAUNCHED();
        int found = 0;
        B200_CUDA(cudaMemcpyAsync(&found, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        if(!found)
            return aoclsparse_status_invalid_index_value;
        // derived copies hold the old value (the reference drops them too, auxiliary.hpp:463-471)
        drop_derived_copies(A);
        return aoclsparse_status_success;
    }
}

using namespace b200;

extern "C" {

aoclsparse_status aoclsparse_sdotmv(const aoclsparse_operation op, const float alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const float *x, const float beta, float *y, float *d)
{
    return dotmv_entry<float>(op, alpha, A, descr, x, beta, y, d);
}
aoclsparse_status aoclsparse_ddotmv(const aoclsparse_operation op, const double alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const double *x, const double beta, double *y, double *d)
{
    return dotmv_entry<double>(op, alpha, A, descr, x, beta, y, d);
}
aoclsparse_status aoclsparse_cdotmv(const aoclsparse_operation op, const aoclsparse_float_complex alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const aoclsparse_float_complex *x, const aoclsparse_float_complex beta, aoclsparse_float_complex *y, aoclsparse_float_complex *d)
{
    return dotmv_entry<float2>(op, make_float2(alpha.real, alpha.imag), A, descr, reinterpret_cast<const float2 *>(x), make_float2(beta.real, beta.imag), reinterpret_cast<float2 *>(y), reinterpret_cast<float2 *>(d));
}
aoclsparse_status aoclsparse_zdotmv(const aoclsparse_operation op, const aoclsparse_double_complex alpha, aoclsparse_matrix A, const aoclsparse_mat_descr descr, const aoclsparse_double_complex *x, const aoclsparse_double_complex beta, aoclsparse_double_complex *y, aoclsparse_double_complex *d)
{
    return dotmv_entry<double2>(op, make_double2(alpha.real, alpha.imag), A, descr, reinterpret_cast<const double2 *>(x), make_double2(beta.real, beta.imag), reinterpret_cast<double2 *>(y), reinterpret_cast<double2 *>(d));
}
aoclsparse_status aoclsparse_sset_value(aoclsparse_matrix A, aoclsparse_int row_idx, aoclsparse_int col_idx, float val)
{
    return set_value_entry<float>(A, row_idx, col_idx, val);
}
aoclsparse_status aoclsparse_dset_value(aoclsparse_matrix A, aoclsparse_int row_idx, aoclsparse_int col_idx, double val)
{
    return set_value_entry<double>(A, row_idx, col_idx, val);
}
aoclsparse_status aoclsparse_cset_value(aoclsparse_matrix A, aoclsparse_int row_idx, aoclsparse_int col_idx, aoclsparse_float_complex val)
{
    return set_value_entry<float2>(A, row_idx, col_idx, make_float2(val.real, val.imag));
}
aoclsparse_status aoclsparse_zset_value(aoclsparse_matrix A, aoclsparse_int row_idx, aoclsparse_int col_idx, aoclsparse_double_complex val)
{
    return set_value_entry<double2>(A, row_idx, col_idx, make_double2(val.real, val.imag));
}

aoclsparse_status aoclsparse_scsrmv(aoclsparse_operation       trans,
                                    const float               *alpha,
                                    aoclsparse_int             m,
                                    aoclsparse_int             n,
                                    aoclsparse_int             nnz,
                                    const float               *csr_val,
                                    const aoclsparse_int      *csr_col_ind,
                                    const aoclsparse_int      *csr_row_ptr,
                                    const aoclsparse_mat_descr descr,
                                    const float               *x,
                                    const float               *beta,
                                    float                     *y)
{
    return csrmv_legacy<float>(trans, alpha, m, n, nnz, csr_val, csr_col_ind, csr_row_ptr, descr, x, beta, y);
}

aoclsparse_status aoclsparse_dcsrmv(aoclsparse_operation       trans,
                                    const double              *alpha,
                                    aoclsparse_int             m,
                                    aoclsparse_int             n,
                                    aoclsparse_int             nnz,
                                    const double              *csr_val,
                                    const aoclsparse_int      *csr_col_ind,
                                    const aoclsparse_int      *csr_row_ptr,
                                    const aoclsparse_mat_descr descr,
                                    const double              *x,
                                    const double              *beta,
                                    double                    *y)
{
    return csrmv_legacy<double>(trans, alpha, m, n, nnz, csr_val, csr_col_ind, csr_row_ptr, descr, x, beta, y);
}

aoclsparse_status aoclsparse_smv(aoclsparse_operation       op,
                                 const float               *alpha,
                                 aoclsparse_matrix          A,
                                 const aoclsparse_mat_descr descr,
                                 const float               *x,
                                 const float               *beta,
                                 float                     *y)
{
    return mv_entry<float>(op, alpha, A, descr, x, beta, y);
}

aoclsparse_status aoclsparse_dmv(aoclsparse_operation       op,
                                 const double              *alpha,
                                 aoclsparse_matrix          A,
                                 const aoclsparse_mat_descr descr,
                                 const double              *x,
                                 const double              *beta,
                                 double                    *y)
{
    return mv_entry<double>(op, alpha, A, descr, x, beta, y);
}

aoclsparse_status aoclsparse_cmv(aoclsparse_operation            op,
                                 const aoclsparse_float_complex *alpha,
                                 aoclsparse_matrix               A,
                                 const aoclsparse_mat_descr      descr,
                                 const aoclsparse_float_complex *x,
                                 const aoclsparse_float_complex *beta,
                                 aoclsparse_float_complex       *y)
{
    return mv_entry<float2>(op,
                            reinterpret_cast<const float2 *>(alpha),
                            A,
                            descr,
                            reinterpret_cast<const float2 *>(x),
                            reinterpret_cast<const float2 *>(beta),
                            reinterpret_cast<float2 *>(y));
}

aoclsparse_status aoclsparse_zmv(aoclsparse_operation             op,
                                 const aoclsparse_double_complex *alpha,
                                 aoclsparse_matrix                A,
                                 const aoclsparse_mat_descr       descr,
                                 const aoclsparse_double_complex *x,
                                 const aoclsparse_double_complex *beta,
                                 aoclsparse_double_complex       *y)
{
    return mv_entry<double2>(op,
                             reinterpret_cast<const double2 *>(alpha),
                             A,
                             descr,
                             reinterpret_cast<const double2 *>(x),
                             reinterpret_cast<const double2 *>(beta),
                             reinterpret_cast<double2 *>(y));
}

aoclsparse_status aoclsparse_b200_dmv_rows(const double              *alpha,
                                           aoclsparse_matrix          A,
                                           const aoclsparse_mat_descr descr,
                                           const double              *x,
                                           const double              *beta,
                                           double                    *y,
                                           aoclsparse_int             row_begin,
                                           aoclsparse_int             row_end)
{
    return mv_rows_entry<double>(alpha, A, descr, x, beta, y, row_begin, row_end);
}

aoclsparse_status aoclsparse_b200_smv_rows(const float               *alpha,
                                           aoclsparse_matrix          A,
                                           const aoclsparse_mat_descr descr,
                                           const float               *x,
                                           const float               *beta,
                                           float                     *y,
                                           aoclsparse_int             row_begin,
                                           aoclsparse_int             row_end)
{
    return mv_rows_entry<float>(alpha, A, descr, x, beta, y, row_begin, row_end);
}

aoclsparse_status aoclsparse_b200_dmv_rows_push(const double              *alpha,
                                                aoclsparse_matrix          A,
                                                const aoclsparse_mat_descr descr,
                                                const double              *x,
                                                const double              *beta,
                                                double                    *y,
                                                aoclsparse_int             row_begin,
                                                aoclsparse_int             row_end,
                                                double                    *push_dst)
{
    if(!push_dst)
        return aoclsparse_status_invalid_pointer;
    return mv_rows_entry<double>(alpha, A, descr, x, beta, y, row_begin, row_end, push_dst);
}

aoclsparse_status aoclsparse_b200_signal(void *flag, unsigned value)
{
    if(!flag)
        return aoclsparse_status_invalid_pointer;
    signal_flag_kernel<<<1, 1, 0, current_stream()>>>(static_cast<volatile unsigned *>(flag), value);
    B200_LAUNCHED();
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_wait(const void *flag, unsigned value, unsigned *timed_out)
{
    if(!flag)
        return aoclsparse_status_invalid_pointer;
    wait_flag_kernel<<<1, 1, 0, current_stream()>>>(static_cast<const volatile unsigned *>(flag), value, timed_out);
    B200_LAUNCHED();
    return aoclsparse_status_success;
}

// device memory that other processes on this node can map (cudaIpc*): the x windows and flags of the
// row-sharded iteration live in such buffers so that neighbours can store into them over NVLink
aoclsparse_status aoclsparse_b200_ipc_alloc(size_t bytes, void **dptr, unsigned char handle[64])
{
    if(!dptr || !handle)
        return aoclsparse_status_invalid_pointer;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    B200_CUDA(cudaMalloc(dptr, bytes));
    B200_CUDA(cudaMemset(*dptr, 0, bytes));
    cudaIpcMemHandle_t h;
    B200_CUDA(cudaIpcGetMemHandle(&h, *dptr));
    memcpy(handle, &h, 64);
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_ipc_open(const unsigned char handle[64], void **dptr)
{
    if(!dptr || !handle)
        return aoclsparse_status_invalid_pointer;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    B200_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_memcpy(void *dst, const void *src, size_t bytes)
{
    if(!dst || !src)
        return aoclsparse_status_invalid_pointer;
    B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, current_stream()));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_ipc_close(void *dptr)
{
    if(dptr)
        B200_CUDA(cudaIpcCloseMemHandle(dptr));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_ipc_free(void *dptr)
{
    if(dptr)
        B200_CUDA(cudaFree(dptr));
    return aoclsparse_status_success;
}

}

// AOCLSPARSE_B200_SHARD_FENCE: 1 (default) boundary CTAs release at GPU scope and only the last one of a side fences at
// system scope; 0 every boundary CTA fences at system scope (spmv_sharded.cuh, boundary_release)
static int shard_cta_fence_gpu()
{
    static const int v = [] {
        const char *e = getenv("AOCLSPARSE_B200_SHARD_FENCE");
        return e ? (atoi(e) != 0 ? 1 : 0) : 1;
    }();
    return v;
}

// One launch per iteration of the row-sharded product: multiply + halo push + flags (spmv_sharded.cuh).
// kc = launches of this kernel on these counters so far, this one included (the flags carry ctl->k, which may run ahead
// of kc when other events -- the initial halo publication of shard.cu -- take a number as well)
aoclsparse_status b200::sharded_step_launch(const double                  *alpha,
                                            aoclsparse_matrix              A,
                                            const aoclsparse_mat_descr     descr,
                                            const double                  *x,
                                            double                        *y,
                                            const aoclsparse_b200_halo_ctl *ctl,
                                            unsigned                        kc)
{
    if(!alpha || !A || !descr || !x || !y || !ctl || !ctl->counters)
        return aoclsparse_status_invalid_pointer;
    if(A->mats.empty() || A->mats[0] == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(descr->base != A->base)
        return aoclsparse_status_invalid_value;
    if(A->val_type != aoclsparse_dmat)
        return aoclsparse_status_wrong_type;
    if(descr->type != aoclsparse_matrix_type_general || ctl->k == 0)
        return aoclsparse_status_invalid_value;
    if(A->is_csc)
        return aoclsparse_status_not_implemented;
    cudaStream_t st = current_stream();
    B200_TRY(ensure_plan(A, st));
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const dev_csr                      &M = *A->mats[0];
    const row_block_plan               &P0 = M.plan;
    // entry-code copy: the kernel walks the block plan built for it (same cuts)
    const bool                          ec = P0.n_codes > 0 && P0.n_ecodes > 0 && !P0.ecodes_stale && P0.eplan && P0.eplan->cut_block.size() == 2;
    const row_block_plan               &P  = ec ? *P0.eplan : P0;
    // needs exactly the cuts [h, m-h] and a plan whose blocks are all thread-per-row
    if(A->row_cuts.size() != 2 || P.cut_block.size() != 2 || P.n_strat[STRAT_THREAD] != P.n_blocks)
        return aoclsparse_status_not_implemented;
    halo_ctl hc;
    hc.left_done     = static_cast<const unsigned *>(ctl->left_done);
    hc.right_done    = static_cast<const unsigned *>(ctl->right_done);
    hc.to_left_done  = static_cast<unsigned *>(ctl->to_left_done);
    hc.to_right_done = static_cast<unsigned *>(ctl->to_right_done);
    hc.counters      = static_cast<unsigned *>(ctl->counters);
    hc.k                 = ctl->k;
    hc.kc                = kc;
    hc.n_first           = P.cut_block[0];
    hc.last_begin        = P.cut_block[1];
    hc.n_last            = P.n_blocks - P.cut_block[1];
    hc.n_blocks          = P.n_blocks;
    hc.first_rows        = A->row_cuts[0];
    hc.last_row0         = A->row_cuts[1];
    // the window holds a halo of row_cuts[0] entries on every side that has a neighbour
    hc.own_lo = (A->win_hi >= 0 ? (int)A->win_lo : 0) + (hc.left_done ? (int)A->row_cuts[0] : 0);
    hc.own_hi = hc.own_lo + (int)A->m;
    hc.cta_fence_gpu = shard_cta_fence_gpu();
    if((hc.left_done && !ctl->push_left) || (hc.right_done && !ctl->push_right))
        return aoclsparse_status_invalid_pointer;
    if(A->win_hi >= 0)
        x = x - A->win_lo;
    const bool   coded = P0.n_codes > 0 && !ec;
    const aoclsparse_int e_nnz = P.max_block_nnz > 0 ? P.max_block_nnz : P.block_nnz; // entry codes: the largest block sizes the buffer
    const int    cap   = ec ? spmv_ecoded_cap(e_nnz) : P.block_nnz + (coded ? 32 : 8);
    const size_t smem  = ec ? spmv_ecoded_smem_bytes(sizeof(double), e_nnz)
                            : (coded ? spmv_coded_smem_bytes(sizeof(double), P.block_nnz) : spmv_smem_bytes(sizeof(double), P.block_nnz));
    auto         kern  = ec ? spmv_sharded_step_kernel<double, false, true>
                            : (coded ? spmv_sharded_step_kernel<double, true, false> : spmv_sharded_step_kernel<double, false, false>);
    static std::atomic<size_t> configured[3] = {{0}, {0}, {0}};
    const int                  variant       = ec ? 2 : (coded ? 1 : 0);
    if(configured[variant].load() < smem)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[variant].store(smem);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3((unsigned)P.n_blocks);
    cfg.blockDim           = dim3(256);
    cfg.dynamicSmemBytes   = smem;
    cfg.stream             = st;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    // more than one wave only: see launch_row_blocks
    cfg.numAttrs = (P.pdl && (long long)P.n_blocks > resident_ctas(kern, 256, smem)) ? 1 : 0;
    B200_CUDA(cudaLaunchKernelEx(&cfg,
                                 kern,
                                 (const int4 *)P.desc.as<int4>(),
                                 cap,
                                 (const aoclsparse_int *)M.row_ptr.as<aoclsparse_int>(),
                                 (const aoclsparse_int *)M.col_idx.as<aoclsparse_int>(),
                                 (const double *)M.val.as<double>(),
                                 x,
                                 y,
                                 *alpha,
                                 static_cast<double *>(ctl->push_left),
                                 static_cast<double *>(ctl->push_right),
                                 hc,
                                 (const unsigned char *)(ec ? P0.ecodes.as<unsigned char>() : P0.codes.as<unsigned char>()),
                                 (const int *)(ec ? P0.etab_off.as<int>() : P0.code_offsets.as<int>()),
                                 (const double *)P0.etab_val.as<double>(),
                                 (int)(ec ? P0.n_ecodes : P0.n_codes)));
    B200_LAUNCHED();
    return aoclsparse_status_success;
}

// k iterations of the row-sharded product in ONE cooperative launch (spmv_sharded_iterate_kernel).  w_cur / w_nxt are
// the two x windows (the current iterate lives in w_cur); push_*[i] are the neighbours' halo slots of window i
// (0 = the window w_cur maps to in the neighbour, 1 = the other one).  *grid_out = CTAs launched (the grid barrier
// counter advances by (iterations - 1) * grid).
// true when the shard's multiply runs on the entry-code copy: its iterations are launched one step kernel each
// (see the note at spmv_sharded_iterate_kernel)
bool b200::sharded_prefers_steps(aoclsparse_matrix A)
{
    if(!A || A->mats.empty() || A->mats[0] == nullptr)
        return true;
    if(ensure_plan(A, current_stream()) != aoclsparse_status_success)
        return true;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const row_block_plan               &P0 = A->mats[0]->plan;
    return P0.n_codes > 0 && P0.n_ecodes > 0 && !P0.ecodes_stale && P0.eplan && P0.eplan->cut_block.size() == 2;
}

aoclsparse_status b200::sharded_iterate_launch(double                     alpha,
                                               aoclsparse_matrix          A,
                                               const aoclsparse_mat_descr descr,
                                               double                    *w_cur,
                                               double                    *w_nxt,
                                               long long                  own_offset,
                                               const sharded_iterate_args &args,
                                               int                        iterations,
                                               int                       *grid_out)
{
    if(!A || !descr || !w_cur || !w_nxt || !args.counters || !grid_out)
        return aoclsparse_status_invalid_pointer;
    if(A->mats.empty() || A->mats[0] == nullptr)
        return aoclsparse_status_invalid_pointer;
    if(descr->base != A->base)
        return aoclsparse_status_invalid_value;
    if(A->val_type != aoclsparse_dmat)
        return aoclsparse_status_wrong_type;
    if(descr->type != aoclsparse_matrix_type_general || args.k0 == 0 || iterations < 1)
        return aoclsparse_status_invalid_value;
    if(A->is_csc)
        return aoclsparse_status_not_implemented;
    cudaStream_t st = current_stream();
    B200_TRY(ensure_plan(A, st));
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const dev_csr                      &M = *A->mats[0];
    const row_block_plan               &P0 = M.plan;
    // (entry-coded shards run one step kernel per iteration: sharded_prefers_steps)
    const row_block_plan               &P  = P0;
    if(A->row_cuts.size() != 2 || P.cut_block.size() != 2 || P.n_strat[STRAT_THREAD] != P.n_blocks || P.n_blocks < 1)
        return aoclsparse_status_not_implemented;
    iterate_ctl hc;
    hc.left_done     = static_cast<const unsigned *>(args.left_done);
    hc.right_done    = static_cast<const unsigned *>(args.right_done);
    hc.to_left_done  = static_cast<unsigned *>(args.to_left_done);
    hc.to_right_done = static_cast<unsigned *>(args.to_right_done);
    hc.counters      = static_cast<unsigned *>(args.counters);
    hc.k0            = args.k0;
    hc.kc0           = args.kc0;
    hc.bar0          = args.bar0;
    hc.iters         = iterations;
    hc.n_first       = P.cut_block[0];
    hc.last_begin    = P.cut_block[1];
    hc.n_last        = P.n_blocks - P.cut_block[1];
    hc.n_blocks      = P.n_blocks;
    hc.last_row0     = A->row_cuts[1];
    hc.own_lo        = (A->win_hi >= 0 ? (int)A->win_lo : 0) + (hc.left_done ? (int)A->row_cuts[0] : 0);
    hc.own_hi        = hc.own_lo + (int)A->m;
    hc.cta_fence_gpu = shard_cta_fence_gpu();
    if((hc.left_done && (!args.push_left[0] || !args.push_left[1])) || (hc.right_done && (!args.push_right[0] || !args.push_right[1])))
        return aoclsparse_status_invalid_pointer;
    const long long shift = A->win_hi >= 0 ? (long long)A->win_lo : 0;
    // iteration 0 reads w_cur and writes the own rows of w_nxt; the neighbours receive into THEIR w_nxt (slot 1)
    const double *x0 = w_cur - shift, *x1 = w_nxt - shift;
    double       *y0 = w_nxt + own_offset, *y1 = w_cur + own_offset;
    const bool   coded = P0.n_codes > 0;
    int          cap   = P.block_nnz + (coded ? 32 : 8);
    const size_t smem  = coded ? spmv_coded_smem_bytes(sizeof(double), P.block_nnz) : spmv_smem_bytes(sizeof(double), P.block_nnz);
    auto         kern  = coded ? spmv_sharded_iterate_kernel<double, true> : spmv_sharded_iterate_kernel<double, false>;
    static std::atomic<size_t> configured[2] = {{0}, {0}};
    const int                  variant       = coded ? 1 : 0;
    if(configured[variant].load() < smem)
    {
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[variant].store(smem);
    }
    long long grid = resident_ctas(kern, 256, smem);
    if(const char *e = getenv("AOCLSPARSE_B200_ITER_GRID"))
    {
        const long long v = atoll(e);
        if(v > 0 && v < grid)
            grid = v;
    }
    if(grid > P.n_blocks)
        grid = P.n_blocks;
    // every CTA walks blocks c, c+G, ...: with G = ceil(blocks / rounds) all CTAs do the same number of rounds (the last
    // one short by < rounds blocks in total) instead of some doing a whole extra block while the others wait at the barrier
    {
        const long long rounds = (P.n_blocks + grid - 1) / grid;
        grid                   = (P.n_blocks + rounds - 1) / rounds;
    }
    const int4           *p_desc  = P.desc.as<int4>();
    const aoclsparse_int *p_rp    = M.row_ptr.as<aoclsparse_int>();
    const aoclsparse_int *p_col   = M.col_idx.as<aoclsparse_int>();
    const double         *p_val   = M.val.as<double>();
    double               *pl0     = static_cast<double *>(args.push_left[1]); // iteration 0 stores into the neighbours' "next" window
    double               *pl1     = static_cast<double *>(args.push_left[0]);
    double               *pr0     = static_cast<double *>(args.push_right[1]);
    double               *pr1     = static_cast<double *>(args.push_right[0]);
    const unsigned char  *p_codes = P0.codes.as<unsigned char>();
    const int            *p_off   = P0.code_offsets.as<int>();
    int                   n_table = (int)P0.n_codes;
    void *params[] = {&p_desc, &cap, &p_rp, &p_col, &p_val, &x0, &x1, &y0, &y1, &alpha, &pl0, &pl1, &pr0, &pr1, &hc, &p_codes, &p_off, &n_table};
    B200_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3((unsigned)grid), dim3(256), params, smem, st));
    B200_LAUNCHED();
    *grid_out = (int)grid;
    return aoclsparse_status_success;
}

extern "C" {
aoclsparse_status aoclsparse_b200_dmv_sharded_step(const double                  *alpha,
                                                   aoclsparse_matrix              A,
                                                   const aoclsparse_mat_descr     descr,
                                                   const double                  *x,
                                                   double                        *y,
                                                   const aoclsparse_b200_halo_ctl *ctl)
{
    if(!ctl)
        return aoclsparse_status_invalid_pointer;
    return b200::sharded_step_launch(alpha, A, descr, x, y, ctl, ctl->k);
}
}
