// itsol.cu -- conjugate gradients on the device: the consumer of "SpMV x 100" in the reference.
//
// Reference: the iterative-solver suite of library/src/solvers/aoclsparse_itsol_functions.{hpp,cpp}: a handle with an
// option registry (aoclsparse_itsol_list_options.hpp:63-239), a reverse-communication CG state machine
// (aoclsparse_cg_rci_solve, itsol_functions.hpp:633-870) and a forward interface that drives it with aoclsparse::mv on a
// symmetric / lower descriptor (aoclsparse_cg_solve, :1369-1500).  Kept here: aoclsparse_itsol_{s,d,c,z}_init,
// aoclsparse_itsol_destroy, aoclsparse_itsol_option_set (all registered options, same names, bounds, defaults and
// normalisation), aoclsparse_itsol_{s,d,c,z}_rci_input, aoclsparse_itsol_{s,d,c,z}_rci_solve and
// aoclsparse_itsol_{s,d,c,z}_solve for the CG method with no or a user preconditioner (complex handles: the section
// "complex conjugate gradients" below).  GMRES and the symmetric Gauss-Seidel preconditioner return
// aoclsparse_status_not_implemented.
//
// B200: the state machine and its scalar logic (tolerances, iteration limit, breakdown tests, rinfo) are the reference's;
// every vector operation is a CUDA kernel on vectors that never leave the device in the forward interface:
//   r = -b, p = x                      cg_start_kernel
//   r += q, |r|^2                      cg_residual_kernel        (block partials in double; the last block adds them in
//                                                                 index order and stores the sum into mapped host memory)
//   r.z , p.q                          dot_kernel
//   p = beta p - z                     cg_direction_kernel
//   x += alpha p, r += alpha q, |r|^2  cg_step_kernel
// The work vectors r, p, q, z live in CUDA managed memory: the reverse-communication interface hands pointers to them to
// the caller (*u, *v), and a host caller may read / write them as ordinary memory while a GPU caller passes them straight
// to aoclsparse_?mv; in the forward interface nothing touches them from the host, so they stay resident in HBM.
// Two scalars per iteration travel to the host (p.q and |r|), because the reference tests convergence and breakdown
// every iteration and reports them through rinfo.
#include "common.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <complex>
#include <limits>
#include <string>

namespace b200
{
    namespace
    {
        constexpr int RED_THREADS = 256, RED_BLOCKS = 148 * 4;

        enum cg_task
        {
            task_start = 0,
            task_init_res,
            task_check_conv,
            task_start_iter,
            task_compute_beta,
            task_take_step
        };

        // ---------------------------------------------------------------- options (itsol_list_options.hpp:63-239)
        template <typename R>
        R expected_precision(R scale)
        {
            // aoclsparse_utils.hpp:560-580: scale * safeguard * sqrt(2 eps), safeguard = 1 (double) / 2 (float)
            const R eps = std::numeric_limits<R>::epsilon();
            return scale * (sizeof(R) == 4 ? (R)2.0 : (R)1.0) * std::sqrt((R)2.0 * eps);
        }

        // trim, squeeze runs of blanks to one space, lower case (OptionUtility::PrepareString, itsol_options.hpp:104-114)
        std::string prepare(const char *s)
        {
            std::string out;
            bool        blank = false;
            for(const char *p = s; *p; ++p)
            {
                if(std::isspace((unsigned char)*p))
                {
                    blank = !out.empty();
                    continue;
                }
                if(blank)
                    out.push_back(' ');
                blank = false;
                out.push_back((char)std::tolower((unsigned char)*p));
            }
            return out;
        }

        template <typename R>
        struct options
        {
            int  solver        = 0; // 0 CG, 1 GMRES
            int  cg_maxit      = 500;
            R    cg_rtol       = expected_precision<R>((R)2.0);
            R    cg_atol       = expected_precision<R>((R)1.0);
            int  cg_precond    = 0; // 0 none, 1 user, 3 symmetric Gauss-Seidel
            int  gm_maxit      = 150;
            R    gm_rtol       = expected_precision<R>((R)2.0);
            R    gm_atol       = expected_precision<R>((R)1.0);
            int  gm_precond    = 0;
            int  gm_restart    = 20;
            bool locked        = false;

            // the registered options with their current values (aoclsparse_itsol_handle_prn_options)
            void print() const
            {
                static const char *methods[] = {"cg", "gmres"};
                static const char *pre[]     = {"none", "user", "gs", "sym gs", "sgs", "ilu0"};
                printf("Begin Options\n");
                printf("   cg iteration limit            = %d\n", cg_maxit);
                printf("   cg rel tolerance              = %.6e\n", (double)cg_rtol);
                printf("   cg abs tolerance              = %.6e\n", (double)cg_atol);
                printf("   cg preconditioner             = %s\n", pre[cg_precond >= 0 && cg_precond < 6 ? cg_precond : 0]);
                printf("   gmres iteration limit         = %d\n", gm_maxit);
                printf("   gmres rel tolerance           = %.6e\n", (double)gm_rtol);
                printf("   gmres abs tolerance           = %.6e\n", (double)gm_atol);
                printf("   gmres restart iterations      = %d\n", gm_restart);
                printf("   gmres preconditioner          = %s\n", pre[gm_precond >= 0 && gm_precond < 6 ? gm_precond : 0]);
                printf("   iterative method              = %s\n", methods[solver == 1 ? 1 : 0]);
                printf("End Options\n");
            }

            // 0 ok, 1 out of range, 2 bad value, 3 unknown option, 4 locked (OptionRegistry::SetOption return codes)
            int set(const std::string &name, const char *raw)
            {
                int *ip = nullptr;
                R   *rp = nullptr;
                if(name == "cg iteration limit")
                    ip = &cg_maxit;
                else if(name == "gmres iteration limit")
                    ip = &gm_maxit;
                else if(name == "gmres restart iterations")
                    ip = &gm_restart;
                else if(name == "cg rel tolerance")
                    rp = &cg_rtol;
                else if(name == "cg abs tolerance")
                    rp = &cg_atol;
                else if(name == "gmres rel tolerance")
                    rp = &gm_rtol;
                else if(name == "gmres abs tolerance")
                    rp = &gm_atol;
                if(ip || rp)
                {
                    if(locked)
                        return 4;
                    char *end = nullptr;
                    if(ip)
                    {
                        const long v = std::strtol(raw, &end, 10);
                        if(end == raw)
                            return 2;
                        if(v < 1 || v > 0x7fffffffL)
                            return 1;
                        *ip = (int)v;
                    }
                    else
                    {
                        const double v = std::strtod(raw, &end);
                        if(end == raw)
                            return 2;
                        if(!(v >= 0.0))
                            return 1;
                        *rp = (R)v;
                    }
                    return 0;
                }
                const std::string v = prepare(raw);
                if(name == "iterative method")
                {
                    if(locked)
                        return 4;
                    if(v == "cg" || v == "pcg")
                        solver = 0;
                    else if(v == "gmres" || v == "gm res")
                        solver = 1;
                    else
                        return 2;
                    return 0;
                }
                if(name == "cg preconditioner")
                {
                    if(locked)
                        return 4;
                    if(v == "none")
                        cg_precond = 0;
                    else if(v == "user")
                        cg_precond = 1;
                    else if(v == "gs" || v == "symgs" || v == "sgs")
                        cg_precond = 3;
                    else
                        return 2;
                    return 0;
                }
                if(name == "gmres preconditioner")
                {
                    if(locked)
                        return 4;
                    if(v == "none")
                        gm_precond = 0;
                    else if(v == "user")
                        gm_precond = 1;
                    else if(v == "ilu0")
                        gm_precond = 2;
                    else
                        return 2;
                    return 0;
                }
                return 3;
            }
        };

        // ---------------------------------------------------------------- vector kernels
        // where a reduction lands: per-block partials, a ticket counter, and the final value in page-locked host memory
        // that the device writes directly (one 8-byte store instead of a finishing launch + a memcpy)
        struct reduce_out
        {
            double   *partial;
            unsigned *ticket;
            double   *result; // device address of mapped host memory
        };

        // block sum -> partial[blockIdx.x]; the LAST block to arrive adds the partials in index order (deterministic)
        template <typename T>
        __device__ __forceinline__ void block_sum_store(double s, reduce_out o)
        {
            __shared__ double sh[RED_THREADS / 32];
            __shared__ bool   last;
#pragma unroll
            for(int k = 16; k > 0; k >>= 1)
                s += __shfl_down_sync(0xffffffffu, s, k);
            if((threadIdx.x & 31) == 0)
                sh[threadIdx.x >> 5] = s;
            __syncthreads();
            if(threadIdx.x < 32)
            {
                s = threadIdx.x < RED_THREADS / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
                for(int k = 16; k > 0; k >>= 1)
                    s += __shfl_down_sync(0xffffffffu, s, k);
                if(threadIdx.x == 0)
                {
                    o.partial[blockIdx.x] = s;
                    __threadfence();
                    last = atomicAdd(o.ticket, 1u) == gridDim.x - 1;
                }
            }
            __syncthreads();
            if(!last)
                return;
            __threadfence();
            double t = 0;
            for(int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS)
                t += __ldcg(o.partial + i);
            __shared__ double fin[RED_THREADS];
            fin[threadIdx.x] = t;
            __syncthreads();
            for(int k = RED_THREADS / 2; k > 0; k >>= 1)
            {
                if(threadIdx.x < k)
                    fin[threadIdx.x] += fin[threadIdx.x + k];
                __syncthreads();
            }
            if(threadIdx.x == 0)
            {
                *o.result = fin[0];
                *o.ticket = 0; // ready for the next reduction on this stream
                __threadfence_system();
            }
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cg_start_kernel(long long n, const T *__restrict__ b, const T *__restrict__ x,
                                                                      T *__restrict__ r, T *__restrict__ p, reduce_out partial)
        {
            double s = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
            {
                const T bi = b[i];
                r[i]       = -bi;
                p[i]       = x[i];
                s += (double)bi * (double)bi;
            }
            block_sum_store<T>(s, partial);
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cg_residual_kernel(long long n, T *__restrict__ r, const T *__restrict__ q,
                                                                         T *__restrict__ p, reduce_out partial)
        {
            double s = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
            {
                const T ri = r[i] + q[i];
                r[i]       = ri;
                p[i]       = (T)0;
                s += (double)ri * (double)ri;
            }
            block_sum_store<T>(s, partial);
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) dot_kernel(long long n, const T *__restrict__ a, const T *__restrict__ b, reduce_out partial)
        {
            double s = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
                s += (double)a[i] * (double)b[i];
            block_sum_store<T>(s, partial);
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cg_direction_kernel(long long n, T beta, T *__restrict__ p, const T *__restrict__ z)
        {
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
                p[i] = beta * p[i] - z[i];
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cg_step_kernel(long long n, T alpha, const T *__restrict__ p, const T *__restrict__ q,
                                                                     T *__restrict__ x, T *__restrict__ r, reduce_out partial)
        {
            double s = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
            {
                x[i] += alpha * p[i];
                const T ri = r[i] + alpha * q[i];
                r[i]       = ri;
                s += (double)ri * (double)ri;
            }
            block_sum_store<T>(s, partial);
        }

        // ---------------------------------------------------------------- device-driven solve (forward interface, no callbacks)
        // The scalar logic of the state machine moves to the device: the LAST block of every reduction kernel runs it
        // (same comparisons, same order, same precision as cg_rci below), so an iteration needs no host round trip; the
        // host enqueues iterations in small batches and looks at `status` after each batch.  Kernels of iterations that
        // come after the stopping one see status != 0 and do nothing (the products in between are wasted, never used).
        template <typename T>
        struct cg_dev_state
        {
            T   rz, alpha, beta, rnorm2, bnorm2, brtol, atol, rtol;
            int niter, maxit;
            int status; // 0 running, 1 converged, 2 iteration limit, 3 breakdown, 4 NaN in b, 5 NaN residual
        };

        template <typename T>
        __device__ __forceinline__ bool dev_nearzero_or_negative(T v)
        {
            return v <= (T)1e-2 * (T)2.0 * (sizeof(T) == 4 ? (T)1.1920928955078125e-7 : (T)2.220446049250313e-16);
        }

        // task_check_conv + task_start_iter + the scalar part of task_compute_beta (unpreconditioned: z = r, r.z = |r|^2)
        template <typename T>
        __device__ __forceinline__ void cg_check_advance(cg_dev_state<T> *st)
        {
            if((T)0 < st->atol && st->rnorm2 <= st->atol)
                st->status = 1;
            else if((T)0 < st->rtol && st->rnorm2 <= st->brtol)
                st->status = 1;
            else if(st->maxit > 0 && st->niter > st->maxit)
                st->status = 2;
            else
            {
                st->niter++;
                const T rz_new = st->rnorm2 * st->rnorm2;
                if(dev_nearzero_or_negative(st->rz))
                    st->status = 3;
                else
                {
                    st->beta = rz_new / st->rz;
                    st->rz   = rz_new;
                }
            }
        }

        enum
        {
            EPI_START = 0, // sum = |b|^2
            EPI_RESID,     // sum = |r0|^2
            EPI_PQ,        // sum = p.q
            EPI_STEP       // sum = |r|^2 after the update
        };

        template <typename T, int EPI>
        __device__ __forceinline__ void cg_epilogue(cg_dev_state<T> *st, double sum)
        {
            if(EPI == EPI_START)
            {
                st->bnorm2 = (T)sqrt(sum);
                if(st->bnorm2 != st->bnorm2)
                    st->status = 4;
                st->brtol = st->rtol * st->bnorm2;
                st->niter = 0;
            }
            else if(EPI == EPI_RESID || EPI == EPI_STEP)
            {
                st->rnorm2 = (T)sqrt(sum);
                if(st->rnorm2 != st->rnorm2)
                {
                    st->status = 5;
                    return;
                }
                if(EPI == EPI_RESID)
                    st->rz = (T)1;
                cg_check_advance(st);
            }
            else
            {
                const T pq = (T)sum;
                if(dev_nearzero_or_negative(pq) || pq == (T)0)
                    st->status = 3; // A is not positive definite
                else
                    st->alpha = st->rz / pq;
            }
        }

        // block partial -> last block: ordered sum -> scalar logic on the device
        template <typename T, int EPI>
        __device__ __forceinline__ void block_sum_epilogue(double s, double *partial, unsigned *ticket, cg_dev_state<T> *st)
        {
            __shared__ double sh[RED_THREADS / 32];
            __shared__ bool   last;
            __shared__ double fin[RED_THREADS];
#pragma unroll
            for(int k = 16; k > 0; k >>= 1)
                s += __shfl_down_sync(0xffffffffu, s, k);
            if((threadIdx.x & 31) == 0)
                sh[threadIdx.x >> 5] = s;
            __syncthreads();
            if(threadIdx.x < 32)
            {
                s = threadIdx.x < RED_THREADS / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
                for(int k = 16; k > 0; k >>= 1)
                    s += __shfl_down_sync(0xffffffffu, s, k);
                if(threadIdx.x == 0)
                {
                    partial[blockIdx.x] = s;
                    __threadfence();
                    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
                }
            }
            __syncthreads();
            if(!last)
                return;
            __threadfence();
            double t = 0;
            for(int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS)
                t += __ldcg(partial + i);
            fin[threadIdx.x] = t;
            __syncthreads();
            for(int k = RED_THREADS / 2; k > 0; k >>= 1)
            {
                if(threadIdx.x < k)
                    fin[threadIdx.x] += fin[threadIdx.x + k];
                __syncthreads();
            }
            if(threadIdx.x == 0)
            {
                cg_epilogue<T, EPI>(st, fin[0]);
                *ticket = 0;
                __threadfence();
            }
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cgd_start_kernel(long long n, const T *__restrict__ b, const T *__restrict__ x,
                                                                       T *__restrict__ r, T *__restrict__ p, double *partial,
                                                                       unsigned *ticket, cg_dev_state<T> *st)
        {
            double s = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
            {
                const T bi = b[i];
                r[i]       = -bi;
                p[i]       = x[i];
                s += (double)bi * (double)bi;
            }
            block_sum_epilogue<T, EPI_START>(s, partial, ticket, st);
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cgd_residual_kernel(long long n, T *__restrict__ r, const T *__restrict__ q,
                                                                          T *__restrict__ p, double *partial, unsigned *ticket,
                                                                          cg_dev_state<T> *st)
        {
            if(st->status)
                return;
            double s = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
            {
                const T ri = r[i] + q[i];
                r[i]       = ri;
                p[i]       = (T)0;
                s += (double)ri * (double)ri;
            }
            block_sum_epilogue<T, EPI_RESID>(s, partial, ticket, st);
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cgd_direction_kernel(long long n, T *__restrict__ p, const T *__restrict__ r,
                                                                           const cg_dev_state<T> *st)
        {
            if(st->status)
                return;
            const T beta = st->beta;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
                p[i] = beta * p[i] - r[i];
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cgd_dot_kernel(long long n, const T *__restrict__ p, const T *__restrict__ q,
                                                                     double *partial, unsigned *ticket, cg_dev_state<T> *st)
        {
            if(st->status)
                return;
            double s = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
                s += (double)p[i] * (double)q[i];
            block_sum_epilogue<T, EPI_PQ>(s, partial, ticket, st);
        }

        template <typename T>
        __global__ void __launch_bounds__(RED_THREADS) cgd_step_kernel(long long n, const T *__restrict__ p, const T *__restrict__ q,
                                                                      T *__restrict__ x, T *__restrict__ r, double *partial,
                                                                      unsigned *ticket, cg_dev_state<T> *st)
        {
            if(st->status)
                return;
            const T alpha = st->alpha;
            double  s     = 0;
            for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS)
            {
                x[i] += alpha * p[i];
                const T ri = r[i] + alpha * q[i];
                r[i]       = ri;
                s += (double)ri * (double)ri;
            }
            block_sum_epilogue<T, EPI_STEP>(s, partial, ticket, st);
        }


        struct managed_buf
        {
            void  *p     = nullptr;
            size_t bytes = 0;
            managed_buf() = default;
            managed_buf(const managed_buf &)            = delete;
            managed_buf &operator=(const managed_buf &) = delete;
            ~managed_buf()
            {
                release();
            }
            void release()
            {
                if(p)
                    cudaFree(p);
                p     = nullptr;
                bytes = 0;
            }
            aoclsparse_status alloc(size_t n, cudaStream_t st, bool managed = true)
            {
                release();
                if(!managed)
                {
                    B200_CUDA(cudaMalloc(&p, n ? n : 16));
                    bytes = n;
                    return aoclsparse_status_success;
                }
                B200_CUDA(cudaMallocManaged(&p, n ? n : 16));
                bytes = n;
                int dev = 0;
                cudaGetDevice(&dev);
                // first touch on the device; a failure here (no prefetch support) is harmless
                if(cudaMemPrefetchAsync(p, n ? n : 16, dev, st) != cudaSuccess)
                    cudaGetLastError();
                return aoclsparse_status_success;
            }
            template <typename T>
            T *as() const
            {
                return static_cast<T *>(p);
            }
        };

        template <typename T>
        struct itsol_data
        {
            options<T>     opts;
            aoclsparse_int n       = 0;
            bool           have_b  = false;
            bool           solving = false;
            bool           host_visible = true; // work vectors in managed memory (reverse communication, callbacks)
            bool           vectors_managed = true; // what the vectors currently allocated are
            managed_buf    b, r, p, q, z, xw; // xw: device-side copy of a host-resident x
            dev_buf        partial, ticket;
            double        *h_result = nullptr, *d_result = nullptr; // mapped page-locked scalar
            ~itsol_data()
            {
                if(h_result)
                    cudaFreeHost(h_result);
            }
            // CG state (cg_data, aoclsparse_itsol_data.hpp)
            int  task = task_start, niter = 0, precond = 0, maxit = 500;
            T    rtol = 0, atol = 0, rnorm2 = 0, bnorm2 = 0, brtol = 0, rz = 0, alpha = 0, beta = 0;
            T   *x_dev       = nullptr; // where the solver updates x
            T   *x_user      = nullptr; // the caller's x when it is host memory (copied back before every return)
        };

        template <typename T>
        aoclsparse_status reduce_to_host(itsol_data<T> *it, cudaStream_t st, double &out)
        {
            B200_CUDA(cudaStreamSynchronize(st));
            out = *static_cast<volatile double *>(it->h_result);
            return aoclsparse_status_success;
        }

        template <typename T>
        aoclsparse_status sync_x_to_user(itsol_data<T> *it, cudaStream_t st)
        {
            if(it->x_user && it->n > 0)
            {
                B200_CUDA(cudaMemcpyAsync(it->x_user, it->x_dev, sizeof(T) * (size_t)it->n, cudaMemcpyDeviceToHost, st));
                B200_CUDA(cudaStreamSynchronize(st));
            }
            return aoclsparse_status_success;
        }

        template <typename T>
        bool negative_or_nearzero(T v)
        {
            // aoclsparse_is_negative_or_nearzero, aoclsparse_utils.hpp:627-640 (scale 1e-2)
            return v <= (T)1e-2 * (T)2.0 * std::numeric_limits<T>::epsilon();
        }

        template <typename T>
        aoclsparse_status rci_input(itsol_data<T> *it, aoclsparse_int n, const T *b)
        {
            // aoclsparse_itsol_rci_input, itsol_functions.hpp:296-330
            if(it == nullptr)
                return aoclsparse_status_internal_error;
            if(n < 0)
                return aoclsparse_status_invalid_value;
            if(!b)
                return aoclsparse_status_invalid_pointer;
            cudaStream_t st = current_stream();
            it->have_b      = false;
            // the work vectors are kept from one solve to the next when size and memory kind are unchanged
            if(it->n != n || !it->b.p || it->vectors_managed != it->host_visible)
            {
                B200_TRY(it->b.alloc(sizeof(T) * (size_t)n, st, it->host_visible));
                for(managed_buf *m : {&it->r, &it->p, &it->q, &it->z})
                    B200_TRY(m->alloc(sizeof(T) * (size_t)n, st, it->host_visible));
                it->vectors_managed = it->host_visible;
                it->xw.release();
            }
            B200_TRY(it->partial.alloc(sizeof(double) * RED_BLOCKS));
            B200_TRY(it->ticket.alloc(sizeof(unsigned)));
            B200_CUDA(cudaMemsetAsync(it->ticket.p, 0, sizeof(unsigned), st));
            if(!it->h_result)
            {
                B200_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&it->h_result), sizeof(double), cudaHostAllocMapped));
                B200_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&it->d_result), it->h_result, 0));
            }
            if(n > 0)
                B200_CUDA(cudaMemcpyAsync(it->b.p, b, sizeof(T) * (size_t)n, cudaMemcpyDefault, st));
            B200_CUDA(cudaStreamSynchronize(st));
            it->n       = n;
            it->have_b  = true;
            it->solving = false;
            return aoclsparse_status_success;
        }

        template <typename T>
        aoclsparse_status solver_init(itsol_data<T> *it)
        {
            // aoclsparse_itsol_solver_init + aoclsparse_cg_data_options, itsol_functions.hpp:205-218,335-380
            if(it->opts.solver != 0)
                return aoclsparse_status_not_implemented; // GMRES
            it->task    = task_start;
            it->precond = it->opts.cg_precond;
            it->rtol    = it->opts.cg_rtol;
            it->atol    = it->opts.cg_atol;
            it->maxit   = it->opts.cg_maxit;
            return aoclsparse_status_success;
        }

        // aoclsparse_cg_rci_solve, itsol_functions.hpp:633-870 -- same tasks, same scalar logic, vector work on the device
        template <typename T>
        aoclsparse_status cg_rci(itsol_data<T> *it, aoclsparse_itsol_rci_job *ircomm, T **u, T **v, T *x, T rinfo[100])
        {
            aoclsparse_status exit_status = aoclsparse_status_success;
            cudaStream_t      st          = current_stream();
            const long long   n           = it->n;
            T                *r = it->r.template as<T>(), *p = it->p.template as<T>(), *q = it->q.template as<T>(),
              *z = it->z.template as<T>();
            const reduce_out partial{it->partial.template as<double>(), it->ticket.template as<unsigned>(), it->d_result};
            double  red     = 0;
            bool    loop;
            if(it->task != task_start && *ircomm == aoclsparse_rci_interrupt)
            {
                *ircomm = aoclsparse_rci_stop;
                sync_x_to_user(it, st);
                return aoclsparse_status_user_stop;
            }
            do
            {
                loop = false;
                switch(it->task)
                {
                case task_start:
                    for(int i = 0; i < 100; ++i)
                        rinfo[i] = (T)0;
                    it->niter = 0;
                    // where x lives while solving
                    if(is_device_accessible(x))
                    {
                        it->x_dev  = x;
                        it->x_user = nullptr;
                    }
                    else
                    {
                        B200_TRY(it->xw.alloc(sizeof(T) * (size_t)n, st, it->host_visible));
                        if(n > 0)
                            B200_CUDA(cudaMemcpyAsync(it->xw.p, x, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, st));
                        it->x_dev  = it->xw.template as<T>();
                        it->x_user = x;
                    }
                    cg_start_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, it->b.template as<T>(), it->x_dev, r, p, partial);
                    B200_LAUNCHED();
                    B200_TRY(reduce_to_host(it, st, red));
                    it->bnorm2 = (T)std::sqrt(red);
                    if(it->bnorm2 != it->bnorm2)
                        return aoclsparse_status_invalid_value; // b is rubbish
                    rinfo[1]  = it->bnorm2;
                    it->brtol = it->rtol * it->bnorm2;
                    *ircomm   = aoclsparse_rci_mv;
                    it->task  = task_init_res;
                    *u        = p;
                    *v        = q;
                    break;

                case task_init_res:
                    // q = A x, r = -b  ->  r = A x - b, p = 0
                    cg_residual_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, r, q, p, partial);
                    B200_LAUNCHED();
                    B200_TRY(reduce_to_host(it, st, red));
                    it->rnorm2 = (T)std::sqrt(red);
                    if(it->rnorm2 != it->rnorm2)
                    {
                        exit_status = aoclsparse_status_numerical_error;
                        break;
                    }
                    rinfo[0] = it->rnorm2;
                    it->rz   = (T)1;
                    it->task = task_check_conv;
                    // fall through
                case task_check_conv:
                    *u = r;
                    *v = nullptr;
                    if((T)0 < it->atol && it->rnorm2 <= it->atol)
                    {
                        *ircomm = aoclsparse_rci_stop;
                        break;
                    }
                    if((T)0 < it->rtol && it->rnorm2 <= it->brtol)
                    {
                        *ircomm = aoclsparse_rci_stop;
                        break;
                    }
                    if(it->maxit > 0 && it->niter > it->maxit)
                    {
                        *ircomm     = aoclsparse_rci_stop;
                        exit_status = aoclsparse_status_maxit;
                        break;
                    }
                    it->task = task_start_iter;
                    *ircomm  = aoclsparse_rci_stopping_criterion;
                    break;

                case task_start_iter:
                    it->niter++;
                    rinfo[30] = (T)it->niter;
                    it->task  = task_compute_beta;
                    if(it->precond)
                    {
                        *ircomm = aoclsparse_rci_precond;
                        *u      = r;
                        *v      = z;
                        break;
                    }
                    // unpreconditioned: z = r, no copy needed -- r.z = |r|^2 is already known
                    // fall through
                case task_compute_beta:
                {
                    T        rz_new;
                    const T *zz = it->precond ? z : r;
                    if(it->precond)
                    {
                        dot_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, r, z, partial);
                        B200_LAUNCHED();
                        B200_TRY(reduce_to_host(it, st, red));
                        rz_new = (T)red;
                    }
                    else
                        rz_new = it->rnorm2 * it->rnorm2;
                    if(negative_or_nearzero(it->rz))
                        return aoclsparse_status_numerical_error;
                    it->beta = rz_new / it->rz;
                    it->rz   = rz_new;
                    cg_direction_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, it->beta, p, zz);
                    B200_LAUNCHED();
                    *ircomm  = aoclsparse_rci_mv;
                    it->task = task_take_step;
                    *u       = p;
                    *v       = q;
                    break;
                }

                case task_take_step:
                {
                    dot_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, p, q, partial);
                    B200_LAUNCHED();
                    B200_TRY(reduce_to_host(it, st, red));
                    const T pq = (T)red;
                    if(negative_or_nearzero(pq) || pq == (T)0)
                        return aoclsparse_status_numerical_error; // A is not positive definite
                    it->alpha = it->rz / pq;
                    cg_step_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, it->alpha, p, q, it->x_dev, r, partial);
                    B200_LAUNCHED();
                    B200_TRY(reduce_to_host(it, st, red));
                    it->rnorm2 = (T)std::sqrt(red);
                    if(it->rnorm2 != it->rnorm2)
                    {
                        exit_status = aoclsparse_status_numerical_error;
                        break;
                    }
                    rinfo[0] = it->rnorm2;
                    loop     = true;
                    it->task = task_check_conv;
                    break;
                }

                default:
                    *ircomm = aoclsparse_rci_stop;
                    return aoclsparse_status_internal_error;
                }
            } while(loop);
            // a host-resident x is brought up to date whenever the caller may look at it
            if(*ircomm == aoclsparse_rci_stop || *ircomm == aoclsparse_rci_stopping_criterion || exit_status != aoclsparse_status_success)
                B200_TRY(sync_x_to_user(it, st));
            else
                B200_CUDA(cudaStreamSynchronize(st)); // *u / *v are about to be read or written by the caller
            return exit_status;
        }

        // aoclsparse_itsol_rci_solve, itsol_functions.hpp:486-553
        template <typename T>
        aoclsparse_status rci_solve(itsol_data<T> *it, aoclsparse_itsol_rci_job *ircomm, T **u, T **v, T *x, T rinfo[100])
        {
            if(ircomm == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(it == nullptr)
            {
                *ircomm = aoclsparse_rci_stop;
                return aoclsparse_status_internal_error;
            }
            if(u == nullptr || v == nullptr || x == nullptr || rinfo == nullptr || !it->have_b)
            {
                *ircomm = aoclsparse_rci_stop;
                return aoclsparse_status_invalid_pointer;
            }
            aoclsparse_status status;
            if(!it->solving && !it->vectors_managed)
            {
                // The last forward solve ran without callbacks and left the work vectors in plain device memory; this
                // interface hands *u / *v to a caller who may read and write them from the host (the reference allows
                // rci_solve after solve without a new rci_input): move b to managed memory, re-allocate the others.
                cudaStream_t    st    = current_stream();
                const size_t    bytes = sizeof(T) * (size_t)it->n;
                managed_buf     nb;
                status = nb.alloc(bytes, st, true);
                if(status == aoclsparse_status_success && bytes
                   && cudaMemcpyAsync(nb.p, it->b.p, bytes, cudaMemcpyDefault, st) != cudaSuccess)
                    status = aoclsparse_status_internal_error;
                if(status == aoclsparse_status_success && cudaStreamSynchronize(st) != cudaSuccess)
                    status = aoclsparse_status_internal_error;
                for(managed_buf *m : {&it->r, &it->p, &it->q, &it->z})
                    if(status == aoclsparse_status_success)
                        status = m->alloc(bytes, st, true);
                if(status != aoclsparse_status_success)
                {
                    *ircomm = aoclsparse_rci_stop;
                    return status;
                }
                std::swap(it->b.p, nb.p);
                std::swap(it->b.bytes, nb.bytes);
                it->vectors_managed = true;
                it->xw.release();
            }
            if(!it->solving)
            {
                status = solver_init(it);
                if(status != aoclsparse_status_success)
                {
                    *ircomm = aoclsparse_rci_stop;
                    return status;
                }
                it->solving     = true;
                it->opts.locked = true;
            }
            status = cg_rci(it, ircomm, u, v, x, rinfo);
            if(status != aoclsparse_status_success)
                *ircomm = aoclsparse_rci_stop;
            if(*ircomm == aoclsparse_rci_stop)
            {
                it->solving     = false;
                it->opts.locked = false;
            }
            return status;
        }

        template <typename T>
        struct mv_of;
        template <>
        struct mv_of<float>
        {
            static aoclsparse_status call(const float *a, aoclsparse_matrix A, const aoclsparse_mat_descr d, const float *x, const float *b, float *y)
            {
                return aoclsparse_smv(aoclsparse_operation_none, a, A, d, x, b, y);
            }
        };
        template <>
        struct mv_of<double>
        {
            static aoclsparse_status call(const double *a, aoclsparse_matrix A, const aoclsparse_mat_descr d, const double *x, const double *b, double *y)
            {
                return aoclsparse_dmv(aoclsparse_operation_none, a, A, d, x, b, y);
            }
        };

        // Unpreconditioned CG without callbacks, driven from the device (see cg_dev_state above).  `it` has been prepared by
        // rci_input + solver_init; the matrix has been hinted and optimized.
        template <typename T>
        aoclsparse_status device_driven_cg(itsol_data<T> *it, aoclsparse_matrix mat, const aoclsparse_mat_descr descr, T *x, T rinfo[100])
        {
            cudaStream_t    st = current_stream();
            const long long n  = it->n;
            T              *r = it->r.template as<T>(), *p = it->p.template as<T>(), *q = it->q.template as<T>();
            T              *xd     = x;
            bool            x_host = false;
            if(!is_device_accessible(x))
            {
                if(!it->xw.p || it->xw.bytes != sizeof(T) * (size_t)n)
                    B200_TRY(it->xw.alloc(sizeof(T) * (size_t)n, st, false));
                if(n > 0)
                    B200_CUDA(cudaMemcpyAsync(it->xw.p, x, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, st));
                xd     = it->xw.template as<T>();
                x_host = true;
            }
            dev_buf dstate;
            B200_TRY(dstate.alloc(sizeof(cg_dev_state<T>)));
            cg_dev_state<T> h{};
            h.atol  = it->atol;
            h.rtol  = it->rtol;
            h.maxit = it->maxit;
            B200_CUDA(cudaMemcpyAsync(dstate.p, &h, sizeof(h), cudaMemcpyHostToDevice, st));
            cg_dev_state<T> *ds      = dstate.as<cg_dev_state<T>>();
            double          *partial = it->partial.template as<double>();
            unsigned        *ticket  = it->ticket.template as<unsigned>();
            const T          one = (T)1, zero = (T)0;
            cgd_start_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, it->b.template as<T>(), xd, r, p, partial, ticket, ds);
            B200_LAUNCHED();
            if(mv_of<T>::call(&one, mat, descr, p, &zero, q) != aoclsparse_status_success)
                return aoclsparse_status_internal_error;
            cgd_residual_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, r, q, p, partial, ticket, ds);
            B200_LAUNCHED();
            constexpr int BATCH = 4;
            while(true)
            {
                B200_CUDA(cudaMemcpyAsync(&h, dstate.p, sizeof(h), cudaMemcpyDeviceToHost, st));
                B200_CUDA(cudaStreamSynchronize(st));
                if(h.status != 0)
                    break;
                for(int k = 0; k < BATCH; ++k)
                {
                    cgd_direction_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, p, r, ds);
                    B200_LAUNCHED();
                    if(mv_of<T>::call(&one, mat, descr, p, &zero, q) != aoclsparse_status_success)
                        return aoclsparse_status_internal_error;
                    cgd_dot_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, p, q, partial, ticket, ds);
                    B200_LAUNCHED();
                    cgd_step_kernel<T><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, p, q, xd, r, partial, ticket, ds);
                    B200_LAUNCHED();
                }
            }
            if(x_host && n > 0)
            {
                B200_CUDA(cudaMemcpyAsync(x, xd, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, st));
                B200_CUDA(cudaStreamSynchronize(st));
            }
            rinfo[0]  = h.rnorm2;
            rinfo[1]  = h.bnorm2;
            rinfo[30] = (T)h.niter;
            switch(h.status)
            {
            case 1:
                return aoclsparse_status_success;
            case 2:
                return aoclsparse_status_maxit;
            case 4:
                return aoclsparse_status_invalid_value;
            default:
                return aoclsparse_status_numerical_error;
            }
        }

        // aoclsparse_itsol_solve + aoclsparse_cg_solve, itsol_functions.hpp:555-624,1369-1500
        template <typename T>
        aoclsparse_status forward_solve(itsol_data<T>             *it,
                                        aoclsparse_int             n,
                                        aoclsparse_matrix          mat,
                                        const aoclsparse_mat_descr descr,
                                        const T                   *b,
                                        T                         *x,
                                        T                          rinfo[100],
                                        aoclsparse_int             precond(aoclsparse_int, aoclsparse_int, const T *, T *, void *),
                                        aoclsparse_int             monit(aoclsparse_int, const T *, const T *, T *, void *),
                                        void                      *udata)
        {
            if(it == nullptr)
                return aoclsparse_status_internal_error;
            if(x == nullptr || rinfo == nullptr)
                return aoclsparse_status_invalid_pointer;
            for(int i = 0; i < 100; ++i)
                rinfo[i] = (T)0;
            // without callbacks nothing outside this library sees the work vectors: plain device memory
            // (AOCLSPARSE_B200_ITSOL_MANAGED=1 keeps them managed, for measurements)
            const char *em   = getenv("AOCLSPARSE_B200_ITSOL_MANAGED");
            it->host_visible = precond != nullptr || monit != nullptr || (em && atoi(em) != 0);
            const aoclsparse_status in_st = rci_input(it, n, b);
            it->host_visible              = true; // the reverse-communication entry points always hand out managed memory
            B200_TRY(in_st);
            B200_TRY(solver_init(it));
            if(mat == nullptr || descr == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(mat->m != n || mat->n != n)
                return aoclsparse_status_invalid_size;
            if(descr->type != aoclsparse_matrix_type_symmetric || descr->fill_mode != aoclsparse_fill_mode_lower)
                return aoclsparse_status_invalid_value;
            if(it->precond == 1 && precond == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(it->precond == 3)
                return aoclsparse_status_not_implemented; // symmetric Gauss-Seidel: two triangular solves per iteration
            // the symmetric product runs as a plain streaming gather on the expanded copy (spmv.cu, expand.cu): hint and
            // optimize once per matrix
            {
                const int want  = get_doid(false, descr->type, descr->fill_mode, aoclsparse_operation_none);
                bool      ready = false;
                {
                    std::shared_lock<std::shared_mutex> rl(mat->guard);
                    for(const hint &hh : mat->hints)
                        ready = ready || (hh.act == 1 && hh.doid == want && hh.done);
                }
                if(!ready && aoclsparse_set_mv_hint(mat, aoclsparse_operation_none, descr, 100) == aoclsparse_status_success)
                    aoclsparse_optimize(mat);
            }
            // no callback, no preconditioner: the whole loop runs without host round trips
            // (AOCLSPARSE_B200_ITSOL_HOST_DRIVEN=1 keeps the host-driven state machine, for comparison)
            {
                const char *eh = getenv("AOCLSPARSE_B200_ITSOL_HOST_DRIVEN");
                if(!precond && !monit && it->precond == 0 && !(eh && atoi(eh) != 0))
                    return device_driven_cg<T>(it, mat, descr, x, rinfo);
            }
            it->solving     = true;
            it->opts.locked = true;
            aoclsparse_itsol_rci_job ircomm      = aoclsparse_rci_start;
            T                       *u = nullptr, *v = nullptr;
            const T                  one = (T)1, zero = (T)0;
            aoclsparse_status        exit_status = aoclsparse_status_success;
            while(ircomm != aoclsparse_rci_stop)
            {
                exit_status = rci_solve(it, &ircomm, &u, &v, x, rinfo);
                if(exit_status != aoclsparse_status_success && ircomm != aoclsparse_rci_stop)
                    break;
                if(ircomm == aoclsparse_rci_mv)
                {
                    if(mv_of<T>::call(&one, mat, descr, u, &zero, v) != aoclsparse_status_success)
                    {
                        exit_status = aoclsparse_status_internal_error;
                        break;
                    }
                }
                else if(ircomm == aoclsparse_rci_precond)
                {
                    // user routine on managed memory: readable and writable from the host
                    if(precond(0, n, u, v, udata) != 0)
                        ircomm = aoclsparse_rci_interrupt;
                }
                else if(ircomm == aoclsparse_rci_stopping_criterion && monit)
                {
                    // monit(n, x, r, rinfo, udata): x as the caller knows it (kept up to date by cg_rci), r managed
                    if(monit(n, it->x_user ? it->x_user : it->x_dev, u, rinfo, udata) != 0)
                        ircomm = aoclsparse_rci_interrupt;
                }
            }
            it->solving     = false;
            it->opts.locked = false;
            return exit_status;
        }

        // ================================================================ complex conjugate gradients (c / z handles)
        // The reference instantiates the SAME state machine for std::complex (itsol_functions.hpp:633-870, handles made by
        // aoclsparse_itsol_{c,z}_init, itsol_functions.cpp:168-230).  What that means for complex data, restated here:
        //   * r.z and p.q are UNCONJUGATED sums of products (:795-797, :822-824); with a complex SYMMETRIC matrix -- the
        //     forward interface insists on a symmetric, lower descriptor (:1411-1417) -- that is the conjugate-orthogonal
        //     recurrence, not the Hermitian one;
        //   * rz starts at (1, 1) (:718-725); the breakdown tests compare |rz| and |p.q| with the near-zero bound
        //     (aoclsparse_utils.hpp:627-640); tolerances, norms and rinfo are real.
        // Host-driven: the scalar logic runs on the host as in cg_rci, every vector operation is a kernel; two complex
        // scalars and one norm per iteration travel back through mapped memory.  Work vectors are managed memory (a host
        // caller of the reverse-communication interface reads and writes *u / *v as ordinary memory).
        template <typename R>
        struct cplx_types;
        template <>
        struct cplx_types<float>
        {
            using dev = float2;
            using api = aoclsparse_float_complex;
        };
        template <>
        struct cplx_types<double>
        {
            using dev = double2;
            using api = aoclsparse_double_complex;
        };

        template <typename C>
        __device__ __forceinline__ C c_mul(C a, C b)
        {
            C r;
            r.x = a.x * b.x - a.y * b.y;
            r.y = a.x * b.y + a.y * b.x;
            return r;
        }

        // two block sums -> partial[2 b], partial[2 b + 1]; the LAST block adds the partials in index order and stores
        // both sums into result[0..1] (mapped host memory)
        __device__ __forceinline__ void block_sum_store2(double s0, double s1, reduce_out o)
        {
            __shared__ double sh[2][RED_THREADS / 32];
            __shared__ bool   last;
            __shared__ double fin[2][RED_THREADS];
#pragma unroll
            for(int k = 16; k > 0; k >>= 1)
            {
                s0 += __shfl_down_sync(0xffffffffu, s0, k);
                s1 += __shfl_down_sync(0xffffffffu, s1, k);
            }
            if((threadIdx.x & 31) == 0)
            {
                sh[0][threadIdx.x >> 5] = s0;
                sh[1][threadIdx.x >> 5] = s1;
            }
            __syncthreads();
            if(threadIdx.x < 32)
            {
                s0 = threadIdx.x < RED_THREADS / 32 ? sh[0][threadIdx.x] : 0.0;
                s1 = threadIdx.x < RED_THREADS / 32 ? sh[1][threadIdx.x] : 0.0;
#pragma unroll
                for(int k = 16; k > 0; k >>= 1)
                {
                    s0 += __shfl_down_sync(0xffffffffu, s0, k);
                    s1 += __shfl_down_sync(0xffffffffu, s1, k);
                }
                if(threadIdx.x == 0)
                {
                    o.partial[2 * blockIdx.x]     = s0;
                    o.partial[2 * blockIdx.x + 1] = s1;
                    __threadfence();
                    last = atomicAdd(o.ticket, 1u) == gridDim.x - 1;
                }
            }
            __syncthreads();
            if(!last)
                return;
            __threadfence();
            double t0 = 0, t1 = 0;
            for(int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS)
            {
                t0 += __ldcg(o.partial + 2 * i);
                t1 += __ldcg(o.partial + 2 * i + 1);
            }
            fin[0][threadIdx.x] = t0;
            fin[1][threadIdx.x] = t1;
            __syncthreads();
            for(int k = RED_THREADS / 2; k > 0; k >>= 1)
            {
                if(threadIdx.x < k)
                {
                    fin[0][threadIdx.x] += fin[0][threadIdx.x + k];
                    fin[1][threadIdx.x] += fin[1][threadIdx.x + k];
                }
                __syncthreads();
            }
            if(threadIdx.x == 0)
            {
                o.result[0] = fin[0][0];
                o.result[1] = fin[1][0];
                *o.ticket   = 0;
                __threadfence_system();
            }
        }

#define B200_GRID_STRIDE(i, n) \
    for(long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < (n); i += (long long)gridDim.x * RED_THREADS)

        // r = -b, p = x, |b|^2
        template <typename C>
        __global__ void __launch_bounds__(RED_THREADS) ccg_start_kernel(long long n, const C *__restrict__ b, const C *__restrict__ x,
                                                                       C *__restrict__ r, C *__restrict__ p, reduce_out out)
        {
            double s = 0;
            B200_GRID_STRIDE(i, n)
            {
                const C bi = b[i];
                C       ri;
                ri.x = -bi.x;
                ri.y = -bi.y;
                r[i] = ri;
                p[i] = x[i];
                s += (double)bi.x * (double)bi.x + (double)bi.y * (double)bi.y;
            }
            block_sum_store2(s, 0.0, out);
        }

        // r += q, p = 0, |r|^2
        template <typename C>
        __global__ void __launch_bounds__(RED_THREADS) ccg_residual_kernel(long long n, C *__restrict__ r, const C *__restrict__ q,
                                                                          C *__restrict__ p, reduce_out out)
        {
            double s = 0;
            B200_GRID_STRIDE(i, n)
            {
                C       ri = r[i];
                const C qi = q[i];
                ri.x += qi.x;
                ri.y += qi.y;
                r[i] = ri;
                C zero;
                zero.x = 0;
                zero.y = 0;
                p[i]   = zero;
                s += (double)ri.x * (double)ri.x + (double)ri.y * (double)ri.y;
            }
            block_sum_store2(s, 0.0, out);
        }

        // sum of a[i] * b[i], NOT conjugated
        template <typename C>
        __global__ void __launch_bounds__(RED_THREADS) cdot_kernel(long long n, const C *__restrict__ a, const C *__restrict__ b, reduce_out out)
        {
            double s0 = 0, s1 = 0;
            B200_GRID_STRIDE(i, n)
            {
                const C ai = a[i], bi = b[i];
                s0 += (double)ai.x * (double)bi.x - (double)ai.y * (double)bi.y;
                s1 += (double)ai.x * (double)bi.y + (double)ai.y * (double)bi.x;
            }
            block_sum_store2(s0, s1, out);
        }

        // p = beta p - z
        template <typename C>
        __global__ void __launch_bounds__(RED_THREADS) ccg_direction_kernel(long long n, C beta, C *__restrict__ p, const C *__restrict__ z)
        {
            B200_GRID_STRIDE(i, n)
            {
                C       t  = c_mul(beta, p[i]);
                const C zi = z[i];
                t.x -= zi.x;
                t.y -= zi.y;
                p[i] = t;
            }
        }

        // x += alpha p, r += alpha q, |r|^2
        template <typename C>
        __global__ void __launch_bounds__(RED_THREADS) ccg_step_kernel(long long n, C alpha, const C *__restrict__ p, const C *__restrict__ q,
                                                                      C *__restrict__ x, C *__restrict__ r, reduce_out out)
        {
            double s = 0;
            B200_GRID_STRIDE(i, n)
            {
                const C ap = c_mul(alpha, p[i]), aq = c_mul(alpha, q[i]);
                C       xi = x[i], ri = r[i];
                xi.x += ap.x;
                xi.y += ap.y;
                ri.x += aq.x;
                ri.y += aq.y;
                x[i] = xi;
                r[i] = ri;
                s += (double)ri.x * (double)ri.x + (double)ri.y * (double)ri.y;
            }
            block_sum_store2(s, 0.0, out);
        }
#undef B200_GRID_STRIDE

        template <typename R>
        struct itsol_cdata
        {
            using C = typename cplx_types<R>::dev;
            options<R>      opts;
            aoclsparse_int  n       = 0;
            bool            have_b  = false;
            bool            solving = false;
            managed_buf     b, r, p, q, z, xw;
            dev_buf         partial, ticket;
            double         *h_result = nullptr, *d_result = nullptr; // two mapped page-locked doubles
            int             task = task_start, niter = 0, precond = 0, maxit = 500;
            R               rtol = 0, atol = 0, rnorm2 = 0, bnorm2 = 0, brtol = 0;
            std::complex<R> rz, alpha, beta;
            C              *x_dev = nullptr, *x_user = nullptr;
            ~itsol_cdata()
            {
                if(h_result)
                    cudaFreeHost(h_result);
            }
        };

        template <typename R>
        aoclsparse_status crci_input(itsol_cdata<R> *it, aoclsparse_int n, const void *b)
        {
            using C = typename cplx_types<R>::dev;
            if(it == nullptr)
                return aoclsparse_status_internal_error;
            if(n < 0)
                return aoclsparse_status_invalid_value;
            if(!b)
                return aoclsparse_status_invalid_pointer;
            cudaStream_t st = current_stream();
            it->have_b      = false;
            if(it->n != n || !it->b.p)
            {
                for(managed_buf *m : {&it->b, &it->r, &it->p, &it->q, &it->z})
                    B200_TRY(m->alloc(sizeof(C) * (size_t)n, st, true));
                it->xw.release();
            }
            B200_TRY(it->partial.alloc(sizeof(double) * 2 * RED_BLOCKS));
            B200_TRY(it->ticket.alloc(sizeof(unsigned)));
            B200_CUDA(cudaMemsetAsync(it->ticket.p, 0, sizeof(unsigned), st));
            if(!it->h_result)
            {
                B200_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&it->h_result), 2 * sizeof(double), cudaHostAllocMapped));
                B200_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&it->d_result), it->h_result, 0));
            }
            if(n > 0)
                B200_CUDA(cudaMemcpyAsync(it->b.p, b, sizeof(C) * (size_t)n, cudaMemcpyDefault, st));
            B200_CUDA(cudaStreamSynchronize(st));
            it->n       = n;
            it->have_b  = true;
            it->solving = false;
            return aoclsparse_status_success;
        }

        template <typename R>
        aoclsparse_status ccg_rci(itsol_cdata<R> *it, aoclsparse_itsol_rci_job *ircomm, void **u, void **v, void *x_, R rinfo[100])
        {
            using C  = typename cplx_types<R>::dev;
            using SC = std::complex<R>;
            cudaStream_t        st = current_stream();
            const aoclsparse_int n  = it->n;
            C                  *x  = static_cast<C *>(x_);
            C *r = it->r.template as<C>(), *p = it->p.template as<C>(), *q = it->q.template as<C>(), *z = it->z.template as<C>();
            reduce_out out{it->partial.template as<double>(), it->ticket.template as<unsigned>(), it->d_result};
            auto       fetch = [&](double &re, double &im) -> aoclsparse_status {
                B200_CUDA(cudaStreamSynchronize(st));
                re = static_cast<volatile double *>(it->h_result)[0];
                im = static_cast<volatile double *>(it->h_result)[1];
                return aoclsparse_status_success;
            };
            auto sync_x = [&]() -> aoclsparse_status {
                if(it->x_user && n > 0)
                    B200_CUDA(cudaMemcpyAsync(it->x_user, it->x_dev, sizeof(C) * (size_t)n, cudaMemcpyDeviceToHost, st));
                B200_CUDA(cudaStreamSynchronize(st));
                return aoclsparse_status_success;
            };
            const R eps_tol = (R)1e-2 * (R)2.0 * std::numeric_limits<R>::epsilon();
            if(it->task != task_start && *ircomm == aoclsparse_rci_interrupt)
            {
                B200_TRY(sync_x());
                *ircomm = aoclsparse_rci_stop;
                return aoclsparse_status_user_stop;
            }
            aoclsparse_status exit_status = aoclsparse_status_success;
            bool              loop;
            double            re = 0, im = 0;
            do
            {
                loop = false;
                switch(it->task)
                {
                case task_start:
                    for(int i = 0; i < 100; ++i)
                        rinfo[i] = (R)0;
                    it->niter = 0;
                    if(is_device_accessible(x))
                    {
                        it->x_dev  = x;
                        it->x_user = nullptr;
                    }
                    else
                    {
                        B200_TRY(it->xw.alloc(sizeof(C) * (size_t)n, st, true));
                        if(n > 0)
                            B200_CUDA(cudaMemcpyAsync(it->xw.p, x, sizeof(C) * (size_t)n, cudaMemcpyHostToDevice, st));
                        it->x_dev  = it->xw.template as<C>();
                        it->x_user = x;
                    }
                    ccg_start_kernel<C><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, it->b.template as<C>(), it->x_dev, r, p, out);
                    B200_LAUNCHED();
                    B200_TRY(fetch(re, im));
                    it->bnorm2 = (R)std::sqrt(re);
                    if(it->bnorm2 != it->bnorm2)
                        return aoclsparse_status_invalid_value; // b is rubbish
                    rinfo[1]  = it->bnorm2;
                    it->brtol = it->rtol * it->bnorm2;
                    *ircomm   = aoclsparse_rci_mv;
                    it->task  = task_init_res;
                    *u        = p;
                    *v        = q;
                    break;

                case task_init_res:
                    ccg_residual_kernel<C><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, r, q, p, out);
                    B200_LAUNCHED();
                    B200_TRY(fetch(re, im));
                    it->rnorm2 = (R)std::sqrt(re);
                    if(it->rnorm2 != it->rnorm2)
                    {
                        exit_status = aoclsparse_status_numerical_error;
                        break;
                    }
                    rinfo[0] = it->rnorm2;
                    it->rz   = SC((R)1, (R)1);
                    it->task = task_check_conv;
                    // fall through
                case task_check_conv:
                    *u = r;
                    *v = nullptr;
                    if((R)0 < it->atol && it->rnorm2 <= it->atol)
                    {
                        *ircomm = aoclsparse_rci_stop;
                        break;
                    }
                    if((R)0 < it->rtol && it->rnorm2 <= it->brtol)
                    {
                        *ircomm = aoclsparse_rci_stop;
                        break;
                    }
                    if(it->maxit > 0 && it->niter > it->maxit)
                    {
                        *ircomm     = aoclsparse_rci_stop;
                        exit_status = aoclsparse_status_maxit;
                        break;
                    }
                    it->task = task_start_iter;
                    *ircomm  = aoclsparse_rci_stopping_criterion;
                    break;

                case task_start_iter:
                    it->niter++;
                    rinfo[30] = (R)it->niter;
                    it->task  = task_compute_beta;
                    if(it->precond)
                    {
                        *ircomm = aoclsparse_rci_precond;
                        *u      = r;
                        *v      = z;
                        break;
                    }
                    // unpreconditioned: z = r (no copy; the dot product below reads r twice)
                    // fall through
                case task_compute_beta:
                {
                    const C *zz = it->precond ? z : r;
                    cdot_kernel<C><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, r, zz, out);
                    B200_LAUNCHED();
                    B200_TRY(fetch(re, im));
                    const SC rz_new((R)re, (R)im);
                    if(std::abs(it->rz) <= eps_tol)
                        return aoclsparse_status_numerical_error;
                    it->beta = rz_new / it->rz;
                    it->rz   = rz_new;
                    C beta;
                    beta.x = it->beta.real();
                    beta.y = it->beta.imag();
                    ccg_direction_kernel<C><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, beta, p, zz);
                    B200_LAUNCHED();
                    *ircomm  = aoclsparse_rci_mv;
                    it->task = task_take_step;
                    *u       = p;
                    *v       = q;
                    break;
                }

                case task_take_step:
                {
                    cdot_kernel<C><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, p, q, out);
                    B200_LAUNCHED();
                    B200_TRY(fetch(re, im));
                    const SC pq((R)re, (R)im);
                    if(std::abs(pq) <= eps_tol || pq == SC((R)0, (R)0))
                        return aoclsparse_status_numerical_error;
                    it->alpha = it->rz / pq;
                    C alpha;
                    alpha.x = it->alpha.real();
                    alpha.y = it->alpha.imag();
                    ccg_step_kernel<C><<<RED_BLOCKS, RED_THREADS, 0, st>>>(n, alpha, p, q, it->x_dev, r, out);
                    B200_LAUNCHED();
                    B200_TRY(fetch(re, im));
                    it->rnorm2 = (R)std::sqrt(re);
                    if(it->rnorm2 != it->rnorm2)
                    {
                        exit_status = aoclsparse_status_numerical_error;
                        break;
                    }
                    rinfo[0] = it->rnorm2;
                    loop     = true;
                    it->task = task_check_conv;
                    break;
                }

                default:
                    *ircomm = aoclsparse_rci_stop;
                    return aoclsparse_status_internal_error;
                }
            } while(loop);
            if(*ircomm == aoclsparse_rci_stop || *ircomm == aoclsparse_rci_stopping_criterion || exit_status != aoclsparse_status_success)
                B200_TRY(sync_x());
            else
                B200_CUDA(cudaStreamSynchronize(st));
            return exit_status;
        }

        template <typename R>
        aoclsparse_status crci_solve(itsol_cdata<R> *it, aoclsparse_itsol_rci_job *ircomm, void **u, void **v, void *x, R rinfo[100])
        {
            if(ircomm == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(it == nullptr)
            {
                *ircomm = aoclsparse_rci_stop;
                return aoclsparse_status_internal_error;
            }
            if(u == nullptr || v == nullptr || x == nullptr || rinfo == nullptr || !it->have_b)
            {
                *ircomm = aoclsparse_rci_stop;
                return aoclsparse_status_invalid_pointer;
            }
            aoclsparse_status status;
            if(!it->solving)
            {
                if(it->opts.solver != 0)
                {
                    *ircomm = aoclsparse_rci_stop;
                    return aoclsparse_status_not_implemented; // GMRES
                }
                it->task        = task_start;
                it->precond     = it->opts.cg_precond;
                it->rtol        = it->opts.cg_rtol;
                it->atol        = it->opts.cg_atol;
                it->maxit       = it->opts.cg_maxit;
                it->solving     = true;
                it->opts.locked = true;
            }
            status = ccg_rci(it, ircomm, u, v, x, rinfo);
            if(status != aoclsparse_status_success)
                *ircomm = aoclsparse_rci_stop;
            if(*ircomm == aoclsparse_rci_stop)
            {
                it->solving     = false;
                it->opts.locked = false;
            }
            return status;
        }

        inline aoclsparse_status cmv_call(float, aoclsparse_matrix A, const aoclsparse_mat_descr d, const void *x, void *y)
        {
            const aoclsparse_float_complex one{1.0f, 0.0f}, zero{0.0f, 0.0f};
            return aoclsparse_cmv(aoclsparse_operation_none, &one, A, d, static_cast<const aoclsparse_float_complex *>(x), &zero,
                                  static_cast<aoclsparse_float_complex *>(y));
        }
        inline aoclsparse_status cmv_call(double, aoclsparse_matrix A, const aoclsparse_mat_descr d, const void *x, void *y)
        {
            const aoclsparse_double_complex one{1.0, 0.0}, zero{0.0, 0.0};
            return aoclsparse_zmv(aoclsparse_operation_none, &one, A, d, static_cast<const aoclsparse_double_complex *>(x), &zero,
                                  static_cast<aoclsparse_double_complex *>(y));
        }

        // aoclsparse_cg_solve for complex data, itsol_functions.hpp:1369-1500 (same checks in the same order as forward_solve)
        template <typename R>
        aoclsparse_status cforward_solve(itsol_cdata<R>            *it,
                                         aoclsparse_int             n,
                                         aoclsparse_matrix          mat,
                                         const aoclsparse_mat_descr descr,
                                         const void                *b,
                                         void                      *x,
                                         R                          rinfo[100],
                                         aoclsparse_int (*precond)(aoclsparse_int, aoclsparse_int, const void *, void *, void *),
                                         aoclsparse_int (*monit)(aoclsparse_int, const void *, const void *, R *, void *),
                                         void *udata)
        {
            using C = typename cplx_types<R>::dev;
            if(it == nullptr)
                return aoclsparse_status_internal_error;
            if(x == nullptr || rinfo == nullptr)
                return aoclsparse_status_invalid_pointer;
            for(int i = 0; i < 100; ++i)
                rinfo[i] = (R)0;
            B200_TRY(crci_input(it, n, b));
            if(it->opts.solver != 0)
                return aoclsparse_status_not_implemented; // GMRES
            if(mat == nullptr || descr == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(mat->m != n || mat->n != n)
                return aoclsparse_status_invalid_size;
            if(descr->type != aoclsparse_matrix_type_symmetric || descr->fill_mode != aoclsparse_fill_mode_lower)
                return aoclsparse_status_invalid_value;
            if(it->opts.cg_precond == 1 && precond == nullptr)
                return aoclsparse_status_invalid_pointer;
            if(it->opts.cg_precond == 3)
                return aoclsparse_status_not_implemented; // symmetric Gauss-Seidel
            {
                const int want  = get_doid(true, descr->type, descr->fill_mode, aoclsparse_operation_none);
                bool      ready = false;
                {
                    std::shared_lock<std::shared_mutex> rl(mat->guard);
                    for(const hint &hh : mat->hints)
                        ready = ready || (hh.act == 1 && hh.doid == want && hh.done);
                }
                if(!ready && aoclsparse_set_mv_hint(mat, aoclsparse_operation_none, descr, 100) == aoclsparse_status_success)
                    aoclsparse_optimize(mat);
            }
            aoclsparse_itsol_rci_job ircomm      = aoclsparse_rci_start;
            void                    *u = nullptr, *v = nullptr;
            aoclsparse_status        exit_status = aoclsparse_status_success;
            while(ircomm != aoclsparse_rci_stop)
            {
                exit_status = crci_solve(it, &ircomm, &u, &v, x, rinfo);
                if(exit_status != aoclsparse_status_success && ircomm != aoclsparse_rci_stop)
                    break;
                if(ircomm == aoclsparse_rci_mv)
                {
                    if(cmv_call(R(), mat, descr, u, v) != aoclsparse_status_success)
                    {
                        exit_status = aoclsparse_status_internal_error;
                        break;
                    }
                }
                else if(ircomm == aoclsparse_rci_precond)
                {
                    if(precond(0, n, u, v, udata) != 0)
                        ircomm = aoclsparse_rci_interrupt;
                }
                else if(ircomm == aoclsparse_rci_stopping_criterion && monit)
                {
                    if(monit(n, it->x_user ? static_cast<const void *>(it->x_user) : static_cast<const void *>(it->x_dev), u, rinfo, udata) != 0)
                        ircomm = aoclsparse_rci_interrupt;
                }
            }
            it->solving     = false;
            it->opts.locked = false;
            (void)sizeof(C);
            return exit_status;
        }
    }
}

using namespace b200;

struct _aoclsparse_itsol_handle
{
    aoclsparse_matrix_data_type type;
    itsol_data<float>          *s = nullptr;
    itsol_data<double>         *d = nullptr;
    itsol_cdata<float>         *c = nullptr;
    itsol_cdata<double>        *z = nullptr;
};

extern "C" {
aoclsparse_status aoclsparse_itsol_s_init(aoclsparse_itsol_handle *handle)
{
    if(handle == nullptr)
        return aoclsparse_status_invalid_pointer;
    *handle = new(std::nothrow) _aoclsparse_itsol_handle;
    if(!*handle)
        return aoclsparse_status_memory_error;
    (*handle)->type = aoclsparse_smat;
    (*handle)->s    = new(std::nothrow) itsol_data<float>;
    if(!(*handle)->s)
    {
        aoclsparse_itsol_destroy(handle);
        return aoclsparse_status_memory_error;
    }
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_itsol_d_init(aoclsparse_itsol_handle *handle)
{
    if(handle == nullptr)
        return aoclsparse_status_invalid_pointer;
    *handle = new(std::nothrow) _aoclsparse_itsol_handle;
    if(!*handle)
        return aoclsparse_status_memory_error;
    (*handle)->type = aoclsparse_dmat;
    (*handle)->d    = new(std::nothrow) itsol_data<double>;
    if(!(*handle)->d)
    {
        aoclsparse_itsol_destroy(handle);
        return aoclsparse_status_memory_error;
    }
    return aoclsparse_status_success;
}

// complex handles (itsol_functions.cpp:168-230): conjugate gradients only, host-driven (see ccg_rci)
aoclsparse_status aoclsparse_itsol_c_init(aoclsparse_itsol_handle *handle)
{
    if(handle == nullptr)
        return aoclsparse_status_invalid_pointer;
    *handle = new(std::nothrow) _aoclsparse_itsol_handle;
    if(!*handle)
        return aoclsparse_status_memory_error;
    (*handle)->type = aoclsparse_cmat;
    (*handle)->c    = new(std::nothrow) itsol_cdata<float>;
    if(!(*handle)->c)
    {
        aoclsparse_itsol_destroy(handle);
        return aoclsparse_status_memory_error;
    }
    return aoclsparse_status_success;
}
aoclsparse_status aoclsparse_itsol_z_init(aoclsparse_itsol_handle *handle)
{
    if(handle == nullptr)
        return aoclsparse_status_invalid_pointer;
    *handle = new(std::nothrow) _aoclsparse_itsol_handle;
    if(!*handle)
        return aoclsparse_status_memory_error;
    (*handle)->type = aoclsparse_zmat;
    (*handle)->z    = new(std::nothrow) itsol_cdata<double>;
    if(!(*handle)->z)
    {
        aoclsparse_itsol_destroy(handle);
        return aoclsparse_status_memory_error;
    }
    return aoclsparse_status_success;
}

void aoclsparse_itsol_destroy(aoclsparse_itsol_handle *handle)
{
    if(handle && *handle)
    {
        delete(*handle)->s;
        delete(*handle)->d;
        delete(*handle)->c;
        delete(*handle)->z;
        delete *handle;
        *handle = nullptr;
    }
}

// aoclsparse_solvers.h:147: prints the option registry of the handle to the standard output
void aoclsparse_itsol_handle_prn_options(aoclsparse_itsol_handle handle)
{
    if(!handle)
        return;
    if(handle->type == aoclsparse_dmat && handle->d)
        handle->d->opts.print();
    else if(handle->type == aoclsparse_smat && handle->s)
        handle->s->opts.print();
    else if(handle->type == aoclsparse_cmat && handle->c)
        handle->c->opts.print();
    else if(handle->type == aoclsparse_zmat && handle->z)
        handle->z->opts.print();
}

aoclsparse_status aoclsparse_itsol_option_set(aoclsparse_itsol_handle handle, const char *option, const char *value)
{
    // handle_parse_option, itsol_functions.hpp:1622-1715: every failure of the registry maps to invalid_value
    if(handle == nullptr)
        return aoclsparse_status_invalid_pointer;
    if((handle->type == aoclsparse_dmat && !handle->d) || (handle->type == aoclsparse_smat && !handle->s)
       || (handle->type == aoclsparse_cmat && !handle->c) || (handle->type == aoclsparse_zmat && !handle->z))
        return aoclsparse_status_internal_error;
    if(!option || !value)
        return aoclsparse_status_invalid_pointer;
    const std::string name = prepare(option);
    int               flag;
    switch(handle->type)
    {
    case aoclsparse_dmat:
        flag = handle->d->opts.set(name, value);
        break;
    case aoclsparse_smat:
        flag = handle->s->opts.set(name, value);
        break;
    case aoclsparse_cmat:
        flag = handle->c->opts.set(name, value);
        break;
    default:
        flag = handle->z->opts.set(name, value);
        break;
    }
    return flag == 0 ? aoclsparse_status_success : aoclsparse_status_invalid_value;
}

aoclsparse_status aoclsparse_itsol_d_rci_input(aoclsparse_itsol_handle handle, aoclsparse_int n, const double *b)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_dmat)
        return aoclsparse_status_wrong_type;
    return rci_input(handle->d, n, b);
}
aoclsparse_status aoclsparse_itsol_s_rci_input(aoclsparse_itsol_handle handle, aoclsparse_int n, const float *b)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_smat)
        return aoclsparse_status_wrong_type;
    return rci_input(handle->s, n, b);
}

aoclsparse_status aoclsparse_itsol_d_rci_solve(aoclsparse_itsol_handle handle, aoclsparse_itsol_rci_job *ircomm, double **u, double **v, double *x, double rinfo[100])
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_dmat)
        return aoclsparse_status_wrong_type;
    return rci_solve(handle->d, ircomm, u, v, x, rinfo);
}
aoclsparse_status aoclsparse_itsol_s_rci_solve(aoclsparse_itsol_handle handle, aoclsparse_itsol_rci_job *ircomm, float **u, float **v, float *x, float rinfo[100])
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_smat)
        return aoclsparse_status_wrong_type;
    return rci_solve(handle->s, ircomm, u, v, x, rinfo);
}

aoclsparse_status aoclsparse_itsol_d_solve(aoclsparse_itsol_handle    handle,
                                           aoclsparse_int             n,
                                           aoclsparse_matrix          mat,
                                           const aoclsparse_mat_descr descr,
                                           const double              *b,
                                           double                    *x,
                                           double                     rinfo[100],
                                           aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const double *u, double *v, void *udata),
                                           aoclsparse_int monit(aoclsparse_int n, const double *x, const double *r, double rinfo[100], void *udata),
                                           void *udata)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_dmat)
        return aoclsparse_status_wrong_type;
    return forward_solve<double>(handle->d, n, mat, descr, b, x, rinfo, precond, monit, udata);
}
aoclsparse_status aoclsparse_itsol_s_solve(aoclsparse_itsol_handle    handle,
                                           aoclsparse_int             n,
                                           aoclsparse_matrix          mat,
                                           const aoclsparse_mat_descr descr,
                                           const float               *b,
                                           float                     *x,
                                           float                      rinfo[100],
                                           aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const float *u, float *v, void *udata),
                                           aoclsparse_int monit(aoclsparse_int n, const float *x, const float *r, float rinfo[100], void *udata),
                                           void *udata)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_smat)
        return aoclsparse_status_wrong_type;
    return forward_solve<float>(handle->s, n, mat, descr, b, x, rinfo, precond, monit, udata);
}

// ---- complex entry points (aoclsparse_solvers.h:276-283, 395-408, 537-575)
aoclsparse_status aoclsparse_itsol_c_rci_input(aoclsparse_itsol_handle handle, aoclsparse_int n, const aoclsparse_float_complex *b)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_cmat)
        return aoclsparse_status_wrong_type;
    return crci_input(handle->c, n, b);
}
aoclsparse_status aoclsparse_itsol_z_rci_input(aoclsparse_itsol_handle handle, aoclsparse_int n, const aoclsparse_double_complex *b)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_zmat)
        return aoclsparse_status_wrong_type;
    return crci_input(handle->z, n, b);
}
aoclsparse_status aoclsparse_itsol_c_rci_solve(aoclsparse_itsol_handle    handle,
                                               aoclsparse_itsol_rci_job  *ircomm,
                                               aoclsparse_float_complex **u,
                                               aoclsparse_float_complex **v,
                                               aoclsparse_float_complex  *x,
                                               float                      rinfo[100])
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_cmat)
        return aoclsparse_status_wrong_type;
    return crci_solve(handle->c, ircomm, reinterpret_cast<void **>(u), reinterpret_cast<void **>(v), x, rinfo);
}
aoclsparse_status aoclsparse_itsol_z_rci_solve(aoclsparse_itsol_handle     handle,
                                               aoclsparse_itsol_rci_job   *ircomm,
                                               aoclsparse_double_complex **u,
                                               aoclsparse_double_complex **v,
                                               aoclsparse_double_complex  *x,
                                               double                      rinfo[100])
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_zmat)
        return aoclsparse_status_wrong_type;
    return crci_solve(handle->z, ircomm, reinterpret_cast<void **>(u), reinterpret_cast<void **>(v), x, rinfo);
}
aoclsparse_status aoclsparse_itsol_c_solve(
    aoclsparse_itsol_handle         handle,
    aoclsparse_int                  n,
    aoclsparse_matrix               mat,
    const aoclsparse_mat_descr      descr,
    const aoclsparse_float_complex *b,
    aoclsparse_float_complex       *x,
    float                           rinfo[100],
    aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const aoclsparse_float_complex *u, aoclsparse_float_complex *v, void *udata),
    aoclsparse_int monit(aoclsparse_int n, const aoclsparse_float_complex *x, const aoclsparse_float_complex *r, float rinfo[100], void *udata),
    void *udata)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_cmat)
        return aoclsparse_status_wrong_type;
    return cforward_solve<float>(handle->c, n, mat, descr, b, x, rinfo,
                                 reinterpret_cast<aoclsparse_int (*)(aoclsparse_int, aoclsparse_int, const void *, void *, void *)>(precond),
                                 reinterpret_cast<aoclsparse_int (*)(aoclsparse_int, const void *, const void *, float *, void *)>(monit), udata);
}
aoclsparse_status aoclsparse_itsol_z_solve(
    aoclsparse_itsol_handle          handle,
    aoclsparse_int                   n,
    aoclsparse_matrix                mat,
    const aoclsparse_mat_descr       descr,
    const aoclsparse_double_complex *b,
    aoclsparse_double_complex       *x,
    double                           rinfo[100],
    aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const aoclsparse_double_complex *u, aoclsparse_double_complex *v, void *udata),
    aoclsparse_int monit(aoclsparse_int n, const aoclsparse_double_complex *x, const aoclsparse_double_complex *r, double rinfo[100], void *udata),
    void *udata)
{
    if(!handle)
        return aoclsparse_status_invalid_pointer;
    if(handle->type != aoclsparse_zmat)
        return aoclsparse_status_wrong_type;
    return cforward_solve<double>(handle->z, n, mat, descr, b, x, rinfo,
                                  reinterpret_cast<aoclsparse_int (*)(aoclsparse_int, aoclsparse_int, const void *, void *, void *)>(precond),
                                  reinterpret_cast<aoclsparse_int (*)(aoclsparse_int, const void *, const void *, double *, void *)>(monit), udata);
}
}
