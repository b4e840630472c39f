// common.hpp -- handle layout, small utilities and the value-type algebra shared by every
// translation unit of libaoclsparse_b200.so.
//
// Reference counterparts (for orientation only; nothing here is derived from their text):
//   _aoclsparse_mat_descr      library/src/include/aoclsparse_descr.h:36-46
//   _aoclsparse_matrix, csr    library/src/include/aoclsparse_mat_structures.hpp:774-859,147-251
//   aoclsparse_optimize_data   library/src/include/aoclsparse_mat_structures.hpp:53-68
#pragma once
#include "aoclsparse.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <shared_mutex>
#include <memory>
#include <vector>

namespace b200
{
    // ---------------------------------------------------------------- error plumbing
    void               note_cuda_error(cudaError_t e, const char *where);
    extern std::atomic<unsigned long long> g_launches;

    inline aoclsparse_status cuda_status(cudaError_t e, const char *where)
    {
        if(e == cudaSuccess)
            return aoclsparse_status_success;
        note_cuda_error(e, where);
        return (e == cudaErrorMemoryAllocation) ? aoclsparse_status_memory_error
                                                : aoclsparse_status_internal_error;
    }

#define B200_STR2(x) #x
#define B200_STR(x) B200_STR2(x)
#define B200_CUDA(call)                                                                     \
    do                                                                                      \
    {                                                                                       \
        cudaError_t e__ = (call);                                                           \
        if(e__ != cudaSuccess)                                                              \
            return ::b200::cuda_status(e__, __FILE__ ":" B200_STR(__LINE__));               \
    } while(0)
#define B200_TRY(call)                                                                      \
    do                                                                                      \
    {                                                                                       \
        aoclsparse_status s__ = (call);                                                     \
        if(s__ != aoclsparse_status_success)                                                \
            return s__;                                                                     \
    } while(0)
    // call after every kernel launch
#define B200_LAUNCHED()                                                                     \
    do                                                                                      \
    {                                                                                       \
        ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
        B200_CUDA(cudaGetLastError());                                                      \
    } while(0)

    cudaStream_t current_stream();

    // ---------------------------------------------------------------- device memory
    // Owning device allocation.  Allocations are padded so that the 16-byte bulk copies of the
    // multiply kernels may read a few elements past the logical end.
    struct dev_buf
    {
        void  *p     = nullptr;
        size_t bytes = 0;
        dev_buf()    = default;
        dev_buf(const dev_buf &)            = delete;
        dev_buf &operator=(const dev_buf &) = delete;
        dev_buf(dev_buf &&o) noexcept
        {
            p       = o.p;
            bytes   = o.bytes;
            o.p     = nullptr;
            o.bytes = 0;
        }
        dev_buf &operator=(dev_buf &&o) noexcept
        {
            if(this != &o)
            {
                release();
                p       = o.p;
                bytes   = o.bytes;
                o.p     = nullptr;
                o.bytes = 0;
            }
            return *this;
        }
        ~dev_buf()
        {
            release();
        }
        aoclsparse_status alloc(size_t n_bytes)
        {
            release();
            size_t padded = ((n_bytes + 255) / 256) * 256 + 256;
            cudaError_t e = cudaMalloc(&p, padded);
            if(e != cudaSuccess)
            {
                p = nullptr;
                return cuda_status(e, "cudaMalloc");
            }
            bytes = n_bytes;
            return aoclsparse_status_success;
        }
        void release()
        {
            if(p)
                cudaFree(p);
            p     = nullptr;
            bytes = 0;
        }
        template <typename T>
        T *as() const
        {
            return static_cast<T *>(p);
        }
    };

    // true if the pointer can be dereferenced by a kernel running on the current device
    bool is_device_accessible(const void *p);
    void *pinned_host_device_ptr(const void *p); // device address of page-locked host memory, else nullptr

    // ---------------------------------------------------------------- value-type algebra
    template <typename T>
    struct vt; // value traits

    template <>
    struct vt<float>
    {
        using real                               = float;
        static constexpr bool is_complex         = false;
        static constexpr int  data_type          = aoclsparse_smat;
        static __host__ __device__ float zero()
        {
            return 0.f;
        }
        static __host__ __device__ float one()
        {
            return 1.f;
        }
    };
    template <>
    struct vt<double>
    {
        using real                               = double;
        static constexpr bool is_complex         = false;
        static constexpr int  data_type          = aoclsparse_dmat;
        static __host__ __device__ double zero()
        {
            return 0.0;
        }
        static __host__ __device__ double one()
        {
            return 1.0;
        }
    };
    template <>
    struct vt<float2>
    {
        using real                               = float;
        static constexpr bool is_complex         = true;
        static constexpr int  data_type          = aoclsparse_cmat;
        static __host__ __device__ float2 zero()
        {
            return make_float2(0.f, 0.f);
        }
        static __host__ __device__ float2 one()
        {
            return make_float2(1.f, 0.f);
        }
    };
    template <>
    struct vt<double2>
    {
        using real                               = double;
        static constexpr bool is_complex         = true;
        static constexpr int  data_type          = aoclsparse_zmat;
        static __host__ __device__ double2 zero()
        {
            return make_double2(0.0, 0.0);
        }
        static __host__ __device__ double2 one()
        {
            return make_double2(1.0, 0.0);
        }
    };

    // acc + a*b
    __host__ __device__ inline float mad(float a, float b, float acc)
    {
#ifdef __CUDA_ARCH__
        return fmaf(a, b, acc);
#else
        return a * b + acc;
#endif
    }
    __host__ __device__ inline double mad(double a, double b, double acc)
    {
#ifdef __CUDA_ARCH__
        return fma(a, b, acc);
#else
        return a * b + acc;
#endif
    }
    __host__ __device__ inline float2 mad(float2 a, float2 b, float2 acc)
    {
        acc.x = mad(a.x, b.x, acc.x);
        acc.x = mad(-a.y, b.y, acc.x);
        acc.y = mad(a.x, b.y, acc.y);
        acc.y = mad(a.y, b.x, acc.y);
        return acc;
    }
    __host__ __device__ inline double2 mad(double2 a, double2 b, double2 acc)
    {
        acc.x = mad(a.x, b.x, acc.x);
        acc.x = mad(-a.y, b.y, acc.x);
        acc.y = mad(a.x, b.y, acc.y);
        acc.y = mad(a.y, b.x, acc.y);
        return acc;
    }
    __host__ __device__ inline float mul(float a, float b)
    {
        return a * b;
    }
    __host__ __device__ inline double mul(double a, double b)
    {
        return a * b;
    }
    __host__ __device__ inline float2 mul(float2 a, float2 b)
    {
        return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    __host__ __device__ inline double2 mul(double2 a, double2 b)
    {
        return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    __host__ __device__ inline float add(float a, float b)
    {
        return a + b;
    }
    __host__ __device__ inline double add(double a, double b)
    {
        return a + b;
    }
    __host__ __device__ inline float2 add(float2 a, float2 b)
    {
        return make_float2(a.x + b.x, a.y + b.y);
    }
    __host__ __device__ inline double2 add(double2 a, double2 b)
    {
        return make_double2(a.x + b.x, a.y + b.y);
    }
    __host__ __device__ inline float cj(float a)
    {
        return a;
    }
    __host__ __device__ inline double cj(double a)
    {
        return a;
    }
    __host__ __device__ inline float2 cj(float2 a)
    {
        return make_float2(a.x, -a.y);
    }
    __host__ __device__ inline double2 cj(double2 a)
    {
        return make_double2(a.x, -a.y);
    }
    __host__ __device__ inline bool is_zero(float a)
    {
        return a == 0.f;
    }
    __host__ __device__ inline bool is_zero(double a)
    {
        return a == 0.0;
    }
    __host__ __device__ inline bool is_zero(float2 a)
    {
        return a.x == 0.f && a.y == 0.f;
    }
    __host__ __device__ inline bool is_zero(double2 a)
    {
        return a.x == 0.0 && a.y == 0.0;
    }
    __host__ __device__ inline bool is_one(float a)
    {
        return a == 1.f;
    }
    __host__ __device__ inline bool is_one(double a)
    {
        return a == 1.0;
    }
    __host__ __device__ inline bool is_one(float2 a)
    {
        return a.x == 1.f && a.y == 0.f;
    }
    __host__ __device__ inline bool is_one(double2 a)
    {
        return a.x == 1.0 && a.y == 0.0;
    }

    inline size_t value_size(int val_type)
    {
        switch(val_type)
        {
        case aoclsparse_dmat:
            return 8;
        case aoclsparse_smat:
            return 4;
        case aoclsparse_cmat:
            return 8;
        default:
            return 16;
        }
    }

    // ---------------------------------------------------------------- dispatch ids
    // Same numbering as aoclsparse::doid (aoclsparse_mtx_dispatcher.hpp:41-74): [group:3][op:2].
    enum : int
    {
        DOID_GN = 0,
        DOID_GC = 1,
        DOID_GT = 2,
        DOID_GH = 3,
        DOID_LEN = 20
    };
    int get_doid(bool complex_type, int descr_type, int fill_mode, int op);

    // ---------------------------------------------------------------- plan (analysis result)
    constexpr int CODE_TABLE_MAX = 256; // entries of the diagonal-code table (one byte per code)
    enum : int
    {
        STRAT_THREAD  = 0, // one thread per row, operands read from the staged chunk
        STRAT_WARP    = 1, // one warp (or sub-warp) per row
        STRAT_PRODUCT = 2, // CTA-wide element products, then per-row segmented sums
        STRAT_LONG    = 3  // block is one segment of a row split across several CTAs
    };

    struct row_block_plan
    {
        aoclsparse_int block_nnz  = 0; // nnz capacity T of a row block
        aoclsparse_int block_rows = 0; // row capacity R of a row block
        aoclsparse_int n_blocks   = 0;
        aoclsparse_int n_long_rows = 0;
        aoclsparse_int n_long_segments = 0;
        aoclsparse_int n_strat[4] = {0, 0, 0, 0};
        aoclsparse_int max_block_rows = 0; // most rows any block holds (sizes the staged row_ptr slice)
        aoclsparse_int max_block_nnz  = 0; // most entries any non-split block holds (<= block_nnz; sizes entry-code buffers)
        int            pdl              = 1; // programmatic dependent launch of the multiply kernel (tuning knob)
        // diagonal-code copy of col_idx (plan.cu, build_diag_codes): one byte per stored entry indexing the table of the
        // matrix's distinct (col - row) offsets.  Built by aoclsparse_optimize when every block is thread-per-row and
        // there are at most 256 distinct offsets (stencils, banded matrices); the multiply then streams 1 instead of 4
        // index bytes per entry and decodes the identical column.
        aoclsparse_int n_codes = 0; // table entries in use; 0 = not built / not applicable
        int            code_state = 0; // 0 not analysed, 1 analysed (built or found not applicable)
        dev_buf        codes;        // unsigned char[nnz]
        dev_buf        code_offsets; // int[256], ascending, unused tail repeats the last entry
        // entry-code copy (plan.cu, build_entry_codes), only next to the diagonal-code copy: one byte per stored entry
        // indexing the table of the matrix's distinct (col - row, value) pairs (<= 256; a constant-coefficient stencil has
        // as many as it has points).  The multiply then streams 1 instead of 4 + sizeof(T) bytes per entry and decodes the
        // identical column and value.  Depends on the VALUES: marked stale when they change and rebuilt before the next
        // multiply (api.cu, ensure_plan).
        aoclsparse_int n_ecodes     = 0;     // table entries in use; 0 = not built / not applicable
        bool           ecodes_stale = false; // the stored values changed since the copy was built
        dev_buf        ecodes;               // unsigned char[nnz]
        dev_buf        etab_off;             // int[256]: col - row of every pair (ascending; ties by value pattern)
        dev_buf        etab_val;             // T[256]: value of every pair
        // the entry-coded kernels run on a block plan of their own (one staged byte per entry: much larger blocks,
        // bounded by rows); everything else -- other kernels, sub-range launches -- keeps this plan
        std::unique_ptr<row_block_plan> eplan;
        int            threads     = 256; // CTA size of the multiply kernel (tuning knob)
        int            stream_hint = 1;   // tag the val/col stream evict-first in L2 (tuning knob)
        dev_buf        desc;      // int4 per block: first row, end row, first nnz, end nnz
        dev_buf        kind;      // int per block: strategy | slot << 4
        dev_buf        long_rows; // int4 per long row: row, first slot, n segments, unused
        dev_buf        partials;  // one value per long segment
        std::vector<aoclsparse_int> cut_block; // block index at which each row cut starts (+ ends)
        // host-staged multiplies (x, y in host memory): the blocks cut into a few chunks with the x prefix each needs,
        // so that H2D of x, the kernels and D2H of y pipeline on three streams (spmv.cu: mv_host_pipelined)
        struct host_chunk
        {
            aoclsparse_int b0, b1, row0, row1, x_hi; // blocks, rows, one past the largest column used so far
        };
        std::vector<host_chunk> host_chunks;
        bool                    host_chunks_ready = false;
        bool           valid = false;
    };

    // box tiles of a lattice-structured (stencil) matrix for csrmm: see mesh_tiles.cu
    struct mesh_tiles
    {
        int       state = 0;             // 0 not analysed, 1 analysed and not usable, 2 ready
        int       box[3]    = {0, 0, 0}; // tile extent along the three grid directions
        long long stride[3] = {0, 0, 0}; // row-index strides of the directions (stride[0] == 1)
        int       dims[3]   = {0, 0, 0}; // grid extent
        int       rows_per_tile = 0, n_tiles = 0, max_distinct = 0, max_walk = 0, max_vals = 0, max_runs = 0;
        size_t    row_bytes    = 0; // n * sizeof(T) the box was sized for
        long long walk_entries = 0; // slots of all walk planes (padding included)
        long long val_entries  = 0; // slots of all value planes (padding included)
        long long n_runs_total = 0; // runs of all tiles + one terminator per tile
        double    reuse = 0.0, fill = 0.0; // stored entries per staged B row; stored entries per value-plane slot
        dev_buf   desc; // int4 per tile: distinct B rows, runs, U | V << 16 (longest walk / value stream of a group), first run
        dev_buf   off;  // 2 long long per tile: first walk slot, first value slot
        dev_buf   walk; // unsigned[walk_entries]: tile t, entry j, group g at off[2t] + j * groups + g: slot | rowmask << 16 | first value << 20
        dev_buf   val;  // T[val_entries]: tile t, i-th value of group g at off[2t+1] + i * groups + g
        dev_buf   rows; // int[n_tiles * rows_per_tile]: matrix row of (tile, lr) or -1
        dev_buf   runs; // int2[n_runs_total]: first column, first slot of every run; terminator (-1, distinct)
    };

    // one device-resident CSR (always 0-based on the device)
    struct dev_csr
    {
        aoclsparse_int m = 0, n = 0, nnz = 0;
        int            doid = DOID_GN; // what this copy represents relative to the user's matrix
        dev_buf        row_ptr, col_idx, val;
        row_block_plan plan;
        // csrmm's tile copy, built on the first row-major multiply that can use it (under tiles_mu; immutable once
        // ready, dropped under the handle's write lock when the values change)
        mutable mesh_tiles tiles;
        mutable std::mutex tiles_mu;
    };

    // "clean CSR" of the reference's analysis (clean.cu): rows grouped lower | diagonal | upper, diagonals present
    struct clean_csr
    {
        bool           valid       = false;
        bool           is_internal = false; // false: the input already is the clean matrix (arrays below unused)
        aoclsparse_int nnz         = 0;
        dev_buf        row_ptr, col_idx, val, idiag, iurow;
    };

    struct hint
    {
        int            act; // 1 mv, 3 mm, ... (aoclsparse_hinted_action numbering)
        int            trans;
        int            type;
        int            fill_mode;
        int            doid;
        aoclsparse_int nop;
        aoclsparse_int kid;
        bool           done = false;
    };
}

struct _aoclsparse_mat_descr
{
    aoclsparse_matrix_type type      = aoclsparse_matrix_type_general;
    aoclsparse_fill_mode   fill_mode = aoclsparse_fill_mode_lower;
    aoclsparse_diag_type   diag_type = aoclsparse_diag_type_non_unit;
    aoclsparse_index_base  base      = aoclsparse_index_base_zero;
};

struct _aoclsparse_matrix
{
    aoclsparse_int                m = 0, n = 0, nnz = 0;
    aoclsparse_index_base         base         = aoclsparse_index_base_zero;
    aoclsparse_matrix_data_type   val_type     = aoclsparse_dmat;
    aoclsparse_matrix_format_type input_format = aoclsparse_csr_mat;
    aoclsparse_matrix_sort        sort         = aoclsparse_unknown_sort;
    bool                          fulldiag     = false;
    aoclsparse_int                min_col = 0, max_col = -1, max_row_nnz = 0;
    aoclsparse_memory_usage       mem_policy = aoclsparse_memory_usage_unrestricted;
    int                           device     = 0;
    std::atomic<int>              lazy_copy_calls{0};   // un-hinted products that would profit from a derived copy (spmv.cu)
    bool                          is_csc     = false; // created from CSC arrays: mats[0] stores the TRANSPOSE (n x m CSR)

    std::vector<b200::hint>       hints; // most recent first, like the reference's linked list
    std::vector<b200::dev_csr *>  mats;  // mats[0] is the user's matrix
    // aoclsparse_sp2m, both operands transposed: the product (B A) whose transpose is this matrix, kept between the
    // nnz_count and finalize stages (and for repeated finalize calls)
    std::unique_ptr<b200::dev_csr> sp2m_product;
    b200::clean_csr               clean; // built on demand (aoclsparse_b200_get_clean_csr)
    std::vector<aoclsparse_int>   row_cuts;
    aoclsparse_int                win_lo = 0, win_hi = -1; // x window (win_hi < 0: whole vector)
    // host mirror handed out by aoclsparse_export_?csr (refreshed by every export call, owned by the handle)
    std::vector<aoclsparse_int>   host_row_ptr, host_col;
    std::vector<unsigned char>    host_val;
    mutable std::shared_mutex     guard;

    ~_aoclsparse_matrix()
    {
        for(auto *c : mats)
            delete c;
    }
};

namespace b200
{
    // check.cu -- GPU restatement of aoclsparse_mat_check_internal
    struct check_result
    {
        aoclsparse_status status;
        int               sort;
        int               fulldiag;
        aoclsparse_int    min_col, max_col, max_row_nnz;
    };
    aoclsparse_status check_csr_device(aoclsparse_int        m,
                                       aoclsparse_int        n,
                                       aoclsparse_int        nnz,
                                       int                   base,
                                       const aoclsparse_int *d_row_ptr,
                                       const aoclsparse_int *d_col,
                                       check_result         &out,
                                       cudaStream_t          st);
    aoclsparse_status rebase_to_zero(aoclsparse_int  m,
                                     aoclsparse_int  nnz,
                                     aoclsparse_int *d_row_ptr,
                                     aoclsparse_int *d_col,
                                     cudaStream_t    st);

    // plan.cu -- row-block analysis
    aoclsparse_status build_plan(dev_csr                           &A,
                                 size_t                             elem_size,
                                 aoclsparse_int                     max_row_nnz, // longest row, < 0 if unknown
                                 aoclsparse_int                     forced_strategy,
                                 const std::vector<aoclsparse_int> &row_cuts,
                                 cudaStream_t                       st,
                                 aoclsparse_int                     block_nnz_override = 0,
                                 int                                coded              = 0, // block size for: 0 val + col, 1 val + column codes, 2 entry codes
                                 row_block_plan                    *target             = nullptr); // default: A.plan
    // plan + (when the matrix has at most 256 distinct col - row offsets and every block comes out thread-per-row) the
    // diagonal-code copy, with the block size that copy wants; falls back to the plain plan otherwise
    aoclsparse_status build_plan_with_codes(dev_csr                           &A,
                                            size_t                             elem_size,
                                            aoclsparse_int                     max_row_nnz,
                                            aoclsparse_int                     forced_strategy,
                                            const std::vector<aoclsparse_int> &row_cuts,
                                            cudaStream_t                       st);
    aoclsparse_status probe_diag_offsets(const dev_csr &A, std::vector<int> &offs, cudaStream_t st);
    // optional second pass of the analysis: the diagonal-code copy of A's column indices (needs a valid plan)
    aoclsparse_status build_diag_codes(dev_csr &A, cudaStream_t st);
    // optional third pass: the entry-code copy (needs the diagonal-code copy); (re)probes the values
    aoclsparse_status build_entry_codes(dev_csr &A, size_t elem_size, aoclsparse_int max_row_nnz, const std::vector<aoclsparse_int> &row_cuts, cudaStream_t st);
    // the sorted distinct bit patterns of A's values (zero-extended to 64 bits), empty when there are more than 256
    aoclsparse_status probe_values(const dev_csr &A, size_t elem_size, std::vector<unsigned long long> &vals, cudaStream_t st);
    void              plan_parameters(size_t          elem_size,
                                      aoclsparse_int  m,
                                      aoclsparse_int  nnz,
                                      aoclsparse_int  max_row_nnz,
                                      aoclsparse_int &block_nnz,
                                      aoclsparse_int &block_rows,
                                      int             coded = 0);

    // transpose.cu -- device csr -> csc (= transposed csr), optional conjugation
    aoclsparse_status transpose_csr(const dev_csr &A, int val_type, bool conj, dev_csr &out, cudaStream_t st);

    // expand.cu -- general CSR copy F (or conj F) of a symmetric / hermitian matrix described by one
    // stored triangle; built on first use, cached in the handle.  op folds in as in get_doid.
    aoclsparse_status get_expanded_copy(aoclsparse_matrix            A,
                                        const _aoclsparse_mat_descr &descr,
                                        aoclsparse_operation         op,
                                        const dev_csr              *&out,
                                        cudaStream_t                 st);

    // api.cu -- device handle for the handle-free legacy entry (no structural rejection of duplicates)
    aoclsparse_status create_temp_csr(aoclsparse_matrix    *mat,
                                      int                   val_type,
                                      aoclsparse_index_base base,
                                      aoclsparse_int        M,
                                      aoclsparse_int        N,
                                      aoclsparse_int        nnz,
                                      const aoclsparse_int *row_ptr,
                                      const aoclsparse_int *col_idx,
                                      const void           *val);

    // clean.cu
    aoclsparse_status ensure_clean(aoclsparse_matrix A, cudaStream_t st);

    // api.cu -- after the stored VALUES changed: derived copies (transposed / expanded / clean) hold stale values and are
    // dropped, hints become pending again; the row-block plan depends on the pattern only and stays (caller holds the
    // write lock)
    void drop_derived_copies(aoclsparse_matrix A);

    // spmv.cu -- one launch of the fused multiply + halo push + flags kernel (aoclsparse_b200_dmv_sharded_step)
    aoclsparse_status sharded_step_launch(const double                   *alpha,
                                          aoclsparse_matrix               A,
                                          const aoclsparse_mat_descr      descr,
                                          const double                   *x,
                                          double                         *y,
                                          const aoclsparse_b200_halo_ctl *ctl,
                                          unsigned                        kc);

    // spmv.cu -- k iterations in one cooperative launch (spmv_sharded_iterate_kernel); see shard.cu
    struct sharded_iterate_args
    {
        const void *left_done = nullptr, *right_done = nullptr;     // local flags the neighbours write
        void       *to_left_done = nullptr, *to_right_done = nullptr; // the neighbours' flags
        void       *counters = nullptr;                              // 4 words: boundary counts, grid barrier, timeout
        void       *push_left[2] = {nullptr, nullptr}, *push_right[2] = {nullptr, nullptr}; // [0] neighbour's window `cur`, [1] `nxt`
        unsigned    k0 = 0, kc0 = 0, bar0 = 0;
    };
    bool              sharded_prefers_steps(aoclsparse_matrix A);
    aoclsparse_status sharded_iterate_launch(double                      alpha,
                                             aoclsparse_matrix           A,
                                             const aoclsparse_mat_descr  descr,
                                             double                     *w_cur,
                                             double                     *w_nxt,
                                             long long                   own_offset,
                                             const sharded_iterate_args &args,
                                             int                         iterations,
                                             int                        *grid_out);

    // mesh_tiles.cu
    bool              detect_lattice(const std::vector<int> &offs, long long m, long long &s1, long long &s2, int &ndim);
    aoclsparse_status build_mesh_tiles(const dev_csr &A, size_t elem_size, size_t row_bytes, cudaStream_t st);
    size_t            mesh_tiles_smem(const mesh_tiles &M, size_t row_bytes, size_t elem_size);
    template <typename T>
    aoclsparse_status launch_mm_tiles(const dev_csr &A, const T *B, long long ldb, T *C, long long ldc, int n, T alpha, T beta, cudaStream_t st);

    // obtains (building on first use) the plan of mats[0]
    aoclsparse_status ensure_plan(aoclsparse_matrix A, cudaStream_t st);
}
