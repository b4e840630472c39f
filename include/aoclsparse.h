/* aoclsparse.h -- C-ABI drop-in boundary of the B200-native CSR SpMV / SpMM path.
 *
 * Every symbol below has the SAME name, argument order, argument meaning, enum values and
 * status codes as the AOCL-Sparse v5.3.2 entry it replaces, so that a program written against
 * the reference's <aoclsparse.h> for this path recompiles and relinks against
 * libaoclsparse_b200.so unchanged.  Each block cites the reference declaration it stands in
 * for (paths relative to the reference tree).  Semantics that differ because the matrix lives
 * in GPU memory are marked "B200:".
 *
 * The numeric enum values are part of the ABI (reference: library/include/aoclsparse_types.h).
 */
#ifndef AOCLSPARSE_H_B200_
#define AOCLSPARSE_H_B200_

#include <stddef.h>
#include <stdint.h>

#ifndef DLL_PUBLIC
#define DLL_PUBLIC __attribute__((__visibility__("default")))
#endif

/* aoclsparse_types.h:54-58 -- the reference is built either LP64 (aoclsparse_int = int32, the default) or ILP64
 * (-Daoclsparse_ILP64: int64).  This library is the drop-in for the LP64 build ONLY: the device kernels stage 32-bit
 * column indices and 32-bit row-pointer windows (int4 block descriptors, 4-byte TMA granules), which is what lets every
 * BASELINE configuration stream 4 (or 1, with the diagonal-code copy) index bytes per stored entry.  The largest
 * configuration (937 951 232 entries) fits; byte offsets are formed in 64 bits.  A caller compiled for the ILP64
 * variant would pass 64-bit arrays to entry points that read 32-bit ones, so that combination is refused here at
 * compile time rather than failing at run time.  An ILP64 variant would be a second set of kernel instantiations (not a
 * recompile), and matrices that need it (> 2^31-1 entries, > 25 GB of indices) are sharded by rows first (shard.cu). */
#if defined(aoclsparse_ILP64)
#error "libaoclsparse_b200 replaces the LP64 build of AOCL-Sparse (aoclsparse_int = int32_t); do not define aoclsparse_ILP64"
#endif
typedef int32_t aoclsparse_int;

/* aoclsparse_types.h:77-99 -- interleaved (re, im) pairs, layout-compatible with C99 / std::complex. */
typedef struct
{
    float real;
    float imag;
} aoclsparse_float_complex;
typedef struct
{
    double real;
    double imag;
} aoclsparse_double_complex;

/* aoclsparse_types.h:114,140 -- opaque handles. */
typedef struct _aoclsparse_mat_descr *aoclsparse_mat_descr;
typedef struct _aoclsparse_matrix    *aoclsparse_matrix;

#ifdef __cplusplus
extern "C" {
#endif

/* aoclsparse_types.h:151-160 */
typedef enum aoclsparse_operation_
{
    aoclsparse_operation_none                = 111,
    aoclsparse_operation_transpose           = 112,
    aoclsparse_operation_conjugate_transpose = 113
} aoclsparse_operation;

/* aoclsparse_types.h:166-169 */
typedef enum aoclsparse_index_base_
{
    aoclsparse_index_base_zero = 0,
    aoclsparse_index_base_one  = 1
} aoclsparse_index_base;

/* aoclsparse_types.h:175-187 */
typedef enum aoclsparse_matrix_type_
{
    aoclsparse_matrix_type_general    = 0,
    aoclsparse_matrix_type_symmetric  = 1,
    aoclsparse_matrix_type_hermitian  = 2,
    aoclsparse_matrix_type_triangular = 3
} aoclsparse_matrix_type;

/* aoclsparse_types.h:193-199 */
typedef enum aoclsparse_matrix_data_type_
{
    aoclsparse_dmat = 0,
    aoclsparse_smat = 1,
    aoclsparse_cmat = 2,
    aoclsparse_zmat = 3
} aoclsparse_matrix_data_type;

/* aoclsparse_types.h:215-240 -- only aoclsparse_csr_mat is produced by this library. */
typedef enum aoclsparse_matrix_format_type_
{
    aoclsparse_csr_mat          = 0,
    aoclsparse_ell_mat          = 1,
    aoclsparse_ellt_mat         = 2,
    aoclsparse_ellt_csr_hyb_mat = 3,
    aoclsparse_ell_csr_hyb_mat  = 4,
    aoclsparse_dia_mat          = 5,
    aoclsparse_csr_mat_br4      = 6,
    aoclsparse_coo_mat          = 7,
    aoclsparse_tcsr_mat         = 8,
    aoclsparse_blkcsr_mat       = 9,
    aoclsparse_bsr_mat          = 10,
    aoclsparse_uninitialized_mat
} aoclsparse_matrix_format_type;

/* aoclsparse_types.h:248-257 */
typedef enum aoclsparse_diag_type_
{
    aoclsparse_diag_type_non_unit = 0,
    aoclsparse_diag_type_unit     = 1,
    aoclsparse_diag_type_zero     = 2
} aoclsparse_diag_type;

/* aoclsparse_types.h:265-269 */
typedef enum aoclsparse_fill_mode_
{
    aoclsparse_fill_mode_lower = 0,
    aoclsparse_fill_mode_upper = 1
} aoclsparse_fill_mode;

/* aoclsparse_types.h:276-280 */
typedef enum aoclsparse_order_
{
    aoclsparse_order_row    = 0,
    aoclsparse_order_column = 1
} aoclsparse_order;

/* aoclsparse_types.h:304-324 */
typedef enum aoclsparse_status_
{
    aoclsparse_status_success             = 0,
    aoclsparse_status_not_implemented     = 1,
    aoclsparse_status_invalid_pointer     = 2,
    aoclsparse_status_invalid_size        = 3,
    aoclsparse_status_internal_error      = 4,
    aoclsparse_status_invalid_value       = 5,
    aoclsparse_status_invalid_index_value = 6,
    aoclsparse_status_maxit               = 7,
    aoclsparse_status_user_stop           = 8,
    aoclsparse_status_wrong_type          = 9,
    aoclsparse_status_memory_error        = 10,
    aoclsparse_status_numerical_error     = 11,
    aoclsparse_status_invalid_operation   = 12,
    aoclsparse_status_unsorted_input      = 13,
    aoclsparse_status_invalid_kid         = 14
} aoclsparse_status;

/* aoclsparse_types.h:331-343 (only used by the sparse x sparse entry, aoclsparse_spmm's relatives) */
typedef enum aoclsparse_request_
{
    aoclsparse_stage_nnz_count        = 0,
    aoclsparse_stage_finalize         = 1,
    aoclsparse_stage_full_computation = 2
} aoclsparse_request;

/* aoclsparse_types.h:384-389 */
typedef enum aoclsparse_memory_usage_
{
    aoclsparse_memory_usage_minimal      = 0,
    aoclsparse_memory_usage_unrestricted = 1
} aoclsparse_memory_usage;

/* aoclsparse_types.h:396-403 -- classification computed by aoclsparse_create_?csr. */
typedef enum aoclsparse_matrix_sort_
{
    aoclsparse_unknown_sort     = 0,
    aoclsparse_fully_sorted     = 1,
    aoclsparse_partially_sorted = 2,
    aoclsparse_unsorted         = 3
} aoclsparse_matrix_sort;

/* ------------------------------------------------------------------------------------------
 * Version string.  Replaces aoclsparse_get_version (aoclsparse_auxiliary.h:47).
 * ---------------------------------------------------------------------------------------- */
DLL_PUBLIC const char *aoclsparse_get_version(void);

/* ------------------------------------------------------------------------------------------
 * Matrix descriptor.  Replaces aoclsparse_auxiliary.h:131-309
 * (implementation: library/src/extra/aoclsparse_auxiliary.cpp:191-362).
 * Defaults: general / lower / non_unit / base zero.  Getters return the default on NULL.
 * ---------------------------------------------------------------------------------------- */
DLL_PUBLIC aoclsparse_status aoclsparse_create_mat_descr(aoclsparse_mat_descr *descr);
DLL_PUBLIC aoclsparse_status aoclsparse_copy_mat_descr(aoclsparse_mat_descr       dest,
                                                       const aoclsparse_mat_descr src);
DLL_PUBLIC aoclsparse_status aoclsparse_destroy_mat_descr(aoclsparse_mat_descr descr);
DLL_PUBLIC aoclsparse_status aoclsparse_set_mat_index_base(aoclsparse_mat_descr  descr,
                                                           aoclsparse_index_base base);
DLL_PUBLIC aoclsparse_index_base aoclsparse_get_mat_index_base(const aoclsparse_mat_descr descr);
DLL_PUBLIC aoclsparse_status     aoclsparse_set_mat_type(aoclsparse_mat_descr   descr,
                                                         aoclsparse_matrix_type type);
DLL_PUBLIC aoclsparse_matrix_type aoclsparse_get_mat_type(const aoclsparse_mat_descr descr);
DLL_PUBLIC aoclsparse_status      aoclsparse_set_mat_fill_mode(aoclsparse_mat_descr descr,
                                                               aoclsparse_fill_mode fill_mode);
DLL_PUBLIC aoclsparse_fill_mode aoclsparse_get_mat_fill_mode(const aoclsparse_mat_descr descr);
DLL_PUBLIC aoclsparse_status    aoclsparse_set_mat_diag_type(aoclsparse_mat_descr descr,
                                                             aoclsparse_diag_type diag_type);
DLL_PUBLIC aoclsparse_diag_type aoclsparse_get_mat_diag_type(const aoclsparse_mat_descr descr);

/* ------------------------------------------------------------------------------------------
 * CSR matrix handle.  Replaces aoclsparse_create_{s,d,c,z}csr (aoclsparse_auxiliary.h:399-437;
 * implementation library/src/create/aoclsparse_create.cpp:33-95) and aoclsparse_destroy
 * (aoclsparse_auxiliary.h:836).
 *
 * Validation and its precedence are those of aoclsparse_mat_check_internal
 * (library/src/analysis/aoclsparse_csr_util.cpp:124-279): NULL arrays -> invalid_pointer;
 * negative M/N/nnz -> invalid_size; row_ptr[0]!=base, row_ptr[M]-base!=nnz, decreasing row_ptr ->
 * invalid_value; then, in storage order, the first of {column out of [0,N) -> invalid_index_value,
 * second diagonal entry in a row -> invalid_value}.  *mat is set to NULL before validating.
 *
 * B200: the reference handle aliases the caller's host arrays; this handle uploads them once and
 * owns device-resident copies (row_ptr, col_idx, val) which aoclsparse_destroy frees.  row_ptr /
 * col_idx / val may be host pointers or CUDA device/managed pointers.  Later edits to the caller's
 * arrays are not seen; use aoclsparse_?update_values to change values.
 * ---------------------------------------------------------------------------------------- */
DLL_PUBLIC aoclsparse_status aoclsparse_create_scsr(aoclsparse_matrix    *mat,
                                                    aoclsparse_index_base base,
                                                    aoclsparse_int        M,
                                                    aoclsparse_int        N,
                                                    aoclsparse_int        nnz,
                                                    aoclsparse_int       *row_ptr,
                                                    aoclsparse_int       *col_idx,
                                                    float                *val);
DLL_PUBLIC aoclsparse_status aoclsparse_create_dcsr(aoclsparse_matrix    *mat,
                                                    aoclsparse_index_base base,
                                                    aoclsparse_int        M,
                                                    aoclsparse_int        N,
                                                    aoclsparse_int        nnz,
                                                    aoclsparse_int       *row_ptr,
                                                    aoclsparse_int       *col_idx,
                                                    double               *val);
DLL_PUBLIC aoclsparse_status aoclsparse_create_ccsr(aoclsparse_matrix        *mat,
                                                    aoclsparse_index_base     base,
                                                    aoclsparse_int            M,
                                                    aoclsparse_int            N,
                                                    aoclsparse_int            nnz,
                                                    aoclsparse_int           *row_ptr,
                                                    aoclsparse_int           *col_idx,
                                                    aoclsparse_float_complex *val);
DLL_PUBLIC aoclsparse_status aoclsparse_create_zcsr(aoclsparse_matrix         *mat,
                                                    aoclsparse_index_base      base,
                                                    aoclsparse_int             M,
                                                    aoclsparse_int             N,
                                                    aoclsparse_int             nnz,
                                                    aoclsparse_int            *row_ptr,
                                                    aoclsparse_int            *col_idx,
                                                    aoclsparse_double_complex *val);

/* CSC input: the M x N matrix is given by columns (col_ptr of N+1 entries, row_idx, val).  Replaces
 * aoclsparse_create_?csc (aoclsparse_auxiliary.h:843-917; aoclsparse_create_csc_t,
 * library/src/extra/aoclsparse_auxiliary.cpp:1030-1089).  As in the reference the arrays are validated
 * and kept as the CSR of the transpose (N rows, M columns); aoclsparse_?mv (every descriptor type and
 * operation), aoclsparse_?dotmv, aoclsparse_?csrmm (general / symmetric / hermitian),
 * aoclsparse_?set_value and aoclsparse_?update_values then work on the handle exactly as on a CSR one
 * (aoclsparse_mv.cpp:160-200 and aoclsparse_csrmm.hpp:496-555 describe the transposition rules).
 * B200-only extensions that need the rows of A (x windows, row cuts, ?mv_rows, the sharded step)
 * return aoclsparse_status_not_implemented for a CSC handle. */
DLL_PUBLIC aoclsparse_status aoclsparse_create_scsc(aoclsparse_matrix    *mat,
                                                    aoclsparse_index_base base,
                                                    aoclsparse_int        M,
                                                    aoclsparse_int        N,
                                                    aoclsparse_int        nnz,
                                                    aoclsparse_int       *col_ptr,
                                                    aoclsparse_int       *row_idx,
                                                    float                *val);
DLL_PUBLIC aoclsparse_status aoclsparse_create_dcsc(aoclsparse_matrix    *mat,
                                                    aoclsparse_index_base base,
                                                    aoclsparse_int        M,
                                                    aoclsparse_int        N,
                                                    aoclsparse_int        nnz,
                                                    aoclsparse_int       *col_ptr,
                                                    aoclsparse_int       *row_idx,
                                                    double               *val);
DLL_PUBLIC aoclsparse_status aoclsparse_create_ccsc(aoclsparse_matrix        *mat,
                                                    aoclsparse_index_base     base,
                                                    aoclsparse_int            M,
                                                    aoclsparse_int            N,
                                                    aoclsparse_int            nnz,
                                                    aoclsparse_int           *col_ptr,
                                                    aoclsparse_int           *row_idx,
                                                    aoclsparse_float_complex *val);
DLL_PUBLIC aoclsparse_status aoclsparse_create_zcsc(aoclsparse_matrix         *mat,
                                                    aoclsparse_index_base      base,
                                                    aoclsparse_int             M,
                                                    aoclsparse_int             N,
                                                    aoclsparse_int             nnz,
                                                    aoclsparse_int            *col_ptr,
                                                    aoclsparse_int            *row_idx,
                                                    aoclsparse_double_complex *val);
DLL_PUBLIC aoclsparse_status aoclsparse_destroy(aoclsparse_matrix *mat);

/* Value refresh keeping the pattern (and the analysis).  Replaces aoclsparse_?update_values
 * (aoclsparse_auxiliary.h:311-358): len must equal nnz, val replaces all stored values. */
DLL_PUBLIC aoclsparse_status aoclsparse_supdate_values(aoclsparse_matrix A, aoclsparse_int len, float *val);
DLL_PUBLIC aoclsparse_status aoclsparse_dupdate_values(aoclsparse_matrix A, aoclsparse_int len, double *val);
DLL_PUBLIC aoclsparse_status aoclsparse_cupdate_values(aoclsparse_matrix         A,
                                                       aoclsparse_int            len,
                                                       aoclsparse_float_complex *val);
DLL_PUBLIC aoclsparse_status aoclsparse_zupdate_values(aoclsparse_matrix          A,
                                                       aoclsparse_int             len,
                                                       aoclsparse_double_complex *val);

/* Single-entry update.  Replaces aoclsparse_?set_value (aoclsparse_auxiliary.h:700-746; implementation
 * library/src/extra/aoclsparse_auxiliary.hpp:388-473): (row_idx, col_idx) are given in the matrix' index base;
 * out of range -> invalid_value, value type mismatch -> wrong_type, entry not stored -> invalid_index_value.
 * B200: updates the device copy in place (first match in storage order) and drops derived copies. */
DLL_PUBLIC aoclsparse_status aoclsparse_sset_value(aoclsparse_matrix A, aoclsparse_int row_idx, aoclsparse_int col_idx, float val);
DLL_PUBLIC aoclsparse_status aoclsparse_dset_value(aoclsparse_matrix A, aoclsparse_int row_idx, aoclsparse_int col_idx, double val);
DLL_PUBLIC aoclsparse_status aoclsparse_cset_value(aoclsparse_matrix        A,
                                                   aoclsparse_int           row_idx,
                                                   aoclsparse_int           col_idx,
                                                   aoclsparse_float_complex val);
DLL_PUBLIC aoclsparse_status aoclsparse_zset_value(aoclsparse_matrix         A,
                                                   aoclsparse_int            row_idx,
                                                   aoclsparse_int            col_idx,
                                                   aoclsparse_double_complex val);

/* ------------------------------------------------------------------------------------------
 * Hints and analysis.  Replaces aoclsparse_analysis.h:57 (optimize), :88-110 (mv / mv_kid / mm
 * hints), :279 (memory hint); implementation library/src/analysis/aoclsparse_analysis.cpp:426-747.
 *
 * B200: aoclsparse_optimize is a GPU analysis pass.  It cuts the rows into nnz-balanced row blocks
 * (<= a fixed number of non-zeros per CTA; rows longer than that are split across CTAs), bins every
 * block by its row-length profile into a thread-per-row, warp-per-row or CTA-per-row strategy, and
 * -- when a transposed / symmetric / hermitian mv or mm hint was given and the memory policy is
 * unrestricted -- materialises the transposed / expanded device copy that hint needs.
 * kid (aoclsparse_set_mv_hint_kid) forces the strategy of every row block: -1 auto, 0 thread-per-row,
 * 1 warp-per-row, 2 CTA-wide product + segmented sum; any other kid makes aoclsparse_?mv return
 * aoclsparse_status_invalid_kid, as an unavailable kernel id does in the reference.
 * ---------------------------------------------------------------------------------------- */
DLL_PUBLIC aoclsparse_status aoclsparse_optimize(aoclsparse_matrix mat);
DLL_PUBLIC aoclsparse_status aoclsparse_set_mv_hint(aoclsparse_matrix          mat,
                                                    aoclsparse_operation       trans,
                                                    const aoclsparse_mat_descr descr,
                                                    aoclsparse_int             expected_no_of_calls);
DLL_PUBLIC aoclsparse_status aoclsparse_set_mv_hint_kid(aoclsparse_matrix          mat,
                                                        aoclsparse_operation       trans,
                                                        const aoclsparse_mat_descr descr,
                                                        aoclsparse_int expected_no_of_calls,
                                                        aoclsparse_int kid);
DLL_PUBLIC aoclsparse_status aoclsparse_set_mm_hint(aoclsparse_matrix          mat,
                                                    aoclsparse_operation       trans,
                                                    const aoclsparse_mat_descr descr,
                                                    aoclsparse_int             expected_no_of_calls);
DLL_PUBLIC aoclsparse_status aoclsparse_set_memory_hint(aoclsparse_matrix             mat,
                                                        const aoclsparse_memory_usage policy);
/* Hints for operations next to the path (aoclsparse_analysis.h:102-161,202-206): validated and recorded like the
 * ones above; a dotmv hint counts as a multiply hint for aoclsparse_optimize, the others only take their place in
 * the hint list (triangular solves, smoothers and the LU / SOR preconditioners are outside this library's path). */
DLL_PUBLIC aoclsparse_status aoclsparse_set_sv_hint(aoclsparse_matrix          mat,
                                                    aoclsparse_operation       trans,
                                                    const aoclsparse_mat_descr descr,
                                                    aoclsparse_int             expected_no_of_calls);
DLL_PUBLIC aoclsparse_status aoclsparse_set_2m_hint(aoclsparse_matrix          mat,
                                                    aoclsparse_operation       trans,
                                                    const aoclsparse_mat_descr descr,
                                                    aoclsparse_int             expected_no_of_calls);
DLL_PUBLIC aoclsparse_status aoclsparse_set_lu_smoother_hint(aoclsparse_matrix          mat,
                                                             aoclsparse_operation       trans,
                                                             const aoclsparse_mat_descr descr,
                                                             aoclsparse_int             expected_no_of_calls);
DLL_PUBLIC aoclsparse_status aoclsparse_set_sm_hint(aoclsparse_matrix          mat,
                                                    aoclsparse_operation       trans,
                                                    const aoclsparse_mat_descr descr,
                                                    const aoclsparse_order     order,
                                                    const aoclsparse_int       expected_no_of_calls);
DLL_PUBLIC aoclsparse_status aoclsparse_set_dotmv_hint(aoclsparse_matrix          mat,
                                                       aoclsparse_operation       trans,
                                                       const aoclsparse_mat_descr descr,
                                                       aoclsparse_int             expected_no_of_calls);
DLL_PUBLIC aoclsparse_status aoclsparse_set_symgs_hint(aoclsparse_matrix          mat,
                                                       aoclsparse_operation       trans,
                                                       const aoclsparse_mat_descr descr,
                                                       aoclsparse_int             expected_no_of_calls);

/* ------------------------------------------------------------------------------------------
 * Sparse matrix - vector product  y = alpha * op(A) * x + beta * y.
 * Replaces aoclsparse_{s,d,c,z}mv (aoclsparse_functions.h:1279-1313; front end
 * library/src/level2/aoclsparse_mv.cpp:41-349, kernels library/src/level2/aoclsparse_csrmv_kr.hpp,
 * aoclsparse_csrmv_avx512.cpp, aoclsparse_csrmv_kt.cpp).
 *
 * Error precedence (mv.cpp:55-121): NULL alpha/beta, A, descr, x/y -> invalid_pointer; descr base
 * != matrix base -> invalid_value; bad op -> invalid_value; value type mismatch -> wrong_type; bad
 * descr type -> invalid_value; symmetric/hermitian on non-square -> invalid_size; real type with a
 * hermitian descr -> not_implemented; empty matrix -> y = beta*y, success.  beta == 0 overwrites y
 * without reading it (NaN/Inf in y are ignored).
 *
 * B200: x and y may be device/managed pointers (used in place; the call returns after the kernel is
 * enqueued on the calling thread's stream, see aoclsparse_b200_set_stream) or plain host pointers
 * (staged through device scratch; the call returns after y has been copied back).
 * ---------------------------------------------------------------------------------------- */
DLL_PUBLIC aoclsparse_status aoclsparse_smv(aoclsparse_operation       op,
                                            const float               *alpha,
                                            aoclsparse_matrix          A,
                                            const aoclsparse_mat_descr descr,
                                            const float               *x,
                                            const float               *beta,
                                            float                     *y);
DLL_PUBLIC aoclsparse_status aoclsparse_dmv(aoclsparse_operation       op,
                                            const double              *alpha,
                                            aoclsparse_matrix          A,
                                            const aoclsparse_mat_descr descr,
                                            const double              *x,
                                            const double              *beta,
                                            double                    *y);
DLL_PUBLIC aoclsparse_status aoclsparse_cmv(aoclsparse_operation            op,
                                            const aoclsparse_float_complex *alpha,
                                            aoclsparse_matrix               A,
                                            const aoclsparse_mat_descr      descr,
                                            const aoclsparse_float_complex *x,
                                            const aoclsparse_float_complex *beta,
                                            aoclsparse_float_complex       *y);
DLL_PUBLIC aoclsparse_status aoclsparse_zmv(aoclsparse_operation             op,
                                            const aoclsparse_double_complex *alpha,
                                            aoclsparse_matrix                A,
                                            const aoclsparse_mat_descr       descr,
                                            const aoclsparse_double_complex *x,
                                            const aoclsparse_double_complex *beta,
                                            aoclsparse_double_complex       *y);

/* Fused product and dot:  y = alpha * op(A) * x + beta * y,  *d = sum_{i < min(m,n)} conj(x_i) * y_i.
 * Replaces aoclsparse_{s,d,c,z}dotmv (aoclsparse_functions.h:1761-1800; library/src/level2/aoclsparse_dotmv.hpp:30-62:
 * NULL d or A -> invalid_pointer, then every check of aoclsparse_?mv).  alpha / beta by value.  d may be a host or a
 * device pointer. */
DLL_PUBLIC aoclsparse_status aoclsparse_sdotmv(const aoclsparse_operation op,
                                               const float                alpha,
                                               aoclsparse_matrix          A,
                                               const aoclsparse_mat_descr descr,
                                               const float               *x,
                                               const float                beta,
                                               float                     *y,
                                               float                     *d);
DLL_PUBLIC aoclsparse_status aoclsparse_ddotmv(const aoclsparse_operation op,
                                               const double               alpha,
                                               aoclsparse_matrix          A,
                                               const aoclsparse_mat_descr descr,
                                               const double              *x,
                                               const double               beta,
                                               double                    *y,
                                               double                    *d);
DLL_PUBLIC aoclsparse_status aoclsparse_cdotmv(const aoclsparse_operation      op,
                                               const aoclsparse_float_complex  alpha,
                                               aoclsparse_matrix               A,
                                               const aoclsparse_mat_descr      descr,
                                               const aoclsparse_float_complex *x,
                                               const aoclsparse_float_complex  beta,
                                               aoclsparse_float_complex       *y,
                                               aoclsparse_float_complex       *d);
DLL_PUBLIC aoclsparse_status aoclsparse_zdotmv(const aoclsparse_operation       op,
                                               const aoclsparse_double_complex  alpha,
                                               aoclsparse_matrix                A,
                                               const aoclsparse_mat_descr       descr,
                                               const aoclsparse_double_complex *x,
                                               const aoclsparse_double_complex  beta,
                                               aoclsparse_double_complex       *y,
                                               aoclsparse_double_complex       *d);

/* Handle-free legacy entry.  Replaces aoclsparse_{s,d}csrmv (aoclsparse_functions.h:695-721;
 * library/src/level2/aoclsparse_csrmv.cpp:30-63, checks in aoclsparse_csrmv.hpp:63-110): general
 * and symmetric descriptors only (others: not_implemented); a symmetric descriptor means "lower
 * triangle with its diagonal stored", fill_mode / diag_type are ignored (aoclsparse_csrmv_symm,
 * library/src/level2/aoclsparse_csrmv_kr.hpp:41-91).
 * B200: the CSR arrays are uploaded and analysed on every call -- correct, and as slow as that sounds;
 * iterated use belongs on aoclsparse_create_?csr + aoclsparse_?mv. */
DLL_PUBLIC aoclsparse_status aoclsparse_scsrmv(aoclsparse_operation       trans,
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
                                               float                     *y);
DLL_PUBLIC aoclsparse_status aoclsparse_dcsrmv(aoclsparse_operation       trans,
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
                                               double                    *y);

/* ------------------------------------------------------------------------------------------
 * Sparse (CSR) x dense product  C = alpha * op(A) * B + beta * C.
 * Replaces aoclsparse_{s,d,c,z}csrmm (aoclsparse_functions.h:2460-2510) and the _kid variants
 * (:3397-3452); front end library/src/level3/aoclsparse_csrmm.hpp:429-838, kernels
 * aoclsparse_csrmm_kt.cpp and aoclsparse_csrmm.hpp:36-427.
 *
 * op(A) is m x k; B is k x n and C is m x n, stored row-major (ld = row stride) or column-major
 * (ld = column stride) as selected by order.  Error precedence (csrmm.hpp:447-618): NULL A/B/C/descr
 * -> invalid_pointer; bad op -> invalid_value; triangular descr -> not_implemented; symmetric /
 * hermitian on non-square -> invalid_size; bad order -> invalid_value; value type mismatch ->
 * wrong_type; base mismatch -> invalid_value; n<0 -> invalid_size; any of m,n,k == 0 or (alpha==0 and
 * beta==1) -> success, C untouched; ldb / ldc too small or dim*ld overflowing aoclsparse_int ->
 * invalid_size.  Padding elements of C (between n and ldc) are never written.
 * ---------------------------------------------------------------------------------------- */
DLL_PUBLIC aoclsparse_status aoclsparse_scsrmm(aoclsparse_operation       op,
                                               const float                alpha,
                                               const aoclsparse_matrix    A,
                                               const aoclsparse_mat_descr descr,
                                               aoclsparse_order           order,
                                               const float               *B,
                                               aoclsparse_int             n,
                                               aoclsparse_int             ldb,
                                               const float                beta,
                                               float                     *C,
                                               aoclsparse_int             ldc);
DLL_PUBLIC aoclsparse_status aoclsparse_dcsrmm(aoclsparse_operation       op,
                                               const double               alpha,
                                               const aoclsparse_matrix    A,
                                               const aoclsparse_mat_descr descr,
                                               aoclsparse_order           order,
                                               const double              *B,
                                               aoclsparse_int             n,
                                               aoclsparse_int             ldb,
                                               const double               beta,
                                               double                    *C,
                                               aoclsparse_int             ldc);
DLL_PUBLIC aoclsparse_status aoclsparse_ccsrmm(aoclsparse_operation            op,
                                               const aoclsparse_float_complex  alpha,
                                               const aoclsparse_matrix         A,
                                               const aoclsparse_mat_descr      descr,
                                               aoclsparse_order                order,
                                               const aoclsparse_float_complex *B,
                                               aoclsparse_int                  n,
                                               aoclsparse_int                  ldb,
                                               const aoclsparse_float_complex  beta,
                                               aoclsparse_float_complex       *C,
                                               aoclsparse_int                  ldc);
DLL_PUBLIC aoclsparse_status aoclsparse_zcsrmm(aoclsparse_operation             op,
                                               const aoclsparse_double_complex  alpha,
                                               const aoclsparse_matrix          A,
                                               const aoclsparse_mat_descr       descr,
                                               aoclsparse_order                 order,
                                               const aoclsparse_double_complex *B,
                                               aoclsparse_int                   n,
                                               aoclsparse_int                   ldb,
                                               const aoclsparse_double_complex  beta,
                                               aoclsparse_double_complex       *C,
                                               aoclsparse_int                   ldc);
DLL_PUBLIC aoclsparse_status aoclsparse_scsrmm_kid(aoclsparse_operation       op,
                                                   const float                alpha,
                                                   const aoclsparse_matrix    A,
                                                   const aoclsparse_mat_descr descr,
                                                   aoclsparse_order           order,
                                                   const float               *B,
                                                   aoclsparse_int             n,
                                                   aoclsparse_int             ldb,
                                                   const float                beta,
                                                   float                     *C,
                                                   aoclsparse_int             ldc,
                                                   const aoclsparse_int       kid);
DLL_PUBLIC aoclsparse_status aoclsparse_dcsrmm_kid(aoclsparse_operation       op,
                                                   const double               alpha,
                                                   const aoclsparse_matrix    A,
                                                   const aoclsparse_mat_descr descr,
                                                   aoclsparse_order           order,
                                                   const double              *B,
                                                   aoclsparse_int             n,
                                                   aoclsparse_int             ldb,
                                                   const double               beta,
                                                   double                    *C,
                                                   aoclsparse_int             ldc,
                                                   const aoclsparse_int       kid);
DLL_PUBLIC aoclsparse_status aoclsparse_ccsrmm_kid(aoclsparse_operation            op,
                                                   const aoclsparse_float_complex  alpha,
                                                   const aoclsparse_matrix         A,
                                                   const aoclsparse_mat_descr      descr,
                                                   aoclsparse_order                order,
                                                   const aoclsparse_float_complex *B,
                                                   aoclsparse_int                  n,
                                                   aoclsparse_int                  ldb,
                                                   const aoclsparse_float_complex  beta,
                                                   aoclsparse_float_complex       *C,
                                                   aoclsparse_int                  ldc,
                                                   const aoclsparse_int            kid);
DLL_PUBLIC aoclsparse_status aoclsparse_zcsrmm_kid(aoclsparse_operation             op,
                                                   const aoclsparse_double_complex  alpha,
                                                   const aoclsparse_matrix          A,
                                                   const aoclsparse_mat_descr       descr,
                                                   aoclsparse_order                 order,
                                                   const aoclsparse_double_complex *B,
                                                   aoclsparse_int                   n,
                                                   aoclsparse_int                   ldb,
                                                   const aoclsparse_double_complex  beta,
                                                   aoclsparse_double_complex       *C,
                                                   aoclsparse_int                   ldc,
                                                   const aoclsparse_int             kid);

/* ------------------------------------------------------------------------------------------
 * Sparse x sparse -> NEW sparse CSR matrix (SpGEMM), SURVEY.md section 8(f) row 4.
 *
 * aoclsparse_sp2m  C = op(A) op(B)   (aoclsparse_functions.h:2103-2210; aoclsparse::sp2m<T>,
 *                                     library/src/level3/aoclsparse_csr2m.cpp:592-860)
 * aoclsparse_spmm  C = op(A) B       (aoclsparse_functions.h:2212-2261;
 *                                     library/src/level3/aoclsparse_spmm.cpp:27-67)
 * A and B are CSR or CSC handles of the same value type, general descriptors only.  request selects
 * the single-stage product (aoclsparse_stage_full_computation) or the two-stage one: nnz_count
 * creates *C with its row pointers (column / value arrays allocated, not filled), finalize fills
 * them and may be repeated after the VALUES of A / B changed.  *C is library-owned (free it with
 * aoclsparse_destroy), always zero-based.  Validation order and codes follow the reference: NULL
 * A/B/C -> invalid_pointer; differing value types -> wrong_type; NULL descriptor ->
 * invalid_pointer; descriptor base != matrix base -> invalid_value; non-general descriptor ->
 * not_implemented; bad op -> invalid_value; inner dimensions differ -> invalid_size; an empty
 * operand gives an empty C (success); more than 2^31-1 entries in C -> invalid_size.
 *
 * B200: the product is computed on the device with per-row hash tables (csrc/spgemm.cu); the arrays
 * of *C live in device memory and the handle can be used directly with aoclsparse_?mv /
 * aoclsparse_?csrmm / aoclsparse_sp2m.  Unlike the reference (first-touch order) the column indices
 * of every row of C are ascending.  Value sums are accumulated with atomic adds: their rounding may
 * differ in the last bits between runs.
 * ---------------------------------------------------------------------------------------- */
DLL_PUBLIC aoclsparse_status aoclsparse_sp2m(aoclsparse_operation       opA,
                                             const aoclsparse_mat_descr descrA,
                                             const aoclsparse_matrix    A,
                                             aoclsparse_operation       opB,
                                             const aoclsparse_mat_descr descrB,
                                             const aoclsparse_matrix    B,
                                             const aoclsparse_request   request,
                                             aoclsparse_matrix         *C);
DLL_PUBLIC aoclsparse_status aoclsparse_spmm(aoclsparse_operation    opA,
                                             const aoclsparse_matrix A,
                                             const aoclsparse_matrix B,
                                             aoclsparse_matrix      *C);

/* ------------------------------------------------------------------------------------------
 * Conjugate gradients -- the iterated-SpMV consumer of the path.
 * Replaces the CG part of the reference's iterative-solver suite (aoclsparse_solvers.h:114-570;
 * library/src/solvers/aoclsparse_itsol_functions.{hpp,cpp}):
 *   aoclsparse_itsol_{s,d}_init / aoclsparse_itsol_destroy             problem handle
 *   aoclsparse_itsol_option_set                                          every option the reference registers
 *       (aoclsparse_itsol_list_options.hpp:63-239: "iterative method", "cg iteration limit" [500],
 *       "cg rel tolerance" [2 s sqrt(2 eps)], "cg abs tolerance" [s sqrt(2 eps)], "cg preconditioner",
 *       and the gmres ones); names and string values are trimmed, blank-squeezed and case-folded
 *       like the reference; unknown option / bad or out-of-range value -> invalid_value
 *   aoclsparse_itsol_{s,d}_solve       forward interface: A symmetric, descriptor lower (else
 *       invalid_value), optional user preconditioner and monitor callbacks
 *   aoclsparse_itsol_{s,d}_rci_input / _rci_solve   reverse communication (jobs as in the reference)
 * rinfo[0] = |A x - b|, rinfo[1] = |b|, rinfo[30] = iterations (itsol_functions.hpp:36-38).
 * Exit codes: success, aoclsparse_status_maxit, aoclsparse_status_user_stop,
 * aoclsparse_status_numerical_error (A not positive definite / breakdown), as in the reference.
 * Not provided (aoclsparse_status_not_implemented): GMRES, the symmetric Gauss-Seidel
 * preconditioner, complex handles.
 *
 * B200: all vector work runs on the device (csrc/itsol.cu).  b and x may be host or device arrays.
 * The work vectors handed out through *u / *v (and passed to the callbacks) are CUDA managed memory:
 * valid as host pointers AND as device pointers for aoclsparse_?mv.
 * ---------------------------------------------------------------------------------------- */
typedef struct _aoclsparse_itsol_handle *aoclsparse_itsol_handle;

/* aoclsparse_solvers.h:114-134 */
typedef enum aoclsparse_itsol_rci_job_
{
    aoclsparse_rci_interrupt          = -1,
    aoclsparse_rci_stop               = 0,
    aoclsparse_rci_start              = 1,
    aoclsparse_rci_mv                 = 2,
    aoclsparse_rci_precond            = 3,
    aoclsparse_rci_stopping_criterion = 4
} aoclsparse_itsol_rci_job;

DLL_PUBLIC aoclsparse_status aoclsparse_itsol_s_init(aoclsparse_itsol_handle *handle);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_d_init(aoclsparse_itsol_handle *handle);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_c_init(aoclsparse_itsol_handle *handle);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_z_init(aoclsparse_itsol_handle *handle);
DLL_PUBLIC void              aoclsparse_itsol_destroy(aoclsparse_itsol_handle *handle);
/* aoclsparse_solvers.h:147 -- prints the handle's options and their values to the standard output */
DLL_PUBLIC void              aoclsparse_itsol_handle_prn_options(aoclsparse_itsol_handle handle);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_option_set(aoclsparse_itsol_handle handle,
                                                         const char             *option,
                                                         const char             *value);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_d_rci_input(aoclsparse_itsol_handle handle,
                                                          aoclsparse_int          n,
                                                          const double           *b);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_s_rci_input(aoclsparse_itsol_handle handle,
                                                          aoclsparse_int          n,
                                                          const float            *b);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_d_rci_solve(aoclsparse_itsol_handle   handle,
                                                          aoclsparse_itsol_rci_job *ircomm,
                                                          double                  **u,
                                                          double                  **v,
                                                          double                   *x,
                                                          double                    rinfo[100]);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_s_rci_solve(aoclsparse_itsol_handle   handle,
                                                          aoclsparse_itsol_rci_job *ircomm,
                                                          float                   **u,
                                                          float                   **v,
                                                          float                    *x,
                                                          float                     rinfo[100]);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_d_solve(
    aoclsparse_itsol_handle    handle,
    aoclsparse_int             n,
    aoclsparse_matrix          mat,
    const aoclsparse_mat_descr descr,
    const double              *b,
    double                    *x,
    double                     rinfo[100],
    aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const double *u, double *v, void *udata),
    aoclsparse_int monit(aoclsparse_int n, const double *x, const double *r, double rinfo[100], void *udata),
    void *udata);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_s_solve(
    aoclsparse_itsol_handle    handle,
    aoclsparse_int             n,
    aoclsparse_matrix          mat,
    const aoclsparse_mat_descr descr,
    const float               *b,
    float                     *x,
    float                      rinfo[100],
    aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const float *u, float *v, void *udata),
    aoclsparse_int monit(aoclsparse_int n, const float *x, const float *r, float rinfo[100], void *udata),
    void *udata);
/* complex handles, aoclsparse_solvers.h:276-283, 395-408, 537-575 (conjugate gradients: the reference's recurrence with
 * unconjugated dot products, for complex SYMMETRIC matrices; rinfo and the tolerances are real) */
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_c_rci_input(aoclsparse_itsol_handle         handle,
                                                          aoclsparse_int                  n,
                                                          const aoclsparse_float_complex *b);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_z_rci_input(aoclsparse_itsol_handle          handle,
                                                          aoclsparse_int                   n,
                                                          const aoclsparse_double_complex *b);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_c_rci_solve(aoclsparse_itsol_handle    handle,
                                                          aoclsparse_itsol_rci_job  *ircomm,
                                                          aoclsparse_float_complex **u,
                                                          aoclsparse_float_complex **v,
                                                          aoclsparse_float_complex  *x,
                                                          float                      rinfo[100]);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_z_rci_solve(aoclsparse_itsol_handle     handle,
                                                          aoclsparse_itsol_rci_job   *ircomm,
                                                          aoclsparse_double_complex **u,
                                                          aoclsparse_double_complex **v,
                                                          aoclsparse_double_complex  *x,
                                                          double                      rinfo[100]);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_c_solve(
    aoclsparse_itsol_handle         handle,
    aoclsparse_int                  n,
    aoclsparse_matrix               mat,
    const aoclsparse_mat_descr      descr,
    const aoclsparse_float_complex *b,
    aoclsparse_float_complex       *x,
    float                           rinfo[100],
    aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const aoclsparse_float_complex *u, aoclsparse_float_complex *v, void *udata),
    aoclsparse_int monit(aoclsparse_int n, const aoclsparse_float_complex *x, const aoclsparse_float_complex *r, float rinfo[100], void *udata),
    void *udata);
DLL_PUBLIC aoclsparse_status aoclsparse_itsol_z_solve(
    aoclsparse_itsol_handle          handle,
    aoclsparse_int                   n,
    aoclsparse_matrix                mat,
    const aoclsparse_mat_descr       descr,
    const aoclsparse_double_complex *b,
    aoclsparse_double_complex       *x,
    double                           rinfo[100],
    aoclsparse_int precond(aoclsparse_int flag, aoclsparse_int n, const aoclsparse_double_complex *u, aoclsparse_double_complex *v, void *udata),
    aoclsparse_int monit(aoclsparse_int n, const aoclsparse_double_complex *x, const aoclsparse_double_complex *r, double rinfo[100], void *udata),
    void *udata);

/* Read access to the CSR arrays of a handle.  Replaces aoclsparse_export_?csr
 * (aoclsparse_auxiliary.h:744-820; aoclsparse_export_csr_t, library/src/extra/
 * aoclsparse_auxiliary.cpp:1303-1350): NULL argument -> invalid_pointer, value type mismatch ->
 * wrong_type, a handle without CSR arrays (created from CSC) -> invalid_value.  The arrays are in
 * the handle's own index base, owned by the handle and valid until the next export call on it or
 * aoclsparse_destroy.  B200: the handle's arrays live on the device, so every call copies them into
 * a host mirror; aoclsparse_b200_export_device_csr (aoclsparse_b200.h) returns the device arrays. */
DLL_PUBLIC aoclsparse_status aoclsparse_export_scsr(const aoclsparse_matrix mat,
                                                    aoclsparse_index_base  *base,
                                                    aoclsparse_int         *m,
                                                    aoclsparse_int         *n,
                                                    aoclsparse_int         *nnz,
                                                    aoclsparse_int        **row_ptr,
                                                    aoclsparse_int        **col_ind,
                                                    float                 **val);
DLL_PUBLIC aoclsparse_status aoclsparse_export_dcsr(const aoclsparse_matrix mat,
                                                    aoclsparse_index_base  *base,
                                                    aoclsparse_int         *m,
                                                    aoclsparse_int         *n,
                                                    aoclsparse_int         *nnz,
                                                    aoclsparse_int        **row_ptr,
                                                    aoclsparse_int        **col_ind,
                                                    double                **val);
DLL_PUBLIC aoclsparse_status aoclsparse_export_ccsr(const aoclsparse_matrix    mat,
                                                    aoclsparse_index_base     *base,
                                                    aoclsparse_int            *m,
                                                    aoclsparse_int            *n,
                                                    aoclsparse_int            *nnz,
                                                    aoclsparse_int           **row_ptr,
                                                    aoclsparse_int           **col_ind,
                                                    aoclsparse_float_complex **val);
DLL_PUBLIC aoclsparse_status aoclsparse_export_zcsr(const aoclsparse_matrix     mat,
                                                    aoclsparse_index_base      *base,
                                                    aoclsparse_int             *m,
                                                    aoclsparse_int             *n,
                                                    aoclsparse_int             *nnz,
                                                    aoclsparse_int            **row_ptr,
                                                    aoclsparse_int            **col_ind,
                                                    aoclsparse_double_complex **val);

/* Array-level CSR -> CSC conversion.  Replaces aoclsparse_?csr2csc (aoclsparse_convert.h:430-530;
 * aoclsparse_csr2csc_template, library/src/conversion/aoclsparse_convert.hpp:553-660): only the index
 * base of descr is used; NULL descr -> invalid_pointer, negative sizes -> invalid_size, an empty matrix
 * fills csc_col_ptr with baseCSC, bad bases -> invalid_value, NULL arrays -> invalid_pointer.  Row
 * indices ascend inside every column; repeated entries keep their order.  B200: the conversion runs on
 * the device (csrc/transpose.cu); every array may be host or device memory. */
DLL_PUBLIC aoclsparse_status aoclsparse_scsr2csc(aoclsparse_int             m,
                                                 aoclsparse_int             n,
                                                 aoclsparse_int             nnz,
                                                 const aoclsparse_mat_descr descr,
                                                 aoclsparse_index_base      baseCSC,
                                                 const aoclsparse_int      *csr_row_ptr,
                                                 const aoclsparse_int      *csr_col_ind,
                                                 const float *csr_val,
                                                 aoclsparse_int            *csc_row_ind,
                                                 aoclsparse_int            *csc_col_ptr,
                                                 float *csc_val);
DLL_PUBLIC aoclsparse_status aoclsparse_dcsr2csc(aoclsparse_int             m,
                                                 aoclsparse_int             n,
                                                 aoclsparse_int             nnz,
                                                 const aoclsparse_mat_descr descr,
                                                 aoclsparse_index_base      baseCSC,
                                                 const aoclsparse_int      *csr_row_ptr,
                                                 const aoclsparse_int      *csr_col_ind,
                                                 const double *csr_val,
                                                 aoclsparse_int            *csc_row_ind,
                                                 aoclsparse_int            *csc_col_ptr,
                                                 double *csc_val);
DLL_PUBLIC aoclsparse_status aoclsparse_ccsr2csc(aoclsparse_int             m,
                                                 aoclsparse_int             n,
                                                 aoclsparse_int             nnz,
                                                 const aoclsparse_mat_descr descr,
                                                 aoclsparse_index_base      baseCSC,
                                                 const aoclsparse_int      *csr_row_ptr,
                                                 const aoclsparse_int      *csr_col_ind,
                                                 const aoclsparse_float_complex *csr_val,
                                                 aoclsparse_int            *csc_row_ind,
                                                 aoclsparse_int            *csc_col_ptr,
                                                 aoclsparse_float_complex *csc_val);
DLL_PUBLIC aoclsparse_status aoclsparse_zcsr2csc(aoclsparse_int             m,
                                                 aoclsparse_int             n,
                                                 aoclsparse_int             nnz,
                                                 const aoclsparse_mat_descr descr,
                                                 aoclsparse_index_base      baseCSC,
                                                 const aoclsparse_int      *csr_row_ptr,
                                                 const aoclsparse_int      *csr_col_ind,
                                                 const aoclsparse_double_complex *csr_val,
                                                 aoclsparse_int            *csc_row_ind,
                                                 aoclsparse_int            *csc_col_ptr,
                                                 aoclsparse_double_complex *csc_val);

/* Ascending column indices inside every row (row indices inside every column for a CSC handle), values
 * moved along.  Replaces aoclsparse_order_mat (aoclsparse_auxiliary.h:1015-1035; library/src/extra/
 * aoclsparse_auxiliary.cpp:840-878).  B200: sorts the device copy; the caller's arrays are not touched. */
DLL_PUBLIC aoclsparse_status aoclsparse_order_mat(aoclsparse_matrix mat);

#ifdef __cplusplus
}
#endif

#include "aoclsparse_b200.h"

#endif /* AOCLSPARSE_H_B200_ */
