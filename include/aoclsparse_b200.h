/* aoclsparse_b200.h -- extensions that have no counterpart in the reference API.
 *
 * The reference (AOCL-Sparse v5.3.2) is a single-process CPU library: it has no notion of a CUDA
 * stream, of where x / y live, of a multi-GPU row partition, and it only exposes its analysis
 * results to white-box tests that reach into the handle.  Everything a caller needs for those is
 * collected here under the aoclsparse_b200_ prefix.  Plain C ABI, plain pointers and sizes.
 */
#ifndef AOCLSPARSE_B200_H_
#define AOCLSPARSE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- execution control ------------------------------------------------------------------ */

/* CUDA stream (a cudaStream_t / CUstream passed as void*) on which the CALLING HOST THREAD's
 * subsequent aoclsparse_* calls enqueue their work.  NULL selects the legacy default stream. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_set_stream(void *cuda_stream);
DLL_PUBLIC void             *aoclsparse_b200_get_stream(void);

/* Last CUDA error text seen by the calling thread ("" if none); for diagnosing internal_error. */
DLL_PUBLIC const char *aoclsparse_b200_last_error(void);

/* Number of kernels this library has launched so far in this process (all threads). */
DLL_PUBLIC unsigned long long aoclsparse_b200_launch_count(void);

/* ---- analysis read-back (black-box replacement for the reference's white-box test hooks) ---- */

/* Facts the reference stores in _aoclsparse_matrix / aoclsparse::csr and that its unit tests read
 * directly (createcsr_tests.cpp:77-99, optimize_tests.cpp:42-58), plus the GPU plan summary. */
typedef struct aoclsparse_b200_matrix_info_
{
    aoclsparse_int m, n, nnz;
    int            base;         /* aoclsparse_index_base of the user's arrays                  */
    int            val_type;     /* aoclsparse_matrix_data_type                                 */
    int            sort;         /* aoclsparse_matrix_sort, as aoclsparse_mat_check_internal    */
    int            fulldiag;     /* 1 if every row i < min(m,n) stores its diagonal             */
    aoclsparse_int min_col;      /* smallest / largest 0-based column index present, or n / -1  */
    aoclsparse_int max_col;
    aoclsparse_int max_row_nnz;  /* longest row                                                 */
    int            optimized;    /* 1 once a plan exists (aoclsparse_optimize or first use)     */
    int            n_hints;      /* hints recorded so far                                       */
    int            n_copies;     /* device CSR copies held (1 = the input only)                 */
    aoclsparse_int block_nnz;    /* plan: nnz capacity of one row block (CTA)                   */
    aoclsparse_int block_rows;   /* plan: row capacity of one row block                         */
    aoclsparse_int n_blocks;     /* plan: number of row blocks (= CTAs of the main kernel)      */
    aoclsparse_int n_thread_blocks; /* blocks binned thread-per-row                             */
    aoclsparse_int n_warp_blocks;   /* blocks binned warp-per-row                               */
    aoclsparse_int n_product_blocks; /* blocks binned CTA-wide product + segmented sum          */
    aoclsparse_int n_long_segments; /* blocks that are one segment of a row split across CTAs   */
    aoclsparse_int n_long_rows;     /* rows split across CTAs                                   */
    aoclsparse_int n_diag_codes;    /* diagonal-code copy of col_idx (one byte per entry indexing a table of the
                                       distinct col - row offsets): table entries, 0 = not built / not applicable */
    aoclsparse_int n_entry_codes;   /* entry-code copy (one byte per entry indexing a table of the distinct (col - row, value)
                                       pairs), only next to the diagonal-code copy: table entries, 0 = not built         */
    aoclsparse_int e_block_nnz;     /* block plan of the entry-coded kernels (0 when there is no entry-code copy): entry   */
    aoclsparse_int e_block_rows;    /* capacity, row capacity and number of blocks; every block is thread-per-row         */
    aoclsparse_int e_n_blocks;
} aoclsparse_b200_matrix_info;

/* Diagonal-code copy of the stored column indices, built by aoclsparse_optimize for banded / stencil matrices (every
 * row block thread-per-row, at most 256 distinct col - row offsets): offsets[0..*n_codes) ascending, codes[p] = index of
 * (col_idx[p] - row of p) in that table.  *n_codes = 0 when it was not built.  offsets / codes may be NULL.  Integer
 * metadata with no counterpart in the reference: pinned bit for bit by oracle/csr_oracle.c::oracle_diag_codes. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_diag_codes(const aoclsparse_matrix A,
                                                            aoclsparse_int         *n_codes,
                                                            aoclsparse_int         *offsets,
                                                            unsigned char          *codes);

/* Entry-code copy of the stored entries, built by aoclsparse_optimize next to the diagonal-code copy when the matrix holds
 * at most 256 distinct (col - row, value) pairs (values compared as bit patterns; 4- and 8-byte value types) -- a
 * constant-coefficient stencil holds as many as it has points: pair i is (offsets[i], values[i]), ascending by offset, then
 * by value pattern; ecodes[p] = index of entry p's pair.  The multiply then streams ONE byte per stored entry and decodes
 * the identical column and value.  *n_pairs = 0 when it was not built.  offsets / values (elements of the handle's value
 * type) / ecodes may be NULL.  aoclsparse_?update_values / aoclsparse_?set_value make the copy stale; it is encoded again
 * before the next multiply.  Integer metadata with no counterpart in the reference: pinned bit for bit by
 * oracle/csr_oracle.c::oracle_entry_codes. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_entry_codes(const aoclsparse_matrix A,
                                                             aoclsparse_int         *n_pairs,
                                                             aoclsparse_int         *offsets,
                                                             void                   *values,
                                                             unsigned char          *ecodes);
/* Row blocks the entry-coded kernels run on (same layout as aoclsparse_b200_get_plan; *n_blocks = 0 when there is no
 * entry-code copy).  One staged byte per entry allows much larger blocks than the handle's main plan, which the other
 * kernels keep using. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_entry_plan(const aoclsparse_matrix A,
                                                            aoclsparse_int          capacity,
                                                            aoclsparse_int         *block_desc,
                                                            aoclsparse_int         *block_kind,
                                                            aoclsparse_int         *n_blocks);

/* Box-tile copy used by row-major aoclsparse_?csrmm on grid (stencil) matrices (csrc/mesh_tiles.cu): built on the first
 * multiply that can use it.  The matrix's distinct col - row offsets are read as a lattice {a + b*stride[1] + c*stride[2]},
 * rows are grouped into boxes box[0] x box[1] x box[2] of that grid (rows_per_tile rows).  A tile stores
 *   - its distinct columns as RUNS of consecutive columns (the kernel stages those B rows in shared memory; an entry's
 *     SLOT is the position of its column among them),
 *   - per ROW GROUP (rows_per_group consecutive rows of the tile) a WALK: the ascending columns at least one row of the
 *     group stores, each as slot | rowmask << 16 | (position of its first value) << 20 and closed by a zero entry, and a
 *     VALUE STREAM: walk entry by walk entry, row by row, the values.
 *     Both are planes over the tile's groups: walk[j][g], val[i][g].
 * state: 0 not analysed yet, 1 analysed and not usable (no lattice, unsorted rows, too little re-use), 2 ready.
 * Integer metadata with no counterpart in the reference: pinned bit for bit by tests/mesh_tiles_ref.py. */
typedef struct aoclsparse_b200_mm_tiles_info_
{
    int       state;
    int       box[3];
    long long stride[3];
    int       dims[3];
    int       rows_per_tile, rows_per_group, n_tiles, max_distinct, max_walk, max_vals, max_runs;
    long long walk_entries; /* slots of all walk planes, padding included               */
    long long val_entries;  /* slots of all value planes, padding included              */
    long long n_runs_total; /* runs of all tiles plus one terminator per tile            */
    long long row_bytes;    /* n * sizeof(T) of the multiply the boxes were sized for    */
    double    reuse, fill;  /* stored entries per staged B row; per value-plane slot     */
} aoclsparse_b200_mm_tiles_info;
DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_mm_tiles_info(const aoclsparse_matrix A, aoclsparse_b200_mm_tiles_info *info);
/* host copies of the tile arrays (any pointer may be NULL): desc 4 ints per tile {distinct, runs, U | V << 16, first
 * run}; off 2 per tile {first walk slot, first value slot}; walk / val per slot; rows per (tile, row in tile); runs 2 ints
 * per run {first column, first slot}, each tile's list closed by {-1, distinct} */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_mm_tiles(const aoclsparse_matrix A,
                                                          int                   *desc,
                                                          long long             *off,
                                                          unsigned              *walk,
                                                          void                  *val,
                                                          int                   *rows,
                                                          int                   *runs);

/* value type of a handle (aoclsparse_matrix_data_type), -1 for NULL */
DLL_PUBLIC int aoclsparse_b200_value_type(const aoclsparse_matrix A);

/* Device view (always zero-based) of a handle's stored arrays -- e.g. the result of aoclsparse_sp2m --
 * for consumers that stay on the GPU.  For a handle created from CSC arrays these are the arrays of
 * the transpose (n rows).  Valid until the handle is modified or destroyed. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_export_device_csr(const aoclsparse_matrix mat,
                                                               aoclsparse_int         *m,
                                                               aoclsparse_int         *n,
                                                               aoclsparse_int         *nnz,
                                                               const aoclsparse_int  **row_ptr,
                                                               const aoclsparse_int  **col_ind,
                                                               const void            **val);

DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_matrix_info(const aoclsparse_matrix      A,
                                                             aoclsparse_b200_matrix_info *info);

/* Copies the row-block plan of the input CSR copy to host arrays (each may be NULL to skip):
 *   block_desc[4*b + {0,1,2,3}] = first row, end row, first nnz, end nnz of block b (0-based)
 *   block_kind[b]               = strategy (0 thread, 1 warp, 2 product, 3 long segment)
 *                                 | (slot << 4), slot = partial-sum slot of a long segment
 * capacity = number of blocks the arrays can hold; *n_blocks receives the true count. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_plan(const aoclsparse_matrix A,
                                                      aoclsparse_int          capacity,
                                                      aoclsparse_int         *block_desc,
                                                      aoclsparse_int         *block_kind,
                                                      aoclsparse_int         *n_blocks);

/* The reference's "clean CSR" (aoclsparse_csr_csc_optimize, library/src/analysis/aoclsparse_csr_util.hpp:765-967;
 * golden tables tests/unit_tests/hint_tests.cpp:72-140): rows grouped lower | diagonal | upper, an explicit zero
 * inserted for every missing diagonal of rows i < n, idiag[i] / iurow[i] = position of the diagonal / of the first
 * strictly-upper entry of row i.  Built on the device on first request.  *is_internal = 0 means the input already
 * was clean (arrays are the caller's own, in the caller's index base, positions based likewise), 1 means a sorted /
 * filled base-0 copy.  Host output arrays, each may be NULL; call once with NULLs to learn *nnz.  val has the
 * matrix' value type.  The multiply kernels do not use this structure (they filter by comparing col with row). */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_get_clean_csr(aoclsparse_matrix A,
                                                           aoclsparse_int   *nnz,
                                                           int              *is_internal,
                                                           aoclsparse_int   *row_ptr,
                                                           aoclsparse_int   *col_idx,
                                                           void             *val,
                                                           aoclsparse_int   *idiag,
                                                           aoclsparse_int   *iurow);

/* Dispatch id the reference derives from (descriptor, operation, value type):
 * aoclsparse::get_doid<T> (library/src/include/aoclsparse_mtx_dispatcher.hpp:79-143).  Returns the
 * same integer (0..19) or 20 for an invalid combination. */
DLL_PUBLIC int aoclsparse_b200_doid(const aoclsparse_mat_descr descr, aoclsparse_operation op, int val_type);

/* Dispatch id a kernel must run with when the stored copy has id mat_doid and the request is
 * req_doid: aoclsparse::get_effective_doid (aoclsparse_mtx_dispatcher.hpp:311-353; table pinned by
 * tests/unit_tests/doid_score_tests.cpp:236-286).  20 = incompatible. */
DLL_PUBLIC int aoclsparse_b200_effective_doid(int mat_doid, int req_doid);

/* ---- row-sharded multi-GPU use (one process per GPU; the caller owns the exchange of x) ----- */

/* Declares that the x passed to the following aoclsparse_?mv calls (op = none) on this handle is
 * NOT the whole vector but the window of global columns [col_lo, col_hi): x[0] is column col_lo.
 * The matrix keeps its global column indices.  Fails with invalid_index_value if a stored column
 * falls outside the window.  col_lo = 0, col_hi = n restores normal behaviour. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_set_x_window(aoclsparse_matrix A,
                                                          aoclsparse_int    col_lo,
                                                          aoclsparse_int    col_hi);

/* Forces row-block boundaries at the given rows (ascending, within (0, m)) so that row ranges can
 * be multiplied separately; must be called before aoclsparse_optimize. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_set_row_cuts(aoclsparse_matrix     A,
                                                          aoclsparse_int        n_cuts,
                                                          const aoclsparse_int *cuts);

/* y[row_begin:row_end) = alpha * A[row_begin:row_end, :] * x + beta * y[row_begin:row_end) for a
 * general matrix, op = none.  row_begin / row_end must be 0, m or one of the row cuts.  Lets the
 * caller overlap boundary rows, the halo exchange and interior rows on different streams. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_dmv_rows(const double              *alpha,
                                                      aoclsparse_matrix          A,
                                                      const aoclsparse_mat_descr descr,
                                                      const double              *x,
                                                      const double              *beta,
                                                      double                    *y,
                                                      aoclsparse_int             row_begin,
                                                      aoclsparse_int             row_end);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_smv_rows(const float               *alpha,
                                                      aoclsparse_matrix          A,
                                                      const aoclsparse_mat_descr descr,
                                                      const float               *x,
                                                      const float               *beta,
                                                      float                     *y,
                                                      aoclsparse_int             row_begin,
                                                      aoclsparse_int             row_end);

/* Fused compute + halo push: as aoclsparse_b200_dmv_rows, and every computed y[r] is ALSO stored to
 * push_dst[r - row_begin].  push_dst may be memory of a PEER GPU mapped into this process
 * (aoclsparse_b200_ipc_open): the boundary rows of a slab write the neighbour's halo of the next x
 * straight from the multiply kernel's epilogue over NVLink, with no separate copy or collective. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_dmv_rows_push(const double              *alpha,
                                                           aoclsparse_matrix          A,
                                                           const aoclsparse_mat_descr descr,
                                                           const double              *x,
                                                           const double              *beta,
                                                           double                    *y,
                                                           aoclsparse_int             row_begin,
                                                           aoclsparse_int             row_end,
                                                           double                    *push_dst);

/* The whole iteration x_{k+1}[own rows] = alpha * A_local * x_k in ONE launch: boundary CTAs (grid order: first
 * boundary, last boundary, interior) wait in-kernel until the facing boundary of the neighbour has completed
 * iteration k-1, store their results both locally and into the neighbour's halo over NVLink, and the last CTA of each
 * side publishes "boundary done k" to that neighbour; interior CTAs overlap with all of that.
 * Needs row cuts {h, m-h}, every row block binned thread-per-row (else aoclsparse_status_not_implemented: use the
 * _rows / _rows_push / _signal / _wait calls), beta = 0.  Flags are 32-bit words in ipc memory; counters are 4 words
 * of local device memory zeroed once before iteration 1 ([3] is set if a flag wait gave up). */
typedef struct aoclsparse_b200_halo_ctl_
{
    const void *left_done, *right_done;       /* local flags written by the neighbours; NULL = no neighbour       */
    void       *to_left_done, *to_right_done; /* the neighbours' flags this rank writes                            */
    void       *counters;
    void       *push_left, *push_right; /* neighbours' halo of x_{k+1} receiving my first / last boundary rows     */
    unsigned    k;                      /* iteration number, 1, 2, 3, ...                                          */
} aoclsparse_b200_halo_ctl;
DLL_PUBLIC aoclsparse_status aoclsparse_b200_dmv_sharded_step(const double                  *alpha,
                                                              aoclsparse_matrix              A,
                                                              const aoclsparse_mat_descr     descr,
                                                              const double                  *x,
                                                              double                        *y,
                                                              const aoclsparse_b200_halo_ctl *ctl);

/* Stream-ordered cross-GPU flags (32-bit words in ipc memory): signal stores `value` after everything the
 * calling thread's stream did before is visible system-wide; wait holds the stream until *flag >= value
 * (wrap-safe), setting *timed_out (may be NULL) and giving up after ~4 s. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_signal(void *flag, unsigned value);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_wait(const void *flag, unsigned value, unsigned *timed_out);

/* Device buffers other processes of this node can map (cudaIpcGetMemHandle / cudaIpcOpenMemHandle). */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_ipc_alloc(size_t bytes, void **dptr, unsigned char handle[64]);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_ipc_open(const unsigned char handle[64], void **dptr);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_ipc_close(void *dptr);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on the calling thread's stream (local or ipc-mapped memory) */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_memcpy(void *dst, const void *src, size_t bytes);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_ipc_free(void *dptr);

/* ---- the row-sharded iteration as one object (csrc/shard.cu) -------------------------------------------------------
 * x <- alpha * A * x, iterated, A split by rows over `world` GPUs of one node, for banded matrices (every stored column
 * within `halo` of its row, e.g. one grid plane of a stencil).  The reference has no multi-device interface (SURVEY.md
 * section 2a); this is the small C extension section 8(e) asks for: create-sharded, iterate-k, gather.
 *
 * Each rank -- a process with its own GPU, or a device driven by the same process -- creates an ordinary handle for
 * its rows [row_lo, row_lo + m) of the n x n matrix (aoclsparse_create_dcsr with m x n and GLOBAL column indices, on
 * the device that is current), then:
 *     aoclsparse_b200_shard_create   sets the x window and row cuts, hints + optimizes the handle, allocates the two
 *                                    x windows and the flag block
 *     aoclsparse_b200_shard_export   fills a 256-byte link record; ranks exchange these by whatever transport they have
 *                                    (MPI_Allgather, torch.distributed, a file): no collective library is linked here
 *     aoclsparse_b200_shard_connect  maps the two neighbours' windows and flags (cudaIpc for other processes, peer
 *                                    access inside one process)
 *     aoclsparse_b200_shard_set_x    own slice of x_0 from a host or device array (or _x_ptr to fill it in place)
 *     aoclsparse_b200_shard_publish  boundary planes of x_0 into the neighbours' halos (stream-ordered, flagged)
 *     aoclsparse_b200_shard_iterate  k iterations, each ONE kernel launch that multiplies, stores the boundary rows into
 *                                    the neighbours' halos over NVLink and hands over the flags (no collective, no
 *                                    host synchronisation; asynchronous on the calling thread's stream, or the shard's
 *                                    own stream when none was set with aoclsparse_b200_set_stream)
 *     aoclsparse_b200_shard_get_x    own slice of the current iterate to a host or device array (synchronises;
 *                                    internal_error if a neighbour never showed up and a flag wait gave up)
 * Ranks need no barrier between these calls; they only must all have finished iterating (get_x / synchronize on every
 * rank) before any of them calls set_x again.  A host thread that drives several shards itself must interleave
 * iterate calls of at most a few dozen iterations per shard (a device queue holds a bounded number of launches, and a
 * launch of one shard waits in-kernel for the previous iteration of its neighbours).
 * double only; world = 1 degenerates to plain aoclsparse_dmv ping-pong. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_device_count(int *count);
/* cudaSetDevice for callers that do not link the CUDA runtime: handles are created on the device that is current */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_set_device(int device);
typedef struct _aoclsparse_b200_shard *aoclsparse_b200_shard;
#define AOCLSPARSE_B200_SHARD_LINK_BYTES 320
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_create(aoclsparse_b200_shard     *shard,
                                                          aoclsparse_matrix          A,
                                                          const aoclsparse_mat_descr descr,
                                                          int                        rank,
                                                          int                        world,
                                                          aoclsparse_int             row_lo,
                                                          aoclsparse_int             halo);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_export(aoclsparse_b200_shard shard,
                                                          unsigned char         link[AOCLSPARSE_B200_SHARD_LINK_BYTES]);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_connect(aoclsparse_b200_shard shard,
                                                           const unsigned char  *left_link,
                                                           const unsigned char  *right_link);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_set_x(aoclsparse_b200_shard shard, const double *x_own);
/* device address of the own slice of the current iterate (m doubles), to be filled / read in place */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_x_ptr(aoclsparse_b200_shard shard, double **own);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_publish(aoclsparse_b200_shard shard);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_iterate(aoclsparse_b200_shard shard, double alpha, int iterations);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_get_x(aoclsparse_b200_shard shard, double *dst);
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_synchronize(aoclsparse_b200_shard shard);
/* frees the windows and flags; the matrix handle stays the caller's */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_shard_destroy(aoclsparse_b200_shard *shard);

/* ---- synthetic matrices of BASELINE.json, generated directly in device memory --------------- */

/* d-dimensional (dims = 2 or 3) stencil on an nx*ny*nz grid (nz = 1 for 2-D), points = 5, 7 or 27,
 * rows [row_lo, row_hi) of the global matrix only (for sharding), diagonal = points-1, off-diagonals
 * = -1, columns sorted, base 0.  Pass row_ptr == NULL to only get *nnz (then allocate and call
 * again).  All pointers are DEVICE pointers; row_ptr has (row_hi-row_lo)+1 entries starting at 0. */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_gen_stencil(int             points,
                                                         aoclsparse_int  nx,
                                                         aoclsparse_int  ny,
                                                         aoclsparse_int  nz,
                                                         long long       row_lo,
                                                         long long       row_hi,
                                                         long long      *nnz,
                                                         aoclsparse_int *row_ptr,
                                                         aoclsparse_int *col_idx,
                                                         double         *val);

/* u(seed, i) = 2 * (splitmix64(seed * 0x100000001B3 ^ i) >> 11) * 2^-53 - 1 for i in [first, first+count),
 * written to a DEVICE array as double (elem_size 8) or float (elem_size 4).  SURVEY.md section 8(d). */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_gen_uniform(unsigned long long seed,
                                                         long long          first,
                                                         long long          count,
                                                         int                elem_size,
                                                         void              *out);

/* R-MAT (Graph500 parameters a,b,c,d = 0.57,0.19,0.19,0.05) edge keys  row << 32 | col  for edges
 * [first, first+count) of a 2^scale-vertex graph, written to a DEVICE int64 array; the caller sorts and
 * de-duplicates them (duplicate diagonal entries are rejected by create). */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_gen_rmat_keys(unsigned long long seed,
                                                           int                scale,
                                                           long long          first,
                                                           long long          count,
                                                           long long         *keys);

/* Sorted, unique keys -> CSR (DEVICE arrays; row_ptr has 2^scale+1 entries, base 0) with values
 * a_ij = (float) u(seed, i * 2^scale + j). */
DLL_PUBLIC aoclsparse_status aoclsparse_b200_rmat_keys_to_csr(unsigned long long seed,
                                                              int                scale,
                                                              long long          count,
                                                              const long long   *keys,
                                                              aoclsparse_int    *row_ptr,
                                                              aoclsparse_int    *col_idx,
                                                              float             *val);

#ifdef __cplusplus
}
#endif
#endif /* AOCLSPARSE_B200_H_ */
