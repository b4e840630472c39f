// mesh_tiles.cu -- csrmm for matrices of regular grids: BOX TILES of rows with their distinct B rows staged once in
// shared memory (north_star: "shared-memory or TMA staging of dense-operand tiles for csrmm").
//
// Why: the row-block csrmm kernel (csrmm.cu) reads one B row (n * sizeof(T) bytes) per stored entry through L1, and the
// L1 delivers at most ~62-70 B/clk/SM for such gathers (profiles/r01_summary.md) -- 0.79 ms on BASELINE config 4 against
// a 0.27 ms HBM floor.  Consecutive rows of a stencil matrix share most of their columns, but a block of R CONSECUTIVE
// rows of a 3-D grid still names ~9 (R + 2) distinct B rows (3.4x the algorithmic traffic out of L2).  A box of
// 8 x 4 x 3 grid points names 300 distinct B rows for its 96 matrix rows (3.1 per row instead of 27 loads per row).
//
// What the analysis does (aoclsparse_optimize's job in the reference is to build format copies too,
// library/src/analysis/aoclsparse_analysis.cpp:192-385):
//   1. the sorted distinct (col - row) offsets of the matrix (plan.cu, probe_diag_offsets; <= 256 of them) are read as a
//      lattice  { a + b*s1 + c*s2 }:  s1 = row stride of the second grid direction, s2 of the third (detect_lattice);
//      row r is the grid point (r % s1, (r / s1) % (s2 / s1), r / s2).  A matrix without that structure keeps the
//      row-block kernel.  Correctness never depends on the guess: tiles are built from the stored columns.
//   2. rows are grouped into boxes X x Y x Z of that grid (RT = X*Y*Z rows, a multiple of 32, <= 96).  Per tile the
//      device collects the columns of its rows (k-way merge of the sorted rows), keeps the distinct ones (= the B rows the
//      tile needs, in ascending order, cut into RUNS of consecutive columns) and rewrites the tile's entries per ROW
//      GROUP (GRP = 2 rows) as a WALK over the union of the group's columns:
//            walk[j][group]  32 bit: 16-bit slot (position of the column among the tile's distinct columns), which rows
//                            of the group store that column, and where their values start; an all-zero entry ends a walk
//            val[i][group]   the values in walk order
//      so the multiply needs neither col_idx nor row_ptr, and one B row read from shared memory serves both rows of a
//      group.  Entries keep their stored (ascending) order inside a row.
//   3. csrmm_mesh_tiles_kernel: PERSISTENT, double-buffered, warp-specialised -- one CTA per SM walks tiles
//      blockIdx.x + k * gridDim.x; its last warp is the producer (TMA bulk copies of the walk / value planes, the row list
//      and -- one copy per run -- the tile's B rows, mbarrier "full" / "empty" per buffer), the other warps are consumers:
//      a team of 8 lanes owns a row group, lane h holds the 16-byte chunks h, h + 8, ... of the group's output rows and
//      reads B out of shared memory with conflict-free 128-byte wavefronts; C rows are stored as whole 128-byte lines.
//      The copies of tile i + 1 overlap the multiply of tile i (see the comment at the kernel).
// Per-row arithmetic (order of the multiply-adds, alpha / beta handling) is that of csrmm_row_major_vec_kernel, so both
// kernels return identical bits.
#include "spmv_kernels.cuh"

#include <algorithm>
#include <climits>
#include <cstdlib>

namespace b200
{
    namespace
    {
        constexpr int RT_MAX    = 96;   // rows per tile (x threads per row = CTA size of the multiply)
        constexpr int KCAP      = 4096; // column keys a tile may hold (RT * longest row)
        constexpr int TILE_NT   = 256;  // threads of the analysis kernels
        constexpr int KEY_EMPTY = INT_MAX;

        struct lattice
        {
            long long m;
            long long s1, s2;     // row strides of the 2nd / 3rd grid direction (the 1st has stride 1)
            int       nx, ny, nz; // grid extent
            int       X, Y, Z;    // box extent
            int       tx, ty, tz; // boxes per direction
        };

        __host__ __device__ inline long long tile_row(const lattice &g, int tile, int local)
        {
            const int bx = tile % g.tx, by = (tile / g.tx) % g.ty, bz = tile / (g.tx * g.ty);
            const int lx = local % g.X, ly = (local / g.X) % g.Y, lz = local / (g.X * g.Y);
            const int x = bx * g.X + lx, y = by * g.Y + ly, z = bz * g.Z + lz;
            if(x >= g.nx || y >= g.ny || z >= g.nz)
                return -1;
            const long long r = (long long)x + (long long)y * g.s1 + (long long)z * g.s2;
            return r < g.m ? r : -1;
        }

        // ascending bitonic sort of keys[0, KCAP) by the whole CTA
        __device__ void sort_keys(int *keys)
        {
            for(int k = 2; k <= KCAP; k <<= 1)
                for(int j = k >> 1; j > 0; j >>= 1)
                {
                    for(int i = threadIdx.x; i < KCAP; i += TILE_NT)
                    {
                        const int ixj = i ^ j;
                        if(ixj > i)
                        {
                            const int  a = keys[i], b = keys[ixj];
                            const bool up = (i & k) == 0;
                            if((a > b) == up)
                            {
                                keys[i]   = b;
                                keys[ixj] = a;
                            }
                        }
                    }
                    __syncthreads();
                }
        }

        constexpr int GRP = 2; // rows of a row group: consecutive rows of a tile that share one walk over their columns

        // columns of the tile's rows into keys[row * lmax + j]; s_info: [0] longest row, [1] rows, [4] a row is not
        // strictly ascending (the group walk needs every row's columns in ascending order, without repeats)
        __device__ void gather_keys(const lattice &g,
                                    int            tile,
                                    int            RT,
                                    int            lmax,
                                    const aoclsparse_int *__restrict__ rp,
                                    const aoclsparse_int *__restrict__ col,
                                    int *keys,
                                    int *s_info)
        {
            for(int i = threadIdx.x; i < KCAP; i += TILE_NT)
                keys[i] = KEY_EMPTY;
            if(threadIdx.x < 8)
                s_info[threadIdx.x] = 0;
            __syncthreads();
            for(int lr = threadIdx.x; lr < RT; lr += TILE_NT)
            {
                const long long r = tile_row(g, tile, lr);
                if(r < 0)
                    continue;
                const int s = rp[r], len = rp[r + 1] - s;
                atomicMax(&s_info[0], len);
                atomicAdd(&s_info[1], 1);
                int prev = -1;
                for(int j = 0; j < len && j < lmax; ++j)
                {
                    const int c        = col[s + j];
                    keys[lr * lmax + j] = c;
                    if(c <= prev)
                        s_info[4] = 1;
                    prev = c;
                }
            }
            __syncthreads();
        }

        // One CTA per tile.  FILL == false: counts[tile] = {distinct columns, runs, U | V << 16, rows | unusable << 30}
        // (U / V = longest union walk / value stream of a row group).  FILL == true: writes the tile's arrays.
        //   distinct columns: the tile's columns sorted, repeats removed (ukeys); a RUN is a maximal stretch of consecutive
        //   columns, stored as (first column, first slot), the tile's list closed by (-1, distinct).
        //   row group grp = rows GRP*grp .. GRP*grp+GRP-1 of the tile: its WALK is the ascending list of the columns at
        //   least one of its rows stores, each as slot | rowmask << 16 | (position of its first value) << 20, closed by a
        //   zero entry; its VALUE STREAM holds, walk entry by walk entry
        //   and row by row (ascending), the stored values.  Both are laid out as planes over the groups:
        //   walk[j][grp] at sm_off + j*NG + grp (zero = padding), val[i][grp] at val_off + i*NG + grp.
        template <int ES, bool FILL>
        __global__ void __launch_bounds__(TILE_NT) tile_build_kernel(lattice g,
                                                                     int     RT,
                                                                     int     lmax,
                                                                     const aoclsparse_int *__restrict__ rp,
                                                                     const aoclsparse_int *__restrict__ col,
                                                                     const unsigned char *__restrict__ val,
                                                                     int4 *counts,
                                                                     const int4 *__restrict__ tdesc,
                                                                     const long long *__restrict__ toff,
                                                                     unsigned      *twalk,
                                                                     unsigned char *tval,
                                                                     int           *trows,
                                                                     int2          *truns)
        {
            __shared__ int keys[KCAP];
            __shared__ int ukeys[KCAP];
            __shared__ int s_info[8];
            __shared__ int s_scan[TILE_NT + 1], s_rscan[TILE_NT + 1];
            const int      tile = blockIdx.x, tid = threadIdx.x;
            const int      NG   = RT / GRP;
            gather_keys(g, tile, RT, lmax, rp, col, keys, s_info);
            sort_keys(keys);
            // compaction: thread t owns keys[t*PER, (t+1)*PER)
            constexpr int PER = KCAP / TILE_NT;
            int           nd = 0, nr = 0;
            for(int q = 0; q < PER; ++q)
            {
                const int i = tid * PER + q, k = keys[i];
                if(k == KEY_EMPTY)
                    continue;
                const int prev = i > 0 ? keys[i - 1] : KEY_EMPTY;
                if(i == 0 || k != prev)
                {
                    ++nd;
                    if(i == 0 || k != prev + 1)
                        ++nr;
                }
            }
            s_scan[tid + 1]  = nd;
            s_rscan[tid + 1] = nr;
            if(tid == 0)
                s_scan[0] = s_rscan[0] = 0;
            __syncthreads();
            if(tid == 0)
                for(int t = 1; t <= TILE_NT; ++t)
                {
                    s_scan[t] += s_scan[t - 1];
                    s_rscan[t] += s_rscan[t - 1];
                }
            __syncthreads();
            const int n_distinct = s_scan[TILE_NT], n_runs = s_rscan[TILE_NT];
            int4      d          = make_int4(0, 0, 0, 0);
            if(FILL)
                d = tdesc[tile]; // {distinct, runs, U | V << 16, first run}
            {
                int u = s_scan[tid], rn = s_rscan[tid];
                for(int q = 0; q < PER; ++q)
                {
                    const int i = tid * PER + q, k = keys[i];
                    if(k == KEY_EMPTY)
                        continue;
                    const int prev = i > 0 ? keys[i - 1] : KEY_EMPTY;
                    if(i == 0 || k != prev)
                    {
                        ukeys[u] = k;
                        if(i == 0 || k != prev + 1)
                        {
                            if(FILL)
                                truns[d.w + rn] = make_int2(k, u); // first column of the run, its first slot
                            ++rn;
                        }
                        ++u;
                    }
                }
            }
            if(FILL && tid == 0)
                truns[d.w + n_runs] = make_int2(-1, n_distinct); // terminator: the slot one past the last run
            __syncthreads();
            // one thread per row group: k-way merge of its rows' (strictly ascending) column lists
            const int       Ut = FILL ? (d.z & 0xffff) : 0, Vt = FILL ? (int)((unsigned)d.z >> 16) : 0;
            const long long sm_off = FILL ? toff[2 * tile] : 0, val_off = FILL ? toff[2 * tile + 1] : 0;
            for(int grp = tid; grp < NG; grp += TILE_NT)
            {
                int U = 0, V = 0;
                int start[GRP], len[GRP], used[GRP];
                for(int i = 0; i < GRP; ++i)
                {
                    const long long r = tile_row(g, tile, grp * GRP + i);
                    start[i]          = r >= 0 ? rp[r] : 0;
                    len[i]            = r >= 0 ? min(rp[r + 1] - rp[r], lmax) : 0;
                    used[i]           = 0;
                }
                for(;;)
                {
                    int c = KEY_EMPTY;
                    for(int i = 0; i < GRP; ++i)
                        if(used[i] < len[i])
                            c = min(c, col[start[i] + used[i]]);
                    if(c == KEY_EMPTY)
                        break;
                    unsigned m = 0;
                    for(int i = 0; i < GRP; ++i)
                        if(used[i] < len[i] && col[start[i] + used[i]] == c)
                            m |= 1u << i;
                    if(FILL)
                    {
                        int lo = 0, hi = n_distinct - 1;
                        while(lo < hi)
                        {
                            const int mid = (lo + hi) >> 1;
                            if(ukeys[mid] < c)
                                lo = mid + 1;
                            else
                                hi = mid;
                        }
                        twalk[sm_off + (long long)U * NG + grp] = (unsigned)lo | (m << 16) | ((unsigned)V << 20);
                    }
                    for(int i = 0; i < GRP; ++i)
                        if(m & (1u << i))
                        {
                            if(FILL)
                            {
                                const long long src = (long long)(start[i] + used[i]) * ES, dst = (val_off + (long long)V * NG + grp) * ES;
                                for(int b = 0; b < ES; b += 4)
                                    *reinterpret_cast<unsigned *>(tval + dst + b) = *reinterpret_cast<const unsigned *>(val + src + b);
                            }
                            ++used[i];
                            ++V;
                        }
                    ++U;
                }
                if(FILL)
                {
                    for(int j = U; j < Ut; ++j)
                        twalk[sm_off + (long long)j * NG + grp] = 0u;
                    for(int i = V; i < Vt; ++i)
                        for(int b = 0; b < ES; b += 4)
                            *reinterpret_cast<unsigned *>(tval + (val_off + (long long)i * NG + grp) * ES + b) = 0u;
                }
                else
                {
                    atomicMax(&s_info[5], U + 1); // + the all-zero entry that ends the walk
                    atomicMax(&s_info[6], V);
                }
            }
            if(FILL)
                for(int lr = tid; lr < RT; lr += TILE_NT)
                    trows[(long long)tile * RT + lr] = (int)tile_row(g, tile, lr);
            __syncthreads();
            if(!FILL && tid == 0)
            {
                const int bad = (s_info[4] || s_info[0] > lmax) ? 1 : 0;
                counts[tile]  = make_int4(n_distinct, n_runs, s_info[5] | (s_info[6] << 16), s_info[1] | (bad << 30));
            }
        }

        // ---------------------------------------------------------------------------------------------------------
        template <typename T>
        struct __align__(16) chunk16
        {
            static constexpr int N = 16 / sizeof(T);
            T                    v[N];
        };

        __device__ __forceinline__ void mbar_arrive(uint64_t *bar)
        {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        }

        // PERSISTENT, DOUBLE BUFFERED, WARP SPECIALISED: a CTA (one per SM) walks tiles blockIdx.x, blockIdx.x + gridDim.x,
        // ...  Its last warp is the PRODUCER: for every tile it waits until the staging buffer is empty, announces the
        // tile's byte count on the buffer's "full" mbarrier and issues the TMA copies (walk / value planes, row list, one
        // copy per run of B rows).  The other warps are CONSUMERS: wait for "full", multiply out of shared memory, arrive
        // (one lane per warp) on the buffer's "empty" mbarrier, store their part of C.  No CTA-wide barrier: a warp that
        // finishes a tile early starts the next one while the others still multiply or store, and the copies of tile i+1
        // overlap the multiply of tile i.
        // A TEAM of 8 lanes (a quarter-warp) owns one row group: lane h holds the 16-byte chunks h, h+8, ... of the GRP
        // output rows.  Per walk entry the team reads ONE B row out of shared memory (8 lanes x 16 bytes = one 128-byte
        // wavefront per chunk index, conflict-free) and uses it for every row of the group that stores that column: B rows
        // are re-used in registers across the rows of a group (a 27-point stencil: 36 B-row reads per 2 rows instead of 54).
        // Values are broadcast loads (the 8 lanes of a team read the same address).
        // Every output entry is formed by the multiply-adds of its row in ascending column order, the order the row-block
        // kernel uses for sorted rows: identical bits.
        template <typename T, int CH>
        __global__ void __launch_bounds__(RT_MAX / GRP * 8 + 32, 1)
            csrmm_mesh_tiles_kernel(const int4 *__restrict__ tdesc,
                                    const long long *__restrict__ toff,
                                    const unsigned *__restrict__ twalk,
                                    const T *__restrict__ tval,
                                    const int *__restrict__ trows,
                                    const int2 *__restrict__ truns,
                                    const T *__restrict__ B,
                                    long long ldb,
                                    T *__restrict__ C,
                                    long long ldc,
                                    int       n_tiles,
                                    int       RT,
                                    int       btile_rows, // rows of a staged B tile buffer
                                    int       max_walk,
                                    int       max_vals,
                                    int       n_buf,      // staging buffers: 2, or 1 when two do not fit
                                    T         alpha,
                                    T         beta,
                                    int       beta_zero,
                                    int       b_contiguous)
        {
            constexpr int VEC       = chunk16<T>::N;
            constexpr int W         = 8 * CH; // 16-byte chunks per B / C row
            constexpr int ROW_BYTES = W * 16;
            extern __shared__ __align__(128) unsigned char smem_raw[];
            uint64_t    *full     = reinterpret_cast<uint64_t *>(smem_raw);     // full[0], full[1]
            uint64_t    *empty    = reinterpret_cast<uint64_t *>(smem_raw) + 2; // empty[0], empty[1]
            const int    NG       = RT / GRP;
            const size_t walk_b   = (size_t)max_walk * NG * 4, vals_b = (size_t)max_vals * NG * sizeof(T);
            const size_t buf_size = ((size_t)btile_rows * ROW_BYTES + walk_b + vals_b + (size_t)RT * 4 + 127) & ~(size_t)127;
            auto         btile_of = [&](int q) { return smem_raw + 128 + (size_t)q * buf_size; };

            const int tid = threadIdx.x, n_cons = NG * 8, lane = tid & 31;
            const int stride = (int)gridDim.x;
            if(tid == 0)
            {
                mbar_init(full, 1);
                mbar_init(full + 1, 1);
                mbar_init(empty, (unsigned)(n_cons / 32));
                mbar_init(empty + 1, (unsigned)(n_cons / 32));
                mbar_init_fence();
            }
            __syncthreads();

            if(tid >= n_cons)
            {
                // ---------------- producer warp
                int it = 0;
                for(int tile = (int)blockIdx.x; tile < n_tiles; tile += stride, ++it)
                {
                    const int q = n_buf == 2 ? (it & 1) : 0;
                    const int u = n_buf == 2 ? (it >> 1) : it; // how often buffer q has been used before
                    if(u > 0)
                        mbar_wait(empty + q, (unsigned)(u - 1) & 1u);
                    unsigned char *btile = btile_of(q);
                    unsigned      *swalk = reinterpret_cast<unsigned *>(btile + (size_t)btile_rows * ROW_BYTES);
                    T             *sval  = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(swalk) + walk_b);
                    int           *srows = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(sval) + vals_b);
                    const int4     d     = tdesc[tile]; // {distinct, runs, U | V << 16, first run}
                    if(lane == 0)
                    {
                        const long long sm_off = toff[2 * tile], val_off = toff[2 * tile + 1];
                        const unsigned  U = (unsigned)d.z & 0xffffu, V = (unsigned)d.z >> 16;
                        const unsigned  total = (unsigned)d.x * ROW_BYTES + U * (unsigned)NG * 4u + V * (unsigned)NG * (unsigned)sizeof(T) + (unsigned)RT * 4u;
                        mbar_expect_tx(full + q, total);
                        if(U > 0)
                            bulk_load_stream(swalk, twalk + sm_off, U * (unsigned)NG * 4u, full + q);
                        if(V > 0)
                            bulk_load_stream(sval, tval + val_off, V * (unsigned)NG * (unsigned)sizeof(T), full + q);
                        bulk_load_stream(srows, trows + (long long)tile * RT, (unsigned)(RT * 4), full + q);
                    }
                    __syncwarp(); // the byte count is announced before any other lane's copy can complete
                    // B rows: one bulk copy per run of consecutive columns (per row when B's rows are not adjacent in memory)
                    for(int i = lane; i < d.y; i += 32)
                    {
                        const int2 run = truns[d.w + i];
                        const int  cnt = truns[d.w + i + 1].y - run.y;
                        if(b_contiguous)
                            bulk_load(btile + (size_t)run.y * ROW_BYTES, B + (long long)run.x * ldb, (unsigned)cnt * ROW_BYTES, full + q);
                        else
                            for(int qq = 0; qq < cnt; ++qq)
                                bulk_load(btile + (size_t)(run.y + qq) * ROW_BYTES, B + (long long)(run.x + qq) * ldb, ROW_BYTES, full + q);
                    }
                }
                return;
            }

            // ---------------- consumer warps: team = tid / 8 owns row group `team`, lane h = tid % 8
            const int grp = tid >> 3, h = tid & 7;
            int       it  = 0;
            for(int tile = (int)blockIdx.x; tile < n_tiles; tile += stride, ++it)
            {
                const int            q     = n_buf == 2 ? (it & 1) : 0;
                const int            u     = n_buf == 2 ? (it >> 1) : it;
                const unsigned char *btile = btile_of(q);
                const unsigned      *swalk = reinterpret_cast<const unsigned *>(btile + (size_t)btile_rows * ROW_BYTES);
                const T             *sval  = reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(swalk) + walk_b);
                const int           *srows = reinterpret_cast<const int *>(reinterpret_cast<const unsigned char *>(sval) + vals_b);
                mbar_wait(full + q, (unsigned)u & 1u);

                int rows[GRP];
#pragma unroll
                for(int i = 0; i < GRP; ++i)
                    rows[i] = srows[grp * GRP + i];
                chunk16<T> acc[GRP][CH];
#pragma unroll
                for(int i = 0; i < GRP; ++i)
#pragma unroll
                    for(int k = 0; k < CH; ++k)
#pragma unroll
                        for(int qq = 0; qq < VEC; ++qq)
                            acc[i][k].v[qq] = vt<T>::zero();
                const unsigned      *wp = swalk + grp;
                const T             *vg = sval + grp;
                const unsigned char *bh = btile + h * 16;
                unsigned             w  = *wp;
                // a walk entry: slot (bits 0-15) | row mask (16-19) | position of its first value in the group's value
                // stream (20-31); an all-zero entry ends the walk (every tile stores one after its longest walk)
                while(w >> 16)
                {
                    const unsigned char *rowp = bh + (size_t)(w & 0xffffu) * ROW_BYTES;
                    const T             *vp   = vg + (w >> 20) * NG;
                    const unsigned       m    = w >> 16;
                    wp += NG;
                    w = *wp; // next walk entry, in flight during the multiply-adds
                    chunk16<T> x[CH];
#pragma unroll
                    for(int k = 0; k < CH; ++k)
                        x[k] = *reinterpret_cast<const chunk16<T> *>(rowp + k * 128);
                    // the entry's values sit at vp[0 .. popc(mask)) (plane stride NG), in row order; a load past the group's
                    // stream stays inside the staging buffer and is never used
                    T v[GRP];
#pragma unroll
                    for(int i = 0; i < GRP; ++i)
                        v[i] = vp[__popc(m & ((1u << i) - 1u)) * NG];
#pragma unroll
                    for(int i = 0; i < GRP; ++i)
                        if(m & (1u << i))
                        {
#pragma unroll
                            for(int k = 0; k < CH; ++k)
#pragma unroll
                                for(int qq = 0; qq < VEC; ++qq)
                                    acc[i][k].v[qq] = mad(v[i], x[k].v[qq], acc[i][k].v[qq]);
                        }
                }
                // this warp is done with the buffer
                __syncwarp();
                if(lane == 0)
                    mbar_arrive(empty + q);
                // the team's rows of C: 8 lanes x 16 bytes = 128 contiguous bytes per store
#pragma unroll
                for(int i = 0; i < GRP; ++i)
                {
                    if(rows[i] < 0)
                        continue;
                    T *crow = C + (long long)rows[i] * ldc;
#pragma unroll
                    for(int k = 0; k < CH; ++k)
                    {
                        T         *cp = crow + (h + 8 * k) * VEC;
                        chunk16<T> o;
#pragma unroll
                        for(int qq = 0; qq < VEC; ++qq)
                            o.v[qq] = mul(alpha, acc[i][k].v[qq]);
                        if(!beta_zero)
                        {
                            const chunk16<T> old = *reinterpret_cast<const chunk16<T> *>(cp);
#pragma unroll
                            for(int qq = 0; qq < VEC; ++qq)
                                o.v[qq] = mad(beta, old.v[qq], o.v[qq]);
                        }
                        *reinterpret_cast<chunk16<T> *>(cp) = o;
                    }
                }
            }
        }

        // one staging buffer (B tile + the tile's walk / value planes + row list), rounded to 128 bytes
        size_t tile_buffer_bytes(int RT, int max_distinct, int max_walk, int max_vals, size_t row_bytes, size_t elem_size)
        {
            const size_t NG = (size_t)RT / GRP;
            return ((size_t)max_distinct * row_bytes + (size_t)max_walk * NG * 4 + (size_t)max_vals * NG * elem_size + (size_t)RT * 4 + 127) & ~(size_t)127;
        }
        size_t tile_smem_bytes(int RT, int max_distinct, int max_walk, int max_vals, size_t row_bytes, size_t elem_size)
        {
            return 128 + tile_buffer_bytes(RT, max_distinct, max_walk, max_vals, row_bytes, elem_size);
        }
    }

    // sorted distinct offsets -> lattice strides.  Accepts offset sets of the form { a + b*s1 + c*s2 : |a| <= ra, |b| <= rb,
    // |c| <= rc } (box stencils) and subsets of them that still contain +-1, +-s1, +-s2 (star stencils); ndim = 1..3
    bool detect_lattice(const std::vector<int> &offs, long long m, long long &s1, long long &s2, int &ndim)
    {
        s1 = s2 = 0;
        ndim    = 0;
        if(offs.empty() || m <= 0)
            return false;
        std::vector<long long> pos;
        for(int o : offs)
            if(o > 0)
                pos.push_back(o);
        auto has = [&](long long o) { return o >= INT_MIN && o <= INT_MAX && std::binary_search(offs.begin(), offs.end(), (int)o); };
        if(pos.empty())
        {
            ndim = 1; // diagonal / lower banded with unit steps only
            return true;
        }
        if(pos[0] != 1 && !(has(-1)))
        {
            // no unit step at all: still a 1-d "lattice" of consecutive rows (block rows share few columns, harmless)
            ndim = 1;
            return true;
        }
        long long r0 = 0; // radius along the first direction
        while(has(r0 + 1) || has(-(r0 + 1)))
            ++r0;
        // first offset beyond the first direction's reach
        auto next_beyond = [&](long long reach) -> long long {
            for(long long o : pos)
                if(o > reach)
                    return o;
            return 0;
        };
        long long q = next_beyond(r0);
        if(q == 0)
        {
            ndim = 1;
            return true;
        }
        // s1 is one of q .. q + r0 (q = s1 - a for a box stencil); it must be a stored offset or its mirror must be, and
        // the offsets within r0 of it must be symmetric about it
        auto symmetric_about = [&](long long c, long long reach) {
            for(long long a = 1; a <= reach; ++a)
                if(has(c + a) != has(c - a))
                    return false;
            return true;
        };
        for(long long a = 0; a <= r0 && s1 == 0; ++a)
            if((has(q + a) || has(-(q + a))) && symmetric_about(q + a, r0))
                s1 = q + a;
        if(s1 <= r0 || s1 > m)
            return false;
        long long r1 = 1;
        while(has((r1 + 1) * s1) || has(-(r1 + 1) * s1))
            ++r1;
        const long long reach1 = r1 * s1 + r0;
        q                      = next_beyond(reach1);
        if(q == 0)
        {
            ndim = 2;
            return true;
        }
        for(long long b = 0; b <= r1 && s2 == 0; ++b)
            for(long long a = 0; a <= r0 && s2 == 0; ++a)
            {
                const long long c = q + b * s1 + a;
                if((has(c) || has(-c)) && symmetric_about(c, r0) && (has(c + s1) || has(-(c + s1))) == (has(c - s1) || has(-(c - s1))))
                    s2 = c;
            }
        if(s2 <= reach1 || s2 % s1 != 0 || s2 > m)
            return false;
        // nothing may lie beyond the third direction's reach of a few planes
        long long r2 = 1;
        while(has((r2 + 1) * s2) || has(-(r2 + 1) * s2))
            ++r2;
        if(next_beyond(r2 * s2 + reach1) != 0)
            return false;
        ndim = 3;
        return true;
    }

    namespace
    {
        // box extents for a grid of `ndim` directions: X = 8 along the unit-stride direction, (Y, Z) the pair with the
        // fewest distinct B rows per matrix row (estimated with a one-point halo) whose tile fits `budget` bytes; the walk
        // of a row group is estimated as max_len * (GRP + 2) / 3 entries (a row of a radius-1 stencil shares two thirds of
        // its columns with its neighbour), its value stream as GRP * max_len
        void choose_box(int ndim, int nx, int ny, int nz, int max_len, size_t row_bytes, size_t elem_size, size_t budget, int box[3])
        {
            if(ndim == 1)
            {
                box[0] = 64;
                box[1] = box[2] = 1;
                return;
            }
            const int X     = 8;
            double    best  = 1e30;
            box[0]          = X;
            box[1]          = 4;
            box[2]          = 1;
            const int ys[]  = {1, 2, 4, 8, 16};
            const int zs[]  = {1, 2, 3, 4, 6, 8};
            const int walk  = (max_len * (GRP + 2) + 2) / 3, vals = GRP * max_len;
            for(int Y : ys)
                for(int Z : zs)
                {
                    if(ndim == 2 && Z != 1)
                        continue;
                    const int RT = X * Y * Z;
                    if(RT % 32 != 0 || RT > RT_MAX || RT * max_len > KCAP)
                        continue;
                    if(Y > std::max(ny, 1) * 2 || Z > std::max(nz, 1) * 2)
                        continue;
                    const int    distinct = (X + 2) * (Y + 2) * (ndim == 3 ? Z + 2 : 1);
                    const size_t smem     = tile_smem_bytes(RT, distinct, walk, vals, row_bytes, elem_size);
                    if(smem > budget)
                        continue;
                    const double cost = (double)distinct / RT;
                    if(cost < best - 1e-9)
                    {
                        best   = cost;
                        box[1] = Y;
                        box[2] = Z;
                    }
                }
        }
    }

    size_t mesh_tiles_smem(const mesh_tiles &M, size_t row_bytes, size_t elem_size)
    {
        return tile_smem_bytes(M.rows_per_tile, M.max_distinct, M.max_walk, M.max_vals, row_bytes, elem_size);
    }

    aoclsparse_status build_mesh_tiles(const dev_csr &A, size_t elem_size, size_t row_bytes, cudaStream_t st)
    {
        mesh_tiles &M = A.tiles;
        M             = mesh_tiles();
        M.state       = 1; // tried; stays "unusable" unless everything below succeeds
        M.row_bytes   = row_bytes;
        if(A.m <= 0 || A.nnz <= 0)
            return aoclsparse_status_success;
        std::vector<int> offs;
        B200_TRY(probe_diag_offsets(A, offs, st));
        if(offs.empty())
            return aoclsparse_status_success;
        long long s1 = 0, s2 = 0;
        int       ndim = 0;
        if(!detect_lattice(offs, A.m, s1, s2, ndim))
            return aoclsparse_status_success;
        // longest row: the offsets are distinct per row (rows without repeated columns), so their number bounds it
        const int lmax = (int)offs.size();
        if(lmax > 64)
            return aoclsparse_status_success;
        lattice g;
        g.m  = A.m;
        g.s1 = ndim >= 2 ? s1 : (long long)A.m;
        g.s2 = ndim >= 3 ? s2 : (ndim >= 2 ? ((A.m + s1 - 1) / s1) * s1 : (long long)A.m);
        g.nx = (int)std::min<long long>(g.s1, A.m);
        g.ny = ndim >= 2 ? (int)(ndim >= 3 ? s2 / s1 : (A.m + s1 - 1) / s1) : 1;
        g.nz = ndim >= 3 ? (int)((A.m + s2 - 1) / s2) : 1;
        size_t budget = 112 * 1024;
        if(const char *e = getenv("AOCLSPARSE_B200_MM_TILE_SMEM"))
            budget = (size_t)atoll(e);
        int box[3];
        choose_box(ndim, g.nx, g.ny, g.nz, lmax, row_bytes, elem_size, budget, box);
        if(const char *e = getenv("AOCLSPARSE_B200_MM_TILE_BOX"))
        {
            int a = 0, b = 0, c = 0;
            if(sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a > 0 && b > 0 && c > 0 && (a * b * c) % 32 == 0 && a * b * c <= RT_MAX)
            {
                box[0] = a;
                box[1] = b;
                box[2] = c;
            }
        }
        dev_buf           d_counts;
        std::vector<int4> counts;
        int               RT = 0;
        for(int attempt = 0; attempt < 4; ++attempt)
        {
            g.X = box[0];
            g.Y = box[1];
            g.Z = box[2];
            RT  = g.X * g.Y * g.Z;
            if(RT % 32 != 0 || RT > RT_MAX || (long long)RT * lmax > KCAP)
                return aoclsparse_status_success;
            g.tx               = (g.nx + g.X - 1) / g.X;
            g.ty               = (g.ny + g.Y - 1) / g.Y;
            g.tz               = (g.nz + g.Z - 1) / g.Z;
            const long long nt = (long long)g.tx * g.ty * g.tz;
            if(nt <= 0 || nt > (1ll << 30))
                return aoclsparse_status_success;
            B200_TRY(d_counts.alloc(sizeof(int4) * (size_t)nt));
            tile_build_kernel<4, false><<<(unsigned)nt, TILE_NT, 0, st>>>(g, RT, lmax, A.row_ptr.as<aoclsparse_int>(),
                                                                          A.col_idx.as<aoclsparse_int>(), nullptr, d_counts.as<int4>(),
                                                                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            B200_LAUNCHED();
            counts.resize((size_t)nt);
            B200_CUDA(cudaMemcpyAsync(counts.data(), d_counts.p, sizeof(int4) * (size_t)nt, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            int       max_d = 0, max_u = 0, max_v = 0, max_r = 0, bad = 0;
            long long rows = 0;
            for(const int4 &c : counts)
            {
                max_d = std::max(max_d, c.x);
                max_r = std::max(max_r, c.y);
                max_u = std::max(max_u, c.z & 0xffff);
                max_v = std::max(max_v, (int)((unsigned)c.z >> 16));
                rows += c.w & 0x3fffffff;
                bad |= c.w >> 30;
            }
            // not a partition of the rows, a row that is not strictly ascending or longer than the offset table: no tiles
            if(rows != A.m || bad || max_d > 65535)
                return aoclsparse_status_success;
            M.max_distinct = max_d;
            M.max_walk     = max_u;
            M.max_vals     = max_v;
            M.max_runs     = max_r;
            M.n_tiles      = (int)nt;
            if(tile_smem_bytes(RT, max_d, max_u, max_v, row_bytes, elem_size) <= budget)
                break;
            // the real tiles need more than the estimate: halve the box along its last direction that is > 1
            if(box[2] > 1)
                box[2] = (box[2] + 1) / 2;
            else if(box[1] > 1)
                box[1] /= 2;
            else
                return aoclsparse_status_success;
            if(attempt == 3)
                return aoclsparse_status_success;
        }
        M.box[0] = g.X;
        M.box[1] = g.Y;
        M.box[2] = g.Z;
        M.stride[0]     = 1;
        M.stride[1]     = g.s1;
        M.stride[2]     = g.s2;
        M.dims[0]       = g.nx;
        M.dims[1]       = g.ny;
        M.dims[2]       = g.nz;
        M.rows_per_tile = RT;
        const int NG    = RT / GRP;
        // staging is only worth it when a tile re-uses its B rows: distinct rows per matrix row well below the row length
        {
            double sum_d = 0, sum_v = 0;
            for(const int4 &c : counts)
            {
                sum_d += c.x;
                sum_v += (double)((unsigned)c.z >> 16) * NG;
            }
            const double reuse = (double)A.nnz / std::max(1.0, sum_d); // stored entries per staged B row
            const double fill  = (double)A.nnz / std::max(1.0, sum_v); // stored entries per slot of the value planes
            M.reuse            = reuse;
            M.fill             = fill;
            if(reuse < 3.0 || fill < 0.7)
                return aoclsparse_status_success;
        }
        // per-tile descriptors: {distinct, runs, U | V << 16, first run (+1 terminator per tile)}; first walk / value slot
        std::vector<int4>      hdesc(counts.size());
        std::vector<long long> hoff(2 * counts.size());
        long long              walk = 0, vals = 0, run = 0;
        for(size_t t = 0; t < counts.size(); ++t)
        {
            hdesc[t]        = make_int4(counts[t].x, counts[t].y, counts[t].z, (int)run);
            hoff[2 * t]     = walk;
            hoff[2 * t + 1] = vals;
            walk += (long long)(counts[t].z & 0xffff) * NG;
            vals += (long long)((unsigned)counts[t].z >> 16) * NG;
            run += counts[t].y + 1;
            if(run > INT_MAX - 4096)
                return aoclsparse_status_success;
        }
        M.walk_entries = walk;
        M.val_entries  = vals;
        B200_TRY(M.desc.alloc(sizeof(int4) * hdesc.size()));
        B200_TRY(M.off.alloc(sizeof(long long) * hoff.size()));
        B200_TRY(M.walk.alloc((size_t)std::max<long long>(walk, 1) * 4));
        B200_TRY(M.val.alloc((size_t)std::max<long long>(vals, 1) * elem_size));
        B200_TRY(M.rows.alloc(sizeof(int) * (size_t)M.n_tiles * RT));
        B200_TRY(M.runs.alloc(sizeof(int2) * (size_t)run));
        M.n_runs_total = run;
        B200_CUDA(cudaMemcpyAsync(M.desc.p, hdesc.data(), sizeof(int4) * hdesc.size(), cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(M.off.p, hoff.data(), sizeof(long long) * hoff.size(), cudaMemcpyHostToDevice, st));
#define B200_TILE_FILL(ES)                                                                                                      \
    tile_build_kernel<ES, true><<<(unsigned)M.n_tiles, TILE_NT, 0, st>>>(g, RT, lmax, A.row_ptr.as<aoclsparse_int>(),          \
                                                                          A.col_idx.as<aoclsparse_int>(), A.val.as<unsigned char>(), \
                                                                          nullptr, M.desc.as<int4>(), M.off.as<long long>(),      \
                                                                          M.walk.as<unsigned>(), M.val.as<unsigned char>(),       \
                                                                          M.rows.as<int>(), M.runs.as<int2>())
        if(elem_size == 4)
            B200_TILE_FILL(4);
        else if(elem_size == 8)
            B200_TILE_FILL(8);
        else
            B200_TILE_FILL(16);
#undef B200_TILE_FILL
        B200_LAUNCHED();
        B200_CUDA(cudaStreamSynchronize(st)); // hdesc / hoff are read by the copies above
        M.state = 2;
        return aoclsparse_status_success;
    }

    template <typename T>
    aoclsparse_status launch_mm_tiles(const dev_csr &A, const T *B, long long ldb, T *C, long long ldc, int n, T alpha, T beta, cudaStream_t st)
    {
        const mesh_tiles &M         = A.tiles;
        const size_t      row_bytes = (size_t)n * sizeof(T);
        const int         RT        = M.rows_per_tile;
        const size_t      one       = tile_buffer_bytes(RT, M.max_distinct, M.max_walk, M.max_vals, row_bytes, sizeof(T));
        const int         n_buf     = (128 + 2 * one <= (size_t)227 * 1024) ? 2 : 1;
        const size_t      smem      = 128 + (size_t)n_buf * one;
        const int         bz        = is_zero(beta) ? 1 : 0;
        const int         contig    = ldb == (long long)n ? 1 : 0;
        int               sms       = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid    = std::min(M.n_tiles, sms);
        const int threads = RT / GRP * 8 + 32;
#define B200_MM_TILES(CH)                                                                                                \
    {                                                                                                                    \
        static std::atomic<size_t> cfg{0};                                                                               \
        if(cfg.load() < smem)                                                                                            \
        {                                                                                                                \
            B200_CUDA(cudaFuncSetAttribute(csrmm_mesh_tiles_kernel<T, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            cfg.store(smem);                                                                                             \
        }                                                                                                                \
        csrmm_mesh_tiles_kernel<T, CH><<<(unsigned)grid, threads, smem, st>>>(                                           \
            M.desc.as<int4>(), M.off.as<long long>(), M.walk.as<unsigned>(), M.val.as<T>(), M.rows.as<int>(),             \
            M.runs.as<int2>(), B, ldb, C, ldc, M.n_tiles, RT, M.max_distinct, M.max_walk, M.max_vals, n_buf, alpha, beta, \
            bz, contig);                                                                                                 \
    }
        if(row_bytes == 128)
            B200_MM_TILES(1)
        else if(row_bytes == 256)
            B200_MM_TILES(2)
        else if(row_bytes == 512)
            B200_MM_TILES(4)
        else
            return aoclsparse_status_internal_error;
#undef B200_MM_TILES
        B200_LAUNCHED();
        return aoclsparse_status_success;
    }

    template aoclsparse_status launch_mm_tiles<float>(const dev_csr &, const float *, long long, float *, long long, int, float, float, cudaStream_t);
    template aoclsparse_status launch_mm_tiles<double>(const dev_csr &, const double *, long long, double *, long long, int, double, double, cudaStream_t);
    template aoclsparse_status launch_mm_tiles<float2>(const dev_csr &, const float2 *, long long, float2 *, long long, int, float2, float2, cudaStream_t);
    template aoclsparse_status launch_mm_tiles<double2>(const dev_csr &, const double2 *, long long, double2 *, long long, int, double2, double2, cudaStream_t);
}

extern "C" {
aoclsparse_status aoclsparse_b200_get_mm_tiles_info(const aoclsparse_matrix A, aoclsparse_b200_mm_tiles_info *info)
{
    if(!A || !info)
        return aoclsparse_status_invalid_pointer;
    if(A->mats.empty() || A->mats[0] == nullptr)
        return aoclsparse_status_invalid_pointer;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const b200::dev_csr                &D = *A->mats[0];
    std::lock_guard<std::mutex>         lk(D.tiles_mu);
    const b200::mesh_tiles             &M = D.tiles;
    memset(info, 0, sizeof(*info));
    info->state = M.state;
    for(int i = 0; i < 3; ++i)
    {
        info->box[i]    = M.box[i];
        info->stride[i] = M.stride[i];
        info->dims[i]   = M.dims[i];
    }
    info->rows_per_tile  = M.rows_per_tile;
    info->rows_per_group = 2;
    info->n_tiles        = M.n_tiles;
    info->max_distinct   = M.max_distinct;
    info->max_walk       = M.max_walk;
    info->max_vals       = M.max_vals;
    info->max_runs       = M.max_runs;
    info->walk_entries   = M.walk_entries;
    info->val_entries    = M.val_entries;
    info->n_runs_total   = M.n_runs_total;
    info->row_bytes      = (long long)M.row_bytes;
    info->reuse          = M.reuse;
    info->fill           = M.fill;
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_get_mm_tiles(const aoclsparse_matrix A, int *desc, long long *off, unsigned *walk, void *val, int *rows, int *runs)
{
    if(!A || A->mats.empty() || A->mats[0] == nullptr)
        return aoclsparse_status_invalid_pointer;
    std::shared_lock<std::shared_mutex> rl(A->guard);
    const b200::dev_csr                &D = *A->mats[0];
    std::lock_guard<std::mutex>         lk(D.tiles_mu);
    const b200::mesh_tiles             &M = D.tiles;
    if(M.state != 2)
        return aoclsparse_status_invalid_operation;
    cudaStream_t st = b200::current_stream();
    const size_t es = b200::value_size(A->val_type);
    const size_t tr = (size_t)M.n_tiles * M.rows_per_tile;
    if(desc)
        B200_CUDA(cudaMemcpyAsync(desc, M.desc.p, sizeof(int4) * (size_t)M.n_tiles, cudaMemcpyDeviceToHost, st));
    if(off)
        B200_CUDA(cudaMemcpyAsync(off, M.off.p, sizeof(long long) * 2 * (size_t)M.n_tiles, cudaMemcpyDeviceToHost, st));
    if(walk)
        B200_CUDA(cudaMemcpyAsync(walk, M.walk.p, 4 * (size_t)M.walk_entries, cudaMemcpyDeviceToHost, st));
    if(val)
        B200_CUDA(cudaMemcpyAsync(val, M.val.p, es * (size_t)M.val_entries, cudaMemcpyDeviceToHost, st));
    if(rows)
        B200_CUDA(cudaMemcpyAsync(rows, M.rows.p, sizeof(int) * tr, cudaMemcpyDeviceToHost, st));
    if(runs)
        B200_CUDA(cudaMemcpyAsync(runs, M.runs.p, sizeof(int2) * (size_t)M.n_runs_total, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return aoclsparse_status_success;
}
}
