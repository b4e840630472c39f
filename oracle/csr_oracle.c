/* csr_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A scalar, single-threaded CPU restatement, in plain C, of what the reference (AOCL-Sparse v5.3.2,
 * /root/reference) computes on the CSR SpMV / SpMM path.  It exists so that the parity tests have a
 * checker that travels to the GPU box, where /root/reference does not exist.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * library (aocl-sparse_b200/) never does.
 *
 * PARITY IS PINNED: tests/test_oracle.py checks every function below against
 *   (a) the reference's own golden vectors lifted into tests/golden/ (mv_tests.cpp, csrmv_tests.cpp,
 *       csrmm_tests.cpp, createcsr_tests.cpp, doid_score_tests.cpp, sample_*.c), and
 *   (b) outputs of the reference itself, compiled from its own sources into
 *       oracle/_ref/libaoclsparse_ref.so (oracle/Makefile), on seeded random inputs.
 *
 * Each function cites the reference file:line it follows.
 */
#include <complex.h>
#include <stdlib.h>
#include <string.h>

/* ---- status / enum values: library/include/aoclsparse_types.h:304-324,396-403 ---- */
enum
{
    ST_SUCCESS = 0,
    ST_INVALID_POINTER = 2,
    ST_INVALID_SIZE = 3,
    ST_INVALID_VALUE = 5,
    ST_INVALID_INDEX_VALUE = 6
};
enum
{
    SORT_FULL = 1,
    SORT_PARTIAL = 2,
    SORT_NONE = 3
};

/* CSR validation + classification.
 * Follows aoclsparse_mat_check_internal, library/src/analysis/aoclsparse_csr_util.cpp:124-279
 * (shape_general, fast_chck == false): same tests in the same order, so the FIRST failing entry in
 * storage order decides the status.  val may be any non-NULL pointer (values are never read). */
int oracle_mat_check(int m, int n, int nnz, const int *rp, const int *col, const void *val, int base,
                     int *sort_out, int *fulldiag_out)
{
    if(!rp || !col || !val)
        return ST_INVALID_POINTER;
    if(m < 0 || n < 0 || nnz < 0)
        return ST_INVALID_SIZE;
    if(rp[0] - base != 0)
        return ST_INVALID_VALUE;
    if(rp[m] - base != nnz)
        return ST_INVALID_VALUE;
    for(int i = 1; i <= m; ++i)
        if(rp[i - 1] > rp[i])
            return ST_INVALID_VALUE;
    int sort = SORT_FULL, fulldiag = 1;
    for(int i = 0; i < m; ++i)
    {
        int seen_diag = 0, seen_upper = 0, prev = -1;
        for(int p = rp[i] - base; p < rp[i + 1] - base; ++p)
        {
            const int j = col[p] - base;
            if(j < 0 || j > n - 1)
                return ST_INVALID_INDEX_VALUE;
            if(sort != SORT_NONE)
            {
                if(prev > j)
                    sort = SORT_PARTIAL;
                else
                    prev = j;
                if((j <= i && seen_upper) || (j < i && seen_diag))
                    sort = SORT_NONE;
            }
            if(j > i)
                seen_upper = 1;
            else if(j == i)
            {
                if(seen_diag)
                    return ST_INVALID_VALUE;
                seen_diag = 1;
            }
        }
        if(!seen_diag && i < n)
            fulldiag = 0;
    }
    *sort_out     = sort;
    *fulldiag_out = fulldiag;
    return ST_SUCCESS;
}

/* Dispatch id of (descriptor type, fill mode, operation) for a real or complex value type.
 * Follows aoclsparse::get_doid<T>, library/src/include/aoclsparse_mtx_dispatcher.hpp:79-143
 * (numbering :41-74).  20 = invalid. */
int oracle_doid(int is_complex, int type, int fill, int op)
{
    int opv = op - 111;
    if(opv < 0 || opv > 2)
        return 20;
    if(!is_complex)
    {
        if(opv == 2)
            opv = 1;
        if(type == 2)
            type = 1;
    }
    if(type == 1 && opv == 1)
        opv = 0;
    else if(type == 2 && opv == 2)
        opv = 0;
    static const int bits[3] = {0, 2, 3};
    switch(type)
    {
    case 0:
        return bits[opv];
    case 1:
        return 4 + 2 * fill + (opv >> 1);
    case 2:
        return 8 + 2 * fill + (opv ^ fill);
    case 3:
        return 12 + 4 * fill + bits[opv];
    }
    return 20;
}

/* "Clean CSR" of a double matrix: rows grouped lower | diagonal | upper, missing diagonals of rows i < n inserted as
 * explicit zeros, idiag / iurow positions.  Follows aoclsparse_csr_csc_optimize<T>
 * (library/src/analysis/aoclsparse_csr_util.hpp:765-967): if the input is group-ordered
 * (aoclsparse_csr_csc_check_sort_diag, aoclsparse_csr_util.cpp:290-364) with a full diagonal it IS the clean matrix
 * (is_internal = 0, the caller's base kept); otherwise a base-0 copy is made, rows are sorted by column when not
 * group-ordered (aoclsparse_sort_idx_val, csr_util.hpp:99-160) and diagonals are filled in front of the first
 * upper entry (aoclsparse_csr_csc_fill_diag, csr_util.hpp:166-279); indices from aoclsparse_csr_csc_indices
 * (aoclsparse_csr_util.cpp:389-458).  Output arrays must hold nnz + min(m,n) entries.  Returns the clean nnz. */
int oracle_clean_csr(int m, int n, int base, const int *rp, const int *col, const double *val, int *is_internal,
                     int *orp, int *ocol, double *oval, int *idiag, int *iurow)
{
    int grouped = 1, fulldiag = 1;
    for(int i = 0; i < m && grouped; ++i)
    {
        int lower = 1, found = 0;
        for(int p = rp[i] - base; p < rp[i + 1] - base; ++p)
        {
            const int j = col[p] - base;
            if(j == i)
            {
                found   = 1;
                grouped = grouped && lower;
                lower   = 0;
            }
            else if(lower)
                lower = j < i;
            else
                grouped = grouped && (j > i);
        }
        if(!found && i < n)
            fulldiag = 0;
    }
    if(!grouped)
    {
        fulldiag = 1;
        for(int i = 0; i < m; ++i)
        {
            int found = 0;
            for(int p = rp[i] - base; p < rp[i + 1] - base; ++p)
                found = found || (col[p] - base == i);
            if(!found && i < n)
                fulldiag = 0;
        }
    }
    const int nnz = rp[m] - base;
    int       ob  = 0; /* base of the output */
    if(grouped && fulldiag)
    {
        *is_internal = 0;
        ob           = base;
        for(int i = 0; i <= m; ++i)
            orp[i] = rp[i];
        for(int p = 0; p < nnz; ++p)
        {
            ocol[p] = col[p];
            oval[p] = val[p];
        }
    }
    else
    {
        *is_internal = 1;
        /* copy to base 0, insertion-sort rows by column when not group-ordered */
        int    *c0 = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
        double *v0 = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
        for(int p = 0; p < nnz; ++p)
        {
            c0[p] = col[p] - base;
            v0[p] = val[p];
        }
        if(!grouped)
            for(int i = 0; i < m; ++i)
                for(int p = rp[i] - base + 1; p < rp[i + 1] - base; ++p)
                {
                    int    cc = c0[p], q = p - 1;
                    double vv = v0[p];
                    while(q >= rp[i] - base && c0[q] > cc)
                    {
                        c0[q + 1] = c0[q];
                        v0[q + 1] = v0[q];
                        --q;
                    }
                    c0[q + 1] = cc;
                    v0[q + 1] = vv;
                }
        int q = 0;
        for(int i = 0; i < m; ++i)
        {
            orp[i]   = q;
            int done = !(i < n), has = 0;
            for(int p = rp[i] - base; p < rp[i + 1] - base; ++p)
                has = has || (c0[p] == i);
            done = done || has;
            for(int p = rp[i] - base; p < rp[i + 1] - base; ++p)
            {
                if(!done && c0[p] > i)
                {
                    ocol[q]   = i;
                    oval[q++] = 0.0;
                    done      = 1;
                }
                ocol[q]   = c0[p];
                oval[q++] = v0[p];
            }
            if(!done)
            {
                ocol[q]   = i;
                oval[q++] = 0.0;
            }
        }
        orp[m] = q;
        free(c0);
        free(v0);
    }
    for(int i = 0; i < m; ++i)
    {
        int p = orp[i] - ob;
        const int e = orp[i + 1] - ob;
        while(p < e && ocol[p] - ob < i)
            ++p;
        idiag[i] = p + ob;
        iurow[i] = ((p < e && ocol[p] - ob == i) ? p + 1 : p) + ob;
    }
    return orp[m] - ob;
}

/* Row-block plan: restates the SPEC at the top of aocl-sparse_b200/csrc/plan.cu (this integer
 * metadata has no reference counterpart; the GPU analysis must reproduce it bit for bit).
 * rp is 0-based.  desc (4 ints per block) and kind may be NULL to only count.  Returns n_blocks. */
int oracle_plan(int m, const int *rp, int T, int R, int forced, int n_cuts, const int *cuts,
                int capacity, int *desc, int *kind, int *n_long_rows, int *n_long_segments)
{
    const long long S   = 64LL * T;
    const int       nnz = m > 0 ? rp[m] : 0;
    int             nb = 0, nlr = 0, nls = 0;
    if(m == 0)
    {
        *n_long_rows = *n_long_segments = 0;
        return 0;
    }
    /* segment boundaries */
    long long ngrid = ((long long)nnz + S - 1) / S - 1;
    if(ngrid < 0)
        ngrid = 0;
    int  nbounds = 0;
    int *bounds  = (int *)malloc(sizeof(int) * (size_t)(ngrid + n_cuts + 2));
    bounds[nbounds++] = 0;
    for(long long kq = 0; kq < ngrid; ++kq)
    {
        const long long target = (kq + 1) * S;
        int             lo = 0, hi = m;
        while(lo < hi)
        {
            int mid = lo + (hi - lo) / 2;
            if((long long)rp[mid] >= target)
                hi = mid;
            else
                lo = mid + 1;
        }
        bounds[nbounds++] = lo;
    }
    for(int c = 0; c < n_cuts; ++c)
        bounds[nbounds++] = cuts[c];
    bounds[nbounds++] = m;
    /* sort + unique (insertion sort: the list is short and nearly sorted) */
    for(int i = 1; i < nbounds; ++i)
    {
        int v = bounds[i], j = i - 1;
        while(j >= 0 && bounds[j] > v)
        {
            bounds[j + 1] = bounds[j];
            --j;
        }
        bounds[j + 1] = v;
    }
    int u = 0;
    for(int i = 0; i < nbounds; ++i)
        if(i == 0 || bounds[i] != bounds[u - 1])
            bounds[u++] = bounds[i];
    nbounds = u;

    for(int s = 0; s + 1 < nbounds; ++s)
    {
        const int sb = bounds[s + 1];
        int       r  = bounds[s];
        while(r < sb)
        {
            const int p0 = rp[r], len = rp[r + 1] - rp[r];
            if(len > T)
            {
                const int q = (int)(((long long)len + T - 1) / T);
                for(int g = 0; g < q; ++g)
                {
                    long long a = (long long)p0 + (long long)g * T, b = a + T;
                    if(b > (long long)p0 + len)
                        b = (long long)p0 + len;
                    if(desc && nb + g < capacity)
                    {
                        desc[4 * (nb + g) + 0] = r;
                        desc[4 * (nb + g) + 1] = r + 1;
                        desc[4 * (nb + g) + 2] = (int)a;
                        desc[4 * (nb + g) + 3] = (int)b;
                        kind[nb + g]           = 3 | ((nls + g) << 4);
                    }
                }
                nb += q;
                nls += q;
                nlr += 1;
                r += 1;
            }
            else
            {
                int hi = (sb - r > R) ? r + R : sb, lo = r + 1;
                const long long lim = (long long)p0 + T;
                while(lo < hi)
                {
                    int mid = lo + (hi - lo + 1) / 2;
                    if((long long)rp[mid] <= lim)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                if(desc && nb < capacity)
                {
                    int L = 0;
                    for(int q = r; q < lo; ++q)
                        if(rp[q + 1] - rp[q] > L)
                            L = rp[q + 1] - rp[q];
                    const long long nr = lo - r, nz = rp[lo] - p0;
                    int             kk;
                    if(forced >= 0)
                        kk = forced;
                    else if(L <= 64 && (long long)L * nr <= 2 * nz + nr)
                        kk = 0;
                    else if(nz >= 48 * nr && (long long)L * nr <= 4 * nz)
                        kk = 1;
                    else
                        kk = 2;
                    desc[4 * nb + 0] = r;
                    desc[4 * nb + 1] = lo;
                    desc[4 * nb + 2] = p0;
                    desc[4 * nb + 3] = rp[lo];
                    kind[nb]         = kk;
                }
                nb += 1;
                r = lo;
            }
        }
    }
    free(bounds);
    *n_long_rows     = nlr;
    *n_long_segments = nls;
    return nb;
}

/* plan_parameters + wave-aware block size of plan.cu (same arithmetic, same candidate order) */
static long long ctas_per_wave(int elem_size, int T, int coded)
{
    const long long smem = coded == 2 ? 16 + (long long)((T + 32 + 15) & ~15) + 256 * (elem_size >= 8 ? 16LL : 8LL) + 1024
                           : coded    ? 16 + (long long)(T + 32) * (long long)(elem_size + 1) + 1024 + 1024
                                      : 16 + (long long)(T + 8) * (long long)(elem_size + 4) + 1024;
    long long       c    = 232448 / smem;
    if(c > 8)
        c = 8;
    if(c < 1)
        c = 1;
    return 148 * c;
}

/* coded == 1: the plan serves the diagonal-code copy (elem_size + 1 staged bytes per entry)
 * coded == 2: the block plan of the entry-coded kernels (1 staged byte per entry): blocks end at a row count -- as many
 *             rows as still leave 8 waves of CTAs, within [512, 2048]; matrices of 1..8 waves: the row count 512 - 8k
 *             (k = 0..24) minimising ceil(1.015 blocks / CTAs per wave) * rows; less than one wave: one block per
 *             resident CTA, at least 64 rows */
void oracle_plan_parameters(int elem_size, int m, int nnz, int max_row_nnz, const int *rp, int n_cuts, const int *cuts,
                            int coded, int *T, int *R)
{
    if(coded == 2)
    {
        const int       t     = (24576 - 4096 - 32) / 256 * 256;
        const long long slots = ctas_per_wave(elem_size, t, 2);
        long long       r     = (long long)m / (8 * slots) / 64 * 64;
        r                     = r < 512 ? 512 : (r > 2048 ? 2048 : r);
        if((long long)m < slots * 512) /* less than one wave: one block per resident CTA, at least 64 rows */
        {
            r = (((long long)m + slots - 1) / slots + 31) / 32 * 32;
            r = r < 64 ? 64 : r;
        }
        *T                    = t;
        *R                    = (int)r;
        if(m > 0 && rp && r >= 512 && (long long)m < 8 * slots * r && (long long)m >= slots * r)
        {
            long long best_cost = -1;
            int       best_r    = (int)r;
            for(int k = 0; k <= 24; ++k)
            {
                const int       rk = (int)r - 8 * k;
                int             a, b;
                const long long nb   = oracle_plan(m, rp, t, rk, -1, n_cuts, cuts, 0, 0, 0, &a, &b);
                const long long cost = (((nb * 203 + 199) / 200 + slots - 1) / slots) * (long long)rk;
                if(best_cost < 0 || cost < best_cost)
                {
                    best_cost = cost;
                    best_r    = rk;
                }
            }
            *R = best_r;
        }
        return;
    }
    int             t    = (24576 / (elem_size + 4)) / 512 * 512;
    if(coded)
        t = (24576 / (elem_size + 1) - 32) / 256 * 256;
    if(elem_size >= 16)
        t = 1536; /* 16-byte values: measured best, with 128-thread CTAs */
    const long long mean = m > 0 ? (long long)nnz / m : 0;
    if((long long)max_row_nnz > 16 * (mean > 1 ? mean : 1))
        t = ((7152 / (elem_size + 4)) - 8) / 128 * 128; /* skewed rows: 8 CTAs inside the 64 KB carve-out, L1 kept large */
    if(t < 512)
        t = 512;
    while(t >= 1024 && (long long)nnz < (long long)t * 148 * 8)
        t -= 512;
    *R = 1024;
    if(m > 0 && rp && (long long)nnz < 8 * ctas_per_wave(elem_size, t, coded) * (long long)t
       && (long long)nnz >= ctas_per_wave(elem_size, t, coded) * (long long)t)
    {
        long long best_cost = -1;
        int       best_t    = t;
        for(int k = 0; k <= 16; ++k)
        {
            const int       tk = t + 32 * k;
            int             a, b;
            const long long nb   = oracle_plan(m, rp, tk, *R, -1, n_cuts, cuts, 0, 0, 0, &a, &b);
            const long long wave = ctas_per_wave(elem_size, tk, coded);
            const long long cost = (((nb * 203 + 199) / 200 + wave - 1) / wave) * (long long)tk;
            if(best_cost < 0 || cost < best_cost)
            {
                best_cost = cost;
                best_t    = tk;
            }
        }
        t = best_t;
    }
    *T = t;
}

/* ---- multiply kernels, instantiated for the four value types ---- */
/* Diagonal-code copy built by aoclsparse_optimize (aocl-sparse_b200/csrc/plan.cu, build_diag_codes): GPU-only
 * integer metadata with no counterpart in the reference, so the written SPEC is restated here:
 *   D = ascending distinct values of col[p] - r over all stored entries (0-based arrays); if 1 <= |D| <= 256,
 *   codes[p] = index of (col[p] - r) in D; otherwise not applicable (returns 0).
 * Returns |D| (0: not applicable); offsets[256] and codes[nnz] are written when it is not 0. */
static int cmp_int(const void *a, const void *b)
{
    const int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}
int oracle_diag_codes(int m, const int *rp, const int *col, int *offsets, unsigned char *codes)
{
    int d[257];
    int nd = 0;
    for(int r = 0; r < m; ++r)
        for(int p = rp[r]; p < rp[r + 1]; ++p)
        {
            const int off = col[p] - r;
            int       k   = 0;
            while(k < nd && d[k] != off)
                ++k;
            if(k == nd)
            {
                if(nd == 256)
                    return 0;
                d[nd++] = off;
            }
        }
    if(nd == 0)
        return 0;
    qsort(d, (size_t)nd, sizeof(int), cmp_int);
    for(int i = 0; i < nd; ++i)
        offsets[i] = d[i];
    for(int r = 0; r < m; ++r)
        for(int p = rp[r]; p < rp[r + 1]; ++p)
        {
            const int off = col[p] - r;
            int       k   = 0;
            while(d[k] != off)
                ++k;
            codes[p] = (unsigned char)k;
        }
    return nd;
}

/* Entry-code copy built by aoclsparse_optimize next to the diagonal-code copy (aocl-sparse_b200/csrc/plan.cu,
 * build_entry_codes): GPU-only integer metadata with no counterpart in the reference, so the written SPEC is restated here:
 *   D = the diagonal-code table (above); V = ascending distinct bit patterns of the stored values (passed zero-extended to
 *   64 bits); Q = ascending distinct pairs (index in D, index in V) over all stored entries, ordered by the first, then
 *   the second component.  Applicable iff 1 <= |D|, |V|, |Q| <= 256 and no value has the all-ones 64-bit pattern:
 *   ecodes[p] = index of p's pair in Q, pair_off[i] = D[Q[i].first], pair_val[i] = V[Q[i].second].
 * Returns |Q| (0: not applicable). */
static int cmp_u64(const void *a, const void *b)
{
    const unsigned long long x = *(const unsigned long long *)a, y = *(const unsigned long long *)b;
    return (x > y) - (x < y);
}
int oracle_entry_codes(int m, const int *rp, const int *col, const unsigned long long *val_bits, int *pair_off,
                       unsigned long long *pair_val, unsigned char *ecodes)
{
    const int nnz = rp[m];
    if(nnz <= 0)
        return 0;
    int            d[256];
    unsigned char *dc = (unsigned char *)malloc((size_t)nnz);
    const int      nd = oracle_diag_codes(m, rp, col, d, dc);
    if(nd == 0)
    {
        free(dc);
        return 0;
    }
    unsigned long long v[257];
    int                nv = 0;
    for(int p = 0; p < nnz; ++p)
    {
        if(val_bits[p] == 0xffffffffffffffffull)
        {
            free(dc);
            return 0;
        }
        int k = 0;
        while(k < nv && v[k] != val_bits[p])
            ++k;
        if(k == nv)
        {
            if(nv == 256)
            {
                free(dc);
                return 0;
            }
            v[nv++] = val_bits[p];
        }
    }
    qsort(v, (size_t)nv, sizeof(v[0]), cmp_u64);
    unsigned char *seen = (unsigned char *)calloc(65536, 1);
    int           *pid  = (int *)malloc(sizeof(int) * (size_t)nnz);
    for(int p = 0; p < nnz; ++p)
    {
        int k = 0;
        while(v[k] != val_bits[p])
            ++k;
        pid[p]       = ((int)dc[p] << 8) | k;
        seen[pid[p]] = 1;
    }
    int rank[65536];
    int nq = 0;
    for(int q = 0; q < 65536; ++q)
        if(seen[q])
            rank[q] = nq++;
    if(nq <= 256)
    {
        for(int q = 0; q < 65536; ++q)
            if(seen[q])
            {
                pair_off[rank[q]] = d[q >> 8];
                pair_val[rank[q]] = v[q & 255];
            }
        for(int p = 0; p < nnz; ++p)
            ecodes[p] = (unsigned char)rank[pid[p]];
    }
    free(dc);
    free(seen);
    free(pid);
    return nq <= 256 ? nq : 0;
}

/* Row pointers of C = A B for two CSR operands in the orientation of the product: the number of distinct column
 * indices reached from every row of A (aoclsparse_csr2m_nnz_count, library/src/level3/aoclsparse_csr2m.cpp:46-305),
 * then a 64-bit prefix sum.  Returns 0, or 3 (invalid size) when the total does not fit a 32-bit int (:236-241). */
int oracle_csr2m_count(int m, int n, int baseA, const int *rpA, const int *colA, int baseB, const int *rpB,
                       const int *colB, int *rpC)
{
    int *mark = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for(int j = 0; j < n; ++j)
        mark[j] = -1;
    long long run = 0;
    rpC[0]        = 0;
    for(int i = 0; i < m; ++i)
    {
        int cnt = 0;
        for(int p = rpA[i] - baseA; p < rpA[i + 1] - baseA; ++p)
        {
            const int k = colA[p] - baseA;
            for(int q = rpB[k] - baseB; q < rpB[k + 1] - baseB; ++q)
            {
                const int j = colB[q] - baseB;
                if(mark[j] != i)
                {
                    mark[j] = i;
                    ++cnt;
                }
            }
        }
        run += cnt;
        rpC[i + 1] = (int)run;
    }
    free(mark);
    return run > 2147483647LL ? 3 : 0;
}

#define CAT2(a, b) a##_##b
#define CAT(a, b) CAT2(a, b)
#define NAME(f) CAT(f, SUF)

#define T float
#define SUF s
#define CONJ(v) (v)
#define IS_COMPLEX 0
#include "csr_oracle_impl.inc"
#undef T
#undef SUF
#undef CONJ
#undef IS_COMPLEX

#define T double
#define SUF d
#define CONJ(v) (v)
#define IS_COMPLEX 0
#include "csr_oracle_impl.inc"
#undef T
#undef SUF
#undef CONJ
#undef IS_COMPLEX

#define T float _Complex
#define SUF c
#define CONJ(v) conjf(v)
#define IS_COMPLEX 1
#include "csr_oracle_impl.inc"
#undef T
#undef SUF
#undef CONJ
#undef IS_COMPLEX

#define T double _Complex
#define SUF z
#define CONJ(v) conj(v)
#define IS_COMPLEX 1
#include "csr_oracle_impl.inc"
#undef T
#undef SUF
#undef CONJ
#undef IS_COMPLEX
