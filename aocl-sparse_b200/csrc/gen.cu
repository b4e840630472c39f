// gen.cu -- the synthetic inputs of BASELINE.json, generated directly in device memory so that the
// large configurations (3D 7-point 512^3: 938 M entries) never exist on the host.
// Definitions: SURVEY.md section 8(d).  tests/gen_np.py holds the numpy twins used on small sizes.
#include "common.hpp"

namespace b200
{
    namespace
    {
        __host__ __device__ inline unsigned long long splitmix64(unsigned long long z)
        {
            z += 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        }
        __host__ __device__ inline double u01(unsigned long long seed, unsigned long long i)
        {
            return (double)(splitmix64((seed * 0x100000001B3ull) ^ i) >> 11) * (1.0 / 9007199254740992.0);
        }

        struct stencil_geom
        {
            int       points;
            long long nx, ny, nz;
        };

        // number of stencil neighbours (including the centre) of grid point (ix,iy,iz) that lie inside
        __device__ inline int stencil_row_nnz(const stencil_geom &g, long long ix, long long iy, long long iz)
        {
            const int cx = 1 + (ix > 0) + (ix < g.nx - 1);
            const int cy = 1 + (iy > 0) + (iy < g.ny - 1);
            const int cz = 1 + (iz > 0) + (iz < g.nz - 1);
            if(g.points == 27)
                return cx * cy * cz;
            // 5-point (2-D) and 7-point (3-D): axis neighbours only
            return 1 + (cx - 1) + (cy - 1) + (g.nz > 1 ? (cz - 1) : 0);
        }

        __global__ void stencil_count_kernel(stencil_geom g, long long row_lo, long long nrows, int *counts)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < nrows; i += (long long)gridDim.x * blockDim.x)
            {
                const long long r  = row_lo + i;
                const long long ix = r % g.nx, iy = (r / g.nx) % g.ny, iz = r / (g.nx * g.ny);
                counts[i]          = stencil_row_nnz(g, ix, iy, iz);
            }
        }

        // row_ptr[i] = closed-form prefix of the counts: computed by a chunked scan (one thread per chunk)
        __global__ void chunk_sums_kernel(long long nrows, int chunk, const int *counts, long long *sums)
        {
            long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            long long nchunks = (nrows + chunk - 1) / chunk;
            if(c >= nchunks)
                return;
            long long s = 0;
            for(long long i = c * chunk; i < nrows && i < (c + 1) * (long long)chunk; ++i)
                s += counts[i];
            sums[c] = s;
        }
        __global__ void chunk_scan_kernel(long long nchunks, long long *sums)
        {
            // single thread: nchunks is small (rows / 4096)
            if(blockIdx.x || threadIdx.x)
                return;
            long long run = 0;
            for(long long c = 0; c < nchunks; ++c)
            {
                long long v = sums[c];
                sums[c]     = run;
                run += v;
            }
            sums[nchunks] = run;
        }
        __global__ void chunk_fill_kernel(long long nrows, int chunk, const int *counts, const long long *sums, int *row_ptr)
        {
            long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            long long nchunks = (nrows + chunk - 1) / chunk;
            if(c >= nchunks)
                return;
            long long s = sums[c];
            for(long long i = c * chunk; i < nrows && i < (c + 1) * (long long)chunk; ++i)
            {
                row_ptr[i] = (int)s;
                s += counts[i];
            }
            if(c == nchunks - 1)
                row_ptr[nrows] = (int)s;
        }

        __global__ void stencil_fill_kernel(stencil_geom g, long long row_lo, long long nrows, const int *row_ptr, int *col, double *val)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < nrows; i += (long long)gridDim.x * blockDim.x)
            {
                const long long r  = row_lo + i;
                const long long ix = r % g.nx, iy = (r / g.nx) % g.ny, iz = r / (g.nx * g.ny);
                int             p  = row_ptr[i];
                const double    dg = (double)(g.points - 1);
                for(int dz = -1; dz <= 1; ++dz)
                {
                    const long long z = iz + dz;
                    if(z < 0 || z >= g.nz)
                        continue;
                    for(int dy = -1; dy <= 1; ++dy)
                    {
                        const long long y = iy + dy;
                        if(y < 0 || y >= g.ny)
                            continue;
                        for(int dx = -1; dx <= 1; ++dx)
                        {
                            const long long x = ix + dx;
                            if(x < 0 || x >= g.nx)
                                continue;
                            const int nz_off = (dx != 0) + (dy != 0) + (dz != 0);
                            if(g.points != 27 && nz_off > 1)
                                continue;
                            col[p] = (int)((z * g.ny + y) * g.nx + x);
                            val[p] = (nz_off == 0) ? dg : -1.0;
                            ++p;
                        }
                    }
                }
            }
        }

        template <typename T>
        __global__ void uniform_kernel(unsigned long long seed, long long first, long long count, T *out)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < count; i += (long long)gridDim.x * blockDim.x)
                out[i] = (T)(2.0 * u01(seed, (unsigned long long)(first + i)) - 1.0);
        }

        // Graph500-style R-MAT edge (a,b,c,d) = (0.57,0.19,0.19,0.05): key = row << 32 | col
        __global__ void rmat_kernel(unsigned long long seed, int scale, long long first, long long count, long long *keys)
        {
            long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            for(; i < count; i += (long long)gridDim.x * blockDim.x)
            {
                const unsigned long long e = (unsigned long long)(first + i);
                unsigned long long row = 0, col = 0;
                for(int l = 0; l < scale; ++l)
                {
                    const double u = u01(seed, e * 64ull + (unsigned long long)l);
                    const int    q = (u < 0.57) ? 0 : ((u < 0.76) ? 1 : ((u < 0.95) ? 2 : 3));
                    row            = (row << 1) | (unsigned long long)(q >> 1);
                    col            = (col << 1) | (unsigned long long)(q & 1);
                }
                keys[i] = (long long)((row << 32) | col);
            }
        }

        __global__ void rmat_csr_kernel(unsigned long long seed, int scale, long long count, const long long *__restrict__ keys, int *row_ptr, int *col, float *val)
        {
            const long long nrows = 1LL << scale;
            long long       i     = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            const long long stride = (long long)gridDim.x * blockDim.x;
            for(long long p = i; p < count; p += stride)
            {
                const unsigned long long k = (unsigned long long)keys[p];
                const unsigned long long r = k >> 32, c = k & 0xffffffffull;
                col[p] = (int)c;
                val[p] = (float)(2.0 * u01(seed, (r << scale) + c) - 1.0);
            }
            for(long long r = i; r <= nrows; r += stride)
            {
                // number of keys with row < r
                const unsigned long long target = (unsigned long long)r << 32;
                long long lo = 0, hi = count;
                while(lo < hi)
                {
                    long long mid = lo + (hi - lo) / 2;
                    if((unsigned long long)keys[mid] < target)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                row_ptr[r] = (int)lo;
            }
        }

        inline unsigned grid_for(long long n, int tpb)
        {
            long long b = (n + tpb - 1) / tpb;
            if(b > 148LL * 32)
                b = 148LL * 32;
            return (unsigned)(b < 1 ? 1 : b);
        }
    }
}

using namespace b200;

extern "C" {

aoclsparse_status aoclsparse_b200_gen_stencil(int             points,
                                              aoclsparse_int  nx,
                                              aoclsparse_int  ny,
                                              aoclsparse_int  nz,
                                              long long       row_lo,
                                              long long       row_hi,
                                              long long      *nnz,
                                              aoclsparse_int *row_ptr,
                                              aoclsparse_int *col_idx,
                                              double         *val)
{
    if(!nnz)
        return aoclsparse_status_invalid_pointer;
    if(!((points == 5 && nz == 1) || ((points == 7 || points == 27) && nz >= 1)) || nx < 1 || ny < 1)
        return aoclsparse_status_invalid_value;
    const long long total = (long long)nx * ny * nz;
    if(row_lo < 0 || row_hi > total || row_lo > row_hi)
        return aoclsparse_status_invalid_size;
    const long long nrows = row_hi - row_lo;
    cudaStream_t    st    = current_stream();
    stencil_geom    g{points, nx, ny, nz};
    if(nrows == 0)
    {
        *nnz = 0;
        if(row_ptr)
            B200_CUDA(cudaMemsetAsync(row_ptr, 0, sizeof(aoclsparse_int), st));
        return aoclsparse_status_success;
    }
    const int       chunk   = 4096;
    const long long nchunks = (nrows + chunk - 1) / chunk;
    dev_buf         counts, sums;
    B200_TRY(counts.alloc(sizeof(int) * (size_t)nrows));
    B200_TRY(sums.alloc(sizeof(long long) * (size_t)(nchunks + 1)));
    stencil_count_kernel<<<grid_for(nrows, 256), 256, 0, st>>>(g, row_lo, nrows, counts.as<int>());
    B200_LAUNCHED();
    chunk_sums_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(nrows, chunk, counts.as<int>(), sums.as<long long>());
    B200_LAUNCHED();
    chunk_scan_kernel<<<1, 1, 0, st>>>(nchunks, sums.as<long long>());
    B200_LAUNCHED();
    long long total_nnz = 0;
    B200_CUDA(cudaMemcpyAsync(&total_nnz, sums.as<long long>() + nchunks, sizeof(long long), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    *nnz = total_nnz;
    if(!row_ptr)
        return aoclsparse_status_success;
    if(!col_idx || !val)
        return aoclsparse_status_invalid_pointer;
    if(total_nnz > 0x7fffffffLL)
        return aoclsparse_status_invalid_size;
    chunk_fill_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(
        nrows, chunk, counts.as<int>(), sums.as<long long>(), row_ptr);
    B200_LAUNCHED();
    stencil_fill_kernel<<<grid_for(nrows, 128), 128, 0, st>>>(g, row_lo, nrows, row_ptr, col_idx, val);
    B200_LAUNCHED();
    B200_CUDA(cudaStreamSynchronize(st));
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_gen_uniform(unsigned long long seed, long long first, long long count, int elem_size, void *out)
{
    if(!out)
        return aoclsparse_status_invalid_pointer;
    if(count < 0 || (elem_size != 4 && elem_size != 8))
        return aoclsparse_status_invalid_value;
    if(count == 0)
        return aoclsparse_status_success;
    cudaStream_t st = current_stream();
    if(elem_size == 8)
        uniform_kernel<double><<<grid_for(count, 256), 256, 0, st>>>(seed, first, count, (double *)out);
    else
        uniform_kernel<float><<<grid_for(count, 256), 256, 0, st>>>(seed, first, count, (float *)out);
    B200_LAUNCHED();
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_gen_rmat_keys(unsigned long long seed, int scale, long long first, long long count, long long *keys)
{
    if(!keys)
        return aoclsparse_status_invalid_pointer;
    if(scale < 1 || scale > 31 || count < 0)
        return aoclsparse_status_invalid_value;
    if(count == 0)
        return aoclsparse_status_success;
    cudaStream_t st = current_stream();
    rmat_kernel<<<grid_for(count, 256), 256, 0, st>>>(seed, scale, first, count, keys);
    B200_LAUNCHED();
    return aoclsparse_status_success;
}

aoclsparse_status aoclsparse_b200_rmat_keys_to_csr(unsigned long long seed, int scale, long long count, const long long *keys, aoclsparse_int *row_ptr, aoclsparse_int *col_idx, float *val)
{
    if(!keys || !row_ptr || !col_idx || !val)
        return aoclsparse_status_invalid_pointer;
    if(scale < 1 || scale > 30 || count < 0 || count > 0x7fffffffLL)
        return aoclsparse_status_invalid_value;
    cudaStream_t    st = current_stream();
    const long long w  = count > (1LL << scale) ? count : (1LL << scale) + 1;
    rmat_csr_kernel<<<grid_for(w, 256), 256, 0, st>>>(seed, scale, count, keys, row_ptr, col_idx, val);
    B200_LAUNCHED();
    return aoclsparse_status_success;
}
}
