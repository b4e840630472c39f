"""numpy twins of the synthetic-input generators; small sizes only.

The device generators in aocl-sparse_b200/csrc/gen.cu must produce the same arrays bit for bit
(tests/test_parity_gpu.py::test_generators_match_numpy).
"""
import numpy as np

M64 = (1 << 64) - 1


def splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def u01(seed, idx):
    """(splitmix64(seed*0x100000001B3 ^ i) >> 11) * 2^-53 for an array of indices"""
    with np.errstate(over="ignore"):
        idx = np.asarray(idx, dtype=np.uint64)
        s = np.uint64((seed * 0x100000001B3) & M64)
        return (splitmix64(s ^ idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform(seed, first, count, dtype=np.float64):
    """u(seed, i) in (-1, 1), i in [first, first+count)"""
    return (2.0 * u01(seed, np.arange(first, first + count, dtype=np.uint64)) - 1.0).astype(dtype)


def stencil(points, nx, ny, nz=1, row_lo=0, row_hi=None, dtype=np.float64):
    """5-point (2-D), 7-point or 27-point (3-D) stencil rows [row_lo,row_hi): diag = points-1, off = -1,
    columns ascending, base 0.  Returns (row_ptr, col, val) with row_ptr starting at 0.
    Offsets are visited in ascending column order, so an entry's slot in its row is the number of
    earlier offsets that fall inside the grid: no sort is needed."""
    total = nx * ny * nz
    if row_hi is None:
        row_hi = total
    r = np.arange(row_lo, row_hi, dtype=np.int64)
    ix, iy, iz = r % nx, (r // nx) % ny, r // (nx * ny)
    offs = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                off = (dx != 0) + (dy != 0) + (dz != 0)
                if (points != 27 and off > 1) or (nz == 1 and dz != 0):
                    continue
                offs.append((dx, dy, dz, off))
    oks = []
    cnt = np.zeros(len(r), dtype=np.int64)
    for (dx, dy, dz, off) in offs:
        x, y, z = ix + dx, iy + dy, iz + dz
        ok = (x >= 0) & (x < nx) & (y >= 0) & (y < ny) & (z >= 0) & (z < nz)
        oks.append((ok, cnt.copy()))
        cnt += ok
    rp = np.zeros(len(r) + 1, dtype=np.int64)
    np.cumsum(cnt, out=rp[1:])
    col = np.empty(rp[-1], dtype=np.int32)
    val = np.empty(rp[-1], dtype=dtype)
    for (dx, dy, dz, off), (ok, before) in zip(offs, oks):
        pos = (rp[:-1] + before)[ok]
        col[pos] = (((iz + dz) * ny + (iy + dy)) * nx + (ix + dx))[ok]
        val[pos] = float(points - 1) if off == 0 else -1.0
    return rp.astype(np.int32), col, val


def rmat_keys(seed, scale, first, count):
    """Graph500 R-MAT edges (a,b,c,d)=(0.57,0.19,0.19,0.05): key = row<<32 | col"""
    e = np.arange(first, first + count, dtype=np.uint64)
    row = np.zeros(count, dtype=np.uint64)
    col = np.zeros(count, dtype=np.uint64)
    for l in range(scale):
        u = u01(seed, e * np.uint64(64) + np.uint64(l))
        q = np.where(u < 0.57, 0, np.where(u < 0.76, 1, np.where(u < 0.95, 2, 3))).astype(np.uint64)
        row = (row << np.uint64(1)) | (q >> np.uint64(1))
        col = (col << np.uint64(1)) | (q & np.uint64(1))
    return ((row << np.uint64(32)) | col).astype(np.int64)


def rmat_csr(scale, edgefactor=16, seed=20240, val_seed=4, dtype=np.float32):
    """sorted, de-duplicated R-MAT CSR with a_ij = u(val_seed, i*2^scale + j)"""
    n = 1 << scale
    keys = np.unique(rmat_keys(seed, scale, 0, edgefactor * n))
    row = (keys >> 32).astype(np.int64)
    col = (keys & 0xFFFFFFFF).astype(np.int64)
    rp = np.searchsorted(row, np.arange(n + 1), side="left").astype(np.int32)
    val = (2.0 * u01(val_seed, (row.astype(np.uint64) << np.uint64(scale)) + col.astype(np.uint64)) - 1.0).astype(dtype)
    return rp, col.astype(np.int32), val


def random_csr(rng, m, n, density=0.2, dtype=np.float64, sort="full", ensure_diag=False, base=0,
               empty_rows=0.1):
    """random CSR in the spirit of the reference's functional-test matrices (values ~ Normal(-1, 1),
    tests/include/aoclsparse_random.hpp:96-131); sort in {full, partial, none}."""
    rp = [0]
    cols, vals = [], []
    cplx = np.issubdtype(dtype, np.complexfloating)
    for i in range(m):
        if n == 0 or rng.random() < empty_rows:
            c = np.zeros(0, dtype=np.int64)
        else:
            k = rng.binomial(n, density)
            c = np.sort(rng.choice(n, size=k, replace=False))
        if ensure_diag and i < n and i not in c:
            c = np.sort(np.append(c, i))
        if sort == "partial" and len(c) > 1:
            lo, hi, d = c[c < i], c[c > i], c[c == i]
            rng.shuffle(lo)
            rng.shuffle(hi)
            c = np.concatenate([lo, d, hi])
        elif sort == "none" and len(c) > 1:
            rng.shuffle(c)
        v = rng.normal(-1.0, 1.0, size=len(c))
        if cplx:
            v = v + 1j * rng.normal(-1.0, 1.0, size=len(c))
        cols.append(c)
        vals.append(v)
        rp.append(rp[-1] + len(c))
    col = (np.concatenate(cols) if cols else np.zeros(0)).astype(np.int32) + base
    val = (np.concatenate(vals) if vals else np.zeros(0)).astype(dtype)
    return (np.asarray(rp, dtype=np.int32) + base), col, val
