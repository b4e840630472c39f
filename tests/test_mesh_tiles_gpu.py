"""GPU tests of csrmm's box-tile path for grid (stencil) matrices (aocl-sparse_b200/csrc/mesh_tiles.cu):
results against the plain-C oracle within the north_star tolerance, and BIT-identical to the row-block kernel (the two
kernels form every output entry with the same sequence of multiply-adds)."""
import os

import numpy as np
import pytest

import gen_np
from conftest import TOL

pytestmark = pytest.mark.gpu

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def _values(rng, val, dt):
    v = (val * rng.uniform(0.5, 1.5, size=len(val))).astype(dt)
    if np.issubdtype(dt, np.complexfloating):
        v = (v + 1j * rng.normal(size=len(val))).astype(dt)
    return v


def _run(lib, p, rp, col, val, m, ncols, B, n, ldb, C0, ldc, alpha, beta, mode):
    import torch
    os.environ["AOCLSPARSE_B200_MM_TILES"] = str(mode)
    try:
        st, h = lib.create_csr(p, 0, m, ncols, len(col), rp, col, val)
        assert st == 0, (st, lib.last_error())
        d = lib.create_descr()
        dB, dC = torch.from_numpy(B).cuda(), torch.from_numpy(C0).cuda()
        for _ in range(2):  # the second call runs on the cached tiles
            dC.copy_(torch.from_numpy(C0))
            assert lib.csrmm(p, 111, alpha, h, d, 0, dB.data_ptr(), n, ldb, beta, dC.data_ptr(), ldc) == 0, lib.last_error()
        torch.cuda.synchronize()
        info = lib.mm_tiles_info(h)
        lib.destroy(h)
        lib.destroy_descr(d)
        return dC.cpu().numpy(), info
    finally:
        os.environ.pop("AOCLSPARSE_B200_MM_TILES", None)


GRIDS = [
    ("27pt 3-D", 27, 20, 12, 10, 3, (8, None, None)),
    ("7pt 3-D", 7, 16, 16, 9, 3, None),
    ("27pt flat", 27, 40, 8, 6, 3, None),
    ("5pt 2-D", 5, 33, 21, 1, 2, None),
]


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
@pytest.mark.parametrize("grid", GRIDS, ids=[g[0] for g in GRIDS])
def test_tiles_match_oracle_and_row_block_kernel(lib, oracle, p, grid):
    import scipy.sparse as sp
    _, pts, nx, ny, nz, ndim, _ = grid
    rng = np.random.default_rng(11)
    dt = DT[p]
    rp, col, val = gen_np.stencil(pts, nx, ny, nz)
    val = _values(rng, val, dt)
    m = len(rp) - 1
    Aabs = sp.csr_matrix((np.abs(val), col, rp), shape=(m, m))
    elem = np.dtype(dt).itemsize
    used = 0
    for row_bytes, pad in ((128, 0), (256, 0), (512, 0), (256, 2 * (16 // elem) if elem < 16 else 2)):
        n = row_bytes // elem
        ldb = ldc = n + pad
        B = rng.normal(size=m * ldb).astype(dt)
        C0 = rng.normal(size=m * ldc).astype(dt)
        if p in "cz":
            B = (B + 1j * rng.normal(size=len(B))).astype(dt)
            C0 = (C0 + 1j * rng.normal(size=len(C0))).astype(dt)
        for alpha, beta in ((1.0, 0.0), (0.5, -1.5)):
            Co = C0.copy()
            assert oracle.csrmm(111, alpha, m, m, 0, rp, col, val, 0, 0, 0, 0, B, n, ldb, beta, Co, ldc) == 0
            tiled, info = _run(lib, p, rp, col, val, m, m, B, n, ldb, C0, ldc, alpha, beta, 2)
            plain, info0 = _run(lib, p, rp, col, val, m, m, B, n, ldb, C0, ldc, alpha, beta, 0)
            assert info0["state"] == 0
            got, want = tiled.reshape(m, ldc), Co.reshape(m, ldc)
            den = abs(alpha) * (Aabs @ np.abs(B.reshape(m, ldb)[:, :n])) + np.abs(beta * C0.reshape(m, ldc)[:, :n]) + 1e-300
            assert np.all(np.abs(got[:, :n] - want[:, :n]) <= TOL[np.dtype(dt)] * den), (p, n, pad)
            assert np.array_equal(got[:, n:], C0.reshape(m, ldc)[:, n:])  # padding untouched
            assert np.array_equal(tiled.view(np.uint8), plain.view(np.uint8)), "tile kernel and row-block kernel differ in bits"
            if info["state"] == 2:
                used += 1
                assert info["rows_per_tile"] % 32 == 0 and info["rows_per_tile"] <= 96
                assert info["strides"][1] == (nx if ndim >= 2 else m)
                if ndim == 3:
                    assert info["strides"][2] == nx * ny
    if pts == 27:
        assert used > 0, "a 27-point grid matrix must take the tile path"


def test_tiles_survive_nan_and_inf_in_unreferenced_rows(lib):
    """B rows no stored entry names must not leak into C (padding entries of the ELL planes are skipped, not multiplied)"""
    rng = np.random.default_rng(5)
    rp, col, val = gen_np.stencil(27, 16, 8, 8)
    m = len(rp) - 1
    # drop every entry of column 77 and make B's row 77 poisonous
    keep = col != 77
    rows = np.repeat(np.arange(m), np.diff(rp))
    rp2 = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows[keep], minlength=m), out=rp2[1:])
    rp2 = rp2.astype(np.int32)
    col2, val2 = col[keep], val[keep]
    n = 32
    B = rng.normal(size=m * n)
    B.reshape(m, n)[77, :] = np.inf
    C0 = np.zeros(m * n)
    out, info = _run(lib, "d", rp2, col2, val2, m, m, B, n, n, C0, n, 1.0, 0.0, 2)
    assert np.all(np.isfinite(out))


def test_non_grid_matrix_keeps_the_row_block_kernel(lib, oracle):
    rng = np.random.default_rng(3)
    m = 4000
    import scipy.sparse as sp
    A = sp.random(m, m, density=0.004, random_state=7, format="csr", dtype=np.float64)
    A.sort_indices()
    rp, col, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    n = 32
    B = rng.normal(size=m * n)
    C0 = np.zeros(m * n)
    out, info = _run(lib, "d", rp, col, val, m, m, B, n, n, C0, n, 1.0, 0.0, 2)
    assert info["state"] == 1  # analysed, not a lattice
    Co = C0.copy()
    assert oracle.csrmm(111, 1.0, m, m, 0, rp, col, val, 0, 0, 0, 0, B, n, n, 0.0, Co, n) == 0
    den = (abs(A) @ np.abs(B.reshape(m, n))) + 1e-300
    assert np.all(np.abs(out.reshape(m, n) - Co.reshape(m, n)) <= 1e-12 * den)


def test_update_values_refreshes_the_tiles(lib):
    import torch
    os.environ["AOCLSPARSE_B200_MM_TILES"] = "2"
    try:
        rng = np.random.default_rng(9)
        rp, col, val = gen_np.stencil(27, 16, 8, 8)
        m = len(rp) - 1
        n = 32
        st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
        assert st == 0
        d = lib.create_descr()
        B = torch.from_numpy(rng.normal(size=m * n)).cuda()
        C1 = torch.zeros(m * n, dtype=torch.float64, device="cuda")
        C2 = torch.zeros_like(C1)
        assert lib.csrmm("d", 111, 1.0, h, d, 0, B.data_ptr(), n, n, 0.0, C1.data_ptr(), n) == 0
        assert lib.mm_tiles_info(h)["state"] == 2
        assert lib.update_values("d", h, len(val), 2.0 * val) == 0
        assert lib.csrmm("d", 111, 1.0, h, d, 0, B.data_ptr(), n, n, 0.0, C2.data_ptr(), n) == 0
        torch.cuda.synchronize()
        assert torch.equal(C2, 2.0 * C1)
        lib.destroy(h)
        lib.destroy_descr(d)
    finally:
        os.environ.pop("AOCLSPARSE_B200_MM_TILES", None)


@pytest.mark.parametrize("p", ["s", "d", "z"])
@pytest.mark.parametrize("grid", [(27, 24, 12, 9), (7, 16, 16, 9), (5, 33, 21, 1), (27, 24, 9, 5)], ids=["27pt", "7pt", "5pt", "27pt-ragged"])
def test_tile_arrays_bit_exact_against_the_cpu_restatement(lib, p, grid):
    """the tile analysis is integer work with no counterpart in the reference: lattice strides, box, runs, walks, row
    lists and the re-ordered values must equal tests/mesh_tiles_ref.py bit for bit"""
    import torch
    import mesh_tiles_ref as ref
    pts, nx, ny, nz = grid
    rng = np.random.default_rng(17)
    dt = DT[p]
    rp, col, val = gen_np.stencil(pts, nx, ny, nz)
    val = _values(rng, val, dt)
    m = len(rp) - 1
    elem = np.dtype(dt).itemsize
    n = 256 // elem
    os.environ["AOCLSPARSE_B200_MM_TILES"] = "2"
    try:
        st, h = lib.create_csr(p, 0, m, m, len(col), rp, col, val)
        assert st == 0
        d = lib.create_descr()
        B = torch.zeros(m * n, dtype=torch.from_numpy(np.zeros(1, dtype=dt)).dtype, device="cuda")
        Cc = torch.zeros_like(B)
        assert lib.csrmm(p, 111, 1.0, h, d, 0, B.data_ptr(), n, n, 0.0, Cc.data_ptr(), n) == 0, lib.last_error()
        torch.cuda.synchronize()
        info = lib.mm_tiles_info(h)
        offs = ref.diag_offsets(rp, col)
        ok, ndim, s1, s2 = ref.detect_lattice(offs, m)
        assert ok
        gs1, gs2, gx, gy, gz = ref.grid_of(m, ndim, s1, s2)
        assert info["strides"] == [1, gs1, gs2] and info["dims"] == [gx, gy, gz]
        box = ref.choose_box(ndim, gx, gy, gz, len(offs), 256, elem)
        want = ref.build(rp, col, val, box, ndim, s1, s2)
        if info["state"] != 2:
            # the library declined (too little re-use, or too much padding on a grid the boxes do not divide)
            assert (pts, nx, ny, nz) != (27, 24, 12, 9)
            return
        assert info["box"] == box and info["rows_per_group"] == ref.GRP
        got = lib.mm_tiles(h, dt)
        for k in ("desc", "off", "rows", "runs", "walk"):
            assert np.array_equal(got[k].reshape(-1), want[k].reshape(-1)), k
        assert np.array_equal(got["val"].view(np.uint8), want["val"].view(np.uint8))
        lib.destroy(h)
        lib.destroy_descr(d)
    finally:
        os.environ.pop("AOCLSPARSE_B200_MM_TILES", None)
