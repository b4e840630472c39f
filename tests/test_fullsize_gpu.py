"""BASELINE.json's configurations at FULL size, CUDA path against the reference's own build on the same host arrays.

`oracle/_ref/libaoclsparse_ref.so` (the reference compiled from its sources by oracle/Makefile) travels to the GPU box
as a prebuilt file.  Every case builds the synthetic matrix of SURVEY.md section 8(d) in device memory, copies the
three CSR arrays to the host ONCE, and hands those same host arrays to both libraries through the same ctypes binding
(the reference aliases them, the product uploads them).  The comparison is the parity protocol of section 8(d), the
way the reference's own harness verifies (tests/include/aoclsparse_check.hpp:35-128 compares against its host
product): per output entry |y_gpu - y_ref| <= tol * (sum_j |a_ij||x_j| + |beta*y0_i|), tol = 1e-12 (double) /
1e-5 (float); config 3 is also compared with an fp64 accumulation, which separates the reference's float rounding on hub
rows from ours.  The scale-24 R-MAT case is the only one whose plan holds rows of more than 32 segments (the strided
branch of finish_long_rows_kernel) and, with config 5's 9.4e8 entries, the one that needs 64-bit offset arithmetic.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _row_ids(rp_t):
    import torch
    m = rp_t.numel() - 1
    return torch.repeat_interleave(torch.arange(m, device=rp_t.device), (rp_t[1:] - rp_t[:-1]).long())


def _row_scale(rp_t, col_t, val_t, x_t, rows=None):
    """sum_j |a_ij| |x_j| per row in fp64 (device)"""
    import torch
    rows = _row_ids(rp_t) if rows is None else rows
    prods = val_t.abs().double() * x_t.abs().double()[col_t.long()]
    return torch.zeros(rp_t.numel() - 1, dtype=torch.float64, device=rp_t.device).index_add_(0, rows, prods)


def _ref_mv(reflib, p, m, n, rp, col, val, x, alpha, beta, y0):
    st, h = reflib.create_csr(p, 0, m, n, len(col), rp, col, val)
    assert st == 0, st
    d = reflib.create_descr()
    y = y0.copy()
    assert reflib.mv(p, 111, alpha, h, d, x, beta, y) == 0
    reflib.destroy_descr(d)
    reflib.destroy(h)
    return y


def _gpu_handle(lib, p, m, n, rp, col, val, hint="mv"):
    st, h = lib.create_csr(p, 0, m, n, len(col), rp, col, val)  # host arrays: the drop-in call
    assert st == 0, (st, lib.last_error())
    d = lib.create_descr()
    if hint == "mv":
        assert lib.set_mv_hint(h, 111, d, 1000) == 0
    elif hint == "mm":
        assert lib.set_mm_hint(h, 111, d, 1000) == 0
    if hint:
        assert lib.optimize(h) == 0, lib.last_error()
    return h, d


def _stencil_host(lib, pts, nx, ny, nz, lo=0, hi=None):
    import torch
    total = nx * ny * nz
    hi = total if hi is None else hi
    nnz = C.c_longlong(0)
    assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, lo, hi, C.byref(nnz), None, None, None) == 0
    rp = torch.empty(hi - lo + 1, dtype=torch.int32, device="cuda")
    col = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
    assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, lo, hi, C.byref(nnz), rp.data_ptr(), col.data_ptr(),
                                               val.data_ptr()) == 0, lib.last_error()
    torch.cuda.synchronize()
    return (rp, col, val), (rp.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy())


def _uniform(lib, seed, first, count, dt):
    import torch
    t = torch.empty(count, dtype=dt, device="cuda")
    assert lib.lib.aoclsparse_b200_gen_uniform(seed, first, count, t.element_size(), t.data_ptr()) == 0
    torch.cuda.synchronize()
    return t


@pytest.mark.parametrize("cfg", [("c1", 5, 1000, 1000, 1, 4996000, 1.0, 0.5), ("c2", 27, 128, 128, 128, 55742968, 1.0, 0.0)])
def test_full_size_stencil_mv_vs_reference(lib, reflib, cfg):
    """configs 1 and 2: aoclsparse_dmv at BASELINE size against the reference's aoclsparse_dmv"""
    import torch
    name, pts, nx, ny, nz, want_nnz, alpha, beta = cfg
    (rp_t, col_t, val_t), (rp, col, val) = _stencil_host(lib, pts, nx, ny, nz)
    m = len(rp) - 1
    assert len(col) == want_nnz
    x_t = _uniform(lib, 1, 0, m, torch.float64)
    y0_t = _uniform(lib, 2, 0, m, torch.float64)
    x, y0 = x_t.cpu().numpy(), y0_t.cpu().numpy()
    yref = _ref_mv(reflib, "d", m, m, rp, col, val, x, alpha, beta, y0)
    h, d = _gpu_handle(lib, "d", m, m, rp, col, val)
    y_t = y0_t.clone() if beta else torch.full((m,), float("nan"), dtype=torch.float64, device="cuda")
    assert lib.mv("d", 111, alpha, h, d, x_t.data_ptr(), beta, y_t.data_ptr()) == 0, lib.last_error()
    torch.cuda.synchronize()
    den = abs(alpha) * _row_scale(rp_t, col_t, val_t, x_t) + abs(beta) * y0_t.abs()
    err = float(torch.max((y_t - torch.from_numpy(yref).cuda()).abs() / den))
    assert err <= 1e-12, (name, err)
    # the same call with HOST vectors (what an unmodified caller of the reference passes)
    yh = y0.copy()
    assert lib.mv("d", 111, alpha, h, d, x, beta, yh) == 0, lib.last_error()
    assert np.array_equal(yh, y_t.cpu().numpy()), "host-staged and device-resident products differ"
    lib.destroy(h)
    lib.destroy_descr(d)


def test_full_size_rmat24_float_vs_reference(lib, reflib):
    """config 3: R-MAT scale 24, float, against the reference's aoclsparse_smv and against an fp64 accumulation"""
    import torch

    import bench
    m, n, nnz, rp_t, col_t, val_t = bench.device_matrix(lib, bench.WORKLOADS["c3"])
    assert m == 1 << 24 and 0.9 * 16 * m < nnz <= 16 * m
    rp, col, val = rp_t.cpu().numpy(), col_t.cpu().numpy(), val_t.cpu().numpy()
    x_t = _uniform(lib, 1, 0, n, torch.float32)
    x = x_t.cpu().numpy()
    yref = _ref_mv(reflib, "s", m, n, rp, col, val, x, 1.0, 0.0, np.zeros(m, np.float32))
    h, d = _gpu_handle(lib, "s", m, n, rp, col, val)
    info = lib.matrix_info(h)
    # hub rows: the longest row must span more than 32 segments, so that finish_long_rows_kernel's lane-strided
    # loop over a row's partial sums runs more than once per lane
    assert info.n_long_rows > 0 and info.max_row_nnz > 32 * info.block_nnz, (info.max_row_nnz, info.block_nnz)
    y_t = torch.full((m,), float("nan"), dtype=torch.float32, device="cuda")
    assert lib.mv("s", 111, 1.0, h, d, x_t.data_ptr(), 0.0, y_t.data_ptr()) == 0, lib.last_error()
    torch.cuda.synchronize()
    rows = _row_ids(rp_t)
    den = _row_scale(rp_t, col_t, val_t, x_t, rows)
    y64 = torch.zeros(m, dtype=torch.float64, device="cuda").index_add_(0, rows, val_t.double() * x_t.double()[col_t.long()])
    del rows
    safe = torch.where(den > 0, den, torch.ones_like(den))
    err_ref = float(torch.max((y_t.double() - torch.from_numpy(yref).cuda().double()).abs() / safe))
    err_64 = float(torch.max((y_t.double() - y64).abs() / safe))
    ref_64 = float(torch.max((torch.from_numpy(yref).cuda().double() - y64).abs() / safe))
    print(f"rmat24: gpu-ref {err_ref:.3e}  gpu-fp64 {err_64:.3e}  ref-fp64 {ref_64:.3e}")
    assert err_64 <= 1e-5, err_64
    assert err_ref <= 1e-5, (err_ref, ref_64)
    empty = (rp_t[1:] == rp_t[:-1])
    assert bool(torch.all(y_t[empty] == 0))  # beta = 0: empty rows are overwritten with zeros, NaN in y not read
    # run-to-run reproducible (fixed summation order, no atomics)
    y2 = torch.empty_like(y_t)
    assert lib.mv("s", 111, 1.0, h, d, x_t.data_ptr(), 0.0, y2.data_ptr()) == 0
    torch.cuda.synchronize()
    assert torch.equal(y_t, y2)
    lib.destroy(h)
    lib.destroy_descr(d)


def test_full_size_csrmm_vs_reference(lib, reflib):
    """config 4: 27-point 128^3 times a dense 2 097 152 x 32 row-major block against the reference's aoclsparse_dcsrmm"""
    import torch
    (rp_t, col_t, val_t), (rp, col, val) = _stencil_host(lib, 27, 128, 128, 128)
    m, nr = len(rp) - 1, 32
    B_t = _uniform(lib, 3, 0, m * nr, torch.float64)
    B = B_t.cpu().numpy()
    st, hr = reflib.create_csr("d", 0, m, m, len(col), rp, col, val)
    assert st == 0
    dr = reflib.create_descr()
    Cref = np.zeros(m * nr)
    assert reflib.csrmm("d", 111, 1.0, hr, dr, 0, B, nr, nr, 0.0, Cref, nr) == 0
    reflib.destroy(hr)
    reflib.destroy_descr(dr)
    h, d = _gpu_handle(lib, "d", m, m, rp, col, val, hint="mm")
    C_t = torch.full((m * nr,), float("nan"), dtype=torch.float64, device="cuda")
    assert lib.csrmm("d", 111, 1.0, h, d, 0, B_t.data_ptr(), nr, nr, 0.0, C_t.data_ptr(), nr) == 0, lib.last_error()
    torch.cuda.synchronize()
    Aabs = torch.sparse_csr_tensor(rp_t.long(), col_t.long(), val_t.abs(), size=(m, m))
    den = (Aabs @ B_t.abs().view(m, nr)).view(-1)
    err = float(torch.max((C_t - torch.from_numpy(Cref).cuda()).abs() / den))
    assert err <= 1e-12, err
    lib.destroy(h)
    lib.destroy_descr(d)


def test_full_size_c5_slab_vs_reference(lib, reflib):
    """config 5: eight grid planes (2 097 152 rows) of the 512^3 7-point stencil -- the slab a rank owns -- against the
    reference on the same rows, (a) as an m x n rectangle with the whole x, (b) the way the sharded iteration uses it:
    x window [lo - plane, hi + plane) and row cuts at the first / last plane; (a) and (b) must agree bit for bit"""
    import torch
    nx = 512
    plane, total = nx * nx, nx ** 3
    lo, hi = 256 * plane, 264 * plane
    (rp_t, col_t, val_t), (rp, col, val) = _stencil_host(lib, 7, nx, nx, nx, lo, hi)
    m = hi - lo
    x_t = _uniform(lib, 1, 0, total, torch.float64)
    x = x_t.cpu().numpy()
    alpha = 1.0 / 12.0
    yref = _ref_mv(reflib, "d", m, total, rp, col, val, x, alpha, 0.0, np.zeros(m))
    h, d = _gpu_handle(lib, "d", m, total, rp, col, val)
    y_t = torch.empty(m, dtype=torch.float64, device="cuda")
    assert lib.mv("d", 111, alpha, h, d, x_t.data_ptr(), 0.0, y_t.data_ptr()) == 0, lib.last_error()
    torch.cuda.synchronize()
    den = alpha * _row_scale(rp_t, col_t, val_t, x_t)
    err = float(torch.max((y_t - torch.from_numpy(yref).cuda()).abs() / den))
    assert err <= 1e-12, err
    lib.destroy(h)
    h, _ = _gpu_handle(lib, "d", m, total, rp, col, val, hint=None)
    assert lib.set_x_window(h, lo - plane, hi + plane) == 0
    assert lib.set_row_cuts(h, [plane, m - plane]) == 0
    assert lib.set_mv_hint(h, 111, d, 1000) == 0 and lib.optimize(h) == 0
    yw = torch.empty_like(y_t)
    xw = x_t[lo - plane: hi + plane].clone()
    assert lib.mv("d", 111, alpha, h, d, xw.data_ptr(), 0.0, yw.data_ptr()) == 0, lib.last_error()
    torch.cuda.synchronize()
    assert torch.equal(yw, y_t)
    lib.destroy(h)
    lib.destroy_descr(d)


def test_full_size_c5_whole_matrix_row_sums(lib):
    """config 5 on one GPU at full size (134 217 728 rows, 937 951 232 entries: entry offsets beyond 2^29 elements,
    byte offsets beyond 2^32): A * 1 = 7 - nnz(row) exactly, and the first iterate of the bench recurrence on three
    probe planes against the host stencil formula"""
    import torch
    nx = 512
    total = nx ** 3
    nnz = C.c_longlong(0)
    assert lib.lib.aoclsparse_b200_gen_stencil(7, nx, nx, nx, 0, total, C.byref(nnz), None, None, None) == 0
    assert nnz.value == 937951232
    rp = torch.empty(total + 1, dtype=torch.int32, device="cuda")
    col = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
    assert lib.lib.aoclsparse_b200_gen_stencil(7, nx, nx, nx, 0, total, C.byref(nnz), rp.data_ptr(), col.data_ptr(),
                                               val.data_ptr()) == 0
    st, h = lib.create_csr("d", 0, total, total, nnz.value, rp.data_ptr(), col.data_ptr(), val.data_ptr())
    assert st == 0, lib.last_error()
    del col, val
    torch.cuda.empty_cache()
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 111, d, 1000) == 0 and lib.optimize(h) == 0
    ones = torch.ones(total, dtype=torch.float64, device="cuda")
    y = torch.empty(total, dtype=torch.float64, device="cuda")
    assert lib.mv("d", 111, 1.0, h, d, ones.data_ptr(), 0.0, y.data_ptr()) == 0, lib.last_error()
    torch.cuda.synchronize()
    assert torch.equal(y, 7.0 - (rp[1:] - rp[:-1]).to(torch.float64))
    del ones
    x = _uniform(lib, 1, 0, total, torch.float64)
    assert lib.mv("d", 111, 1.0 / 12.0, h, d, x.data_ptr(), 0.0, y.data_ptr()) == 0
    torch.cuda.synchronize()
    g = x.view(nx, nx, nx)
    for z in (0, 255, 511):  # first, interior and last plane (the last one sits past 2^31 bytes of col_idx)
        want = 6.0 * g[z].clone()
        for dz in (-1, 1):
            if 0 <= z + dz < nx:
                want -= g[z + dz]
        want[1:, :] -= g[z][:-1, :]
        want[:-1, :] -= g[z][1:, :]
        want[:, 1:] -= g[z][:, :-1]
        want[:, :-1] -= g[z][:, 1:]
        got = y.view(nx, nx, nx)[z]
        assert float(torch.max((got - want / 12.0).abs())) <= 1e-12 * 12.0 / 12.0 * 7, z
    lib.destroy(h)
    lib.destroy_descr(d)
