#!/usr/bin/env python
"""Breadth check of the multiply paths on the 27-point 128^3 matrix (device-resident operands, CUDA events):
csrmm across value types / widths / layouts, and the mv paths that are not config 2 (beta != 0, un-hinted transpose =
atomic scatter, CSC handle).  Effective GB/s by the reference's byte models."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import capi  # noqa: E402
import gen_np  # noqa: E402

TY = {"s": (np.float32, 4), "d": (np.float64, 8), "c": (np.complex64, 8), "z": (np.complex128, 16)}


def timeit(fn, reps):
    import torch
    for _ in range(40):  # past the lazy-copy threshold of un-hinted transposed / symmetric products
        assert fn() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if __name__ == "__main__":
    import torch
    tt = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}
    lib = capi.AoclSparse()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    rp, col, val = gen_np.stencil(27, 128, 128, 128)
    m, nnz = len(rp) - 1, len(col)
    d = lib.create_descr()
    for p, n, order in [("d", 32, 0), ("s", 32, 0), ("c", 32, 0), ("z", 32, 0), ("d", 8, 0), ("d", 16, 0), ("d", 64, 0),
                        ("d", 128, 0), ("s", 64, 0), ("d", 32, 1), ("d", 8, 1), ("d", 33, 0)]:
        dt, elem = TY[p]
        st, h = lib.create_csr(p, 0, m, m, nnz, rp, col, val.astype(dt))
        assert st == 0
        assert lib.set_mm_hint(h, 111, d, 100) == 0 and lib.optimize(h) == 0
        B = torch.ones(m * n, dtype=tt[p], device="cuda")
        Cm = torch.zeros(m * n, dtype=tt[p], device="cuda")
        ld = n if order == 0 else m
        ms = timeit(lambda: lib.csrmm(p, 111, 1.0, h, d, order, B.data_ptr(), n, ld, 0.0, Cm.data_ptr(), ld), 10)
        byts = (m + 1 + nnz) * 4 + (nnz + 2 * m * n) * elem
        fl = 2.0 * nnz * n * (4 if p in "cz" else 1)
        print(f"csrmm {p} n={n:3d} {'row' if order == 0 else 'col'}-major: {ms:7.3f} ms  {byts/ms/1e6:6.0f} GB/s ({byts/ms/1e6/65.51:5.1f} %)  "
              f"{fl/ms/1e9:6.2f} TFLOP/s")
        lib.destroy(h)
        del B, Cm
    # mv variants, double
    st, h = lib.create_csr("d", 0, m, m, nnz, rp, col, val)
    x = torch.ones(m, dtype=torch.float64, device="cuda")
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    byts = (m + 1 + nnz) * 4 + (2 * m + nnz) * 8
    ms = timeit(lambda: lib.mv("d", 111, 1.0, h, d, x.data_ptr(), 0.5, y.data_ptr()), 30)
    print(f"mv d beta=0.5 (no hint):            {ms*1e3:7.1f} us {(byts + 8*m)/ms/1e6:6.0f} GB/s")
    ms = timeit(lambda: lib.mv("d", 112, 1.0, h, d, x.data_ptr(), 0.0, y.data_ptr()), 10)
    print(f"mv d transpose, NOT hinted (scatter): {ms*1e3:7.1f} us {byts/ms/1e6:6.0f} GB/s")
    dsym = lib.create_descr(1, 0, 0, 0)
    ms = timeit(lambda: lib.mv("d", 111, 1.0, h, dsym, x.data_ptr(), 0.0, y.data_ptr()), 10)
    print(f"mv d symmetric lower, NOT hinted:   {ms*1e3:7.1f} us")
    dtri = lib.create_descr(3, 0, 0, 0)
    ms = timeit(lambda: lib.mv("d", 111, 1.0, h, dtri, x.data_ptr(), 0.0, y.data_ptr()), 10)
    print(f"mv d triangular lower:              {ms*1e3:7.1f} us")
    lib.destroy(h)
    import scipy.sparse as sp
    Ac = sp.csr_matrix((val, col, rp)).tocsc()
    st, hc = lib.create_csc("d", 0, m, m, nnz, Ac.indptr.astype(np.int32), Ac.indices.astype(np.int32), Ac.data)
    assert st == 0
    ms = timeit(lambda: lib.mv("d", 111, 1.0, hc, d, x.data_ptr(), 0.0, y.data_ptr()), 5)
    print(f"mv d CSC handle, NOT hinted:        {ms*1e3:7.1f} us")
    assert lib.set_mv_hint(hc, 111, d, 100) == 0 and lib.optimize(hc) == 0
    ms = timeit(lambda: lib.mv("d", 111, 1.0, hc, d, x.data_ptr(), 0.0, y.data_ptr()), 30)
    print(f"mv d CSC handle, hinted:            {ms*1e3:7.1f} us {byts/ms/1e6:6.0f} GB/s")
