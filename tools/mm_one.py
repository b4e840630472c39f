#!/usr/bin/env python
"""one csrmm timing on the 27-point 128^3 matrix: python tools/mm_one.py <s|d|c|z> <n> [row|col]; knobs from the environment"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import capi  # noqa: E402
import gen_np  # noqa: E402

if __name__ == "__main__":
    import torch
    p, n = sys.argv[1], int(sys.argv[2])
    order = 1 if (len(sys.argv) > 3 and sys.argv[3] == "col") else 0
    dt, elem, tdt = {"s": (np.float32, 4, torch.float32), "d": (np.float64, 8, torch.float64),
                     "c": (np.complex64, 8, torch.complex64), "z": (np.complex128, 16, torch.complex128)}[p]
    lib = capi.AoclSparse()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    rp, col, val = gen_np.stencil(27, 128, 128, 128)
    m, nnz = len(rp) - 1, len(col)
    d = lib.create_descr()
    st, h = lib.create_csr(p, 0, m, m, nnz, rp, col, val.astype(dt))
    assert lib.set_mm_hint(h, 111, d, 100) == 0 and lib.optimize(h) == 0
    B = torch.ones(m * n, dtype=tdt, device="cuda")
    Cm = torch.zeros(m * n, dtype=tdt, device="cuda")
    ld = n if order == 0 else m
    call = lambda: lib.csrmm(p, 111, 1.0, h, d, order, B.data_ptr(), n, ld, 0.0, Cm.data_ptr(), ld)  # noqa: E731
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    knobs = {k[15:]: v for k, v in os.environ.items() if k.startswith("AOCLSPARSE_B200_")}
    print(f"csrmm {p} n={n} {'col' if order else 'row'} {knobs}: {ms:.3f} ms")
