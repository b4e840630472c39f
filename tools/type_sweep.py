#!/usr/bin/env python
"""Effective bandwidth of aoclsparse_?mv for all four value types (and op = transpose on a hinted handle) on the 27-point
128^3 matrix, device-resident operands, CUDA events.  Shows whether the kernel holds its roofline beyond config 2's double."""
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
    lib = capi.AoclSparse()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    rp, col, val = gen_np.stencil(27, 128, 128, 128)
    m, nnz = len(rp) - 1, len(col)
    for p, dt, tdt, elem in (("s", np.float32, torch.float32, 4), ("d", np.float64, torch.float64, 8),
                             ("c", np.complex64, torch.complex64, 8), ("z", np.complex128, torch.complex128, 16)):
        st, h = lib.create_csr(p, 0, m, m, nnz, rp, col, val.astype(dt))
        assert st == 0
        for op, tname, mtype in ((111, "none", 0), (112, "transpose (hinted)", 0), (111, "symmetric lower (hinted)", 1)):
            d = lib.create_descr(mtype, 0, 0, 0)
            assert lib.set_mv_hint(h, op, d, 100) == 0 and lib.optimize(h) == 0
            x = torch.ones(m, dtype=tdt, device="cuda")
            y = torch.zeros(m, dtype=tdt, device="cuda")
            for _ in range(5):
                assert lib.mv(p, op, 1.0, h, d, x.data_ptr(), 0.0, y.data_ptr()) == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                lib.mv(p, op, 1.0, h, d, x.data_ptr(), 0.0, y.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 50
            byts = (m + 1 + nnz) * 4 + (2 * m + nnz) * elem
            print(f"{p} op={tname:26s} {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s  ({byts/ms/1e6/6551*100:5.1f} % of 6551)")
            lib.destroy_descr(d)
        lib.destroy(h)
