#!/usr/bin/env python
"""double-complex SpMV on the 27-point 128^3 matrix under the plan knobs given in the environment (one line of output)"""
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
    p = sys.argv[1] if len(sys.argv) > 1 else "z"
    dt, tdt, elem = {"z": (np.complex128, torch.complex128, 16), "c": (np.complex64, torch.complex64, 8),
                     "d": (np.float64, torch.float64, 8), "s": (np.float32, torch.float32, 4)}[p]
    lib = capi.AoclSparse()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    rp, col, val = gen_np.stencil(27, 128, 128, 128)
    m, nnz = len(rp) - 1, len(col)
    st, h = lib.create_csr(p, 0, m, m, nnz, rp, col, val.astype(dt))
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 111, d, 100) == 0 and lib.optimize(h) == 0
    info = lib.matrix_info(h)
    x = torch.ones(m, dtype=tdt, device="cuda")
    y = torch.zeros(m, dtype=tdt, device="cuda")
    for _ in range(5):
        assert lib.mv(p, 111, 1.0, h, d, x.data_ptr(), 0.0, y.data_ptr()) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        lib.mv(p, 111, 1.0, h, d, x.data_ptr(), 0.0, y.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    byts = (m + 1 + nnz) * 4 + (2 * m + nnz) * elem
    knobs = {k: v for k, v in os.environ.items() if k.startswith("AOCLSPARSE_B200_")}
    print(f"{p} {knobs} T={info.block_nnz} blocks={info.n_blocks} thread/warp/product={info.n_thread_blocks}/{info.n_warp_blocks}/{info.n_product_blocks}: "
          f"{ms*1e3:7.1f} us {byts/ms/1e6:6.0f} GB/s")
