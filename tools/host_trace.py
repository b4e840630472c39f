#!/usr/bin/env python
"""timeline of one host-resident aoclsparse_dmv on config 2 (AOCLSPARSE_B200_HOST_TRACE=1 makes the library print it)"""
import os
import sys
import time

import numpy as np

os.environ["AOCLSPARSE_B200_HOST_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import capi  # noqa: E402
import gen_np  # noqa: E402

if __name__ == "__main__":
    import torch
    lib = capi.AoclSparse()
    rp, col, val = gen_np.stencil(27, 128, 128, 128)
    m = len(rp) - 1
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    d = lib.create_descr()
    assert lib.set_mv_hint(h, 111, d, 100) == 0 and lib.optimize(h) == 0
    hx = torch.ones(m, dtype=torch.float64).pin_memory()
    hy = torch.zeros(m, dtype=torch.float64).pin_memory()
    for i in range(6):
        t0 = time.perf_counter()
        assert lib.mv("d", 111, 1.0, h, d, hx.data_ptr(), 0.0, hy.data_ptr()) == 0
        print(f"call {i}: {1e6*(time.perf_counter()-t0):.0f} us wall", file=sys.stderr)
