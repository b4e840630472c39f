#!/usr/bin/env python
"""Times C = A A (aoclsparse_spmm) for the 27-point stencil on the GPU library and, beside it, on the reference's own
CPU build (oracle/_ref) when present.  A measurement aid for DESIGN.md / profiles/, not part of bench.py's contract."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import capi  # noqa: E402
import gen_np  # noqa: E402


def run(lib, rp, col, val, reps, sync=None):
    m = len(rp) - 1
    st, h = lib.create_csr("d", 0, m, m, len(col), rp, col, val)
    assert st == 0
    best, nnz = 1e30, 0
    for _ in range(reps):
        if sync:
            sync()
        t0 = time.perf_counter()
        st, c = lib.spmm(111, h, h)
        if sync:
            sync()
        dt = time.perf_counter() - t0
        assert st == 0, st
        best = min(best, dt)
        d0 = lib.create_descr()
        st, c2 = lib.sp2m(111, d0, h, 111, d0, h, 0)   # nnz_count stage alone
        lib.destroy_descr(d0)
        lib.destroy(c2)
        lib.destroy(c)
    t0 = time.perf_counter()
    d0 = lib.create_descr()
    st, c2 = lib.sp2m(111, d0, h, 111, d0, h, 0)
    if sync:
        sync()
    t_count = time.perf_counter() - t0
    info_nnz = None
    if hasattr(lib.lib, "aoclsparse_b200_get_matrix_info"):
        info_nnz = lib.matrix_info(c2).nnz
    lib.destroy(c2)
    lib.destroy_descr(d0)
    lib.destroy(h)
    return best, t_count, info_nnz


if __name__ == "__main__":
    import torch
    for nx in (64, 128):
        rp, col, val = gen_np.stencil(27, nx, nx, nx)
        products = 0
        lens = np.diff(rp)
        products = int(np.sum(lens[col]))
        gpu = capi.AoclSparse()
        t, tc, nnzc = run(gpu, rp, col, val, 3, torch.cuda.synchronize)
        print(f"27-pt {nx}^3 squared: rows {len(rp)-1}, nnz(A) {len(col)}, products {products}, nnz(C) {nnzc}: "
              f"GPU full {t*1e3:.1f} ms (nnz_count stage {tc*1e3:.1f} ms), {2*products/t/1e9:.1f} GFLOP/s")
        ref_path = os.path.join(ROOT, "oracle", "_ref", "libaoclsparse_ref.so")
        if os.path.exists(ref_path) and nx == 64:
            ref = capi.AoclSparse(ref_path)
            t, tc, _ = run(ref, rp, col, val, 2)
            print(f"   reference CPU ({os.cpu_count()} hw threads): full {t*1e3:.1f} ms (nnz_count {tc*1e3:.1f} ms), "
                  f"{2*products/t/1e9:.2f} GFLOP/s")
