#!/usr/bin/env python
"""Times aoclsparse_itsol_d_solve (conjugate gradients, device-resident b / x) on the 27-point 128^3 matrix of config 2,
and the reference's own CG (oracle/_ref) on the same system on the host.  A measurement aid for profiles/."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import capi  # noqa: E402
import gen_np  # noqa: E402

if __name__ == "__main__":
    import torch
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rp, col, val = gen_np.stencil(27, nx, nx, nx)
    n = len(rp) - 1
    b = gen_np.uniform(1, 0, n)
    for name, path, dev in (("GPU", None, True), ("reference CPU", os.path.join(ROOT, "oracle", "_ref", "libaoclsparse_ref.so"), False)):
        if path and not os.path.exists(path):
            continue
        lib = capi.AoclSparse(path) if path else capi.AoclSparse()
        st, A = lib.create_csr("d", 0, n, n, len(col), rp, col, val)
        assert st == 0
        d = lib.create_descr(1, 0, 0, 0)
        st, it = lib.itsol_init("d")
        assert lib.itsol_option_set(it, "cg iteration limit", "50") == 0
        assert lib.itsol_option_set(it, "cg rel tolerance", "1e-30") == 0 and lib.itsol_option_set(it, "cg abs tolerance", "0") == 0
        rinfo = np.zeros(100)
        best = 1e30
        for rep in range(3 if dev else 1):
            if dev:
                db, dx = torch.from_numpy(b).cuda(), torch.zeros(n, dtype=torch.float64, device="cuda")
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                st = lib.itsol_solve("d", it, n, A, d, db.data_ptr(), dx.data_ptr(), rinfo)
                torch.cuda.synchronize()
            else:
                x = np.zeros(n)
                t0 = time.perf_counter()
                st = lib.itsol_solve("d", it, n, A, d, b, x, rinfo)
            best = min(best, time.perf_counter() - t0)
        iters = int(rinfo[30])
        print(f"{name}: CG on 27-pt {nx}^3 ({n} unknowns, {len(col)} entries): status {st}, {iters} iterations, "
              f"{best*1e3:.1f} ms = {best*1e3/iters:.3f} ms / iteration, |r| = {rinfo[0]:.3e}")
