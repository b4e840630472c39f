#!/usr/bin/env python
"""Is config 1 (17-18 us per SpMV) limited by the host's enqueue rate?  Times the same 5-set rotation of aoclsparse_dmv
calls (a) issued one by one from Python and (b) captured once into a CUDA graph and replayed."""
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
    lib = capi.AoclSparse()
    s = torch.cuda.Stream()
    rp, col, val = gen_np.stencil(5, 1000, 1000, 1)
    m, nnz = len(rp) - 1, len(col)
    d = lib.create_descr()
    sets = []
    for k in range(5):
        st, h = lib.create_csr("d", 0, m, m, nnz, rp, col, val)
        assert st == 0 and lib.set_mv_hint(h, 111, d, 1000) == 0 and lib.optimize(h) == 0
        sets.append((h, torch.ones(m, dtype=torch.float64, device="cuda"), torch.zeros(m, dtype=torch.float64, device="cuda")))
    byts = (m + 1 + nnz) * 4 + (3 * m + nnz) * 8

    def step(i):
        h, x, y = sets[i % 5]
        return lib.mv("d", 111, 1.0, h, d, x.data_ptr(), 0.5, y.data_ptr())
    with torch.cuda.stream(s):
        lib.set_stream(s.cuda_stream)
        for i in range(20):
            assert step(i) == 0
        s.synchronize()
        # host-side cost of one call (no GPU wait): enqueue 2000 calls, measure wall time until the last enqueue returns
        t0 = time.perf_counter()
        for i in range(2000):
            step(i)
        t_enq = (time.perf_counter() - t0) / 2000
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for i in range(500):
            step(i)
        e1.record(s)
        s.synchronize()
        loop_us = e0.elapsed_time(e1) / 500 * 1e3
        g = torch.cuda.CUDAGraph()
        ok = True
        try:
            with torch.cuda.graph(g, stream=s):
                for i in range(100):
                    assert step(i) == 0, lib.last_error()
        except Exception as ex:  # capture not possible: say so
            ok = False
            print("graph capture failed:", repr(ex)[:300])
        if ok:
            for _ in range(3):
                g.replay()
            s.synchronize()
            e0.record(s)
            for _ in range(5):
                g.replay()
            e1.record(s)
            s.synchronize()
            graph_us = e0.elapsed_time(e1) / 500 * 1e3
            print(f"C1: host enqueue {t_enq*1e6:.1f} us/call; python loop {loop_us:.2f} us/SpMV ({byts/loop_us/1e3:.0f} GB/s); "
                  f"graph replay {graph_us:.2f} us/SpMV ({byts/graph_us/1e3:.0f} GB/s)")
        else:
            print(f"C1: host enqueue {t_enq*1e6:.1f} us/call; python loop {loop_us:.2f} us/SpMV")
