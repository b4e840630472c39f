#!/usr/bin/env python
"""Experiment (round 2): config 3 (R-MAT scale 24, float) with the columns RENUMBERED so that the most referenced ones
are contiguous -- x' = x[perm] is then gathered instead of x.  The hot columns of an R-MAT matrix are scattered (each
shares its 32-byte sector with seven cold ones), so L1 holds them badly (hit rate 6 %); packed together, the hottest
K columns are K*4 bytes of contiguous lines that stay resident in the 192 KB of L1 the row-block kernel leaves.
Measures only the multiply (same library kernel, remapped col_idx, permuted x) plus the cost of the permutation of x;
prints one line per variant: label, ms per product, ms of the x permutation, GFLOP/s incl. permutation.
Usage: python tools/experiments/c3_hotfirst.py [scale]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import capi  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
lib = capi.AoclSparse()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
lib.set_stream(stream.cuda_stream)
wl = dict(bench.WORKLOADS["c3"], rmat=scale)
m, n, nnz, rp, col, val = bench.device_matrix(lib, wl)
x = torch.rand(n, dtype=torch.float32, device="cuda")
d = lib.create_descr()
STEPS = 30


def run(colx, xx, label, perm=None):
    st, A = lib.create_csr("s", 0, m, n, nnz, rp.data_ptr(), colx.data_ptr(), val.data_ptr())
    assert st == 0, lib.last_error()
    assert lib.set_mv_hint(A, 111, d, 1000) == 0 and lib.optimize(A) == 0
    y = torch.empty(m, dtype=torch.float32, device="cuda")
    for _ in range(5):
        assert lib.mv("s", 111, 1.0, A, d, xx.data_ptr(), 0.0, y.data_ptr()) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(STEPS):
        assert lib.mv("s", 111, 1.0, A, d, xx.data_ptr(), 0.0, y.data_ptr()) == 0
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    pms = 0.0
    if perm is not None:
        xp = torch.empty_like(x)
        for _ in range(3):
            torch.index_select(x, 0, perm, out=xp)
        e0.record(stream)
        for _ in range(STEPS):
            torch.index_select(x, 0, perm, out=xp)
        e1.record(stream)
        torch.cuda.synchronize()
        pms = e0.elapsed_time(e1) / STEPS
    print("## %-34s %.4f ms  perm %.4f ms  -> %.1f GFLOP/s  y.sum %.6e" % (label, ms, pms, 2.0 * nnz / ((ms + pms) * 1e-3) / 1e9,
                                                                          float(y.double().sum())), flush=True)
    lib.destroy(A)
    return y


y0 = run(col, x, "storage order")
cnt = torch.bincount(col.long(), minlength=n)
order = torch.argsort(cnt, descending=True, stable=True)
csum = torch.cumsum(cnt[order], 0)
for K in (4096, 16384, 32768, 65536, 262144, n):
    if K > n:
        continue
    hot = torch.zeros(n, dtype=torch.bool, device="cuda")
    hot[order[:K]] = True
    if K == n:
        perm = order  # every column by falling reference count
        label = "all columns by count"
    else:
        # the K hottest first (by falling count), the others after them in their original order
        perm = torch.cat([order[:K], torch.nonzero(~hot).flatten()])
        label = "hottest %d first (mass %.3f)" % (K, float(csum[K - 1]) / nnz)
    newpos = torch.empty(n, dtype=torch.int64, device="cuda")
    newpos[perm] = torch.arange(n, device="cuda")
    col2 = newpos[col.long()].to(torch.int32)
    xp = x[perm].contiguous()
    y = run(col2, xp, label, perm)
    err = float((y.double() - y0.double()).abs().max())
    print("   max |y - y0| = %.3e" % err, flush=True)
    del col2, xp, newpos, perm, hot
