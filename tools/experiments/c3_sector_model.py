"""Offline model of the x-gather sector traffic of config 3 (R-MAT scale S): how many distinct 32-byte sectors a
warp-wide gather touches in storage order, with the entries of a block sorted by column, and per block.
Usage: python c3_sector_model.py [scale] (writes /tmp/c3/keys_<scale>.npy as a cache)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import gen_np  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cache = f"/tmp/c3/keys_{scale}.npy"
t0 = time.time()
if os.path.exists(cache):
    keys = np.load(cache)
else:
    n = 1 << scale
    parts = []
    chunk = 1 << 24
    for first in range(0, 16 * n, chunk):
        parts.append(np.unique(gen_np.rmat_keys(20240, scale, first, chunk)))
    keys = np.unique(np.concatenate(parts))
    del parts
    np.save(cache, keys)
print("nnz", len(keys), "gen s", round(time.time() - t0, 1), flush=True)
col = (keys & 0xFFFFFFFF).astype(np.int32)
row = (keys >> 32).astype(np.int32)
nnz = len(col)
del keys


def distinct_per_group(sec, g=32):
    """sum over groups of g consecutive entries of the number of distinct values (sec need not be sorted)"""
    k = len(sec) // g * g
    s = np.sort(sec[:k].reshape(-1, g), axis=1)
    return int((s[:, 1:] != s[:, :-1]).sum() + s.shape[0])


for shift, name in ((3, "sector(8 floats)"), (5, "line(32 floats)")):
    sec = col >> shift
    base = distinct_per_group(sec)
    print(f"{name}: storage order: {base/ (nnz/32):.2f} per warp request, total {base/1e6:.1f} M", flush=True)
    for T in (2048, 4096, 8192, 16384, 32768):
        k = nnz // T * T
        s = np.sort(sec[:k].reshape(-1, T), axis=1)
        tot_sorted = distinct_per_group(s.reshape(-1))
        tot_block = int((s[:, 1:] != s[:, :-1]).sum() + s.shape[0])
        print(f"  T={T}: sorted-in-block {tot_sorted/(k/32):.2f} per warp request, total {tot_sorted/1e6:.1f} M;"
              f" distinct per block total {tot_block/1e6:.1f} M ({tot_block/k:.3f} per entry)", flush=True)
# hot-column coverage
cnt = np.bincount(col, minlength=1 << scale)
srt = np.sort(cnt)[::-1]
cs = np.cumsum(srt)
for K in (8192, 24576, 49152, 98304, 196608, 393216, 1 << 20):
    print(f"top {K} columns carry {cs[K-1]/nnz:.3f} of the entries")
# by 128-byte line of x (32 floats)
cntl = np.bincount(col >> 5, minlength=1 << (scale - 5))
srt = np.sort(cntl)[::-1]
cs = np.cumsum(srt)
for K in (256, 768, 1024, 2048, 4096, 8192, 16384, 32768):
    print(f"top {K} lines ({K*128//1024} KB) carry {cs[K-1]/nnz:.3f} of the entries")
rl = np.bincount(row, minlength=1 << scale)
print("rows: empty", float((rl == 0).mean()), "max", int(rl.max()), "rows>2048:", int((rl > 2048).sum()), "rows>65536:", int((rl > 65536).sum()))
