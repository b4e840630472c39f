"""Offline LRU model of one SM's L1 on config 3's x gathers: natural column numbering against columns renumbered by
popularity (hot first).  Lines of 128 bytes filled by 32-byte sectors; a gather hits when line and sector are present.
Usage: python c3_l1_model.py  (needs /tmp/c3/keys_24.npy from c3_sector_model.py)"""
import collections
import sys

import numpy as np

keys = np.load("/tmp/c3/keys_24.npy", mmap_mode="r")
nnz = len(keys)
col_all = None
cnt = np.zeros(1 << 24, np.int64)
step = 1 << 26
for a in range(0, nnz, step):
    c = (np.asarray(keys[a:a + step]) & 0xFFFFFFFF).astype(np.int64)
    cnt += np.bincount(c, minlength=1 << 24)
order = np.argsort(-cnt, kind="stable")
rank = np.empty(1 << 24, np.int64)
rank[order] = np.arange(1 << 24)
print("popularity ranks ready", flush=True)


def simulate(cols, lines_cap):
    cache = collections.OrderedDict()  # line -> sector mask
    hits = 0
    for c in cols:
        ln, sec = c >> 5, 1 << ((c >> 3) & 3)
        m = cache.get(ln)
        if m is not None:
            cache.move_to_end(ln)
            if m & sec:
                hits += 1
            else:
                cache[ln] = m | sec
        else:
            cache[ln] = sec
            if len(cache) > lines_cap:
                cache.popitem(last=False)
    return hits / len(cols)


# the stream one SM sees: 8 resident blocks of T entries, interleaved in warp-sized groups of 32 entries
T = 2048
for start_frac in (0.1, 0.5, 0.9):
    a0 = int(nnz * start_frac) // T * T
    n_rounds = 24
    chunk = (np.asarray(keys[a0:a0 + 8 * T * n_rounds]) & 0xFFFFFFFF).astype(np.int64)
    blocks = chunk.reshape(n_rounds, 8, T // 32, 32)           # round, cta, warp-group, lane
    stream = blocks.transpose(0, 2, 1, 3).reshape(-1)          # within a round the 8 CTAs alternate group by group
    for name, s in (("natural", stream), ("by popularity", rank[stream])):
        for cap in (512, 1024, 1536):
            print(f"start {start_frac}: {name:14s} L1 {cap * 128 // 1024:4d} KB: hit rate {simulate(s.tolist(), cap):.3f}", flush=True)
