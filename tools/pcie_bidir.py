#!/usr/bin/env python
"""PCIe microbenchmark for the e2e analysis: 16 MB host->device alone, device->host alone, and both at once on two
streams (pinned host memory).  Says what the floor of a host-resident aoclsparse_dmv on C2 is."""
import time
import torch

n = 2097152
hx = torch.empty(n, dtype=torch.float64).pin_memory()
hy = torch.empty(n, dtype=torch.float64).pin_memory()
dx = torch.empty(n, dtype=torch.float64, device="cuda")
dy = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    with torch.cuda.stream(s1):
        dx.copy_(hx, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        hy.copy_(dy, non_blocking=True)


def both():
    h2d()
    d2h()


def both_chunked(k=8):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1):
            dx[i * c:(i + 1) * c].copy_(hx[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2):
            hy[i * c:(i + 1) * c].copy_(dy[i * c:(i + 1) * c], non_blocking=True)


a, b, c, d = timeit(h2d), timeit(d2h), timeit(both), timeit(both_chunked)
print(f"16 MB H2D {a:.3f} ms ({16.777/a:.1f} GB/s)  D2H {b:.3f} ms ({16.777/b:.1f} GB/s)  both at once {c:.3f} ms  "
      f"both, 8 chunks each {d:.3f} ms")
