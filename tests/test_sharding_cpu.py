"""CPU (gloo, world_size 2 and 3) test of the host logic behind the row-sharded iterated SpMV:
partition, halo sizing, halo exchange, and that the sharded iteration reproduces the un-sharded one.
The multiply itself is stood in for by scipy here; the GPU multiply on a slab with an x window is covered
by tests/test_parity_gpu.py::test_row_sharded_window_and_row_ranges."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, iters, q):
    import scipy.sparse as sp

    import gen_np
    import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nx, ny, nz = 6, 5, 12
    plane, n = nx * ny, nx * ny * nz
    slab = sharding.make_slab(n, world, rank, halo=plane, granularity=plane)
    rp, col, val = gen_np.stencil(7, nx, ny, nz, slab.row_lo, slab.row_hi)
    assert sharding.halo_needed(int(col.min()), int(col.max()), slab.row_lo, slab.row_hi) <= plane
    A = sp.csr_matrix((val, col - slab.win_lo, rp), shape=(slab.rows, slab.win_hi - slab.win_lo))
    cur = torch.zeros(slab.win_hi - slab.win_lo, dtype=torch.float64)
    nxt = torch.zeros_like(cur)
    off = slab.own_offset
    cur[off: off + slab.rows] = torch.from_numpy(gen_np.uniform(1, slab.row_lo, slab.rows))
    for r in sharding.exchange_halo(slab, cur):
        r.wait()
    for _ in range(iters):
        nxt[off: off + slab.rows] = torch.from_numpy((A @ cur.numpy()) / 12.0)
        for r in sharding.exchange_halo(slab, nxt):
            r.wait()
        cur, nxt = nxt, cur
    q.put((rank, slab.row_lo, cur[off: off + slab.rows].numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_iteration_matches_unsharded(world):
    import scipy.sparse as sp

    import gen_np
    iters = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, iters, q)) for r in range(world)]
    [p.start() for p in procs]
    parts = sorted([q.get(timeout=120) for _ in range(world)])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    got = np.concatenate([p[2] for p in parts])
    nx, ny, nz = 6, 5, 12
    rp, col, val = gen_np.stencil(7, nx, ny, nz)
    A = sp.csr_matrix((val, col, rp))
    x = gen_np.uniform(1, 0, nx * ny * nz)
    for _ in range(iters):
        x = (A @ x) / 12.0
    assert np.max(np.abs(got - x)) <= 1e-14 * np.max(np.abs(x)) * 10


def test_partition_properties():
    import sharding
    for n, g in ((512 ** 3, 512 ** 2), (1000, 1), (96, 8)):
        for world in (1, 2, 4, 8):
            edges = [sharding.partition_rows(n, world, r, g) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            assert all(lo % g == 0 and hi % g == 0 for lo, hi in edges)
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= g
    s = sharding.make_slab(4096, 4, 0, 64)
    assert (s.win_lo, s.win_hi, s.has_left, s.has_right, s.own_offset) == (0, 1024 + 64, False, True, 0)
    s = sharding.make_slab(4096, 4, 3, 64)
    assert (s.win_lo, s.win_hi, s.has_left, s.has_right, s.own_offset) == (3072 - 64, 4096, True, False, 64)
    s = sharding.make_slab(4096, 1, 0, 64)
    assert s.halo == 0 and not s.has_left and not s.has_right


def _gather_worker(rank, world, port, iters, q):
    import scipy.sparse as sp

    import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = _skewed_matrix()
    cuts = sharding.partition_by_nnz(A.indptr, world)
    plan = sharding.GatherPlan(cuts, rank, torch.float64, "cpu")
    mine = A[plan.row_lo: plan.row_hi]
    x = torch.from_numpy(np.linspace(-1.0, 1.0, A.shape[0]))
    for _ in range(iters):
        plan.own_slice().copy_(torch.from_numpy(mine @ x.numpy()) / 50.0)
        x = plan.gather(torch.empty_like(x))
    q.put((rank, x.numpy().copy(), cuts))
    dist.barrier()
    dist.destroy_process_group()


def _skewed_matrix():
    """power-law-ish square matrix: a few very long rows, many short ones (deterministic)"""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    n = 400
    deg = np.minimum((n * 0.6 / (1 + np.arange(n)) ** 0.9).astype(int) + 1, n)
    rows = np.repeat(np.arange(n), deg)
    cols = np.concatenate([rng.choice(n, d, replace=False) for d in deg])
    return sp.csr_matrix((rng.normal(size=len(rows)), (rows, cols)), shape=(n, n))


@pytest.mark.parametrize("world", [2, 3])
def test_allgather_mode_with_nnz_balanced_rows(world):
    """general matrices: rows split at equal-nnz points, whole x all-gathered after every multiply"""
    iters = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, iters, q)) for r in range(world)]
    [p.start() for p in procs]
    parts = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    A = _skewed_matrix()
    x = np.linspace(-1.0, 1.0, A.shape[0])
    for _ in range(iters):
        x = (A @ x) / 50.0
    for _, got, cuts in parts:  # every rank ends with the same, complete vector
        assert np.max(np.abs(got - x)) <= 1e-13 * max(1.0, np.max(np.abs(x)))
    cuts = parts[0][2]
    per_rank = np.diff(A.indptr[cuts])
    assert cuts[0] == 0 and cuts[-1] == A.shape[0] and all(a <= b for a, b in zip(cuts, cuts[1:]))
    # balanced to within the longest row
    assert per_rank.max() - per_rank.min() <= 2 * np.diff(A.indptr).max()
    rows_per_rank = np.diff(cuts)
    assert rows_per_rank.max() > 2 * rows_per_rank.min()  # i.e. NOT an equal-rows split


def test_partition_by_nnz_edge_cases():
    import sharding
    assert sharding.partition_by_nnz([0, 0, 0, 0], 2) == [0, 0, 3] or sharding.partition_by_nnz([0, 0, 0, 0], 2)[-1] == 3
    assert sharding.partition_by_nnz([0, 5], 4)[0] == 0 and sharding.partition_by_nnz([0, 5], 4)[-1] == 1
    rp = np.arange(0, 101, 10)
    assert sharding.partition_by_nnz(rp, 5) == [0, 2, 4, 6, 8, 10]
    assert sharding.partition_by_nnz(rp, 1) == [0, 10]


def _transposed_worker(rank, world, port, q):
    import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = _skewed_matrix()
    n = A.shape[0]
    cuts = sharding.partition_by_nnz(A.indptr, world)
    ycuts = [n * r // world for r in range(world + 1)]
    plan = sharding.TransposedPlan(ycuts, rank, torch.float64, "cpu")
    mine = A[cuts[rank]: cuts[rank + 1]]
    x = np.linspace(-1.0, 1.0, n)[cuts[rank]: cuts[rank + 1]]
    plan.partial.copy_(torch.from_numpy(mine.T @ x))
    y = plan.reduce()
    q.put((rank, y.numpy().copy(), ycuts))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_transposed_product_on_row_shards(world):
    """y = A^T x with A and x row-sharded: full-length partial products reduced and scattered"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_transposed_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    parts = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    A = _skewed_matrix()
    want = A.T @ np.linspace(-1.0, 1.0, A.shape[0])
    got = np.concatenate([p[1] for p in parts])
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
