"""Launched under torch.distributed.run with N >= 2 GPUs (tests/test_multi_gpu.py does that): a small
config-5-shaped problem (3D 7-point stencil, x <- A x / 12 iterated) row-sharded over the ranks, with the halo
exchanged (a) by NCCL send/recv and (b) by the fused peer push of the multiply kernel; both must reproduce the
un-sharded product computed on the host, to the parity tolerance."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import capi  # noqa: E402
import gen_np  # noqa: E402
import sharding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    lib = capi.AoclSparse()
    lib.set_stream(torch.cuda.current_stream().cuda_stream)
    nx, ny, nz, iters = 48, 40, 16 * world, 7
    plane, n = nx * ny, nx * ny * nz
    slab = sharding.make_slab(n, world, rank, halo=plane, granularity=plane)
    rp, col, val = gen_np.stencil(7, nx, ny, nz, slab.row_lo, slab.row_hi)
    m = slab.rows
    st, A = lib.create_csr("d", 0, m, n, len(col), rp, col, val)
    assert st == 0
    info = lib.matrix_info(A)
    assert sharding.halo_needed(info.min_col, info.max_col, slab.row_lo, slab.row_hi) == plane
    d = lib.create_descr()
    assert lib.set_x_window(A, slab.win_lo, slab.win_hi) == 0
    assert lib.set_row_cuts(A, [plane, m - plane]) == 0
    assert lib.set_mv_hint(A, 111, d, 100) == 0 and lib.optimize(A) == 0
    x0 = gen_np.uniform(1, slab.row_lo, m)
    off, wlen = slab.own_offset, slab.win_hi - slab.win_lo
    results = {}

    # (a) NCCL send/recv halo
    bufs = [torch.zeros(wlen, dtype=torch.float64, device="cuda") for _ in range(2)]
    bufs[0][off: off + m] = torch.from_numpy(x0).cuda()
    for r in sharding.exchange_halo(slab, bufs[0]):
        r.wait()
    cur = 0
    for _ in range(iters):
        assert lib.mv("d", 111, 1.0 / 12, A, d, bufs[cur].data_ptr(), 0.0, bufs[1 - cur][off:].data_ptr()) == 0
        for r in sharding.exchange_halo(slab, bufs[1 - cur]):
            r.wait()
        cur = 1 - cur
    torch.cuda.synchronize()
    results["nccl"] = bufs[cur][off: off + m].cpu().numpy()

    # (b) fused peer push
    peer = sharding.PeerHalo(lib, slab, 8)
    assert lib.memcpy(peer.own_ptr(0), torch.from_numpy(x0).cuda().data_ptr(), m * 8) == 0
    torch.cuda.synchronize()
    dist.barrier()
    peer.initial_push(0)
    dist.barrier()
    for k in range(1, iters + 1):
        peer.iteration(k, 1.0 / 12, A, d, 0.0)
    torch.cuda.synchronize()
    dist.barrier()
    out = torch.empty(m, dtype=torch.float64, device="cuda")
    assert lib.memcpy(out.data_ptr(), peer.own_ptr(iters % 2), m * 8) == 0
    tout = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.memcpy(tout.data_ptr(), peer.timeout_ptr, 4) == 0
    torch.cuda.synchronize()
    assert int(tout.item()) == 0, "a flag wait timed out"
    results["push"] = out.cpu().numpy()

    # (c) the whole iteration in one kernel: multiply + push + flags
    peer2 = sharding.PeerHalo(lib, slab, 8)
    assert lib.memcpy(peer2.own_ptr(0), torch.from_numpy(x0).cuda().data_ptr(), m * 8) == 0
    torch.cuda.synchronize()
    dist.barrier()
    peer2.initial_push(0)
    dist.barrier()
    for k in range(1, iters + 1):
        assert peer2.iteration_fused(k, 1.0 / 12, A, d) == 0, lib.last_error()
    torch.cuda.synchronize()
    dist.barrier()
    assert peer2.timed_out() == 0, "a flag wait timed out"
    out2 = torch.empty(m, dtype=torch.float64, device="cuda")
    assert lib.memcpy(out2.data_ptr(), peer2.own_ptr(iters % 2), m * 8) == 0
    torch.cuda.synchronize()
    results["fused"] = out2.cpu().numpy()

    # un-sharded truth on the host
    import scipy.sparse as sp
    rpg, colg, valg = gen_np.stencil(7, nx, ny, nz)
    Ag = sp.csr_matrix((valg, colg, rpg))
    x = gen_np.uniform(1, 0, n)
    for _ in range(iters):
        x = (Ag @ x) / 12.0
    want = x[slab.row_lo: slab.row_hi]
    scale = np.max(np.abs(want))
    for mode, got in results.items():
        err = float(np.max(np.abs(got - want)) / scale)
        assert err <= 1e-12 * iters, (mode, err)
    assert np.array_equal(results["nccl"], results["push"])  # same arithmetic in the same order: bit-identical
    assert np.array_equal(results["nccl"], results["fused"])
    dist.barrier()
    lib.destroy(A)

    # (d) a matrix without band structure ("general"): rows split at equal-nnz points, global column
    # indices, whole x all-gathered (NCCL) after every multiply; float, skewed row lengths
    rng = np.random.default_rng(11)
    ng = 30000
    deg = np.minimum((ng * 0.3 / (1 + np.arange(ng)) ** 0.8).astype(np.int64) + 2, ng)
    rows = np.repeat(np.arange(ng), deg)
    cols = np.concatenate([rng.choice(ng, int(dg), replace=False) for dg in deg])
    G = sp.csr_matrix((rng.normal(size=len(rows)).astype(np.float32), (rows, cols)), shape=(ng, ng))
    G.sort_indices()
    cuts = sharding.partition_by_nnz(G.indptr, world)
    plan = sharding.GatherPlan(cuts, rank, torch.float32, "cuda")
    mine = G[plan.row_lo: plan.row_hi]
    st, Ag2 = lib.create_csr("s", 0, mine.shape[0], ng, mine.nnz, mine.indptr.astype(np.int32),
                             mine.indices.astype(np.int32), mine.data)
    assert st == 0, lib.last_error()
    assert lib.set_mv_hint(Ag2, 111, d, 100) == 0 and lib.optimize(Ag2) == 0
    xg0 = np.linspace(-1.0, 1.0, ng).astype(np.float32)
    xg = torch.from_numpy(xg0).cuda()
    nxt = torch.empty_like(xg)
    giters = 3
    for _ in range(giters):
        assert lib.mv("s", 111, 0.01, Ag2, d, xg.data_ptr(), 0.0, plan.own_slice().data_ptr()) == 0, lib.last_error()
        plan.gather(nxt)
        xg, nxt = nxt, xg
    torch.cuda.synchronize()
    want = xg0.astype(np.float64)
    den = np.abs(want)
    G64, Gabs = G.astype(np.float64), abs(G).astype(np.float64)
    for _ in range(giters):
        den = 0.01 * (Gabs @ den)
        want = 0.01 * (G64 @ want)
    got = xg.cpu().numpy().astype(np.float64)
    err = float(np.max(np.abs(got - want) / np.where(den > 0, den, 1.0)))
    assert err <= 1e-5 * giters, ("allgather", err)
    per_rank = np.diff(G.indptr[cuts])
    assert per_rank.max() - per_rank.min() <= 2 * int(deg.max())
    # (e) transposed product on the same row shards: full-length partial vectors, NCCL reduce-scatter
    ycuts = [ng * r // world for r in range(world + 1)]
    tplan = sharding.TransposedPlan(ycuts, rank, torch.float32, "cuda")
    x_own = torch.from_numpy(xg0[plan.row_lo: plan.row_hi].copy()).cuda()
    assert lib.mv("s", 112, 1.0, Ag2, d, x_own.data_ptr(), 0.0, tplan.partial.data_ptr()) == 0, lib.last_error()
    y_own = tplan.reduce()
    torch.cuda.synchronize()
    want_t = (G64.T @ xg0.astype(np.float64))[ycuts[rank]: ycuts[rank + 1]]
    den_t = (Gabs.T @ np.abs(xg0.astype(np.float64)))[ycuts[rank]: ycuts[rank + 1]]
    err_t = float(np.max(np.abs(y_own.cpu().numpy().astype(np.float64) - want_t) / np.where(den_t > 0, den_t, 1.0)))
    assert err_t <= 1e-5, ("transposed", err_t)
    lib.destroy(Ag2)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
