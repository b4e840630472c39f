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
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
