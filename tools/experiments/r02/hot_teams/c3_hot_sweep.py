#!/usr/bin/env python
"""config 3 (R-MAT, float): the row-block kernel against the hot-column-table kernel (csrc/hot.cu) on ONE matrix built once:
block size x table entries x team size, each result compared bit for bit with the plain kernel's.

    python tools/c3_hot_sweep.py [scale=24] [steps=30]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import capi  # noqa: E402

if __name__ == "__main__":
    import torch
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    lib = capi.AoclSparse()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    lib.set_stream(stream.cuda_stream)
    wl = dict(bench.WORKLOADS["c3"], rmat=scale)
    m, n, nnz, rp, col, val = bench.device_matrix(lib, wl)
    st, A = lib.create_csr("s", 0, m, n, nnz, rp.data_ptr(), col.data_ptr(), val.data_ptr())
    assert st == 0
    del rp, col, val
    torch.cuda.empty_cache()
    d = lib.create_descr()
    os.environ["AOCLSPARSE_B200_HOT"] = "0"
    assert lib.set_mv_hint(A, 111, d, 1000) == 0 and lib.optimize(A) == 0
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    lib.lib.aoclsparse_b200_gen_uniform(1, 0, n, 4, x.data_ptr())
    y = torch.empty(m, dtype=torch.float32, device="cuda")
    g_bytes, g_flops = bench.spmv_bytes_flops(m, n, nnz, 4, False)
    peak, _ = bench.measured_peak()

    def run(label):
        for _ in range(3):
            assert lib.mv("s", 111, 1.0, A, d, x.data_ptr(), 0.0, y.data_ptr()) == 0, lib.last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            lib.mv("s", 111, 1.0, A, d, x.data_ptr(), 0.0, y.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        info = lib.matrix_info(A)
        return ms, info

    ref = None
    block_sizes = [int(v) for v in os.environ.get("SWEEP_T", "768,512,384,1024").split(",")]
    tables = [int(v) for v in os.environ.get("SWEEP_TABLE", "0,6144,8192,12288,16384,20480,24576").split(",")]
    teams = [int(v) for v in os.environ.get("SWEEP_TEAM", "64,128,256").split(",")]
    for T in block_sizes:
        os.environ["AOCLSPARSE_B200_BLOCK_NNZ"] = str(T)
        assert lib.set_row_cuts(A, []) == 0  # invalidates the plan: the next multiply rebuilds it with this block size
        ref = None  # the summation order depends on the block size: compare within one block size
        for entries in tables:
            for team in (teams if entries else [0]):
                if team and T > 768:
                    continue  # a thread of the team gathers all of its (at most 8) entries of a block at once
                assert lib.set_hot_table(A, entries, team) == 0, lib.last_error()
                ms, info = run("")
                if ref is None:
                    ref = y.clone()
                same = bool(torch.equal(y.view(torch.int32), ref.view(torch.int32)))
                err = float((y.double() - ref.double()).abs().max().item())
                print(f"scale {scale} T={info.block_nnz:5d} blocks={info.n_blocks:7d} table={info.hot_entries:6d} mass={info.hot_mass_ppm / 1e4:5.2f}% "
                      f"team={team:3d}: {ms:.4f} ms  {g_flops / ms / 1e6:7.1f} GFLOP/s  {g_bytes / ms / 1e6 / peak:.4f} of measured HBM peak  "
                      f"bitwise==plain: {same} max|diff| {err:.2e}", flush=True)
    lib.destroy(A)
