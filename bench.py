#!/usr/bin/env python
"""bench.py -- CSR SpMV / SpMM throughput of libaoclsparse_b200.so on B200, one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

A "step" is one pass of the hot path (one aoclsparse_?mv / aoclsparse_dcsrmm call) over the synthetic
matrix of a BASELINE.json configuration (SURVEY.md section 8(d) defines the inputs):

  c1  2D 5-point Laplacian 1000^2, double, y = A x + 0.5 y
  c2  3D 27-point stencil 128^3, double, set_mv_hint + optimize, y = A x
  c3  R-MAT scale 24 edge factor 16, float, y = A x
  c4  csrmm: c2's matrix times a dense 2 097 152 x 32 row-major block, double
  c5  3D 7-point stencil 512^3, double, rows sharded over the N GPUs, x_{k+1} = A x_k / 12 with a halo
      exchange per step                                                     <- the line's `value` at every N

BASELINE.json quotes its metric "at 1/2/4/8 B200", i.e. on c5 (it fits one GPU: 11.3 GB of CSR), so c5 is what
`value` measures at N = 1 as well: the 1 -> 8 series is one workload.  With no --workload at N = 1 the other four
configurations are measured in the same run and carried in the `configs` object (each with its own value,
ms_per_step, roofline, e2e and cpu_baseline).

Keys follow the driver contract: `value` = whole-job GFLOP/s with operands resident in HBM, timed with
CUDA events, max over ranks; `e2e` = the same metric through the C ABI with HOST x / y (pinned; the pageable
figure beside it), the host<->device copies inside the timed region; `roofline` = algorithmic bytes per launch
(the reference's own byte model, tests/include/aoclsparse_gbyte.hpp:39-86) / measured launch duration against
the measured HBM copy bandwidth; `cpu_baseline` = the reference's own CPU implementation (oracle/_ref, built
from the reference's sources) timed on this box's host cores in a child process with the OpenMP environment of
SURVEY.md section 8(d).  --impl reference prints that CPU run as its own line.  At N > 1 the timed loop is followed
by a bitwise check of the fused multiply + halo-push kernel against the un-fused aoclsparse_dmv + NCCL halo path
(`parity`).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "aocl-sparse_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "CSR dmv GFLOP/s + effective HBM GB/s (% of 8 TB/s) at 1/2/4/8 B200"
WORKLOADS = {
    "c1": dict(name="c1: 2D 5-point Laplacian 1000^2, CSR double, y=A*x+0.5*y (aoclsparse_dmv)", kind="mv",
               stencil=(5, 1000, 1000, 1), prefix="d", alpha=1.0, beta=0.5),
    "c2": dict(name="c2: 3D 27-point stencil 128^3, CSR double, set_mv_hint+optimize, y=A*x (aoclsparse_dmv)", kind="mv",
               stencil=(27, 128, 128, 128), prefix="d", alpha=1.0, beta=0.0),
    "c3": dict(name="c3: R-MAT scale 24 edgefactor 16, CSR float, y=A*x (aoclsparse_smv)", kind="mv", rmat=24,
               prefix="s", alpha=1.0, beta=0.0),
    "c4": dict(name="c4: csrmm, 3D 27-point 128^3 CSR double x dense 2097152x32 row-major (aoclsparse_dcsrmm)", kind="mm",
               stencil=(27, 128, 128, 128), prefix="d", alpha=1.0, beta=0.0, n_rhs=32),
    "c5": dict(name="c5: 3D 7-point stencil 512^3, CSR double, row-sharded, x<-A*x/12 iterated, halo exchange",
               kind="mv", stencil=(7, 512, 512, 512), prefix="d", alpha=1.0 / 12.0, beta=0.0, sharded=True),
}
# diagnostic only (not a BASELINE configuration): the slab one of 8 ranks owns, alone on one GPU, un-sharded
WORKLOADS["c5slab8"] = dict(name="diagnostic: 512x512x64 slab of c5's grid on one GPU", kind="mv", stencil=(7, 512, 512, 64),
                            prefix="d", alpha=1.0 / 12.0, beta=0.0)
# diagnostic only: c5's grid with 64 planes per rank at ANY world size (what one of 8 ranks owns at the full size)
WORKLOADS["c5q"] = dict(name="diagnostic: 7-point stencil 512x512x(64*ranks), row-sharded, 64 planes per rank", kind="mv",
                        stencil=(7, 512, 512, 64), prefix="d", alpha=1.0 / 12.0, beta=0.0, sharded=True, planes_per_rank=64)
ELEM = {"s": 4, "d": 8, "c": 8, "z": 16}
L2_BYTES = 126 * 1000 * 1000
C5_SAMPLE_PLANES = 64  # CPU arm on c5: the slab one of 8 ranks owns (1/8 of the matrix, 16.8 M rows, 117 M entries)


def spmv_bytes_flops(m, n, nnz, elem, beta_nonzero):
    """tests/include/aoclsparse_gbyte.hpp:39-45, aoclsparse_flops.hpp:30-34 of the reference"""
    b = (m + 1 + nnz) * 4 + (m + n + nnz + (m if beta_nonzero else 0)) * elem
    f = 2 * nnz + (m if beta_nonzero else 0)
    return b, f


def spmm_bytes_flops(m, k, n, nnz, elem, beta_nonzero):
    """tests/include/aoclsparse_gbyte.hpp:76-86, aoclsparse_flops.hpp:49-58 of the reference"""
    b = (m + 1) * 4 + nnz * 4 + (nnz + k * n + m * n + (m * n if beta_nonzero else 0)) * elem
    f = 2 * nnz * n + (m * n if beta_nonzero else 0)
    return b, f


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md); MEASURED_PEAKS.json absent"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------------------------------
def host_topology():
    """(cpu model, {socket: [one hardware thread per physical core]}) restricted to the CPUs this process may use"""
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    allowed = sorted(os.sched_getaffinity(0))
    sockets = {}
    seen = set()
    for c in allowed:
        base = f"/sys/devices/system/cpu/cpu{c}/topology"
        try:
            pkg = int(open(f"{base}/physical_package_id").read())
            core = int(open(f"{base}/core_id").read())
        except (OSError, ValueError):
            pkg, core = 0, c
        if (pkg, core) in seen:
            continue
        seen.add((pkg, core))
        sockets.setdefault(pkg, []).append(c)
    return model, sockets, len(allowed)


def reference_env():
    """OpenMP environment of SURVEY.md 8(d) for the reference's kernels: all physical cores of one socket, bound.
    Returns (env updates, cpus to pin the process to, description)."""
    model, sockets, logical = host_topology()
    pkg = min(sockets) if sockets else 0
    cpus = sockets.get(pkg, sorted(os.sched_getaffinity(0)))
    nt = len(cpus)
    env = {"OMP_NUM_THREADS": str(nt), "AOCLSPARSE_NUM_THREADS": str(nt), "OMP_PROC_BIND": "close",
           "OMP_PLACES": "cores", "OMP_DYNAMIC": "false"}
    desc = {"cpu_model": model, "sockets": len(sockets) or 1, "physical_cores_socket0": nt, "logical_cpus": logical}
    return env, cpus, desc


def host_matrix(wl):
    """numpy CSR of the workload (a slab of it for c5 / a smaller scale for c3), plus a description"""
    import gen_np
    if "stencil" in wl:
        pts, nx, ny, nz = wl["stencil"]
        if wl.get("sharded") and nz > C5_SAMPLE_PLANES:
            lo = (nz // 2) * nx * ny
            hi = lo + C5_SAMPLE_PLANES * nx * ny
            rp, col, val = gen_np.stencil(pts, nx, ny, nz, lo, hi)
            return rp, col, val, hi - lo, nx * ny * nz, (f"SAMPLE: rows [{lo},{hi}) ({C5_SAMPLE_PLANES} of {nz} grid planes, the slab "
                                                        f"one of {nz // C5_SAMPLE_PLANES} ranks owns) of the matrix, whole x")
        rp, col, val = gen_np.stencil(pts, nx, ny, nz)
        return rp, col, val, len(rp) - 1, len(rp) - 1, "the full matrix"
    scale = min(wl["rmat"], 20)
    rp, col, val = gen_np.rmat_csr(scale)
    return rp, col, val, len(rp) - 1, len(rp) - 1, f"SAMPLE: R-MAT scale {scale} (same generator, fewer vertices)"


def run_reference_cpu(wl, steps, warmup):
    """times oracle/_ref/libaoclsparse_ref.so (the reference compiled from its own sources); falls back to
    the scalar oracle port only if that file is absent.  Returns (gflops, ms_per_step, info dict).  Runs inside the
    child process prepared by reference_child(): pinned to the physical cores of one socket before the arrays are
    first touched, OpenMP environment set before libgomp loads."""
    import capi
    import gen_np
    import oracle_py
    p = wl["prefix"]
    dt = {"s": np.float32, "d": np.float64}[p]
    rp, col, val, m, n, sample = host_matrix(wl)
    nnz = len(col)
    x = gen_np.uniform(1, 0, n, dt)
    if wl["kind"] == "mv":
        _, flops = spmv_bytes_flops(m, n, nnz, ELEM[p], wl["beta"] != 0)
    else:
        _, flops = spmm_bytes_flops(m, n, wl["n_rhs"], nnz, ELEM[p], wl["beta"] != 0)

    def timed(call, w, k):
        ts = []
        for i in range(w + k):
            t0 = time.perf_counter()
            assert call() == 0
            if i >= w:
                ts.append(time.perf_counter() - t0)
        return ts
    extra = {}
    if os.path.exists(oracle_py.REF_SO):
        kind = "reference"
        ref = capi.AoclSparse(oracle_py.REF_SO)
        nt = C.c_int32(0)
        isa, tl, arch, upd = C.create_string_buffer(32), C.create_string_buffer(32), C.create_string_buffer(32), C.c_bool(False)
        ref.lib.aoclsparse_debug_get(isa, C.byref(nt), tl, C.byref(upd), arch)  # what the reference itself will use
        threads = int(nt.value)
        t0 = time.perf_counter()
        st, h = ref.create_csr(p, 0, m, n, nnz, rp, col, val)
        extra["create_ms"] = round((time.perf_counter() - t0) * 1e3, 3)  # the serial O(nnz) validation scan
        assert st == 0, st
        d = ref.create_descr()
        if wl["kind"] == "mv":
            y = np.zeros(m, dt)
            call = lambda: ref.mv(p, 111, wl["alpha"], h, d, x, wl["beta"], y)  # noqa: E731
        else:
            nr = wl["n_rhs"]
            B = gen_np.uniform(3, 0, n * nr, dt)
            Cm = np.zeros(m * nr, dt)
            call = lambda: ref.csrmm(p, 111, wl["alpha"], h, d, 0, B, nr, nr, wl["beta"], Cm, nr)  # noqa: E731
        times = timed(call, warmup, steps)
        note = "no optimize"
        if wl["kind"] == "mv":
            # time hint+optimize as well and report the faster (double: br4 format, float: clean copy)
            t0 = time.perf_counter()
            assert ref.set_mv_hint(h, 111, d, 1000) == 0 and ref.optimize(h) == 0
            extra["optimize_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            t2 = timed(call, warmup, steps)
            extra["ms_no_optimize"] = round(statistics.median(times) * 1e3, 4)
            extra["ms_after_optimize"] = round(statistics.median(t2) * 1e3, 4)
            if statistics.median(t2) < statistics.median(times):
                times, note = t2, "after set_mv_hint+optimize"
        ref.destroy(h)
        extra["isa"] = isa.value.decode(errors="replace")
    else:
        kind, threads, note = "port", 1, "scalar oracle port"
        orc = oracle_py.Oracle()
        y = np.zeros(m, dt)
        times = timed(lambda: orc.csrmv(111, wl["alpha"], m, n, 0, rp, col, val, 0, 0, 0, x, wl["beta"], y) or 0,
                      min(warmup, 1), min(steps, 3))
    t = statistics.median(times)
    info = {"kind": kind, "cores": threads,
            "sample": f"{sample}; median of {len(times)} calls (best {min(times) * 1e3:.3f} ms), {note}", **extra}
    return flops / t / 1e9, t * 1e3, info


def reference_child(workload, steps, warmup, one_thread=False, timeout=900):
    """runs `bench.py --impl reference` on one workload in a child process whose OpenMP environment and CPU affinity
    are set BEFORE anything loads libgomp (torch.distributed.run exports OMP_NUM_THREADS=1; torch itself loads an
    OpenMP runtime) and returns the parsed JSON line"""
    env = dict(os.environ)
    upd, cpus, desc = reference_env()
    env.update(upd)
    if one_thread:
        env.update({"OMP_NUM_THREADS": "1", "AOCLSPARSE_NUM_THREADS": "1"})
        cpus = cpus[:1]
    env["BENCH_REF_CHILD"] = ",".join(str(c) for c in cpus)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
                          "--steps", str(steps), "--warmup", str(warmup)], capture_output=True, text=True, env=env,
                         timeout=timeout)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    if out.returncode != 0 or not lines:
        raise RuntimeError(f"reference child failed rc={out.returncode}: {out.stderr[-800:]}")
    return json.loads(lines[-1])


def cpu_baseline_for(workload):
    """cpu_baseline object of a GPU line: the reference on the physical cores of one socket + a 1-thread figure"""
    try:
        j = reference_child(workload, steps=10, warmup=3)
        cpu = dict(j["cpu_baseline"])
        cpu["ms_per_step"] = j["ms_per_step"]
        try:
            j1 = reference_child(workload, steps=3, warmup=1, one_thread=True)
            cpu["one_thread"] = {"value": j1["value"], "ms_per_step": j1["ms_per_step"]}
        except Exception as ex:
            cpu["one_thread"] = {"value": None, "error": repr(ex)[:200]}
        return cpu
    except Exception as ex:  # the baseline is a reported number, not a gate
        return {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex!r}"[:400]}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def device_matrix(lib, wl, row_lo=None, row_hi=None):
    """CSR of the workload (rows [row_lo,row_hi) for the sharded one) generated in device memory"""
    import torch
    if "stencil" in wl:
        pts, nx, ny, nz = wl["stencil"]
        total = nx * ny * nz
        lo = 0 if row_lo is None else row_lo
        hi = total if row_hi is None else row_hi
        nnz = C.c_longlong(0)
        assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, lo, hi, C.byref(nnz), None, None, None) == 0
        rp = torch.empty(hi - lo + 1, dtype=torch.int32, device="cuda")
        col = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
        val = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
        assert lib.lib.aoclsparse_b200_gen_stencil(pts, nx, ny, nz, lo, hi, C.byref(nnz), rp.data_ptr(), col.data_ptr(),
                                                   val.data_ptr()) == 0, lib.last_error()
        return hi - lo, total, nnz.value, rp, col, val
    scale = wl["rmat"]
    n = 1 << scale
    nedges = 16 * n
    keys = torch.empty(nedges, dtype=torch.int64, device="cuda")
    assert lib.lib.aoclsparse_b200_gen_rmat_keys(20240, scale, 0, nedges, keys.data_ptr()) == 0
    keys = torch.unique(keys)  # sort + de-duplicate (plumbing; not on the multiply path)
    nnz = keys.numel()
    rp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    col = torch.empty(nnz, dtype=torch.int32, device="cuda")
    val = torch.empty(nnz, dtype=torch.float32, device="cuda")
    assert lib.lib.aoclsparse_b200_rmat_keys_to_csr(4, scale, nnz, keys.data_ptr(), rp.data_ptr(), col.data_ptr(),
                                                    val.data_ptr()) == 0
    torch.cuda.synchronize()
    del keys
    return n, n, nnz, rp, col, val


def make_shard(lib, A, d, slab, rank, world):
    """aoclsparse_b200_shard_* object for this rank's slab, linked to its neighbours (the 320-byte link records travel
    through torch.distributed; the library itself links no collective)"""
    import torch.distributed as dist
    st, S = lib.shard_create(A, d, rank, world, slab.row_lo, slab.halo)
    assert st == 0, (st, lib.last_error())
    links = [None] * world
    dist.all_gather_object(links, lib.shard_export(S))
    st = lib.shard_connect(S, links[rank - 1] if rank > 0 else None, links[rank + 1] if rank < world - 1 else None)
    assert st == 0, (st, lib.last_error())
    return S


def sharded_parity(lib, wl, slab, A, d, fused_runner, iters):
    """fused multiply + halo push (one launch per iteration, flags in-kernel) against the un-fused path (plain
    aoclsparse_dmv on the windowed handle + NCCL send/recv of the halos) from the same x_0 for `iters` iterations:
    the two own slices must be bit-identical on every rank.  Returns the parity object of the JSON line."""
    import torch
    import torch.distributed as dist

    import sharding
    elem = ELEM[wl["prefix"]]
    m, off, wlen = slab.rows, slab.own_offset, slab.win_hi - slab.win_lo
    alpha = wl["alpha"]
    # un-fused: plain product, halos by NCCL
    bufs = [torch.zeros(wlen, dtype=torch.float64, device="cuda") for _ in range(2)]
    lib.lib.aoclsparse_b200_gen_uniform(1, slab.row_lo, m, elem, bufs[0][off:].data_ptr())
    for r in sharding.exchange_halo(slab, bufs[0]):
        r.wait()
    cur = 0
    for _ in range(iters):
        s = lib.mv("d", 111, alpha, A, d, bufs[cur].data_ptr(), 0.0, bufs[1 - cur][off:].data_ptr())
        assert s == 0, (s, lib.last_error())
        for r in sharding.exchange_halo(slab, bufs[1 - cur]):
            r.wait()
        cur = 1 - cur
    torch.cuda.synchronize()
    plain = bufs[cur][off: off + m]
    # fused: fresh windows and flags
    fused, timed_out = fused_runner(iters)
    mism = int((fused.view(torch.int64) != plain.view(torch.int64)).sum().item())
    err = float((fused - plain).abs().max().item())
    stats = torch.tensor([float(mism), err, float(timed_out)], dtype=torch.float64, device="cuda")
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    nrm = torch.tensor([float((fused * fused).sum().item())], dtype=torch.float64, device="cuda")
    dist.all_reduce(nrm)
    return {"checked": True, "iterations": iters, "mismatching_entries_max_over_ranks": int(stats[0].item()),
            "max_err": float(stats[1].item()), "flag_wait_timed_out": int(stats[2].item()),
            "x_norm2": float(nrm.item()) ** 0.5,
            "what": "fused spmv_sharded_step_kernel (peer stores + in-kernel flags) vs plain aoclsparse_dmv + NCCL halo "
                    "send/recv from the same x_0, own slices compared bitwise on every rank"}


def measure(args, key, primary):
    """one workload on the GPU(s); returns the JSON object of that workload"""
    import torch
    import torch.distributed as dist

    import capi
    import sharding

    wl = WORKLOADS[key]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lib = capi.AoclSparse()
    p = wl["prefix"]
    tdt = torch.float64 if p == "d" else torch.float32
    elem = ELEM[p]
    # one explicit (non-default) stream for everything: the library's calls, torch's helpers and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    lib.set_stream(stream.cuda_stream)
    # the small configurations run tens of microseconds per step: give them enough steps for a stable average
    steps = args.steps if (primary or key == "c3") else max(args.steps, 100)
    warmup = args.warmup if primary else max(args.warmup, 10)

    # ---- matrix, handle, analysis (outside the timed region; their times are reported)
    sharded = bool(wl.get("sharded"))
    if sharded:
        pts, nx, ny, nz = wl["stencil"]
        if wl.get("planes_per_rank"):
            nz = wl["planes_per_rank"] * world
            wl = dict(wl, stencil=(pts, nx, ny, nz))
        plane = nx * ny
        slab = sharding.make_slab(nx * ny * nz, world, rank, halo=plane, granularity=plane)
        m, n_glob, nnz, rp, col, val = device_matrix(lib, wl, slab.row_lo, slab.row_hi)
    else:
        assert world == 1, f"workload {key} does not shard: run it with --gpus 1"
        m, n_glob, nnz, rp, col, val = device_matrix(lib, wl)
        slab = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, A = lib.create_csr(p, 0, m, n_glob, nnz, rp.data_ptr(), col.data_ptr(), val.data_ptr())
    torch.cuda.synchronize()
    create_ms = (time.perf_counter() - t0) * 1e3
    assert st == 0, (st, lib.last_error())
    create_host_ms = None
    if not sharded:
        # the drop-in call: CSR arrays in pageable HOST memory (upload + validation + classification)
        hrp, hcol, hval = rp.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy()
        t0 = time.perf_counter()
        st2, A2 = lib.create_csr(p, 0, m, n_glob, nnz, hrp, hcol, hval)
        torch.cuda.synchronize()
        create_host_ms = (time.perf_counter() - t0) * 1e3
        assert st2 == 0
        lib.destroy(A2)
        del hrp, hcol, hval
    handles = [A]
    del rp, col, val  # the handle owns device copies
    torch.cuda.empty_cache()
    d = lib.create_descr()
    # halo exchange of the sharded workload: shard-c-abi (default) = the library's own C object (csrc/shard.cu): one
    # kernel per iteration that multiplies, stores the boundary rows into the neighbours' x windows over NVLink and
    # hands over the flags; p2p-fused = the same kernel driven from sharding.py; p2p-push = same stores, separate
    # boundary / interior launches and flag kernels; nccl = send/recv baseline
    halo_mode = "none" if (not sharded or world == 1) else os.environ.get("BENCH_HALO", "shard-c-abi")
    shard = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if halo_mode == "shard-c-abi":
        shard = make_shard(lib, A, d, slab, rank, world)  # window, cuts, hint, optimize + windows, flags, peer maps
    else:
        if sharded:
            info0 = lib.matrix_info(A)
            assert sharding.halo_needed(info0.min_col, info0.max_col, slab.row_lo, slab.row_hi) <= slab.halo
            if slab.win_lo != 0 or slab.win_hi != n_glob:
                assert lib.set_x_window(A, slab.win_lo, slab.win_hi) == 0
            cuts = [c for c in (plane, m - plane) if 0 < c < m] if world > 1 else []
            if cuts:
                assert lib.set_row_cuts(A, sorted(set(cuts))) == 0
        assert lib.set_mv_hint(A, 111, d, 1000) == 0 if wl["kind"] == "mv" else lib.set_mm_hint(A, 111, d, 1000) == 0
        assert lib.optimize(A) == 0, lib.last_error()
    torch.cuda.synchronize()
    optimize_ms = (time.perf_counter() - t0) * 1e3
    info = lib.matrix_info(A)

    alpha, beta = wl["alpha"], wl["beta"]
    peer, bufs, n_sets = None, None, 1
    # ---- operands resident in HBM
    if wl["kind"] == "mm":
        nr = wl["n_rhs"]
        B = torch.empty(n_glob * nr, dtype=tdt, device="cuda")
        lib.lib.aoclsparse_b200_gen_uniform(3, 0, n_glob * nr, elem, B.data_ptr())
        Cm = torch.zeros(m * nr, dtype=tdt, device="cuda")

        def step(i):
            s = lib.csrmm(p, 111, alpha, A, d, 0, B.data_ptr(), nr, nr, beta, Cm.data_ptr(), nr)
            assert s == 0, (s, lib.last_error())
        g_bytes, g_flops = spmm_bytes_flops(m, n_glob, nr, nnz, elem, beta != 0)
        g_nnz = nnz
        launches_per_step = 1
    elif not sharded:
        g_bytes, g_flops = spmv_bytes_flops(m, n_glob, nnz, elem, beta != 0)
        g_nnz = nnz
        # cold-L2 protocol: a working set below 2x L2 would be served from the 126 MB L2 on
        # back-to-back launches, so rotate over enough independent (A, x, y) sets to exceed it
        n_sets = 1 if g_bytes > 2 * L2_BYTES else int(-(-3 * L2_BYTES // g_bytes))
        sets = []
        for k in range(n_sets):
            if k == 0:
                Ak = A
            else:
                mk, _, nnzk, rpk, colk, valk = device_matrix(lib, wl)
                stk, Ak = lib.create_csr(p, 0, mk, n_glob, nnzk, rpk.data_ptr(), colk.data_ptr(), valk.data_ptr())
                assert stk == 0
                del rpk, colk, valk
                assert lib.set_mv_hint(Ak, 111, d, 1000) == 0 and lib.optimize(Ak) == 0
                handles.append(Ak)
            xk = torch.empty(n_glob, dtype=tdt, device="cuda")
            lib.lib.aoclsparse_b200_gen_uniform(1, 0, n_glob, elem, xk.data_ptr())
            yk = torch.empty(m, dtype=tdt, device="cuda")
            lib.lib.aoclsparse_b200_gen_uniform(2, 0, m, elem, yk.data_ptr())
            sets.append((Ak, xk, yk))
        x, y = sets[0][1], sets[0][2]

        def step(i):
            Ak, xk, yk = sets[i % n_sets]
            s = lib.mv(p, 111, alpha, Ak, d, xk.data_ptr(), beta, yk.data_ptr())
            assert s == 0, (s, lib.last_error())
        launches_per_step = 1 + (1 if info.n_long_rows else 0)
    else:
        wlen = slab.win_hi - slab.win_lo
        off = slab.own_offset
        if shard is not None:
            lib.lib.aoclsparse_b200_gen_uniform(1, slab.row_lo, m, elem, C.c_void_p(lib.shard_x_ptr(shard)))
            torch.cuda.synchronize()
            assert lib.lib.aoclsparse_b200_shard_publish(shard) == 0, lib.last_error()
        elif halo_mode in ("p2p-push", "p2p-fused"):
            # x windows live in ipc memory; boundary rows store into the neighbours' halos from the kernel epilogue
            peer = sharding.PeerHalo(lib, slab, elem)
            lib.lib.aoclsparse_b200_gen_uniform(1, slab.row_lo, m, elem, C.c_void_p(peer.own_ptr(0)))
            torch.cuda.synchronize()
            dist.barrier()
            peer.initial_push(0)
            dist.barrier()
        else:
            bufs = [torch.zeros(wlen, dtype=tdt, device="cuda") for _ in range(2)]
            lib.lib.aoclsparse_b200_gen_uniform(1, slab.row_lo, m, elem, bufs[0][off:].data_ptr())
        comm_stream = torch.cuda.Stream()
        state = {"cur": 0, "k": 0}
        if world > 1 and peer is None and shard is None:
            for r in sharding.exchange_halo(slab, bufs[0]):
                r.wait()
            torch.cuda.synchronize()
            dist.barrier()

        def step(i):
            if shard is not None:
                s = lib.lib.aoclsparse_b200_shard_iterate(shard, alpha, 1)
                assert s == 0, (s, lib.last_error())
                return
            if peer is not None:
                state["k"] += 1
                if halo_mode == "p2p-fused":
                    s = peer.iteration_fused(state["k"], alpha, A, d)
                    assert s == 0, (s, lib.last_error())
                else:
                    peer.iteration(state["k"], alpha, A, d, beta)
                return
            cur, nxt = bufs[state["cur"]], bufs[1 - state["cur"]]
            ydst = nxt[off:].data_ptr()
            if world == 1:
                s = lib.mv(p, 111, alpha, A, d, cur.data_ptr(), beta, ydst)
                assert s == 0, (s, lib.last_error())
            else:
                # boundary planes first, so their exchange overlaps the interior rows
                for (r0, r1) in ((0, plane), (m - plane, m)):
                    s = lib.mv_rows(p, alpha, A, d, cur.data_ptr(), beta, ydst, r0, r1)
                    assert s == 0, (s, lib.last_error())
                ev = torch.cuda.Event()
                ev.record(stream)
                with torch.cuda.stream(comm_stream):
                    comm_stream.wait_event(ev)
                    reqs = sharding.exchange_halo(slab, nxt)
                s = lib.mv_rows(p, alpha, A, d, cur.data_ptr(), beta, ydst, plane, m - plane)
                assert s == 0, (s, lib.last_error())
                for r in reqs:
                    r.wait()  # makes the current (compute) stream wait for the NCCL work
            state["cur"] = 1 - state["cur"]
        # whole-job algorithmic work of one global SpMV (all ranks together)
        nnz_t = torch.tensor([nnz], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(nnz_t)
        g_nnz = int(nnz_t.item())
        g_bytes, g_flops = spmv_bytes_flops(n_glob, n_glob, g_nnz, elem, beta != 0)
        launches_per_step = 1 if (world == 1 or halo_mode in ("p2p-fused", "shard-c-abi")) else (
            3 if peer is None else 3 + 4 * len(peer.peer) + 2 * len(peer.peer))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then K timed steps bracketed by barrier + synchronize, CUDA events on the launch stream
    # the sharded object iterates k times per call (aoclsparse_b200_shard_iterate: one persistent kernel, grid barrier
    # between iterations); BENCH_SHARD_BATCH=0 calls it once per step instead (one launch per iteration)
    batched = shard is not None and world > 1 and os.environ.get("BENCH_SHARD_BATCH", "1") != "0"

    def run_steps(n):
        if batched:
            s = lib.lib.aoclsparse_b200_shard_iterate(shard, alpha, n)
            assert s == 0, (s, lib.last_error())
        else:
            for i in range(n):
                step(i)
    run_steps(warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    e0.record(stream)
    run_steps(steps)
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3  # host clock around the same region: cross-check of the event time
    dev_ms = e0.elapsed_time(e1)
    if dev_ms < 0.5 * wall_ms and wall_ms > 5.0:
        sys.stderr.write("bench: CUDA-event time %.3f ms is far below the host clock %.3f ms around the same region -- "
                         "work not on the timed stream?\n" % (dev_ms, wall_ms))
    launches = lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps

    tiles = lib.mm_tiles_info(A) if wl["kind"] == "mm" else None

    # ---- N > 1: every rank's slab ALONE (plain aoclsparse_dmv on its window, no flags, no peer stores), all ranks at the
    # same time: separates device-to-device variation and the lock-step with the slowest rank from the cost of the exchange
    rank_alone_ms = None
    if sharded and world > 1 and shard is not None:
        xa = lib.shard_x_ptr(shard) - slab.own_offset * elem
        ya = torch.empty(m, dtype=tdt, device="cuda")
        for _ in range(5):
            assert lib.mv(p, 111, alpha, A, d, xa, beta, ya.data_ptr()) == 0
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(50):
            assert lib.mv(p, 111, alpha, A, d, xa, beta, ya.data_ptr()) == 0
        a1.record(stream)
        barrier()
        t = torch.tensor([a0.elapsed_time(a1) / 50.0], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_alone_ms = [round(float(v.item()), 5) for v in allt]
        del ya

    # ---- launch duration for the roofline: a step is one launch of the dominant kernel (plus, for split rows,
    # the small finish kernel), so the average over the timed region (CUDA events on the launch stream) is the
    # kernel's average launch duration; isolated launches are timed as well and reported beside it
    if sharded and world > 1:
        kern_ms, iso_ms = None, None
    else:
        kern_ms = dev_ms / steps
        ks = []
        for i in range(min(steps, 20)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            step(i)
            b.record(stream)
            b.synchronize()
            ks.append(a.elapsed_time(b))
        iso_ms = statistics.median(ks)

    # ---- end to end through the C ABI with HOST buffers: H2D x, multiply, D2H y per step; page-locked buffers
    # (the headline) and plain pageable ones (what an unmodified caller's malloc'ed arrays are)
    if wl["kind"] == "mv":
        xlen = (slab.win_hi - slab.win_lo) if sharded else n_glob
        hx = torch.empty(xlen, dtype=tdt).pin_memory()
        hy = torch.zeros(m, dtype=tdt).pin_memory()
        if shard is not None:
            assert lib.memcpy(hx.data_ptr(), lib.shard_x_ptr(shard) - off * elem, xlen * elem) == 0
            torch.cuda.synchronize()
        elif sharded and bufs is None and peer is not None:
            assert lib.memcpy(hx.data_ptr(), peer.w_ptr[0], xlen * elem) == 0
            torch.cuda.synchronize()
        else:
            hx.copy_(bufs[0] if sharded else x)
        px, py = hx.numpy().copy(), np.zeros(m, hx.numpy().dtype)
        calls = {"pinned": (lambda: lib.mv(p, 111, alpha, A, d, hx.data_ptr(), beta, hy.data_ptr())),
                 "pageable": (lambda: lib.mv(p, 111, alpha, A, d, px, beta, py))}
        h2d, d2h = xlen * elem + (m * elem if beta != 0 else 0), m * elem
    else:
        hB = torch.empty(n_glob * nr, dtype=tdt).pin_memory()
        hB.copy_(B)
        hC = torch.zeros(m * nr, dtype=tdt).pin_memory()
        pB, pC = hB.numpy().copy(), np.zeros(m * nr, hB.numpy().dtype)
        calls = {"pinned": (lambda: lib.csrmm(p, 111, alpha, A, d, 0, hB.data_ptr(), nr, nr, beta, hC.data_ptr(), nr)),
                 "pageable": (lambda: lib.csrmm(p, 111, alpha, A, d, 0, pB, nr, nr, beta, pC, nr))}
        h2d, d2h = (n_glob * nr + (m * nr if beta != 0 else 0)) * elem, m * nr * elem
    e2e_res = {}
    for name, call in calls.items():
        e2e_steps = max(3, min(steps, 20)) if name == "pinned" else max(3, min(steps, 5))
        for _ in range(2):
            assert call() == 0, lib.last_error()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            assert call() == 0
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_res[name] = (float(t.item()), e2e_steps)
    e2e_ms = e2e_res["pinned"][0]
    e2e = {"value": round(g_flops / (e2e_ms * 1e-3) / 1e9, 3), "unit": "GFLOP/s", "ms_per_step": round(e2e_ms, 4),
           "steps": e2e_res["pinned"][1], "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "pageable": {"value": round(g_flops / (e2e_res["pageable"][0] * 1e-3) / 1e9, 3),
                        "ms_per_step": round(e2e_res["pageable"][0], 4), "steps": e2e_res["pageable"][1],
                        "how": "same call, x / y (B / C) in plain malloc'ed (numpy) host memory"},
           "how": "aoclsparse_?mv / csrmm called with page-locked HOST x,y (B,C): chunked H2D / kernel / D2H pipeline "
                  "on three streams inside the call, result on the host when it returns"}

    # ---- N > 1: bitwise check of the fused path against the un-fused one (both from x_0, >= 100 iterations)
    parity = None
    if sharded and world > 1 and halo_mode in ("p2p-fused", "shard-c-abi"):
        def fused_runner(iters):
            out = torch.empty(m, dtype=torch.float64, device="cuda")
            if halo_mode == "shard-c-abi":
                dist.barrier()  # every rank is done with the timed shard before the windows are refilled
                lib.lib.aoclsparse_b200_gen_uniform(1, slab.row_lo, m, elem, C.c_void_p(lib.shard_x_ptr(shard)))
                torch.cuda.synchronize()
                dist.barrier()
                assert lib.lib.aoclsparse_b200_shard_publish(shard) == 0
                assert lib.lib.aoclsparse_b200_shard_iterate(shard, alpha, iters) == 0, lib.last_error()
                st = lib.lib.aoclsparse_b200_shard_get_x(shard, C.c_void_p(out.data_ptr()))
                return out, 0 if st == 0 else 1
            peer2 = sharding.PeerHalo(lib, slab, elem)
            lib.lib.aoclsparse_b200_gen_uniform(1, slab.row_lo, m, elem, C.c_void_p(peer2.own_ptr(0)))
            torch.cuda.synchronize()
            dist.barrier()
            peer2.initial_push(0)
            dist.barrier()
            for k in range(1, iters + 1):
                s = peer2.iteration_fused(k, alpha, A, d)
                assert s == 0, (s, lib.last_error())
            torch.cuda.synchronize()
            dist.barrier()
            assert lib.memcpy(out.data_ptr(), peer2.own_ptr(iters % 2), m * elem) == 0
            torch.cuda.synchronize()
            return out, peer2.timed_out()
        parity = sharded_parity(lib, wl, slab, A, d, fused_runner, max(100, steps))
    if shard is not None:
        lib.lib.aoclsparse_b200_shard_synchronize(shard)
        if world > 1:
            dist.barrier()
        lib.lib.aoclsparse_b200_shard_destroy(C.byref(shard))

    for hdl in handles:
        lib.destroy(hdl)
    lib.destroy_descr(d)
    if rank != 0:
        return None

    peak, peak_src = measured_peak()
    value = g_flops / (ms_per_step * 1e-3) / 1e9
    eff_gbs = g_bytes / (ms_per_step * 1e-3) / 1e9
    if sharded:
        l_bytes, _ = spmv_bytes_flops(m, slab.win_hi - slab.win_lo, nnz, elem, beta != 0)
    else:
        l_bytes = g_bytes
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get(key)
        traffic_src = ("static: profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                       "capture of this kernel, " + str(tj.get("_captured", "round 1")) + "); not re-measured in this run")
    roof = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src, "traffic": traffic,
            "traffic_source": traffic_src,
            "kernel": (("spmv_sharded_iterate_kernel" if (batched and launches < steps) else "spmv_sharded_step_kernel")
                       if (sharded and world > 1 and halo_mode in ("p2p-fused", "shard-c-abi"))
                       else "spmv_row_blocks_kernel") if wl["kind"] == "mv" else
            ("csrmm_mesh_tiles_kernel" if (tiles and tiles["state"] == 2) else "csrmm_row_major_vec_kernel"),
            "algorithmic_bytes_per_launch": int(l_bytes)}
    if kern_ms:
        roof["achieved"] = round(l_bytes / (kern_ms * 1e-3) / 1e9, 1)
        roof["launch_ms"] = round(kern_ms, 5)
        roof["isolated_launch_ms"] = round(iso_ms, 5)
    else:
        roof["achieved"] = round(eff_gbs / world, 1)
        roof["launch_ms"] = None
        roof["note"] = "per-GPU share of the step (interior + boundary CTAs of one launch overlap the halo exchange)"
    roof["frac"] = round(roof["achieved"] / peak, 4)
    roof["frac_of_nominal_8TBs"] = round(roof["achieved"] / 8000.0, 4)
    if tiles and tiles["state"] == 2:
        roof["tiles"] = {k: tiles[k] for k in ("box", "strides", "rows_per_tile", "rows_per_group", "n_tiles", "max_distinct",
                                               "max_walk", "max_vals", "reuse", "fill")}
        roof["tiles"]["what"] = ("box tiles of the grid (csrc/mesh_tiles.cu): a tile's distinct B rows are staged once in shared "
                                 "memory; entries are stored as 16-bit slots + values re-ordered per row pair, "
                                 "%.1f bytes per stored entry instead of 12" %
                                 ((tiles["walk_entries"] * 4 + tiles["val_entries"] * elem) / max(1, nnz)))
    if info.n_diag_codes > 0 and wl["kind"] == "mv":
        # disclosure: aoclsparse_optimize built the diagonal-code copy of col_idx (1 byte per entry instead of 4, the
        # decoded columns are bit-identical), so the kernel STREAMS fewer bytes than the reference's byte model counts;
        # `achieved` / `frac` stay on the algorithmic bytes (they may exceed 1), `streamed_frac` is the honest
        # utilisation of HBM
        # ... and, for matrices with at most 256 distinct (col - row, value) pairs (constant-coefficient stencils), the
        # entry-code copy: ONE byte per stored entry instead of 4 + sizeof(value); decoded columns and values are
        # bit-identical
        entry_coded = info.n_entry_codes > 0
        streamed = l_bytes - (3 + elem if entry_coded else 3) * nnz
        roof["streamed_bytes_per_launch"] = int(streamed)
        roof["streamed_frac"] = round(streamed / (l_bytes / roof["achieved"]) / peak, 4)
        roof["compression"] = (("entry-code stream: %d distinct (col-row offset, value) pairs, 1 byte per stored entry instead of %d"
                                % (info.n_entry_codes, 4 + elem)) if entry_coded else
                               ("diagonal-code column stream: %d distinct col-row offsets, 1 index byte per entry"
                                % info.n_diag_codes))

    out = {
        "metric": METRIC,
        "value": round(value, 3), "unit": "GFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
        "host_clock_ms_per_step": round(wall_ms / steps, 5),
        "scaling": "strong" if sharded else "weak", "vs_baseline": None,
        "dtype": {"s": "f32", "d": "f64"}[p], "data": "synthetic",
        "config": {"workload": wl["name"], "rows": int(n_glob if sharded else m), "nnz": int(g_nnz),
                   "parallelism": ((f"row slabs x{world}, halo: {halo_mode}"
                                    + ((", all timed iterations in ONE aoclsparse_b200_shard_iterate call ("
                                        + ("persistent kernel, grid barrier between iterations" if launches < steps else
                                           "entry-coded shard: one fused step kernel per iteration, launched back to back "
                                           "by the call, no host synchronisation") + ")") if batched else ""))
                                   if sharded else "single GPU"),
                   "l2": ("operands larger than L2: %.0f MB streamed per step vs 126 MB L2" % (l_bytes / 1e6))
                   if (sharded or wl["kind"] == "mm" or n_sets == 1) else
                   ("rotating over %d independent (A,x,y) sets, %.0f MB in total vs 126 MB L2" % (n_sets, n_sets * l_bytes / 1e6)),
                   "plan": {"block_nnz": info.block_nnz, "blocks": info.n_blocks, "thread": info.n_thread_blocks,
                            "warp": info.n_warp_blocks, "product": info.n_product_blocks,
                            "long_segments": info.n_long_segments, "long_rows": info.n_long_rows,
                            "diag_code_table": info.n_diag_codes, "entry_code_table": info.n_entry_codes,
                            "entry_plan": ({"block_nnz": info.e_block_nnz, "block_rows": info.e_block_rows, "blocks": info.e_n_blocks}
                                           if info.n_entry_codes > 0 else None)},
                   "optimize_ms": round(optimize_ms, 2),
                   "create_ms": round(create_ms, 2),
                   "create_from_pageable_host_ms": None if create_host_ms is None else round(create_host_ms, 2)},
        "effective_gbs": round(eff_gbs, 1), "effective_frac_of_8TBs": round(eff_gbs / (8000.0 * world), 4),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "gpu_launches_per_step": (round(launches / steps, 4) if batched else launches_per_step), "roofline": roof,
    }
    if parity is not None:
        out["parity"] = parity
    if rank_alone_ms is not None:
        out["rank_alone_ms"] = {"per_rank": rank_alone_ms, "max": max(rank_alone_ms),
                                "what": "each rank's slab alone (plain aoclsparse_dmv on its window, no exchange), all ranks "
                                        "concurrently: the lock-step iteration cannot be faster than the slowest of these"}
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    primary = args.workload or "c5"
    out = measure(args, primary, True)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_for(primary)
        else:
            out["cpu_baseline"] = None
        if args.workload is None and world == 1:
            # the other BASELINE.json configurations, same protocol, in the same run
            out["configs"] = {}
            for key in ("c1", "c2", "c3", "c4"):
                try:
                    o = measure(args, key, False)
                    sub = {k: o[k] for k in ("value", "unit", "ms_per_step", "steps", "warmup", "dtype", "effective_gbs",
                                             "effective_frac_of_8TBs", "roofline", "e2e", "gpu_launches", "config")}
                    sub["cpu_baseline"] = None if args.no_cpu_baseline else cpu_baseline_for(key)
                    out["configs"][key] = sub
                    out["gpu_launches"] += o["gpu_launches"]
                except Exception as ex:  # an extra line must never take the primary one down
                    out["configs"][key] = {"value": None, "error": repr(ex)[:400]}
                torch.cuda.empty_cache()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    key = args.workload or "c5"
    wl = WORKLOADS[key]
    if "BENCH_REF_CHILD" not in os.environ:
        # re-run in a child whose OpenMP environment / affinity is right from the first instruction
        j = reference_child(key, args.steps, args.warmup)
        j["n_gpus"] = args.gpus
        try:
            j1 = reference_child(key, steps=3, warmup=1, one_thread=True)
            j["cpu_baseline"]["one_thread"] = {"value": j1["value"], "ms_per_step": j1["ms_per_step"]}
        except Exception as ex:
            j["cpu_baseline"]["one_thread"] = {"value": None, "error": repr(ex)[:200]}
        print(json.dumps(j), flush=True)
        return
    cpus = [int(c) for c in os.environ["BENCH_REF_CHILD"].split(",") if c]
    _, _, desc = reference_env()
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)  # before the arrays are first touched: pages land next to the threads
        except OSError:
            pass
    gf, ms, info = run_reference_cpu(wl, steps=args.steps, warmup=args.warmup)
    info.update({"host": desc, "omp": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "AOCLSPARSE_NUM_THREADS",
                                                                      "OMP_PROC_BIND", "OMP_PLACES")},
                 "pinned_to_cpus": len(cpus),
                 "first_touch": "process pinned to the cores its OpenMP threads run on before the arrays are generated"})
    out = {
        "impl": "reference",
        "metric": METRIC,
        "value": round(gf, 3), "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong" if wl.get("sharded") else "weak",
        "vs_baseline": None, "dtype": {"s": "f32", "d": "f64"}[wl["prefix"]], "data": "synthetic",
        "config": {"workload": wl["name"], "sample": info["sample"].split(";")[0]},
        "cpu_baseline": dict(value=round(gf, 3), unit="GFLOP/s", **info),
        "e2e": {"value": round(gf, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
