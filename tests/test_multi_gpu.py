"""GPU test needing >= 2 devices: spawns tests/multi_gpu_check.py under torch.distributed.run."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_row_sharded_iteration_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("stream", ["entry-coded", "diagonal-coded", "plain"])
@pytest.mark.parametrize("world", [2, 3])
def test_shard_c_abi_from_plain_c(tmp_path, world, stream):
    """tests/shard_check.c: the row-sharded iteration through aoclsparse_b200_shard_* from a C program, no Python and no
    collective library in the loop -- `world` shards in one process, on `world` GPUs when the box has them (else they
    share a device; the flag protocol is the same).  The program compares with a one-shard run bit for bit and with the
    host recurrence to 1e-12 * iterations, and returns non-zero on mismatch.
    `stream`: the matrix stream the shards multiply from -- the entry-code copy (one fused step kernel per iteration),
    or, with that copy switched off, the diagonal-code copy / the plain 32-bit column stream (both iterate in the
    persistent cooperative kernel when every shard has a GPU of its own)."""
    import shutil
    cc = shutil.which("gcc") or shutil.which("cc") or "/usr/bin/gcc"
    exe = str(tmp_path / "shard_check")
    libdir = os.path.join(ROOT, "aocl-sparse_b200")
    subprocess.run([cc, "-O2", os.path.join(ROOT, "tests", "shard_check.c"), "-I", os.path.join(ROOT, "include"), "-L", libdir,
                    "-laoclsparse_b200", "-lm", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    env = dict(os.environ)
    if stream != "entry-coded":
        env["AOCLSPARSE_B200_ENTRY_CODES"] = "0"
    if stream == "plain":
        env["AOCLSPARSE_B200_DIAG_CODES"] = "0"
    out = subprocess.run([exe, str(world), "25"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "SHARD_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
