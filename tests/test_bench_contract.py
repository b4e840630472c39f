"""CPU test of bench.py's output contract on the arm that needs no GPU (`--impl reference`): one JSON line with the keys
the driver reads, the reference's own build as the thing measured, and rank != 0 staying silent under a multi-rank
launch."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    import oracle_py
    if not os.path.exists(oracle_py.REF_SO):
        pytest.skip("oracle/_ref not built")
    lines = _run(["--impl", "reference", "--workload", "c1", "--steps", "3", "--warmup", "3"])
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["unit"] == "GFLOP/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["value"] > 0 and j["steps"] == 3 and j["n_gpus"] == 1 and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["config"]["workload"].startswith("c1")
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["sample"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]


def test_reference_arm_other_ranks_stay_silent():
    lines = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert lines == []
