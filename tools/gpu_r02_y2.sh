#!/bin/bash
# round 2, run Y2 (2 GPUs, after the step-kernel dispatch of entry-coded shards): multi-GPU tests (Python-driven check + the C program driving two GPUs without Python),
# c5 on 2 GPUs (timed iterations in one aoclsparse_b200_shard_iterate call + bitwise parity against the un-fused path),
# reference arm under torchrun
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r02_tests_y2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_y2.log
tail -4 gpurun_out/r02_tests_y2.log
: > gpurun_out/r02_y2.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 >> gpurun_out/r02_y2.jsonl 2>> gpurun_out/r02_y2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 10 --warmup 3 >> gpurun_out/r02_y2.jsonl 2>> gpurun_out/r02_y2.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_y2.jsonl'):
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j.get('impl','b200'), j['value'], j['ms_per_step'], j['n_gpus'], (j.get('roofline') or {}).get('frac'), 'e2e', j['e2e'].get('ms_per_step'), j.get('parity'), (j.get('rank_alone_ms') or {}).get('per_rank'), (j.get('cpu_baseline') or {}).get('cores'))
PY
tail -5 gpurun_out/r02_y2.err
