#!/bin/bash
# round 2, run J (2 GPUs): persistent k-iteration sharded kernel -- C test (2 shards / 2 GPUs, bitwise vs one shard), bench with
# all iterations in one call vs one call per step, 100-iteration bitwise parity against the un-fused path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "shard or multi_gpu or sharded" > gpurun_out/r02_tests_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_j.log
tail -6 gpurun_out/r02_tests_j.log
: > gpurun_out/r02_j.jsonl
for b in 1 0; do
  echo "## c5 2gpu shard-c-abi batch=$b" >> gpurun_out/r02_j.jsonl
  BENCH_SHARD_BATCH=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 >> gpurun_out/r02_j.jsonl 2>> gpurun_out/r02_j.err
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_j.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j.get('host_clock_ms_per_step'), j['n_gpus'], j['roofline']['frac'], j['gpu_launches'], 'e2e', j['e2e']['ms_per_step'], j.get('parity'))
PY
tail -8 gpurun_out/r02_j.err
