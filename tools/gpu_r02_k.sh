#!/bin/bash
# round 2, run K (8 GPUs): c5 on 8 GPUs, persistent k-iteration kernel vs one launch per iteration; the 4-GPU point
mkdir -p gpurun_out
: > gpurun_out/r02_k.jsonl
for b in 1 0; do
  echo "## c5 8gpu shard-c-abi batch=$b" >> gpurun_out/r02_k.jsonl
  BENCH_SHARD_BATCH=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 10 >> gpurun_out/r02_k.jsonl 2>> gpurun_out/r02_k.err
done
echo "## c5 4gpu shard-c-abi batch=1" >> gpurun_out/r02_k.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 100 --warmup 10 >> gpurun_out/r02_k.jsonl 2>> gpurun_out/r02_k.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_k.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j.get('host_clock_ms_per_step'), j['n_gpus'], j['roofline']['frac'], j['gpu_launches'], 'e2e', j['e2e']['ms_per_step'], (j.get('parity') or {}).get('max_err'), (j.get('parity') or {}).get('iterations'))
PY
tail -4 gpurun_out/r02_k.err
