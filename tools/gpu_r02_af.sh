#!/bin/bash
# round 2, run AF (4 GPUs): c5 on 4 GPUs, final state (entry-coded shards, one fused step kernel per iteration), driver protocol
mkdir -p gpurun_out
: > gpurun_out/r02_af.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 >> gpurun_out/r02_af.jsonl 2>> gpurun_out/r02_af.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_af.jsonl'):
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j.get('impl','b200'), j['value'], j['ms_per_step'], j['n_gpus'], (j.get('roofline') or {}).get('frac'), j.get('parity'), (j.get('rank_alone_ms') or {}).get('per_rank'), j['gpu_launches'])
PY
tail -3 gpurun_out/r02_af.err
