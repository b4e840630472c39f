#!/bin/bash
# round 2, run R (2 GPUs): whole GPU suite on 2 GPUs, default bench (all five configs), c5 on 2 GPUs, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_r.log
tail -4 gpurun_out/r02_tests_r.log
: > gpurun_out/r02_r.jsonl
echo "## default n=1" >> gpurun_out/r02_r.jsonl
( time timeout 900 python bench.py >> gpurun_out/r02_r.jsonl 2>> gpurun_out/r02_r.err ) 2>> gpurun_out/r02_r.err
echo "## c5 2gpu" >> gpurun_out/r02_r.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 >> gpurun_out/r02_r.jsonl 2>> gpurun_out/r02_r.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_r.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j.get('host_clock_ms_per_step'), j['n_gpus'], j['roofline']['frac'], 'e2e', j['e2e']['ms_per_step'], (j.get('parity') or {}).get('max_err'), (j.get('rank_alone_ms') or {}).get('per_rank'))
        for k,v in (j.get('configs') or {}).items():
            print('   ', k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline'].get('streamed_frac'), v['roofline'].get('kernel'), 'e2e', v['e2e']['ms_per_step'])
PY
tail -8 gpurun_out/r02_r.err
