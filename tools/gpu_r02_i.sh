#!/bin/bash
# round 2, run I (2 GPUs): full GPU suite on 2 GPUs, c5 on 2 GPUs through the shard C ABI (event time vs host clock), default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_i.log
tail -4 gpurun_out/r02_tests_i.log
: > gpurun_out/r02_i.jsonl
for mode in shard-c-abi p2p-fused; do
  echo "## c5 2gpu $mode" >> gpurun_out/r02_i.jsonl
  BENCH_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 >> gpurun_out/r02_i.jsonl 2>> gpurun_out/r02_i.err
done
echo "## default n=1" >> gpurun_out/r02_i.jsonl
( time timeout 900 python bench.py >> gpurun_out/r02_i.jsonl 2>> gpurun_out/r02_i.err ) 2>> gpurun_out/r02_i.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_i.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j.get('host_clock_ms_per_step'), j['n_gpus'], j['roofline']['frac'], 'e2e', j['e2e']['ms_per_step'], (j.get('parity') or {}).get('max_err'))
        for k,v in (j.get('configs') or {}).items():
            print('   ', k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline'].get('streamed_frac'), 'e2e', v['e2e']['ms_per_step'])
PY
tail -8 gpurun_out/r02_i.err
