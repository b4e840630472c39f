#!/bin/bash
# round 2, run C (2 GPUs): full GPU test suite, coded block-size sweep, c5 on 1 and 2 GPUs through the C-ABI shard object
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_c.log
tail -4 gpurun_out/r02_tests_c.log
: > gpurun_out/r02_c.jsonl
for w in c2 c5; do
  for t in 2304 2560 2816 2944; do
    echo "## $w T=$t coded" >> gpurun_out/r02_c.jsonl
    AOCLSPARSE_B200_BLOCK_NNZ=$t timeout 300 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline >> gpurun_out/r02_c.jsonl 2>> gpurun_out/r02_c.err
  done
done
echo "## c1 default" >> gpurun_out/r02_c.jsonl
timeout 300 python bench.py --workload c1 --steps 200 --warmup 20 --no-cpu-baseline >> gpurun_out/r02_c.jsonl 2>> gpurun_out/r02_c.err
for mode in shard-c-abi p2p-fused; do
  echo "## c5 2gpu $mode" >> gpurun_out/r02_c.jsonl
  BENCH_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 >> gpurun_out/r02_c.jsonl 2>> gpurun_out/r02_c.err
done
echo "## reference arm n=2" >> gpurun_out/r02_c.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 >> gpurun_out/r02_c.jsonl 2>> gpurun_out/r02_c.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_c.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        if j.get('impl') == 'reference':
            print('REF', j['value'], j['ms_per_step'], j['cpu_baseline']['cores'], j['cpu_baseline'].get('one_thread')); continue
        print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline'].get('streamed_frac'), j['config']['plan']['block_nnz'], 'e2e', j['e2e']['ms_per_step'], j.get('parity'))
PY
tail -5 gpurun_out/r02_c.err
