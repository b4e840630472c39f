#!/bin/bash
# round 2, run Z (1 GPU): entry-code copy (1 byte per stored entry for matrices with <= 256 distinct (col - row, value) pairs): A/B parity,
# c1 / c2 / c5 with and without, block-rows / threads sweep
mkdir -p gpurun_out
timeout 600 python tools/vc_ab.py > gpurun_out/r02_vc_ab.txt 2>&1; tail -15 gpurun_out/r02_vc_ab.txt
: > gpurun_out/r02_z.jsonl
run() { echo "## $1" >> gpurun_out/r02_z.jsonl; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 50 --warmup 5 >> gpurun_out/r02_z.jsonl 2>> gpurun_out/r02_z.err; }
for w in c1 c2 c5; do
  run "$w ENTRY_CODES=0" $w AOCLSPARSE_B200_ENTRY_CODES=0
  run "$w ENTRY_CODES=1" $w X=1
  for r in 256 768 1024 1536 2048; do run "$w R=$r" $w AOCLSPARSE_B200_BLOCK_ROWS=$r; done
  run "$w NT=128" $w AOCLSPARSE_B200_THREADS=128
  run "$w NT=128 R=256" $w AOCLSPARSE_B200_THREADS=128 AOCLSPARSE_B200_BLOCK_ROWS=256
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_z.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['config']['plan'], 'e2e', j['e2e']['ms_per_step'])
PY
tail -5 gpurun_out/r02_z.err
