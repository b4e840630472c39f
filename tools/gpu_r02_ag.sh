#!/bin/bash
# round 2, run AG (1 GPU): entry-coded kernel, code bytes read as 32-bit words + funnel shift (new) vs byte loads (prev): A/B of two builds
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "codes" > gpurun_out/r02_tests_ah.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_ah.log
tail -3 gpurun_out/r02_tests_ah.log
: > gpurun_out/r02_ag.jsonl
run() { echo "## $1" >> gpurun_out/r02_ag.jsonl; w=$2; shift; shift; env "$@" timeout 200 python bench.py --workload $w --no-cpu-baseline --steps 50 --warmup 5 >> gpurun_out/r02_ag.jsonl 2>> gpurun_out/r02_ag.err; }
for w in c5 c2 c1; do
  run "$w new" $w X=1
  run "$w prev" $w AOCLSPARSE_B200_LIB=/root/repo/tools/lib_prev.so.bin
  run "$w new again" $w X=1
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_ag.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['roofline']['frac'])
PY
tail -3 gpurun_out/r02_ag.err
