#!/bin/bash
# round 2, run Z4 (1 GPU): entry-coded kernel, pair table {value, offset} in one 16-byte load vs two tables (A/B of two builds)
mkdir -p gpurun_out
: > gpurun_out/r02_z4.jsonl
run() { echo "## $1" >> gpurun_out/r02_z4.jsonl; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 50 --warmup 5 >> gpurun_out/r02_z4.jsonl 2>> gpurun_out/r02_z4.err; }
for w in c5 c2 c1; do
  run "$w pair" $w AOCLSPARSE_B200_LIB=/root/repo/tools/lib_pair.so.bin
  run "$w split" $w AOCLSPARSE_B200_LIB=/root/repo/tools/lib_split.so.bin
  run "$w pair again" $w AOCLSPARSE_B200_LIB=/root/repo/tools/lib_pair.so.bin
done
python - <<'PY'
import json
for ln in open('gpurun_out/r02_z4.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['roofline']['frac'])
PY
tail -3 gpurun_out/r02_z4.err
