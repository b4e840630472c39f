#!/bin/bash
# end-of-round evidence: launch list of the default bench command under ncu, and clean (unprofiled) bench lines of all
# single-GPU workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_default.csv \
   python bench.py --no-cpu-baseline > gpurun_out/launches_default.log 2>&1
: > gpurun_out/bench_final.jsonl
python bench.py >> gpurun_out/bench_final.jsonl 2>gpurun_out/bench_final.err
for w in c1 c3 c4 c5; do
  python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_final.jsonl 2>>gpurun_out/bench_final.err
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_final.jsonl'):
    if l.startswith('{'):
        j = json.loads(l)
        print(j['config']['workload'][:40], j['value'], j['ms_per_step'], j['roofline']['frac'], (j.get('e2e') or {}).get('value'))
PY
