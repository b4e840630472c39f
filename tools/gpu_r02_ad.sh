#!/bin/bash
# round 2, run AD (1 GPU): whole GPU suite, default bench (c5 + configs c1..c4), reference arm, ncu launch list of the
# default bench command -- state after the pair table and the step-kernel dispatch of entry-coded shards
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_ad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_ad.log
tail -4 gpurun_out/r02_tests_ad.log
( time timeout 600 python bench.py > gpurun_out/r02_ad_bench.json 2> gpurun_out/r02_ad_bench.err ) 2>> gpurun_out/r02_ad_bench.err
( time timeout 400 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_ad_ref.json 2> gpurun_out/r02_ad_ref.err ) 2>> gpurun_out/r02_ad_ref.err
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/r02_ad_bench.json').read().strip().splitlines()[-1])
    print('c5', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline'].get('streamed_frac'), 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'cpu', (j.get('cpu_baseline') or {}).get('value'))
    for k, v in (j.get('configs') or {}).items():
        if v.get('value') is None:
            print(k, v); continue
        print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline'].get('streamed_frac'), v['roofline'].get('kernel'), 'e2e', v['e2e']['value'], v['e2e']['ms_per_step'], 'cpu', (v.get('cpu_baseline') or {}).get('value'))
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -3 gpurun_out/r02_ad_bench.err
tail -c 400 gpurun_out/r02_ad_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ad_launches.log 2>&1
tail -2 gpurun_out/r02_ad_launches.log | cut -c1-300
