#!/bin/bash
# round 2, run AI (1 GPU): whole GPU suite + default bench on the final tree (complex CG, word-wise code reads for long rows)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_ai.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_ai.log
tail -4 gpurun_out/r02_tests_ai.log
( time timeout 400 python bench.py > gpurun_out/r02_ai_bench.json 2> gpurun_out/r02_ai_bench.err ) 2>> gpurun_out/r02_ai_bench.err
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/r02_ai_bench.json').read().strip().splitlines()[-1])
    print('c5', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline'].get('streamed_frac'), 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'cpu', (j.get('cpu_baseline') or {}).get('value'))
    for k, v in (j.get('configs') or {}).items():
        if v.get('value') is None:
            print(k, v); continue
        print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline'].get('streamed_frac'), v['roofline'].get('kernel'), 'e2e', v['e2e']['value'], v['e2e']['ms_per_step'], 'cpu', (v.get('cpu_baseline') or {}).get('value'))
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -4 gpurun_out/r02_ai_bench.err
