#!/bin/bash
# round 2, run Z2 (1 GPU): whole GPU suite with the entry-code copy in place, default bench, ncu --set full of the
# entry-coded kernel on c5 / c2 / c1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_z2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_z2.log
tail -6 gpurun_out/r02_tests_z2.log
( time timeout 900 python bench.py > gpurun_out/r02_z2_bench.json 2> gpurun_out/r02_z2_bench.err ) 2>> gpurun_out/r02_z2_bench.err
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/r02_z2_bench.json').read().strip().splitlines()[-1])
    print('c5', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline'].get('streamed_frac'), 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'cpu', (j.get('cpu_baseline') or {}).get('value'), j['config']['plan'].get('entry_plan'), 'optimize', j['config']['optimize_ms'])
    for k, v in (j.get('configs') or {}).items():
        if v.get('value') is None:
            print(k, v); continue
        print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline'].get('streamed_frac'), v['roofline'].get('kernel'), 'e2e', v['e2e']['value'], v['e2e']['ms_per_step'], 'cpu', (v.get('cpu_baseline') or {}).get('value'), v['config']['plan'].get('entry_plan'), 'optimize', v['config']['optimize_ms'])
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -3 gpurun_out/r02_z2_bench.err
cap() { # name, kernel regex, skip, workload
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -o "gpurun_out/r02_ncu_ec_$1" -f \
    python bench.py --workload "$4" --steps 3 --warmup 3 --no-cpu-baseline > "gpurun_out/r02_z2_ncu_$1.log" 2>&1
}
cap c5 spmv_row_blocks 3 c5
cap c2 spmv_row_blocks 4 c2
cap c1 spmv_row_blocks 4 c1
ls -la gpurun_out/r02_ncu_ec_*.ncu-rep
