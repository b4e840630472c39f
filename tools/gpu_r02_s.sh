#!/bin/bash
# round 2, run S (1 GPU): whole GPU suite, default bench (c5 + configs c1..c4), reference arm, ncu launch list of the
# default bench command, ncu --set full captures of the dominant kernel of c1 / c2 / c3 / c4 / c5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_s.log
tail -4 gpurun_out/r02_tests_s.log
( time timeout 900 python bench.py > gpurun_out/r02_s_bench.json 2> gpurun_out/r02_s_bench.err ) 2>> gpurun_out/r02_s_bench.err
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_s_ref.json 2> gpurun_out/r02_s_ref.err ) 2>> gpurun_out/r02_s_ref.err
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/r02_s_bench.json').read().strip().splitlines()[-1])
    print('c5', j['value'], j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'cpu', (j.get('cpu_baseline') or {}).get('value'))
    for k, v in (j.get('configs') or {}).items():
        if v.get('value') is None:
            print(k, v); continue
        print(k, v['value'], v['ms_per_step'], v['roofline']['frac'], v['roofline'].get('streamed_frac'), v['roofline'].get('kernel'), 'e2e', v['e2e']['value'], v['e2e']['ms_per_step'], 'cpu', (v.get('cpu_baseline') or {}).get('value'))
except Exception as ex:
    print('bench parse failed', ex)
PY
tail -3 gpurun_out/r02_s_bench.err
tail -c 600 gpurun_out/r02_s_ref.json
# launch list of the default bench command (serialised, cold cache: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s_launches.log 2>&1
cap() { # name, kernel regex, skip, workload
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -o "gpurun_out/r02_ncu_$1" -f \
    python bench.py --workload "$4" --steps 3 --warmup 3 --no-cpu-baseline > "gpurun_out/r02_s_ncu_$1.log" 2>&1
}
cap c4 csrmm_mesh_tiles 3 c4
cap c3 spmv_row_blocks 3 c3
cap c1 spmv_row_blocks 20 c1
cap c2 spmv_row_blocks 10 c2
cap c5 spmv_row_blocks 3 c5
ls -la gpurun_out/*.ncu-rep
