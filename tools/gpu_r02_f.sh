#!/bin/bash
# round 2, run F: hot-table pipeline kernel with gather teams: correctness + sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "hot_table or rmat" > gpurun_out/r02_tests_f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_f.log
tail -4 gpurun_out/r02_tests_f.log
: > gpurun_out/r02_f.jsonl
run() { echo "## $1" >> gpurun_out/r02_f.jsonl; shift; env "$@" timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/r02_f.jsonl 2>> gpurun_out/r02_f.err; }
run "teams2 st5 T2048" X=1
run "teams2 st4 T2048" AOCLSPARSE_B200_HOT_STAGES=4
run "teams2 st6 T2048" AOCLSPARSE_B200_HOT_STAGES=6
run "teams3 st6 T2048" AOCLSPARSE_B200_HOT_TEAMS=3 AOCLSPARSE_B200_HOT_STAGES=6
run "teams3 st5 T1536" AOCLSPARSE_B200_HOT_TEAMS=3 AOCLSPARSE_B200_HOT_STAGES=5 AOCLSPARSE_B200_BLOCK_NNZ=1536
run "teams1 st3 T3072" AOCLSPARSE_B200_HOT_TEAMS=1 AOCLSPARSE_B200_HOT_STAGES=3 AOCLSPARSE_B200_BLOCK_NNZ=3072
run "teams2 st5 T3072" AOCLSPARSE_B200_BLOCK_NNZ=3072
run "teams4 st7 T1024" AOCLSPARSE_B200_HOT_TEAMS=4 AOCLSPARSE_B200_HOT_STAGES=7 AOCLSPARSE_B200_BLOCK_NNZ=1024
python - <<'PY'
import json
for ln in open('gpurun_out/r02_f.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        pl=j['config']['plan']
        print(j['value'], j['ms_per_step'], j['roofline']['frac'], pl['block_nnz'], pl['hot_table_entries'], pl['hot_table_mass'])
PY
tail -5 gpurun_out/r02_f.err
