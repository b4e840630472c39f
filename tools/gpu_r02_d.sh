#!/bin/bash
# round 2, run D: hot-table pipeline kernel (c3) correctness + sweep, the c5 slab alone, reference samples, shard C test
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "hot_table or rmat or reference_sample or shard_c_abi or config3" > gpurun_out/r02_tests_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_d.log
tail -15 gpurun_out/r02_tests_d.log
: > gpurun_out/r02_d.jsonl
run() { echo "## $1" >> gpurun_out/r02_d.jsonl; shift; env "$@" timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/r02_d.jsonl 2>> gpurun_out/r02_d.err; }
run "c3 hot default" X=1
run "c3 plain" AOCLSPARSE_B200_HOT=0
run "c3 hot stages=2" AOCLSPARSE_B200_HOT_STAGES=2
run "c3 hot stages=4" AOCLSPARSE_B200_HOT_STAGES=4
run "c3 hot T=2048" AOCLSPARSE_B200_BLOCK_NNZ=2048
run "c3 hot T=4096" AOCLSPARSE_B200_BLOCK_NNZ=4096
run "c3 hot table=24576" AOCLSPARSE_B200_HOT_TABLE=24576
run "c3 hot table=16384" AOCLSPARSE_B200_HOT_TABLE=16384
echo "## c5slab8" >> gpurun_out/r02_d.jsonl
timeout 300 python bench.py --workload c5slab8 --steps 100 --warmup 10 --no-cpu-baseline >> gpurun_out/r02_d.jsonl 2>> gpurun_out/r02_d.err
python - <<'PY'
import json
for ln in open('gpurun_out/r02_d.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline'].get('streamed_frac'), j['config']['plan'], 'opt', j['config']['optimize_ms'])
PY
tail -5 gpurun_out/r02_d.err
