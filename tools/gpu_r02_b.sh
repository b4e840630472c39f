#!/bin/bash
# round 2, run B: diagonal-code copy A/B on c1 / c2 / c5, block-size sweep with codes, row-load microbenchmark
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "diag_codes or plan_bit_exact or fused_push or config or host_vectors" > gpurun_out/r02_tests_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_b.log
tail -3 gpurun_out/r02_tests_b.log
: > gpurun_out/r02_dcc.jsonl
for w in c1 c2 c5; do
  for dc in 1 0; do
    echo "## $w DIAG_CODES=$dc" >> gpurun_out/r02_dcc.jsonl
    AOCLSPARSE_B200_DIAG_CODES=$dc timeout 300 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline >> gpurun_out/r02_dcc.jsonl 2>> gpurun_out/r02_dcc.err
  done
done
for w in c1 c2 c5; do
  for t in 1536 2560 3072 4096; do
    echo "## $w T=$t coded" >> gpurun_out/r02_dcc.jsonl
    AOCLSPARSE_B200_BLOCK_NNZ=$t timeout 300 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu-baseline >> gpurun_out/r02_dcc.jsonl 2>> gpurun_out/r02_dcc.err
  done
done
timeout 300 tools/bin/microbench_rowload > gpurun_out/r02_microbench_rowload.txt 2>&1
cat gpurun_out/r02_microbench_rowload.txt
python - <<'PY'
import json
for ln in open('gpurun_out/r02_dcc.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln); print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline'].get('isolated_launch_ms'), j['config']['plan']['block_nnz'], 'e2e', j['e2e']['ms_per_step'], j['e2e']['pageable']['ms_per_step'])
PY
