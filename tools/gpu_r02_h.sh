#!/bin/bash
# round 2, run H: coded kernel with 128-thread CTAs (16 per SM) and smaller blocks on c1 / c2 / c5; c3 with the new skewed block size
mkdir -p gpurun_out
: > gpurun_out/r02_h.jsonl
run() { w=$1; echo "## $w $2" >> gpurun_out/r02_h.jsonl; shift; shift; env "$@" timeout 600 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline >> gpurun_out/r02_h.jsonl 2>> gpurun_out/r02_h.err; }
for w in c1 c2 c5; do
  run $w "default" X=1
  run $w "NT128 T1280" AOCLSPARSE_B200_THREADS=128 AOCLSPARSE_B200_BLOCK_NNZ=1280
  run $w "NT128 T1024" AOCLSPARSE_B200_THREADS=128 AOCLSPARSE_B200_BLOCK_NNZ=1024
  run $w "NT128 T1536" AOCLSPARSE_B200_THREADS=128 AOCLSPARSE_B200_BLOCK_NNZ=1536
  run $w "NT128 T2560" AOCLSPARSE_B200_THREADS=128 AOCLSPARSE_B200_BLOCK_NNZ=2560
done
run c3 "default (T768)" X=1
python - <<'PY'
import json
for ln in open('gpurun_out/r02_h.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        pl=j['config']['plan']
        print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline'].get('streamed_frac'), pl['block_nnz'], pl['blocks'])
PY
tail -3 gpurun_out/r02_h.err
