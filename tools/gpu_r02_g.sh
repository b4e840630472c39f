#!/bin/bash
# round 2, run G: c3 with the row-block kernel: how much shared memory is carved out of L1 decides how many gather
# misses an SM can have in flight (each pending miss holds an L1 line): sweep block size x CTA size
mkdir -p gpurun_out
: > gpurun_out/r02_g.jsonl
run() { echo "## $1" >> gpurun_out/r02_g.jsonl; shift; env AOCLSPARSE_B200_HOT=0 "$@" timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/r02_g.jsonl 2>> gpurun_out/r02_g.err; }
run "T2048 NT256 (base)" X=1
run "T768 NT256" AOCLSPARSE_B200_BLOCK_NNZ=768
run "T896 NT256" AOCLSPARSE_B200_BLOCK_NNZ=896
run "T1280 NT256" AOCLSPARSE_B200_BLOCK_NNZ=1280
run "T1664 NT512" AOCLSPARSE_B200_BLOCK_NNZ=1664 AOCLSPARSE_B200_THREADS=512
run "T1792 NT512" AOCLSPARSE_B200_BLOCK_NNZ=1792 AOCLSPARSE_B200_THREADS=512
run "T2048 NT512" AOCLSPARSE_B200_BLOCK_NNZ=2048 AOCLSPARSE_B200_THREADS=512
run "T3584 NT512" AOCLSPARSE_B200_BLOCK_NNZ=3584 AOCLSPARSE_B200_THREADS=512
run "T832 NT128" AOCLSPARSE_B200_BLOCK_NNZ=832 AOCLSPARSE_B200_THREADS=128
run "T384 NT128" AOCLSPARSE_B200_BLOCK_NNZ=384 AOCLSPARSE_B200_THREADS=128
python - <<'PY'
import json
for ln in open('gpurun_out/r02_g.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        pl=j['config']['plan']
        print(j['value'], j['ms_per_step'], j['roofline']['frac'], pl['block_nnz'], pl['blocks'], pl['long_segments'])
PY
tail -3 gpurun_out/r02_g.err
