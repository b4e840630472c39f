#!/bin/bash
# persistent pipelined kernel: ring depth / CTAs per SM sweep (run under gpurun)
for w in c1 c2 c5; do
  r=$(python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1)
  echo "$w baseline $(echo $r | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["effective_gbs"])')"
  for ST in 2 3 4 6; do for CT in 1 2 3; do
    r=$(AOCLSPARSE_B200_PIPELINE=1 AOCLSPARSE_B200_PIPE_STAGES=$ST AOCLSPARSE_B200_PIPE_CTAS=$CT python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1)
    echo "$w stages=$ST ctas=$CT $(echo $r | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["effective_gbs"])' 2>/dev/null)"
  done; done
done
