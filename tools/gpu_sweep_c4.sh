#!/bin/bash
# csrmm (config 4) sweep: lanes per row, non-zeros in flight, block size (run under gpurun)
for T in 1024 2048 4096; do for LPR in 8 16; do for U in 2 4; do
  r=$(AOCLSPARSE_B200_BLOCK_NNZ=$T AOCLSPARSE_B200_MM_LPR=$LPR AOCLSPARSE_B200_MM_UNROLL=$U python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
  echo "T=$T LPR=$LPR U=$U $(echo $r | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["effective_gbs"])' 2>/dev/null)"
done; done; done
