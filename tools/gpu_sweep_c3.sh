#!/bin/bash
# C3 (R-MAT) strategy / block-size sweep, run under gpurun
for KID in -1 0 1 2; do for T in 1024 2048 3072 6144; do
  r=$(AOCLSPARSE_B200_FORCE_KID=$KID AOCLSPARSE_B200_BLOCK_NNZ=$T python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
  echo "KID=$KID T=$T $(echo $r | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["config"]["plan"])' 2>/dev/null)"
done; done
