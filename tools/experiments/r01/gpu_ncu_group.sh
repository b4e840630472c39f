#!/bin/bash
# one ncu capture each of the row-grouped csrmm kernel (K=4 and K=2, one vector per lane) on C4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in 4 2; do
  AOCLSPARSE_B200_MM_GROUP=$k AOCLSPARSE_B200_MM_GROUP_NV=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:csrmm_grouped -s 3 -c 1 -o gpurun_out/group_k$k -f python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_group_k$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
