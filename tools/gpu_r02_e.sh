#!/bin/bash
# round 2, run E: ncu capture of the hot-table pipeline kernel on c3 (source-level stalls)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_hot_pipeline -s 2 -c 1 -o gpurun_out/r02_hot -f python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_e.log 2>&1
tail -3 gpurun_out/r02_e.log
ls -la gpurun_out/r02_hot.ncu-rep
