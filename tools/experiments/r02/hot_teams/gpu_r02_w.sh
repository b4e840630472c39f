#!/bin/bash
# round 2, run W (1 GPU): hot-column-table kernel v3 (software-pipelined teams: gathers of block i+1 fly while block i is summed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "hot_column or config3" > gpurun_out/r02_tests_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_w.log
tail -4 gpurun_out/r02_tests_w.log
timeout 900 python tools/c3_hot_sweep.py 24 20 > gpurun_out/r02_c3_hot_sweep_v3.txt 2> gpurun_out/r02_c3_hot_sweep_v3.err
cat gpurun_out/r02_c3_hot_sweep_v3.txt; tail -5 gpurun_out/r02_c3_hot_sweep_v3.err
SWEEP_T=512 SWEEP_TABLE=12288 SWEEP_TEAM=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_hot_teams -s 3 -c 1 \
   -o gpurun_out/r02_ncu_c3_hot_v3 -f python tools/c3_hot_sweep.py 24 3 > gpurun_out/r02_v_ncu_c3_hot.log 2>&1
ls -la gpurun_out/*.ncu-rep
