#!/bin/bash
# round 2, run X (1 GPU): hot-column-table kernel v4 (warp-specialised teams: 3 gather warps + 1 reduce warp, mbarrier hand-over)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "hot_column or config3" > gpurun_out/r02_tests_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_x.log
tail -4 gpurun_out/r02_tests_x.log
SWEEP_TEAM=128 SWEEP_T=768,512,384 timeout 900 python tools/c3_hot_sweep.py 24 20 > gpurun_out/r02_c3_hot_sweep_v4b.txt 2> gpurun_out/r02_c3_hot_sweep_v4b.err
cat gpurun_out/r02_c3_hot_sweep_v4b.txt; tail -5 gpurun_out/r02_c3_hot_sweep_v4b.err
SWEEP_T=768 SWEEP_TABLE=12288 SWEEP_TEAM=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_hot_teams -s 3 -c 1 \
   -o gpurun_out/r02_ncu_c3_hot_v4b -f python tools/c3_hot_sweep.py 24 3 > gpurun_out/r02_v_ncu_c3_hot.log 2>&1
ls -la gpurun_out/*.ncu-rep
