#!/bin/bash
# round 2, run T (1 GPU): hot-column-table kernel (csrc/hot.cu): parity test, sweep on the scale-24 R-MAT matrix,
# ncu --set full of one configuration; c1 / c2 captures of the timed (whole-matrix) launches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "hot_column or config3" > gpurun_out/r02_tests_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_t.log
tail -4 gpurun_out/r02_tests_t.log
timeout 900 python tools/c3_hot_sweep.py 24 30 > gpurun_out/r02_c3_hot_sweep.txt 2> gpurun_out/r02_c3_hot_sweep.err
cat gpurun_out/r02_c3_hot_sweep.txt; tail -5 gpurun_out/r02_c3_hot_sweep.err
SWEEP_T=768 SWEEP_TABLE=16384 SWEEP_TEAM=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_hot_teams -s 3 -c 1 \
   -o gpurun_out/r02_ncu_c3_hot -f python tools/c3_hot_sweep.py 24 3 > gpurun_out/r02_t_ncu_c3_hot.log 2>&1
cap() { # name, kernel regex, skip, workload
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -o "gpurun_out/r02_ncu_$1" -f \
    python bench.py --workload "$4" --steps 3 --warmup 3 --no-cpu-baseline > "gpurun_out/r02_t_ncu_$1.log" 2>&1
}
cap c1 spmv_row_blocks 4 c1
cap c2 spmv_row_blocks 4 c2
ls -la gpurun_out/*.ncu-rep
