#!/bin/bash
# end-of-round ncu --set full captures: the SpMV kernel on C2 (double) and on the same matrix in double complex, and the
# column-major csrmm kernel; summaries are produced on the build host with tools/ncu_summary.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_row_blocks -s 20 -c 1 -o gpurun_out/final_c2 -f \
   python bench.py --steps 20 --warmup 5 --workload c2 --no-cpu-baseline > gpurun_out/ncu_final_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_row_blocks -s 10 -c 1 -o gpurun_out/final_z -f \
   python tools/z_sweep.py z > gpurun_out/ncu_final_z.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csrmm_col_major -s 3 -c 1 -o gpurun_out/final_mmcol -f \
   python tools/mm_one.py d 32 col > gpurun_out/ncu_final_mmcol.log 2>&1
ls -la gpurun_out/final_*.ncu-rep
