#!/bin/bash
# round 2, run O: persistent double-buffered csrmm tile kernel: tests, c4 with 4x4 / 2x8 threads, boxes, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mesh_tiles_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q -k "tile or csrmm or mm" > gpurun_out/r02_tests_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_o.log
tail -5 gpurun_out/r02_tests_o.log
: > gpurun_out/r02_o.jsonl
run() { echo "## $1" >> gpurun_out/r02_o.jsonl; shift; env "$@" timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/r02_o.jsonl 2>> gpurun_out/r02_o.err; }
run "c4 4x4 default box" X=1
run "c4 box 8,4,3 forced" AOCLSPARSE_B200_MM_TILE_BOX=8,4,3
run "c4 box 8,4,2" AOCLSPARSE_B200_MM_TILE_BOX=8,4,2
run "c4 box 8,2,2" AOCLSPARSE_B200_MM_TILE_BOX=8,2,2
run "c4 box 8,3,4" AOCLSPARSE_B200_MM_TILE_BOX=8,3,4
run "c4 box 8,2,6" AOCLSPARSE_B200_MM_TILE_BOX=8,2,6
run "c4 box 16,2,3" AOCLSPARSE_B200_MM_TILE_BOX=16,2,3
python - <<'PY'
import json
for ln in open('gpurun_out/r02_o.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['roofline']['frac'])
PY
tail -5 gpurun_out/r02_o.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csrmm_mesh_tiles -s 3 -c 1 -o gpurun_out/r02_c4_tiles_p -f python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_o_ncu.log 2>&1
