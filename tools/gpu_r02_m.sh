#!/bin/bash
# round 2, run M: csrmm box tiles -- tests, c4 bench with tiles on / off and a few boxes, ncu of the tile kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mesh_tiles_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py -x -q -k "tiles or csrmm or mm" > gpurun_out/r02_tests_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_m.log
tail -15 gpurun_out/r02_tests_m.log
: > gpurun_out/r02_m.jsonl
run() { echo "## $1" >> gpurun_out/r02_m.jsonl; shift; env "$@" timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline >> gpurun_out/r02_m.jsonl 2>> gpurun_out/r02_m.err; }
run "c4 tiles default" X=1
run "c4 tiles off" AOCLSPARSE_B200_MM_TILES=0
run "c4 box 8,4,2" AOCLSPARSE_B200_MM_TILE_BOX=8,4,2
run "c4 box 8,4,4 (1 CTA/SM)" AOCLSPARSE_B200_MM_TILE_BOX=8,4,4 AOCLSPARSE_B200_MM_TILE_SMEM=230000
run "c4 box 8,2,2" AOCLSPARSE_B200_MM_TILE_BOX=8,2,2
run "c4 box 16,2,2" AOCLSPARSE_B200_MM_TILE_BOX=16,2,2
run "c4 box 32,2,1" AOCLSPARSE_B200_MM_TILE_BOX=32,2,1
python - <<'PY'
import json
for ln in open('gpurun_out/r02_m.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['roofline']['frac'], 'e2e', j['e2e']['ms_per_step'])
PY
tail -5 gpurun_out/r02_m.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csrmm_mesh_tiles -s 3 -c 1 -o gpurun_out/r02_c4_tiles -f python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_m_ncu.log 2>&1
ls -la gpurun_out/r02_c4_tiles.ncu-rep
