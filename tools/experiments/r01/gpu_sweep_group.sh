#!/bin/bash
# csrmm row-grouped kernel sweep (C4): group size x vectors per lane x staged entries per block
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/sweep_group.txt
: > $out
run() {
  echo "== $*" >> $out
  env "$@" python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline 2>>gpurun_out/sweep_group.err | python -c "
import sys, json
for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        j = json.loads(l); print(j['ms_per_step'], j['value'], j['roofline']['achieved'])" >> $out
}
python -m pytest tests/test_parity_gpu.py -x -q -k "csrmm or mm_sweep or kat_mm or csc" 2>&1 | tail -3 >> $out
run AOCLSPARSE_B200_MM_GROUP=0
for k in 4 2; do
  for nv in 2 1; do
    run AOCLSPARSE_B200_MM_GROUP=$k AOCLSPARSE_B200_MM_GROUP_NV=$nv
    for t in 512 768 1024 1280; do
      run AOCLSPARSE_B200_MM_GROUP=$k AOCLSPARSE_B200_MM_GROUP_NV=$nv AOCLSPARSE_B200_MM_GROUP_NNZ=$t
    done
  done
done
cat $out
