#!/bin/bash
# Tuning sweep run on the GPU box under gpurun: one bench line per (block nnz, CTA threads, L2 hint).
# Usage: tools/gpu_sweep.sh <workload> [steps]
wl=${1:-c2}; steps=${2:-50}
out=gpurun_out/sweep_${wl}.jsonl
: > $out
for T in 1024 2048 4096 8192; do for TH in 128 256 512; do for H in 0 1; do
  r=$(AOCLSPARSE_B200_BLOCK_NNZ=$T AOCLSPARSE_B200_THREADS=$TH AOCLSPARSE_B200_L2HINT=$H \
      python bench.py --workload $wl --steps $steps --warmup 5 --no-cpu-baseline 2>&1 | tail -1)
  echo "{\"T\":$T,\"threads\":$TH,\"l2hint\":$H,\"r\":$r}" >> $out
  echo "T=$T TH=$TH H=$H $(echo $r | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["effective_gbs"], d["roofline"]["achieved"], d["roofline"]["frac"])' 2>/dev/null)"
done; done; done
