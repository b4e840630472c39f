#!/bin/bash
# round 2, run P (4 GPUs): c5q (64 planes per rank) after halo-only L2 loads + acq_rel fences; persistent vs per-step
mkdir -p gpurun_out
: > gpurun_out/r02_p.jsonl
run() { echo "## $1" >> gpurun_out/r02_p.jsonl; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload c5q --gpus 4 --steps 100 --warmup 10 >> gpurun_out/r02_p.jsonl 2>> gpurun_out/r02_p.err; }
run "c5q 4gpu persistent" X=1
run "c5q 4gpu one launch per iteration" BENCH_SHARD_BATCH=0
run "c5q 4gpu persistent, local stores, no flags" AOCLSPARSE_B200_SHARD_DEBUG=3
python - <<'PY'
import json
for ln in open('gpurun_out/r02_p.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['n_gpus'], j['gpu_launches'], (j.get('rank_alone_ms') or {}).get('per_rank'), (j.get('parity') or {}).get('mismatching_entries_max_over_ranks'))
PY
tail -4 gpurun_out/r02_p.err
