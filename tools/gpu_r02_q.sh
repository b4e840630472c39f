#!/bin/bash
# round 2, run Q (4 GPUs): c5q, one system-scope fence per side (GPU-scope release per boundary CTA) vs one per boundary CTA
mkdir -p gpurun_out
: > gpurun_out/r02_q.jsonl
run() { echo "## $1" >> gpurun_out/r02_q.jsonl; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload c5q --gpus 4 --steps 100 --warmup 10 >> gpurun_out/r02_q.jsonl 2>> gpurun_out/r02_q.err; }
run "c5q 4gpu persistent fence=gpu+last sys" X=1
run "c5q 4gpu persistent fence=sys per CTA" AOCLSPARSE_B200_SHARD_FENCE=0
run "c5q 4gpu per-step fence=gpu+last sys" BENCH_SHARD_BATCH=0
python - <<'PY'
import json
for ln in open('gpurun_out/r02_q.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['n_gpus'], j['gpu_launches'], (j.get('rank_alone_ms') or {}).get('per_rank'), (j.get('parity') or {}).get('mismatching_entries_max_over_ranks'), (j.get('parity') or {}).get('iterations'))
PY
tail -4 gpurun_out/r02_q.err
