#!/bin/bash
# round 2, run AB (2 GPUs): persistent entry-coded sharded kernel, rows per block (occupancy with two code buffers),
# one launch per iteration, and the cost of the exchange (local stores, no flags)
mkdir -p gpurun_out
: > gpurun_out/r02_ab.jsonl
run() { echo "## $1" >> gpurun_out/r02_ab.jsonl; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline >> gpurun_out/r02_ab.jsonl 2>> gpurun_out/r02_ab.err; }
run "c5 2gpu default" X=1
run "c5 2gpu R=1664" AOCLSPARSE_B200_BLOCK_ROWS=1664
run "c5 2gpu R=1024" AOCLSPARSE_B200_BLOCK_ROWS=1024
run "c5 2gpu R=512" AOCLSPARSE_B200_BLOCK_ROWS=512
run "c5 2gpu one launch per iteration" BENCH_SHARD_BATCH=0
run "c5 2gpu R=1664 local stores, no flags" AOCLSPARSE_B200_BLOCK_ROWS=1664 AOCLSPARSE_B200_SHARD_DEBUG=3
python - <<'PY'
import json
for ln in open('gpurun_out/r02_ab.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['gpu_launches'], (j.get('rank_alone_ms') or {}).get('per_rank'), (j.get('parity') or {}).get('mismatching_entries_max_over_ranks'), j['config']['plan'].get('entry_plan'))
PY
tail -4 gpurun_out/r02_ab.err
