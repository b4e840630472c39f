#!/bin/bash
# round 2, run AC (2 GPUs): c5q (64 planes per rank = what one of 8 ranks owns at full size): persistent kernel vs one
# launch per iteration, entry-coded, with the shard plan rule (>= 24 rounds) and with the single-GPU rule (8 waves)
mkdir -p gpurun_out
: > gpurun_out/r02_ac.jsonl
run() { echo "## $1" >> gpurun_out/r02_ac.jsonl; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --workload c5q --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline >> gpurun_out/r02_ac.jsonl 2>> gpurun_out/r02_ac.err; }
run "c5q 2gpu persistent (R by shard rule)" X=1
run "c5q 2gpu step kernels inside one iterate call (R by shard rule)" AOCLSPARSE_B200_SHARD_PERSISTENT=0
run "c5q 2gpu step kernels, R=1728" AOCLSPARSE_B200_SHARD_PERSISTENT=0 AOCLSPARSE_B200_BLOCK_ROWS=1728
run "c5q 2gpu step kernels, R=1024" AOCLSPARSE_B200_SHARD_PERSISTENT=0 AOCLSPARSE_B200_BLOCK_ROWS=1024
run "c5q 2gpu persistent, R=1024" AOCLSPARSE_B200_BLOCK_ROWS=1024
python - <<'PY'
import json
for ln in open('gpurun_out/r02_ac.jsonl'):
    if ln.startswith('##'): print(ln.strip(), end='  '); continue
    if ln.startswith('{'):
        j=json.loads(ln)
        print(j['value'], j['ms_per_step'], j['gpu_launches'], (j.get('rank_alone_ms') or {}).get('per_rank'), (j.get('parity') or {}).get('mismatching_entries_max_over_ranks'), j['config']['plan'].get('entry_plan'))
PY
tail -4 gpurun_out/r02_ac.err
