#!/bin/bash
cd "$(dirname "$0")/.."
for p in s d c z; do for th in 128 256; do
  AOCLSPARSE_B200_THREADS=$th python tools/z_sweep.py $p 2>&1 | tail -1
done; done
AOCLSPARSE_B200_THREADS=128 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('c5 threads=128', j['value'], j['ms_per_step'])"
AOCLSPARSE_B200_THREADS=128 python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('c1 threads=128', j['value'], j['ms_per_step'])"
AOCLSPARSE_B200_THREADS=128 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('c3 threads=128', j['value'], j['ms_per_step'])"
