#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "d 32" "s 32" "z 32" "d 64" "d 16" "c 32"; do
  for lpr in 4 8 16; do for nv in 2 4; do
    AOCLSPARSE_B200_MM_LPR=$lpr AOCLSPARSE_B200_MM_NV=$nv python tools/mm_one.py $cfg 2>&1 | tail -1
  done; done
done
