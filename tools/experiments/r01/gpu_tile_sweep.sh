#!/bin/bash
cd "$(dirname "$0")/.."
for t in 512 384 256 192 128; do for nt in 128 256; do
  AOCLSPARSE_B200_MM_TILE_NNZ=$t AOCLSPARSE_B200_MM_TILE_THREADS=$nt timeout 300 python tools/mm_one.py d 32 2>&1 | tail -1
done; done
AOCLSPARSE_B200_MM_TILES=0 timeout 300 python tools/mm_one.py d 32 2>&1 | tail -1
