#!/bin/bash
# round 2, run AA (1 GPU): config 3 with the hottest columns renumbered to the front (tools/experiments/c3_hotfirst.py)
mkdir -p gpurun_out
timeout 60 python tools/experiments/c3_hotfirst.py 18 > gpurun_out/r02_aa_c3_hotfirst_s18.txt 2>&1
tail -3 gpurun_out/r02_aa_c3_hotfirst_s18.txt
timeout 400 python tools/experiments/c3_hotfirst.py 24 > gpurun_out/r02_aa_c3_hotfirst.txt 2>&1
cat gpurun_out/r02_aa_c3_hotfirst.txt | tail -30
