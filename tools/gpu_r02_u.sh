#!/bin/bash
# round 2, run U (1 GPU): gather rate out of a cluster-wide shared-memory table (DSMEM), alone and mixed with L2 gathers
mkdir -p gpurun_out tools/bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/microbench_dsmem_gather tools/microbench_dsmem_gather.cu
timeout 600 tools/bin/microbench_dsmem_gather > gpurun_out/r02_microbench_dsmem_gather.txt 2>&1
cat gpurun_out/r02_microbench_dsmem_gather.txt
