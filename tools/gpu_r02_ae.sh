#!/bin/bash
# round 2, run AE (2 GPUs): multi-GPU tests (C shard program on all three matrix streams) + the itsol tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_itsol_gpu.py -x -q -m gpu > gpurun_out/r02_tests_ae.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_ae.log
tail -15 gpurun_out/r02_tests_ae.log
