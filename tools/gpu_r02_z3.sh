#!/bin/bash
# round 2, run Z3 (1 GPU): whole GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_z3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests_z3.log
tail -6 gpurun_out/r02_tests_z3.log
