#!/bin/bash
# gpurun --timeout 120 -- 'bash scripts/experiments/run_gather4_test.sh'
mkdir -p gpurun_out
cd scripts/experiments
timeout 60 nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather4_shift_test gather4_shift_test.cu -lcuda 2>&1 | tail -5
(timeout 20 /tmp/gather4_shift_test 2>&1 | tail -20) | tee ../../gpurun_out/gather4_shift_test.log
