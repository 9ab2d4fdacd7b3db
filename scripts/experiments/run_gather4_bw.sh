#!/bin/bash
# gpurun --timeout 300 -- 'bash scripts/experiments/run_gather4_bw.sh'
mkdir -p gpurun_out
timeout 100 nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather4_bw scripts/experiments/gather4_bw.cu -lcuda 2>&1 | grep -i error
(timeout 90 /tmp/gather4_bw 2>&1 | tail -20) | tee gpurun_out/gather4_bw.log
