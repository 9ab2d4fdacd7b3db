#!/bin/bash
mkdir -p gpurun_out
for l in 1 2 3 4; do R4R_CONV_LAG=$l timeout 40 python scripts/conv_bench.py --dist amazon --iters 9 2>&1 | tail -1 | cut -c1-200; done | tee gpurun_out/conv_lag.log
(timeout 90 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv_pool_tensor_core or doc_plan" --tb=short 2>&1 | tail -3) | tee gpurun_out/t_conv.log
