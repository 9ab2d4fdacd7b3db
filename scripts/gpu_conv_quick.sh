#!/bin/bash
# Conv kernel change: its parity tests, then the micro-benchmark with per-role cycle counters.
#   gpurun --timeout 300 -- 'bash scripts/gpu_conv_quick.sh'      (inner timeouts sum to < 200 s)
mkdir -p gpurun_out
(timeout 90 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -x -q -k "conv" --tb=short 2>&1 | tail -4) | tee gpurun_out/t_conv.log
for p in 1 0; do CONV_PROF=1 R4R_DOC_PLAN=$p timeout 45 python scripts/conv_bench.py --dist amazon --iters 9 2>&1 | tail -7; done | tee gpurun_out/conv_bench.log
