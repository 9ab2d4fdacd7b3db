#!/bin/bash
# conv kernel: parity tests, then the micro-benchmark
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv" --tb=short 2>&1 | tail -5) | tee gpurun_out/t_conv.log
for dist in amazon uniform; do
  timeout 120 python scripts/conv_bench.py --dist $dist 2>&1 | tail -1
done | tee gpurun_out/conv_sweep.log
