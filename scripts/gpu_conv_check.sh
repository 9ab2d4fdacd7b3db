#!/bin/bash
# conv kernel: parity tests, then the micro-benchmark with per-role cycle counters
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv" --tb=short 2>&1 | tail -5) | tee gpurun_out/t_conv.log
for d in 0 "$@"; do
  CONV_PROF=1 R4R_CONV_DBG=$d timeout 120 python scripts/conv_bench.py --dist amazon 2>&1 | tail -6
done | tee gpurun_out/conv_sweep.log
