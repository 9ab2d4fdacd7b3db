#!/bin/bash
# conv kernel check after a change: parity tests first (each step under its own short timeout), then the
# determinism probe, then the micro-benchmark with per-role cycle counters
mkdir -p gpurun_out
(timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x --tb=short -k "conv or doc_plan" 2>&1 | tail -8) | tee gpurun_out/t_conv.log
(timeout 120 python scripts/conv_repro.py --reps 30 2>&1 | tail -8) | tee gpurun_out/repro.log
for d in amazon nopad; do
  (CONV_PROF=1 timeout 100 python scripts/conv_bench.py --dist $d 2>&1 | tail -8) | tee gpurun_out/conv_bench_$d.log
done
