#!/bin/bash
# doc-plan change: parity tests, conv micro-bench with the plan off/on, bench line
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q --tb=short 2>&1 | tail -15) | tee gpurun_out/t_kernels.log
for p in 0 1; do
  CONV_PROF=1 R4R_DOC_PLAN=$p timeout 120 python scripts/conv_bench.py --dist amazon 2>&1 | tail -8
done | tee gpurun_out/conv_plan.log
(timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.log
