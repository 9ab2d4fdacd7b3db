#!/bin/bash
# full GPU test suite + smoke + bench line
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25) | tee gpurun_out/t_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4) | tee gpurun_out/t_smoke.log
(timeout 600 python bench.py --steps 400 --warmup 10 "$@" 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.log
