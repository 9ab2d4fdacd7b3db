#!/bin/bash
# One GPU: full parity suite + smoke + bench line.   gpurun --timeout 1200 -- 'bash scripts/gpu_all.sh'
# Every step runs under its OWN short timeout (a hung kernel must not eat the GPU budget).
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25) | tee gpurun_out/t_gpu.log
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -4) | tee gpurun_out/t_smoke.log
(timeout 300 python bench.py "$@" 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json
tail -n 5 gpurun_out/bench_err.log
