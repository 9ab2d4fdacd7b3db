#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -25) | tee gpurun_out/t_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) | tee gpurun_out/t_smoke.log
(timeout 400 python bench.py 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json
tail -n 5 gpurun_out/bench_err.log
(timeout 400 python bench.py --docs padded --no-cpu-baseline 2> gpurun_out/bench_err_p.log) | tee gpurun_out/bench_padded.json | cut -c1-1500
tail -n 5 gpurun_out/bench_err_p.log
