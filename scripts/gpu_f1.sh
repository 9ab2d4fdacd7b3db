#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15) | tee gpurun_out/t_gpu.log
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json | cut -c1-1300
tail -n 5 gpurun_out/bench_err.log
