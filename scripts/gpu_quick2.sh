#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_readers.py tests/test_gpu_models.py -m gpu -x -q --tb=short 2>&1 | tail -6) | tee gpurun_out/t_quick.log
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json | cut -c1-900
tail -n 5 gpurun_out/bench_err.log
