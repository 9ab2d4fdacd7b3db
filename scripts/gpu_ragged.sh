#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_readers.py tests/test_gpu_kernels.py -m gpu -x -q --tb=short 2>&1 | tail -12) | tee gpurun_out/t_ragged.log
for r in "" "--ragged"; do CONV_PROF=1 timeout 120 python scripts/conv_bench.py --dist amazon $r 2>&1 | tail -7; done | tee gpurun_out/conv_ragged.log
(timeout 400 python bench.py --no-cpu-baseline 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json | cut -c1-1400
tail -n 5 gpurun_out/bench_err.log
