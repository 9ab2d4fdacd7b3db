#!/bin/bash
# single GPU: sharded-path tests (virtual ranks, world 1) and the device-side cost of the sharded lookup
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q --tb=short 2>&1 | tail -8) | tee gpurun_out/t_sharded1.log
(timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2> gpurun_out/bench_err.log) | tee gpurun_out/bench_plain.json | cut -c1-260
(timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --force-shard 2> gpurun_out/bench_err2.log) | tee gpurun_out/bench_shard1.json | cut -c1-260
tail -n 3 gpurun_out/bench_err.log; tail -n 3 gpurun_out/bench_err2.log
