#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q --tb=short 2>&1 | tail -8) | tee gpurun_out/t_sharded1.log
(timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --force-shard 2> gpurun_out/bench_err2.log) | tee gpurun_out/bench_shard1.json | cut -c1-260
tail -n 3 gpurun_out/bench_err2.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
(timeout 300 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/launches_shard1.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --force-shard > gpurun_out/ncu_shard1.log 2>&1)
tail -c 200 gpurun_out/ncu_shard1.log
