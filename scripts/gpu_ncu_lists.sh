#!/bin/bash
# ncu launch lists (time + DRAM bytes per launch) of a bench step: plain and sharded-at-world-1
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
(timeout 300 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/launches_plain.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1)
(timeout 300 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/launches_shard1.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --force-shard > gpurun_out/ncu_shard1.log 2>&1)
tail -c 300 gpurun_out/ncu_plain.log; tail -c 300 gpurun_out/ncu_shard1.log
