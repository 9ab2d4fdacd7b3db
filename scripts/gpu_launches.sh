#!/bin/bash
# ncu launch lists (time + DRAM bytes per launch) of a few bench steps per model -> gpurun_out/launches_<model>.csv
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for m in "$@"; do
  f=$(echo $m | tr '+' 'p')
  (timeout 500 ncu --metrics $M --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$f.csv \
     python bench.py --model $m --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-fp32 > gpurun_out/ncu_list_$f.log 2>&1)
  tail -c 200 gpurun_out/ncu_list_$f.log; echo
done
