#!/bin/bash
# ncu launch list of a bench step + full capture of the conv kernel
mkdir -p gpurun_out
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1)
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pool_tc -s 2 -c 1 -f -o gpurun_out/prof_conv3 \
   python scripts/conv_bench.py --iters 1 > gpurun_out/ncu_full.log 2>&1)
tail -2 gpurun_out/ncu_bench.log | cut -c1-300; tail -2 gpurun_out/ncu_full.log
