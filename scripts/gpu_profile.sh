#!/bin/bash
# evidence for profiles/: ncu full captures of the two top kernels + launch list (time, DRAM bytes) of the bench step
mkdir -p gpurun_out
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_pool_tc -s 3 -c 1 -f -o gpurun_out/prof_conv_final \
   python scripts/conv_bench.py --iters 2 > gpurun_out/ncu_conv.log 2>&1)
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_half -s 6 -c 1 -f -o gpurun_out/prof_wgrad_final \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_wgrad.log 2>&1)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
(timeout 400 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1)
tail -n 2 gpurun_out/ncu_conv.log | cut -c1-200; tail -n 2 gpurun_out/ncu_wgrad.log | cut -c1-200; tail -c 300 gpurun_out/ncu_list.log
