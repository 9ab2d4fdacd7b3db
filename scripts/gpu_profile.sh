#!/bin/bash
# evidence for profiles/: ncu launch list (time + DRAM bytes) of the bench step, ncu --set full of the conv and wgrad kernels
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
(timeout 400 ncu --metrics $M --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager --no-fp32 > gpurun_out/ncu_list.log 2>&1)
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_pool_tc -s 3 -c 1 -f -o gpurun_out/prof_conv \
   python scripts/conv_bench.py --iters 2 > gpurun_out/ncu_conv.log 2>&1)
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 4 -c 1 -f -o gpurun_out/prof_wgrad \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-fp32 > gpurun_out/ncu_wgrad.log 2>&1)
tail -c 200 gpurun_out/ncu_list.log; tail -n 2 gpurun_out/ncu_conv.log | cut -c1-200; tail -n 2 gpurun_out/ncu_wgrad.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
