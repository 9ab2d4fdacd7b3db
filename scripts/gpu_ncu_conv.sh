#!/bin/bash
# full ncu capture of the conv kernel (doc plan on) from the micro-benchmark
mkdir -p gpurun_out
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pool_tc -s 3 -c 1 -f -o gpurun_out/prof_conv_plan \
   python scripts/conv_bench.py --iters 2 > gpurun_out/ncu_full.log 2>&1)
tail -3 gpurun_out/ncu_full.log
