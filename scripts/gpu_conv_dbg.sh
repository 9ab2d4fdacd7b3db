#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 8 6 12 14; do
  R4R_CONV_DBG=$d timeout 120 python scripts/conv_bench.py --dist amazon 2>&1 | tail -1
done | tee gpurun_out/conv_dbg.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pool_tc -s 2 -c 1 -f -o gpurun_out/prof_conv2 python scripts/conv_bench.py --iters 1 > gpurun_out/ncu_full2.log 2>&1
