#!/bin/bash
# First GPU visit of a change: parity tests, bench line, ncu launch list, ncu full capture of the conv kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
(timeout 900 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15) > gpurun_out/t_gpu.log
(timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_err.log) > gpurun_out/bench.json
tail -5 gpurun_out/bench_err.log
(timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench_err.log) > gpurun_out/bench_ref.json
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1)
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pool_tc -s 4 -c 2 -f -o gpurun_out/prof_conv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1)
cat gpurun_out/t_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/ncu_bench.log; tail -3 gpurun_out/ncu_full.log
