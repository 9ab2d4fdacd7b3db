#!/bin/bash
# bench lines of every model (1 GPU).   gpurun --timeout 1500 -- 'bash scripts/gpu_bench_models.sh'
mkdir -p gpurun_out
for m in deepconn deepconn++ NARRE transnet++; do
  f=$(echo $m | tr '+' 'p')
  (timeout 400 python bench.py --model $m "$@" 2> gpurun_out/bench_${f}_err.log) | tee gpurun_out/bench_$f.json | cut -c1-600
  tail -n 3 gpurun_out/bench_${f}_err.log
done
