#!/bin/bash
# bench line (+ optional extra args), saved to gpurun_out/bench.json
mkdir -p gpurun_out
(timeout 900 python bench.py --steps 20 --warmup 5 "$@" 2> gpurun_out/bench_err.log) | tee gpurun_out/bench.json
tail -3 gpurun_out/bench_err.log
