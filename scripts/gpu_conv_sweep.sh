#!/bin/bash
mkdir -p gpurun_out
for ld in cg ca; do for dist in amazon nopad uniform; do
  R4R_CONV_LD=$ld timeout 120 python scripts/conv_bench.py --dist $dist 2>&1 | tail -1
done; done | tee gpurun_out/conv_sweep.log
