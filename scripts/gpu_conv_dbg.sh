#!/bin/bash
mkdir -p gpurun_out
for d in 0 14 "$@"; do
  CONV_PROF=1 R4R_CONV_DBG=$d timeout 120 python scripts/conv_bench.py --dist amazon 2>&1 | tail -6
done | tee gpurun_out/conv_dbg.log
