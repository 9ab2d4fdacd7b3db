#!/bin/bash
# determinism probe of the conv kernel over the producer lags; optional R4R_LIB variants
mkdir -p gpurun_out
{
for lag in 1 2 4; do
  R4R_CONV_LAG=$lag timeout 120 python scripts/conv_repro.py --reps ${REPS:-60} 2>&1 | tail -6
done
R4R_CONV_LAG=1 timeout 120 python scripts/conv_repro.py --reps ${REPS:-60} --mode bf16 2>&1 | tail -6
} | tee gpurun_out/repro.log
