#!/bin/bash
# determinism probe of the conv kernel: the same full-size launch REPS times in f16 and bf16; optional R4R_LIB variants
mkdir -p gpurun_out
{
timeout 120 python scripts/conv_repro.py --reps ${REPS:-60} 2>&1 | tail -6
timeout 120 python scripts/conv_repro.py --reps ${REPS:-60} --mode bf16 2>&1 | tail -6
} | tee gpurun_out/repro.log
