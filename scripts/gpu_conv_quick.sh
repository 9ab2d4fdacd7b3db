#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "conv" --tb=short 2>&1 | tail -4) | tee gpurun_out/t_conv.log
for p in 1 0; do CONV_PROF=1 R4R_DOC_PLAN=$p timeout 120 python scripts/conv_bench.py --dist amazon 2>&1 | tail -7; done | tee gpurun_out/conv_nacc.log
