#!/bin/bash
# 2-GPU visit (every step under a short timeout): sharded tests incl. 2-rank parity (NCCL + P2P), 2-GPU bench lines
mkdir -p gpurun_out
(timeout 420 python -m pytest tests/test_gpu_sharded.py -q --tb=short 2>&1 | tail -30) | tee gpurun_out/t_sharded2.log
for tr in nccl p2p; do
  (timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM%10)) \
     bench.py --gpus 2 --transport $tr 2> gpurun_out/bench2_$tr.err) | tee gpurun_out/bench2_$tr.json | cut -c1-250
  tail -n 4 gpurun_out/bench2_$tr.err | cut -c1-300
done
