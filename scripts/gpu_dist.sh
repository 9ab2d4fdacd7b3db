#!/bin/bash
# N-GPU visit (gpurun --gpus N): N-rank parity (NCCL), sharded bench lines with and without prefetched lookups
N=${1:-2}
mkdir -p gpurun_out
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$((RANDOM%10)) \
   tests/dist_parity.py --transport nccl 2>&1 | grep -E "dist_parity|DIST_PARITY|Error|error|assert" | tail -20) | tee gpurun_out/dist_parity_$N.log
for f in "" "--no-prefetch" "--table replicated"; do
  tag=$(echo "$f" | tr -d ' -')
  (timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((RANDOM%10)) \
     bench.py --gpus $N --no-eager $f 2> gpurun_out/bench${N}_$tag.err) | tee gpurun_out/bench${N}_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N $f: value %.3fM  ms/step %.3f  e2e %.3fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"
  tail -n 2 gpurun_out/bench${N}_$tag.err | cut -c1-300
done
