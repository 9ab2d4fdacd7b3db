#!/bin/bash
# 8-GPU visit: 4- and 8-rank parity, scaling lines at N = 8 (sharded with prefetch, replicated) and N = 4
mkdir -p gpurun_out
for N in 4 8; do
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$((RANDOM%10)) \
   tests/dist_parity.py --transport nccl 2>&1 | grep -E "dist_parity|DIST_PARITY|Error|error|assert" | tail -14) | tee gpurun_out/dist_parity_$N.log
done
run() { N=$1; shift; tag=$1; shift
  (timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$((RANDOM%10)) \
     bench.py --gpus $N --no-eager "$@" 2> gpurun_out/bench${N}_$tag.err) | grep '^{' | tee gpurun_out/bench${N}_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N $tag: value %.3fM  ms/step %.3f  e2e %.3fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"
  tail -n 2 gpurun_out/bench${N}_$tag.err | cut -c1-200
}
run 8 sharded
run 8 replicated --table replicated
run 4 sharded
run 8 narre --model NARRE
