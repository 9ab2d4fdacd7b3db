for c in 70 68 66 62 70; do
 (R4R_CONV_CLUSTERS=$c timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM%10)) \
     bench.py --gpus 2 --no-eager 2>/dev/null) | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('clusters $c: value %.3fM  ms/step %.3f  e2e %.3fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"
done
