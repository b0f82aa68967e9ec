#!/bin/bash
N=${1:-8}
for ov in 0 1; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N --only-resgld --resgld-overlap $ov --resgld-steps 400 2>/dev/null | grep us_per_step | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N overlap=$ov us/step', round(d['us_per_step'],1))"
done
