#!/bin/bash
N=${1:-2}
for pull in 0 1; do
  SGMC_HOST_PULL=$pull timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 500 --no-resgld 2>/dev/null | grep '"metric"' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=$N pull=$pull value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), d['e2e']['host_link'])"
done
