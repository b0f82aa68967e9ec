#!/bin/bash
nproc; uptime
for t in 16 12 8 16 12 8; do
  SGMC_GATHER_THREADS=$t timeout 300 python bench.py --no-cpu-baseline --no-resgld --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('threads=$t e2e', round(d['e2e']['value']/1e6,2), 'M  device', round(d['value']/1e6,2))"
done
uptime
