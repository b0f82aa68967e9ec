#!/bin/bash
# e2e leg of the bench for the staging-ring variants
for rep in 1 2; do for t in 16 8; do
  SGMC_RING_WC=0 SGMC_GATHER_THREADS=$t timeout 300 python bench.py --no-cpu-baseline --no-resgld --steps 500 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('threads=$t e2e', round(d['e2e']['value']/1e6,2), 'M', d['e2e']['host_link'], 'device', round(d['value']/1e6,2))"
done; done
