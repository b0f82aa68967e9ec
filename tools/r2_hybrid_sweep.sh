#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_api.py -m gpu -q -x --timeout 200 -k "streaming" 2>&1 | tail -3
for f in 0 0.125 0.25 0.375 0.5 0.25; do
  SGMC_HOST_PULL_FRACTION=$f SGMC_PULL_CTAS=${CT:-8} timeout 300 python bench.py --no-cpu-baseline --no-resgld --steps 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('frac=$f e2e', [round(x/1e6,1) for x in d['e2e']['runs']], d['e2e']['host_link'])"
done
