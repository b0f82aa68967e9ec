#!/bin/bash
mkdir -p gpurun_out
MODE=step PATHS=tc_parity timeout 90 python tools/r2_timeline.py > gpurun_out/r2_timeline3.log 2>&1; head -8 gpurun_out/r2_timeline3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 48 --csv --log-file gpurun_out/r2_launches4.csv python bench.py --steps 100 --warmup 50 --no-cpu-baseline > gpurun_out/r2_ncu_bench4.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches4.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, v in agg.items(): print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v)/1e3:8.2f} us")
PY
