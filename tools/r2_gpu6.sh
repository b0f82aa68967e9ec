#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 120 -c 40 --csv --log-file gpurun_out/r2_launches6.csv python bench.py --steps 100 --warmup 50 --no-cpu-baseline > gpurun_out/r2_ncu_bench6.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches6.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); mi = hdr.index("Metric Name")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    try: agg[r[ki][:60]][r[mi]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
for k, m in agg.items():
    print(f"{k:62s}", " ".join(f"{n.split('__')[1][:14]}={sum(v)/len(v):12.1f}" for n, v in m.items()), "n=", len(next(iter(m.values()))))
PY
