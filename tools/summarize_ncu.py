#!/usr/bin/env python
"""Summarise ncu artefacts into small tracked files under profiles/.

  summarize_ncu.py launches <launches.csv> <out.csv>      per-kernel totals / shares
  summarize_ncu.py report <file.ncu-rep> <out.md>         key metrics per captured launch
  summarize_ncu.py traffic <file.ncu-rep> <out.json>      dram bytes per launch of the fused
                                                          update (bench.py's roofline.traffic)
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(src, dst):
  rows = [r for r in csv.reader(open(src)) if len(r) > 10]
  hdr = rows[0]
  ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
  agg = OrderedDict()
  for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
      continue
    name = r[ki].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
  total = sum(a[1] for a in agg.values())
  with open(dst, "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_ns", "avg_ns", "share_of_gpu_time"])
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      w.writerow([k, n, int(t), int(t / n), f"{t / total:.4f}"])
  print(open(dst).read())


def report(src, dst):
  raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  ki = hdr.index("Kernel Name")
  with open(dst, "w") as f:
    f.write(f"# ncu --set full summary of `{src}`\n\n")
    for r in data:
      f.write(f"## {r[ki][:100]}\n\n| metric | value | unit |\n|---|---|---|\n")
      for k in KEYS:
        if k in hdr:
          i = hdr.index(k)
          f.write(f"| {k} | {r[i]} | {units[i]} |\n")
      f.write("\n")
  print(open(dst).read()[:3000])


def traffic(src, dst):
  import json
  raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"],
                       capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
  ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
  it = hdr.index("gpu__time_duration.sum")
  rd = [float(r[ir].replace(",", "")) * mult[units[ir]] for r in data]
  wr = [float(r[iw].replace(",", "")) * mult[units[iw]] for r in data]
  out = {"source": src, "kernel": data[0][hdr.index("Kernel Name")][:80], "launches": len(data),
         "dram_read_bytes_per_launch": sum(rd) / len(rd),
         "dram_write_bytes_per_launch": sum(wr) / len(wr),
         "k_noise_pass_rms_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
         "us_per_launch_under_ncu": sum(float(r[it].replace(",", "")) for r in data) / len(data),
         "note": "ncu --set full --clock-control none: the caches are flushed before every "
                 "replay pass, so all loads come from DRAM, and the launch's own stores are "
                 "still in the write-back L2 when the measurement ends (they show up as DRAM "
                 "writes of LATER kernels: see r02_warm_traffic_raw.csv for a warm run)"}
  json.dump(out, open(dst, "w"), indent=1)
  print(json.dumps(out, indent=1))


if __name__ == "__main__":
  {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
