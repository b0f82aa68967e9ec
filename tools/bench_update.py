#!/usr/bin/env python
"""Micro-benchmark of the fused update kernels (HBM roofline).

Times each kernel on rotating state sets larger than L2 with CUDA events and
prints achieved algorithmic GB/s against MEASURED_PEAKS.json.  Also the command
to run under ncu (`--kernels sgld_rms --reps 2`)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--chains", type=int, default=4096)
  ap.add_argument("--params", type=int, default=1024)
  ap.add_argument("--reps", type=int, default=20)
  ap.add_argument("--sets", type=int, default=6)
  ap.add_argument("--kernels", default="sgld,sgld_rms,sghmc,obabo_a,obabo_b,normal_like")
  ap.add_argument("--layout", default="original")
  ap.add_argument("--sizes", default="", help="comma-separated leaf sizes (overrides --params)")
  a = ap.parse_args()
  device.set_device(0)
  s = Stream.create()
  device.set_current_stream(s)
  leaf_sizes = [int(x) for x in a.sizes.split(",")] if a.sizes else [a.params]
  C, P, R = a.chains, sum(leaf_sizes), a.sets
  pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
  peak = json.load(open(pk)).get("hbm_gbs", 6650.0) if os.path.exists(pk) else 6650.0
  sets = [dict(t=DA.zeros((C, P)), v=DA.full((C, P), 1.0), g=DA.full((C, P), 0.01),
               p=DA.zeros((C, P))) for _ in range(R)]
  ke = DA.zeros((C,))
  kk = [ops.prng_keys(range(C)), DA((C, 2), np.uint32)]
  sizes = leaf_sizes
  eps = 1e-3

  def run(name, i):
    b = sets[i % R]
    kin, kout = kk[i % 2], kk[(i + 1) % 2]
    if name == "sgld":
      ops.sgld_update(b["t"], b["g"], kin, kout, sizes, eps, 1.0, layout=a.layout)
    elif name == "sgld_rms":
      ops.sgld_update(b["t"], b["g"], kin, kout, sizes, eps, 1.0, v=b["v"],
                      layout=a.layout)
    elif name == "sghmc":
      ops.sghmc_step(b["t"], b["p"], b["g"], kin, kout, sizes, eps, 0.9,
                     layout=a.layout)
    elif name == "obabo_a":
      ops.obabo_pass_a(b["t"], b["p"], b["g"], ke, kin, kout, sizes, eps,
                       layout=a.layout)
    elif name == "obabo_b":
      ops.obabo_pass_b(b["p"], b["g"], ke, kin, sizes, eps, layout=a.layout)
    elif name == "normal_like":
      ops.normal_like(kin, sizes, a.layout, out=b["t"])

  bytes_pp = {"sgld": 12, "sgld_rms": 20, "sghmc": 20, "obabo_a": 20,
              "obabo_b": 12, "normal_like": 4}
  out = {}
  for name in a.kernels.split(","):
    for i in range(R):
      run(name, i)
    s.sync()
    e0, e1 = Event(), Event()
    e0.record(s)
    for i in range(a.reps * R):
      run(name, i)
    e1.record(s)
    e1.sync()
    us = e0.elapsed_ms(e1) * 1e3 / (a.reps * R)
    gbs = C * P * bytes_pp[name] / (us * 1e-6) / 1e9
    out[name] = {"us": round(us, 2), "GBps": round(gbs, 1),
                 "frac_of_measured_hbm": round(gbs / peak, 3),
                 "Gparams_per_s": round(C * P / us / 1e3, 2)}
    print(name, out[name], flush=True)
  print(json.dumps({"chains": C, "params": P, "hbm_peak": peak, "kernels": out}))


if __name__ == "__main__":
  main()
