#!/usr/bin/env python
"""Micro-benchmark of sgmc_glm_potential_grad per path (C2 shapes by default):
CUDA-event time per call, algorithmic TFLOP/s (4*n*d*C per call) against the
measured bf16 peak in MEASURED_PEAKS.json."""
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
  ap.add_argument("--features", type=int, default=1024)
  ap.add_argument("--batch", type=int, default=1024)
  ap.add_argument("--observations", type=int, default=100000)
  ap.add_argument("--paths", default="tc_parity,tc_throughput,simt")
  ap.add_argument("--reps", type=int, default=20)
  a = ap.parse_args()
  device.set_device(0)
  s = Stream.create()
  device.set_current_stream(s)
  C, d, n, N = a.chains, a.features, a.batch, a.observations
  pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
  peaks = json.load(open(pk)) if os.path.exists(pk) else {}
  X, y, _ = ops.synth_logistic_data(0, N, d)
  rng = np.random.default_rng(0)
  theta = DA.from_numpy((rng.standard_normal((C, d)) * 0.3).astype(np.float32))
  idx = DA((n,), np.int32)
  dk = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
  ops.minibatch_draw(dk[0], dk[1], idx, N)
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  flops = 4.0 * n * d * C
  out = {}
  for path in a.paths.split(","):
    ws = ops.glm_workspace(C, n, d, path)
    for _ in range(3):
      ops.glm_potential_grad(spec, theta, X, y, idx, N, U, var, g, workspace=ws, path=path)
    s.sync()
    e0, e1 = Event(), Event()
    e0.record(s)
    for _ in range(a.reps):
      ops.glm_potential_grad(spec, theta, X, y, idx, N, U, var, g, workspace=ws, path=path)
    e1.record(s)
    e1.sync()
    us = e0.elapsed_ms(e1) * 1e3 / a.reps
    tf = flops / (us * 1e-6) / 1e12
    out[path] = {"us": round(us, 1), "algorithmic_TFLOPs": round(tf, 1),
                 "frac_of_bf16_burst": round(tf / peaks.get("bf16_tflops", 1590.0), 3)}
    print(path, out[path], flush=True)
  print(json.dumps({"C": C, "d": d, "n": n, "paths": out}))


if __name__ == "__main__":
  main()
