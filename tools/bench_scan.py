#!/usr/bin/env python
"""Device-resident C2 scan timing: K pSGLD steps through sgmc_glm_sgld_scan_device (the
native scan solver.mcmc hands the whole lax.scan to), CUDA events around the call."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--path", default="tc_parity")
a = ap.parse_args()
device.set_device(0)
s = Stream.create()
device.set_current_stream(s)
C, d, n, N = 4096, 1024, 1024, 1_000_000
X, y, _ = ops.synth_logistic_data(0, N, d)
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(X))
theta, v, g = DA.zeros((C, d)), DA.full((C, d), 1.0), DA.zeros((C, d))
U, var = DA((C,), np.float32), DA((C,), np.float32)
keys = [ops.prng_keys(range(C)), DA((C, 2), np.uint32)]
dk = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
idx = DA((n,), np.int32)
ws = ops.glm_workspace(C, n, d, a.path)
K = a.steps
eps = np.full(K, 1e-3, np.float32)
tau = np.ones(K, np.float32)
for rep in range(a.reps + 1):
  e0, e1 = Event(), Event()
  e0.record(s)
  ops.glm_sgld_scan_device(spec, theta, X, y, N, n, U, var, g, keys[0], keys[1], [d], eps, tau,
                           np.zeros(K, np.uint8), None, None, 0, data_key_a=dk[0], data_key_b=dk[1], idx_buf=idx,
                           idx_all=None, v=v, alpha=0.9, lmbd=1e-5, workspace=ws, path=a.path)
  e1.record(s)
  e1.sync()
  ms = e0.elapsed_ms(e1)
  if rep:
    print(f"options={os.environ.get('SGMC_OPTIONS', '')!r} {K} steps: {ms * 1e3 / K:.2f} us/step "
          f"{C * K / (ms * 1e-3) / 1e6:.2f} M chain-steps/s", flush=True)
