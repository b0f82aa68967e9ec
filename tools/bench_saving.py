#!/usr/bin/env python
"""Cost of keeping samples at C2 scale (4096 chains x 1024 parameters = 16.8 MB
per kept sample): the host ring of jax_sgmc_b200.io (d2d into a staging slot on
the sampling stream, D2H on a copy stream) against not saving at all."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jax_sgmc_b200 import _lib, device, io, ops
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream

C, d, n, N, steps = 4096, 1024, 1024, 1_000_000, 600
_lib.load()
device.set_device(0)
stream = Stream.create()
device.set_current_stream(stream)
X, y, _ = ops.synth_logistic_data(0, N, d)
theta, v, grad = DA.zeros((C, d)), DA.full((C, d), 1.0), DA.zeros((C, d))
keys = [ops.prng_keys(range(C)), DA((C, 2), np.uint32)]
dkey = [DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)]
idx = DA((n,), np.int32)
U, var = DA((C,), np.float32), DA((C,), np.float32)
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(X))
ws = ops.glm_workspace(C, n, d, "tc_parity")
k = 0


def step():
  global k
  ops.minibatch_draw(dkey[k % 2], dkey[(k + 1) % 2], idx, N)
  ops.glm_sgld_step(spec, theta, X, y, idx, N, U, var, grad, keys[k % 2], keys[(k + 1) % 2],
                    1e-3, 1.0, v=v, workspace=ws, path="tc_parity", write_grad=False)
  k += 1


for _ in range(50):
  step()
stream.sync()
out = []
for every in (0, 10, 4, 2, 1):
  kept = steps // every if every else 0
  ring = io._HostRing(kept, C, d) if every else None
  e0, e1 = Event(), Event()
  e0.record(stream)
  cnt = 0
  for i in range(steps):
    step()
    if every and (i + 1) % every == 0 and cnt < kept:
      ring.push(cnt, theta)
      cnt += 1
  e1.record(stream)
  e1.sync()
  if ring is not None:
    ring.finish(cnt)
  ms = e0.elapsed_ms(e1) / steps
  out.append({"keep_every": every, "kept": cnt, "us_per_step": ms * 1e3,
              "chain_steps_per_s": C / (ms * 1e-3),
              "d2h_gbs": (cnt * C * d * 4) / (ms * steps * 1e-3) / 1e9})
  print(json.dumps(out[-1]), flush=True)
