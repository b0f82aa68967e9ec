#!/usr/bin/env python
"""Experiment: advance the C2 chains as G independent groups on G streams
(chains are independent, so kernels of different groups may overlap: tensor-bound
GEMM tiles of one group next to the ALU-bound update of another)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jax_sgmc_b200 import _lib, device, ops
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream

p = argparse.ArgumentParser()
p.add_argument("--groups", type=int, nargs="+", default=[1, 2, 4])
p.add_argument("--steps", type=int, default=1000)
p.add_argument("--offset", type=int, default=0, help="stagger group g by g*offset host steps")
a = p.parse_args()
C, d, n, N = 4096, 1024, 1024, 1_000_000
_lib.load()
device.set_device(0)
main = Stream.create()
device.set_current_stream(main)
X, y, _ = ops.synth_logistic_data(0, N, d)
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(X))
main.sync()

for G in a.groups:
  Cg = C // G
  grp = []
  for g in range(G):
    s = Stream.create()
    grp.append(dict(
        s=s, theta=DA.zeros((Cg, d)), v=DA.full((Cg, d), 1.0), grad=DA.zeros((Cg, d)),
        keys=[ops.prng_keys(range(g * Cg, (g + 1) * Cg)), DA((Cg, 2), np.uint32)],
        dkey=[DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)], idx=DA((n,), np.int32),
        U=DA((Cg,), np.float32), var=DA((Cg,), np.float32),
        ws=ops.glm_workspace(Cg, n, d, "tc_parity"), k=0, ev=Event()))
  device.synchronize()

  def step(q):
    k = q["k"]
    st = q["s"]
    ops.minibatch_draw(q["dkey"][k % 2], q["dkey"][(k + 1) % 2], q["idx"], N, stream=st)
    ops.glm_potential_grad(spec, q["theta"], X, y, q["idx"], N, q["U"], q["var"], q["grad"],
                           workspace=q["ws"], path="tc_parity", stream=st)
    ops.sgld_update(q["theta"], q["grad"], q["keys"][k % 2], q["keys"][(k + 1) % 2], [d],
                    1e-3, 1.0, v=q["v"], alpha=0.9, lmbd=1e-5, stream=st)
    q["k"] = k + 1

  def run(steps):
    e0, e1 = Event(), Event()
    e0.record(main)
    for q in grp:
      q["s"].wait_event(e0)
    for i in range(steps):
      for q in grp:
        step(q)
    for q in grp:
      q["ev"].record(q["s"])
      main.wait_event(q["ev"])
    e1.record(main)
    e1.sync()
    return e0.elapsed_ms(e1)

  run(50)
  ms = run(a.steps)
  print(f"groups={G}: {ms / a.steps * 1e3:.1f} us/step  {C * a.steps / ms * 1e3 / 1e6:.2f} M chain-steps/s",
        flush=True)
