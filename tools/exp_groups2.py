#!/usr/bin/env python
"""Experiment: the C2 chains as G independent groups, every group its own native scan
(sgmc_glm_sgld_scan_device on C / G chains) on its own stream from its own host thread,
the potential kernel of a group capped at 74 / G CTA pairs -- so the HBM-bound update of
one group can run under the tensor-bound potential of another."""
import argparse
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jax_sgmc_b200 import _lib, device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--groups", type=int, nargs="+", default=[1, 2])
p.add_argument("--steps", type=int, default=1000)
p.add_argument("--pairs", type=int, default=0, help="pair cap per group (0: 74 // G)")
p.add_argument("--stagger", type=float, default=0.5, help="start group g after g*stagger steps")
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
K = a.steps
eps, tau, keep = np.full(K, 1e-3, np.float32), np.ones(K, np.float32), np.zeros(K, np.uint8)

for G in a.groups:
  Cg = C // G
  ops.set_option(ops.OPT_TC_MAX_PAIRS, 0 if G == 1 else (a.pairs or 74 // G))
  grp = []
  for g in range(G):
    grp.append(dict(
        s=Stream.create(), theta=DA.zeros((Cg, d)), v=DA.full((Cg, d), 1.0), grad=DA.zeros((Cg, d)),
        keys=[ops.prng_keys(range(g * Cg, (g + 1) * Cg)), DA((Cg, 2), np.uint32)],
        dkey=[DA.from_numpy(ops.prng_key(0)), DA((2,), np.uint32)], idx=DA((n,), np.int32),
        U=DA((Cg,), np.float32), var=DA((Cg,), np.float32),
        ws=ops.glm_workspace(Cg, n, d, "tc_parity")))
  device.synchronize()

  def scan(q, k, delay=0.0):
    device.set_device(0)
    if delay:
      time.sleep(delay)
    ops.glm_sgld_scan_device(spec, q["theta"], X, y, N, n, q["U"], q["var"], q["grad"], q["keys"][0],
                             q["keys"][1], [d], eps[:k], tau[:k], keep[:k], None, None, 0,
                             data_key_a=q["dkey"][0], data_key_b=q["dkey"][1], idx_buf=q["idx"],
                             idx_all=None, v=q["v"], alpha=0.9, lmbd=1e-5, workspace=q["ws"],
                             path="tc_parity", stream=q["s"])
    q["s"].sync()

  def run(k):
    ths = [threading.Thread(target=scan, args=(q, k, g * a.stagger * 85e-6)) for g, q in enumerate(grp)]
    t0 = time.perf_counter()
    for t in ths:
      t.start()
    for t in ths:
      t.join()
    return time.perf_counter() - t0

  run(100)
  for rep in range(2):
    s = run(K)
    print(f"groups={G} pairs/group={ops.get_option(ops.OPT_TC_MAX_PAIRS) if hasattr(ops, 'get_option') else '?'}: "
          f"{s / K * 1e6:.1f} us per step of all {C} chains = {C * K / s / 1e6:.2f} M chain-steps/s", flush=True)
