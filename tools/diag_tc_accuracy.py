#!/usr/bin/env python
"""Accuracy diagnostic: z = theta.x recovered from ell for the GLM paths vs a
float64 reference, as a function of the contraction length d."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA  # noqa: E402

device.set_device(0)
rng = np.random.default_rng(0)
for d in (64, 256, 1024, 4096):
  C, n, N = 256, 512, 2048
  X = (rng.standard_normal((N, d)) / np.sqrt(d)).astype(np.float32)
  theta = (rng.standard_normal((C, d)) * 3).astype(np.float32)
  y = np.ones(N, np.float32)                 # ell = z - softplus(z) = -softplus(-z)
  idx = np.arange(n, dtype=np.int32)
  z64 = theta.astype(np.float64) @ X[:n].astype(np.float64).T
  ell64 = -np.logaddexp(0, -z64)
  g64 = None
  spec = ops.glm_spec("logistic", d, 0)
  for path in ("simt", "tc_parity", "tc_throughput"):
    U, var = DA((C,), np.float32), DA((C,), np.float32)
    g, ell = DA((C, d), np.float32), DA((C, n), np.float32)
    ops.glm_potential_grad(spec, DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y),
                           DA.from_numpy(idx), N, U, var, g, ell, path=path)
    e = ell.numpy().astype(np.float64) - ell64
    # gradient reference in f64
    sig = 1 / (1 + np.exp(-z64))
    g64 = ((1 - sig) * (-N / n)) @ X[:n].astype(np.float64)
    ge = g.numpy().astype(np.float64) - g64
    print(f"d={d:5d} {path:14s} ell err: mean {e.mean():+.2e} rms {np.sqrt((e**2).mean()):.2e} "
          f"max {np.abs(e).max():.2e} | grad err/scale: rms "
          f"{np.sqrt((ge**2).mean()) / np.abs(g64).max():.2e} max "
          f"{np.abs(ge).max() / np.abs(g64).max():.2e} mean {ge.mean() / np.abs(g64).max():+.2e}",
          flush=True)
