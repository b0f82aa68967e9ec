#!/usr/bin/env python
"""Why one cold-start pSGLD step (v = 1, |g| up to 1e3) cannot be compared at 1e-5 of
|theta|: the same step computed from an fp64 gradient, from the fp32 SIMT gradient and
from the tensor-core (parity) gradient, all three pushed through the SAME fp64 update
formula (integrator.py:882-912, adaption.py:270-291).  Prints the spread of theta'
against the fp64 result for both device gradients, next to the gradient errors."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA  # noqa: E402

device.set_device(0)
C, d, n, N = 512, 1024, 1024, 1_000_000
X, y, _ = ops.synth_logistic_data(0, 65536, d)
hX, hy = X.numpy(), y.numpy()
rng = np.random.default_rng(0)
theta = (rng.standard_normal((C, d)) * 0.05).astype(np.float32)
idx = rng.integers(0, 65536, n).astype(np.int32)
spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                    prior_scale=10.0, x_absmax=ops.absmax(X))
Xb, yb = hX[idx].astype(np.float64), hy[idx].astype(np.float64)
z = theta.astype(np.float64) @ Xb.T
sig = 1 / (1 + np.exp(-z))
g64 = ((yb[None] - sig) * (-N / n)) @ Xb + theta.astype(np.float64) / 100.0
out = {}
for path in ("simt", "tc_parity"):
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  ops.glm_potential_grad(spec, DA.from_numpy(theta), X, y, DA.from_numpy(idx), N, U, var, g,
                         path=path)
  out[path] = g.numpy().astype(np.float64)


def step(g, eps=1e-3, v=1.0, alpha=0.9, lmbd=1e-5):
  vn = alpha * v + (1 - alpha) * g * g
  G = 1.0 / (lmbd + np.sqrt(vn))
  return theta.astype(np.float64) + G * (-eps * g)       # noise term is common: omitted


ref = step(g64)
scale_t = np.abs(ref).max(axis=1, keepdims=True)
scale_g = np.abs(g64).max(axis=1, keepdims=True)
for path, g in out.items():
  print(f"{path:10s}: grad err / row max {np.abs(g - g64).max() / 1:.3e} abs, "
        f"{(np.abs(g - g64) / scale_g).max():.2e} rel | theta' err / row max|theta'| "
        f"cold v=1: {(np.abs(step(g) - ref) / scale_t).max():.2e}   "
        f"adapted v=g^2: {(np.abs(step(g, v=g64 * g64) - step(g64, v=g64 * g64)) / scale_t).max():.2e}")
