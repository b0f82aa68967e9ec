#!/usr/bin/env python
"""Quick device-side check of the tensor-core potential against the fp32 SIMT
path (same device, same inputs): max relative errors of U, var, grad per shape.
Usage: r2_quick.py [SGMC_OPTIONS string]"""
import os
import sys

if len(sys.argv) > 1:
  os.environ["SGMC_OPTIONS"] = sys.argv[1]
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA  # noqa: E402

device.set_device(0)
print("options:", os.environ.get("SGMC_OPTIONS", ""), flush=True)
for (C, n, d) in [(128, 256, 64), (256, 512, 256), (130, 264, 72), (512, 1024, 1024),
                  (4096, 1024, 1024)]:
  rng = np.random.default_rng(C + n + d)
  N = 3000
  X = (rng.standard_normal((N, d)) / np.sqrt(d)).astype(np.float32)
  w = rng.standard_normal(d).astype(np.float32)
  y = (rng.random(N) < 1 / (1 + np.exp(-(X @ w)))).astype(np.float32)
  theta = (rng.standard_normal((C, d)) * 0.7).astype(np.float32)
  theta[0] *= 1e-3
  theta[-1] *= 30.0
  idx = rng.integers(0, N, n).astype(np.int32)
  spec = ops.glm_spec("logistic", d, w_off=0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  out = {}
  for path in ("simt", "tc_parity", "tc_throughput"):
    U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
    ell = DA((C, n), np.float32)
    ops.glm_potential_grad(spec, DA.from_numpy(theta), DA.from_numpy(X), DA.from_numpy(y),
                           DA.from_numpy(idx), N, U, var, g, ell, path=path)
    device.synchronize()
    out[path] = (U.numpy(), var.numpy(), g.numpy(), ell.numpy())
  U0, v0, g0, l0 = out["simt"]
  for path in ("tc_parity", "tc_throughput"):
    U1, v1, g1, l1 = out[path]
    gs = np.abs(g0).max(axis=1, keepdims=True)
    print(f"C={C} n={n} d={d} {path}: U {np.abs(U1 / U0 - 1).max():.2e} var "
          f"{np.abs(v1 / v0 - 1).max():.2e} grad {(np.abs(g1 - g0) / gs).max():.2e} ell "
          f"{np.abs(l1 - l0).max() / np.abs(l0).max():.2e}", flush=True)
print("done", flush=True)
