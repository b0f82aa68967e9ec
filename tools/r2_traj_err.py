#!/usr/bin/env python
"""Diagnostic: relative trajectory error of the tensor-core parity path after
1 000 pSGLD / SGLD steps against the oracle (the quantity the parity tests
bound by 1e-5), per kernel variant (SGMC_OPTIONS selects the variant)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA  # noqa: E402
from oracle import data as odata, prng, scheduler as osched, sgmc as osgmc  # noqa: E402

device.set_device(0)


def run(rms, path, d=64, C=128, n=64, N=5000, K=1000, exact=False):
  ops.set_option(ops.OPT_EXACT_UPDATE_MATH, 1 if exact else 0)
  X, y, _ = odata.logistic_dataset(N, d, seed=0)
  theta0 = np.zeros((C, d), np.float32)
  keys = np.stack([prng.PRNGKey(c) for c in range(C)])
  eps = osched.polynomial_step_size_first_last(K, 1e-3 if not rms else 2e-2,
                                               1e-4 if not rms else 2e-3)
  spec = ops.glm_spec("logistic", d, w_off=0, prior="gaussian", prior_off=0,
                      prior_size=d, prior_scale=10.0)
  pot = osgmc.minibatch_potential(osgmc.Logistic(d, 0), osgmc.Prior("gaussian", 0, d, 10.0))
  dX, dy = DA.from_numpy(X), DA.from_numpy(y)
  d_theta = DA.from_numpy(theta0)
  d_v = DA.from_numpy(np.ones_like(theta0)) if rms else None
  d_k = [DA.from_numpy(keys), DA((C, 2), np.uint32)]
  d_dk = [DA.from_numpy(prng.PRNGKey(0)), DA((2,), np.uint32)]
  d_idx = DA((n,), np.int32)
  d_U, d_var, d_g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  ws = ops.glm_workspace(C, n, d, path)
  for k in range(K):
    ops.minibatch_draw(d_dk[k % 2], d_dk[(k + 1) % 2], d_idx, N)
    ops.glm_sgld_step(spec, d_theta, dX, dy, d_idx, N, d_U, d_var, d_g, d_k[k % 2],
                      d_k[(k + 1) % 2], eps[k], 1.0, v=d_v, workspace=ws, path=path)
  st = osgmc.langevin_init(theta0, keys, rms=rms)
  dk = prng.PRNGKey(0)
  for k in range(K):
    dk, idx = odata.device_draw(dk, n, N)
    Xb, yb = X[idx], y[idx]
    st = osgmc.langevin_update(st, lambda th: pot(th, (Xb, yb), N), [d], eps[k], 1.0)
  got = d_theta.numpy()
  err = np.abs(got - st.theta).max() / np.abs(st.theta).max()
  uerr = np.abs(d_U.numpy() - st.potential).max() / np.abs(st.potential).max()
  return err, uerr


for d, n in ((64, 64), (256, 128)):
  for rms in (False, True):
    for path in ("simt", "tc_parity"):
      for exact in (False, True):
        e, u = run(rms, path, d=d, n=n, exact=exact)
        print(f"d={d} n={n} rms={rms} path={path} exact_update={exact}: "
              f"traj err {e:.3e}  U err {u:.3e}", flush=True)
