#!/usr/bin/env python
"""C3 timing: the 784-512-512-10 classifier potential + gradient for 256 chains on a
shared minibatch of 256 (sgmc_mlp_potential_grad), and one SGHMC outer step (5 leapfrog
steps = 5 gradients + 6 fused update passes) through integrator.friction_leapfrog."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import data, device, glm, integrator, nn, ops, potential, scheduler  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA, Event, Stream  # noqa: E402
from jax_sgmc_b200.tree_util import ChainTree  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=256)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
device.set_device(0)
s = Stream.create()
device.set_current_stream(s)
sizes, C, n, N = (784, 512, 512, 10), a.chains, a.batch, 60000
rng = np.random.default_rng(0)
X = DA.from_numpy(rng.random((N, 784)).astype(np.float32))
y = DA.from_numpy(rng.integers(0, 10, N).astype(np.float32))
tree = nn.init_params(ops.prng_key(0), sizes)
one = ChainTree.from_trees([tree])
P = one.n_params
theta = DA.from_numpy(np.tile(one.flat.numpy(), (C, 1)))
sample = ChainTree.like(one, theta)
pot = potential.minibatch_potential(glm.GaussianPrior(10.0), nn.MLPClassifier())
spec = nn.resolve(pot.likelihood, pot.prior, sample, 1.0)
ws = ops.mlp_workspace(spec, C, n)
idx = DA.from_numpy(rng.integers(0, N, n).astype(np.int32))
U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, P), np.float32)
for _ in range(2):
  ops.mlp_potential_grad(spec, theta, X, y, idx, N, U, var, g, workspace=ws)
e0, e1 = Event(), Event()
e0.record(s)
for _ in range(a.reps):
  ops.mlp_potential_grad(spec, theta, X, y, idx, N, U, var, g, workspace=ws)
e1.record(s)
e1.sync()
ms = e0.elapsed_ms(e1) / a.reps
flops = C * n * 2.0 * (3 * (784 * 512 + 512 * 512 + 512 * 10) - 784 * 512)
print(f"mlp potential+grad: {ms * 1e3:.1f} us per call, {flops / (ms * 1e-3) / 1e12:.1f} TFLOP/s "
      f"(fp32 FFMA), {C} chains x batch {n}", flush=True)
loader = data.DeviceNumpyDataLoader(x=X, y=y)
init, integrate, get = integrator.friction_leapfrog(pot, data.random_reference_data(loader, 1, n),
                                                    steps=5, friction=1.0)
state = init(sample, key=np.stack([ops.prng_key(c) for c in range(C)]))
sch = scheduler.schedule(np.float32(1e-5), np.float32(1.0), 1.0, True)
state = integrate(state, sch)
s.sync()
e0.record(s)
for _ in range(3):
  state = integrate(state, sch)
e1.record(s)
e1.sync()
ms = e0.elapsed_ms(e1) / 3
print(f"SGHMC outer step (5 leapfrog steps): {ms:.2f} ms -> {C * 5 / (ms * 1e-3):.0f} chain-leapfrog-steps/s",
      flush=True)
