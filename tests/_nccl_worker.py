"""One rank of the 2-GPU NCCL tests (tests/test_gpu_multi.py starts RANK = 0, 1):
(1) sgmc_glm_potential_grad_row_sharded with ncclAllReduce == the single-process
emulation (partials of both ranks summed on the host), bit for bit at R = 2, and == the
unsharded evaluation within the parity tolerance; (2) ncclAllGather moves the rows."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jax_sgmc_b200 import device, dist, ops  # noqa: E402
from jax_sgmc_b200.device import DeviceArray as DA  # noqa: E402

rank, world, local = dist.env_rank_world()
device.set_device(local)
ctl = dist.SocketCommunicator()
nccl = dist.NcclCommunicator.from_control_plane(ctl)

# ---- all-gather ------------------------------------------------------------------------
send = DA.from_numpy(np.full((5,), rank + 1, np.float32))
recv = DA.zeros((world, 5))
nccl.allgather(send, recv)
device.synchronize()
assert np.array_equal(recv.numpy(), np.arange(1, world + 1, dtype=np.float32)[:, None] * np.ones(5))

# ---- row-sharded GLM gradient -------------------------------------------------------------
for path, (C, d, n, N) in (("simt", (6, 24, 64, 400)), ("tc_parity", (128, 64, 256, 2000))):
  rng = np.random.default_rng(3)
  X = (rng.standard_normal((N, d)) / np.sqrt(d)).astype(np.float32)
  y = (rng.random(N) < 0.5).astype(np.float32)
  theta = (rng.standard_normal((C, d)) * 0.3).astype(np.float32)
  idx = rng.integers(0, N, n).astype(np.int32)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=10.0)
  dX, dy, dt, di = DA.from_numpy(X), DA.from_numpy(y), DA.from_numpy(theta), DA.from_numpy(idx)
  U, var, g = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  ops.glm_potential_grad_row_sharded(spec, dt, dX, dy, di, N, U, var, g, n, rank, world,
                                     nccl._comm.value, path=path)
  device.synchronize()
  # single-process emulation: both ranks' partials, summed in rank order
  parts = []
  for r in range(world):
    Ur, vr, gr = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
    _, scratch = ops.glm_potential_grad_row_sharded(spec, dt, dX, dy, di, N, Ur, vr, gr, n, r,
                                                    world, None, path=path)
    device.synchronize()
    n_r = n // world
    parts.append((gr.numpy(), scratch.numpy()[C * n_r:C * n_r + 3 * C].reshape(C, 3)))
  g_sum, ex_sum = parts[0][0].copy(), parts[0][1].copy()
  for gp, ep in parts[1:]:
    g_sum, ex_sum = g_sum + gp, ex_sum + ep
  U2, v2 = DA((C,), np.float32), DA((C,), np.float32)
  ops.glm_row_shard_finalize(DA.from_numpy(ex_sum), n, U2, v2)
  device.synchronize()
  if world == 2:
    assert np.array_equal(g.numpy().view(np.uint32), g_sum.view(np.uint32)), path
    assert np.array_equal(U.numpy().view(np.uint32), U2.numpy().view(np.uint32)), path
    assert np.array_equal(var.numpy().view(np.uint32), v2.numpy().view(np.uint32)), path
  # the unsharded evaluation
  U0, v0, g0 = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  ops.glm_potential_grad(spec, dt, dX, dy, di, N, U0, v0, g0, path=path)
  device.synchronize()
  np.testing.assert_allclose(U.numpy(), U0.numpy(), rtol=1e-5)
  np.testing.assert_allclose(var.numpy(), v0.numpy(), rtol=2e-4)
  scale = np.abs(g0.numpy()).max(axis=1, keepdims=True)
  assert (np.abs(g.numpy() - g0.numpy()) / scale).max() < 1e-5, path
# ---- CNN potential with the minibatch rows sharded over the ranks (configs[4]) -------------
from jax_sgmc_b200 import data, glm, nn, potential  # noqa: E402
from jax_sgmc_b200.tree_util import ChainTree  # noqa: E402

rng = np.random.default_rng(11)
N, n, C = 96, 32, 3
X = rng.random((N, 8, 8, 3)).astype(np.float32)
y = rng.integers(0, 5, N).astype(np.float32)
loader = data.DeviceNumpyDataLoader(x=X, y=y)
trees = [nn.init_cnn_params(ops.prng_key(c), (8, 8, 3), (4, 6), (2, 1), 5) for c in range(C)]
sample = ChainTree.from_trees(trees)
init_fn, get_fn, _ = data.random_reference_data(loader, 1, n)
_, ref = get_fn(init_fn(), information=True)           # same key on every rank: same minibatch
lik, prior = nn.CNNClassifier(strides=(2, 1)), glm.GaussianPrior(3.0)
whole = potential.minibatch_potential(prior, lik)
(U0, _), g0 = whole.value_and_grad(sample, ref)
U0, g0 = U0.numpy(), g0.flat.numpy()
sharded = potential.minibatch_potential(prior, lik).shard_rows(nccl)
(U1, _), g1 = sharded.value_and_grad(sample, ref)
device.synchronize()
np.testing.assert_allclose(U1.numpy(), U0, rtol=1e-5)
scale = np.abs(g0).max(axis=1, keepdims=True)
assert (np.abs(g1.flat.numpy() - g0) / scale).max() < 1e-5
# every rank holds the identical reduced gradient (they apply the identical update)
both = DA.zeros((world,) + g0.shape)
nccl.allgather(g1.flat, both)
device.synchronize()
assert np.array_equal(both.numpy()[0].view(np.uint32), both.numpy()[1].view(np.uint32))
ctl.barrier()
print(f"rank {rank} ok", flush=True)
