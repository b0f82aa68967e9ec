"""Multi-GPU paths with real NCCL / peer memory (skipped on boxes with one device):
the row-sharded gradient all-reduce (tests/_nccl_worker.py) and the sharded reSGLD
ladder against the single-process run, bit for bit, over both transports
(tools/run_sharded_resgld.py --check)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _run_two(script, *extra):
  port, ctl_port = _free_port(), _free_port()
  procs = []
  for r in range(2):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r),
               WORLD_SIZE="2", LOCAL_RANK=str(r), SGMC_CONTROL_PORT=str(ctl_port),
               NCCL_DEBUG="WARN")
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, script), *extra], env=env,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
  outs = [p.communicate(timeout=600)[0] for p in procs]
  for r, (p, o) in enumerate(zip(procs, outs)):
    assert p.returncode == 0, f"rank {r}:\n" + o[-4000:]
  return outs


@pytest.fixture(scope="module")
def two_gpus(gpu):
  if gpu.device_count() < 2:
    pytest.skip("needs two CUDA devices")
  return gpu


def test_row_sharded_gradient_allreduce_two_gpus(two_gpus):
  outs = _run_two("tests/_nccl_worker.py")
  assert "rank 0 ok" in outs[0] and "rank 1 ok" in outs[1]


@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_sharded_resgld_equals_single_process_two_gpus(two_gpus, transport):
  extra = ["--check"] + (["--p2p"] if transport == "p2p" else [])
  outs = _run_two("tools/run_sharded_resgld.py", *extra)
  for r in range(2):
    assert f"rank {r}: sharded == single-process: True" in outs[r]


def test_row_shard_partials_sum_to_the_unsharded_evaluation(gpu):
  """Single device: the R partial evaluations (no communicator) summed on the host equal
  the unsharded potential / variance / gradient within the parity tolerance."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray as DA
  C, d, n, N = 16, 40, 96, 600
  rng = np.random.default_rng(0)
  X = (rng.standard_normal((N, d)) / np.sqrt(d)).astype(np.float32)
  y = (rng.random(N) < 0.5).astype(np.float32)
  theta = (rng.standard_normal((C, d)) * 0.5).astype(np.float32)
  idx = rng.integers(0, N, n).astype(np.int32)
  spec = ops.glm_spec("logistic", d, 0, prior="gaussian", prior_off=0, prior_size=d,
                      prior_scale=2.0)
  dX, dy, dt, di = DA.from_numpy(X), DA.from_numpy(y), DA.from_numpy(theta), DA.from_numpy(idx)
  U0, v0, g0 = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
  ops.glm_potential_grad(spec, dt, dX, dy, di, N, U0, v0, g0, path="simt")
  for R in (2, 4):
    g_sum, ex_sum = np.zeros((C, d), np.float32), np.zeros((C, 3), np.float32)
    for r in range(R):
      Ur, vr, gr = DA((C,), np.float32), DA((C,), np.float32), DA((C, d), np.float32)
      _, scratch = ops.glm_potential_grad_row_sharded(spec, dt, dX, dy, di, N, Ur, vr, gr, n, r, R,
                                                      None, path="simt")
      n_r = n // R
      g_sum += gr.numpy()
      ex_sum += scratch.numpy()[C * n_r:C * n_r + 3 * C].reshape(C, 3)
    U, var = DA((C,), np.float32), DA((C,), np.float32)
    ops.glm_row_shard_finalize(DA.from_numpy(ex_sum), n, U, var)
    np.testing.assert_allclose(U.numpy(), U0.numpy(), rtol=1e-5)
    np.testing.assert_allclose(var.numpy(), v0.numpy(), rtol=2e-4)
    scale = np.abs(g0.numpy()).max(axis=1, keepdims=True)
    assert (np.abs(g_sum - g0.numpy()) / scale).max() < 1e-5
