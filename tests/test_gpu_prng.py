"""GPU parity: in-kernel jax.random (threefry2x32 / split / bits / uniform /
normal / randint / random_tree) is BIT-EXACT against the oracle and the public
jax.random vectors.  All calls go through the C ABI."""
import json
import os

import numpy as np
import pytest

from oracle import data as odata
from oracle import prng
from oracle import sgmc as osgmc

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden",
                                   "prng_public.json")))
LAYOUTS = ["original", "partitionable"]


def _keys(seeds):
  return np.stack([prng.PRNGKey(s) for s in seeds])


def test_public_vectors_on_device(gpu):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray
  for case in GOLD["split"]:
    k = DeviceArray.from_numpy(prng.PRNGKey(case["seed"])[None])
    assert ops.split(k, case["num"]).numpy()[0].tolist() == case["out"]
  for case in GOLD["uniform_scalar"]:
    k = DeviceArray.from_numpy(prng.PRNGKey(case["seed"])[None])
    assert ops.uniform(k, 1).numpy()[0, 0] == np.float32(case["out"])
  for case in GOLD["normal"]:
    key = prng.PRNGKey(case["key_from"]["seed"])
    if "split_index" in case["key_from"]:
      key = prng.split(key)[case["key_from"]["split_index"]]
    n = int(np.prod(case["shape"])) if case["shape"] else 1
    got = ops.normal(DeviceArray.from_numpy(key[None]), n).numpy()[0]
    want = np.array(case["out"], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("n", [1, 2, 7, 8, 1000, 4097])
def test_bits_uniform_normal_bit_exact(gpu, layout, n):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray
  keys = _keys([0, 1, 42, 2 ** 31 + 5])
  dk = DeviceArray.from_numpy(keys)
  assert np.array_equal(ops.random_bits(dk, n, layout).numpy(),
                        prng.random_bits(keys, n, layout))
  u = ops.uniform(dk, n, -10.0, 10.0, layout).numpy()
  assert np.array_equal(u.view(np.uint32),
                        prng.uniform(keys, (n,), -10, 10, layout).view(np.uint32))
  z = ops.normal(dk, n, layout).numpy()
  assert np.array_equal(z.view(np.uint32),
                        prng.normal(keys, (n,), layout).view(np.uint32))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("num", [1, 2, 3, 5, 64])
def test_split_bit_exact(gpu, layout, num):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray
  keys = _keys(range(9))
  got = ops.split(DeviceArray.from_numpy(keys), num, layout).numpy()
  assert np.array_equal(got, prng.split(keys, num, layout))


def test_normal_one_million_bit_exact(gpu):
  """>= 10^6 samples, both erf_inv branches (|u| close to 1 included)."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray
  keys = _keys([2024])
  n = 1 << 20
  got = ops.normal(DeviceArray.from_numpy(keys), n).numpy()
  want = prng.normal(keys, (n,))
  assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
  assert np.abs(want).max() > 4.0      # tail branch (w >= 5) exercised


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("N,n", [(10, 3), (1000, 10), (1_000_000, 1024)])
def test_minibatch_indices_bit_exact(gpu, layout, N, n):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray
  key = prng.PRNGKey(0)
  k_in = DeviceArray.from_numpy(key)
  k_out = DeviceArray((2,), np.uint32)
  idx = DeviceArray((n,), np.int32)
  for _ in range(3):                        # chained draws
    ops.minibatch_draw(k_in, k_out, idx, N, layout)
    key, want = odata.device_draw(key, n, N, layout)
    assert np.array_equal(idx.numpy(), want)
    assert np.array_equal(k_out.numpy(), key)
    k_in, k_out = k_out, k_in
  got = ops.randint(DeviceArray.from_numpy(prng.PRNGKey(9)), n, -5, N, layout)
  assert np.array_equal(got.numpy(), prng.randint(prng.PRNGKey(9), (n,), -5, N,
                                                  layout))


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("sizes", [[1024], [1, 4], [8, 16, 24], [7], [3, 10, 5120],
                                   [1] * 5, [401408, 512, 10]])
def test_random_tree_bit_exact(gpu, layout, sizes):
  """integrator.random_tree: vectorised and ragged leaves, several chains."""
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray
  C = 3 if sum(sizes) > 100000 else 37
  keys = _keys(range(100, 100 + C))
  got = ops.normal_like(DeviceArray.from_numpy(keys), sizes, layout).numpy()
  want = osgmc.random_tree_flat(keys, sizes, layout)
  assert got.shape == want.shape
  assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_gather_rows(gpu):
  from jax_sgmc_b200 import ops
  from jax_sgmc_b200.device import DeviceArray
  rng = np.random.default_rng(0)
  for cols in (1, 4, 5, 1024):
    src = rng.standard_normal((50, cols)).astype(np.float32)
    idx = rng.integers(0, 50, 17).astype(np.int32)
    out = ops.gather_rows(DeviceArray.from_numpy(src), DeviceArray.from_numpy(idx))
    assert np.array_equal(out.numpy(), src[idx])
