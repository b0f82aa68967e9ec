"""The two restatements of the oracle check each other: C (hardware fmaf,
fesetround) vs NumPy (emulated single-rounding FMA / add.rz) -- bit for bit."""
import numpy as np

from oracle import cnative, prng
from oracle import sgmc as osgmc


def _keys(n, base=0):
  return np.stack([prng.PRNGKey(base + i) for i in range(n)])


def test_c_noise_equals_numpy_noise():
  keys = _keys(5, 3)
  for sizes in ([1024], [1, 4], [7, 2, 33], [65536]):
    a = cnative.normal_like(keys, sizes)
    b = osgmc.random_tree_flat(keys, sizes)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), sizes


def test_c_noise_public_vector():
  out = cnative.normal_like(prng.PRNGKey(0)[None], [3])
  # random_tree splits first: compare with the NumPy path instead of raw normal
  want = osgmc.random_tree_flat(prng.PRNGKey(0)[None], [3])
  assert np.array_equal(out, want)


def test_c_sgld_and_psgld_step_equal_numpy():
  rng = np.random.default_rng(0)
  sizes = [1, 4, 19]
  C, P = 6, sum(sizes)
  for rms in (False, True):
    theta = rng.standard_normal((C, P)).astype(np.float32)
    grad = (rng.standard_normal((C, P)) * 3).astype(np.float32)
    v = (np.abs(rng.standard_normal((C, P))) + 0.1).astype(np.float32) if rms else None
    keys = _keys(C, 11)
    st = osgmc.LangevinState(theta.copy(), keys.copy(), None if v is None else v.copy(),
                             np.zeros(C, np.float32), np.ones(C, np.float32))
    want = osgmc.langevin_update(
        st, lambda th: (np.zeros(C, np.float32), np.zeros((C, 2), np.float32), grad),
        sizes, 0.0123, 1.7)
    t2, k2 = theta.copy(), keys.copy()
    v2 = None if v is None else v.copy()
    cnative.sgld_step(t2, v2, grad, k2, sizes, 0.0123, 1.7)
    assert np.array_equal(k2, want.key)
    assert np.array_equal(t2.view(np.uint32), want.theta.view(np.uint32))
    if rms:
      assert np.array_equal(v2.view(np.uint32), want.v.view(np.uint32))
